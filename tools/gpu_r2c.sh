#!/bin/bash
# Round 2, call C: SAH-optimal wide collapse. Hit-ID parity, then kernel times: sorted per-child stack vs node-group stack on the new trees.
out=gpurun_out
mkdir -p $out
V=pltfalcor_b200/build/variants
timeout 900 python -m pytest tests/test_gpu_trace.py -q -x > $out/r2c_pytest_trace.log 2>&1; echo "pytest exit $?" >> $out/r2c_pytest_trace.log
tail -4 $out/r2c_pytest_trace.log
for v in base groups groups6; do for s in cornell grating city; do PLTPT_LIB=$V/$v/libpltpt.so timeout 300 python tools/kernel_times.py $s 1920 1080 4 | sed "s|^|$v |"; done; done > $out/r2c_kernel_times.log 2>&1
cat $out/r2c_kernel_times.log | cut -c1-520
timeout 900 python -m pytest tests/test_gpu_integrator.py -q -x -k "bounce_records or single_frame" > $out/r2c_pytest_int.log 2>&1; tail -3 $out/r2c_pytest_int.log
