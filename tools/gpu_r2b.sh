#!/bin/bash
# Round 2, call B: node-group traversal. Hit-ID parity (all scenes, both drivers, 10 M triangles), then kernel times of the variants.
out=gpurun_out
mkdir -p $out
V=pltfalcor_b200/build/variants
timeout 900 python -m pytest tests/test_gpu_trace.py -q > $out/r2b_pytest_trace.log 2>&1; echo "pytest exit $?" >> $out/r2b_pytest_trace.log
tail -4 $out/r2b_pytest_trace.log
for v in base trav5 trav6 steps8 refill16; do for s in cornell grating city; do PLTPT_LIB=$V/$v/libpltpt.so timeout 300 python tools/kernel_times.py $s 1920 1080 4 | sed "s|^|$v |"; done; done > $out/r2b_kernel_times.log 2>&1
cat $out/r2b_kernel_times.log | cut -c1-420
for s in grating city; do timeout 300 python tools/mnee_profile.py $s 1920 1080 2 | cut -c1-400; done > $out/r2b_mnee_profile.log 2>&1
cat $out/r2b_mnee_profile.log
timeout 900 python -m pytest tests/test_gpu_integrator.py -q -x -k "bounce_records or single_frame" > $out/r2b_pytest_int.log 2>&1; tail -3 $out/r2b_pytest_int.log
