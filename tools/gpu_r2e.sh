#!/bin/bash
# Round 2, call E: specialised replay kernels and BSDF sweep with the four-column beam (variants), BSDF parity of the new sweep kernel.
out=gpurun_out
mkdir -p $out
V=pltfalcor_b200/build/variants
timeout 900 python -m pytest tests/test_gpu_bsdf.py tests/test_gpu_physics.py -q -x > $out/r2e_pytest_bsdf.log 2>&1; echo "pytest exit $?" >> $out/r2e_pytest_bsdf.log; tail -3 $out/r2e_pytest_bsdf.log
for v in base split split8 sweep2 sweep3; do timeout 300 python tools/variants.py one $V/$v/libpltpt.so 16; done > $out/r2e_variants.log 2>&1
cut -c1-700 $out/r2e_variants.log
for v in base split; do for s in city grating; do PLTPT_LIB=$V/$v/libpltpt.so timeout 300 python tools/kernel_times.py $s 1920 1080 4 | sed "s|^|$v |"; done; done > $out/r2e_kernel_times.log 2>&1
cut -c1-500 $out/r2e_kernel_times.log
