#!/bin/bash
# Round 2, call J: which of the three approximate-math flags breaks parity in the connect stage (each alone on stage_connect.cu).
out=gpurun_out
mkdir -p $out
V=pltfalcor_b200/build/variants
for v in c_fmad c_div c_sqrt; do timeout 300 python tools/variants.py one $V/$v/libpltpt.so 8; done > $out/r2j_variants.log 2>&1
cut -c1-700 $out/r2j_variants.log
for v in c_fmad c_div c_sqrt; do
  PLTPT_LIB=$PWD/$V/$v/libpltpt.so PLTPT_PARITY_TAG=r2j_${v}_ timeout 900 python -m pytest tests/test_gpu_integrator.py -q -k "single_frame or option_variants" > $out/r2j_pytest_$v.log 2>&1
  tail -3 $out/r2j_pytest_$v.log
done
