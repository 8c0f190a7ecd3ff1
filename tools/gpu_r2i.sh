#!/bin/bash
# Round 2, call I: replay step with approximate division / rsqrt / folded Gaussians / exp2-log2 power (fm_u), occupancy and unrolling re-tested on top of it.
out=gpurun_out
mkdir -p $out
V=pltfalcor_b200/build/variants
for v in fm_u fm_mb5 fm_mb3 fm_unroll; do timeout 300 python tools/variants.py one $V/$v/libpltpt.so 16; done > $out/r2i_variants.log 2>&1
cut -c1-700 $out/r2i_variants.log
for v in fm_u; do
  PLTPT_LIB=$PWD/$V/$v/libpltpt.so PLTPT_PARITY_TAG=r2i_${v}_ timeout 900 python -m pytest tests/test_gpu_integrator.py -q -k "single_frame or option_variants" > $out/r2i_pytest_$v.log 2>&1
  tail -3 $out/r2i_pytest_$v.log
done
