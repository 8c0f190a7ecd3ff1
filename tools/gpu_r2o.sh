#!/bin/bash
# Round 2, call O: traversal variants -- intersected children pushed from the slab-test loop with only the nearest moved to the top
# (push_near), leaf triangles of a round tested by all lanes together (defer), both (defer_push): per-kernel times on the large scenes,
# bit-exact hit-ID tests and integrator parity for each.
out=gpurun_out
mkdir -p $out
V=pltfalcor_b200/build/variants
for v in base push_near defer defer_push; do for s in city grating layered; do PLTPT_LIB=$PWD/$V/$v/libpltpt.so timeout 300 python tools/kernel_times.py $s 1920 1080 4 | sed "s|^|$v |"; done; done > $out/r2o_kernel_times.log 2>&1; cut -c1-560 $out/r2o_kernel_times.log
for v in push_near defer defer_push; do
  PLTPT_LIB=$PWD/$V/$v/libpltpt.so PLTPT_PARITY_TAG=r2o_${v}_ timeout 900 python -m pytest tests/test_gpu_trace.py tests/test_gpu_integrator.py -q -k "not 10m" > $out/r2o_pytest_$v.log 2>&1; echo "pytest exit $?" >> $out/r2o_pytest_$v.log; tail -4 $out/r2o_pytest_$v.log
done
for v in base defer_push; do timeout 300 python tools/variants.py one $V/$v/libpltpt.so 16; done > $out/r2o_variants.log 2>&1; cut -c1-600 $out/r2o_variants.log
