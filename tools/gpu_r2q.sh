#!/bin/bash
# Round 2, call Q: PLOC binary tree (radius 8; 4 and 16 as variants) under the SAH-optimal wide collapse against the Karras radix tree
# (bvh_lbvh): bit-exact hit IDs through every traversal driver (incl. the 10 M-triangle scene), SAH cost, build time, nodes per path,
# per-kernel times on the large scenes.
out=gpurun_out
mkdir -p $out
V=pltfalcor_b200/build/variants
timeout 900 python -m pytest tests/test_gpu_trace.py -q -x > $out/r2q_pytest_trace.log 2>&1; echo "pytest exit $?" >> $out/r2q_pytest_trace.log; tail -5 $out/r2q_pytest_trace.log
for v in base bvh_lbvh ploc_r4 ploc_r16; do for s in city grating layered; do PLTPT_LIB=$PWD/$V/$v/libpltpt.so timeout 300 python tools/kernel_times.py $s 1920 1080 4 | sed "s|^|$v |"; done; done > $out/r2q_kernel_times.log 2>&1; cut -c1-640 $out/r2q_kernel_times.log
timeout 900 python -m pytest tests/test_gpu_integrator.py -q -x > $out/r2q_pytest_integrator.log 2>&1; echo "pytest exit $?" >> $out/r2q_pytest_integrator.log; tail -5 $out/r2q_pytest_integrator.log
