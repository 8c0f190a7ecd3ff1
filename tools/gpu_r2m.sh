#!/bin/bash
# Round 2, call M: every solve-stage kernel instantiated without the textured path (manifold walks back to their former time?), persistent-lane
# traversal parameters (refill threshold, steps per round), normalisation with one reciprocal (experiment).
out=gpurun_out
mkdir -p $out
V=pltfalcor_b200/build/variants
for v in base nrm_rcp; do timeout 300 python tools/variants.py one $V/$v/libpltpt.so 16; done > $out/r2m_variants.log 2>&1; cut -c1-700 $out/r2m_variants.log
for v in base trav_r4 trav_r16 trav_s2 trav_s8; do for s in grating city; do PLTPT_LIB=$PWD/$V/$v/libpltpt.so timeout 300 python tools/kernel_times.py $s 1920 1080 4 | sed "s|^|$v |"; done; done > $out/r2m_kernel_times.log 2>&1; cut -c1-520 $out/r2m_kernel_times.log
PLTPT_LIB=$PWD/$V/base/libpltpt.so timeout 300 python tools/kernel_times.py layered 1920 1080 4 >> $out/r2m_kernel_times.log 2>&1
timeout 900 python -m pytest tests/test_gpu_textures.py tests/test_gpu_integrator.py -q -x > $out/r2m_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2m_pytest.log; tail -4 $out/r2m_pytest.log
