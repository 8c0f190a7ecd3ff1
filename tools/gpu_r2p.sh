#!/bin/bash
# Round 2, call P: adopted traversal (persistent lanes: nearest child first, leaf triangles of a round tested together, culled entries
# skipped at pop; tiny-scene kernels: ordered push) against the round-1 driver (trav_legacy) and with 6 / 8 node steps per round;
# hit-ID tests through all three query modes + integrator + texture parity on the product build.
out=gpurun_out
mkdir -p $out
V=pltfalcor_b200/build/variants
for v in base trav_legacy defer_s6 defer_s8; do for s in city grating layered; do PLTPT_LIB=$PWD/$V/$v/libpltpt.so timeout 300 python tools/kernel_times.py $s 1920 1080 4 | sed "s|^|$v |"; done; done > $out/r2p_kernel_times.log 2>&1; cut -c1-560 $out/r2p_kernel_times.log
timeout 300 python tools/variants.py one $V/base/libpltpt.so 16 > $out/r2p_variants.log 2>&1; cut -c1-600 $out/r2p_variants.log
timeout 1200 python -m pytest tests/test_gpu_trace.py tests/test_gpu_integrator.py tests/test_gpu_textures.py tests/test_gpu_accumulate.py -q > $out/r2p_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2p_pytest.log; tail -6 $out/r2p_pytest.log
