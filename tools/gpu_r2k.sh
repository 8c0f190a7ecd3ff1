#!/bin/bash
# Round 2, call K: textured materials / normal maps / alpha cut-outs (new GPU tests), full GPU parity suite with the approximate-math replay
# step, per-kernel times (does the textured branch cost the untextured scenes anything?), default bench.
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_textures.py -q > $out/r2k_pytest_textures.log 2>&1; echo "pytest exit $?" >> $out/r2k_pytest_textures.log; tail -15 $out/r2k_pytest_textures.log
timeout 300 python tools/variants.py one pltfalcor_b200/libpltpt.so 16 > $out/r2k_variants.log 2>&1; cut -c1-700 $out/r2k_variants.log
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_trace.py::test_closest_hit_bit_exact_10m_triangles > $out/r2k_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/r2k_pytest_gpu.log; tail -8 $out/r2k_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > $out/r2k_bench_n1.json 2> $out/r2k_bench_n1.err; echo "bench exit $?" >> $out/r2k_bench_n1.err; tail -3 $out/r2k_bench_n1.err; cut -c1-1200 $out/r2k_bench_n1.json
