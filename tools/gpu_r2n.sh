#!/bin/bash
# Round 2, call N: normalize() as one reciprocal + three products (oracle and CUDA path alike): full GPU parity suite; BSDF sweep kernels
# compiled with the replay step's approximate-math flags (variant): BSDF tolerance tests + sweep rate; default bench.
out=gpurun_out
mkdir -p $out
V=pltfalcor_b200/build/variants
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_trace.py::test_closest_hit_bit_exact_10m_triangles > $out/r2n_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/r2n_pytest_gpu.log; tail -8 $out/r2n_pytest_gpu.log
cp $out/parity_metrics.json $out/r2n_parity_metrics.json
PLTPT_LIB=$PWD/$V/sweep_fm/libpltpt.so PLTPT_PARITY_TAG=r2n_sweepfm_ timeout 900 python -m pytest tests/test_gpu_bsdf.py tests/test_gpu_physics.py -q > $out/r2n_pytest_sweepfm.log 2>&1; echo "pytest exit $?" >> $out/r2n_pytest_sweepfm.log; tail -12 $out/r2n_pytest_sweepfm.log
for v in sweep_fm; do timeout 300 python tools/variants.py one $V/$v/libpltpt.so 8; done > $out/r2n_variants.log 2>&1; cut -c1-400 $out/r2n_variants.log
timeout 900 python bench.py --steps 20 --warmup 3 > $out/r2n_bench_n1.json 2> $out/r2n_bench_n1.err; echo "bench exit $?" >> $out/r2n_bench_n1.err; tail -3 $out/r2n_bench_n1.err; cut -c1-300 $out/r2n_bench_n1.json
