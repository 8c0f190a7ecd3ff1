#!/bin/bash
# Round 2, call D: full GPU parity suite, smoke, bench (default + reference arm + the other configurations, short).
out=gpurun_out
mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q > $out/r2d_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/r2d_pytest_gpu.log
tail -6 $out/r2d_pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/r2d_smoke.log 2>&1; echo "smoke exit $?" >> $out/r2d_smoke.log; tail -4 $out/r2d_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $out/r2d_bench_n1.json 2> $out/r2d_bench_n1.err; echo "bench exit $?" >> $out/r2d_bench_n1.err; tail -3 $out/r2d_bench_n1.err; cut -c1-900 $out/r2d_bench_n1.json
timeout 300 python bench.py --impl reference --steps 8 --warmup 1 > $out/r2d_bench_ref.json 2> $out/r2d_bench_ref.err; cut -c1-300 $out/r2d_bench_ref.json
for c in cornell_closed grating layered city; do timeout 900 python bench.py --config $c --steps 6 --warmup 3 > $out/r2d_bench_$c.json 2> $out/r2d_bench_$c.err; echo "$c exit $?"; cut -c1-400 $out/r2d_bench_$c.json; tail -2 $out/r2d_bench_$c.err; done
