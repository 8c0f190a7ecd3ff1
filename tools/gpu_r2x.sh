#!/bin/bash
# Round 2, call X: final build of the round (call R plus four frames in flight, BVH refit, Mitsuba textures, accumulate
# modes): full GPU parity suite, smoke, bench (default + reference arm + the other configurations).
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/r2x_smi.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q > $out/r2x_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/r2x_pytest_gpu.log; tail -12 $out/r2x_pytest_gpu.log
cp $out/parity_metrics.json $out/r2x_parity_metrics.json; cp $out/image_parity.json $out/r2x_image_parity.json
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/r2x_smoke.log 2>&1; echo "smoke exit $?" >> $out/r2x_smoke.log; tail -4 $out/r2x_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $out/r2x_bench_n1.json 2> $out/r2x_bench_n1.err; echo "bench exit $?" >> $out/r2x_bench_n1.err; tail -3 $out/r2x_bench_n1.err; cut -c1-900 $out/r2x_bench_n1.json
timeout 300 python bench.py --impl reference --steps 8 --warmup 1 > $out/r2x_bench_ref.json 2> $out/r2x_bench_ref.err; cut -c1-300 $out/r2x_bench_ref.json
for c in cornell_closed grating layered city; do timeout 900 python bench.py --config $c --steps 6 --warmup 3 > $out/r2x_bench_$c.json 2> $out/r2x_bench_$c.err; echo "$c exit $?"; cut -c1-330 $out/r2x_bench_$c.json; tail -2 $out/r2x_bench_$c.err; done
