#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench line, ncu launch list, full ncu captures of the top kernels.
# usage (from the repo root on the GPU box): bash tools/gpu_round.sh <tag> [skip-ncu]
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/${tag}_smoke.log 2>&1; echo "smoke exit $?" >> $out/${tag}_smoke.log
timeout 600 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench exit $?" >> $out/${tag}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 8 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
if [ -z "$2" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_launches_bench.log 2>&1
  bash tools/ncu_capture.sh $tag
fi
tail -3 $out/${tag}_pytest_gpu.log; tail -4 $out/${tag}_smoke.log; cat $out/${tag}_bench_n1.json | cut -c1-600
