#!/bin/bash
# Round 2, call W: BVH refit (hit IDs after a refit against the brute-force oracle, refit vs rebuild time and frame time at 10 M triangles),
# frames in flight 1 .. 4 on the Cornell box, its closed variant and the city.
out=gpurun_out
mkdir -p $out
timeout 400 python -m pytest tests/test_gpu_trace.py -q -s -k "transforms or refit" > $out/r2w_pytest_refit.log 2>&1; echo "pytest exit $?" >> $out/r2w_pytest_refit.log; tail -5 $out/r2w_pytest_refit.log
timeout 200 python -m pytest tests/test_host_shell.py -q -m gpu > $out/r2w_pytest_host.log 2>&1; echo "pytest exit $?" >> $out/r2w_pytest_host.log; tail -3 $out/r2w_pytest_host.log
timeout 200 python tools/overlap_refit.py refit 10000000 > $out/r2w_refit.json 2> $out/r2w_refit.err; cut -c1-1500 $out/r2w_refit.json; tail -3 $out/r2w_refit.err
for cfg in "cornell 1920 1080 32" "cornell_closed 1920 1080 16" "city 1920 1080 8"; do
  set -- $cfg
  timeout 200 python tools/overlap_refit.py overlap $1 $2 $3 $4 >> $out/r2w_overlap.jsonl 2>> $out/r2w_overlap.err
done
cat $out/r2w_overlap.jsonl; tail -3 $out/r2w_overlap.err
