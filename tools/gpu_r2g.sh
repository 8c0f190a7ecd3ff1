#!/bin/bash
# Round 2, call G (2 GPUs): single-process multi-GPU entry point against single GPU, torchrun bench at N = 2 (weak / strong / row bands).
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/r2g_smi.txt
timeout 600 python -m pytest tests/test_gpu_integrator.py -q -k "multi_gpu or row_bands" > $out/r2g_pytest_multi.log 2>&1; echo "pytest exit $?" >> $out/r2g_pytest_multi.log; tail -3 $out/r2g_pytest_multi.log
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-kernel-timers "$@" > $out/r2g_bench_$tag.json 2> $out/r2g_bench_$tag.err; echo "$tag exit $?"; cut -c1-260 $out/r2g_bench_$tag.json; python - <<PY
import json
try:
    d=json.load(open("$out/r2g_bench_$tag.json")); print("   value %.1fM e2e %.1fM equal %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["multi_gpu_image_equals_single_gpu_1e-12"]))
except Exception as e: print("   parse failed", e)
PY
}
run n2_weak --steps 10 --warmup 3
run n2_strong --steps 10 --warmup 3 --scaling strong
run n2_rows --steps 10 --warmup 3 --partition rows
run n2_city --steps 4 --warmup 3 --config city
