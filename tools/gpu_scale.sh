#!/bin/bash
# Multi-GPU scaling lines (torchrun, one rank per GPU): usage tools/gpu_scale.sh N
N=$1
out=gpurun_out
mkdir -p $out
run() { tag=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-kernel-timers "$@" > $out/r2v_n${N}_$tag.json 2> $out/r2v_n${N}_$tag.err; echo "n$N $tag exit $?"; python - <<PY
import json
try:
    d=json.load(open("$out/r2v_n${N}_$tag.json")); print("   value %.1fM e2e %.1fM equal %s ms/step %.1f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["multi_gpu_image_equals_single_gpu_1e-12"], d["ms_per_step"]))
except Exception as e: print("   parse failed", e)
PY
}
run cornell_weak --steps 20 --warmup 3
run cornell_strong --steps 32 --warmup 3 --scaling strong
run layered_weak --steps 6 --warmup 3 --config layered
run city_weak --steps 4 --warmup 3 --config city
run city_strong --steps 32 --warmup 3 --config city --scaling strong
