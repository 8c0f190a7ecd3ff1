#!/bin/bash
# Round 2, call Z (two GPUs): the single-process multi-GPU entry and the two-rank bench on the final build (four frames in flight per GPU).
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_integrator.py -q -k "multi" > $out/r2z_pytest_multi.log 2>&1; echo "pytest exit $?" >> $out/r2z_pytest_multi.log; tail -3 $out/r2z_pytest_multi.log
N=2
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-kernel-timers "$@" > $out/r2z_n${N}_$tag.json 2> $out/r2z_n${N}_$tag.err; echo "n$N $tag exit $?"; python - <<PY
import json
try:
    d=json.load(open("$out/r2z_n${N}_$tag.json")); print("   value %.1fM e2e %.1fM equal %s ms/step %.1f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["multi_gpu_image_equals_single_gpu_1e-12"], d["ms_per_step"]))
except Exception as e: print("   parse failed", e)
PY
}
run cornell_weak --steps 20 --warmup 3
run cornell_strong --steps 32 --warmup 3 --scaling strong
run city_weak --steps 4 --warmup 3 --config city
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sweep --no-kernel-timers > $out/r2z_bench_n1.json 2> $out/r2z_bench_n1.err; python -c "
import json; d=json.load(open('$out/r2z_bench_n1.json')); print('n1 value %.1fM e2e %.1fM' % (d['value']/1e6, d['e2e']['value']/1e6), d['e2e']['per_frame_plugin_path'])"
