#!/usr/bin/env python
"""bench.py — PLT paths/s of the PLTPT hot path on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path (default N=1, K=32, W=3, --config cornell)
  python bench.py --impl reference ...                          the reference's CPU path (oracle port) on the host cores
  python bench.py --config {cornell,cornell_closed,grating,layered,city} [--scaling strong] [--partition rows]

Workloads (config.workload), all procedurally generated (BASELINE.json configs[1..4]):
  cornell         configs[1]: Cornell box, PLTDiffuse walls, rough PLT Dielectric + rough Au Conductor boxes, LSPDD area emitter, 1920x1080   (DEFAULT)
  cornell_closed  the same box with a front wall and albedo-0.8 walls: long paths, stresses the O(L^2) forward replay (ours)
  grating         configs[2]: diffraction-grated Au plane + dielectric sphere, area + distant emitter, MNEE through the sphere, 1920x1080
  layered         configs[3]: coated conductors, birefringent thin dielectric, multilayer stacks, 3840x2160
  city            configs[4]: ~10 M triangles, all ten PLT materials, 3840x2160 (BVH build time reported)
All at max depth 8, HWSS 4, NEE + MIS + MNEE + Russian roulette on.  One "step" = FRAMES_PER_STEP frames = that many samples per pixel over
the whole image (one frame = one PLTPT::execute, PLTPT.cpp:357-499), so that the timed region lasts seconds, not milliseconds.

With N GPUs (one process per GPU) the frames are partitioned by sample index -- rank r renders frame indices r, r+N, ... -- or, with
--partition rows, by image tile (every rank renders all frames of its band of rows); the per-GPU fp64 accumulators are summed onto rank 0
with one NCCL reduce.  --scaling weak (default): K steps per GPU; --scaling strong: K steps in total, split over the ranks.

Numbers:
  value  device-resident throughput: all frames through pltpt_render_accumulate, scene and accumulators in HBM, CUDA events, the reduce
         inside the timed region.
  e2e    the same metric end to end per step: camera struct from host memory in, the step's frames accumulated on the device, (N > 1: one NCCL
         reduce of the accumulators,) resolved, ONE RGBA32F image copied to pinned host memory and waited for.  At N = 1 this is the single
         C-ABI call pltpt_render_accumulate_host.  The per-frame plugin path (PLTPT::execute without shared surfaces: three RGBA32F images
         to the host EVERY frame) is reported beside it as e2e.per_frame_*.
  roofline  the kernel class with the largest share of the step, against the measured HBM copy bandwidth (MEASURED_PEAKS.json); kernel
            durations from CUDA events recorded after every launch on the launching stream, frames run one at a time.
  cpu_baseline  the CPU oracle (a restatement of the reference's Slang path; `kind: port`) on the box's host cores, same scene, bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAX_BOUNCES = 8
# name -> (scene factory name, width, height, frames per step, workload description, CPU sample size)
CONFIGS = {
    "cornell": ("cornell_box", 1920, 1080, 16, "cornell_box_plt_dielectric_conductor_lspdd_emitter_1920x1080_depth8 (BASELINE configs[1])", (960, 540)),
    "cornell_closed": ("cornell_box_closed", 1920, 1080, 8, "closed_cornell_box_albedo_0.8_1920x1080_depth8 (deep-path variant of configs[1])", (480, 270)),
    "grating": ("grating_sphere_scene", 1920, 1080, 8, "grated_conductor_plane_dielectric_sphere_mnee_1920x1080_depth8 (BASELINE configs[2])", (480, 270)),
    "layered": ("layered_stress_scene", 3840, 2160, 4, "layered_material_stress_hwss4_3840x2160_depth8 (BASELINE configs[3])", (480, 270)),
    "city": ("city_scene", 3840, 2160, 4, "synthetic_city_10M_triangles_mixed_plt_materials_3840x2160_depth8 (BASELINE configs[4])", (240, 136)),
}
# SURVEY §8(d) algorithmic bytes of the rendering path
F_VERTEX_FETCH = 396   # indices 12 + 3 vertices 96 + instance 32 + matrices 128 + material 128
R_BOUNCE = 40          # bounce record
LIGHT_FETCH = 88       # alias 8 + flux 16 + emissive triangle 64
NODE_BYTES, TRI_BYTES = 80, 48
TRAVERSAL_CLASSES = ("trace", "shadow_trace", "mnee")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def build_scene(config):
    from pltfalcor_b200 import scene
    factory, w, h = CONFIGS[config][:3]
    return getattr(scene, factory)(aspect=w / h)


def path_bytes(stats, frames, paths):
    """per-path algorithmic bytes of the two stages from the measured path statistics."""
    paths = max(paths, 1)
    L = stats["totalPathVertices"] / paths                      # mean recorded vertices per path
    replay = stats["totalReplaySteps"] / paths                  # mean sum_{p<L} p
    bvh = (stats["traversalNodesVisited"] * NODE_BYTES + stats["traversalTrianglesTested"] * TRI_BYTES) / paths
    sample_b = L * (F_VERTEX_FETCH + R_BOUNCE)
    solve_b = L * (R_BOUNCE + F_VERTEX_FETCH + LIGHT_FETCH) + replay * (R_BOUNCE + F_VERTEX_FETCH) + 48
    return dict(mean_vertices=L, mean_replay_steps=replay, sample_bytes=sample_b, solve_bytes=solve_b, bvh_bytes=bvh, total=sample_b + solve_b + bvh)


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture
    (profiles/traffic.json, written by tools/ncu_traffic.py from the raw CSV page); None when no capture is committed."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f)
    e = t.get(kernel)
    return e["dram_bytes_per_launch"] if e else None


def cpu_baseline_sample(config, buffers, cam, opts, seconds=10.0):
    """Times the oracle (CPU restatement of the reference path) on all host cores on a bounded sample of the workload."""
    from oracle import binding as ob
    from pltfalcor_b200 import scene
    osc = ob.OracleScene(buffers)
    threads = ob.hardware_threads()
    w, h = CONFIGS[config][5]
    osc.render_frame(opts, scene.with_jitter(cam, 0, w, h), 0, w, h, threads=threads)  # warm-up (page in, thread pool)
    t0, frames = time.perf_counter(), 0
    while True:
        osc.render_frame(opts, scene.with_jitter(cam, 1 + frames, w, h), 1 + frames, w, h, threads=threads)
        frames += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or frames >= 64:
            break
    return {"value": w * h * frames / dt, "unit": "paths/s", "cores": threads, "kind": "port",
            "sample": f"{frames} frames of the same scene, camera and options at {w}x{h}, depth {MAX_BOUNCES}, {dt:.1f} s wall, oracle/liboracle.so"}


def cpu_bsdf_sweep_sample(ids, sw_buffers, seconds=4.0):
    """BASELINE configs[0] on the CPU oracle: BSDF evals/s (one eval = one wavelength's Mueller matrix) on a bounded sample."""
    from oracle import binding as ob
    osc = ob.OracleScene(sw_buffers)
    threads = ob.hardware_threads()
    rng = np.random.default_rng(1)
    n, groups = 1 << 15, 4
    wi = rng.normal(size=(n, 3)).astype(np.float32); wi /= np.linalg.norm(wi, axis=1, keepdims=True); wi[:, 2] = np.abs(wi[:, 2])
    wo = rng.normal(size=(n, 3)).astype(np.float32); wo /= np.linalg.norm(wo, axis=1, keepdims=True)
    hero = (380 + 340 * rng.random((groups, n))).astype(np.float32)
    out = {}
    for key in ("dielectric", "conductor"):
        osc.bsdf_eval_mueller(ids[key], 0x33, wi[:1024], wo[:1024], hero[:, :1024], threads=threads)
        t0, reps = time.perf_counter(), 0
        while time.perf_counter() - t0 < seconds / 2:
            osc.bsdf_eval_mueller(ids[key], 0x33, wi, wo, hero, threads=threads)
            reps += 1
        dt = time.perf_counter() - t0
        out[key] = {"evals_per_s": reps * n * groups * 4 / dt, "cores": threads, "sample": f"{reps} x 2^15 queries x 16 wavelengths, {dt:.1f} s wall"}
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path. slangc is not available, so the Slang path cannot be compiled
    for the CPU as a whole; the oracle port (C++ restatement, pinned to the reference's shader text by tests/test_oracle_ref.py, all host
    threads) stands in (DESIGN.md 'Oracle').  One step = one frame of the configuration's scene at the bounded CPU sample size."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as ob
    from pltfalcor_b200 import layouts as L, scene
    buffers, cam = build_scene(args.config)
    opts = L.make_options(maxBounces=MAX_BOUNCES)
    osc = ob.OracleScene(buffers)
    threads = ob.hardware_threads()
    W_, H_ = CONFIGS[args.config][1:3]
    w, h = (W_ // 4, H_ // 4) if args.config == "cornell" else CONFIGS[args.config][5]
    for f in range(args.warmup):
        osc.render_frame(opts, scene.with_jitter(cam, f, w, h), f, w, h, threads=threads)
    t0 = time.perf_counter()
    for f in range(args.steps):
        osc.render_frame(opts, scene.with_jitter(cam, 100 + f, w, h), 100 + f, w, h, threads=threads)
    dt = time.perf_counter() - t0
    value = w * h * args.steps / dt
    sample = f"each step = one frame at {w}x{h} (a bounded sample of the {W_}x{H_} pixels, same scene/camera/options)"
    emit({
        "impl": "reference", "metric": "plt_paths_per_s", "value": value, "unit": "paths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CONFIGS[args.config][4], "max_bounces": MAX_BOUNCES, "hwss": 4, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "paths/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def bsdf_sweep(torch, p, ids, hbm_peak):
    """BASELINE configs[0] on the GPU: rough dielectric + rough Au conductor, 2^20 random wi/wo x 16 wavelengths, Mueller output."""
    n, groups = 1 << 20, 4
    g = torch.Generator(device="cuda").manual_seed(1)
    wi = torch.nn.functional.normalize(torch.randn((n, 3), device="cuda", generator=g), dim=1)
    wi[:, 2].abs_()
    wo = torch.nn.functional.normalize(torch.randn((n, 3), device="cuda", generator=g), dim=1)
    hero = 380 + 340 * torch.rand((groups, n), device="cuda", generator=g)
    out = torch.empty((n, groups, 4, 4, 4), device="cuda")  # 1.07 GB > L2
    res = {}
    for key in ("dielectric", "conductor"):
        for _ in range(2):
            p.bsdf_eval_mueller(ids[key], 0x33, wi, wo, hero, out)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        reps = 5
        for _ in range(reps):
            p.bsdf_eval_mueller(ids[key], 0x33, wi, wo, hero, out)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        evals = n * groups * 4
        gbs = evals * 66.5 / (ms * 1e-3) / 1e9  # 66.5 algorithmic bytes per BSDF eval
        res[key] = {"evals_per_s": evals / (ms * 1e-3), "ms": ms, "achieved_gbs": gbs, "hbm_frac": gbs / hbm_peak}
    return res


_RESULT_FD = None


def claim_stdout():
    """Rank 0 must print exactly one JSON line on stdout, but libraries write there too (NCCL prints its version banner to fd 1 on the
    first communicator, whatever NCCL_DEBUG_FILE says).  Keep the real stdout aside for the result and point fd 1 at stderr."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(result):
    line = (json.dumps(result) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cornell", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: K steps per GPU; strong: K steps in total, split over the GPUs")
    ap.add_argument("--partition", default="sample", choices=["sample", "rows"], help="multi-GPU partition: by sample (frame) index, or by image row bands")
    ap.add_argument("--frames-per-step", type=int, default=0, help="frames (samples per pixel) in one step; 0 = the configuration's default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timers", action="store_true", help="skip the per-launch CUDA-event pass (no roofline block)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the BSDF sweep (BASELINE configs[0]) that rides along at N = 1 on the default configuration")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from pltfalcor_b200 import binding, build, layouts as L, partition, scene
    if build.needs_build():
        # normally built beforehand by __graft_entry__.build(); a stale or missing library is rebuilt here (nvcc, about a minute)
        # rather than failing the run -- there is no other implementation to fall back to
        try:
            import fcntl
            with open(build.LIB + ".lock", "w") as lock:  # under torchrun every rank gets here: one builds, the others wait and find it fresh
                fcntl.flock(lock, fcntl.LOCK_EX)
                build.build()
        except Exception as e:  # noqa: BLE001
            raise SystemExit(f"libpltpt.so is stale or missing and could not be rebuilt ({e}): run `python -c 'import __graft_entry__ as g; g.build()'`")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the PLTPT path has no CPU fallback")
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    W, K = max(args.warmup, 3), args.steps
    WIDTH, HEIGHT = CONFIGS[args.config][1:3]
    F = args.frames_per_step or CONFIGS[args.config][3]
    if args.scaling == "strong":
        F = -(-F // world) * world   # a step's frames split evenly over the ranks (a 2-frame step over 8 GPUs would leave six of them idle)
    hbm_peak, peak_src = peaks()

    buffers, cam = build_scene(args.config)
    p = binding.Pltpt(device=local, maxBounces=MAX_BOUNCES)
    p.set_scene(buffers)
    build_stats = p.stats()
    accum = torch.zeros((HEIGHT, WIDTH, 4), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    rows = partition.row_band(rank, world, HEIGHT) if args.partition == "rows" else (0, HEIGHT)

    def render(first_frame, frame_step, frame_count, target):
        """this rank's share of a frame set, into `target` (RGBA64F sums), asynchronously on the bench stream"""
        if frame_count <= 0 or rows[0] >= rows[1]:
            return
        if args.partition == "rows":
            p.render_accumulate_rows(cam, first_frame, frame_step, frame_count, WIDTH, HEIGHT, rows[0], rows[1], target, stream=sptr)
        else:
            p.render_accumulate(cam, first_frame, frame_step, frame_count, WIDTH, HEIGHT, target, stream=sptr)

    def share(first, total_frames_all_ranks):
        """(first_frame, frame_step, frame_count) of this rank for a frame set of `total_frames_all_ranks` frames starting at `first`"""
        if args.partition == "rows":
            return first, 1, total_frames_all_ranks
        return partition.frames_of_rank_strong(rank, world, first, total_frames_all_ranks)

    # frames of the whole job: weak = K steps of F frames per GPU, strong = K steps of F frames in total
    job_frames = K * F * (world if args.scaling == "weak" else 1)
    step_frames = F * (world if args.scaling == "weak" else 1)   # frames (over all ranks) that make up one step's image

    # ---- device-resident timing -------------------------------------------------------------------------------------
    # W warm-up steps of F frames per rank (they also size the work buffers of every frame lane: a lane first used inside the timed region
    # would put its cudaMalloc there)
    render(*share(0, max(W * F, 4) * (world if args.partition == "sample" else 1)), accum)
    accum.zero_()
    if world > 1:
        dist.reduce(accum, dst=0)  # warm-up of the NCCL communicator
        accum.zero_()
    torch.cuda.synchronize()
    p.reset_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    first, step, count = share(1000, job_frames)
    render(first, step, count, accum)
    partition.reduce_accumulators(dist, accum, dst=0)
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    timed_stats = p.stats()
    paths_total = WIDTH * HEIGHT * job_frames
    value = paths_total / (total_ms * 1e-3)
    dropped = torch.tensor([float(timed_stats["droppedConnections"] + timed_stats["traversalStackOverflows"])], device="cuda")
    if world > 1:
        dist.all_reduce(dropped, op=dist.ReduceOp.SUM)
    if float(dropped.item()) != 0:
        raise SystemExit("connections were dropped or a traversal stack overflowed during the timed region: the result would be biased")

    # ---- per-kernel durations: this rank's share of ONE step again with a CUDA event after every launch on the launching stream.  The
    # timed region above keeps up to four frames in flight on their own streams (kernels of different frames share the SMs, so a per-launch duration
    # taken there would include the other frame's work); this pass runs the frames one after the other so that every kernel is timed
    # alone, which is what a roofline fraction needs.  Its frame time is reported as `sequential_ms_per_frame`.
    stats, seq_ms, prof_frames = timed_stats, None, 0
    if rank == 0 and not args.no_kernel_timers:
        scratch = torch.zeros_like(accum)
        p.reset_stats()
        p.set_profiling(True)
        f0_, st_, cnt_ = share(1000, max(step_frames, 2 * world))
        prof_frames = cnt_
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        render(f0_, st_, cnt_, scratch)
        s1.record(stream)
        torch.cuda.synchronize()
        seq_ms = s0.elapsed_time(s1) / max(cnt_, 1)
        stats = p.stats()
        p.set_profiling(False)
        del scratch
    if world > 1:
        dist.barrier()

    # ---- end to end: per step, camera from host memory in -> this rank's share of the step's frames accumulated on the device -> (one NCCL
    # reduce) -> resolve -> ONE RGBA32F image in pinned host memory, waited for.  N = 1: the single C-ABI call pltpt_render_accumulate_host.
    host_img = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32, pin_memory=True)
    step_accum = torch.zeros_like(accum)
    step_color = torch.zeros((HEIGHT, WIDTH, 4), dtype=torch.float32, device="cuda")
    e2e_steps = max(2, min(K, 16))

    def e2e_step(i):
        base = 5000 + i * step_frames
        if world == 1 and args.partition == "sample":
            p.render_accumulate_host(cam, base, 1, step_frames, WIDTH, HEIGHT, host_img)
            return
        step_accum.zero_()
        render(*share(base, step_frames), step_accum)
        partition.reduce_accumulators(dist, step_accum, dst=0)
        if rank == 0:
            p.resolve_accumulated(step_accum, WIDTH, HEIGHT, step_color, stream=sptr)
            host_img.copy_(step_color, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step(-1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = WIDTH * HEIGHT * step_frames * e2e_steps / float(t.item())
    checksum = float(host_img[..., :3].double().mean()) if rank == 0 else None

    # ---- the multi-GPU image must be the single-GPU image: rank 0 renders the last step's whole frame set alone and compares (fp64 sums)
    multi_equals_single = None
    if world > 1:
        if rank == 0:
            alone = torch.zeros_like(accum)
            p.render_accumulate(cam, 5000 + (e2e_steps - 1) * step_frames, 1, step_frames, WIDTH, HEIGHT, alone, stream=sptr)
            torch.cuda.synchronize()
            multi_equals_single = bool(torch.allclose(step_accum, alone, rtol=1e-12, atol=0) and bool((alone[..., 3] == step_frames).all()))
            del alone
        dist.barrier()

    # ---- the per-frame plugin path (PLTPT::execute without shared surfaces): camera in, three RGBA32F images out EVERY frame (N = 1 only)
    per_frame = None
    if world == 1 and args.partition == "sample":
        host = [torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32, pin_memory=True) for _ in range(9)]  # three image sets
        pf_frames = 3 * ((max(4, min(K * F, 48)) + 2) // 3)
        wait = binding.lib().pltpt_host_wait

        def pf_step(f, i, queued):
            c = scene.with_jitter(cam, f, WIDTH, HEIGHT)
            b = 3 * (i % 3)
            if queued:
                binding._check(binding.lib().pltpt_render_frame_host_async(p.h, f, C.byref(c), WIDTH, HEIGHT, host[b].data_ptr(), host[b + 1].data_ptr(), host[b + 2].data_ptr()))
                binding._check(wait(p.h, 2))
            else:
                binding._check(binding.lib().pltpt_render_frame_host(p.h, f, C.byref(c), WIDTH, HEIGHT, host[b].data_ptr(), host[b + 1].data_ptr(), host[b + 2].data_ptr()))

        def pf_run(queued):
            for f in range(8):  # warm-up: both output sets of every frame lane get their buffers before the clock starts
                pf_step(f, f, queued)
            binding._check(wait(p.h, 0))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for f in range(pf_frames):
                pf_step(9000 + f, f, queued)
            binding._check(wait(p.h, 0))
            torch.cuda.synchronize()
            return WIDTH * HEIGHT * pf_frames / (time.perf_counter() - t0)

        sync_value = pf_run(False)
        sync_image = host[6].clone()
        queued_value = pf_run(True)
        per_frame = {"queued_value": queued_value, "synchronous_value": sync_value, "unit": "paths/s", "frames": pf_frames, "d2h_bytes_per_frame": 3 * WIDTH * HEIGHT * 16,
                     "queued_equals_synchronous": bool(torch.equal(sync_image, host[6])),
                     "api": "pltpt_render_frame_host_async + pltpt_host_wait(h, 2) (two frames in flight) / pltpt_render_frame_host"}
        del host

    if rank == 0:
        frames_timed = max(timed_stats["framesRendered"], 1)
        paths_prof = WIDTH * (rows[1] - rows[0]) * max(prof_frames, 1)
        pb = path_bytes(stats, prof_frames, paths_prof) if prof_frames else None
        out = {
            "metric": "plt_paths_per_s", "value": value, "unit": "paths/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIGS[args.config][4], "name": args.config, "width": WIDTH, "height": HEIGHT, "max_bounces": MAX_BOUNCES, "hwss": 4, "nee": True, "mis": True, "mnee": True,
                       "russian_roulette": True, "frames_in_flight": min(4, F), "frames_per_step": F, "paths_per_step": WIDTH * HEIGHT * F * (world if args.scaling == "weak" else 1),
                       "frames_total": job_frames,
                       "partition": ("frame index modulo n_gpus" if args.partition == "sample" else "bands of image rows, every rank renders all frames of its band") +
                                    "; fp64 accumulators summed by one NCCL reduce inside the timed region",
                       "l2": "inputs larger than L2: every frame rewrites and re-reads its %.0f MB bounce buffer (126 MB L2)" % ((MAX_BOUNCES + 1) * WIDTH * HEIGHT * 40 / 1e6),
                       "triangles": int(build_stats["triangleCount"]), "bvh_nodes": int(build_stats["bvhNodes"]), "bvh_build_ms": build_stats["bvhBuildMs"],
                       "bvh_sah_cost": build_stats.get("bvhSahCost")},
            "e2e": {"value": e2e_value, "unit": "paths/s", "h2d_bytes_per_step": C.sizeof(L.PltptCameraData), "d2h_bytes_per_step": WIDTH * HEIGHT * 16, "steps": e2e_steps,
                    "api": ("pltpt_render_accumulate_host: camera in, the step's frames accumulated on the device in fp64, one resolved RGBA32F image out to pinned host memory" if world == 1 and args.partition == "sample"
                            else "per rank pltpt_render_accumulate[_rows] of its share -> one NCCL reduce of the fp64 accumulators -> pltpt_resolve_accumulated -> one RGBA32F image to pinned host memory on rank 0"),
                    "checksum_mean_rgb": checksum, "multi_gpu_image_equals_single_gpu_1e-12": multi_equals_single, "per_frame_plugin_path": per_frame},
            "gpu_launches": int(timed_stats["kernelLaunches"]),
            "clocks": clocks,
        }
        if prof_frames:
            frames = max(stats["framesRendered"], 1)
            solve_ms, sample_ms = stats["solveStageMs"] / frames, stats["sampleStageMs"] / frames
            kms, kcount = stats["kernelMs"], stats["kernelLaunchCount"]
            kernels = {k: {"ms_per_frame": kms[k] / frames, "launches_per_frame": kcount[k] / frames} for k in kms if kcount[k]}
            timed_sum = sum(v["ms_per_frame"] for v in kernels.values())
            for v in kernels.values():
                v["share_of_frame"] = v["ms_per_frame"] / timed_sum if timed_sum else None
            # the dominant kernel class of this configuration.  Replay step: unit = one generalized ray pushed through one recorded vertex
            # (R + F = 436 B).  Traversal kernels (closest-hit trace, shadow rays, manifold-walk trial rays run the same node /
            # triangle code): unit = one BVH node visited or leaf triangle tested, 80 B / 48 B, against their summed duration.
            trav_ms = sum(kms[k] for k in TRAVERSAL_CLASSES)
            bvh_bytes = stats["traversalNodesVisited"] * NODE_BYTES + stats["traversalTrianglesTested"] * TRI_BYTES
            trav_launches = sum(kcount[k] for k in TRAVERSAL_CLASSES)
            # candidates: (kernel name, summed ms, launches, units, algorithmic bytes per unit or None for the traversal group, traffic key)
            cand = {
                "unwind_step": ("k_unwind_step", kms.get("unwind_step", 0.0), kcount.get("unwind_step", 0), stats["unwindJobSteps"], R_BOUNCE + F_VERTEX_FETCH, "k_unwind_step"),
                "shade": ("k_shade", kms.get("shade", 0.0), kcount.get("shade", 0), stats["totalPathVertices"], F_VERTEX_FETCH + R_BOUNCE, "k_shade"),
                "connect_a": ("k_connect_a", kms.get("connect_a", 0.0), kcount.get("connect_a", 0), stats["totalPathVertices"], R_BOUNCE + F_VERTEX_FETCH + LIGHT_FETCH, "k_connect_a"),
                "traversal": ("k_trace + k_shadow_trace + k_mnee_walk (BVH traversal)", trav_ms, trav_launches, stats["traversalNodesVisited"] + stats["traversalTrianglesTested"], None, args.config + ":traversal"),
            }
            per_kernel = {}
            for key, (name, ms_total, launches, units, bpu, tkey) in cand.items():
                if not launches or ms_total <= 0:
                    continue
                alg = bvh_bytes if bpu is None else units * bpu
                per_kernel[key] = {"kernel": name, "ms_per_frame": ms_total / frames, "share_of_frame": ms_total / frames / timed_sum if timed_sum else None, "launches": int(launches),
                                   "units_per_launch": units / launches, "algorithmic_bytes_per_unit": bpu if bpu is not None else "80 B per node visited, 48 B per leaf triangle tested",
                                   "achieved": alg / (ms_total * 1e-3) / 1e9, "frac": alg / (ms_total * 1e-3) / 1e9 / hbm_peak, "traffic_key": tkey}
            # the dominant kernel (largest share of the step) is the one the roofline block reports; the others ride along in roofline.kernels.
            # The three traversal kernels share one set of node / triangle counters, so they are accounted as a group -- but the group is only
            # "dominant" when its largest member is the largest single kernel class of the step.
            single = {k: v["ms_per_frame"] for k, v in per_kernel.items() if k != "traversal"}
            if "traversal" in per_kernel:
                single["traversal"] = max(kms[k] for k in TRAVERSAL_CLASSES) / frames
            top = max(single, key=single.get)
            t = per_kernel[top]
            roof = {"kernel": t["kernel"], "kernel_ms": t["ms_per_frame"] * frames / t["launches"], "launches": t["launches"], "units_per_launch": t["units_per_launch"],
                    "algorithmic_bytes_per_unit": t["algorithmic_bytes_per_unit"], "achieved": t["achieved"], "share_of_frame": t["share_of_frame"], "traffic_key": t["traffic_key"]}
            for v in per_kernel.values():
                v["traffic"] = ncu_traffic(v.pop("traffic_key"))
            roof["kernels"] = per_kernel
            roof["timer"] = "CUDA events after every launch on the launching stream; frames run one at a time (sequential_ms_per_frame) so that every kernel is timed alone"
            traffic = ncu_traffic(roof.pop("traffic_key"))
            out["sequential_ms_per_frame"] = seq_ms
            out["roofline"] = {"bound": "hbm", "achieved": roof["achieved"], "peak": hbm_peak, "unit": "GB/s", "frac": roof["achieved"] / hbm_peak, "traffic": traffic,
                               "peak_source": peak_src, **{k: v for k, v in roof.items() if k != "achieved"},
                               "note": "shading and traversal kernels are issue- and latency-bound (fp32 + MUFU, -fmad=false; divergent traversal), not HBM-bound: see profiles/ for the ncu issue-slot and warp-efficiency evidence",
                               "stages": {"sample_ms": sample_ms, "solve_ms": solve_ms, "sample_bytes_per_path": pb["sample_bytes"] + pb["bvh_bytes"], "solve_bytes_per_path": pb["solve_bytes"],
                                          "whole_frame_achieved": pb["total"] * WIDTH * (rows[1] - rows[0]) / ((sample_ms + solve_ms) * 1e-3) / 1e9}}
            out["kernels"] = kernels
            out["path_stats"] = {"mean_vertices": pb["mean_vertices"], "mean_replay_steps": pb["mean_replay_steps"], "bsdf_evals_per_path": stats["totalBsdfEvals"] / paths_prof,
                                 "bvh_nodes_per_path": stats["traversalNodesVisited"] / paths_prof, "bvh_triangles_per_path": stats["traversalTrianglesTested"] / paths_prof,
                                 "algorithmic_bytes_per_path": pb["total"], "mnee_walks_last_frame": int(stats["mneeWalks"])}
        if world == 1:
            if args.config == "cornell" and not args.no_sweep:
                sw_buffers, ids = scene.sweep_scene()
                q = binding.Pltpt(device=local)
                q.set_scene(sw_buffers)
                out["bsdf_sweep"] = bsdf_sweep(torch, q, ids, hbm_peak)
                q.close()
            if not args.no_cpu_baseline:
                out["cpu_baseline"] = cpu_baseline_sample(args.config, buffers, cam, p.options)
                if "bsdf_sweep" in out:
                    cpu_sw = cpu_bsdf_sweep_sample(ids, sw_buffers)
                    for key, v in cpu_sw.items():
                        out["bsdf_sweep"][key]["cpu_oracle"] = v
                        out["bsdf_sweep"][key]["gpu_over_cpu"] = out["bsdf_sweep"][key]["evals_per_s"] / v["evals_per_s"]
        emit(out)
    p.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
