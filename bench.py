#!/usr/bin/env python
"""bench.py — PLT paths/s of the PLTPT hot path on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path (default N=1, K=32, W=3)
  python bench.py --impl reference ...                          the reference's CPU path (oracle port) on the host cores

Workload (config.workload): BASELINE configs[1] — procedural Cornell box (PLTDiffuse walls, rough PLT Dielectric box,
rough Au PLT Conductor box, LSPDD area emitter), 1920x1080, max depth 8, HWSS 4, NEE + MIS + MNEE on.  One "step" = one
frame = one sample per pixel over the whole image (= one PLTPT::execute, PLTPT.cpp:357-499) = 2 073 600 paths.
With N GPUs every rank renders its own frame indices {rank, rank+N, ...} (sample-index partitioning; scene replicated) and
the per-GPU fp64 accumulation buffers are summed onto rank 0 with one NCCL reduce inside the timed region ("weak" scaling:
K frames per GPU).

Numbers:
  value  device-resident throughput: K frames through pltpt_render_accumulate, scene and accumulators in HBM, CUDA events.
  e2e    the same metric through the host-buffer C-ABI calls a Falcor shell makes without shared surfaces: camera struct H2D and
         the three RGBA32F RenderData images D2H (pinned) every step.  Queued (pltpt_render_frame_host_async + pltpt_host_wait,
         two frames in flight) is the value; the synchronous call (pltpt_render_frame_host) is reported beside it.
  roofline  dominant kernel (k_unwind_step, the forward replay) against the measured HBM copy bandwidth (MEASURED_PEAKS.json);
            its duration comes from CUDA events recorded after every launch on the launching stream inside the timed region.
  cpu_baseline  the CPU oracle (a restatement of the reference's Slang path; `kind: port`) on the box's host cores.
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

WIDTH, HEIGHT, MAX_BOUNCES = 1920, 1080, 8
WORKLOAD = "cornell_box_plt_dielectric_conductor_lspdd_emitter_1920x1080_depth8 (BASELINE configs[1]; 1 spp per step)"
# SURVEY §8(d) algorithmic bytes of the rendering path
F_VERTEX_FETCH = 396   # indices 12 + 3 vertices 96 + instance 32 + matrices 128 + material 128
R_BOUNCE = 40          # bounce record
LIGHT_FETCH = 88       # alias 8 + flux 16 + emissive triangle 64
NODE_BYTES, TRI_BYTES = 80, 48


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


def build_scene():
    from pltfalcor_b200 import scene
    return scene.cornell_box(aspect=WIDTH / HEIGHT)


def path_bytes(stats, frames, paths):
    """per-path algorithmic bytes of the two stages from the measured path statistics."""
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


def cpu_baseline_sample(buffers, cam, opts, seconds=10.0):
    """Times the oracle (CPU restatement of the reference path) on all host cores on a bounded sample of the workload."""
    from oracle import binding as ob
    from pltfalcor_b200 import scene
    osc = ob.OracleScene(buffers)
    threads = ob.hardware_threads()
    w, h = WIDTH // 2, HEIGHT // 2
    osc.render_frame(opts, scene.with_jitter(cam, 0, w, h), 0, w, h, threads=threads)  # warm-up (page in, thread pool)
    t0, frames = time.perf_counter(), 0
    while True:
        osc.render_frame(opts, scene.with_jitter(cam, 1 + frames, w, h), 1 + frames, w, h, threads=threads)
        frames += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or frames >= 64:
            break
    return {"value": w * h * frames / dt, "unit": "paths/s", "cores": threads, "kind": "port",
            "sample": f"{frames} frames of the same Cornell-box workload at {w}x{h} (1/4 of the pixels), depth {MAX_BOUNCES}, {dt:.1f} s wall, oracle/liboracle.so"}


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
    """--impl reference: the reference's CPU implementation of the path. slangc is not available, so the Slang modules cannot
    be compiled for the CPU; the oracle port (C++ restatement, all host threads) stands in (DESIGN.md 'Oracle')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as ob
    from pltfalcor_b200 import layouts as L, scene
    buffers, cam = build_scene()
    opts = L.make_options(maxBounces=MAX_BOUNCES)
    osc = ob.OracleScene(buffers)
    threads = ob.hardware_threads()
    w, h = WIDTH // 4, HEIGHT // 4  # bounded sample per step: 1/16 of the frame's pixels, same camera and scene
    for f in range(args.warmup):
        osc.render_frame(opts, scene.with_jitter(cam, f, w, h), f, w, h, threads=threads)
    t0 = time.perf_counter()
    for f in range(args.steps):
        osc.render_frame(opts, scene.with_jitter(cam, 100 + f, w, h), 100 + f, w, h, threads=threads)
    dt = time.perf_counter() - t0
    value = w * h * args.steps / dt
    sample = f"each step = one frame at {w}x{h} (1/16 of the 1920x1080 pixels, same scene/camera/options)"
    emit({
        "impl": "reference", "metric": "plt_paths_per_s", "value": value, "unit": "paths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "max_bounces": MAX_BOUNCES, "hwss": 4, "sample": sample},
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timers", action="store_true", help="do not record per-launch CUDA events in the timed region (roofline falls back to stage timers)")
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
    hbm_peak, peak_src = peaks()

    buffers, cam = build_scene()
    p = binding.Pltpt(device=local, maxBounces=MAX_BOUNCES)
    p.set_scene(buffers)
    build_stats = p.stats()
    accum = torch.zeros((HEIGHT, WIDTH, 4), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)

    # ---- device-resident timing -------------------------------------------------------------------------------------
    p.render_accumulate(cam, rank, world, W, WIDTH, HEIGHT, accum, stream=sptr)  # warm-up frames (also sizes the work buffers)
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
    first, step, count = partition.frames_of_rank(rank, world, W * world, K)
    p.render_accumulate(cam, first, step, count, WIDTH, HEIGHT, accum, stream=sptr)
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
    paths_per_gpu = WIDTH * HEIGHT * K
    value = paths_per_gpu * world / (total_ms * 1e-3)

    # ---- per-kernel durations: the same K frames again with a CUDA event after every launch on the launching stream.  The timed
    # region above keeps two frames in flight on two streams (kernels of different frames share the SMs, so a per-launch duration
    # taken there would include the other frame's work); this pass runs the frames one after the other so that every kernel is
    # timed alone, which is what a roofline fraction needs.  Its frame time is reported as `sequential_ms_per_step`.
    stats = timed_stats
    seq_ms = None
    if rank == 0 and not args.no_kernel_timers:
        scratch = torch.zeros_like(accum)
        p.reset_stats()
        p.set_profiling(True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        p.render_accumulate(cam, first, step, count, WIDTH, HEIGHT, scratch, stream=sptr)
        s1.record(stream)
        torch.cuda.synchronize()
        seq_ms = s0.elapsed_time(s1) / K
        stats = p.stats()
        p.set_profiling(False)
        del scratch
    if world > 1:
        dist.barrier()

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------------------------------
    # Two ways through the same entry-point family, both timed: the synchronous call (returns with the images on the host) and the
    # queued one (pltpt_render_frame_host_async: the way PLTPT::execute records work and returns; two frames in flight, each with its
    # own pinned image set, pltpt_host_wait inside the timed region).  The queued figure is the e2e value; the synchronous one is kept beside it.
    host = [torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32, pin_memory=True) for _ in range(9)]  # three image sets
    e2e_steps = 3 * ((max(4, min(K, 64)) + 2) // 3)   # a multiple of the three image sets, so both runs end on the same set
    wait = binding.lib().pltpt_host_wait

    def e2e_step(f, i, queued):
        c = scene.with_jitter(cam, f, WIDTH, HEIGHT)
        b = 3 * (i % 3)
        if queued:
            # the consumer pattern of include/pltpt.h: queue frame i, then wait until only two frames are on their way, i.e. frame
            # i-2 is in its host buffers and its image set is free for frame i+1
            binding._check(binding.lib().pltpt_render_frame_host_async(p.h, f, C.byref(c), WIDTH, HEIGHT, host[b].data_ptr(), host[b + 1].data_ptr(), host[b + 2].data_ptr()))
            binding._check(wait(p.h, 2))
        else:
            binding._check(binding.lib().pltpt_render_frame_host(p.h, f, C.byref(c), WIDTH, HEIGHT, host[b].data_ptr(), host[b + 1].data_ptr(), host[b + 2].data_ptr()))

    def e2e_run(queued):
        for f in range(4):
            e2e_step(rank + f * world, f, queued)
        binding._check(wait(p.h, 0))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f in range(e2e_steps):
            e2e_step(1000 + rank + f * world, f, queued)
        binding._check(wait(p.h, 0))
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return WIDTH * HEIGHT * e2e_steps * world / float(t.item())

    e2e_sync_value = e2e_run(False)
    sync_image = host[6].clone()   # last synchronous frame (third image set)
    e2e_value = e2e_run(True)
    e2e_queued_equals_sync = bool(torch.equal(sync_image, host[6]))  # same frame index, same jitter: must be the same bits
    checksum = float(host[6][..., :3].double().mean())

    if rank == 0:
        pb = path_bytes(stats, K, paths_per_gpu)
        frames = max(stats["framesRendered"], 1)
        solve_ms = stats["solveStageMs"] / frames
        sample_ms = stats["sampleStageMs"] / frames
        kms, kcount = stats["kernelMs"], stats["kernelLaunchCount"]
        kernels = {k: {"ms_per_step": kms[k] / frames, "launches_per_step": kcount[k] / frames} for k in kms if kcount[k]}
        timed_sum = sum(v["ms_per_step"] for v in kernels.values())
        for v in kernels.values():
            v["share_of_step"] = v["ms_per_step"] / timed_sum if timed_sum else None
        # dominant kernel: k_unwind_step.  Unit of work = one generalized ray pushed through one recorded vertex (R + F bytes)
        unit_bytes = R_BOUNCE + F_VERTEX_FETCH
        if kcount["unwind_step"]:
            k_ms_total, k_launches, units = kms["unwind_step"], kcount["unwind_step"], stats["unwindJobSteps"]
            roof = {"kernel": "k_unwind_step", "kernel_ms": k_ms_total / k_launches, "launches": int(k_launches), "units_per_launch": units / k_launches,
                    "algorithmic_bytes_per_unit": unit_bytes, "achieved": units * unit_bytes / (k_ms_total * 1e-3) / 1e9, "share_of_step": kernels["unwind_step"]["share_of_step"],
                    "timer": "CUDA events after every launch on the launching stream; same K frames as the timed region, run one frame at a time (sequential_ms_per_step) so that the kernel is timed alone"}
        else:  # --no-kernel-timers: whole solve stage from the stage timers
            roof = {"kernel": "solve stage (k_connect_*, k_shadow_trace, k_mnee, k_sort_queue, k_unwind_step, k_resolve)", "kernel_ms": solve_ms, "launches": int(frames),
                    "units_per_launch": WIDTH * HEIGHT, "algorithmic_bytes_per_unit": pb["solve_bytes"], "achieved": pb["solve_bytes"] * WIDTH * HEIGHT / (solve_ms * 1e-3) / 1e9,
                    "share_of_step": solve_ms / (solve_ms + sample_ms), "timer": "stage CUDA events"}
        traffic = ncu_traffic(roof["kernel"])
        out = {
            "metric": "plt_paths_per_s", "value": value, "unit": "paths/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "max_bounces": MAX_BOUNCES, "hwss": 4, "nee": True, "mis": True, "mnee": True, "russian_roulette": True, "frames_in_flight": 2,
                       "paths_per_step_per_gpu": WIDTH * HEIGHT, "partition": "frame index modulo n_gpus; fp64 accumulators summed by one NCCL reduce inside the timed region",
                       "l2": "inputs larger than L2: every frame rewrites and re-reads its %.0f MB bounce buffer (126 MB L2)" % ((MAX_BOUNCES + 1) * WIDTH * HEIGHT * 40 / 1e6),
                       "triangles": int(build_stats["triangleCount"]), "bvh_nodes": int(build_stats["bvhNodes"]), "bvh_build_ms": build_stats["bvhBuildMs"]},
            "e2e": {"value": e2e_value, "unit": "paths/s", "h2d_bytes_per_step": C.sizeof(L.PltptCameraData), "d2h_bytes_per_step": 3 * WIDTH * HEIGHT * 16, "steps": e2e_steps,
                    "api": "pltpt_render_frame_host_async + pltpt_host_wait (camera in, color/albedo/normal RGBA32F out per frame into pinned host buffers, two frames in flight, pltpt_host_wait(h, 2) after every frame)",
                    "synchronous_value": e2e_sync_value, "synchronous_api": "pltpt_render_frame_host (returns with the three images on the host)",
                    "queued_equals_synchronous": e2e_queued_equals_sync, "checksum_mean_rgb": checksum},
            "gpu_launches": int(timed_stats["kernelLaunches"]),
            "sequential_ms_per_step": seq_ms,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": roof["achieved"], "peak": hbm_peak, "unit": "GB/s", "frac": roof["achieved"] / hbm_peak, "traffic": traffic,
                         "peak_source": peak_src, **{k: v for k, v in roof.items() if k != "achieved"},
                         "note": "shading kernels are issue-bound (fp32 + MUFU, -fmad=false), not HBM-bound: see profiles/ for the ncu issue-slot and warp-efficiency evidence",
                         "stages": {"sample_ms": sample_ms, "solve_ms": solve_ms, "sample_bytes_per_path": pb["sample_bytes"] + pb["bvh_bytes"], "solve_bytes_per_path": pb["solve_bytes"],
                                    "whole_step_achieved": pb["total"] * WIDTH * HEIGHT / ((sample_ms + solve_ms) * 1e-3) / 1e9}},
            "kernels": kernels,
            "path_stats": {"mean_vertices": pb["mean_vertices"], "mean_replay_steps": pb["mean_replay_steps"], "bsdf_evals_per_path": stats["totalBsdfEvals"] / paths_per_gpu,
                           "bsdf_evals_per_s": stats["totalBsdfEvals"] * world / (total_ms * 1e-3), "bvh_nodes_per_path": stats["traversalNodesVisited"] / paths_per_gpu,
                           "bvh_triangles_per_path": stats["traversalTrianglesTested"] / paths_per_gpu, "algorithmic_bytes_per_path": pb["total"]},
        }
        if world == 1:
            sw_buffers, ids = scene.sweep_scene()
            q = binding.Pltpt(device=local)
            q.set_scene(sw_buffers)
            out["bsdf_sweep"] = bsdf_sweep(torch, q, ids, hbm_peak)
            q.close()
            if not args.no_cpu_baseline:
                out["cpu_baseline"] = cpu_baseline_sample(buffers, cam, p.options)
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
