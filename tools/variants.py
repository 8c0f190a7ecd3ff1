"""Builds experimental variants of libpltpt.so (extra nvcc flags per source) under pltfalcor_b200/build/variants/<name>/ and times
them on the bench workload.  Development tool: the product build is pltfalcor_b200/build.py.

  python tools/variants.py build            # here (cross-compile)
  python tools/variants.py bench [frames]   # on the GPU box: per-kernel ms of one 1080p Cornell frame for every built variant
  python tools/variants.py one <lib> [frames]
"""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from pltfalcor_b200 import build as B  # noqa: E402

VDIR = os.path.join(B._HERE, "build", "variants")
_FM = ["-prec-div=false", "-prec-sqrt=false", "-fmad=true"]
_TRAV = ("stage_sample.cu", "stage_connect.cu", "bvh.cu")   # translation units that instantiate the traversal
VARIANTS = {  # name -> {source: extra nvcc flags}; the experiments and their outcomes are listed in DESIGN.md section 3 (the log behind each in profiles/)
    "base": {},
    # -- replay step (product: stage_unwind.cu with _FM + PLT_FAST_MATH, pltfalcor_b200/build.py)
    "unwind_exact": {"stage_unwind.cu": ["-prec-div=true", "-prec-sqrt=true", "-fmad=false", "-DPLT_FAST_MATH=0"]},   # rounds 1-2a: the oracle's rounding in the replay too
    "stage": {"stage_unwind.cu": ["-DPLT_UNWIND_STAGE=1"]},                                                            # cp.async staging of the thread's next item (r2a: worse)
    "ufm_u": {"stage_unwind.cu": ["-use_fast_math"]},                                                                  # r2h: nothing beyond _FM
    "fm_mb5": {"stage_unwind.cu": ["-DPLT_MINB_UNWIND_STEP=5"]}, "fm_mb3": {"stage_unwind.cu": ["-DPLT_MINB_UNWIND_STEP=3"]},  # r2i: 96 / 160 registers
    "fm_unroll": {"stage_unwind.cu": ["-DPLT_LAMBDA_UNROLL=1"]},                                                       # r2i: unrolled wavelength loops
    # -- the same arithmetic elsewhere (rejected: parity)
    "fm_uc": {"stage_connect.cu": _FM}, "c_fmad": {"stage_connect.cu": ["-fmad=true"]}, "c_div": {"stage_connect.cu": ["-prec-div=false"]}, "c_sqrt": {"stage_connect.cu": ["-prec-sqrt=false"]},  # r2h, r2j
    "sweep_fm": {"api.cu": _FM + ["-DPLT_FAST_MATH=1"]},                                                               # r2n: 27 G evals/s, misses the 1e-4 BSDF tolerance
    # -- traversal (product: deferred leaf tests + nearest-first push in the persistent driver, 8 idle lanes to refill, 4 node steps per round)
    "trav_legacy": {f: ["-DPLT_DEFER_LEAVES=0"] for f in _TRAV},                                                       # the round-1 driver (r2o, r2p)
    "trav_r4": {f: ["-DPLT_TRAV_REFILL=4"] for f in _TRAV}, "trav_r16": {f: ["-DPLT_TRAV_REFILL=16"] for f in _TRAV},
    "trav_s2": {f: ["-DPLT_TRAV_STEPS=2"] for f in _TRAV}, "trav_s6": {f: ["-DPLT_TRAV_STEPS=6"] for f in _TRAV}, "trav_s8": {f: ["-DPLT_TRAV_STEPS=8"] for f in _TRAV},
    # -- BVH builder (product: PLOC radius 8 + SAH-optimal collapse)
    "bvh_lbvh": {"bvh.cu": ["-DPLT_BVH_PLOC=0"]}, "ploc_r16": {"bvh.cu": ["-DPLT_PLOC_RADIUS=16"]}, "ploc_r4": {"bvh.cu": ["-DPLT_PLOC_RADIUS=4"]},   # r2q
    "mneeprof": {"stage_connect.cu": ["-DPLT_MNEE_PHASE_TIMERS=1"]},               # per-round phase timers of k_mnee_walk (tools/mnee_profile.py)
}


def build_variant(name, flags):
    d = os.path.join(VDIR, name)
    os.makedirs(d, exist_ok=True)
    nvcc = B._nvcc()

    def one(src):
        obj = os.path.join(d, src.replace(".cu", ".o"))
        extra = flags.get(src, [])
        if extra:
            extra = B.SOURCE_FLAGS.get(src, []) + extra   # (a later -D of the same macro wins)
        # a later -opt=value / -DNAME=value of the same option or macro wins (nvcc rejects a repeated single-value option)
        last = {}
        for f in extra:
            last[f.split("=")[0]] = f
        extra = [f for f in extra if last[f.split("=")[0]] is f]
        base = [f for f in B.NVCC_FLAGS if f.split("=")[0] not in last or not f.startswith("-fmad")]
        if not extra:  # unchanged source: reuse the product object when it is fresh
            prod = os.path.join(B._HERE, "build", src.replace(".cu", ".o"))
            if os.path.exists(prod) and all(os.path.getmtime(x) <= os.path.getmtime(prod) for x in B._deps()):
                return prod
        r = subprocess.run([nvcc, *base, *extra, "-c", os.path.join(B.CSRC, src), "-o", obj], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(r.stderr)
        return obj

    with ThreadPoolExecutor(len(B.SOURCES)) as ex:
        objs = list(ex.map(one, B.SOURCES))
    lib = os.path.join(d, "libpltpt.so")
    r = subprocess.run([nvcc, "-shared", "-o", lib, *objs, "-lcudart_static", "-Xlinker", "--no-undefined"], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr)
    return lib


def bench_one(lib, frames):
    import torch
    from pltfalcor_b200 import binding, layouts as L, scene
    binding.LIB = lib
    W, H = 1920, 1080
    buffers, cam = scene.cornell_box(aspect=W / H)
    p = binding.Pltpt(maxBounces=8)
    p.set_scene(buffers)
    acc = torch.zeros((H, W, 4), dtype=torch.float64, device="cuda")
    p.render_accumulate(cam, 0, 1, 3, W, H, acc)
    torch.cuda.synchronize()
    acc.zero_()
    p.reset_stats()
    def timed():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        p.render_accumulate(cam, 3, 1, frames, W, H, acc)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / frames

    timed()  # sizes the second lane's buffers
    ms = timed()
    ms_seq = None
    if hasattr(binding.lib(), "pltpt_set_frame_overlap"):
        p.set_frame_overlap(1)
        ms_seq = timed()
        p.set_frame_overlap(0)
    # frame-by-frame host-buffer call (what bench.py's e2e times)
    host = [torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True) for _ in range(3)]
    import ctypes as C
    import time
    def e2e(n, first):
        for f in range(first, first + n):
            c = scene.with_jitter(cam, f, W, H)
            binding._check(binding.lib().pltpt_render_frame_host(p.h, f, C.byref(c), W, H, host[0].data_ptr(), host[1].data_ptr(), host[2].data_ptr()))
    e2e(3, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e(frames, 100)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / frames
    p.set_profiling(True)
    p.reset_stats()
    p.render_accumulate(cam, 3, 1, frames, W, H, acc)
    torch.cuda.synchronize()
    s = p.stats()
    k = {n: round(s["kernelMs"][n] / frames, 3) for n in L.KERNEL_CLASSES if s["kernelLaunchCount"][n]}
    import bench
    sb, ids = scene.sweep_scene()
    q = binding.Pltpt()
    q.set_scene(sb)
    sweep = {k: round(v["evals_per_s"] / 1e9, 3) for k, v in bench.bsdf_sweep(torch, q, ids, 6543.7).items()}
    print(json.dumps({"lib": os.path.relpath(lib, ROOT), "sweep_Gevals_s": sweep, "ms_per_frame": round(ms, 3), "e2e_ms": round(e2e_ms, 3), "ms_sequential": ms_seq and round(ms_seq, 3), "Mpaths_s": round(W * H / ms / 1e3, 1), "mean_rgb": float(acc[..., :3].mean() / acc[..., 3].mean()), "kernels": k}))


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "build"
    if cmd == "build":
        B.build()
        names = sys.argv[2:] or list(VARIANTS)
        for n in names:
            print(n, build_variant(n, VARIANTS[n]))
    elif cmd == "bench":
        frames = sys.argv[2] if len(sys.argv) > 2 else "16"
        for n in sorted(os.listdir(VDIR)):
            lib = os.path.join(VDIR, n, "libpltpt.so")
            if os.path.exists(lib):
                subprocess.run([sys.executable, os.path.abspath(__file__), "one", lib, frames])
    elif cmd == "one":
        bench_one(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 16)
