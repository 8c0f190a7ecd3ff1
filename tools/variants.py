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
VARIANTS = {  # name -> {source: extra nvcc flags}; the experiments and their outcomes are listed in DESIGN.md section 3
    "base": {},
    "stage": {"stage_unwind.cu": ["-DPLT_UNWIND_STAGE=1"]},                          # k_unwind_step with cp.async staging of the thread's next item
    "fm_flags": {"stage_unwind.cu": ["-prec-div=false", "-prec-sqrt=false", "-fmad=true"]},
    "fm_u": {"stage_unwind.cu": ["-prec-div=false", "-prec-sqrt=false", "-fmad=true", "-DPLT_FAST_MATH=1"]},
    "fm_uc": {"stage_unwind.cu": ["-prec-div=false", "-prec-sqrt=false", "-fmad=true", "-DPLT_FAST_MATH=1"], "stage_connect.cu": ["-prec-div=false", "-prec-sqrt=false", "-fmad=true"]},
    "ufm_u": {"stage_unwind.cu": ["-use_fast_math", "-DPLT_FAST_MATH=1"]},
    "fm_mb5": {"stage_unwind.cu": ["-prec-div=false", "-prec-sqrt=false", "-fmad=true", "-DPLT_FAST_MATH=1", "-DPLT_MINB_UNWIND_STEP=5"]},
    "fm_mb3": {"stage_unwind.cu": ["-prec-div=false", "-prec-sqrt=false", "-fmad=true", "-DPLT_FAST_MATH=1", "-DPLT_MINB_UNWIND_STEP=3"]},
    "fm_unroll": {"stage_unwind.cu": ["-prec-div=false", "-prec-sqrt=false", "-fmad=true", "-DPLT_FAST_MATH=1", "-DPLT_LAMBDA_UNROLL=1"]},
    "c_fmad": {"stage_unwind.cu": ["-prec-div=false", "-prec-sqrt=false", "-fmad=true", "-DPLT_FAST_MATH=1"], "stage_connect.cu": ["-fmad=true"]},
    "c_div": {"stage_unwind.cu": ["-prec-div=false", "-prec-sqrt=false", "-fmad=true", "-DPLT_FAST_MATH=1"], "stage_connect.cu": ["-prec-div=false"]},
    "c_sqrt": {"stage_unwind.cu": ["-prec-div=false", "-prec-sqrt=false", "-fmad=true", "-DPLT_FAST_MATH=1"], "stage_connect.cu": ["-prec-sqrt=false"]},
    "nrm_rcp": {"stage_sample.cu": ["-DPLT_NORMALIZE_RCP=1"], "stage_connect.cu": ["-DPLT_NORMALIZE_RCP=1"]},  # (breaks oracle parity unless the oracle follows)
    "trav_r4": {"stage_sample.cu": ["-DPLT_TRAV_REFILL=4"], "stage_connect.cu": ["-DPLT_TRAV_REFILL=4"]},
    "trav_r16": {"stage_sample.cu": ["-DPLT_TRAV_REFILL=16"], "stage_connect.cu": ["-DPLT_TRAV_REFILL=16"]},
    "trav_s2": {"stage_sample.cu": ["-DPLT_TRAV_STEPS=2"], "stage_connect.cu": ["-DPLT_TRAV_STEPS=2"]},
    "trav_s8": {"stage_sample.cu": ["-DPLT_TRAV_STEPS=8"], "stage_connect.cu": ["-DPLT_TRAV_STEPS=8"]},
    "sweep_fm": {"api.cu": ["-prec-div=false", "-prec-sqrt=false", "-fmad=true", "-DPLT_FAST_MATH=1"]},
    "push_near": {"stage_sample.cu": ["-DPLT_PUSH_NEAREST=1"], "stage_connect.cu": ["-DPLT_PUSH_NEAREST=1"], "bvh.cu": ["-DPLT_PUSH_NEAREST=1"]},
    "defer": {"stage_sample.cu": ["-DPLT_DEFER_LEAVES=1"], "stage_connect.cu": ["-DPLT_DEFER_LEAVES=1"], "bvh.cu": ["-DPLT_DEFER_LEAVES=1"]},
    "defer_push": {"stage_sample.cu": ["-DPLT_DEFER_LEAVES=1", "-DPLT_PUSH_NEAREST=1"], "stage_connect.cu": ["-DPLT_DEFER_LEAVES=1", "-DPLT_PUSH_NEAREST=1"], "bvh.cu": ["-DPLT_DEFER_LEAVES=1", "-DPLT_PUSH_NEAREST=1"]},
    "trav_legacy": {"stage_sample.cu": ["-DPLT_DEFER_LEAVES=0"], "stage_connect.cu": ["-DPLT_DEFER_LEAVES=0"], "bvh.cu": ["-DPLT_DEFER_LEAVES=0"]},
    "defer_s8": {"stage_sample.cu": ["-DPLT_TRAV_STEPS=8"], "stage_connect.cu": ["-DPLT_TRAV_STEPS=8"], "bvh.cu": ["-DPLT_TRAV_STEPS=8"]},
    "defer_s6": {"stage_sample.cu": ["-DPLT_TRAV_STEPS=6"], "stage_connect.cu": ["-DPLT_TRAV_STEPS=6"], "bvh.cu": ["-DPLT_TRAV_STEPS=6"]},
    "bvh_lbvh": {"bvh.cu": ["-DPLT_BVH_PLOC=0"]},
    "ploc_r16": {"bvh.cu": ["-DPLT_PLOC_RADIUS=16"]},
    "ploc_r4": {"bvh.cu": ["-DPLT_PLOC_RADIUS=4"]},
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
        base = [f for f in B.NVCC_FLAGS if not (f == "-fmad=false" and "-fmad=true" in extra)]
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
