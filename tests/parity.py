"""Shared helpers for the GPU-vs-oracle parity tests (tests/test_gpu_*.py) and tools/gpu_diag.py."""
import numpy as np

from pltfalcor_b200 import layouts as L


def random_dirs(rng, n, upper=False):
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    if upper:
        d[:, 2] = np.abs(d[:, 2])
    return d.astype(np.float32)


def camera_rays(cam, width, height):
    """Pinhole rays exactly as Camera.slang:46-78 forms them (fp32)."""
    f32 = np.float32
    px, py = np.meshgrid(np.arange(width), np.arange(height))
    p = np.stack([(px.astype(f32) + f32(0.5)) / f32(width), (py.astype(f32) + f32(0.5)) / f32(height)], -1)
    ndc = p * np.array([2, -2], f32) + np.array([-1, 1], f32)
    U, V, W = (np.array(list(getattr(cam, k)), f32) for k in ("cameraU", "cameraV", "cameraW"))
    d = ndc[..., :1] * U + ndc[..., 1:] * V + W
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(f32).reshape(-1, 3)
    rays = np.zeros((len(d), 8), f32)
    rays[:, 0:3] = np.array(list(cam.posW), f32)
    rays[:, 3] = 0.0
    rays[:, 4:7] = d
    rays[:, 7] = np.finfo(f32).max
    return rays


def scene_bounds(buffers):
    """World-space bounds of the procedural scenes (all instance matrices are identity or affine; use transformed vertices)."""
    a = buffers.arrays
    P = a["vertices"]["position"].astype(np.float64)
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    inst = a["geometryInstances"]
    meshes = a["meshes"]
    M = a["worldMatrices"].reshape(-1, 4, 4).astype(np.float64)
    for g in inst:
        md = meshes[g["geometryID"]]
        v = P[g["vbOffset"]: g["vbOffset"] + md["vertexCount"]]
        w = v @ M[g["globalMatrixID"]][:3, :3].T + M[g["globalMatrixID"]][:3, 3]
        lo, hi = np.minimum(lo, w.min(0)), np.maximum(hi, w.max(0))
    return lo, hi


def random_rays(rng, buffers, n, tmax=None):
    lo, hi = scene_bounds(buffers)
    ext = hi - lo
    o = (lo - 0.1 * ext + rng.uniform(size=(n, 3)) * 1.2 * ext).astype(np.float32)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = o
    rays[:, 4:7] = random_dirs(rng, n)
    rays[:, 3] = 0.0
    rays[:, 7] = np.finfo(np.float32).max if tmax is None else rng.uniform(0.05, 1.0, n).astype(np.float32) * np.float32(np.linalg.norm(ext)) * tmax
    return rays


ABS_FLOOR = 1e-8

# Measured parity figures of a test session: tests call record(); tests/conftest.py writes them to gpurun_out/parity_metrics.json at the
# end of the session (and the converged-image figures to gpurun_out/image_parity.json), from where they are committed under profiles/.
METRICS = {}
IMAGE_PARITY = {}


def record(key, value):
    METRICS[key] = value



def mueller_rel_err(a, b, columns=None):
    """Per Mueller matrix: max |a-b| / (max|b| + ABS_FLOOR / 1e-4), i.e. the test `err <= 1e-4` reads
    |a-b| <= 1e-4 * max|b| + 1e-8.  The absolute floor only matters for matrices whose entries are < 1e-4 (total destructive
    interference in the multilayer stack, PSD tails): BSDF*cos values that contribute nothing next to O(1e-3..1e3) lobes.
    `columns` restricts the comparison to some input Stokes components (column c = response to basis vector e_c)."""
    a = a.reshape(-1, 4, 4).astype(np.float64)
    b = b.reshape(-1, 4, 4).astype(np.float64)
    scale = np.abs(b).reshape(-1, 16).max(axis=1) + ABS_FLOOR / 1e-4
    if columns is not None:
        a, b = a[:, :, columns], b[:, :, columns]
    d = np.abs(a - b).reshape(len(a), -1)
    fa, fb = np.isfinite(a).reshape(len(a), -1).all(axis=1), np.isfinite(b).reshape(len(b), -1).all(axis=1)
    err = np.where(fa & fb, np.nan_to_num(d).max(axis=1) / scale, 0.0)
    err[fa != fb] = np.inf
    return err


def compare_bounces(gpu, cpu):
    """gpu, cpu: (H, W, R) arrays of layouts.bounce_dtype. Returns (metrics dict, per-pixel identical-path mask).
    Only the records up to and including the terminating one (miss / terminated flag) are specified."""
    R = gpu.shape[-1]
    ended = (cpu["flags"] & 0x30) != 0
    first_end = np.where(ended.any(-1), ended.argmax(-1), R - 1)
    live = np.arange(R)[None, None, :] <= first_end[..., None]
    rec_same = (gpu["hit"][..., 0] == cpu["hit"][..., 0]) & (gpu["hit"][..., 1] == cpu["hit"][..., 1]) & (gpu["sampledBsdfLobe"] == cpu["sampledBsdfLobe"]) & (gpu["flags"] == cpu["flags"])
    same = (rec_same | ~live).all(axis=-1)
    m = same[..., None] & live
    out = {"pixels": int(same.size), "identical_paths": float(same.mean()), "mean_records": float(live.sum(-1).mean())}
    if m.any():
        bary_g = gpu["hit"][..., 2:].copy().view(np.float32)[m]
        bary_c = cpu["hit"][..., 2:].copy().view(np.float32)[m]
        out["max_bary_abs"] = float(np.abs(bary_g - bary_c).max())
        out["max_wi_abs"] = float(np.abs(gpu["encodedWi"][m] - cpu["encodedWi"][m]).max())
        pc_g, pc_c = gpu["pathContribution"][m].astype(np.float64), cpu["pathContribution"][m].astype(np.float64)
        out["max_contribution_rel"] = float((np.abs(pc_g - pc_c) / np.maximum(np.abs(pc_c), 1e-20)).max())
        lp_g, lp_c = gpu["lastNonDeltaPdf"][m].astype(np.float64), cpu["lastNonDeltaPdf"][m].astype(np.float64)
        out["max_lastpdf_rel"] = float((np.abs(lp_g - lp_c) / np.maximum(np.abs(lp_c), 1e-6)).max())
    return out, same


def compare_color(gpu, cpu, same_paths):
    g = gpu[..., :3].astype(np.float64)
    c = cpu[..., :3].astype(np.float64)
    scale = np.maximum(np.abs(c).max(axis=-1), 1e-4 * max(c.mean(), 1e-12))
    rel = np.abs(g - c).max(axis=-1) / scale
    m = same_paths
    return {
        "mean_gpu": float(g.mean()), "mean_cpu": float(c.mean()),
        "mean_rel_diff": float(abs(g.mean() - c.mean()) / max(c.mean(), 1e-20)),
        "frac_pixels_within_1e-3": float((rel[m] <= 1e-3).mean()) if m.any() else 0.0,
        "frac_pixels_within_1e-2": float((rel[m] <= 1e-2).mean()) if m.any() else 0.0,
        "p99_rel": float(np.quantile(rel[m], 0.99)) if m.any() else float("nan"),
        "nonfinite_gpu": int((~np.isfinite(g)).sum()),
    }


def rel_mse(a, b):
    """relMSE as the reference's ImageCompare defines it (Source/Tools/ImageCompare/ImageCompare.cpp:198-207)."""
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(((a - b) ** 2 / (b ** 2 + 1e-3)).mean())


f32, f64 = np.float32, np.float64


def accumulate_test_frames(mode, h=19, w=33, n=40):
    """Seeded input frames of the AccumulatePass tests: values spanning eight decades -- the single-precision sum visibly loses bits, the compensated
    and double ones must not."""
    rng = np.random.default_rng(7 + mode)
    return [(rng.uniform(0, 1, (h, w, 4)) * 10.0 ** rng.integers(-4, 4, (h, w, 1))).astype(f32) for _ in range(n)]


def accumulate_pass_restatement(mode, frames, moving_average):
    """AccumulatePass stepped frame by frame (Accumulate.cs.slang:73-198; mode 0 Single, 1 SingleCompensated, 2 Double): the same fp32 / fp64 operations in
    the same order, in numpy.  Held to the reference's own shader text by tests/test_oracle_ref_frames.py (tests/golden/ref_accumulate.npz)."""
    h, w, _ = frames[0].shape
    if mode == 2:
        s = np.zeros((h, w, 4), f64)
    else:
        s = np.zeros((h, w, 4), f32)
    corr = np.zeros((h, w, 4), f32)
    outs = []
    for n, c in enumerate(frames):
        if mode == 0:
            wgt = f32(1.0) / f32(n + 1)
            if moving_average:
                o = (s + (c - s) * wgt).astype(f32)
                s = o
            else:
                s = (s + c).astype(f32)
                o = (s * wgt).astype(f32)
        elif mode == 1:
            y = (c - corr).astype(f32)
            nxt = (s + y).astype(f32)
            o = (nxt / f32(n + 1)).astype(f32)
            corr = ((nxt - s).astype(f32) - y).astype(f32)
            s = nxt
        else:
            wgt = f64(1.0) / f64(n + 1)
            if moving_average:
                s = s * (f64(1.0) - wgt) + c.astype(f64) * wgt
                o = s.astype(f32)
            else:
                s = s + c.astype(f64)
                o = (s * wgt).astype(f32)
        outs.append(o)
    return outs


def snorm2x16_steps(a, b):
    """Largest distance, in 16-bit snorm quantisation steps, between two arrays of 2x16-snorm packed words (octahedral normals)."""
    a, b = np.asarray(a, np.uint32), np.asarray(b, np.uint32)
    worst = 0
    for shift in (0, 16):
        da = ((a >> shift) & 0xffff).astype(np.int64)
        db = ((b >> shift) & 0xffff).astype(np.int64)
        da, db = np.where(da >= 32768, da - 65536, da), np.where(db >= 32768, db - 65536, db)
        worst = max(worst, int(np.abs(da - db).max()) if da.size else 0)
    return worst
