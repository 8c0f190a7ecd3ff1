"""Development diagnostics on the GPU box: prints GPU-vs-oracle error distributions with the worst-case inputs.
Usage (under gpurun):  python tools/gpu_diag.py [bsdf] [sample] [bounce] [color]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import binding as ob  # noqa: E402  (diagnostic tool = test infrastructure)
from pltfalcor_b200 import binding, layouts as L, scene  # noqa: E402
from tests import parity  # noqa: E402

np.set_printoptions(precision=6, linewidth=200)
what = set(sys.argv[1:]) or {"bsdf", "sample", "bounce", "color"}


def eval_both(p, osc, mid, lobe, wi, wo, hero):
    out = torch.full((len(wi), hero.shape[0], 4, 4, 4), float("nan"), device="cuda")
    p.bsdf_eval_mueller(mid, lobe, torch.from_numpy(wi).cuda(), torch.from_numpy(wo).cuda(), torch.from_numpy(hero).cuda(), out)
    torch.cuda.synchronize()
    return out.cpu().numpy(), osc.bsdf_eval_mueller(mid, lobe, wi, wo, hero)


def report(name, got, ref, wi, wo, hero):
    err = parity.mueller_rel_err(got, ref)
    nz = np.abs(ref.reshape(-1, 16)).max(axis=1) > 0
    print(f"{name:34s} nonzero {nz.mean():.3f}  max {err.max():.2e}  p99.9 {np.quantile(err, 0.999):.2e}  p99 {np.quantile(err, 0.99):.2e}  median {np.median(err[nz]) if nz.any() else 0:.2e}  frac>1e-4 {(err > 1e-4).mean():.4f}")
    if err.max() > 1e-4:
        w = int(np.argmax(err))
        q, g, j = w // (hero.shape[0] * 4), (w // 4) % hero.shape[0], w % 4
        print(f"    worst: query {q} group {g} lane {j} wi {wi[q]} wo {wo[q]} hero {hero[g, q]}")
        print("    GPU", got.reshape(-1, 16)[w])
        print("    CPU", ref.reshape(-1, 16)[w])


if "bsdf" in what or "sample" in what:
    buffers, ids = scene.sweep_scene()
    p = binding.Pltpt()
    p.set_scene(buffers)
    osc = ob.OracleScene(buffers)
    rng = np.random.default_rng(1)
    N = 4096
    if "bsdf" in what:
        for key, mid in ids.items():
            if key == "emitter":
                continue
            wi = parity.random_dirs(rng, N, upper=True)
            hero = rng.uniform(380, 720, (4, N)).astype(np.float32)
            wo = parity.random_dirs(rng, N, upper=True)
            got, ref = eval_both(p, osc, mid, L.LOBE_NON_DELTA, wi, wo, hero)
            report(key + " nondelta random", got, ref, wi, wo, hero)
            s = osc.bsdf_sample(mid, L.LOBE_NON_DELTA, wi, hero[0], seed=5)
            ok = s[:, 6] > 0
            if ok.sum() > 16:
                wo2 = np.ascontiguousarray(s[ok, 0:3])
                got, ref = eval_both(p, osc, mid, L.LOBE_NON_DELTA, wi[ok], wo2, np.ascontiguousarray(hero[:, ok]))
                report(key + " nondelta sampled-wo", got, ref, wi[ok], wo2, hero[:, ok])
            wr = np.stack([-wi[:, 0], -wi[:, 1], wi[:, 2]], -1).astype(np.float32)
            got, ref = eval_both(p, osc, mid, L.LOBE_DELTA_REFLECTION, wi, wr, hero)
            report(key + " delta-reflection", got, ref, wi, wr, hero)
    if "sample" in what:
        for key, mid in ids.items():
            if key == "emitter":
                continue
            n = 8192
            wi = parity.random_dirs(rng, n, upper=True)
            wl = rng.uniform(380, 720, n).astype(np.float32)
            out = torch.zeros((n, 8), device="cuda")
            p.bsdf_sample(mid, L.LOBE_ALL, torch.from_numpy(wi).cuda(), torch.from_numpy(wl).cuda(), 17, out)
            torch.cuda.synchronize()
            got = out.cpu().numpy()
            ref = osc.bsdf_sample(mid, L.LOBE_ALL, wi, wl, seed=17)
            agree = (got[:, 6] == ref[:, 6]) & (got[:, 4].view(np.uint32) == ref[:, 4].view(np.uint32))
            m = agree & (ref[:, 6] > 0)
            line = f"{key:22s} agree {agree.mean():.4f} valid {m.mean():.3f}"
            if m.any():
                line += f" wo {np.abs(got[m, 0:3] - ref[m, 0:3]).max():.2e}"
                for col, nm in ((3, "pdf"), (5, "weight"), (7, "pdf()")):
                    rel = np.abs(got[m, col].astype(np.float64) - ref[m, col]) / np.maximum(np.abs(ref[m, col]), 1e-6)
                    line += f" {nm} max {rel.max():.2e} p99.9 {np.quantile(rel, 0.999):.2e}"
                    if nm == "pdf()" and rel.max() > 1e-3:
                        w = np.nonzero(m)[0][int(np.argmax(rel))]
                        line += f"\n    worst pdf(): wi {wi[w]} wl {wl[w]} GPU {got[w]} CPU {ref[w]}"
            print(line)
            bad = np.nonzero(~agree)[0][:3]
            for b in bad:
                print("    disagree:", wi[b], wl[b], "GPU", got[b], got[b, 4:5].view(np.uint32), "CPU", ref[b], ref[b, 4:5].view(np.uint32))
    p.close()

if "bounce" in what or "color" in what:
    for name, fn in (("cornell", scene.cornell_box), ("grating", scene.grating_sphere_scene), ("layered", scene.layered_stress_scene)):
        buffers, cam = fn()
        p = binding.Pltpt(maxBounces=8)
        p.set_scene(buffers)
        osc = ob.OracleScene(buffers)
        W, H, frame = 96, 54, 2
        c = scene.with_jitter(cam, frame, W, H)
        gb = p.sample_stage_host(c, frame, W, H)
        cpu = osc.render_frame(p.options, c, frame, W, H, want_bounces=True)
        m, same = parity.compare_bounces(gb, cpu["bounces"])
        print(name, "bounces:", m)
        ys, xs = np.nonzero(~same)
        for y, x in list(zip(ys, xs))[:4]:
            print(f"  pixel ({x},{y})")
            for r in range(9):
                g, cc = gb[y, x, r], cpu["bounces"][y, x, r]
                print(f"    r{r} GPU hit {g['hit'][0]:#x},{g['hit'][1]} bary {g['hit'][2:].view(np.float32)} lobe {g['sampledBsdfLobe']:#x} flags {g['flags']:#x} pc {g['pathContribution']:.6g}"
                      f" | CPU hit {cc['hit'][0]:#x},{cc['hit'][1]} bary {cc['hit'][2:].view(np.float32)} lobe {cc['sampledBsdfLobe']:#x} flags {cc['flags']:#x} pc {cc['pathContribution']:.6g}")
                if (cc["flags"] & 0x30) and (g["flags"] & 0x30):
                    break
        if "color" in what:
            gpu = p.render_frame_host(c, frame, W, H)
            mc = parity.compare_color(gpu["color"], cpu["color"], same)
            print(name, "color:", mc, "stats", p.stats())
            g, cc = gpu["color"][..., :3].astype(np.float64), cpu["color"][..., :3].astype(np.float64)
            rel = np.abs(g - cc).max(-1) / np.maximum(np.abs(cc).max(-1), 1e-4 * cc.mean())
            rel[~same] = 0
            for i in np.argsort(rel.ravel())[::-1][:4]:
                y, x = divmod(int(i), W)
                print(f"  worst color pixel ({x},{y}) rel {rel[y, x]:.3e} GPU {g[y, x]} CPU {cc[y, x]}")
        p.close()
