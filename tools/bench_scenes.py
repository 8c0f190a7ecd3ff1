"""Throughput of BASELINE configs[2..4] at their quoted resolutions with the product's two frames in flight (profiling off):
grating + sphere (config 3, 1080p), layered stack stress (config 4, 4K), 10 M-triangle city (config 5, 1080p and 4K, BVH build time).
One JSON line per run.  Usage: python tools/bench_scenes.py [frames]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pltfalcor_b200 import binding, scene  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
RUNS = [("grating_sphere_mnee (configs[2])", scene.grating_sphere_scene, [(1920, 1080)]),
        ("layered_stress_hwss4 (configs[3])", scene.layered_stress_scene, [(3840, 2160)]),
        ("city_10m (configs[4])", lambda aspect: scene.city_scene(10_000_000, aspect=aspect), [(1920, 1080), (3840, 2160)])]
for name, make, sizes in RUNS:
    buffers, cam = make(aspect=16.0 / 9.0)
    p = binding.Pltpt(maxBounces=8)
    t0 = time.perf_counter()
    p.set_scene(buffers)
    torch.cuda.synchronize()
    t_set = time.perf_counter() - t0
    s = p.stats()
    for (w, h) in sizes:
        acc = torch.zeros((h, w, 4), dtype=torch.float64, device="cuda")
        p.render_accumulate(cam, 0, 1, 4, w, h, acc)   # warm-up: 4 frames (two per lane)
        torch.cuda.synchronize()
        p.reset_stats()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        p.render_accumulate(cam, 4, 1, frames, w, h, acc)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / frames
        st = p.stats()
        paths = w * h * frames
        print(json.dumps({"workload": name, "width": w, "height": h, "frames": frames, "frames_in_flight": 2, "triangles": int(s["triangleCount"]), "bvh_nodes": int(s["bvhNodes"]),
                          "bvh_build_ms": round(s["bvhBuildMs"], 2), "set_scene_s": round(t_set, 3), "ms_per_frame": round(ms, 2), "Mpaths_s": round(w * h / ms / 1e3, 1),
                          "mean_vertices": round(st["totalPathVertices"] / paths, 2), "mnee_walks_last_frame": int(st["mneeWalks"]),
                          "bvh_nodes_per_path": round(st["traversalNodesVisited"] / paths, 1), "stack_overflows": int(st["traversalStackOverflows"]),
                          "dropped_connections": int(st["droppedConnections"]), "mean_rgb": float(acc[..., :3].sum() / acc[..., 3].sum())}), flush=True)
        del acc
    p.close()
    torch.cuda.empty_cache()
