"""BASELINE configs[4] evidence run: the ~10 M-triangle procedural city (all ten PLT materials, instanced spheres), BVH build time
and render throughput at 1080p / 4K.  Usage: python tools/bench_city.py [target_triangles] [frames]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pltfalcor_b200 import binding, scene  # noqa: E402

target = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
t0 = time.perf_counter()
buffers, cam = scene.city_scene(target)
t_gen = time.perf_counter() - t0
p = binding.Pltpt(maxBounces=8)
t0 = time.perf_counter()
p.set_scene(buffers)
torch.cuda.synchronize()
t_set = time.perf_counter() - t0
s = p.stats()
out = {"triangles": int(s["triangleCount"]), "bvh_nodes": int(s["bvhNodes"]), "bvh_leaf_triangles": int(s["bvhLeafTriangles"]), "bvh_build_ms": s["bvhBuildMs"],
       "set_scene_s": round(t_set, 3), "scene_generation_s": round(t_gen, 1), "bvh_bytes": int(s["bvhNodes"]) * 80 + int(s["bvhLeafTriangles"]) * 48}
for (w, h) in ((1920, 1080), (3840, 2160)):
    acc = torch.zeros((h, w, 4), dtype=torch.float64, device="cuda")
    p.render_accumulate(cam, 0, 1, 2, w, h, acc)
    torch.cuda.synchronize()
    p.reset_stats()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    p.render_accumulate(cam, 2, 1, frames, w, h, acc)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / frames
    st = p.stats()
    paths = w * h * frames
    out[f"{w}x{h}"] = {"ms_per_frame": round(ms, 2), "Mpaths_s": round(w * h / ms / 1e3, 1), "mean_vertices": round(st["totalPathVertices"] / paths, 2),
                       "bvh_nodes_per_path": round(st["traversalNodesVisited"] / paths, 1), "bvh_triangles_per_path": round(st["traversalTrianglesTested"] / paths, 1),
                       "stack_overflows": int(st["traversalStackOverflows"]), "dropped_connections": int(st["droppedConnections"]),
                       "mean_rgb": float(acc[..., :3].sum() / acc[..., 3].sum())}
    del acc
print(json.dumps(out))
