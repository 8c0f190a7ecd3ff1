"""Per-kernel ms of one frame of a named scene (CUDA events after every launch).  Usage: kernel_times.py scene [w h] [frames] [k=v options...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pltfalcor_b200 import binding, layouts as L, scene  # noqa: E402

if os.environ.get("PLTPT_LIB"):
    binding.LIB = os.environ["PLTPT_LIB"]  # a variant built by tools/variants.py

name = sys.argv[1]
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 4
opts = {}
for kv in sys.argv[5:]:
    k, v = kv.split("=")
    opts[k] = (v == "True") if v in ("True", "False") else (float(v) if "." in v else int(v))
fn = {"cornell": scene.cornell_box, "grating": scene.grating_sphere_scene, "layered": scene.layered_stress_scene, "env": scene.env_scene,
      "city": lambda aspect: scene.city_scene(10_000_000, aspect=aspect), "city1m": lambda aspect: scene.city_scene(1_000_000, aspect=aspect)}[name]
buffers, cam = fn(aspect=w / h)
kw = dict(maxBounces=8)
kw.update(opts)
p = binding.Pltpt(**kw)
p.set_scene(buffers)
acc = torch.zeros((h, w, 4), dtype=torch.float64, device="cuda")
p.render_accumulate(cam, 0, 1, 2, w, h, acc)
torch.cuda.synchronize()
p.set_profiling(True)
p.reset_stats()
p.render_accumulate(cam, 2, 1, frames, w, h, acc)
torch.cuda.synchronize()
s = p.stats()
paths = w * h * frames
print(json.dumps({"scene": name, "options": opts, "triangles": int(s["triangleCount"]), "bvh_nodes": int(s["bvhNodes"]), "bvh_build_ms": round(s["bvhBuildMs"], 2), "bvh_sah_cost": round(s.get("bvhSahCost", 0.0), 2),
                  "kernels_ms_per_frame": {n: round(s["kernelMs"][n] / frames, 3) for n in L.KERNEL_CLASSES if s["kernelLaunchCount"][n]},
                  "mean_vertices": round(s["totalPathVertices"] / paths, 2), "mnee_iterations_last_frame": int(s["mneeIterations"]), "mnee_walks_last_frame": int(s["mneeWalks"]),
                  "bvh_nodes_per_path": round(s["traversalNodesVisited"] / paths, 1), "bvh_triangles_per_path": round(s["traversalTrianglesTested"] / paths, 1),
                  "unwind_steps_per_path": round(s["unwindJobSteps"] / paths, 2)}))
