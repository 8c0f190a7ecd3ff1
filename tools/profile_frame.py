"""Renders a few frames of the bench workload (BASELINE configs[1]) for profiling under ncu.
Usage: python tools/profile_frame.py [frames] [scene] [width height]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pltfalcor_b200 import binding, scene  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
name = sys.argv[2] if len(sys.argv) > 2 else "cornell"
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
fn = {"cornell": scene.cornell_box, "grating": scene.grating_sphere_scene, "layered": scene.layered_stress_scene}[name]
buffers, cam = fn(aspect=W / H)
p = binding.Pltpt(maxBounces=8)
p.set_scene(buffers)
acc = torch.zeros((H, W, 4), dtype=torch.float64, device="cuda")
p.render_accumulate(cam, 0, 1, frames, W, H, acc)
torch.cuda.synchronize()
s = p.stats()
print({k: s[k] for k in ("framesRendered", "sampleStageMs", "solveStageMs", "totalPathVertices", "totalBsdfEvals", "droppedConnections", "traversalStackOverflows")})
print("mean", float(acc[..., :3].mean()) / frames)
