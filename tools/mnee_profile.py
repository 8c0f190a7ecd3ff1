"""Per-round profile of k_mnee_walk (live walks, advance / trace phase durations) from a library built with
-DPLT_MNEE_PHASE_TIMERS=1 (tools/variants.py build mneeprof).  Usage: PLTPT_LIB=<variant lib> python tools/mnee_profile.py scene [w h] [frames]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pltfalcor_b200 import binding, scene  # noqa: E402

binding.LIB = os.environ.get("PLTPT_LIB", os.path.join(os.path.dirname(binding.LIB), "build", "variants", "mneeprof", "libpltpt.so"))
name = sys.argv[1]
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 2
fn = {"grating": scene.grating_sphere_scene, "layered": scene.layered_stress_scene, "env": scene.env_scene,
      "city": lambda aspect: scene.city_scene(10_000_000, aspect=aspect)}[name]
buffers, cam = fn(aspect=w / h)
p = binding.Pltpt(maxBounces=8)
p.set_scene(buffers)
p.set_frame_overlap(1)
acc = torch.zeros((h, w, 4), dtype=torch.float64, device="cuda")
p.render_accumulate(cam, 0, 1, 1, w, h, acc)
torch.cuda.synchronize()
prof = np.zeros(4 * 128, np.uint64)
binding.lib().pltpt_debug_mnee_profile(prof.ctypes.data_as(C.c_void_p))
before = prof.copy()
p.render_accumulate(cam, 1, 1, frames, w, h, acc)
torch.cuda.synchronize()
binding.lib().pltpt_debug_mnee_profile(prof.ctypes.data_as(C.c_void_p))
d = (prof - before).reshape(128, 4).astype(np.float64) / frames
rounds = [{"round": r, "live_walks": int(d[r, 0]), "advance_us": round(d[r, 1] / 1e3, 1), "trace_us": round(d[r, 2] / 1e3, 1), "rays": int(d[r, 3])} for r in range(128) if d[r, 0] > 0]
print(json.dumps({"scene": name, "width": w, "height": h, "frames": frames, "advance_ms": round(d[:, 1].sum() / 1e6, 3), "trace_ms": round(d[:, 2].sum() / 1e6, 3),
                  "walks": int(d[0, 0]), "trial_rays": int(d[:, 3].sum()), "rounds": rounds}))
