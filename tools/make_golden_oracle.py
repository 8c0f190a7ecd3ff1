"""Pins the oracle against itself: renders small frames of the harness scenes with the CPU oracle and stores colour, guide buffers and
the bounce-record digest under tests/golden/oracle_frames.npz.  Frames of the reference itself are tests/golden/ref_frames.npz (tools/make_golden_ref.py);
this fixture is the oracle's own output: it makes an accidental change of the oracle (the checker of every GPU parity test) visible.  Usage: python tools/make_golden_oracle.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import binding as ob  # noqa: E402
from pltfalcor_b200 import layouts as L, scene  # noqa: E402

W, H, FRAME = 32, 18, 3
CASES = {
    "cornell": (scene.cornell_box, dict(maxBounces=6)),
    "grating": (lambda: scene.grating_sphere_scene(n_lat=16, n_lon=32), dict(maxBounces=6)),
    "layered": (scene.layered_stress_scene, dict(maxBounces=5)),
    "env": (scene.env_scene, dict(maxBounces=5)),
    "cornell_nomis_hwss1": (scene.cornell_box, dict(maxBounces=4, doMIS=False, HWSS=1)),
}


def render(name):
    fn, opts = CASES[name]
    buffers, cam = fn()
    osc = ob.OracleScene(buffers)
    r = osc.render_frame(L.make_options(**opts), scene.with_jitter(cam, FRAME, W, H), FRAME, W, H, want_bounces=True)
    b = r["bounces"]
    return dict(color=r["color"], albedo=r["albedo"], normal=r["normal"], hit=b["hit"].astype(np.uint32), lobe=b["sampledBsdfLobe"].astype(np.uint32), flags=b["flags"].astype(np.uint32))


if __name__ == "__main__":
    out = {}
    for name in CASES:
        for k, v in render(name).items():
            out[f"{name}/{k}"] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "oracle_frames.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")
