"""Development-time script (reads the reference, never shipped data from it verbatim): fits piecewise polynomials (four degree-5 segments) to the two colour maps
the PLTPT debug views use (ccMagma, ccInferno: 256-entry tables with linear interpolation, Source/RenderPasses/PLTPT/
ColourCoding.slang:275-540, 542-807) and writes
  * the coefficients as a C initialiser to stdout (pasted into pltfalcor_b200/csrc/plt_debug.cuh and oracle/plt_debug.h), and
  * tests/golden/colormaps.json: the reference rule evaluated at 33 arguments per map (golden vectors for the fit, tolerance in the test).
Usage: python tools/make_colormap_fit.py [/root/reference]"""
import json
import os
import re
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(ref, "Source/RenderPasses/PLTPT/ColourCoding.slang")).read()


def table(name):
    body = src[src.index(f"float3 {name}(float val)"):]
    body = body[:body.index("val *= float(bins);")]
    rows = re.findall(r"\{\s*([0-9.eE+-]+)\s*,\s*([0-9.eE+-]+)\s*,\s*([0-9.eE+-]+)\s*\}", body)
    t = np.array(rows, np.float64)
    assert t.shape == (256, 3), t.shape
    return t


def reference_rule(t, val):  # ColourCoding.slang:536-539
    v = val * 256.0
    lo = int(np.clip(int(v), 0, 255))
    hi = int(np.clip(int(v + 1), 0, 255))
    f = v - np.floor(v)
    return t[lo] + (t[hi] - t[lo]) * f


golden = {}
for name in ("ccMagma", "ccInferno"):
    t = table(name)
    xs = np.linspace(0.0, 1.0, 2049)[:-1]
    ys = np.array([reference_rule(t, x) for x in xs])
    segs, err = [], 0.0
    for s0 in range(4):  # four segments [s0/4, (s0+1)/4], local argument u = 4 x - s0 in [0, 1], degree 5 each
        m = (xs >= s0 / 4) & (xs <= (s0 + 1) / 4 + 1e-9)
        u = 4 * xs[m] - s0
        coef = [np.polyfit(u, ys[m][:, c], 5)[::-1] for c in range(3)]  # c0 + c1 u + ... + c5 u^5
        err = max(err, max(np.abs(np.polyval(coef[c][::-1], u) - ys[m][:, c]).max() for c in range(3)))
        segs.append(coef)
    print(f"// {name}: four degree-5 least-squares segments of the 256-entry table, max abs error {err:.5f}")
    print("{" + ",\n ".join("{" + ", ".join("{" + ", ".join(f"{v:.7e}f" for v in seg[c]) + "}" for c in range(3)) + "}" for seg in segs) + "},")
    args = np.linspace(0.0, 0.999, 33)
    golden[name] = {"args": [float(a) for a in args], "rgb": [[float(x) for x in reference_rule(t, a)] for a in args]}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "colormaps.json")
with open(out, "w") as f:
    json.dump(golden, f)
