"""The 13 debug views of the pass (DebugViewType.slangh:5-20, pltpt_debug.slangh, pltpt_solve.rt.slang:100-106,318-326,450-475).
CPU: the colour-map fits against golden samples of the reference's tables (tests/golden/colormaps.json, tools/make_colormap_fit.py).
GPU: every view against the oracle on the scenes that exercise it."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from pltfalcor_b200 import layouts as L

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "colormaps.json")
VIEWS = {1: "nans", 2: "inf", 3: "path_length", 4: "hwss_comps", 5: "normals", 6: "albedo", 7: "roughness", 8: "UVs", 9: "coherence_area",
         10: "coherence_anisotropy", 11: "polarization", 12: "mnee_iterations", 13: "mnee_bounces"}


def test_colour_map_fits_match_the_reference_tables(oracle):
    g = json.load(open(GOLDEN))
    for which, name in enumerate(("ccMagma", "ccInferno")):
        args = np.asarray(g[name]["args"], np.float32)
        out = np.zeros((len(args), 3), np.float32)
        oracle.lib().oracle_colour_map(C.c_int(which), C.c_uint32(len(args)), args.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        # four degree-5 segments per channel: max abs error 0.0052 (magma) / 0.0084 (inferno) over the whole table
        assert np.abs(out - np.asarray(g[name]["rgb"])).max() <= 0.01, name
    # out-of-range arguments clamp like the table lookup does
    v = np.asarray([-1.0, 0.0, 1.0, 7.0], np.float32)
    out = np.zeros((4, 3), np.float32)
    oracle.lib().oracle_colour_map(C.c_int(0), C.c_uint32(4), v.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert np.allclose(out[0], out[1]) and np.allclose(out[2], out[3], atol=2e-3)


def test_oracle_debug_views_are_distinct_overlays(oracle):
    from pltfalcor_b200 import scene
    buffers, cam = scene.grating_sphere_scene(n_lat=24, n_lon=48)
    osc = oracle.OracleScene(buffers)
    c = scene.with_jitter(cam, 0, 48, 27)
    base = osc.render_frame(L.make_options(maxBounces=6), c, 0, 48, 27)["color"][..., :3]
    seen = []
    for dv in VIEWS:
        img = osc.render_frame(L.make_options(maxBounces=6, debugView=dv), c, 0, 48, 27)["color"][..., :3]
        assert np.isfinite(img).all(), VIEWS[dv]
        half = osc.render_frame(L.make_options(maxBounces=6, debugView=dv, debugViewIntensity=0.5), c, 0, 48, 27)["color"][..., :3]
        assert np.allclose(half, 0.5 * (base + img), rtol=1e-4, atol=1e-6), VIEWS[dv]  # lerp(color, view, intensity)
        seen.append(img)
    assert np.abs(seen[4] - seen[5]).max() > 0.1  # normals vs albedo
    n = seen[4][seen[4].any(-1)]
    assert np.allclose(np.linalg.norm(n, axis=-1), 1, atol=1e-3)  # the normals view shows unit normals at the first vertex


@pytest.mark.gpu
@pytest.mark.parametrize("dv", sorted(VIEWS))
def test_gpu_debug_view_matches_oracle(dv, torch_cuda, oracle):
    from pltfalcor_b200 import binding, scene
    from tests import parity
    # the grating + sphere scene has delta chains (MNEE views) and polarising surfaces; the Cornell box covers the rest
    buffers, cam = scene.grating_sphere_scene(n_lat=32, n_lon=64) if dv >= 9 else scene.cornell_box()
    p = binding.Pltpt(maxBounces=6, debugView=dv)
    p.set_scene(buffers)
    osc = oracle.OracleScene(buffers)
    w, h, frame = 64, 36, 2
    c = scene.with_jitter(cam, frame, w, h)
    gpu_b = p.sample_stage_host(c, frame, w, h)
    gpu = p.render_frame_host(c, frame, w, h)["color"]
    cpu = osc.render_frame(p.options, c, frame, w, h, want_bounces=True)
    _, same = parity.compare_bounces(gpu_b, cpu["bounces"])
    d = np.abs(gpu[..., :3] - cpu["color"][..., :3]).max(-1)
    assert np.isfinite(gpu).all()
    # identical paths: the overlay is a mean of per-connection colours, each a smooth function of fp32 quantities good to ~1e-4
    assert (d[same] <= 5e-3).mean() >= 0.97, (VIEWS[dv], float(d[same].max()))
    if dv >= 3:
        assert gpu[..., :3].max() > 0, VIEWS[dv]  # the view shows something on this scene
    p.close()


@pytest.mark.gpu
def test_gpu_rejects_unknown_debug_view(torch_cuda):
    from pltfalcor_b200 import binding
    with pytest.raises(binding.PltptError):
        binding.Pltpt(debugView=14)
