"""Frames of the REFERENCE ITSELF, run here: renders small frames of the harness scenes with the reference's own shader text (both
ray-generation programs of the PLTPT pass, transpiled textually by oracle/ref_transpile.py and driven by oracle/ref_frames.cpp; only the
intersection search, the texture unit and the environment importance map are supplied) and stores colour, guide buffers and the bounce
records under tests/golden/ref_frames.npz.  tests/test_oracle_ref_frames.py holds the oracle to them wherever the tests run.
Can only be run where /root/reference lies.  Usage: python tools/make_golden_ref.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pltfalcor_b200 import layouts as L, scene  # noqa: E402

W, H, FRAMES = 32, 18, (3, 4)
CASES = {
    "cornell": (scene.cornell_box, dict(maxBounces=6)),
    "cornell_closed": (scene.cornell_box_closed, dict(maxBounces=8)),
    "grating": (lambda: scene.grating_sphere_scene(n_lat=16, n_lon=32), dict(maxBounces=6)),
    "layered": (scene.layered_stress_scene, dict(maxBounces=5)),
    "env": (scene.env_scene, dict(maxBounces=5)),
    "textured": (scene.textured_scene, dict(maxBounces=5)),
    "analytic": (scene.analytic_lights_scene, dict(maxBounces=5)),
    "city": (lambda: scene.city_scene(4000), dict(maxBounces=5)),
    "cornell_nomis_norr": (scene.cornell_box, dict(maxBounces=4, doMIS=False, doRussianRoulette=False, HWSSDoMIS=False)),
    "cornell_nomnee_uniform_tilesg": (scene.cornell_box, dict(maxBounces=5, doMNEE=False, emissiveSampler=0, NEEUsePerTileSG=True, tileSize=8)),
    # the polarising camera (pltpt_solve.rt.slang:111-125) on the scene with the most polarising surfaces; paths without NEE (every light found by
    # BSDF sampling only); no direct lighting and no emitter importance sampling under an environment map; textured scene without alpha masking
    "grating_polarizer": (lambda: _polarised(scene.grating_sphere_scene(n_lat=16, n_lon=32), 0.8, 0.6), dict(maxBounces=6)),
    "env_nonee": (scene.env_scene, dict(maxBounces=6, doNEE=False)),
    "env_nodirect_noimportance": (scene.env_scene, dict(maxBounces=5, useDirectLights=False, doImportanceSampleEmitters=False)),
    "textured_noalpha": (scene.textured_scene, dict(maxBounces=5, alphaMasking=False)),
    # the other hero-wavelength sample counts (HWSS_SAMPLES, hwss_t = vector<float, n>: separate builds of the transpiled pass)
    "cornell_hwss1": (scene.cornell_box, dict(maxBounces=5, HWSS=1)),
    "grating_hwss2": (lambda: scene.grating_sphere_scene(n_lat=16, n_lon=32), dict(maxBounces=6, HWSS=2)),
    "layered_hwss3_nomis": (scene.layered_stress_scene, dict(maxBounces=5, HWSS=3, HWSSDoMIS=False)),
}


def _polarised(scene_and_cam, polarizer, theta):
    buffers, cam = scene_and_cam
    cam = L.PltptCameraData.from_buffer_copy(bytes(cam))
    cam.polarizer, cam.polarizerTheta = polarizer, theta
    return buffers, cam


# configurations checked on the CPU only (oracle against the reference text; added after the last B200 session of the round)
CPU_ONLY_CASES = {
    "grating_mnee1": (lambda: scene.grating_sphere_scene(n_lat=16, n_lon=32), dict(maxBounces=6, MNEEMaxOccluders=1)),
    "grating_mnee3_loose": (lambda: scene.grating_sphere_scene(n_lat=16, n_lon=32), dict(maxBounces=6, MNEEMaxOccluders=3, MNEEMaxIterations=15, MNEESolverThreshold=1e-3)),
}
KEYS = ("color", "albedo", "normal", "hit", "lobe", "flags", "wi", "contribution", "lastpdf")


def _pack(r):
    b = r["bounces"]
    return dict(color=r["color"], albedo=r["albedo"], normal=r["normal"], hit=b["hit"].astype(np.uint32), lobe=b["sampledBsdfLobe"].astype(np.uint32),
                flags=b["flags"].astype(np.uint32), wi=b["encodedWi"].astype(np.float32), contribution=b["pathContribution"].astype(np.float32),
                lastpdf=b["lastNonDeltaPdf"].astype(np.float32))


def setup(name):
    fn, opts = (CASES.get(name) or CPU_ONLY_CASES[name])
    buffers, cam = fn()
    return buffers, cam, L.make_options(**opts)


def render_reference(name, frame):
    from oracle import ref_binding as rb
    buffers, cam, opts = setup(name)
    return _pack(rb.RefScene(buffers, hwss=int(opts.HWSS), occluders=int(opts.MNEEMaxOccluders)).render_frame(opts, scene.with_jitter(cam, frame, W, H), frame, W, H))


def render_oracle(name, frame):
    from oracle import binding as ob
    buffers, cam, opts = setup(name)
    return _pack(ob.OracleScene(buffers).render_frame(opts, scene.with_jitter(cam, frame, W, H), frame, W, H, want_bounces=True))


def load_fixture():
    """{(case, frame): dict(color, albedo, normal, bounces)} from tests/golden/ref_frames.npz, the bounce records reassembled in the C ABI's layout."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_frames.npz"))
    out = {}
    for name in list(CASES) + list(CPU_ONLY_CASES):
        for frame in FRAMES:
            g = {k: z[f"{name}/{frame}/{k}"] for k in KEYS}
            b = np.zeros(g["lobe"].shape, L.bounce_dtype)
            b["hit"], b["encodedWi"], b["sampledBsdfLobe"], b["flags"] = g["hit"], g["wi"], g["lobe"], g["flags"]
            b["pathContribution"], b["lastNonDeltaPdf"] = g["contribution"], g["lastpdf"]
            out[(name, frame)] = dict(color=g["color"], albedo=g["albedo"], normal=g["normal"], bounces=b)
    return out


ACC_MODES = [(0, False), (0, True), (1, False), (2, False), (2, True)]   # (AccumulatePass precision, moving average)
ACC_STEPS = (0, 1, 2, 7, 39)                                             # of 40 accumulated frames, the ones kept in the fixture


def render_reference_accumulate():
    """AccumulatePass by the reference's own text (Accumulate.cs.slang, transpiled) over the seeded frames of tests/parity.accumulate_test_frames."""
    from oracle import ref_binding as rb
    from tests import parity
    out = {}
    for mode, ma in ACC_MODES:
        outs = rb.accumulate_frames(mode, parity.accumulate_test_frames(mode), ma)
        for k in ACC_STEPS:
            out[f"{mode}/{int(ma)}/{k}"] = outs[k]
    return out


DEBUG_VIEWS = {1: "nans", 2: "inf", 3: "path_length", 4: "hwss_comps", 5: "normals", 6: "albedo", 7: "roughness", 8: "UVs", 9: "coherence_area",
               10: "coherence_anisotropy", 11: "polarization", 12: "mnee_iterations", 13: "mnee_bounces"}   # DebugViewType.slangh:5-20
COLOUR_MAPPED = (3, 4, 9, 10, 12)   # views written through ccMagma / ccInferno (pltpt_debug.slangh:24-29): the product evaluates polynomial fits of the tables
DEBUG_CASES = ("cornell", "grating")
DEBUG_FRAME = 3


def debug_options(name, view):
    opts = dict(CASES[name][1])
    opts.update(debugView=view)
    return L.make_options(**opts)


def render_reference_debug_views():
    """The 13 debug views by the reference's own text compiled as the pass compiles it when a view is selected (DO_DEBUG_VIEW = 1, libref_frames_dbg.so)."""
    from oracle import ref_binding as rb
    out = {}
    for name in DEBUG_CASES:
        buffers, cam, _ = setup(name)
        rs = rb.RefScene(buffers, debug=True)
        for view in DEBUG_VIEWS:
            out[f"{name}/{view}"] = rs.render_frame(debug_options(name, view), scene.with_jitter(cam, DEBUG_FRAME, W, H), DEBUG_FRAME, W, H)["color"]
    return out


def emitters_scene():
    """Many emissive triangles under non-trivial instance transforms: a rotated, anisotropically scaled emissive sphere, an emissive quad, a second
    instance of that quad under a MIRRORING transform (negative determinant: Falcor marks such an instance IsWorldFrontFaceCW, SceneBuilder.cpp, and
    computeFaceNormalAndAreaW flips the face normal, Scene.slang:493-494), and a non-emissive mesh between them."""
    import math
    b = scene.SceneBuilder()
    lamp = b.add_emission_profile(scene.LAMP, 1.0)
    dim = b.add_emission_profile(scene.LAMP, 0.25)
    white = b.add_diffuse((0.7, 0.7, 0.7))
    e1 = b.add_diffuse((0.0, 0.0, 0.0), emission_profile=lamp)
    e2 = b.add_diffuse((0.1, 0.1, 0.1), emission_profile=dim)
    c, s_ = math.cos(0.7), math.sin(0.7)
    R = np.array([[c, 0, s_, 0.3], [0, 1, 0, 1.1], [-s_, 0, c, -0.2], [0, 0, 0, 1]], np.float64) @ np.diag([0.9, 1.7, 0.6, 1.0])
    b.add_mesh(scene.uv_sphere([0.1, 0.2, -0.1], 0.35, n_lat=8, n_lon=12), e1, matrix=R.astype(np.float32))
    b.add_mesh(scene.quad([-2, 0, 2], [2, 0, 2], [2, 0, -2], [-2, 0, -2]), white)
    q = b.add_mesh(scene.quad([-0.3, 1.9, -0.3], [0.3, 1.9, -0.3], [0.3, 1.9, 0.3], [-0.3, 1.9, 0.3]), e2)
    mirror = np.diag([-1.0, 1.0, 1.0, 1.0])
    mirror[:3, 3] = (1.2, -0.4, 0.5)
    b.add_instance(q, e2, mirror.astype(np.float32))
    buffers = b.finalize()
    inst = buffers.arrays["geometryInstances"]
    inst["flags"][len(inst) - 1] |= 0x10   # PLTPT_INSTANCE_FLAG_WORLD_FRONT_FACE_CW: the mirrored instance
    cam = scene.make_camera([0.0, 1.0, 3.5], [0.0, 1.0, 0.0], focal_length=35.0, aspect=16.0 / 9.0)
    return buffers, cam


LIGHT_COLLECTION_CASES = {"cornell": scene.cornell_box, "layered": scene.layered_stress_scene, "emitters": emitters_scene}


def render_reference_light_collections():
    """Emissive-triangle and flux tables as the reference's BuildTriangleList / FinalizeIntegration passes build them."""
    from oracle import ref_binding as rb
    out = {}
    for name, fn in LIGHT_COLLECTION_CASES.items():
        buffers, _ = fn()
        tris, flux = rb.RefScene(buffers).build_light_collection()
        out[f"{name}/tris"], out[f"{name}/flux"] = tris.view(np.uint32).reshape(len(tris), 16), flux.view(np.float32).reshape(len(flux), 4)
    return out


SPECTRAL_SCALES = (1.0, 0.02)


def spectral_tables():
    """(key, path under the reference's Source/Tables, wavelengths, values, computeICdf) of the tables the package ships (pltfalcor_b200/data/spd_tables.json)."""
    import json
    t = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pltfalcor_b200", "data", "spd_tables.json")))
    for name, tab in t["emission"].items():
        yield "emission/" + name, f"emission/{name}.spd", tab["wavelengths"], tab["values"], True
    for name, nk in t["ior"].items():
        for part in ("eta", "k"):
            yield f"ior/{name}.{part}", f"ior/{name}.{part}.spd", nk[part]["wavelengths"], nk[part]["values"], False


def render_reference_spectral_profiles():
    """SpectralProfile records by the reference's own C++ (oracle/ref_spectrum.cpp) from the .spd files where they lie under /root/reference/Source/Tables."""
    from oracle import ref_binding as rb
    out = {}
    for key, rel, _w, _v, icdf in spectral_tables():
        for scale in SPECTRAL_SCALES:
            out[f"{key}/{scale}"] = np.frombuffer(rb.profile_from_spd_file(os.path.join("/root/reference/Source/Tables", rel), scale, icdf).tobytes(), np.uint8)
    return out


def _main():
    out = {}
    for name in list(CASES) + list(CPU_ONLY_CASES):
        for frame in FRAMES:
            for k, v in render_reference(name, frame).items():
                out[f"{name}/{frame}/{k}"] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_frames.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_debug_views.npz")
    np.savez_compressed(path, **render_reference_debug_views())
    print(path, os.path.getsize(path), "bytes")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_light_collection.npz")
    np.savez_compressed(path, **render_reference_light_collections())
    print(path, os.path.getsize(path), "bytes")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_spectral_profiles.npz")
    np.savez_compressed(path, **render_reference_spectral_profiles())
    print(path, os.path.getsize(path), "bytes")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_accumulate.npz")
    np.savez_compressed(path, **render_reference_accumulate())
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    _main()
