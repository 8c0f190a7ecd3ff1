"""The oracle held to FRAMES OF THE REFERENCE ITSELF: tests/golden/ref_frames.npz was rendered by the reference's own shader
text -- both ray-generation programs of the PLTPT pass with everything they import, transpiled by oracle/ref_transpile.py and driven tile by
tile by oracle/ref_frames.cpp; only the intersection search, the texture unit and the environment importance map are supplied -- on ten
configurations of the harness scenes (tools/make_golden_ref.py, run where /root/reference lies).  The fixture travels; this test runs
wherever the CPU suite runs.  The oracle must trace the same paths (every hit word, lobe and flag word of every bounce record) and reproduce
the pixels: measured, all ten cases bit-equal except 2 of 576 pixels of the layered scene (9e-7 relative)."""
import importlib.util
import os

import numpy as np
import pytest

from tests import parity

_spec = importlib.util.spec_from_file_location("make_golden_ref", os.path.join(os.path.dirname(__file__), "..", "tools", "make_golden_ref.py"))
gold = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gold)


@pytest.fixture(scope="module")
def fixture():
    return gold.load_fixture()


@pytest.mark.parametrize("name", list(gold.CASES) + list(gold.CPU_ONLY_CASES))
def test_oracle_reproduces_the_reference_frames(fixture, name):
    for frame in gold.FRAMES:
        ref = fixture[(name, frame)]
        got = gold.render_oracle(name, frame)
        b = np.zeros(got["lobe"].shape, ref["bounces"].dtype)
        b["hit"], b["encodedWi"], b["sampledBsdfLobe"], b["flags"] = got["hit"], got["wi"], got["lobe"], got["flags"]
        b["pathContribution"], b["lastNonDeltaPdf"] = got["contribution"], got["lastpdf"]
        m, same = parity.compare_bounces(b, ref["bounces"])
        assert m["identical_paths"] == 1.0, (name, frame, m)
        assert m["max_bary_abs"] == 0.0 and m["max_wi_abs"] == 0.0 and m["max_contribution_rel"] <= 1e-6 and m["max_lastpdf_rel"] <= 1e-6, (name, frame, m)
        assert np.array_equal(got["albedo"], ref["albedo"]) and np.array_equal(got["normal"], ref["normal"]), (name, frame)
        identical = (got["color"].view(np.uint32) == ref["color"].view(np.uint32)).all(-1)
        assert identical.mean() >= 0.99, (name, frame, float(identical.mean()))
        rel = np.abs(got["color"] - ref["color"]) / np.maximum(np.abs(ref["color"]), 1e-6)
        assert rel.max() <= 1e-5, (name, frame, float(rel.max()))
        assert (ref["color"][..., :3].sum(-1) > 0).mean() >= 0.1, (name, frame)   # the fixture is not a black frame (12 % lit without direct lighting, 64-98 % otherwise)


def test_fixture_is_regenerated_identically_where_the_reference_lies():
    """Where /root/reference exists: the committed fixture is what the reference's text renders now (the generator is deterministic)."""
    from oracle import ref_binding
    if not os.path.isdir("/root/reference"):
        pytest.skip("the reference's sources are not on this machine")
    fx = gold.load_fixture()
    for name in ("cornell", "layered", "env"):
        got = gold.render_reference(name, gold.FRAMES[0])
        ref = fx[(name, gold.FRAMES[0])]
        assert np.array_equal(got["color"], ref["color"]) and np.array_equal(got["hit"], ref["bounces"]["hit"]), name


@pytest.mark.parametrize("mode,moving_average", gold.ACC_MODES)
def test_accumulate_restatement_reproduces_the_reference_pass(mode, moving_average):
    """tests/golden/ref_accumulate.npz: the AccumulatePass of the reference's own text (Accumulate.cs.slang accumulateSingle / accumulateSingleCompensated /
    accumulateDouble, transpiled; tools/make_golden_ref.py).  The numpy restatement the GPU test holds pltpt_accumulate_frame to bit for bit
    (tests/parity.accumulate_pass_restatement) equals it to the last bit in the two single-precision modes.  In the Double mode the weight
    `double curWeight = 1.0 / (gAccumCount + 1)` depends on how the shader compiler types the unsuffixed literal: the transpile follows HLSL's rule (fp32
    literal, fp32 division, widened), restatement and CUDA path divide in fp64 -- at most one ulp of the fp32 output, which is the bound asserted here."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_accumulate.npz"))
    outs = parity.accumulate_pass_restatement(mode, parity.accumulate_test_frames(mode), moving_average)
    for k in gold.ACC_STEPS:
        ref = z[f"{mode}/{int(moving_average)}/{k}"]
        ulp = np.abs(outs[k].view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64)).max()
        assert ulp <= (1 if mode == 2 else 0), (mode, moving_average, k, int(ulp))


@pytest.mark.parametrize("name", gold.DEBUG_CASES)
def test_oracle_debug_views_reproduce_the_reference(oracle, name):
    """tests/golden/ref_debug_views.npz: the pass's 13 debug views rendered by the reference's own text in its DO_DEBUG_VIEW = 1 variant.  Views that
    write values directly (NaN / inf markers, normals, albedo, roughness, UVs, polarisation, MNEE bounce classes) are reproduced bit for bit; views written
    through ccMagma / ccInferno carry the error of the polynomial fits that stand in for the reference's 256-entry tables (<= 0.0084 per channel,
    tests/test_debug_views.py).  This comparison found the one behavioural difference the fits had: a non-finite argument makes the table rule return NaN
    (frac(inf) = NaN), which writeDebugOutput drops -- now reproduced by oracle and CUDA path."""
    from pltfalcor_b200 import scene
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_debug_views.npz"))
    buffers, cam, _ = gold.setup(name)
    osc = oracle.OracleScene(buffers)
    c = scene.with_jitter(cam, gold.DEBUG_FRAME, gold.W, gold.H)
    shown = 0
    for view, vname in gold.DEBUG_VIEWS.items():
        ref = z[f"{name}/{view}"]
        got = osc.render_frame(gold.debug_options(name, view), c, gold.DEBUG_FRAME, gold.W, gold.H)["color"]
        if view in gold.COLOUR_MAPPED:
            assert np.abs(got - ref).max() <= 1e-2, (name, vname, float(np.abs(got - ref).max()))
        else:
            assert np.array_equal(got, ref), (name, vname, float(np.abs(got - ref).max()))
        shown += int(ref[..., :3].max() > 0)
    assert shown >= 10, (name, shown)   # the NaN / inf markers stay black on a healthy render, MNEE bounce classes on a scene without walks


@pytest.mark.parametrize("name", list(gold.LIGHT_COLLECTION_CASES))
def test_harness_light_collection_matches_the_reference_passes(name):
    """tests/golden/ref_light_collection.npz: emissive-triangle and flux tables as the reference's own BuildTriangleList / FinalizeIntegration passes build them
    (transpiled; dispatched per mesh light as LightCollection.cpp does).  The scene harness builds the tables it hands to pltpt_set_scene in numpy (fp64
    transform, rounded): same triangles, same uv bits, areas / flux within fp32 rounding, normals within one snorm step -- except that the harness knows nothing
    of mirrored instances, whose face normal the reference flips back (IsWorldFrontFaceCW, Scene.slang:493-494): the fixture's `emitters` case holds one, and
    pltpt_build_light_collection is held to it on the GPU (tests/test_gpu_ref_frames.py)."""
    from pltfalcor_b200 import layouts as L
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_light_collection.npz"))
    buffers, _ = gold.LIGHT_COLLECTION_CASES[name]()
    ref_t = z[f"{name}/tris"].view(L.emissive_triangle_dtype).reshape(-1)
    ref_f = z[f"{name}/flux"].view(L.emissive_flux_dtype).reshape(-1)
    a, f = buffers.arrays["emissiveTriangles"], buffers.arrays["emissiveFlux"]
    assert len(a) == len(ref_t) > 0
    assert np.allclose(a["posAndTexCoords"][..., :3], ref_t["posAndTexCoords"][..., :3], rtol=1e-6, atol=1e-6)
    assert np.array_equal(a["posAndTexCoords"][..., 3].view(np.uint32), ref_t["posAndTexCoords"][..., 3].view(np.uint32))
    assert np.array_equal(a["materialID"], ref_t["materialID"])
    assert np.allclose(a["area"].view(np.float32), ref_t["area"].view(np.float32), rtol=1e-5)
    flagged = (buffers.arrays["geometryInstances"]["flags"] & 0x10) != 0
    keep = np.ones(len(a), bool)
    for i in np.nonzero(flagged)[0]:   # triangles of mirrored instances: the harness does not flip
        first = int(buffers.arrays["perMeshInstanceOffset"][i])
        keep[first:first + 2] = False
    assert parity.snorm2x16_steps(a["normal"][keep], ref_t["normal"][keep]) <= 1
    if name == "emitters":
        assert (~keep).sum() == 2 and parity.snorm2x16_steps(a["normal"][~keep], ref_t["normal"][~keep]) > 30000   # the flip is in the fixture
        assert sorted(set(ref_t["lightIdx"].tolist())) == [0, 1, 2]
    assert np.allclose(f["flux"], ref_f["flux"], rtol=1e-5) and np.allclose(f["averageRadiance"], ref_f["averageRadiance"], rtol=1e-6)
