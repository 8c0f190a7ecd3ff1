"""Pins the oracle to the reference's own shader text (SURVEY §8c "oracle/_ref").

oracle/ref_transpile.py rewrites the reference's PLT Slang modules -- MathHelpers, SpectrumUtils, Fresnel, Mueller, PLT (the generalized
ray), Interference, FractalSurface, GaussianSurface, Birefringence, DiffractionGrating, Scene.evalSpectrum / sampleSpectrum, the
differential reflect / refract helpers, all ten PLT*BSDF modules, the analytic light samplers (LightData, LightHelpers), the pass's own light sampling
(LightSampling.slangh with computeRayOrigin / computeNewRayOrigin, over the emissive-triangle samplers, the light collection, the environment map and its
importance sampler), the ten PLT material loaders (PLT*Material.slang over MaterialData / BasicMaterialData / TextureHandle: both sides unpack
the same 128-byte records, with Uniform and with Texture-mode handles), the shading data at a hit (MaterialFactory.slang prepareShadingData composed
from ShadingUtils.slang: tangent-space orthonormalisation, RGB / RG normal maps, opacity, double-sided flip; Scene.slang computeVertexData), primary rays (Camera.slang over the same 656-byte CameraData block), Russian roulette,
the manifold solver's geometry and Newton step (differentialShadingGeometry, manifoldOneStep), the WHOLE closest-hit and miss programs of the sample
stage (pltpt_sample.rt.slang over Path.slang and appendBounce: path book-keeping, bounce records, guide outputs), `measure` of the solve shader, the packed-record decoders (snorm /
octahedral normals, PackedEmissiveTriangle, PackedStaticVertexData) and finally BOTH RAY-GENERATION PROGRAMS WHOLE (pltpt_sample.rt.slang main; pltpt_solve.rt.slang main with NEE,
evalNEE, evalEmissive, onMiss, forwardUnwind; Manifolds.slangh traceMNEEOccluders / manifoldNEESample / manifoldForwardUnwind; ShadowRays.slangh's ray programs) rendering
whole frames of random scenes next to the oracle's samplePixel / solvePixel -- only the intersection search behind TraceRay is supplied (DXR hardware in the reference) --, and
the environment importance map's setup pass (EnvMapSamplerSetup.cs.slang, all 512 x 512 texels) --
textually into C++ (where /root/reference lies; nothing of it is
committed) and oracle/ref_compare.cpp runs them next to the oracle's hand-written restatement on >= 10^5 seeded random inputs per
function.  The oracle must reproduce the reference text bit for bit, except where its header documents a deliberate definition:
  * delta-lobe energies alpha = exp(-x) are the correctly rounded fp32 value (<= 1 ulp from expf; `1 - alpha` then amplifies that ulp
    in the scatter lobes and their pdfs),
  * the Airy phasors of the multilayer stack are the correctly rounded fp32 cos / sin (an interference sum that cancels to ~1e-4).
"""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_compare")
N = 100_000

# largest allowed distance in units in the last place; everything not listed must be bit-exact
ULP = {
    "FractalSurface.alpha": 1, "GaussianSurface.alpha": 1, "DiffractionGrating.setup_alpha": 1,
    # (1 - alpha) and products of it: the single ulp of alpha, amplified by the cancellation
    "PLTConductorBSDF.sample": 4, "PLTConductorBSDF.eval": 64, "PLTConductorBSDF.pdf": 1024,
    "PLTOpaqueDielectricBSDF.sample": 4, "PLTOpaqueDielectricBSDF.eval": 64, "PLTOpaqueDielectricBSDF.pdf": 1024,
    "PLTDielectricBSDF.eval": 64, "PLTDielectricBSDF.pdf": 1024, "PLTDielectricBSDF.sample": 4,
    "PLTCoatedConductorBSDF.sample": 4, "PLTCoatedConductorBSDF.eval": 64, "PLTCoatedConductorBSDF.pdf": 1024,
    "PLTCoatedOpaqueDielectricBSDF.sample": 4, "PLTCoatedOpaqueDielectricBSDF.eval": 64, "PLTCoatedOpaqueDielectricBSDF.pdf": 1024,
    "PLTDiffractionGratedConductorBSDF.sample": 16, "PLTDiffractionGratedConductorBSDF.eval": 1024, "PLTDiffractionGratedConductorBSDF.pdf": 1024,
    "PLTDiffractionGratedOpaqueDielectricBSDF.sample": 16, "PLTDiffractionGratedOpaqueDielectricBSDF.eval": 1024, "PLTDiffractionGratedOpaqueDielectricBSDF.pdf": 1024,
}
# the material loaders are compared through the BSDFs they set up: the same bounds as the BSDF's own sample / eval / pdf
for _m, _b in (("Conductor", 1024), ("Dielectric", 1024), ("OpaqueDielectric", 1024), ("CoatedConductor", 1024), ("CoatedOpaqueDielectric", 1024),
               ("DiffractionGratedConductor", 1024), ("DiffractionGratedOpaqueDielectric", 1024)):
    ULP[f"PLT{_m}Material.setupMaterialInstance"] = _b
ULP["PLTMaterials.setupMaterialInstance_textured"] = 1024
# the closest-hit / miss programs multiply the sampled BSDF pdf (see the BSDFs' sample() bounds) into the path state and the bounce record
ULP["pltpt_sample.scatterTriangleMeshClosestHit"] = 16
ULP["pltpt_sample.scatterMiss"] = 16
# whole frames: the bounce records carry the sampled pdf's ulp; the pixel colour is a sum of beams evaluated through (1 - alpha) lobes (the eval bounds above)
ULP["pltpt_sample.main_bounce_records"] = 16
ULP["pltpt_solve.main_color"] = 1024
# relative bound instead of an ulp bound: the multilayer stack's interference sum (correctly rounded phasors in the oracle)
REL = {"PLTMultiLayeredStackBSDF.eval": 5e-3, "PLTMultiLayeredStackMaterial.setupMaterialInstance": 5e-3}

EXPECTED = ["MathHelpers.complex", "MathHelpers.boxmueller_truncated", "SpectrumUtils.rgbToSpectrum", "SpectrumUtils.spectrumToRgb", "Fresnel.fresnel_dielectric", "Fresnel.fresnel_conductor",
            "Mueller.MuellerT", "Mueller.MuellerFresnel", "Mueller.MuellerFresnelRConductor", "Mueller.MuellerFresnelDielectric", "PLT.sourceBeam", "PLT.propagate",
            "PLT.constructLocalFrame_rotateFrame", "PLT.coherence", "PLT.transformCoherence", "Interference.interfere_phaseShift", "FractalSurface.eval_pdf_sample",
            "GaussianSurface.eval_pdf_sample", "Birefringence.fresnel_iso_aniso", "Birefringence.fresnel_aniso_iso", "Birefringence.vectors_in_slab", "Scene.evalSpectrum",
            "Scene.sampleSpectrum", "DifferentialShadingGeometry.reflect_refractDerivatives", "DiffractionGrating.lobeIntensity_diffract", "DiffractionGrating.closestLobeHint",
            "DiffractionGrating.sampleLobe"] + [f"PLT{m}BSDF.{f}" for m in ("Diffuse", "Conductor", "Dielectric", "OpaqueDielectric", "ThinDielectric", "CoatedConductor", "CoatedOpaqueDielectric",
                                                                          "DiffractionGratedConductor", "DiffractionGratedOpaqueDielectric", "MultiLayeredStack") for f in ("eval", "sample", "pdf", "lambdaPdf")] + \
           [f"LightHelpers.sample{t}" for t in ("Point", "Directional", "Distant", "Rect", "Disc", "Sphere")] + \
           ["GeometryHelpers.computeRayOrigin", "ShadowRays.computeNewRayOrigin", "LightSampling.selectLightType", "LightSampling.generateLightSample_analytic",
            "pltpt_solve.measure", "pltpt_solve.MISWeight", "PackedFormats.decodeNormal2x16", "LightCollectionShared.PackedEmissiveTriangle_unpack",
            "SceneTypes.PackedStaticVertexData_unpack", "EmissiveUniformSampler.sampleLight", "EmissiveUniformSampler.evalPdf", "EmissivePowerSampler.sampleLight",
            "EmissivePowerSampler.evalPdf", "LightSampling.generateLightSample_emissive", "EnvMap.eval_toLocal_toWorld", "EnvMapSampler.sample", "EnvMapSampler.evalPdf",
            "LightSampling.generateLightSample_env"] + \
           [f"PLT{m}Material.setupMaterialInstance" for m in ("Diffuse", "Conductor", "Dielectric", "OpaqueDielectric", "ThinDielectric", "CoatedConductor", "CoatedOpaqueDielectric",
                                                               "DiffractionGratedConductor", "DiffractionGratedOpaqueDielectric", "MultiLayeredStack")] + \
           ["MaterialFactory.prepareShadingData", "PLTMaterials.setupMaterialInstance_textured", "Scene.computeVertexData", "Camera.computeRayPinhole", "pltpt.RR",
            "Manifolds.manifoldOneStep", "DifferentialShadingGeometry.differentialShadingGeometry", "pltpt_sample.scatterTriangleMeshClosestHit", "pltpt_sample.scatterMiss",
            "pltpt_sample.main_bounce_records", "pltpt_sample.main_guides", "pltpt_solve.main_color", "EnvMapSamplerSetup.main"]


@pytest.fixture(scope="module")
def report():
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])  # no-op when up to date
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/ref_compare is not built and /root/reference is absent (the comparison can only be generated where the reference lies)")
    out = subprocess.run([BIN, str(N)], capture_output=True, text=True, check=True, timeout=600).stdout
    rep = json.loads(out)
    with open(os.path.join(ROOT, "oracle", "_ref", "report.json"), "w") as f:
        json.dump(rep, f, indent=1)
    return rep


def test_every_module_is_compared(report):
    missing = [k for k in EXPECTED if k not in report]
    assert not missing, missing
    thin = {k: v["n"] for k, v in report.items() if v["n"] < 10_000}
    assert not thin, thin
    assert sum(v["n"] for v in report.values()) >= 70_000_000


def test_oracle_reproduces_the_reference_text(report):
    bad = {}
    for name, v in report.items():
        if v["mismatches"]:
            bad[name] = v           # a NaN / inf / integer / boolean result differs
        elif name in REL:
            if v["max_rel"] > REL[name]:
                bad[name] = v
        elif v["max_ulp"] > ULP.get(name, 0):
            bad[name] = v
    assert not bad, bad
    # whole frames rendered by the reference's two ray-generation programs and by the oracle: the same paths (integer fields of every bounce record
    # equal -- a differing hit, lobe or flag word is a "mismatch" above), the same guide images, and all but a handful of pixels equal to the last bit
    frames = report["pltpt_solve.main_color"]
    assert frames["units"] >= 40_000 and frames["identical_units"] >= 0.999 * frames["units"], frames
    assert report["pltpt_sample.main_guides"]["max_ulp"] == 0
    exact = sum(1 for v in report.values() if v["max_ulp"] == 0 and not v["mismatches"])
    assert exact >= 99, exact   # of 127 compared functions / entry points
