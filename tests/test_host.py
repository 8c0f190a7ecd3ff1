"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/pltpt.h declares,
struct layouts match between the header, numpy and ctypes, options follow PLTPT.h:74-108, and the product fails loudly
without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pltpt.h")


@pytest.fixture(scope="module")
def binding():
    from pltfalcor_b200 import binding, build
    build.build()
    return binding


def header_functions():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"PLTPT_API\s+[\w\s\*]+?\b(pltpt_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points(binding):
    assert header_functions() == sorted(binding.EXPORTS)


def test_library_exports_every_declared_symbol(binding):
    lib = binding.lib()
    for name in header_functions():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", binding.LIB], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    assert set(header_functions()) <= exported
    # nothing but the C ABI leaks out of the library
    assert all(s.startswith("pltpt_") for s in exported), sorted(s for s in exported if not s.startswith("pltpt_"))[:5]


def test_library_is_built_for_sm_100a(binding):
    out = subprocess.run(["cuobjdump", "-lelf", binding.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_struct_layouts_match_header(tmp_path):
    from pltfalcor_b200 import layouts as L
    names = ["PltptOptions", "PltptPackedVertex", "PltptGeometryInstance", "PltptMeshDesc", "PltptMaterialBlob", "PltptSpectralProfile", "PltptLightData",
             "PltptEmissiveTriangle", "PltptEmissiveFlux", "PltptCameraData", "PltptBounce", "PltptSceneDesc", "PltptStats"]
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "pltpt.h"\nint main(void){' + "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in names) +
                   'printf("off_invWeightsSum %zu\\n", __builtin_offsetof(PltptSceneDesc, invWeightsSum));'
                   'printf("off_emissionSpectralId %zu\\n", __builtin_offsetof(PltptMaterialBlob, emissionSpectralId));'
                   'printf("off_texBaseColor %zu\\n", __builtin_offsetof(PltptMaterialBlob, texBaseColor));'
                   'printf("off_cameraU %zu\\n", __builtin_offsetof(PltptCameraData, cameraU));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])  # the header is plain C
    sizes = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    sizes = {k: int(v) for k, v in sizes.items()}
    # Falcor layouts: 32 B vertex / instance / mesh, 128 B material blob, 1052 B profile, 64 B emissive triangle, 40 B bounce
    assert sizes["PltptPackedVertex"] == L.packed_vertex_dtype.itemsize == 32
    assert sizes["PltptGeometryInstance"] == L.geometry_instance_dtype.itemsize == 32
    assert sizes["PltptMeshDesc"] == L.mesh_desc_dtype.itemsize == 32
    assert sizes["PltptMaterialBlob"] == L.material_blob_dtype.itemsize == 128
    assert sizes["PltptSpectralProfile"] == L.spectral_profile_dtype.itemsize == 1052
    assert sizes["PltptLightData"] == L.light_data_dtype.itemsize == 208
    assert sizes["PltptEmissiveTriangle"] == L.emissive_triangle_dtype.itemsize == 64
    assert sizes["PltptEmissiveFlux"] == L.emissive_flux_dtype.itemsize == 16
    assert sizes["PltptBounce"] == L.bounce_dtype.itemsize == 40
    assert sizes["PltptCameraData"] == C.sizeof(L.PltptCameraData) == 656
    assert sizes["PltptOptions"] == C.sizeof(L.PltptOptions)
    assert sizes["PltptSceneDesc"] == C.sizeof(L.PltptSceneDesc)
    assert sizes["PltptStats"] == C.sizeof(L.PltptStats)
    assert sizes["off_invWeightsSum"] == L.PltptSceneDesc.invWeightsSum.offset
    assert sizes["off_emissionSpectralId"] == L.material_blob_dtype.fields["emissionSpectralId"][1]
    assert sizes["off_texBaseColor"] == L.material_blob_dtype.fields["tex"][1]
    assert sizes["off_cameraU"] == L.PltptCameraData.cameraU.offset


def test_default_options_match_reference_defaults(binding):  # PLTPT.h:74-108
    from pltfalcor_b200 import layouts as L
    o = L.PltptOptions()
    binding.lib().pltpt_default_options(C.byref(o))
    for k, v in L.OPTION_DEFAULTS.items():
        got = getattr(o, k)
        assert got == pytest.approx(float(v)), k
    assert (o.maxBounces, o.tileSize, o.HWSS, o.MNEEMaxOccluders, o.MNEEMaxIterations) == (32, 512, 4, 2, 60)


def test_unknown_property_is_rejected():
    from pltfalcor_b200 import layouts as L
    with pytest.raises(KeyError, match="Unknown field 'maxBounce'"):  # PLTPT.cpp:173 logs the same text
        L.make_options(maxBounce=3)


def test_no_cpu_fallback(binding):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(binding.PltptError) as e:
        binding.Pltpt()
    assert e.value.code == 6 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may touch oracle/."""
    pat = re.compile(r"(^|\n)\s*(from|import)\s+oracle\b|liboracle|#include\s+[\"<][^\n]*oracle|oracle\.binding|oracle_\w+\(")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pltfalcor_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(txt), f


def test_halton_jitter_matches_gbuffer_pattern():  # HaltonSamplePattern.cpp:37-67
    from pltfalcor_b200 import scene
    # frame 0 -> halton(0)=0 -> 0+0.5-0-0.5 = 0 ; frame 1 -> (0.5, 1/3) -> wrapped to [-0.5, 0.5)
    assert scene.halton_jitter(0, 100, 100) == (0.0, 0.0)
    jx, jy = scene.halton_jitter(1, 100, 100)
    assert jx == pytest.approx(-0.5 / 100) and jy == pytest.approx((1 / 3) / 100, rel=1e-6)
    assert scene.halton_jitter(17, 64, 64) == scene.halton_jitter(1, 64, 64)


def test_alias_table_induces_flux_pdf(cornell):  # EmissivePowerSampler.cpp:86-170: parity is on the induced pdf, not the bytes
    buffers, _ = cornell
    tab = buffers.arrays["triangleAliasTable"].reshape(-1, 2)
    flux = buffers.arrays["emissiveFlux"]["flux"]
    n = len(flux)
    p = np.zeros(n)
    for i in range(n):
        x, y = int(tab[i, 0]), int(tab[i, 1])
        thr = float(np.array([x >> 16], np.uint16).view(np.float16)[0])
        above, below = ((x & 0xFFFF) << 8) | ((y >> 24) & 0xFF), y & 0xFFFFFF
        p[above] += (1 - thr) / n
        p[below] += thr / n
    np.testing.assert_allclose(p, flux / flux.sum(), atol=2e-3)
    assert buffers.scalars["invWeightsSum"] == pytest.approx(1.0 / flux.sum(), rel=1e-5)
