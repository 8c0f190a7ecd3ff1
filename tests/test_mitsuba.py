"""Mitsuba-style BSDF descriptions -> PLT materials (pltfalcor_b200/mitsuba.py after MitsubaImporter.cpp:286-527, Tables.h:13-39)."""
import math

import numpy as np
import pytest

from pltfalcor_b200 import layouts as L, mitsuba, scene


def _h(x):
    return float(np.float16(x))


def _f16(m, field):
    return m[field].view(np.float16).astype(np.float32)


def test_named_iors_and_roughness_conversion():
    assert mitsuba.lookup_ior({}, "int_ior", "bk7") == (1.5046, 64.0)
    assert mitsuba.lookup_ior({"int_ior": "Diamond"}, "int_ior", "bk7") == (2.419, 55.3)
    assert mitsuba.lookup_ior({"int_ior": 1.7}, "int_ior", "bk7") == (1.7, math.inf)
    assert mitsuba.lookup_ior({"ext_ior": "no such glass"}, "ext_ior", "air")[0] == 1.000277  # falls back to air (Tables.h:52)
    assert mitsuba.convert_gs_sigma2(1000.0) == 0.0 and mitsuba.convert_gs_sigma2(0.0) == 1.0
    assert mitsuba.convert_gs_sigma2(100.0) == pytest.approx(2 * (1 - 0.1 ** (1 / 32)))


def test_bsdf_descriptions_map_to_the_reference_material_slots():
    b = scene.SceneBuilder()
    # rough dielectric with a named glass: IoR, dispersion from the Abbe number, exterior IoR of air
    m = b.materials[mitsuba.add_bsdf(b, {"type": "roughdielectric", "int_ior": "bk7", "roughness": 0.2, "specular_transmittance": (0.9, 0.8, 0.7)})]
    assert (m["header"][0] & 0xffff) == L.MATERIAL_DIELECTRIC
    assert _f16(m, "IoR")[()] == _h(1.5046)
    spec = _f16(m, "specular")
    assert spec[0] == _h(0.52 * 0.5046 / 64.0) and spec[1] == _h(0.2) and spec[3] == _h(1.000277)
    assert np.array_equal(_f16(m, "transmission"), np.float16([0.9, 0.8, 0.7]).astype(np.float32))
    # smooth "dielectric" still gets the sigma2 default (the importer makes no difference, MitsubaImporter.cpp:334-337); no dispersion for a numeric IoR
    m = b.materials[mitsuba.add_bsdf(b, {"type": "dielectric", "int_ior": 1.33})]
    assert _f16(m, "specular")[1] == _h(mitsuba.convert_gs_sigma2(100.0)) and _f16(m, "specular")[0] == 0.0
    # conductor from the tabulated IOR files
    m = b.materials[mitsuba.add_bsdf(b, {"type": "roughconductor", "material": "Au", "roughness": 0.3, "specular_reflectance": 0.9})]
    assert (m["header"][0] & 0xffff) == L.MATERIAL_CONDUCTOR and m["iorNSpectralId"] != L.NO_SPECTRAL_PROFILE and m["iorKSpectralId"] == m["iorNSpectralId"] + 1
    assert _f16(m, "specular")[0] == 3.0 and _f16(m, "specular")[1] == _h(0.3) and _f16(m, "baseColor")[0] == _h(0.9)
    with pytest.raises(ValueError):
        mitsuba.add_bsdf(b, {"type": "conductor"})
    # coated conductor: thickness default 50 um, coat IoR from int_ior, the nested conductor's roughness and tables
    m = b.materials[mitsuba.add_bsdf(b, {"type": "coated_roughconductor", "int_ior": 1.5, "bsdf": {"type": "roughconductor", "material": "Cu", "roughness": 0.05}})]
    assert (m["header"][0] & 0xffff) == L.MATERIAL_COATED_CONDUCTOR
    assert _f16(m, "data1")[0] == 50.0 and _f16(m, "data1")[3] == 1.5 and _f16(m, "specular")[1] == _h(0.05)
    with pytest.raises(ValueError):
        mitsuba.add_bsdf(b, {"type": "coated_conductor", "bsdf": {"type": "diffuse"}})
    # grating: degrees -> radians, height 1 with `q` as its scale, type + lobes packed into data1.w
    m = b.materials[mitsuba.add_bsdf(b, {"type": "grating", "material": "Au", "pitch": 3.0, "lobes": 5, "q": 0.4, "grating_direction": 90.0, "grating_type": "radial_rectangular"})]
    assert (m["header"][0] & 0xffff) == L.MATERIAL_GRATED_CONDUCTOR
    d1 = _f16(m, "data1")
    assert d1[0] == 3.0 and d1[1] == 0.0 and d1[2] == _h(math.pi / 2) and int(m["data1"][3]) == ((0x11 << 8) | 5)
    assert _f16(m, "specular")[0] == _h(0.4) and _f16(m, "data2")[0] == 1.0
    # thin dielectric (with the importer's swapped birefringence properties), two-sided wrapper, unknown plugin
    m = b.materials[mitsuba.add_bsdf(b, {"type": "thindielectric", "thickness": 0.3, "birefringence": 0.1, "optic_axis": (0.0, 0.6, -0.8)})]
    assert _f16(m, "baseColor")[0] == 1.0 and _f16(m, "specular")[0] == _h(0.1) and _f16(m, "specular")[3] == _h(0.3)
    assert _f16(m, "baseColor")[1] == 0.0 and _f16(m, "baseColor")[2] == _h(-0.6)  # axis flipped to z >= 0
    m = b.materials[mitsuba.add_bsdf(b, {"type": "twosided", "bsdf": {"type": "diffuse"}})]
    assert (m["header"][0] >> 28) & 1 and np.allclose(_f16(m, "baseColor")[:3], 0.5)
    m = b.materials[mitsuba.add_bsdf(b, {"type": "plastic"})]
    assert (m["header"][0] & 0xffff) == L.MATERIAL_DIFFUSE and np.array_equal(_f16(m, "baseColor")[:3], [1, 0, 0])
    with pytest.raises(mitsuba.UnsupportedMitsubaFeature):
        mitsuba.add_bsdf(b, {"type": "diffuse", "reflectance": {"type": "bitmap", "filename": "x.png"}})


def test_described_scene_renders_with_the_oracle(oracle):
    """A scene whose materials all come from Mitsuba-style descriptions goes through the whole harness -> oracle path."""
    b = scene.SceneBuilder()
    lamp = b.add_emission_profile(scene.LAMP, 2.0)
    floor = mitsuba.add_bsdf(b, {"type": "diffuse", "reflectance": (0.6, 0.6, 0.6)})
    gold = mitsuba.add_bsdf(b, {"type": "roughconductor", "material": "Au", "roughness": 0.25})
    glass = mitsuba.add_bsdf(b, {"type": "dielectric", "int_ior": "bk7", "roughness": 0.0})
    light = b.add_diffuse((0, 0, 0), emission_profile=lamp)
    b.add_mesh(scene.grid_plane(0.0, -3.0, 3.0, -3.0, 3.0, 2, 2), floor)
    b.add_mesh(scene.uv_sphere([-0.7, 0.5, 0.0], 0.5, 12, 24), gold)
    b.add_mesh(scene.uv_sphere([0.7, 0.5, 0.0], 0.5, 12, 24), glass)
    b.add_mesh(scene.quad([-0.5, 2.5, -0.5], [0.5, 2.5, -0.5], [0.5, 2.5, 0.5], [-0.5, 2.5, 0.5]), light)
    buffers = b.finalize()
    cam = scene.make_camera([0.0, 1.2, 3.5], [0.0, 0.5, 0.0], focal_length=30.0)
    img = oracle.OracleScene(buffers).render_frame(L.make_options(maxBounces=4), scene.with_jitter(cam, 0, 48, 27), 0, 48, 27)["color"]
    assert np.isfinite(img).all() and img[..., :3].mean() > 0
