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
        mitsuba.add_bsdf(b, {"type": "diffuse", "reflectance": {"type": "bitmap", "filename": "x.png"}})   # no image files offline: texels are passed


def _handle(m, slot):
    return int(m["tex"][scene.SceneBuilder.TEX_SLOTS[slot]])


def test_texture_descriptions_go_to_the_importers_slots():
    """buildTexture / lookupTexture (MitsubaImporter.cpp:212-284) and the slots buildBSDF puts them in (:300-372, :480-500)."""
    b = scene.SceneBuilder()
    # checkerboard: 512 x 512 RGBA32F, (x < 256) ^ (y < 256) ? color1 : color0, alpha 1; defaults 0.4 / 0.2
    mid = mitsuba.add_bsdf(b, {"type": "diffuse", "reflectance": {"type": "checkerboard", "color1": (0.9, 0.1, 0.1)}})
    m = b.materials[mid]
    assert _handle(m, "baseColor") == (1 << 29) | 0 and len(b.textures) == 1
    px, fmt = b.textures[0]
    assert fmt == "rgba32f" and px.shape == (512, 512, 4) and (px[..., 3] == 1).all()
    assert np.allclose(px[0, 0, :3], 0.4) and np.allclose(px[0, 300, :3], [0.9, 0.1, 0.1]) and np.allclose(px[300, 0, :3], [0.9, 0.1, 0.1]) and np.allclose(px[300, 300, :3], 0.4)
    assert np.allclose(_f16(m, "baseColor")[:3], 0.5)               # the uniform value stays the parameter's default
    # bitmaps: uint8 is sRGB unless raw, three channels get alpha; float texels are linear
    rgb8 = np.full((4, 8, 3), 128, np.uint8)
    mid = mitsuba.add_bsdf(b, {"type": "roughdielectric", "specular_reflectance": {"type": "bitmap", "texels": rgb8},
                               "specular_transmittance": {"type": "bitmap", "texels": rgb8, "raw": True}})
    m = b.materials[mid]
    assert _handle(m, "baseColor") == (1 << 29) | 1 and _handle(m, "transmission") == (1 << 29) | 2
    assert b.textures[1][1] == "rgba8_srgb" and b.textures[2][1] == "rgba8" and b.textures[1][0].shape == (4, 8, 4) and (b.textures[1][0][..., 3] == 255).all()
    mid = mitsuba.add_bsdf(b, {"type": "conductor", "material": "Cu", "specular_reflectance": {"type": "bitmap", "texels": np.ones((2, 2, 4), np.float32) * 0.5}})
    assert _handle(b.materials[mid], "baseColor") == (1 << 29) | 3 and b.textures[3][1] == "rgba32f"
    # the coated conductor takes over the nested conductor's base-colour texture; the thin dielectric's birefringence map is its Specular slot
    mid = mitsuba.add_bsdf(b, {"type": "coated_conductor", "bsdf": {"type": "conductor", "material": "Au", "specular_reflectance": {"type": "checkerboard"}}})
    assert _handle(b.materials[mid], "baseColor") == (1 << 29) | 4
    mid = mitsuba.add_bsdf(b, {"type": "thindielectric", "birefringence": {"type": "checkerboard", "color0": 0.0, "color1": 0.2}})
    assert _handle(b.materials[mid], "specular") == (1 << 29) | 5
    # normalmap: nested BSDF + map; NormalMapType RGB in the material flags
    mid = mitsuba.add_bsdf(b, {"type": "normalmap", "bsdf": {"type": "roughconductor", "material": "Au", "roughness": 0.1},
                               "normalmap": {"type": "bitmap", "raw": True, "texels": np.tile(np.uint8([128, 128, 255]), (4, 4, 1))}})
    m = b.materials[mid]
    assert (m["header"][0] & 0xffff) == L.MATERIAL_CONDUCTOR and _handle(m, "normalMap") == (1 << 29) | 6 and (int(m["flags"]) >> 1) & 3 == 1
    with pytest.raises(ValueError):
        mitsuba.add_bsdf(b, {"type": "normalmap", "bsdf": {"type": "diffuse"}})
    with pytest.raises(mitsuba.UnsupportedMitsubaFeature):
        mitsuba.add_bsdf(b, {"type": "diffuse", "reflectance": {"type": "scale"}})
    with pytest.raises(ValueError):
        mitsuba.add_bsdf(b, {"type": "diffuse", "reflectance": "red"})


def test_conductor_iors_given_as_curves():
    """eta / k as spectra (MitsubaImporter.cpp:380-387): two profiles through addSpectralProfile, whose computeICdf defaults to true."""
    b = scene.SceneBuilder()
    wl = np.linspace(380.0, 720.0, 18).astype(np.float32)
    mid = mitsuba.add_bsdf(b, {"type": "roughconductor", "roughness": 0.2, "eta": (wl, 0.2 + 0.001 * (wl - 380.0)), "k": (wl, 3.0 + 0.002 * (wl - 380.0))})
    m = b.materials[mid]
    n_id, k_id = int(m["iorNSpectralId"]), int(m["iorKSpectralId"])
    assert (n_id, k_id) == (0, 1) and len(b.profiles) == 2
    ref = scene.gen_spectral_profile(wl, 3.0 + 0.002 * (wl - 380.0), True)
    assert b.profiles[k_id].tobytes() == ref.tobytes() and float(ref["intensity"]) > 1.0   # normalised pdf + its sum, as the importer's call leaves it
    with pytest.raises(mitsuba.UnsupportedMitsubaFeature):
        mitsuba.add_bsdf(b, {"type": "conductor", "eta": (0.2, 0.9, 1.1), "k": (3.9, 2.4, 2.2)})


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


def _textured_described_scene():
    b = scene.SceneBuilder()
    lamp = b.add_emission_profile(scene.LAMP, 2.0)
    floor = mitsuba.add_bsdf(b, {"type": "diffuse", "reflectance": {"type": "checkerboard", "color0": (0.8, 0.8, 0.8), "color1": (0.05, 0.05, 0.05)}})
    bump = np.tile(np.uint8([128, 128, 255]), (8, 8, 1))
    bump[::2, :, 0] = 200
    gold = mitsuba.add_bsdf(b, {"type": "normalmap", "bsdf": {"type": "roughconductor", "material": "Au", "roughness": 0.25}, "map": {"type": "bitmap", "raw": True, "texels": bump}})
    light = b.add_diffuse((0, 0, 0), emission_profile=lamp)
    b.add_mesh(scene.grid_plane(0.0, -3.0, 3.0, -3.0, 3.0, 2, 2), floor)
    b.add_mesh(scene.uv_sphere([0.0, 0.5, 0.0], 0.5, 12, 24), gold)
    b.add_mesh(scene.quad([-0.5, 2.5, -0.5], [0.5, 2.5, -0.5], [0.5, 2.5, 0.5], [-0.5, 2.5, 0.5]), light)
    return b.finalize(), scene.make_camera([0.0, 1.2, 3.5], [0.0, 0.5, 0.0], focal_length=30.0)


def test_textured_described_scene_renders_with_the_oracle(oracle):
    """Checkerboard floor + normal-mapped conductor, both from descriptions, through harness -> texture tables -> oracle."""
    buffers, cam = _textured_described_scene()
    img = oracle.OracleScene(buffers).render_frame(L.make_options(maxBounces=3), scene.with_jitter(cam, 0, 48, 27), 0, 48, 27)["color"]
    assert np.isfinite(img).all() and img[..., :3].mean() > 0
    # the checkerboard is visible: floor pixels of the bottom rows fall into both colours
    row = img[24, :, :3].mean(axis=1)
    assert row.max() > 4 * max(row.min(), 1e-6)


@pytest.mark.gpu
def test_textured_described_scene_on_the_gpu_matches_the_oracle(torch_cuda, oracle):
    """The same described scene through the C ABI: bounce records and one frame against the oracle."""
    from pltfalcor_b200 import binding
    from tests import parity
    buffers, cam = _textured_described_scene()
    p = binding.Pltpt(maxBounces=4)
    p.set_scene(buffers)
    osc = oracle.OracleScene(buffers)
    w, h, frame = 96, 54, 1
    c = scene.with_jitter(cam, frame, w, h)
    gpu_b = p.sample_stage_host(c, frame, w, h)
    gpu = p.render_frame_host(c, frame, w, h)
    cpu = osc.render_frame(p.options, c, frame, w, h, want_bounces=True)
    m, same = parity.compare_bounces(gpu_b, cpu["bounces"])
    mc = parity.compare_color(gpu["color"], cpu["color"], same)
    parity.record("single_frame/mitsuba_textured", dict(mc, identical_paths=m["identical_paths"]))
    assert m["identical_paths"] >= 0.99, m
    assert mc["nonfinite_gpu"] == 0 and mc["frac_pixels_within_1e-3"] >= 0.985 and mc["mean_rel_diff"] <= 0.02, mc
    assert p.stats()["droppedConnections"] == 0
    p.close()
