"""Pins the oracle's leaves against every known-answer procedure the reference's own unit tests hold for this path
(SURVEY §8c): PRNG (Tests/Sampling/PseudorandomTests.cpp:40-185), computeRayOrigin (Tests/Utils/GeometryHelpersTests.cpp:217-251),
SampledSpectrum (Tests/Utils/Color/SampledSpectrumTests.cpp:33-57), XYZ fit (Tests/Utils/Color/SpectrumUtilsTests.cpp:40-80)."""
import ctypes as C
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
M32, M64 = 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF


def _rotl(x, k):
    return ((x << k) | (x >> (32 - k))) & M32


def xoshiro128ss(state, n):  # public reference algorithm (xoshiro.di.unimi.it/xoshiro128starstar.c)
    s = list(state)
    out = []
    for _ in range(n):
        out.append((_rotl((s[0] * 5) & M32, 7) * 9) & M32)
        t = (s[1] << 9) & M32
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]
        s[2] ^= t
        s[3] = _rotl(s[3], 11)
    return out


def splitmix64(x, n):  # public reference algorithm (xoshiro.di.unimi.it/splitmix64.c)
    out = []
    for _ in range(n):
        x = (x + 0x9E3779B97F4A7C15) & M64
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        out.append(z ^ (z >> 31))
    return out


def test_splitmix64_published_vectors(oracle):
    # first outputs for seed 0 as published with the algorithm
    out = np.zeros(3, np.uint64)
    oracle.lib().oracle_splitmix64(C.c_uint64(0), C.c_uint32(3), out.ctypes.data_as(C.c_void_p))
    assert [int(v) for v in out] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    assert splitmix64(0, 3) == [int(v) for v in out]


def test_splitmix64_procedure(oracle):  # PseudorandomTests.cpp:125-185 — many seeds x 64 draws
    rng = np.random.default_rng(1)
    for seed in rng.integers(0, 2**63, 64, dtype=np.uint64):
        out = np.zeros(64, np.uint64)
        oracle.lib().oracle_splitmix64(C.c_uint64(int(seed)), C.c_uint32(64), out.ctypes.data_as(C.c_void_p))
        assert [int(v) for v in out] == splitmix64(int(seed), 64)


def test_xoshiro_procedure(oracle):  # PseudorandomTests.cpp:40-97 — 256 seeds x 64 draws, bit exact
    rng = np.random.default_rng(2)
    for _ in range(256):
        seed = rng.integers(0, 2**32, 4, dtype=np.uint32)
        out = np.zeros(64, np.uint32)
        oracle.lib().oracle_xoshiro(seed.ctypes.data_as(C.c_void_p), C.c_uint32(64), out.ctypes.data_as(C.c_void_p))
        assert [int(v) for v in out] == xoshiro128ss([int(v) for v in seed], 64)


def _interleave(x, y):  # BitTricks.slang:45-61
    def spread(v):
        v &= 0xFFFF
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    return spread(x) | (spread(y) << 1)


def test_sample_generator_seeding(oracle):  # UniformSampleGenerator.slang:50-67
    rng = np.random.default_rng(3)
    for _ in range(64):
        px, py, frame = int(rng.integers(0, 4096)), int(rng.integers(0, 4096)), int(rng.integers(0, 2**31))
        a, b = splitmix64((frame << 32) | _interleave(px, py), 2)
        expect = xoshiro128ss([a & M32, a >> 32, b & M32, b >> 32], 16)
        out = np.zeros(16, np.uint32)
        oracle.lib().oracle_prng(C.c_uint32(px), C.c_uint32(py), C.c_uint32(frame), C.c_uint32(16), out.ctypes.data_as(C.c_void_p))
        assert [int(v) for v in out] == expect
        f = np.zeros(16, np.float32)
        oracle.lib().oracle_sample_next_1d(C.c_uint32(px), C.c_uint32(py), C.c_uint32(frame), C.c_uint32(16), f.ctypes.data_as(C.c_void_p))
        assert np.array_equal(f, (np.array(expect, np.uint32) >> 8).astype(np.float32) * np.float32(2.0**-24))  # SampleGeneratorInterface.slang:48-75


def offset_ray(p, n):
    """'A Fast and Robust Method for Avoiding Self-Intersection' (Ray Tracing Gems ch. 6) — the function the reference test
    compares computeRayOrigin against (GeometryHelpersTests.cpp:40-60), restated in numpy fp32/int32."""
    origin, float_scale, int_scale = np.float32(1 / 32.0), np.float32(1 / 65536.0), np.float32(256.0)
    of_i = (int_scale * n).astype(np.int32)
    p_i = (p.view(np.int32) + np.where(p < 0, -of_i, of_i)).view(np.float32)
    return np.where(np.abs(p) < origin, p + float_scale * n, p_i).astype(np.float32)


def test_compute_ray_origin_bit_exact(oracle):
    # computeRayOrigin is offset_ray with normal*3-scaled constants folded in... the reference test feeds the SAME
    # inputs to both and expects equality, which holds because GeometryHelpers.slang uses origin 1/16, fScale 3/65536,
    # iScale 3*256 == offset_ray evaluated on (p, 3n) with origin 1/16.
    n_tests = 1 << 14
    rng = np.random.default_rng(4)
    scale = (10.0 ** (np.arange(n_tests) / n_tests * 60.0 - 30.0)).astype(np.float32)
    pos = (rng.uniform(-1, 1, (n_tests, 3)).astype(np.float32) * scale[:, None]).astype(np.float32)
    nrm = rng.uniform(-1, 1, (n_tests, 3)).astype(np.float32)
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    out = np.zeros_like(pos)
    oracle.lib().oracle_compute_ray_origin(C.c_uint32(n_tests), pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    origin, f_scale, i_scale = np.float32(1 / 16.0), np.float32(3 / 65536.0), np.float32(3 * 256.0)
    of_i = (i_scale * nrm).astype(np.int32)
    p_i = (pos.view(np.int32) + np.where(pos < 0, -of_i, of_i)).view(np.float32)
    ref = np.where(np.abs(pos) < origin, pos + f_scale * nrm, p_i).astype(np.float32)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    # and the offset always moves the point to the normal's side
    d = np.einsum("ij,ij->i", (out.astype(np.float64) - pos.astype(np.float64)), nrm.astype(np.float64))
    assert (d >= 0).all()


def test_sampled_spectrum_known_answers(oracle):  # SampledSpectrumTests.cpp:33-57
    samples = np.array([0.0, 0.25, 1.0, 0.75, 0.0], np.float32)
    for lam, expected in [(200.0, 0.0), (450.0, 0.125), (600.0, 1.0), (725.0, 0.5625), (800.0, 0.0), (810.0, 0.0)]:
        r = oracle.lib().oracle_sampled_spectrum_eval(C.c_float(400.0), C.c_float(800.0), C.c_uint32(5), samples.ctypes.data_as(C.c_void_p), C.c_float(lam))
        assert r == expected, (lam, r)


def test_wavelength_to_xyz_against_cie1931():  # SpectrumUtilsTests.cpp:40-80 (tolerances from the reference test)
    from oracle import binding
    cie = json.load(open(os.path.join(GOLDEN, "cie1931_1nm.json")))
    table = np.array(cie["xyz"], np.float64)
    n = 20000
    rng = np.random.default_rng(5)
    lam = (300.0 + (np.arange(n) + rng.uniform(0, 1, n)) / n * 600.0).astype(np.float32)
    out = np.zeros((n, 3), np.float32)
    binding.lib().oracle_wavelength_to_xyz(C.c_uint32(n), lam.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    # SampledSpectrum::eval semantics: linear interpolation inside [360,830], zero outside
    x = (lam.astype(np.float64) - cie["start_nm"])
    i0 = np.clip(np.floor(x).astype(int), 0, 469)
    w = x - i0
    ref = table[i0] * (1 - w[:, None]) + table[i0 + 1] * w[:, None]
    ref[(lam < cie["start_nm"]) | (lam > cie["end_nm"])] = 0
    assert (out >= 0).all()
    err = ((ref - out) ** 2).max(axis=0)
    assert err[0] <= 2.0e-4 and err[1] <= 6.6e-5 and err[2] <= 5.2e-4, err


def test_f16_roundtrip_matches_ieee(oracle):
    rng = np.random.default_rng(6)
    x = np.concatenate([rng.normal(0, 1, 4096), rng.normal(0, 1e-6, 512), rng.normal(0, 3e4, 512), [0.0, -0.0, 65504.0, 1e-8, 6.1e-5, 5.96e-8]]).astype(np.float32)
    h = np.zeros(len(x), np.uint16)
    back = np.zeros(len(x), np.float32)
    oracle.lib().oracle_f16_roundtrip(C.c_uint32(len(x)), x.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), back.ctypes.data_as(C.c_void_p))
    with np.errstate(over="ignore"):
        ref = x.astype(np.float16)
    finite = np.isfinite(ref)
    assert np.array_equal(h[finite], ref.view(np.uint16)[finite])
    assert np.array_equal(back[finite], ref.astype(np.float32)[finite])


def test_equal_area_octahedral_roundtrip(oracle):  # MathHelpers.slang:199-240 (bounce records store wi this way)
    rng = np.random.default_rng(7)
    d = rng.normal(size=(4096, 3)).astype(np.float32)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    enc = np.zeros((len(d), 2), np.float32)
    dec = np.zeros_like(d)
    oracle.lib().oracle_oct_equal_area(C.c_uint32(len(d)), d.ctypes.data_as(C.c_void_p), enc.ctypes.data_as(C.c_void_p), dec.ctypes.data_as(C.c_void_p))
    assert (np.abs(enc) <= 1.0 + 1e-6).all()
    assert np.abs(dec - d).max() < 2e-3  # the reference's own maps are only approximately inverse in fp32 near the poles
    assert np.median(np.abs(dec - d)) < 1e-6


def test_hwss_wavelengths(oracle):  # SpectrumSampler.slang:44-54
    out = (C.c_float * 4)()
    oracle.lib().oracle_hwss_wavelengths(C.c_float(700.0), C.c_int(4), out)
    assert list(out) == [700.0, 445.0, 530.0, 615.0]
    oracle.lib().oracle_hwss_wavelengths(C.c_float(380.0), C.c_int(2), out)
    assert list(out)[:2] == [380.0, 550.0]


def test_spectral_profile_two_restatements_agree(oracle):
    """genSpectralProfile (SceneBuilder.cpp:911-960): the C++ oracle and the numpy harness were written independently."""
    from pltfalcor_b200 import layouts as L, scene
    t = scene.bundled_tables()
    cases = [(t["ior"]["Au"]["eta"], False), (t["ior"]["Au"]["k"], False), (t["emission"][scene.LAMP], True)]
    for tab, icdf in cases:
        w, v = np.asarray(tab["wavelengths"], np.float32), np.asarray(tab["values"], np.float32)
        ours = scene.gen_spectral_profile(w, v, compute_icdf=icdf)
        ref = np.zeros(1, L.spectral_profile_dtype)
        oracle.lib().oracle_gen_spectral_profile(C.c_uint32(len(w)), w.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_float(1.0), C.c_int(int(icdf)), ref.ctypes.data_as(C.c_void_p))
        ours = np.asarray(ours).reshape(1)
        assert ours["bins"][0] == ref["bins"][0]
        b = int(ref["bins"][0])
        assert ours["minWavelength"][0] == ref["minWavelength"][0] and ours["maxWavelength"][0] == ref["maxWavelength"][0]
        np.testing.assert_allclose(ours["pdf"][0][:b], ref["pdf"][0][:b], rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(ours["intensity"][0], ref["intensity"][0], rtol=2e-5)
        # `rgb` is host-only metadata (flux integration in LightCollection); the C++ oracle leaves it zero.
        if icdf:
            np.testing.assert_allclose(ours["icdf"][0][:b], ref["icdf"][0][:b], rtol=1e-4, atol=1e-5)


def test_eval_spectrum_interpolates_gold(oracle):
    """Scene.slang:251-294 on the Au table: inside the table the profile must reproduce the .spd rows (linear in between)."""
    from pltfalcor_b200 import scene
    t = scene.bundled_tables()["ior"]["Au"]["eta"]
    w, v = np.asarray(t["wavelengths"], np.float32), np.asarray(t["values"], np.float32)
    prof = np.asarray(scene.gen_spectral_profile(w, v, compute_icdf=False)).reshape(1)
    lib = oracle.lib()
    for lam in np.linspace(385, 715, 67):
        got = lib.oracle_eval_spectrum(prof.ctypes.data_as(C.c_void_p), C.c_float(lam))
        ref = np.interp(lam, w, v)
        assert abs(got - ref) < 0.02 * max(1.0, abs(ref)), (lam, got, ref)  # table is resampled to <= 128 bins
    assert lib.oracle_eval_spectrum(prof.ctypes.data_as(C.c_void_p), C.c_float(300.0)) == 0.0
    assert lib.oracle_eval_spectrum(prof.ctypes.data_as(C.c_void_p), C.c_float(800.0)) == 0.0


def test_oracle_frames_match_their_pinned_copies():
    """tests/golden/oracle_frames.npz (tools/make_golden_oracle.py): the oracle's own small frames.  Not a reference image (none exists,
    DESIGN §5) — a tripwire for accidental changes of the checker: discrete path records identical, floats to 1e-4 (libm may differ)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_golden_oracle", os.path.join(os.path.dirname(__file__), "..", "tools", "make_golden_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_frames.npz"))
    for name in mod.CASES:
        got = mod.render(name)
        same = (got["hit"] == gold[f"{name}/hit"]).all(axis=(-1, -2)) & (got["lobe"] == gold[f"{name}/lobe"]).all(-1) & (got["flags"] == gold[f"{name}/flags"]).all(-1)
        assert same.mean() >= 0.99, name  # a path may fork where a random number sits within an ulp of a branch threshold
        for k in ("color", "albedo", "normal"):
            d = np.abs(got[k] - gold[f"{name}/{k}"])[same]
            assert (d <= 1e-4 * np.maximum(np.abs(gold[f"{name}/{k}"])[same], 1e-3)).mean() >= 0.995, (name, k, float(d.max()))
