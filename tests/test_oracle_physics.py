"""Closed-form optics the CPU oracle must reproduce.  The reference holds no test for its PLT BSDFs (parity unpinned), so the
oracle — the checker of every GPU parity test — is itself checked against textbook results that do not depend on the reference's code:
Fresnel reflectance at normal incidence, Brewster's angle, energy conservation of a lossless interface, Mueller-matrix structure."""
import numpy as np
import pytest

from pltfalcor_b200 import layouts as L


@pytest.fixture(scope="module")
def osc(oracle, sweep):
    buffers, ids = sweep
    return oracle.OracleScene(buffers), ids


def _eta(wl):
    """Cauchy dispersion of the sweep scene's smooth dielectric with its parameters as packed into fp16 (PLTDielectricMaterial.slang:40-57)."""
    B, ior = float(np.float16(0.01)), float(np.float16(1.5))
    return max(1.0, ior - 2.897 * B) + B * (1.0 / (np.asarray(wl, np.float64) * 1e-3)) ** 2


def _incident(theta, phi):
    return np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], -1).astype(np.float32)


def _mirror(wi):
    return np.stack([-wi[:, 0], -wi[:, 1], wi[:, 2]], -1).astype(np.float32)


def _reflect(osc, wi, wl):
    o, ids = osc
    return o.bsdf_eval_mueller(ids["smooth_dielectric"], L.LOBE_DELTA_REFLECTION, wi, _mirror(wi), wl[None], hwss=1)[:, 0, 0].astype(np.float64)


# bsdf_eval_mueller returns, per wavelength, the responses to the four basis Stokes vectors: out[i] = M e_i (column i of the Mueller matrix),
# each expressed in the outgoing beam's frame, which is rotated about the ray with respect to the plane of incidence.  The checks below
# therefore use rotation invariants: m00, the polarizance |(m10, m20)|, the diattenuation |(m01, m02)| and m33.
# The specular colour (1, 1, 1) goes through the 10-bin rgbToSpectrum basis (SpectrumUtils.slang:106-239), whose white is 1 only to
# within 8e-4 between 450 and 580 nm; hence the relative tolerance.
REL = 1.2e-3


def _close(got, want, rel=REL, abs_=2e-7):
    return np.all(np.abs(got - want) <= rel * np.abs(want) + abs_)


def test_fresnel_reflectance_near_normal_incidence(osc):
    n = 64
    wl = np.linspace(400, 700, n).astype(np.float32)
    wi = _incident(np.full(n, 1e-3), np.linspace(0, 6.0, n))
    M = _reflect(osc, wi, wl)
    eta = _eta(wl)
    assert _close(M[:, 0, 0], ((eta - 1) / (eta + 1)) ** 2)
    # at normal incidence s and p are indistinguishable: no polarizance, no diattenuation
    assert np.abs(M[:, 0, 1:3]).max() < 1e-5 and np.abs(M[:, 1:3, 0]).max() < 1e-5


def test_brewster_angle_reflection_is_fully_polarised(osc):
    n = 48
    wl = np.linspace(400, 700, n).astype(np.float32)
    eta = _eta(wl)
    wi = _incident(np.arctan(eta), np.linspace(0.1, 6.0, n))
    M = _reflect(osc, wi, wl)
    # Rp = 0 at Brewster's angle: m00 = Rs / 2 with Rs = ((eta^2 - 1) / (eta^2 + 1))^2, and unpolarised light leaves fully polarised
    rs = ((eta ** 2 - 1) / (eta ** 2 + 1)) ** 2
    assert _close(M[:, 0, 0], rs / 2)
    assert np.abs(np.hypot(M[:, 0, 1], M[:, 0, 2]) / M[:, 0, 0] - 1).max() < 1e-5
    assert np.abs(np.hypot(M[:, 1, 0], M[:, 2, 0]) / M[:, 0, 0] - 1).max() < 1e-5
    assert np.abs(M[:, 3, 3]).max() < 1e-5 and np.abs(M[:, 0, 3]).max() == 0   # no circular component, sqrt(Rs Rp) = 0


def test_fresnel_equations_over_all_angles(osc):
    rng = np.random.default_rng(3)
    n = 2048
    theta = np.arccos(rng.uniform(0.02, 1.0, n))
    wl = rng.uniform(400, 700, n).astype(np.float32)
    wi = _incident(theta, rng.uniform(0, 2 * np.pi, n))
    M = _reflect(osc, wi, wl)
    eta = _eta(wl)
    ci = np.cos(theta)
    ct = np.sqrt(1 - (np.sin(theta) / eta) ** 2)
    rs = ((ci - eta * ct) / (ci + eta * ct)) ** 2
    rp = ((eta * ci - ct) / (eta * ci + ct)) ** 2
    assert _close(M[:, 0, 0], (rs + rp) / 2, abs_=2e-6)
    assert _close(np.hypot(M[:, 0, 1], M[:, 0, 2]), np.abs(rp - rs) / 2, abs_=2e-6)   # polarizance
    assert _close(np.hypot(M[:, 1, 0], M[:, 2, 0]), np.abs(rp - rs) / 2, abs_=2e-6)   # diattenuation
    # real rs, rp below the critical angle: m33 = rs rp, no retardance between linear and circular components
    assert _close(np.abs(M[:, 3, 3]), np.sqrt(rs * rp), abs_=2e-6)


def test_lossless_interface_conserves_energy(osc):
    o, ids = osc
    rng = np.random.default_rng(8)
    wi = _incident(np.arccos(rng.uniform(0.05, 1.0, 2048)), rng.uniform(0, 2 * np.pi, 2048))
    wl = rng.uniform(400, 700, 2048).astype(np.float32)
    s = o.bsdf_sample(ids["smooth_dielectric"], L.LOBE_DELTA & L.LOBE_TRANSMISSION, wi, wl, seed=1)
    ok = s[:, 6] > 0
    assert ok.mean() > 0.5
    wi, wl, wt = wi[ok], wl[ok], np.ascontiguousarray(s[ok, 0:3])
    R = _reflect(osc, wi, wl)[:, 0, 0]
    T = o.bsdf_eval_mueller(ids["smooth_dielectric"], L.LOBE_DISPERSIVE_DELTA_TRANSMISSION, wi, wt, wl[None], hwss=1)[:, 0, 0, 0, 0].astype(np.float64)
    eta = _eta(wl)
    assert np.abs(R + T * eta ** 2 - 1.0).max() < 2e-3   # transmitted radiance carries the 1/eta^2 solid-angle factor
    # the sampled transmission direction obeys Snell's law
    sin_i, sin_t = np.sqrt(1 - wi[:, 2].astype(np.float64) ** 2), np.sqrt(np.maximum(0, 1 - wt[:, 2].astype(np.float64) ** 2))
    assert np.abs(sin_i - eta * sin_t).max() < 2e-3 and (wt[:, 2] < 0).all()


def test_diffuse_is_lambertian_and_depolarising(osc, sweep):
    """PLTDiffuse: f cos = albedo(lambda) / pi * cos for every pair of directions, a pure depolariser, and the albedo spectrum of the
    sweep scene's (0.6, 0.4, 0.2) material passes near its red, green and blue values at 650, 550 and 450 nm."""
    from tests import parity
    o, ids = osc
    rng = np.random.default_rng(2)
    n = 512
    wi, wo = parity.random_dirs(rng, n, upper=True), parity.random_dirs(rng, n, upper=True)
    for lam, channel in ((650, 0.6), (550, 0.4), (450, 0.2)):
        M = o.bsdf_eval_mueller(ids["diffuse"], L.LOBE_DIFFUSE_REFLECTION, wi, wo, np.full((1, n), lam, np.float32), hwss=1)[:, 0, 0].astype(np.float64)
        albedo = M[:, 0, 0] * np.pi / wi[:, 2]
        assert np.ptp(albedo) < 1e-5 * albedo.mean()           # no dependence on either direction beyond the cosine
        assert abs(albedo.mean() - channel) < 0.03
        M[:, 0, 0] = 0
        assert np.abs(M).max() == 0                             # polarised input is lost, unpolarised output only


def test_gold_mirror_follows_the_complex_fresnel_equations(osc):
    """Smooth Au conductor (roughness 0: all energy in the delta lobe) against the complex-index Fresnel equations evaluated in fp64 on the
    tabulated n, k (Source/Tables/ior/Au.*.spd through the committed fixture).  The library resamples the tables to <= 128 bins
    (SceneBuilder.cpp:911-960), hence the tolerance."""
    from pltfalcor_b200 import scene
    o, ids = osc
    t = scene.bundled_tables()["ior"]["Au"]
    rng = np.random.default_rng(4)
    n = 2048
    theta = np.arccos(rng.uniform(0.02, 1.0, n))
    wl = rng.uniform(400, 700, n).astype(np.float32)
    wi = _incident(theta, rng.uniform(0, 2 * np.pi, n))
    M = o.bsdf_eval_mueller(ids["smooth_conductor"], L.LOBE_DELTA_REFLECTION, wi, _mirror(wi), wl[None], hwss=1)[:, 0, 0].astype(np.float64)
    N = np.interp(wl, t["eta"]["wavelengths"], t["eta"]["values"]) + 1j * np.interp(wl, t["k"]["wavelengths"], t["k"]["values"])
    ci, si2 = np.cos(theta), np.sin(theta) ** 2
    root = np.sqrt(N * N - si2)
    rs = (ci - root) / (ci + root)
    rp = (N * N * ci - root) / (N * N * ci + root)
    # lower 2x2 block: the reference multiplies rs and conj(rp) COMPONENT-WISE (`fs*cconj(fp)` on float2, Mueller.slang:76-77; pinned by
    # tests/test_oracle_ref.py), so the block holds Re rs Re rp and -Im rs Im rp instead of Re / Im of rs conj(rp); reproduced as it is
    Rs, Rp, cross = np.abs(rs) ** 2, np.abs(rp) ** 2, rs.real * rp.real - 1j * (rs.imag * rp.imag)
    tol = dict(rel=4e-3, abs_=3.5e-3)   # worst cases sit on gold's absorption edge (500-540 nm) between table knots; the median error is 1e-4
    assert _close(M[:, 0, 0], (Rs + Rp) / 2, **tol)
    assert _close(np.hypot(M[:, 0, 1], M[:, 0, 2]), np.abs(Rs - Rp) / 2, **tol)
    assert _close(np.abs(M[:, 3, 3]), np.abs(cross.real), **tol)                      # linear <-> linear at 45 degrees
    assert _close(np.hypot(M[:, 1, 3], M[:, 2, 3]), np.abs(cross.imag), **tol)        # retardance: linear -> circular
    assert np.median(np.abs(M[:, 0, 0] - (Rs + Rp) / 2)) < 3e-4 and np.median(np.abs(np.hypot(M[:, 1, 3], M[:, 2, 3]) - np.abs(cross.imag))) < 3e-4
    assert M[:, 0, 0].min() > 0.3 and M[:, 0, 0].max() < 1.0                           # gold: 35-40 % in the blue, > 95 % in the red


def test_grating_lobes_obey_the_grating_equation(osc):
    """1-D Au grating, pitch 3 um along x, incidence in the plane normal to the grooves: every sampled diffraction lobe leaves in that
    plane with sin(theta_o) = m lambda / pitch - sin(theta_i), m = +-1..3 equal to the order encoded in bits 8-15 of the lobe word
    (order in the low six bits, 0x40 = sign; PLTDiffractionGratedConductorMaterial.slang:54-66)."""
    o, ids = osc
    rng = np.random.default_rng(6)
    n = 8192
    th = rng.uniform(-1.3, 1.3, n)
    wi = np.stack([np.sin(th), np.zeros(n), np.cos(th)], -1).astype(np.float32)
    wl = rng.uniform(380, 720, n).astype(np.float32)
    s = o.bsdf_sample(ids["grated_conductor"], L.LOBE_ALL, wi, wl, seed=3)
    lobe = s[:, 4].copy().view(np.uint32)
    ok = (s[:, 6] > 0) & ((lobe & L.LOBE_DISPERSIVE_DELTA_REFLECTION) != 0) & (np.abs(s[:, 0:3]).sum(axis=1) > 0)   # evanescent orders come back as a zero direction
    assert ok.sum() > 3000
    wo, w = s[ok, 0:3].astype(np.float64), wi[ok].astype(np.float64)
    order = (wo[:, 0] + w[:, 0]) / (wl[ok].astype(np.float64) * 1e-3 / 3.0)   # wi points away from the surface: sin(theta_i) = -wi.x for the arriving ray
    code = lobe[ok] >> 8
    want = (code & 0x3F).astype(np.float64) * np.where(code & 0x40, 1.0, -1.0)
    assert np.abs(order - want).max() < 2e-4, np.abs(order - want).max()
    assert set(np.unique(np.abs(want))) == {1.0, 2.0, 3.0}
    assert np.abs(wo[:, 1]).max() == 0 and np.abs(np.linalg.norm(wo, axis=1) - 1).max() < 1e-6 and (wo[:, 2] > 0).all()


def test_diffuse_sampling_is_cosine_weighted(osc):
    from tests import parity
    o, ids = osc
    rng = np.random.default_rng(11)
    n = 1 << 16
    wi = parity.random_dirs(rng, n, upper=True)
    s = o.bsdf_sample(ids["diffuse"], L.LOBE_ALL, wi, rng.uniform(400, 700, n).astype(np.float32), seed=5)
    ok = s[:, 6] > 0
    assert ok.mean() > 0.999
    wo, pdf = s[ok, 0:3].astype(np.float64), s[ok, 3].astype(np.float64)   # columns: wo, pdf, lobe bits, weight, valid, pdf() of the pair
    assert np.abs(pdf - wo[:, 2] / np.pi).max() < 1e-6 and (wo[:, 2] > 0).all()
    m = ok.sum()
    assert abs(wo[:, 2].mean() - 2 / 3) < 4 * np.sqrt(1 / 18 / m)            # E[cos] = 2/3, Var[cos] = 1/18 under the cosine density
    assert abs((wo[:, 2] ** 2).mean() - 0.5) < 4 * np.sqrt(1 / 12 / m)        # E[cos^2] = 1/2, Var = 1/12
    phi = np.arctan2(wo[:, 1], wo[:, 0])
    assert abs(np.cos(phi).mean()) < 4 * np.sqrt(0.5 / m) and abs(np.sin(2 * phi).mean()) < 4 * np.sqrt(0.5 / m)   # uniform azimuth
