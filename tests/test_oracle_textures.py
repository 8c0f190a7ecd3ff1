"""The oracle's material-texture filter (oracle/plt_bsdf.h TextureSystem) against an independent numpy restatement of the D3D sampling
rules and against known answers; normal mapping against its closed form.  CPU only."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def tex_scene(oracle):
    from pltfalcor_b200 import scene
    buffers, cam = scene.textured_scene()
    return buffers, cam, oracle.OracleScene(buffers)


def _address(i, n, mode):
    if mode == 0:
        return i % n, True
    if mode == 1:
        p = i % (2 * n)
        return (p if p < n else 2 * n - 1 - p), True
    if mode == 2:
        return min(max(i, 0), n - 1), True
    if mode == 4:
        i = -1 - i if i < 0 else i
        return min(i, n - 1), True
    return i, 0 <= i < n


def _decode(t, fmt):
    t = np.asarray(t)
    if fmt in (1, 2):
        c = t.astype(np.float32) / np.float32(255)
        if fmt == 2:
            rgb = c[..., :3].astype(np.float64)
            c[..., :3] = np.where(rgb <= 0.04045, rgb / 12.92, ((rgb + 0.055) / 1.055) ** 2.4)
        return c
    return t.astype(np.float32)


def _numpy_sample(texels, fmt, sampler, uv):
    linear, au, av, border = sampler
    tex = _decode(texels, fmt)
    H, W = tex.shape[:2]
    out = np.zeros((len(uv), 4), np.float32)

    def texel(x, y):
        x, okx = _address(x, W, au)
        y, oky = _address(y, H, av)
        return tex[y, x] if okx and oky else np.asarray(border, np.float32)

    for k, (u, v) in enumerate(uv):
        if not linear:
            out[k] = texel(int(np.floor(np.float32(u) * np.float32(W))), int(np.floor(np.float32(v) * np.float32(H))))
            continue
        x, y = np.float32(u) * np.float32(W) - np.float32(0.5), np.float32(v) * np.float32(H) - np.float32(0.5)
        x0, y0 = np.floor(x), np.floor(y)
        fx, fy = np.float32(x - x0), np.float32(y - y0)
        ix, iy = int(x0), int(y0)
        top = texel(ix, iy) * (np.float32(1) - fx) + texel(ix + 1, iy) * fx
        bot = texel(ix, iy + 1) * (np.float32(1) - fx) + texel(ix + 1, iy + 1) * fx
        out[k] = top * (np.float32(1) - fy) + bot * fy
    return out


def test_filter_matches_numpy_restatement(tex_scene):
    buffers, cam, osc = tex_scene
    rng = np.random.default_rng(2)
    uv = rng.uniform(-2.5, 3.5, (600, 2)).astype(np.float32)
    samplers = buffers.scalars["samplers"] + [(1, 0, 0, (0, 0, 0, 0))]   # + the default sampler (linear, wrap)
    for tid, (w, h, fmt) in enumerate(buffers.scalars["textures"]):
        for sid, smp in enumerate(samplers):
            got = osc.texture_sample((1 << 29) | tid, sid, uv)
            ref = _numpy_sample(buffers.arrays[f"texture{tid}"], fmt, smp, uv)
            assert np.abs(got - ref).max() <= (3e-6 if fmt == 2 else 1e-6), (tid, sid)


def test_filter_known_answers(tex_scene):
    buffers, cam, osc = tex_scene
    spec = buffers.arrays["texture2"]                         # 32 x 32 RGBA32F
    centres = np.array([[(x + 0.5) / 32, (y + 0.5) / 32] for x, y in ((0, 0), (5, 9), (31, 31))], np.float32)
    got = osc.texture_sample((1 << 29) | 2, 0, centres)
    assert np.array_equal(got, np.stack([spec[0, 0], spec[9, 5], spec[31, 31]]))          # a texel centre returns the texel
    got = osc.texture_sample((1 << 29) | 2, 0, np.array([[6.0 / 32, 9.5 / 32]], np.float32))  # on the edge between texels 5 and 6
    assert np.allclose(got[0], 0.5 * (spec[9, 5] + spec[9, 6]), atol=1e-6)
    a = osc.texture_sample((1 << 29) | 2, 0, np.array([[0.3, 0.7]], np.float32))
    b = osc.texture_sample((1 << 29) | 2, 0, np.array([[2.3, -1.3]], np.float32))           # wrap: period 1
    assert np.allclose(a, b, atol=2e-6)
    far = osc.texture_sample((1 << 29) | 2, 3, np.array([[1.7, 0.5]], np.float32))          # border sampler, outside
    assert np.allclose(far[0], [0.1, 0.6, 0.1, 1.0])
    srgb = osc.texture_sample((1 << 29) | 0, 1, np.array([[(7 + 0.5) / 96, (3 + 0.5) / 64]], np.float32))[0]
    t = buffers.arrays["texture0"][3, 7].astype(np.float64) / 255
    assert np.allclose(srgb[:3], np.where(t[:3] <= 0.04045, t[:3] / 12.92, ((t[:3] + 0.055) / 1.055) ** 2.4), atol=1e-6) and srgb[3] == 1.0
    assert np.allclose(osc.texture_sample(0, 0, centres, uniform=(1, 2, 3, 4)), [1, 2, 3, 4])  # Uniform handle
    assert np.all(osc.texture_sample(2 << 29, 0, centres) == 0)                                 # unknown mode: float4(0)


def test_guides_carry_the_sampled_parameters(tex_scene):
    """First-hit guides of the oracle frame: the floor's albedo is the sRGB texture at the hit's uv, and the normal-mapped back wall's
    shading normal is the tangent-frame transform of the decoded normal-map texel (ShadingUtils.slang:199-246)."""
    from pltfalcor_b200 import layouts as L, scene
    buffers, cam, osc = tex_scene
    opts = L.PltptOptions()
    for k, v in L.OPTION_DEFAULTS.items():
        setattr(opts, k, type(getattr(opts, k))(v))
    opts.maxBounces = 2
    w, h = 96, 54
    r = osc.render_frame(opts, cam, 0, w, h, want_bounces=True)
    alb, nrm, b0 = r["albedo"], r["normal"], r["bounces"][:, :, 0]
    inst = (b0["hit"][..., 0] >> buffers.scalars["primitiveIndexBits"]) & ((1 << max(buffers.scalars["instanceIDBits"], 1)) - 1)
    floor = (inst == 0) & ((b0["hit"][..., 0] >> 29) == 1)
    assert floor.sum() > 200
    assert alb[floor][:, :3].std(axis=0).min() > 0.02          # the checker texture shows in the guide, not one flat colour
    wall = (inst == 2) & ((b0["hit"][..., 0] >> 29) == 1)
    assert wall.sum() > 200
    n = nrm[wall][:, :3]
    assert np.allclose(np.linalg.norm(n, axis=1), 1, atol=1e-4)
    assert n[:, 2].min() > 0.5 and n[:, 2].max() > 0.99 and np.abs(n[:, 0]).max() > 0.15   # bumps tilt the view-space normal, never flip it
