"""Environment map: lat-long evaluation, hierarchical importance map, sample / evalPdf.
CPU tests pin the oracle on identities the reference's construction implies (EnvMapSampler.slang:71-162, EnvMapSamplerSetup.cs.slang:47-75);
the GPU tests compare pltpt_envmap_eval / pltpt_envmap_sample with the oracle."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def env(oracle):
    from pltfalcor_b200 import scene
    buffers, cam = scene.env_scene()
    return buffers, oracle.OracleScene(buffers)


def _dirs(rng, n):
    d = rng.normal(size=(n, 3))
    return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


def test_oracle_eval_at_texel_centres_returns_the_texel(env):
    buffers, osc = env
    tex, s = buffers.arrays["envMapTexels"], buffers.scalars
    h, w = tex.shape[:2]
    rng = np.random.default_rng(0)
    ys, xs = rng.integers(1, h - 1, 200), rng.integers(0, w, 200)
    u, v = (xs + 0.5) / w, (ys + 0.5) / h
    phi, theta = np.pi * (2 * u - 1), np.pi * v                     # latlong_map_to_world (MathHelpers.slang:120-129)
    local = np.stack([np.sin(theta) * np.sin(phi), np.cos(theta), -np.sin(theta) * np.cos(phi)], -1)
    world = local @ np.asarray(s["envMap"]["transform"], np.float64)[:, :3].T
    got = osc.envmap_eval(world.astype(np.float32))[:, :3]
    want = tex[ys, xs, :3] * np.float32(s["envMap"]["intensity"]) * np.asarray(s["envMap"]["tint"], np.float32)
    # the direction -> uv round trip is exact to ~1e-6 texels; neighbouring texels differ by <= 10 % noise
    assert np.allclose(got, want, rtol=2e-3, atol=1e-4)


def test_oracle_importance_map_is_a_normalised_density(env):
    _, osc = env
    imp = osc.envmap_importance()
    assert imp.shape == (512, 512) and (imp > 0).all()
    rng = np.random.default_rng(1)
    s = osc.envmap_sample(rng.random((200_000, 2)).astype(np.float32))
    assert np.allclose(np.linalg.norm(s[:, :3], axis=1), 1, atol=1e-5)
    # E_p[1/p] = measure of the sphere
    assert abs(np.mean(1.0 / s[:, 3].astype(np.float64)) / (4 * np.pi) - 1) < 0.02
    # the pdf of a sampled direction is the pdf evalPdf reports for it (same texel) — except on texel borders
    e = osc.envmap_eval(s[:, :3])
    assert (np.abs(e[:, 3] - s[:, 3]) <= 1e-5 * s[:, 3]).mean() > 0.98
    assert np.array_equal(e[:, :3], s[:, 4:7])
    # brighter texels are sampled proportionally more often: the sun texel region dominates
    assert s[:, 3].max() / np.median(s[:, 3]) > 20


def test_oracle_env_lighting_switch(env, oracle):
    from pltfalcor_b200 import layouts as L, scene
    buffers, osc = env
    _, cam = scene.env_scene()
    on = osc.render_frame(L.make_options(maxBounces=4), scene.with_jitter(cam, 0, 48, 27), 0, 48, 27)["color"]
    off = osc.render_frame(L.make_options(maxBounces=4, useEnvLights=False), scene.with_jitter(cam, 0, 48, 27), 0, 48, 27)["color"]
    assert np.isfinite(on).all() and (on[:3, :, :3] > 0).all()      # top rows look at the sky
    assert (off[:3, :, :3] == 0).all() and on[..., :3].mean() > 5 * off[..., :3].mean()


def test_scene_file_carries_the_env_map(tmp_path, env):
    from pltfalcor_b200 import scene
    buffers, _ = env
    _, cam = scene.env_scene()
    path = tmp_path / "env.pltscn"
    buffers.save(str(path), cam)
    raw = path.read_bytes()
    tex = buffers.arrays["envMapTexels"]
    # the env-map section is followed by the (here empty) texture / sampler section: two u32 counts
    assert raw.endswith(tex.tobytes() + b"\0" * 8) and len(raw) > tex.nbytes + 112 + 16 + 8


@pytest.mark.gpu
def test_gpu_envmap_eval_and_pdf_match_oracle(env, torch_cuda):
    torch = torch_cuda
    from pltfalcor_b200 import binding
    buffers, osc = env
    p = binding.Pltpt()
    p.set_scene(buffers)
    d = _dirs(np.random.default_rng(2), 100_000)
    out = torch.zeros((len(d), 4), device="cuda")
    p.envmap_eval(torch.from_numpy(d).cuda(), out)
    got, want = out.cpu().numpy(), osc.envmap_eval(d)
    rel = np.abs(got[:, :3] - want[:, :3]) / np.maximum(np.abs(want[:, :3]), 1e-6)
    # atan2/acos differ by an ulp between libm and CUDA: uv moves by ~1e-7, the bilinear result by <= 1e-4 relative (tolerance of north_star)
    assert (rel.max(axis=1) <= 1e-4).mean() >= 0.999, rel.max()
    # pdf is piecewise constant: equal except where the direction sits on a texel border of the importance map
    assert (np.abs(got[:, 3] - want[:, 3]) <= 1e-5 * want[:, 3]).mean() >= 0.999
    p.close()


@pytest.mark.gpu
def test_gpu_envmap_sample_matches_oracle(env, torch_cuda):
    torch = torch_cuda
    from pltfalcor_b200 import binding
    buffers, osc = env
    p = binding.Pltpt()
    p.set_scene(buffers)
    rnd = np.random.default_rng(3).random((100_000, 2)).astype(np.float32)
    out = torch.zeros((len(rnd), 8), device="cuda")
    p.envmap_sample(torch.from_numpy(rnd).cuda(), out)
    got, want = out.cpu().numpy(), osc.envmap_sample(rnd)
    same_texel = np.abs(got[:, 3] - want[:, 3]) <= 1e-5 * want[:, 3]
    assert same_texel.mean() >= 0.999                                  # the 9-level warp takes the same branches
    assert np.abs(got[same_texel, :3] - want[same_texel, :3]).max() <= 2e-4   # the warped sub-texel position amplifies rounding by <= 2^9
    rel = np.abs(got[same_texel, 4:7] - want[same_texel, 4:7]) / np.maximum(want[same_texel, 4:7], 1e-6)
    assert (rel.max(axis=1) <= 1e-2).mean() >= 0.999
    p.close()


@pytest.mark.gpu
def test_gpu_env_map_requires_a_map(cornell, torch_cuda):
    torch = torch_cuda
    from pltfalcor_b200 import binding
    buffers, _ = cornell
    p = binding.Pltpt()
    p.set_scene(buffers)
    with pytest.raises(binding.PltptError):
        p.envmap_eval(torch.zeros((4, 3), device="cuda"), torch.zeros((4, 4), device="cuda"))
    p.close()
