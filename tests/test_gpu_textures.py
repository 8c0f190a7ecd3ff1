"""Textured material parameters, normal maps and alpha cut-outs (MaterialSystem.slang:283-291, *Material.slang::setupMaterialInstance,
ShadingUtils.slang:199-246) against the oracle: the texture filter itself, then the sample and solve stages on a scene whose
materials read their slots from textures of every supported format through every address mode."""
import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def loaded(torch_cuda, oracle):
    from pltfalcor_b200 import binding, scene
    buffers, cam = scene.textured_scene()
    p = binding.Pltpt(maxBounces=6)
    p.set_scene(buffers)
    osc = oracle.OracleScene(buffers)
    yield buffers, cam, p, osc
    p.close()


def test_texture_filter_matches_oracle(loaded, torch_cuda):
    """Every texture (RGBA8 sRGB / RGBA8 / RGBA32F / RGBA16F) through every sampler (wrap, clamp, point + mirror / mirror-once, border)
    and the default sampler, at coordinates inside and far outside [0, 1], on texel centres and on texel edges."""
    torch = torch_cuda
    buffers, cam, p, osc = loaded
    rng = np.random.default_rng(11)
    uv = rng.uniform(-2.5, 3.5, (4096, 2)).astype(np.float32)
    uv[:64] = (rng.integers(-20, 60, (64, 2)) + 0.5) / np.float32(16)   # texel centres of the 16-texel textures
    uv[64:128] = rng.integers(-20, 60, (64, 2)) / np.float32(16)        # texel edges
    uv[128] = [0.0, 0.0]; uv[129] = [1.0, 1.0]; uv[130] = [-0.0, 1.0]
    duv = torch.from_numpy(uv).cuda()
    out = torch.zeros((len(uv), 4), device="cuda")
    worst = 0.0
    for tid in range(len(buffers.scalars["textures"])):
        for smp in range(len(buffers.scalars["samplers"]) + 1):
            handle = (1 << 29) | tid
            p.texture_sample(handle, smp, duv, out)
            torch.cuda.synchronize()
            ref = osc.texture_sample(handle, smp, uv)
            got = out.cpu().numpy()
            # identical fp32 operations; only the sRGB decode (powf) may differ in the last bits
            tol = 2e-6 if buffers.scalars["textures"][tid][2] == 2 else 0.0
            err = np.abs(got - ref).max()
            worst = max(worst, float(err))
            assert err <= tol, (tid, smp, err)
    # a Uniform handle returns the uniform value
    p.texture_sample(0, 0, duv[:4], out[:4], uniform=(0.1, 0.2, 0.3, 0.4))
    torch.cuda.synchronize()
    assert np.allclose(out[:4].cpu().numpy(), [0.1, 0.2, 0.3, 0.4])
    parity.record("textures/filter_max_abs_err", {"max_abs_err": worst})


W, H = 96, 54


def test_textured_scene_sample_stage(loaded):
    """Bounce records: alpha cut-outs are passed through, normal-mapped frames and textured roughness steer the same samples."""
    from pltfalcor_b200 import scene
    buffers, cam, p, osc = loaded
    for frame in (0, 3):
        c = scene.with_jitter(cam, frame, W, H)
        gpu = p.sample_stage_host(c, frame, W, H)
        cpu = osc.render_frame(p.options, c, frame, W, H, solve=False, want_bounces=True)["bounces"]
        m, same = parity.compare_bounces(gpu, cpu)
        parity.record(f"bounces/textured/frame{frame}", m)
        # the per-hit material derivation runs CUDA's powf / sqrtf against glibc's in the oracle: a few more paths sit on a branch threshold
        assert m["identical_paths"] >= 0.99, m
        assert m["max_bary_abs"] <= 2e-3 and m["max_wi_abs"] <= 2e-3, m
        # pathContribution is a product of up to seven sampling pdfs, each through the per-hit derivation (measured 1.1e-3; 5.7e-4 untextured)
        assert m["max_contribution_rel"] <= 2e-3, m


def test_textured_scene_single_frame(loaded):
    from pltfalcor_b200 import scene
    buffers, cam, p, osc = loaded
    frame = 2
    c = scene.with_jitter(cam, frame, W, H)
    gpu_b = p.sample_stage_host(c, frame, W, H)
    gpu = p.render_frame_host(c, frame, W, H)
    cpu = osc.render_frame(p.options, c, frame, W, H, want_bounces=True)
    _, same = parity.compare_bounces(gpu_b, cpu["bounces"])
    m = parity.compare_color(gpu["color"], cpu["color"], same)
    m["identical_paths"] = float(same.mean())
    parity.record("single_frame/textured", m)
    assert m["nonfinite_gpu"] == 0
    assert m["frac_pixels_within_1e-3"] >= 0.985 and m["frac_pixels_within_1e-2"] >= 0.985, m
    assert m["mean_rel_diff"] <= 0.02, m
    for k in ("albedo", "normal"):   # guides carry the sampled base colour and the normal-mapped shading normal
        d = np.abs(gpu[k] - cpu[k]).max(axis=-1)
        assert (d[same] <= 2e-3).mean() >= 0.99, k
    assert p.stats()["droppedConnections"] == 0


def test_textures_change_the_image(loaded, torch_cuda):
    """The same scene with every handle reset to Uniform renders a different image: the textured path is really taken."""
    from pltfalcor_b200 import binding, scene
    buffers, cam, p, osc = loaded
    c = scene.with_jitter(cam, 1, W, H)
    a = p.render_frame_host(c, 1, W, H)
    flat, _ = scene.textured_scene()
    flat.arrays["materialData"]["tex"][:] = 0
    q = binding.Pltpt(maxBounces=6)
    q.set_scene(flat)
    b = q.render_frame_host(c, 1, W, H)
    q.close()
    assert np.abs(a["albedo"] - b["albedo"]).max() > 0.2
    assert np.abs(a["normal"] - b["normal"]).max() > 0.05


def test_texture_handles_are_validated(torch_cuda):
    from pltfalcor_b200 import binding, scene
    buffers, _ = scene.textured_scene()
    buffers.arrays["materialData"]["tex"][1, 0] = (1 << 29) | 77   # texture id beyond the table
    p = binding.Pltpt()
    with pytest.raises(binding.PltptError):
        p.set_scene(buffers)
    buffers, _ = scene.textured_scene()
    buffers.arrays["materialData"]["tex"][1, 0] = (1 << 31) | (1 << 29)  # UDIM
    with pytest.raises(binding.PltptError):
        p.set_scene(buffers)
    p.close()
