"""The radiometric-law checks of tests/test_oracle_radiometry.py on the CUDA integrator (pltpt_render_frame_host)."""
import numpy as np
import pytest

from pltfalcor_b200 import layouts as L
from pltfalcor_b200 import scene as S

pytestmark = pytest.mark.gpu


def _lit_plane(h, d, albedo, half=0.05, frames=48):
    from pltfalcor_b200 import binding
    b = S.SceneBuilder()
    lamp = b.add_emission_profile(S.LAMP)
    ground = b.add_diffuse((albedo,) * 3)
    light = b.add_diffuse((0.0, 0.0, 0.0), emission_profile=lamp)
    b.add_mesh(S.quad([-10, 0, 10], [10, 0, 10], [10, 0, -10], [-10, 0, -10]), ground)
    b.add_mesh(S.quad([-half, h, -half], [half, h, -half], [half, h, half], [-half, h, half]), light)
    cam = S.make_camera([d, 3.0, 4.0], [d, 0.0, 0.0], focal_length=2000.0, aspect=1.0)
    p = binding.Pltpt(maxBounces=2)
    p.set_scene(b.finalize())
    w = 8
    acc = np.zeros(3)
    for f in range(frames):
        acc += p.render_frame_host(S.with_jitter(cam, f, w, w), f, w, w, want_guides=False)["color"][..., :3].reshape(-1, 3).mean(axis=0)
    p.close()
    return acc / frames


def test_cuda_direct_lighting_follows_inverse_square_cosine_and_albedo(torch_cuda):
    d = 0.5
    near, far, dark = _lit_plane(1.0, d, 0.5), _lit_plane(2.0, d, 0.5), _lit_plane(1.0, d, 0.25)
    geom = lambda h: h * h / (d * d + h * h) ** 2
    assert near.min() > 0 and far.min() > 0 and dark.min() > 0
    assert np.abs(near / far / (geom(1.0) / geom(2.0)) - 1).max() < 0.03, near / far
    assert np.abs(near / dark / 2.0 - 1).max() < 0.03, near / dark
