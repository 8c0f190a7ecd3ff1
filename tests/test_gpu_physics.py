"""The closed-form optics checks of tests/test_oracle_physics.py run on the CUDA path itself (pltpt_bsdf_eval_mueller): Fresnel at normal
incidence, at Brewster's angle and over all angles, the complex-index Fresnel equations for the Au mirror, energy conservation of the
lossless interface and the Lambertian, depolarising diffuse lobe.  Directions that have to be sampled come from the oracle's sample()."""
import numpy as np
import pytest

from tests import test_oracle_physics as physics

pytestmark = pytest.mark.gpu


class CudaEval:
    """bsdf_eval_mueller with the oracle binding's signature, evaluated by the CUDA library."""

    def __init__(self, torch, p, oracle_scene):
        self.torch, self.p, self.o = torch, p, oracle_scene

    def bsdf_eval_mueller(self, material_id, lobe, wi, wo, hero, hwss=4):
        torch = self.torch
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
        self.p.set_options(HWSS=hwss)   # the output stride follows the handle's HWSS option
        out = torch.full((len(wi), np.shape(hero)[0], hwss, 4, 4), float("nan"), device="cuda")
        self.p.bsdf_eval_mueller(material_id, lobe, dev(wi), dev(wo), dev(hero), out)
        torch.cuda.synchronize()
        return out.cpu().numpy()

    def bsdf_sample(self, *args, **kwargs):
        return self.o.bsdf_sample(*args, **kwargs)


@pytest.fixture(scope="module")
def cuda(torch_cuda, oracle, sweep):
    from pltfalcor_b200 import binding
    buffers, ids = sweep
    p = binding.Pltpt()
    p.set_scene(buffers)
    yield CudaEval(torch_cuda, p, oracle.OracleScene(buffers)), ids
    p.close()


@pytest.mark.parametrize("check", [physics.test_fresnel_reflectance_near_normal_incidence, physics.test_brewster_angle_reflection_is_fully_polarised,
                                   physics.test_fresnel_equations_over_all_angles, physics.test_lossless_interface_conserves_energy,
                                   physics.test_gold_mirror_follows_the_complex_fresnel_equations], ids=lambda f: f.__name__[5:])
def test_cuda_bsdf_against_closed_form(cuda, check):
    check(cuda)


def test_cuda_diffuse_is_lambertian_and_depolarising(cuda, sweep):
    physics.test_diffuse_is_lambertian_and_depolarising(cuda, sweep)
