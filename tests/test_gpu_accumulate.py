"""AccumulatePass stepped frame by frame (pltpt_accumulate_frame; Accumulate.cs.slang:73-198): Single, SingleCompensated and Double
precision and the moving-average forms, against a numpy restatement that performs the same fp32 / fp64 operations in the same order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
f32, f64 = np.float32, np.float64


from tests.parity import accumulate_pass_restatement as _reference, accumulate_test_frames  # noqa: E402


@pytest.mark.parametrize("mode,moving_average", [(0, False), (0, True), (1, False), (2, False), (2, True)])
def test_accumulate_frame_matches_the_pass(torch_cuda, mode, moving_average):
    torch = torch_cuda
    from pltfalcor_b200 import binding
    p = binding.Pltpt()
    h, w, n = 19, 33, 40
    frames = accumulate_test_frames(mode, h, w, n)
    ref = _reference(mode, frames, moving_average)
    last = torch.zeros((h, w, 4), dtype=torch.float64 if mode == 2 else torch.float32, device="cuda")
    corr = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda") if mode == 1 else None
    out = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    for k, c in enumerate(frames):
        p.accumulate_frame(mode, k, torch.from_numpy(c).cuda(), last, out, last_corr=corr, moving_average=moving_average)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), ref[k]), (mode, moving_average, k)
    if not moving_average:  # the three modes agree on the mean to their own precision
        exact = np.mean(np.stack(frames).astype(f64), axis=0)
        tol = {0: 3e-6, 1: 2e-7, 2: 1e-7}[mode]
        assert np.max(np.abs(out.cpu().numpy() - exact) / exact) <= tol
    p.close()


def test_accumulate_frame_rejects_bad_arguments(torch_cuda):
    torch = torch_cuda
    from pltfalcor_b200 import binding
    p = binding.Pltpt()
    a = torch.zeros((4, 4, 4), device="cuda")
    with pytest.raises(binding.PltptError):
        p.accumulate_frame(3, 0, a, a, a)
    with pytest.raises(binding.PltptError):
        p.accumulate_frame(1, 0, a, a, a)              # compensated mode needs the correction buffer
    with pytest.raises(binding.PltptError):
        p.accumulate_frame(1, 0, a, a, a, last_corr=a, moving_average=True)
    p.close()
