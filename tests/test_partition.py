"""Multi-GPU host logic on CPU: two gloo ranks partition the frames of a render by sample index, each sums its own frames
(rendered by the CPU oracle at a tiny size), one reduce puts the sum on rank 0, and it equals the single-process sum."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, FRAMES_PER_RANK, FIRST = 24, 14, 3, 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _render_sum(frames):
    from oracle import binding as ob
    from pltfalcor_b200 import layouts as L, scene
    buffers, cam = scene.cornell_box(aspect=W / H)
    osc = ob.OracleScene(buffers)
    opts = L.make_options(maxBounces=4)
    acc = np.zeros((H, W, 4), np.float64)
    for f in frames:
        img = osc.render_frame(opts, scene.with_jitter(cam, f, W, H), f, W, H, threads=1)["color"]
        acc[..., :3] += img[..., :3].astype(np.float64)
        acc[..., 3] += 1.0
    return acc


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from pltfalcor_b200 import partition
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    frames = partition.frame_list(rank, world, FIRST, FRAMES_PER_RANK)
    gathered = [None] * world
    dist.all_gather_object(gathered, frames)
    acc = torch.from_numpy(_render_sum(frames))
    partition.reduce_accumulators(dist, acc, dst=0)
    if rank == 0:
        np.save(out, acc.numpy())
        np.save(out + ".frames.npy", np.array(gathered, dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


def test_frames_of_rank_is_a_disjoint_cover():
    from pltfalcor_b200 import partition
    for world in (1, 2, 4, 8):
        allf = sorted(f for r in range(world) for f in partition.frame_list(r, world, 7, 5))
        assert allf == list(range(7, 7 + 5 * world))
    with pytest.raises(ValueError):
        partition.frames_of_rank(2, 2, 0, 1)


def test_strong_scaling_and_row_band_partitions_cover_the_job_exactly():
    from pltfalcor_b200 import partition
    for world in (1, 2, 3, 4, 8):
        for total in (0, 1, 7, 64, 65):
            frames = []
            for r in range(world):
                f0, step, n = partition.frames_of_rank_strong(r, world, 11, total)
                frames += [f0 + i * step for i in range(n)]
            assert sorted(frames) == list(range(11, 11 + total)), (world, total)
        for height in (4, 27, 1080, 2160):
            rows = []
            for r in range(world):
                a, b = partition.row_band(r, world, height)
                assert (a % 4 == 0 or a == b) and a <= b <= height   # (an empty band past the image end may start anywhere)
                rows += list(range(a, b))
            assert rows == list(range(height)), (world, height)


def test_two_rank_partition_sums_to_the_single_rank_image(tmp_path, oracle):
    import torch.multiprocessing as mp
    world, port, out = 2, _free_port(), str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = np.load(out)
    frames = np.load(out + ".frames.npy")
    assert sorted(frames.reshape(-1).tolist()) == list(range(FIRST, FIRST + world * FRAMES_PER_RANK))
    ref = _render_sum(range(FIRST, FIRST + world * FRAMES_PER_RANK))
    assert np.all(got[..., 3] == world * FRAMES_PER_RANK)
    # identical samples, different fp64 summation order
    np.testing.assert_allclose(got[..., :3], ref[..., :3], rtol=1e-12, atol=1e-15)
    assert float(ref[..., :3].sum()) > 0
