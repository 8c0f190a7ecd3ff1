"""Partitioning of a render across the GPUs of one box (north-star (e)): by sample index and by image tile.

Every random number of the PLTPT path is a pure function of (pixel, frameCount) (pltpt_sample.rt.slang:154,
pltpt_solve.rt.slang:410-415) and the pass carries no state from frame to frame (PLTPT.cpp:498).  A render of `frames` samples per
pixel can therefore be split
  * by SAMPLE INDEX: rank r of N renders the frame indices {first + r, first + r + N, ...} for all pixels, with the camera jitter the
    single-GPU run would use for that index (pltpt_render_accumulate(first + r, N, count)), or
  * by IMAGE TILE: every rank renders all frames of its band of rows (pltpt_render_accumulate_rows; bands start on multiples of 4 rows,
    the height of the 8x4 primary-ray tiles),
and in both cases the per-rank fp64 accumulators add up to the single-GPU accumulator (disjoint samples, or disjoint pixels with zeros
elsewhere), so one sum -- an NCCL reduce between processes, a peer-memory kernel inside one process -- finishes the image.  No collective
on the data path.  bench.py, the gloo test (tests/test_partition.py) and libpltpt's pltpt_render_accumulate_multi use the same rules."""


def frames_of_rank(rank, world, first, frames_per_rank):
    """(first_frame, frame_step, frame_count) arguments of pltpt_render_accumulate for this rank: `frames_per_rank` frames each
    (weak scaling: the job grows with the number of ranks)."""
    if not (0 <= rank < world) or frames_per_rank < 0:
        raise ValueError("invalid rank/world/frames_per_rank")
    return first + rank, world, frames_per_rank


def frames_of_rank_strong(rank, world, first, total_frames):
    """The same for a FIXED job of `total_frames` frames (strong scaling): rank r takes first + r, + world, ... below first + total."""
    if not (0 <= rank < world) or total_frames < 0:
        raise ValueError("invalid rank/world/total_frames")
    return first + rank, world, total_frames // world + (1 if rank < total_frames % world else 0)


def frame_list(rank, world, first, frames_per_rank):
    f0, step, n = frames_of_rank(rank, world, first, frames_per_rank)
    return [f0 + i * step for i in range(n)]


def row_band(rank, world, height):
    """[row_begin, row_end) of this rank's band of image rows; bands start on multiples of 4 (a band may be empty for tiny images)."""
    if not (0 <= rank < world) or height <= 0:
        raise ValueError("invalid rank/world/height")
    band = ((height + world - 1) // world + 3) & ~3
    return min(height, rank * band), min(height, (rank + 1) * band)


def reduce_accumulators(dist, accum, dst=0):
    """Sums the per-rank RGBA64F accumulators onto rank `dst` (one reduce; NCCL on the GPU box, gloo in the CPU tests)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
