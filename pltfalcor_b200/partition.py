"""Sample-index partitioning of a render across the GPUs of one box.

Every random number of the PLTPT path is a pure function of (pixel, frameCount) (pltpt_sample.rt.slang:154,
pltpt_solve.rt.slang:410-415) and the pass carries no state from frame to frame (PLTPT.cpp:498), so rank r of N renders the
frame indices {first + r, first + r + N, ...} for all pixels, with the camera jitter the single-GPU run would use for that
index, and the per-rank fp64 accumulation buffers are summed once at the end.  No collective on the data path."""


def frames_of_rank(rank, world, first, frames_per_rank):
    """(first_frame, frame_step, frame_count) arguments of pltpt_render_accumulate for this rank."""
    if not (0 <= rank < world) or frames_per_rank < 0:
        raise ValueError("invalid rank/world/frames_per_rank")
    return first + rank, world, frames_per_rank


def frame_list(rank, world, first, frames_per_rank):
    f0, step, n = frames_of_rank(rank, world, first, frames_per_rank)
    return [f0 + i * step for i in range(n)]


def reduce_accumulators(dist, accum, dst=0):
    """Sums the per-rank RGBA64F accumulators onto rank `dst` (one reduce; NCCL on the GPU box, gloo in the CPU tests)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
