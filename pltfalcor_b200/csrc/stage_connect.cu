// stage_connect.cu — first half of the solve stage (reference pltpt_solve.rt.slang:401-476, NEE :245-366, evalEmissive :152-200,
// MNEE Manifolds.slangh): per path, in recorded-vertex order, sample the light, trace the shadow ray / manifold walk, apply MIS and
// emit one connection job per surviving light connection.
#include "plt_integrator.cuh"

namespace plt {
namespace {
__global__ void __launch_bounds__(kStageThreads) k_connect_stage(const __grid_constant__ FrameParams P) {
    uint32_t px, py;
    uint32_t evals = 0, verts = 0;
    TraversalCounters tc = {0, 0};
    if (pixelOfThread(P, px, py)) verts = connectPixel(P, px, py, evals, tc);
    // warp-aggregated statistics (all lanes reach this point)
    warpAdd(P.counters + 0, evals);
    warpAdd(P.counters + 1, verts);
    warpAdd(P.counters + 4, verts * (verts - (verts ? 1u : 0u)) / 2u);
    warpAdd(P.counters + 5, tc.nodes);
    warpAdd(P.counters + 6, tc.tris);
}
}  // namespace
void launchConnectStage(const FrameParams& P, uint32_t blocks, cudaStream_t s) { k_connect_stage<<<blocks, kStageThreads, 0, s>>>(P); }
}  // namespace plt
