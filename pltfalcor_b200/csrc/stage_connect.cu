// stage_connect.cu — first half of the solve stage (reference pltpt_solve.rt.slang:401-476, NEE :245-366, evalEmissive :152-200,
// MNEE Manifolds.slangh).  One launch per recorded-vertex index p: every path that has a hit at vertex p samples the light,
// traces the shadow ray / manifold walk, applies MIS and emits one connection job per surviving light connection.  The work list
// is the sample stage's type-sorted hit list of bounce p, so a warp evaluates one material's pdf / delta sampling.
#include "plt_integrator.cuh"

namespace plt {
namespace {
__global__ void __launch_bounds__(kStageThreads) k_connect(const __grid_constant__ FrameParams P) {
    const uint32_t n = P.levelCount[P.level];
    const uint32_t* list = P.vertexList + size_t(P.level) * P.pixelCount;
    uint32_t evals = 0, verts = 0;
    TraversalCounters tc = {0, 0};
    for (uint32_t base = blockIdx.x * kStageThreads; base < n; base += gridDim.x * kStageThreads) {
        const uint32_t i = base + threadIdx.x;
        if (i < n) {
            uint32_t e = 0;
            TraversalCounters t = {0, 0};
            connectVertex(P, list[i], e, t);
            evals += e; tc.nodes += t.nodes; tc.tris += t.tris; ++verts;
        }
    }
    // warp-aggregated statistics (all lanes reach this point)
    warpAdd(P.counters + 0, evals);
    warpAdd(P.counters + 1, verts);
    warpAdd(P.counters + 4, verts * P.level);  // sum over paths of L(L-1)/2 == sum over recorded vertices of their index
    warpAdd(P.counters + 5, tc.nodes);
    warpAdd(P.counters + 6, tc.tris);
}
}  // namespace
void launchConnect(const FrameParams& P, uint32_t grid, cudaStream_t s) { k_connect<<<grid, kStageThreads, 0, s>>>(P); }
}  // namespace plt
