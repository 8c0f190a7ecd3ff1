// stage_sample.cu — sample stage of PLTPT::execute (reference pltpt_sample.rt.slang:146-185): one backward path per pixel;
// records every bounce (SoA planes) and caches the vertex shading data for the solve stage.
#include "plt_integrator.cuh"

namespace plt {
namespace {
__global__ void __launch_bounds__(kStageThreads) k_sample_stage(const __grid_constant__ FrameParams P) {
    uint32_t px, py;
    TraversalCounters tc = {0, 0};
    if (pixelOfThread(P, px, py)) tc = samplePixel(P, px, py);
    warpAdd(P.counters + 5, tc.nodes);
    warpAdd(P.counters + 6, tc.tris);
}

// Re-assembles the reference's 40 B AoS bounce records (pixel-major, pltpt.slangh:132-136) from the SoA planes.
__global__ void k_gather_bounces(FrameParams P, PltptBounce* __restrict__ out, uint32_t records) {
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel >= P.pixelCount) return;
    for (uint32_t p = 0; p < records; ++p) {
        const size_t o = size_t(p) * P.pixelCount + pixel;
        const uint4 h = P.bHit[o];
        const float4 d = P.bDir[o];
        const float2 q = P.bPdf[o];
        PltptBounce b;
        b.hit[0] = h.x; b.hit[1] = h.y; b.hit[2] = h.z; b.hit[3] = h.w;
        b.encodedWi[0] = d.x; b.encodedWi[1] = d.y;
        b.sampledBsdfLobe = __float_as_uint(d.z); b.flags = __float_as_uint(d.w);
        b.pathContribution = q.x; b.lastNonDeltaPdf = q.y;
        out[size_t(P.outOffset + pixel) * records + p] = b;
    }
}
}  // namespace

void launchSampleStage(const FrameParams& P, uint32_t blocks, cudaStream_t s) { k_sample_stage<<<blocks, kStageThreads, 0, s>>>(P); }
void launchGatherBounces(const FrameParams& P, PltptBounce* out, uint32_t records, cudaStream_t s) {
    k_gather_bounces<<<(P.pixelCount + 127) / 128, 128, 0, s>>>(P, out, records);
}
}  // namespace plt
