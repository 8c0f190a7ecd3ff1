// stage_sample.cu — sample stage of PLTPT::execute (reference pltpt_sample.rt.slang:146-185) as a wavefront:
//   k_raygen            : seeds the path, hero wavelength, pinhole ray (one thread per pixel)
//   k_trace   (bounce b): closest hit of every live path (persistent grid over the trace queue); misses end here; hits are
//                         appended to the shade histogram keyed by the MaterialType that was hit
//   k_shade   (bounce b): walks the hits in MaterialType order (128 same-type paths per CTA => one BSDF per warp): shading setup,
//                         Russian roulette, bsdf.sample, guide buffers, bounce record + vertex cache, next ray -> trace queue
// The type-sorted hit list of bounce b is kept (FrameParams::vertexList): it is exactly the work list of the solve stage's
// connect kernel for vertex b.
#include "plt_integrator.cuh"

#ifndef PLT_MINB_SHADE
#define PLT_MINB_SHADE 4  // 128 registers (128 B of spills): 16 resident warps per SM instead of 12 (-18% kernel time)
#endif
#ifndef PLT_MINB_TRACE
#define PLT_MINB_TRACE 1
#endif
namespace plt {
namespace {
__global__ void __launch_bounds__(256) k_raygen(const __grid_constant__ FrameParams P) {
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel < P.pixelCount) raygenPixel(P, pixel);
}

// Tiny scenes (a handful of BVH nodes: every ray of a warp walks the same two or three nodes) are faster without the refill
// logic: one ray per thread per grid-stride iteration (Cornell box 1.27 vs 1.61 ms per frame; profiles/r1_kernel_times_scenes*.log).
__global__ void __launch_bounds__(kStageThreads, PLT_MINB_TRACE) k_trace_simple(const __grid_constant__ FrameParams P) {
    const bool primary = P.level == 0;
    const uint32_t tilesX = (P.width + 7) / 8, tilesY = (P.tileRows + 3) / 4;
    const uint32_t n = primary ? tilesX * tilesY * 32u : P.levelCount[kTraceCountSlot];
    TraversalCounters tc = {0, 0};
    for (uint32_t base = blockIdx.x * kStageThreads; base < n; base += gridDim.x * kStageThreads) {
        const uint32_t i = base + threadIdx.x;
        uint32_t pixel = 0;
        const bool live = i < n && (primary ? primaryPixel(P, i, pixel) : (pixel = P.traceQueue[i], true));
        if (live) {
            const float4 o = P.psOrigin[pixel], d = P.psDir[pixel];
            Hit h;
            const bool hit = traverseBvh<false, false>(P.sc.bvhNodes, P.sc.bvhTris, mk3(o.x, o.y, o.z), mk3(d.x, d.y, d.z), 0.f, kFltMax, h, reinterpret_cast<uint32_t*>(P.counters + 2), &tc);
            const uint32_t type = tracePathCommit(P, pixel, hit, h);
            if (type != 0u) enqueueJob<false>(P, type, pixel);
        }
    }
    warpAdd(P.counters + 5, tc.nodes);
    warpAdd(P.counters + 6, tc.tris);
}

__global__ void __launch_bounds__(kStageThreads, PLT_MINB_TRACE) k_trace(const __grid_constant__ FrameParams P) {
    const bool primary = P.level == 0;
    // bounce 0 walks the implicit queue of all pixels (8x4 tiles per warp, padded), later bounces the queue the shade kernel filled
    const uint32_t tilesX = (P.width + 7) / 8, tilesY = (P.tileRows + 3) / 4;
    const uint32_t n = primary ? tilesX * tilesY * 32u : P.levelCount[kTraceCountSlot];
    TraversalCounters tc = {0, 0};
    auto pixelOf = [&](uint32_t i, uint32_t& pixel) { return primary ? primaryPixel(P, i, pixel) : (pixel = P.traceQueue[i], true); };
    // persistent lanes: a lane whose ray is done takes the next queue entry (cursor queueCount[QC_TraceCursor], re-zeroed by k_flip_queue)
    travPersistent<false>(P.sc.bvhNodes, P.sc.bvhTris, n, P.queueCount + QC_TraceCursor, reinterpret_cast<uint32_t*>(P.counters + 2), &tc,
        [&](uint32_t i, float3& o, float3& d, float& tmin, float& tmax) {
            uint32_t pixel;
            if (!pixelOf(i, pixel)) return false;
            const float4 ro = P.psOrigin[pixel], rd = P.psDir[pixel];
            o = mk3(ro.x, ro.y, ro.z); d = mk3(rd.x, rd.y, rd.z); tmin = 0.f; tmax = kFltMax;
            return true;
        },
        [&](uint32_t i, bool hit, const Hit& h) {
            uint32_t pixel;
            pixelOf(i, pixel);
            const uint32_t type = tracePathCommit(P, pixel, hit, h);
            if (type != 0u) enqueueJob<false>(P, type, pixel);
        });
    warpAdd(P.counters + 5, tc.nodes);
    warpAdd(P.counters + 6, tc.tris);
}

template <bool kTex>
__global__ void __launch_bounds__(kStageThreads, PLT_MINB_SHADE) k_shade(const __grid_constant__ FrameParams P) {
    const uint32_t n = P.levelCount[P.level];
    const uint32_t* list = P.vertexList + size_t(P.level) * P.pixelCount;
    TraversalCounters tc = {0, 0};
    for (uint32_t base = blockIdx.x * kStageThreads; base < n; base += gridDim.x * kStageThreads) {
        const uint32_t i = base + threadIdx.x;
        if (i < n) {
            const uint32_t pixel = list[i];
            if (shadePath<kTex>(P, pixel, tc)) {
                // append to the trace queue of the next bounce (warp-aggregated)
                const uint32_t active = __activemask();
                const int lane = threadIdx.x & 31, leader = __ffs(active) - 1;
                uint32_t slot = 0;
                if (lane == leader) slot = atomicAdd(P.levelCount + kTraceCountSlot, uint32_t(__popc(active)));
                slot = __shfl_sync(active, slot, leader) + __popc(active & ((1u << lane) - 1u));
                P.traceQueue[slot] = pixel;
            }
        }
    }
    warpAdd(P.counters + 5, tc.nodes);
    warpAdd(P.counters + 6, tc.tris);
}

// Re-assembles the reference's 40 B AoS bounce records (pixel-major, pltpt.slangh:132-136) from the SoA planes.
__global__ void k_gather_bounces(FrameParams P, PltptBounce* __restrict__ out, uint32_t records) {
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel >= P.pixelCount) return;
    bool ended = false;
    for (uint32_t p = 0; p < records; ++p) {
        const size_t o = size_t(p) * P.pixelCount + pixel;
        PltptBounce b = {};
        if (!ended) {  // records after the terminating one are unspecified in the planes: dump them as zero like the reference buffer
            const uint4 h = P.bHit[o];
            const float4 d = P.bDir[o];
            const float2 q = P.bPdf[o];
            b.hit[0] = h.x; b.hit[1] = h.y; b.hit[2] = h.z; b.hit[3] = h.w;
            b.encodedWi[0] = d.x; b.encodedWi[1] = d.y;
            b.sampledBsdfLobe = __float_as_uint(d.z); b.flags = __float_as_uint(d.w);
            b.pathContribution = q.x; b.lastNonDeltaPdf = q.y;
            ended = (b.flags & (RF_Miss | RF_Terminated)) != 0;
        }
        out[size_t(P.outOffset + pixel) * records + p] = b;
    }
}
}  // namespace

void launchRaygen(const FrameParams& P, cudaStream_t s) { k_raygen<<<(P.pixelCount + 255) / 256, 256, 0, s>>>(P); }
void launchTrace(const FrameParams& P, uint32_t grid, cudaStream_t s) {
    if (P.tinyScene) k_trace_simple<<<grid, kStageThreads, 0, s>>>(P);
    else k_trace<<<grid, kStageThreads, 0, s>>>(P);
}
void launchShade(const FrameParams& P, uint32_t grid, cudaStream_t s) {
    if (P.sc.tex.blobs) k_shade<true><<<grid, kStageThreads, 0, s>>>(P);
    else k_shade<false><<<grid, kStageThreads, 0, s>>>(P);
}
void launchGatherBounces(const FrameParams& P, PltptBounce* out, uint32_t records, cudaStream_t s) {
    k_gather_bounces<<<(P.pixelCount + 127) / 128, 128, 0, s>>>(P, out, records);
}
}  // namespace plt
