// stage_connect.cu — first half of the solve stage (reference pltpt_solve.rt.slang:401-476, NEE :245-366, evalEmissive :152-200,
// MNEE Manifolds.slangh) as a wavefront.  Per recorded-vertex index p (the work list is the sample stage's type-sorted hit list
// of bounce p, so a warp evaluates one material):
//   k_connect_a      : light sample + beam sourcing (+ delta sample of the current vertex for MNEE) + MIS weight of the unoccluded
//                      connection; posts one shadow ray
//   k_shadow_trace   : closest-hit (MNEE on: the first occluder matters) or any-hit traversal of the posted rays; a connection whose
//                      answer settles it (visible / blocked by an opaque occluder, nothing else to do at the vertex) is finished right
//                      here from the hand-over item: MIS scale, hand-over to the unwind queues
//   k_connect_b_slow : the others, densely packed: delta-occluder chain, delta sample of the current vertex, emissive hit
// and once per tile
//   k_mnee_walk / k_mnee_finish : manifold walks (Newton + line search, data-dependent trip count) of the connections that crossed
//                      delta occluders, isolated so that they cannot stall the coherent NEE work (north-star (d)).
#include "plt_integrator.cuh"
#include <cooperative_groups.h>
#include <algorithm>
namespace cg = cooperative_groups;

#ifndef PLT_MINB_CONNECT_A
#define PLT_MINB_CONNECT_A 4  // 128 registers: 16 resident warps per SM instead of 12 (-11% kernel time, profiles/r1_variants13_minblocks.log)
#endif
#ifndef PLT_MINB_SHADOW_TRACE
#define PLT_MINB_SHADOW_TRACE 1
#endif
#ifndef PLT_MINB_CONNECT_B
#define PLT_MINB_CONNECT_B 4  // 128 registers: 16 resident warps per SM instead of 12 (-19% kernel time)
#endif
#ifndef PLT_MINB_MNEE
#define PLT_MINB_MNEE 4  // 128 registers: 21.2 vs 23.7 ms on the 10 M-triangle city, 5.98 vs 5.85 ms on the grating scene (profiles/r1_variants18_mnee_occupancy.log)
#endif
#ifndef PLT_MNEE_PHASE_TIMERS
#define PLT_MNEE_PHASE_TIMERS 0  // development builds (tools/variants.py): per-round live-walk counts and phase durations of k_mnee_walk
#endif
#if PLT_MNEE_PHASE_TIMERS
__device__ unsigned long long g_mneeProfile[4 * 128];  // per round: live walks, advance ns, trace ns, rays traced
extern "C" __attribute__((visibility("default"))) int pltpt_debug_mnee_profile(unsigned long long* out) {
    return cudaMemcpyFromSymbol(out, g_mneeProfile, sizeof(g_mneeProfile)) == cudaSuccess ? 0 : 1;
}
__device__ __forceinline__ unsigned long long globalTimerNs() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif
namespace plt {
namespace {
template <bool kTex>
__global__ void __launch_bounds__(kStageThreads, PLT_MINB_CONNECT_A) k_connect_a(const __grid_constant__ FrameParams P) {
    const uint32_t n = P.levelCount[P.level];
    const uint32_t* list = P.vertexList + size_t(P.level) * P.pixelCount;
    if (blockIdx.x == 0 && threadIdx.x == 0) { P.levelCount[kTraceCountSlot] = 0; P.queueCount[QC_ShadowCursor] = 0; }  // deferred-entry count of this level's half B; shadow-trace fetch cursor
    uint32_t evals = 0, verts = 0;
    TraversalCounters tc = {0, 0};
    for (uint32_t base = blockIdx.x * kStageThreads; base < n; base += gridDim.x * kStageThreads) {
        const uint32_t i = base + threadIdx.x;
        if (i < n) { connectA<kTex>(P, i, list[i], evals, tc); ++verts; }
    }
    // warp-aggregated statistics (all lanes reach this point)
    warpAdd(P.counters + 1, verts);
    warpAdd(P.counters + 4, verts * P.level);  // sum over paths of L(L-1)/2 == sum over recorded vertices of their index
}

// onMiss (pltpt_solve.rt.slang:371-396) of every path that left the scene; one launch per tile, only for env-lit scenes.  It draws
// no random numbers and writes fixed contribution slots, so it does not have to run in vertex order.
__global__ void __launch_bounds__(kStageThreads) k_connect_miss(const __grid_constant__ FrameParams P) {
    const uint32_t n = *P.missCount;
    for (uint32_t base = blockIdx.x * kStageThreads; base < n; base += gridDim.x * kStageThreads) {
        const uint32_t i = base + threadIdx.x;
        if (i < n) { const uint32_t e = P.missList[i]; onMiss(P, e & 0x3ffffffu, int(e >> 26)); }
    }
}

// Appends entry `i` of the level's work list to the deferred list of half B (warp-aggregated; the sample stage's trace queue is
// idle during the solve stage and is reused for it; k_connect_a of this level zeroed its count).
__device__ __forceinline__ void deferConnection(const FrameParams& P, uint32_t i) {
    const uint32_t active = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(active) - 1;
    uint32_t at = 0;
    if (lane == leader) at = atomicAdd(P.levelCount + kTraceCountSlot, uint32_t(__popc(active)));
    at = __shfl_sync(active, at, leader) + __popc(active & ((1u << lane) - 1u));
    P.traceQueue[at] = i;
}
// After the shadow ray of entry `i` has been traced: settle the connection at once when the answer suffices (connectBFast: the
// first half of what used to be a separate pass over the same items), otherwise publish the hit for k_connect_b_slow and defer.
__device__ __forceinline__ void shadowAnswer(const FrameParams& P, const uint32_t* list, uint32_t i, const ConnItem& item, bool hit, const Hit& h) {
    const bool anyHit = (item.info.y & CI_ANY_HIT) != 0;
    if (!connectBFast(P, list[i], item, hit, anyHit ? 0u : h.instanceID)) return;
    P.shadowHits[i] = (hit && !anyHit) ? make_uint4(h.instanceID, h.primitiveIndex, __float_as_uint(h.u), __float_as_uint(h.v)) : make_uint4(hit ? 0u : 0xffffffffu, 0u, 0u, 0u);
    P.shadowT[i] = h.t;
    deferConnection(P, i);
}

// Tiny scenes: one ray per thread per grid-stride iteration (see k_trace_simple).
__global__ void __launch_bounds__(kStageThreads, PLT_MINB_SHADOW_TRACE) k_shadow_trace_simple(const __grid_constant__ FrameParams P) {
    const uint32_t n = P.levelCount[P.level];
    const uint32_t* list = P.vertexList + size_t(P.level) * P.pixelCount;
    TraversalCounters tc = {0, 0};
    for (uint32_t base = blockIdx.x * kStageThreads; base < n; base += gridDim.x * kStageThreads) {
        const uint32_t i = base + threadIdx.x;
        if (i >= n) continue;
        const ConnItem item = P.connItems[i];
        if (!(item.info.y & CI_TRACE)) continue;
        const float4 r0 = P.shadowRays[2 * size_t(i)], r1 = P.shadowRays[2 * size_t(i) + 1];
        const float3 o = mk3(r0.x, r0.y, r0.z), d = mk3(r1.x, r1.y, r1.z);
        Hit h;
        const bool hit = (item.info.y & CI_ANY_HIT) ? traverseBvh<true, false>(P.sc.bvhNodes, P.sc.bvhTris, o, d, r0.w, r1.w, h, reinterpret_cast<uint32_t*>(P.counters + 2), &tc)
                                                    : traverseBvh<false, false>(P.sc.bvhNodes, P.sc.bvhTris, o, d, r0.w, r1.w, h, reinterpret_cast<uint32_t*>(P.counters + 2), &tc);
        shadowAnswer(P, list, i, item, hit, h);
    }
    warpAdd(P.counters + 5, tc.nodes);
    warpAdd(P.counters + 6, tc.tris);
}

__global__ void __launch_bounds__(kStageThreads, PLT_MINB_SHADOW_TRACE) k_shadow_trace(const __grid_constant__ FrameParams P) {
    const uint32_t n = P.levelCount[P.level];
    const uint32_t* list = P.vertexList + size_t(P.level) * P.pixelCount;
    TraversalCounters tc = {0, 0};
    uint32_t* cursor = P.queueCount + QC_ShadowCursor;  // zeroed by k_connect_a of this level
    uint32_t* overflow = reinterpret_cast<uint32_t*>(P.counters + 2);
    auto fetch = [&](uint32_t i, float3& o, float3& d, float& tmin, float& tmax) {
        if (!(P.connItems[i].info.y & CI_TRACE)) return false;
        const float4 r0 = P.shadowRays[2 * size_t(i)], r1 = P.shadowRays[2 * size_t(i) + 1];
        o = mk3(r0.x, r0.y, r0.z); d = mk3(r1.x, r1.y, r1.z); tmin = r0.w; tmax = r1.w;
        return true;
    };
    auto commit = [&](uint32_t i, bool hit, const Hit& h) { shadowAnswer(P, list, i, P.connItems[i], hit, h); };
    // half A posts any-hit rays when no manifold walk can follow (MNEE off, or no bounce left for an occluder chain): uniform per launch
    if (!P.doMS || int(P.opt.maxBounces) - int(P.level) <= 0) travPersistent<true>(P.sc.bvhNodes, P.sc.bvhTris, n, cursor, overflow, &tc, fetch, commit);
    else travPersistent<false>(P.sc.bvhNodes, P.sc.bvhTris, n, cursor, overflow, &tc, fetch, commit);
    warpAdd(P.counters + 5, tc.nodes);
    warpAdd(P.counters + 6, tc.tris);
}

// Half B for the entries the shadow-ray kernel could not settle (occluder chain, delta sample of the current vertex, emissive hit),
// densely packed.
template <bool kTex>
__global__ void __launch_bounds__(kStageThreads, PLT_MINB_CONNECT_B) k_connect_b_slow(const __grid_constant__ FrameParams P) {
    const uint32_t n = P.levelCount[kTraceCountSlot];
    const uint32_t* list = P.vertexList + size_t(P.level) * P.pixelCount;
    uint32_t evals = 0;
    TraversalCounters tc = {0, 0};
    for (uint32_t base = blockIdx.x * kStageThreads; base < n; base += gridDim.x * kStageThreads) {
        const uint32_t k = base + threadIdx.x;
        if (k < n) { const uint32_t i = P.traceQueue[k]; connectB<kTex>(P, i, list[i], evals, tc); }
    }
    warpAdd(P.counters + 5, tc.nodes);
    warpAdd(P.counters + 6, tc.tris);
}

// Manifold walks (Manifolds.slangh:165-281), pass 1: ONE cooperative launch that loops over rounds on the device (grid-wide
// barriers instead of ~2 x MNEEMaxIterations host launches that would mostly be empty).  Round = { advance: consume the previous
// trial ray's answer, Newton steps, post the next trial ray (mneeWalkAdvance) | barrier | trace the posted rays | barrier }.
// Live walks are kept in two ping-pong lists, converged ones are listed for pass 2; the three list counts rotate through
// queueCount[QC_MneeRound ..+3) (zeroed per tile; the count two rounds back is re-zeroed by thread 0 while nobody reads it), the number
// of converged walks is queueCount[QC_MneeDone].
template <bool kTex>
__global__ void __launch_bounds__(kStageThreads, PLT_MINB_MNEE) k_mnee_walk(const __grid_constant__ FrameParams P) {
    cg::grid_group grid = cg::this_grid();
    const uint32_t n = min(*P.mneeCount, P.mneeCapacity);
    if (n == 0) return;  // grid-uniform
    uint32_t* counts = P.queueCount + QC_MneeRound;
    uint32_t* doneCount = P.queueCount + QC_MneeDone;
    uint32_t* doneList = P.mneeLists + 2 * size_t(P.mneeCapacity);
    const uint32_t tid = blockIdx.x * kStageThreads + threadIdx.x, stride = gridDim.x * kStageThreads;
    const int lane = threadIdx.x & 31;
    TraversalCounters tc = {0, 0};
    unsigned long long iterations = 0;
    for (uint32_t round = 0;; ++round) {
        const uint32_t* cur = P.mneeLists + size_t(round & 1u) * P.mneeCapacity;
        uint32_t* next = P.mneeLists + size_t((round + 1u) & 1u) * P.mneeCapacity;
        const uint32_t nCur = round == 0 ? n : counts[round % 3u];
        uint32_t* nextCount = counts + (round + 1u) % 3u;
        if (tid == 0) { counts[(round + 2u) % 3u] = 0; P.queueCount[QC_MneeTraceCursor + (round + 2u) % 3u] = 0; }
#if PLT_MNEE_PHASE_TIMERS
        unsigned long long t0 = 0;
        if (tid == 0) t0 = globalTimerNs();
#endif
        for (uint32_t base = blockIdx.x * kStageThreads; base < nCur; base += stride) {  // ---- advance
            const uint32_t k = base + threadIdx.x;
            int r = WALK_FAILED;
            uint32_t item = 0;
            const bool valid = k < nCur;
            if (valid) {
                item = round == 0 ? k : cur[k];
                r = mneeWalkAdvance<kTex>(P, P.mneeItems[item], P.mneeWalks[item], round == 0);
                if (r != WALK_CONTINUE) {
                    iterations += P.mneeWalks[item].iterations;
                    if (P.debugView == DV_MneeIterations) {  // pltpt_solve.rt.slang:318-322
                        float4 ev = mk4(0.f);
                        debugEvent(&ev, ccInferno(float(P.mneeWalks[item].iterations) / float(P.opt.MNEEMaxIterations)));
                        writeDebug(P, P.mneeItems[item].pixel, P.mneeItems[item].bounceIdx, 0, ev);
                    }
                    if (r == WALK_FAILED) writeContribution(P, P.mneeItems[item].pixel, P.mneeItems[item].bounceIdx, 0, mk3(0.f));
                }
            }
            // warp-aggregated appends to the two lists
            const uint32_t go = __ballot_sync(0xffffffffu, valid && r == WALK_CONTINUE), ok = __ballot_sync(0xffffffffu, valid && r == WALK_CONVERGED);
            uint32_t b0 = 0, b1 = 0;
            if (lane == 0) { if (go) b0 = atomicAdd(nextCount, uint32_t(__popc(go))); if (ok) b1 = atomicAdd(doneCount, uint32_t(__popc(ok))); }
            b0 = __shfl_sync(0xffffffffu, b0, 0); b1 = __shfl_sync(0xffffffffu, b1, 0);
            const uint32_t below = (1u << lane) - 1u;
            if (valid && r == WALK_CONTINUE) next[b0 + __popc(go & below)] = item;
            if (valid && r == WALK_CONVERGED) doneList[b1 + __popc(ok & below)] = item;
        }
        grid.sync();
        const uint32_t nNext = *nextCount;
#if PLT_MNEE_PHASE_TIMERS
        unsigned long long t1 = 0;
        if (tid == 0 && round < 128) { t1 = globalTimerNs(); g_mneeProfile[4 * round] += nCur; g_mneeProfile[4 * round + 1] += t1 - t0; g_mneeProfile[4 * round + 3] += nNext; }
#endif
        if (nNext == 0) break;  // grid-uniform: every thread reads the same count after the barrier
        // ---- trace the posted trial rays (persistent lanes; the fetch cursors rotate through queueCount[QC_MneeTraceCursor ..+3) like the counts)
        travPersistent<false>(P.sc.bvhNodes, P.sc.bvhTris, nNext, P.queueCount + QC_MneeTraceCursor + round % 3u, reinterpret_cast<uint32_t*>(P.counters + 2), &tc,
            [&](uint32_t k, float3& o, float3& d, float& tmin, float& tmax) {
                const MneeWalk& w = P.mneeWalks[next[k]];
                const float4 r0 = w.ray0, r1 = w.ray1;
                o = mk3(r0.x, r0.y, r0.z); d = mk3(r1.x, r1.y, r1.z); tmin = r0.w; tmax = r1.w;
                return true;
            },
            [&](uint32_t k, bool hit, const Hit& h) {
                MneeWalk& w = P.mneeWalks[next[k]];
                w.hit = hit ? make_uint4(h.instanceID, h.primitiveIndex, __float_as_uint(h.u), __float_as_uint(h.v)) : make_uint4(0xffffffffu, 0u, 0u, 0u);
                w.t = h.t;
            });
        grid.sync();
#if PLT_MNEE_PHASE_TIMERS
        if (tid == 0 && round < 128) g_mneeProfile[4 * round + 2] += globalTimerNs() - t1;
#endif
    }
    warpAdd(P.counters + 5, tc.nodes);
    warpAdd(P.counters + 6, tc.tris);
    if (iterations) atomicAdd(P.counters + 3, iterations);
}
// Pass 2: converged walks, densely packed: final visibility, replay through the specular chain, MIS, hand-over to the unwind queues.
template <bool kTex>
__global__ void __launch_bounds__(kStageThreads, PLT_MINB_MNEE) k_mnee_finish(const __grid_constant__ FrameParams P) {
    const uint32_t n = P.queueCount[QC_MneeDone];
    const uint32_t* doneList = P.mneeLists + 2 * size_t(P.mneeCapacity);
    uint32_t evals = 0;
    TraversalCounters tc = {0, 0};
    for (uint32_t base = blockIdx.x * kStageThreads; base < n; base += gridDim.x * kStageThreads) {
        const uint32_t i = base + threadIdx.x;
        if (i < n) mneeFinish<kTex>(P, P.mneeItems[doneList[i]], evals, tc);
    }
    warpAdd(P.counters + 0, evals);
    warpAdd(P.counters + 5, tc.nodes);
    warpAdd(P.counters + 6, tc.tris);
}
}  // namespace
// cooperative launch: every CTA must be resident, so the grid is what the device can hold of this kernel (per device: a process may
// hold handles on several GPUs)
uint32_t mneeCooperativeGrid(int device) {
    int perSm = 0, sms = 0, cur = 0;
    cudaGetDevice(&cur);
    if (cur != device) cudaSetDevice(device);
    int perSmTex = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_mnee_walk<false>, kStageThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmTex, k_mnee_walk<true>, kStageThreads, 0);
    perSm = std::min(perSm, perSmTex);  // one grid size for both instantiations (textured / untextured scenes)
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (cur != device) cudaSetDevice(cur);
    return uint32_t(std::max(perSm, 1) * std::max(sms, 1));
}
void launchMnee(const FrameParams& P, uint32_t grid, cudaStream_t s) {
    FrameParams Pc = P;
    void* args[] = {&Pc};
    if (P.sc.tex.blobs) {
        cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&k_mnee_walk<true>), dim3(P.mneeGrid), dim3(kStageThreads), args, 0, s);
        k_mnee_finish<true><<<grid, kStageThreads, 0, s>>>(P);
    } else {
        cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&k_mnee_walk<false>), dim3(P.mneeGrid), dim3(kStageThreads), args, 0, s);
        k_mnee_finish<false><<<grid, kStageThreads, 0, s>>>(P);
    }
}
void launchConnectMiss(const FrameParams& P, uint32_t grid, cudaStream_t s) { k_connect_miss<<<grid, kStageThreads, 0, s>>>(P); }
// (the instantiation without the textured-material path for scenes that have no textured material: plt_scene.cuh loadMaterialAt)
void launchConnectA(const FrameParams& P, uint32_t grid, cudaStream_t s) {
    if (P.sc.tex.blobs) k_connect_a<true><<<grid, kStageThreads, 0, s>>>(P);
    else k_connect_a<false><<<grid, kStageThreads, 0, s>>>(P);
}
void launchShadowTrace(const FrameParams& P, uint32_t grid, cudaStream_t s) {
    if (P.tinyScene) k_shadow_trace_simple<<<grid, kStageThreads, 0, s>>>(P);
    else k_shadow_trace<<<grid, kStageThreads, 0, s>>>(P);
}
void launchConnectB(const FrameParams& P, uint32_t grid, cudaStream_t s) {
    if (P.sc.tex.blobs) k_connect_b_slow<true><<<grid, kStageThreads, 0, s>>>(P);
    else k_connect_b_slow<false><<<grid, kStageThreads, 0, s>>>(P);
}
}  // namespace plt
