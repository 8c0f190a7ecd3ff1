// stage_unwind.cu — second half of the solve stage: the forward replay (reference forwardUnwind, pltpt_solve.rt.slang:48-79,
// evalNEE :219-243, measure :96-147) as a wavefront.  One launch of k_unwind_step advances every live connection job by ONE
// recorded vertex; jobs are walked in MaterialType order, 128 same-type jobs per CTA, so a warp executes one BSDF.
#include "plt_integrator.cuh"

#ifndef PLT_MINB_UNWIND_STEP
#define PLT_MINB_UNWIND_STEP 4  // 128 registers (196 B of spills): 16 resident warps per SM instead of 8 (profiles/r1_variants*.log: -20% kernel time)
#endif
#ifndef PLT_UNWIND_PREFETCH
#define PLT_UNWIND_PREFETCH 0  // measured on B200 (profiles/r1_variants6_prefetch.log): 1.92 ms with, 1.71 ms without -- left off
#endif
namespace plt {
namespace {
__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// Counting-sort scatter of the pending list into material-type order (dst).  Runs between a producer and a consumer launch.
// Each CTA takes chunks of kSortChunk entries: a shared-memory histogram of the chunk, ONE global cursor bump per type per chunk
// (the per-warp version serialised ~65k atomics on the cursor of the dominant type), then the scatter with shared-memory cursors.
constexpr int kSortThreads = 256, kSortItems = 8, kSortChunk = kSortThreads * kSortItems;
__global__ void __launch_bounds__(kSortThreads) k_sort_queue(const __grid_constant__ FrameParams P, SortView view) {
    __shared__ uint32_t offs[kQueueTypes], hist[kQueueTypes], base[kQueueTypes];
    uint32_t* qc = P.queueCount;
    const uint32_t total = qc[QC_Total];
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int t = 0; t < kQueueTypes; ++t) { offs[t] = run; run += qc[QC_Hist + t]; }
    }
    const int lane = threadIdx.x & 31;
    for (uint32_t chunk = blockIdx.x * kSortChunk; chunk < total; chunk += gridDim.x * kSortChunk) {
        if (threadIdx.x < kQueueTypes) hist[threadIdx.x] = 0;
        __syncthreads();
        uint32_t type[kSortItems], rank[kSortItems];
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            const uint32_t i = chunk + k * kSortThreads + threadIdx.x;
            type[k] = i < total ? P.pendingType[i] : 0xffu;
            // warp-aggregated rank within the chunk's per-type run
            const uint32_t peers = __match_any_sync(0xffffffffu, type[k]);
            uint32_t r = 0;
            if (type[k] != 0xffu && lane == __ffs(peers) - 1) r = atomicAdd(&hist[type[k]], uint32_t(__popc(peers)));
            rank[k] = __shfl_sync(0xffffffffu, r, __ffs(peers) - 1) + __popc(peers & ((1u << lane) - 1u));
        }
        __syncthreads();
        if (threadIdx.x < kQueueTypes) base[threadIdx.x] = hist[threadIdx.x] ? atomicAdd(qc + QC_Cursor + threadIdx.x, hist[threadIdx.x]) : 0u;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            const uint32_t i = chunk + k * kSortThreads + threadIdx.x;
            if (type[k] == 0xffu) continue;
            const uint32_t pos = offs[type[k]] + base[type[k]] + rank[k];
            view.dst[pos] = P.pending[i];
            if (view.dstAux) view.dstAux[pos] = P.pendingAux[i];
        }
        __syncthreads();
    }
}
// Publishes the sorted count and resets histogram / cursors / pending count (single thread; stream-ordered after k_sort_queue).
__global__ void k_flip_queue(const __grid_constant__ FrameParams P, SortView view, bool resetTraceQueue) {
    uint32_t* qc = P.queueCount;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        *view.dstCount = qc[QC_Total];
        for (uint32_t t = QC_Hist; t < QC_SortedCount; ++t) qc[t] = 0;  // histogram, total, cursors
        if (resetTraceQueue) { P.levelCount[kTraceCountSlot] = 0; qc[QC_TraceCursor] = 0; }  // trace queue of the next bounce and the trace kernel's fetch cursor
    }
}

// (Round 2 also measured the step as four kernels specialised by {diffuse, other material} x {continues, ends in measure()}: the
// diffuse kernels need 72-80 registers instead of 128, but the four launches per step with their tails cost more than the occupancy
// gains when the kernels run alone -- 1.54 -> 1.83 ms on the Cornell box, 1.53 -> 3.58 ms on the ten-material city -- and only 2.8 % of
// the frame time came back with two frames in flight; rejected, profiles/r2e_variants.log.)
#ifndef PLT_UNWIND_STAGE
#define PLT_UNWIND_STAGE 0  // measured on B200 (profiles/r2a_variants.log): 1.78 ms with the staging, 1.57 ms without -- left off
#endif
#if PLT_UNWIND_STAGE
// Asynchronous global -> shared copies (LDGSTS): the gather of a thread's NEXT item runs under the BSDF arithmetic of its current one
// and costs no registers while in flight (the kernel sits at the 128-register limit that gives 16 warps per SM).
__device__ __forceinline__ void cpAsync16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(smem))), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void stageUnwindItem(const FrameParams& P, UnwindItem* slot, uint32_t id, uint32_t aux) {
    const uint32_t pixel = aux & 0x3ffffffu, q = aux >> 26;
    const size_t o = size_t(q) * P.pixelCount + pixel;
    const float4* J = P.job + size_t(id) * kJobPlanes;
    const float4* V = P.vc + 4 * o;
#pragma unroll
    for (int k = 0; k < kJobPlanes; ++k) cpAsync16(&slot->job[k], J + k);
#pragma unroll
    for (int k = 0; k < 4; ++k) cpAsync16(&slot->vtx[k], V + k);
    cpAsync16(&slot->dir, P.bDir + o);
    cpAsyncCommit();
}
// Each thread owns one 240 B slot (stride 60 words: a quarter-warp's 16 B accesses fall into distinct banks).  The slot is read
// into registers, then refilled for the thread's next item (same thread, program order: no barrier), then the step is computed.
template <bool kTex>
__global__ void __launch_bounds__(kStageThreads, PLT_MINB_UNWIND_STEP) k_unwind_step(const __grid_constant__ FrameParams P) {
    __shared__ UnwindItem slots[kStageThreads];
    const uint32_t n = P.queueCount[QC_SortedCount];
    uint32_t evals = 0;
    const uint32_t stride = gridDim.x * kStageThreads;
    uint32_t i = blockIdx.x * kStageThreads + threadIdx.x;
    UnwindItem* slot = &slots[threadIdx.x];
    uint32_t id = 0, aux = 0, nid = 0, naux = 0;
    if (i < n) { id = P.sorted[i]; aux = P.sortedAux[i]; stageUnwindItem(P, slot, id, aux); }
    if (i + stride < n) { nid = P.sorted[i + stride]; naux = P.sortedAux[i + stride]; }
    for (; i < n; i += stride) {
        cpAsyncWaitAll();
        uint32_t materialID;                            // staged record -> registers (15 LDS.128)
        const VertexData v = readVertexCache(slot->vtx, materialID);
        const float4 d = slot->dir;
        Beam beam;
        JobHeader h;
        loadJob(slot->job, beam, h);
        const uint32_t j = i + stride;
        uint32_t nnid = 0, nnaux = 0;
        if (j < n) {
            stageUnwindItem(P, slot, nid, naux);        // refill for the next iteration: in flight during this item's arithmetic
            const uint32_t j2 = j + stride;
            if (j2 < n) { nnid = P.sorted[j2]; nnaux = P.sortedAux[j2]; }
        }
        uint32_t nextType = 0, nextAux = 0;
        if (unwindStepLoaded<kTex>(P, id, v, materialID, d, beam, h, evals, nextType, nextAux)) enqueueJob(P, nextType, id, nextAux);
        id = nid; aux = naux; nid = nnid; naux = nnaux;
    }
    warpAdd(P.counters + 0, evals);
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(P.counters + 7, (unsigned long long)n);
}
#else
template <bool kTex>
__global__ void __launch_bounds__(kStageThreads, PLT_MINB_UNWIND_STEP) k_unwind_step(const __grid_constant__ FrameParams P) {
    const uint32_t n = P.queueCount[QC_SortedCount];
    uint32_t evals = 0;
    const uint32_t stride = gridDim.x * kStageThreads;
    uint32_t i = blockIdx.x * kStageThreads + threadIdx.x;
    uint32_t id = 0, aux = 0;
    if (i < n) { id = P.sorted[i]; aux = P.sortedAux[i]; }
    for (; i < n; i += stride) {  // (tail threads of the last CTA-row simply skip: no block-wide barrier in this kernel)
        // optional software prefetch of this thread's NEXT item (job 160 B, vertex 64 B, bounce direction 16 B) into L2; an
        // experiment that lost (see PLT_UNWIND_PREFETCH above): the queue entries of the next item are still read one iteration early
        const uint32_t j = i + stride;
        uint32_t nid = 0, naux = 0;
        if (j < n) {
            nid = P.sorted[j]; naux = P.sortedAux[j];
#if PLT_UNWIND_PREFETCH
            const uint32_t npx = naux & 0x3ffffffu, nq = naux >> 26;
            const char* jr = reinterpret_cast<const char*>(P.job + size_t(nid) * kJobPlanes);
            prefetchL2(jr); prefetchL2(jr + 128);
            prefetchL2(P.vc + 4 * (size_t(nq) * P.pixelCount + npx));
            prefetchL2(P.bDir + size_t(nq) * P.pixelCount + npx);
#endif
        }
        uint32_t nextType = 0, nextAux = 0;
        if (unwindStep<kTex>(P, id, aux, evals, nextType, nextAux)) enqueueJob(P, nextType, id, nextAux);
        id = nid; aux = naux;
    }
    warpAdd(P.counters + 0, evals);
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(P.counters + 7, (unsigned long long)n);
}
#endif

// The BSDF query of pltpt_bsdf_eval_mueller evaluated with THIS translation unit's arithmetic (approximate division / square root,
// contracted FMAs, PLT_FAST_MATH) and the renderer's own single-column beam: what k_unwind_step computes for a replayed vertex.  Not a
// benchmark kernel (four evaluations per matrix, plain stores): it exists so that the BSDF tolerance tests can state how far the replay's
// arithmetic is from the oracle's (tests/test_gpu_bsdf.py::test_replay_arithmetic_*).
__global__ void __launch_bounds__(128) k_bsdf_eval_replay_math(DevScene sc, uint32_t materialID, uint32_t lobe, uint32_t n, uint32_t groups, const float* __restrict__ wi,
                                                               const float* __restrict__ wo, const float* __restrict__ hero, float* __restrict__ out, int hwss, float maxBeamOmega) {
    const uint64_t total = uint64_t(n) * groups;
    const DevMaterial m = loadMaterial(sc.materials, materialID);
    const BsdfCtx cx = {sc.profiles, hwss};
    for (uint64_t item = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; item < total; item += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t i = uint32_t(item / groups), g = uint32_t(item % groups);
        const float3 Wi = mk3(__ldg(wi + 3 * size_t(i)), __ldg(wi + 3 * size_t(i) + 1), __ldg(wi + 3 * size_t(i) + 2));
        const float3 Wo = mk3(__ldg(wo + 3 * size_t(i)), __ldg(wo + 3 * size_t(i) + 1), __ldg(wo + 3 * size_t(i) + 2));
        const float4 wl = hwssWavelengths(__ldg(hero + size_t(g) * n + i), hwss);
        SD sd;
        sd.posW = mk3(0.f); sd.V = Wi; sd.N = mk3(0.f, 0.f, 1.f); sd.T = mk3(1.f, 0.f, 0.f); sd.B = mk3(0.f, 1.f, 0.f);
        sd.faceN = sd.N; sd.uv = mk2(0.f, 0.f); sd.materialID = materialID; sd.frontFacing = true;
        float* dst = out + item * size_t(hwss) * 16;
        for (int c = 0; c < 4; ++c) {
            Beam beam = sourceBeamArea(Wi, -Wi, 1.f, 1.f, mk4(1.f), wl, maxBeamOmega, hwss);
            beam.L = mk4(c == 0 ? 1.f : 0.f); beam.L1 = mk4(c == 1 ? 1.f : 0.f); beam.L2 = mk4(c == 2 ? 1.f : 0.f); beam.L3 = mk4(c == 3 ? 1.f : 0.f);
            beam.propagate(sd.posW);
            bsdfEval(m, cx, sd, Wo, lobe, beam);
            for (int j = 0; j < hwss; ++j) {
                const float4 SP = beam.getSP(j);
                dst[j * 16 + 0 + c] = SP.x; dst[j * 16 + 4 + c] = SP.y; dst[j * 16 + 8 + c] = SP.z; dst[j * 16 + 12 + c] = SP.w;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_resolve(const __grid_constant__ FrameParams P) {
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel < P.pixelCount) resolvePixel(P, pixel);
}
}  // namespace

void launchSortFlip(const FrameParams& P, SortView v, bool resetTraceQueue, uint32_t grid, cudaStream_t s) {
    k_sort_queue<<<grid, kSortThreads, 0, s>>>(P, v);
    k_flip_queue<<<1, 32, 0, s>>>(P, v, resetTraceQueue);
}
void launchUnwindStep(const FrameParams& P, uint32_t grid, cudaStream_t s) {
    if (P.sc.tex.blobs) k_unwind_step<true><<<grid, kStageThreads, 0, s>>>(P);
    else k_unwind_step<false><<<grid, kStageThreads, 0, s>>>(P);
}
void launchBsdfEvalReplayMath(const DevScene& sc, uint32_t materialID, uint32_t lobe, uint32_t n, uint32_t groups, const float* wi, const float* wo, const float* hero, float* out, int hwss,
                              float maxBeamOmega, uint32_t grid, cudaStream_t s) {
    k_bsdf_eval_replay_math<<<grid, 128, 0, s>>>(sc, materialID, lobe, n, groups, wi, wo, hero, out, hwss, maxBeamOmega);
}
void launchResolve(const FrameParams& P, cudaStream_t s) { k_resolve<<<(P.pixelCount + 255) / 256, 256, 0, s>>>(P); }
}  // namespace plt
