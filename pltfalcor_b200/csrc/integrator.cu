// integrator.cu — the two stages of PLTPT::execute (reference Source/RenderPasses/PLTPT/PLTPT.cpp:487-495): the sample
// stage records a backward path per pixel, the solve stage sources generalized rays at the lights and replays them forward.
#include "plt_integrator.cuh"

namespace plt {
namespace {

constexpr int kStageThreads = 128;

// Pixels are assigned to threads in 8x4 tiles per warp so that a warp's primary rays (and its bounce-record lines) stay coherent.
PLT_D bool pixelOfThread(const FrameParams& P, uint32_t& px, uint32_t& py) {
    const uint32_t tilesX = (P.width + 7) / 8;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    px = (warp % tilesX) * 8 + (lane & 7);
    py = (warp / tilesX) * 4 + (lane >> 3);
    return px < P.width && py < P.height;
}

PLT_D void warpAdd(unsigned long long* dst, uint32_t v) {  // all 32 lanes must call
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, (unsigned long long)v);
}

__global__ void __launch_bounds__(kStageThreads) k_sample_stage(const __grid_constant__ FrameParams P) {
    uint32_t px, py;
    TraversalCounters tc = {0, 0};
    if (pixelOfThread(P, px, py)) tc = samplePixel(P, px, py);
    warpAdd(P.counters + 5, tc.nodes);
    warpAdd(P.counters + 6, tc.tris);
}

__global__ void __launch_bounds__(kStageThreads) k_solve_stage(const __grid_constant__ FrameParams P) {
    uint32_t px, py;
    uint32_t evals = 0, verts = 0;
    TraversalCounters tc = {0, 0};
    if (pixelOfThread(P, px, py)) verts = solvePixel(P, px, py, evals, tc);
    // warp-aggregated statistics (all lanes reach this point)
    warpAdd(P.counters + 0, evals);
    warpAdd(P.counters + 1, verts);
    warpAdd(P.counters + 4, verts * (verts - (verts ? 1u : 0u)) / 2u);
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
        out[size_t(pixel) * records + p] = b;
    }
}

}  // namespace

static int fillParams(PltptContext* c, FrameParams& P, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t h) {
    P.sc = c->scene;
    P.opt = c->opt;
    P.camPos = mk3(cam->posW[0], cam->posW[1], cam->posW[2]);
    P.camU = mk3(cam->cameraU[0], cam->cameraU[1], cam->cameraU[2]);
    P.camV = mk3(cam->cameraV[0], cam->cameraV[1], cam->cameraV[2]);
    P.camW = mk3(cam->cameraW[0], cam->cameraW[1], cam->cameraW[2]);
    P.jitterX = cam->jitterX; P.jitterY = cam->jitterY;
    P.polarizer = cam->polarizer; P.polarizerTheta = cam->polarizerTheta;
    for (int i = 0; i < 3; ++i) { P.viewRow0[i] = cam->viewMat[i]; P.viewRow1[i] = cam->viewMat[4 + i]; P.viewRow2[i] = cam->viewMat[8 + i]; }
    P.frameCount = frameCount; P.width = w; P.height = h; P.pixelCount = w * h;
    const size_t records = size_t(c->opt.maxBounces + 1) * P.pixelCount;
    PLT_CUDA(c->work[0].ensure(records * sizeof(uint4)));
    PLT_CUDA(c->work[1].ensure(records * sizeof(float4)));
    PLT_CUDA(c->work[2].ensure(records * sizeof(float2)));
    PLT_CUDA(c->work[3].ensure(16 * sizeof(unsigned long long)));
    P.bHit = c->work[0].as<uint4>();
    P.bDir = c->work[1].as<float4>();
    P.bPdf = c->work[2].as<float2>();
    P.counters = c->work[3].as<unsigned long long>();
    c->workW = w; c->workH = h; c->workBounces = c->opt.maxBounces + 1;
    P.useEmissive = c->opt.useEmissiveLights && c->emissiveTriangleCount > 0;
    P.useAnalytic = c->opt.useAnalyticLights && c->lightCount > 0;
    P.doNEE = c->opt.doNEE != 0;
    P.doMS = P.doNEE && c->opt.doMNEE;
    P.doMIS = P.doNEE && c->opt.doMIS;
    P.perTileSG = P.doNEE && c->opt.NEEUsePerTileSG;
    return PLTPT_OK;
}

int renderFrame(PltptContext* c, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t h, float* color, float* albedo, float* normal, double* accum,
                bool solve, cudaStream_t s) {
    FrameParams P{};
    const bool fresh = c->work[3].p == nullptr;
    if (int r = fillParams(c, P, frameCount, cam, w, h)) return r;
    P.color = reinterpret_cast<float4*>(color);
    P.albedo = reinterpret_cast<float4*>(albedo);
    P.normal = reinterpret_cast<float4*>(normal);
    P.accum = accum;
    if (fresh) PLT_CUDA(cudaMemsetAsync(P.counters, 0, 16 * sizeof(unsigned long long), s));
    // snapshot of the cumulative counters before this frame, so that "last frame" statistics can be formed at sync time
    PLT_CUDA(cudaMemcpyAsync(P.counters + 8, P.counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
    const uint32_t warps = ((w + 7) / 8) * ((h + 3) / 4);
    const uint32_t blocks = (warps * 32 + kStageThreads - 1) / kStageThreads;
    // device timers: one event triple per frame from a ring, read back in syncFrameStats()
    if (c->timerEvents.empty()) {
        c->timerEvents.resize(3 * PltptContext::kTimerRing);
        for (auto& e : c->timerEvents) PLT_CUDA(cudaEventCreate(&e));
    }
    if (c->timerCount == PltptContext::kTimerRing) { if (int r = syncFrameStats(c)) return r; }
    cudaEvent_t* ev = &c->timerEvents[3 * c->timerCount];
    cudaEventRecord(ev[0], s);
    k_sample_stage<<<blocks, kStageThreads, 0, s>>>(P);
    cudaEventRecord(ev[1], s);
    c->stats.kernelLaunches++;
    if (solve) {
        k_solve_stage<<<blocks, kStageThreads, 0, s>>>(P);
        c->stats.kernelLaunches++;
    }
    cudaEventRecord(ev[2], s);
    c->timerCount++;
    c->timerStream = s;
    PLT_CUDA(cudaGetLastError());
    c->statsDirty = true;
    return PLTPT_OK;
}

int syncFrameStats(PltptContext* c) {
    PLT_CUDA(cudaSetDevice(c->device));
    if (c->statsDirty && c->work[3].p) {
        PLT_CUDA(cudaStreamSynchronize(c->timerStream));
        unsigned long long host[16];
        PLT_CUDA(cudaMemcpy(host, c->work[3].p, sizeof(host), cudaMemcpyDeviceToHost));
        c->stats.bsdfEvals = host[0] - host[8];       // last frame
        c->stats.pathVertices = host[1] - host[9];
        c->stats.mneeIterations = host[3] - host[11];
        c->stats.traversalStackOverflows += host[2] & 0xffffffffull;
        c->stats.totalBsdfEvals += host[0];
        c->stats.totalPathVertices += host[1];
        c->stats.totalReplaySteps += host[4];
        c->stats.traversalNodesVisited += host[5];
        c->stats.traversalTrianglesTested += host[6];
        PLT_CUDA(cudaMemset(c->work[3].p, 0, 16 * sizeof(unsigned long long)));
        for (uint32_t i = 0; i < c->timerCount; ++i) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, c->timerEvents[3 * i], c->timerEvents[3 * i + 1]);
            cudaEventElapsedTime(&b, c->timerEvents[3 * i + 1], c->timerEvents[3 * i + 2]);
            c->stats.sampleStageMs += a;
            c->stats.solveStageMs += b;
        }
        c->stats.framesRendered += c->timerCount;
        c->timerCount = 0;
        c->statsDirty = false;
    }
    if (c->bvhScratch[7].p) {  // overflow counter of the standalone ray-query kernels (bvh.cu)
        uint32_t* ov = c->bvhScratch[7].as<uint32_t>() + 9;
        uint32_t v = 0;
        PLT_CUDA(cudaMemcpy(&v, ov, 4, cudaMemcpyDeviceToHost));
        if (v) { c->stats.traversalStackOverflows += v; PLT_CUDA(cudaMemset(ov, 0, 4)); }
    }
    return PLTPT_OK;
}

int copyBounces(PltptContext* c, PltptBounce* host, uint32_t w, uint32_t h) {
    FrameParams P{};
    P.pixelCount = w * h;
    P.bHit = c->work[0].as<uint4>(); P.bDir = c->work[1].as<float4>(); P.bPdf = c->work[2].as<float2>();
    const uint32_t records = c->opt.maxBounces + 1;
    const size_t bytes = size_t(P.pixelCount) * records * sizeof(PltptBounce);
    PLT_CUDA(c->work[4].ensure(bytes));
    k_gather_bounces<<<(P.pixelCount + 127) / 128, 128>>>(P, c->work[4].as<PltptBounce>(), records);
    c->stats.kernelLaunches++;
    PLT_CUDA(cudaGetLastError());
    PLT_CUDA(cudaMemcpy(host, c->work[4].p, bytes, cudaMemcpyDeviceToHost));
    c->work[4].release();
    return PLTPT_OK;
}

}  // namespace plt
