// integrator.cu — host side of PLTPT::execute's tile loop (reference Source/RenderPasses/PLTPT/PLTPT.cpp:487-495):
// per row tile, the sample stage records a backward path per pixel, then the solve stage runs as a wavefront
// (connect -> [sort, step] x (maxBounces+1) -> resolve).  Buffer sizing, statistics and the parity dump live here.
#include "plt_integrator.cuh"
#include <algorithm>

namespace plt {

namespace {
// worst-case work-buffer bytes per pixel: bounce planes + vertex cache + jobs (2 per recorded vertex) + queues + contributions
size_t workBytesPerPixel(uint32_t records) {
    // (parked manifold walks: worst case one per recorded vertex -- a connection parks at most once per (pixel, vertex index))
    return size_t(records) * (16 + 16 + 8 + 64 + 1 + 4) + size_t(2 * records) * (kJobPlanes * 16 + 4 + 1 + 4 + 8) + size_t(2 * records) * 16 + 6 * 16 + 4 + size_t(records) * (sizeof(MneeItem) + sizeof(MneeWalk) + 12) + sizeof(ConnItem) + 32 + 16 + 4;
}
constexpr size_t kWorkBudgetBytes = size_t(24) << 30;  // tiles are sized so that the work buffers stay below this
}  // namespace

static int fillParams(PltptContext* c, PltptContext::Lane& ln, FrameParams& P, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t h, uint32_t& tileRows) {
    P.sc = c->scene;
    P.opt = c->opt;
    P.camPos = mk3(cam->posW[0], cam->posW[1], cam->posW[2]);
    P.camU = mk3(cam->cameraU[0], cam->cameraU[1], cam->cameraU[2]);
    P.camV = mk3(cam->cameraV[0], cam->cameraV[1], cam->cameraV[2]);
    P.camW = mk3(cam->cameraW[0], cam->cameraW[1], cam->cameraW[2]);
    P.jitterX = cam->jitterX; P.jitterY = cam->jitterY;
    P.polarizer = cam->polarizer; P.polarizerTheta = cam->polarizerTheta;
    for (int i = 0; i < 3; ++i) { P.viewRow0[i] = cam->viewMat[i]; P.viewRow1[i] = cam->viewMat[4 + i]; P.viewRow2[i] = cam->viewMat[8 + i]; }
    P.frameCount = frameCount; P.width = w; P.height = h;
    const uint32_t recordsPerPixel = c->opt.maxBounces + 1;
    const size_t perRow = workBytesPerPixel(recordsPerPixel) * w;
    const size_t budget = c->workBudgetBytes ? c->workBudgetBytes : kWorkBudgetBytes;
    tileRows = uint32_t(std::min<size_t>(h, std::max<size_t>(4, (budget / perRow) & ~size_t(3))));
    const uint32_t tilePixels = tileRows * w;
    P.pixelCount = tilePixels;
    const size_t records = size_t(recordsPerPixel) * tilePixels;
    const size_t jobCapacity = 2 * records;
    if (jobCapacity >= (size_t(1) << 32)) { setError("tile too large for 32-bit job ids"); return PLTPT_ERR_INVALID_ARGUMENT; }
    PLT_CUDA(ln.work[0].ensure(records * sizeof(uint4)));
    PLT_CUDA(ln.work[1].ensure(records * sizeof(float4)));
    PLT_CUDA(ln.work[2].ensure(records * sizeof(float2)));
    PLT_CUDA(ln.work[3].ensure(16 * sizeof(unsigned long long)));
    PLT_CUDA(ln.work[5].ensure(records * 4 * sizeof(float4)));
    PLT_CUDA(ln.work[9].ensure(records));
    PLT_CUDA(ln.work[10].ensure(jobCapacity * kJobPlanes * sizeof(float4)));
    PLT_CUDA(ln.work[11].ensure(jobCapacity * 4));
    PLT_CUDA(ln.work[12].ensure(jobCapacity));
    PLT_CUDA(ln.work[13].ensure(jobCapacity * 4));
    PLT_CUDA(ln.work[6].ensure(jobCapacity * 4));
    PLT_CUDA(ln.work[7].ensure(jobCapacity * 4));
    PLT_CUDA(ln.work[14].ensure(256 * 4));
    PLT_CUDA(ln.work[15].ensure(jobCapacity * sizeof(float4)));  // contrib: 2 * records float4
    for (int k = 0; k < 7; ++k) PLT_CUDA(ln.work[16 + k].ensure(size_t(tilePixels) * 16));  // path state planes, hit, solve rng x2
    PLT_CUDA(ln.work[4].ensure(size_t(tilePixels) * 4));                                    // trace queue
    PLT_CUDA(ln.work[23].ensure(records * 4));                                               // per-bounce type-sorted hit lists
    PLT_CUDA(ln.work[26].ensure(records * sizeof(MneeItem)));                                // parked manifold walks (worst case: every recorded vertex parks one)
    PLT_CUDA(ln.work[27].ensure(size_t(tilePixels) * sizeof(ConnItem)));
    PLT_CUDA(ln.work[28].ensure(size_t(tilePixels) * 32));
    PLT_CUDA(ln.work[29].ensure(size_t(tilePixels) * 16));
    PLT_CUDA(ln.work[30].ensure(size_t(tilePixels) * 4));
    P.bHit = ln.work[0].as<uint4>();
    P.bDir = ln.work[1].as<float4>();
    P.bPdf = ln.work[2].as<float2>();
    P.counters = ln.work[3].as<unsigned long long>();
    P.vc = ln.work[5].as<float4>();
    P.vMatType = ln.work[9].as<uint8_t>();
    P.job = ln.work[10].as<float4>();
    P.jobCapacity = uint32_t(jobCapacity);
    P.pending = ln.work[11].as<uint32_t>();
    P.pendingType = ln.work[12].as<uint8_t>();
    P.sorted = ln.work[13].as<uint32_t>();
    P.pendingAux = ln.work[6].as<uint32_t>();
    P.sortedAux = ln.work[7].as<uint32_t>();
    P.queueCount = ln.work[14].as<uint32_t>();
    P.jobCount = P.queueCount + QC_JobCount;
    P.contrib = ln.work[15].as<float4>();
    P.debugView = c->opt.debugView;
    P.debugIntensity = c->opt.debugViewIntensity;
    P.debug = nullptr;
    if (P.debugView) {  // debug views only: one event slot per connection slot
        PLT_CUDA(ln.work[32].ensure(jobCapacity * sizeof(float4)));
        P.debug = ln.work[32].as<float4>();
    }
    P.psRng = ln.work[16].as<uint4>(); P.psOrigin = ln.work[17].as<float4>(); P.psDir = ln.work[18].as<float4>(); P.psPdf = ln.work[19].as<float4>();
    P.psHit = ln.work[20].as<uint4>(); P.ssRng = ln.work[21].as<uint4>(); P.ssTileRng = ln.work[22].as<uint4>();
    P.traceQueue = ln.work[4].as<uint32_t>();
    P.vertexList = ln.work[23].as<uint32_t>();
    P.levelCount = P.queueCount + QC_LevelCount;
    P.mneeItems = ln.work[26].as<MneeItem>();
    P.mneeCount = P.queueCount + QC_MneeCount;
    P.mneeCapacity = uint32_t(records);
    PLT_CUDA(ln.work[8].ensure(records * sizeof(MneeWalk)));
    PLT_CUDA(ln.work[31].ensure(size_t(3) * records * 4));
    if (!c->mneeGrid) c->mneeGrid = mneeCooperativeGrid(c->device);
    P.mneeGrid = c->mneeGrid;
    P.mneeWalks = ln.work[8].as<MneeWalk>();
    P.mneeLists = ln.work[31].as<uint32_t>();
    PLT_CUDA(ln.work[24].ensure(size_t(tilePixels) * 4));
    P.missList = ln.work[24].as<uint32_t>();
    P.missCount = P.queueCount + QC_MissCount;
    if (tilePixels >= (1u << 26)) { setError("tile too large for the packed miss list"); return PLTPT_ERR_INVALID_ARGUMENT; }
    P.connItems = ln.work[27].as<ConnItem>();
    P.shadowRays = ln.work[28].as<float4>();
    P.shadowHits = ln.work[29].as<uint4>();
    P.shadowT = ln.work[30].as<float>();
    P.useEmissive = c->opt.useEmissiveLights && c->emissiveTriangleCount > 0;
    P.useAnalytic = c->opt.useAnalyticLights && c->lightCount > 0;
    P.useEnv = c->opt.useEnvLights && c->scene.env.texels != nullptr;  // PLTPT.cpp:235
    P.tinyScene = c->stats.bvhNodes < 256;
    P.doNEE = c->opt.doNEE != 0;
    P.doMS = P.doNEE && c->opt.doMNEE;
    P.doMIS = P.doNEE && c->opt.doMIS;
    P.perTileSG = P.doNEE && c->opt.NEEUsePerTileSG;
    return PLTPT_OK;
}

// One frame = one sample per pixel.  `bouncesOut` (device, optional) receives the reference-layout bounce records.
static int renderTiles(PltptContext* c, PltptContext::Lane& ln, FrameParams& P, uint32_t tileRows, uint32_t rowBegin, uint32_t rowEnd, bool solve, PltptBounce* bouncesOut, cudaStream_t s) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const uint32_t persistentGrid = uint32_t(sms) * 8;
    for (uint32_t y0 = rowBegin; y0 < rowEnd; y0 += tileRows) {
        P.tileY0 = y0;
        P.tileRows = std::min(tileRows, rowEnd - y0);
        P.outOffset = y0 * P.width;
        P.pixelCount = P.tileRows * P.width;
        cudaEvent_t* ev = &ln.timerEvents[3 * ln.timerCount];
        const bool timed = y0 == rowBegin;  // stage timers cover the first tile of a frame (the only one unless the frame is tiled)
        const uint32_t maxB = c->opt.maxBounces;
        if (maxB > 63) { setError("maxBounces > 63 is not supported by the wavefront queues"); return PLTPT_ERR_INVALID_ARGUMENT; }
        if (timed) cudaEventRecord(ev[0], s);
        // ---- sample stage
        PLT_CUDA(cudaMemsetAsync(P.queueCount, 0, 256 * 4, s));
        if (P.debug) PLT_CUDA(cudaMemsetAsync(P.debug, 0, size_t(P.jobCapacity) * sizeof(float4), s));  // slots without an event must read as none
        c->markKernel(0xff, s);
        launchRaygen(P, s);
        c->markKernel(PLTPT_K_RAYGEN, s);
        c->stats.kernelLaunches++;
        for (uint32_t b = 0; b <= maxB; ++b) {
            P.level = b;
            launchTrace(P, persistentGrid, s);
            c->markKernel(PLTPT_K_TRACE, s);
            launchSortFlip(P, SortView{P.vertexList + size_t(b) * P.pixelCount, P.levelCount + b, nullptr}, true, persistentGrid, s);
            c->markKernel(PLTPT_K_SORT, s);
            launchShade(P, persistentGrid, s);
            c->markKernel(PLTPT_K_SHADE, s);
            c->stats.kernelLaunches += 4;
        }
        if (timed) cudaEventRecord(ev[1], s);
        // guide buffers (albedo, normal) are final once the last tile's sample stage has run: the host-buffer path starts their
        // device-to-host copies here, under the solve stage
        if (c->armGuides && y0 + tileRows >= rowEnd) cudaEventRecord(&ln == &c->lanes[0] ? c->guidesReady : ln.guidesReady, s);
        if (bouncesOut) { launchGatherBounces(P, bouncesOut, maxB + 1, s); c->stats.kernelLaunches++; c->markKernel(PLTPT_K_OTHER, s); }
        if (solve) {
            // ---- solve stage: connect per vertex index, then the replay wavefront, then the deterministic per-pixel sum
            for (uint32_t b = 0; b <= maxB; ++b) {
                P.level = b;
                launchConnectA(P, persistentGrid, s);
                c->markKernel(PLTPT_K_CONNECT_A, s);
                launchShadowTrace(P, persistentGrid, s);
                c->markKernel(PLTPT_K_SHADOW_TRACE, s);
                launchConnectB(P, persistentGrid, s);
                c->markKernel(PLTPT_K_CONNECT_B, s);
                c->stats.kernelLaunches += 3;
            }
            if (P.useEnv) { launchConnectMiss(P, persistentGrid, s); c->markKernel(PLTPT_K_CONNECT_A, s); c->stats.kernelLaunches++; }
            if (P.doMS) { launchMnee(P, persistentGrid, s); c->stats.kernelLaunches += 2; c->markKernel(PLTPT_K_MNEE, s); }
            for (uint32_t it = 0; it <= maxB; ++it) {
                launchSortFlip(P, SortView{P.sorted, P.queueCount + QC_SortedCount, P.sortedAux}, false, persistentGrid, s);
                c->markKernel(PLTPT_K_SORT, s);
                launchUnwindStep(P, persistentGrid, s);
                c->markKernel(PLTPT_K_UNWIND_STEP, s);
                c->stats.kernelLaunches += 3;
            }
            launchResolve(P, s);
            c->markKernel(PLTPT_K_RESOLVE, s);
            c->stats.kernelLaunches++;
        }
        if (timed) cudaEventRecord(ev[2], s);
    }
    PLT_CUDA(cudaGetLastError());
    return PLTPT_OK;
}

static int ensureTimers(PltptContext::Lane& ln) {
    if (ln.timerEvents.empty()) {
        ln.timerEvents.resize(3 * PltptContext::kTimerRing);
        for (auto& e : ln.timerEvents) PLT_CUDA(cudaEventCreate(&e));
    }
    return PLTPT_OK;
}

// Rows [rowBegin, rowEnd) of one frame on one lane (rowEnd = 0: the whole frame).
int renderFrame(PltptContext* c, int lane, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t h, float* color, float* albedo, float* normal, double* accum,
                bool solve, cudaStream_t s, uint32_t rowBegin, uint32_t rowEnd) {
    if (rowEnd == 0) rowEnd = h;
    PltptContext::Lane& ln = c->lanes[lane];
    FrameParams P{};
    const bool fresh = ln.work[3].p == nullptr;
    uint32_t tileRows = 0;
    if (int r = fillParams(c, ln, P, frameCount, cam, w, h, tileRows)) return r;
    P.color = reinterpret_cast<float4*>(color);
    P.albedo = reinterpret_cast<float4*>(albedo);
    P.normal = reinterpret_cast<float4*>(normal);
    P.accum = accum;
    if (fresh) PLT_CUDA(cudaMemsetAsync(P.counters, 0, 16 * sizeof(unsigned long long), s));
    // snapshot of the cumulative counters before this frame, so that "last frame" statistics can be formed at sync time
    PLT_CUDA(cudaMemcpyAsync(P.counters + 8, P.counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
    // device timers: one event triple per frame from a ring, read back in syncFrameStats()
    if (int r = ensureTimers(ln)) return r;
    if (ln.timerCount == PltptContext::kTimerRing || c->kernelEventCount > 60000) { if (int r = syncFrameStats(c)) return r; }
    if (int r = renderTiles(c, ln, P, tileRows, rowBegin, rowEnd, solve, nullptr, s)) return r;
    if (!ln.lastUse) PLT_CUDA(cudaEventCreateWithFlags(&ln.lastUse, cudaEventDisableTiming));
    PLT_CUDA(cudaEventRecord(ln.lastUse, s));
    ln.timerCount++;
    ln.timerStream = s;
    ln.statsDirty = true;
    if (rowBegin != 0) ln.partialFrames++;  // the lower part of a split frame: not a frame of its own in the statistics
    return PLTPT_OK;
}

// One frame for a caller that renders frame by frame (pltpt_render_frame, pltpt_render_frame_host): the two lanes take the upper
// and the lower half of the image concurrently (rows are independent: seeds derive from absolute pixel coordinates), so that the
// narrow late-bounce launches of one half share the GPU with the other half's.  `s` is the caller's stream; lane 1 forks from it
// and joins it again.
int renderFrameSplit(PltptContext* c, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t h, float* color, float* albedo, float* normal, cudaStream_t s) {
    const bool split = !c->profiling && c->frameOverlap >= 2 && PltptContext::kLanes >= 2 && h >= 128;
    if (!split) return renderFrame(c, 0, frameCount, cam, w, h, color, albedo, normal, nullptr, true, s, 0, h);
    PltptContext::Lane& l1 = c->lanes[1];
    if (!c->evFork) PLT_CUDA(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
    if (!l1.stream) PLT_CUDA(cudaStreamCreateWithFlags(&l1.stream, cudaStreamNonBlocking));
    if (!l1.done) PLT_CUDA(cudaEventCreateWithFlags(&l1.done, cudaEventDisableTiming));
    if (!l1.guidesReady) PLT_CUDA(cudaEventCreateWithFlags(&l1.guidesReady, cudaEventDisableTiming));
    const uint32_t mid = ((h / 2) + 3u) & ~3u;  // primary rays are traced in 8x4 pixel tiles per warp
    PLT_CUDA(cudaEventRecord(c->evFork, s));
    PLT_CUDA(cudaStreamWaitEvent(l1.stream, c->evFork, 0));
    if (int r = renderFrame(c, 0, frameCount, cam, w, h, color, albedo, normal, nullptr, true, s, 0, mid)) return r;
    if (int r = renderFrame(c, 1, frameCount, cam, w, h, color, albedo, normal, nullptr, true, l1.stream, mid, h)) return r;
    PLT_CUDA(cudaEventRecord(l1.done, l1.stream));
    PLT_CUDA(cudaStreamWaitEvent(s, l1.done, 0));
    c->splitGuides = true;
    return PLTPT_OK;
}

int syncFrameStats(PltptContext* c) {
    PLT_CUDA(cudaSetDevice(c->device));
    bool any = false;
    for (auto& ln : c->lanes) {
        if (!(ln.statsDirty && ln.work[3].p)) continue;
        any = true;
        PLT_CUDA(cudaStreamSynchronize(ln.timerStream));
        unsigned long long host[16];
        PLT_CUDA(cudaMemcpy(host, ln.work[3].p, sizeof(host), cudaMemcpyDeviceToHost));
        if (ln.partialFrames) {  // lower half of a split frame: completes the upper half's "last frame" figures
            c->stats.bsdfEvals += host[0] - host[8];
            c->stats.pathVertices += host[1] - host[9];
            c->stats.mneeIterations += host[3] - host[11];
        } else {
            c->stats.bsdfEvals = host[0] - host[8];       // last frame
            c->stats.pathVertices = host[1] - host[9];
            c->stats.mneeIterations = host[3] - host[11];
        }
        c->stats.traversalStackOverflows += host[2] & 0xffffffffull;
        c->stats.totalBsdfEvals += host[0];
        c->stats.totalPathVertices += host[1];
        c->stats.totalReplaySteps += host[4];
        c->stats.traversalNodesVisited += host[5];
        c->stats.traversalTrianglesTested += host[6];
        c->stats.unwindJobSteps += host[7];
        PLT_CUDA(cudaMemset(ln.work[3].p, 0, 16 * sizeof(unsigned long long)));
        uint32_t dropped = 0;
        PLT_CUDA(cudaMemcpy(&dropped, ln.work[14].as<uint32_t>() + QC_Dropped, 4, cudaMemcpyDeviceToHost));
        c->stats.droppedConnections += dropped;
        uint32_t walks = 0;
        PLT_CUDA(cudaMemcpy(&walks, ln.work[14].as<uint32_t>() + QC_MneeCount, 4, cudaMemcpyDeviceToHost));
        c->stats.mneeWalks = ln.partialFrames ? c->stats.mneeWalks + walks : walks;
        for (uint32_t i = 0; i < ln.timerCount; ++i) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, ln.timerEvents[3 * i], ln.timerEvents[3 * i + 1]);
            cudaEventElapsedTime(&b, ln.timerEvents[3 * i + 1], ln.timerEvents[3 * i + 2]);
            c->stats.sampleStageMs += a;
            c->stats.solveStageMs += b;
        }
        c->stats.framesRendered += ln.timerCount - std::min(ln.timerCount, ln.partialFrames);
        ln.partialFrames = 0;
        ln.timerCount = 0;
        ln.statsDirty = false;
    }
    if (any) {  // per-kernel timers are only recorded on lane 0 (profiling disables the frame overlap)
        for (uint32_t i = 1; i < c->kernelEventCount; ++i) {
            const uint8_t cls = c->kernelEventClass[i];
            if (cls >= PLTPT_KERNEL_CLASS_COUNT) continue;  // 0xff: first event of a tile, nothing precedes it
            float ms = 0.f;
            cudaEventElapsedTime(&ms, c->kernelEvents[i - 1], c->kernelEvents[i]);
            c->stats.kernelMs[cls] += ms;
            c->stats.kernelLaunchCount[cls]++;
        }
        c->kernelEventCount = 0;
    }
    if (c->bvhScratch[7].p) {  // overflow counter of the standalone ray-query kernels (bvh.cu)
        uint32_t* ov = c->bvhScratch[7].as<uint32_t>() + 9;
        uint32_t v = 0;
        PLT_CUDA(cudaMemcpy(&v, ov, 4, cudaMemcpyDeviceToHost));
        if (v) { c->stats.traversalStackOverflows += v; PLT_CUDA(cudaMemset(ov, 0, 4)); }
    }
    return PLTPT_OK;
}

// Sample stage only; writes the reference-layout bounce records of the whole frame to host memory (parity checks).
int sampleStageToHost(PltptContext* c, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t h, float* albedo, float* normal, PltptBounce* host) {
    PltptContext::Lane& ln = c->lanes[0];
    FrameParams P{};
    const bool fresh = ln.work[3].p == nullptr;
    uint32_t tileRows = 0;
    if (int r = fillParams(c, ln, P, frameCount, cam, w, h, tileRows)) return r;
    P.albedo = reinterpret_cast<float4*>(albedo);
    P.normal = reinterpret_cast<float4*>(normal);
    if (fresh) PLT_CUDA(cudaMemset(P.counters, 0, 16 * sizeof(unsigned long long)));
    if (int r = ensureTimers(ln)) return r;
    if (ln.timerCount == PltptContext::kTimerRing) { if (int r = syncFrameStats(c)) return r; }
    const uint32_t records = c->opt.maxBounces + 1;
    const size_t bytes = size_t(w) * h * records * sizeof(PltptBounce);
    PLT_CUDA(ln.work[25].ensure(bytes));
    if (int r = renderTiles(c, ln, P, tileRows, 0, h, false, ln.work[25].as<PltptBounce>(), 0)) return r;
    PLT_CUDA(cudaMemcpy(host, ln.work[25].p, bytes, cudaMemcpyDeviceToHost));
    ln.work[25].release();
    return PLTPT_OK;
}

}  // namespace plt
