// context.h — internal host-side state behind a PltptHandle and the device-side scene view.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "plt_bsdf.cuh"
#include "plt_material.cuh"
#include "plt_envmap.cuh"

namespace plt {

// ---- compressed 8-wide BVH node (80 B), after Ylitie/Karras/Laine 2017 "Efficient Incoherent Ray Traversal on GPUs
// Through Compressed Wide BVHs": child boxes quantised to 8 bits on a per-node power-of-two grid.
struct alignas(16) BvhNode {
    float px, py, pz;          // node AABB origin
    uint8_t ex, ey, ez, imask; // per-axis grid exponent (biased by 127), bit i of imask = child i is an inner node
    uint32_t childBase;        // index of the first inner child node
    uint32_t triBase;          // index of the first leaf triangle
    uint8_t meta[8];           // inner: 0b001xxxxx | (24 + slot); leaf: (#tris in unary, 3 bits) << 5 | offset from triBase; 0 = empty
    uint8_t qlox[8], qloy[8], qloz[8], qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(BvhNode) == 80, "CWBVH node must be 80 B");

// leaf triangle record: world-space vertices (exactly the values the intersection contract prescribes) + ids, 48 B
struct alignas(16) BvhTri {
    float v0x, v0y, v0z; uint32_t instanceID;
    float v1x, v1y, v1z; uint32_t primitiveIndex;
    float v2x, v2y, v2z; uint32_t globalIndex;
};
static_assert(sizeof(BvhTri) == 48, "BvhTri must be 48 B");

// device view of the scene: Falcor buffers (as uploaded) + our derived tables
struct DevScene {
    const PltptPackedVertex* vertices;
    const uint8_t* indexData;
    const PltptGeometryInstance* instances;
    const PltptMeshDesc* meshes;
    const float* world;
    const float* worldIT;
    const DevMaterial* materials;
    const DevProfile* profiles;
    const PltptLightData* lights;
    const PltptEmissiveTriangle* etris;
    const PltptEmissiveFlux* eflux;
    const uint32_t* activeTris;
    const uint32_t* perMeshInstanceOffset;
    const uint2* alias;
    const uint32_t* instTriOffset;  // first global triangle of each instance (instanceCount + 1 entries)
    const BvhNode* bvhNodes;
    const BvhTri* bvhTris;
    uint32_t instanceCount, lightCount, etriCount, activeTriCount, triCount;
    float invWeightsSum;
    uint32_t hitTypeBits, hitInstBits, hitPrimBits;
    DevEnvMap env;  // env.texels == nullptr: the scene has no environment map
    DevTextures tex; // material textures, samplers and the raw material blobs (read only by materials flagged HF_Textured / HF_NormalMap)
};

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    cudaError_t ensure(size_t n) {
        if (n <= bytes) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace plt

#ifndef PLT_LANES
#define PLT_LANES 4   // frame lanes built in (streams + private work buffers, allocated when first used); pltpt_set_frame_overlap picks how many run
#endif
#ifndef PLT_DEFAULT_OVERLAP
#define PLT_DEFAULT_OVERLAP 4   // Cornell 1080p: 389 / 480 / 501 / 505 M paths/s with 1 / 2 / 3 / 4 frames in flight (profiles/r2w_overlap.jsonl)
#endif
struct PltptContext {
    int device = 0;
    PltptOptions opt{};
    bool hasScene = false;
    plt::DevScene scene{};
    // owned device copies
    plt::DevBuf vertices, indexData, instances, meshes, world, worldIT, materials, profiles, lights, etris, eflux, activeTris, perMeshOffset, alias, instTriOffset;
    plt::DevBuf bvhNodes, bvhTris, bvhParent, bvhScratch[8];  // bvhParent: parent wide node of every wide node (bottom-up refit)
    int bvhUpdateMode = 0;                    // PLTPT_BVH_UPDATE_REBUILD / _REFIT (pltpt_set_bvh_update_mode)
    plt::DevBuf envTexels, envImportance;
    plt::DevBuf texBlob, texDescs, texSamplers, materialBlobs;
    // host mirrors needed by the host logic
    std::vector<PltptGeometryInstance> hInstances;
    std::vector<PltptMeshDesc> hMeshes;
    std::vector<uint32_t> hInstTriOffset;
    std::vector<uint32_t> hMaterialEmission;  // emissionSpectralId per material (pltpt_build_light_collection)
    std::vector<float> hMaterialRgb;           // rgb of that emission profile, 3 per material
    bool ownLightCollection = false;           // the light collection was built by pltpt_build_light_collection
    uint32_t emissiveTriangleCount = 0, lightCount = 0, matrixCount = 0;
    int traceMode = 0;                         // pltpt_set_trace_mode: which traversal driver the ray-query entry points use
    uint32_t mneeGrid = 0;                     // cooperative grid of k_mnee_walk on this handle's device (resident CTAs), 0 = not queried yet
    // Frame lanes: frames are independent (every random number derives from (pixel, frameCount), pltpt_sample.rt.slang:154), so
    // pltpt_render_accumulate keeps several frames in flight (frameOverlap, default four), each on its own stream with its own work
    // buffers, counters and timers: the under-filled late bounces of one frame overlap the wide early bounces of the others.  Lane 0
    // runs on the caller's stream.  A frame rendered on its own (pltpt_render_frame) is split over lanes 0 and 1 (renderFrameSplit).
    static constexpr uint32_t kTimerRing = 256;
    static constexpr int kLanes = PLT_LANES;
    struct Lane {
        plt::DevBuf work[40];                  // per-frame work buffers (integrator.cu)
        plt::DevBuf accum;                     // lane-private RGBA64F sums (lanes >= 1), merged into the caller's buffer at the end
        cudaStream_t stream = nullptr;         // own non-blocking stream (lanes >= 1)
        cudaEvent_t done = nullptr, guidesReady = nullptr;
        cudaEvent_t lastUse = nullptr;         // recorded after every frame on the stream it ran on (orders the host-frame queue behind it)
        uint32_t partialFrames = 0;            // lower halves of split frames among timerCount
        bool statsDirty = false;
        std::vector<cudaEvent_t> timerEvents;  // kTimerRing x {start, after sample stage, after solve stage}
        uint32_t timerCount = 0;
        cudaStream_t timerStream = nullptr;
    } lanes[kLanes];
    cudaEvent_t evFork = nullptr;
    cudaEvent_t guidesReady = nullptr, copyDone = nullptr;  // pltpt_render_frame_host: guide-buffer copies run on copyStream under the solve stage
    cudaStream_t copyStream = nullptr;
    bool armGuides = false, splitGuides = false;  // splitGuides: the last frame was split over both lanes (both guidesReady events count)
    static constexpr int kDefaultOverlap = PLT_DEFAULT_OVERLAP < PLT_LANES ? PLT_DEFAULT_OVERLAP : PLT_LANES;
    int frameOverlap = kDefaultOverlap;  // frames in flight (pltpt_set_frame_overlap)
    size_t workBudgetBytes = 0; // per-lane work-buffer budget that sizes the row tiles (0 = default 24 GB; pltpt_set_work_budget)
    plt::DevBuf outColor, outAlbedo, outNormal;
    plt::DevBuf hostAccum;                    // RGBA64F sums of pltpt_render_accumulate_host / _multi
    // pltpt_render_frame_host_async: frame i runs whole on lane i % frameOverlap, on that lane's own stream, into one of that lane's
    // two output image sets; its device-to-host copies run on hostCopyStream behind an event, so the lane starts its next frame
    // (other image set) at once and only waits for a set's previous copy before overwriting that set
    cudaStream_t hostStream[kLanes] = {};
    cudaStream_t hostCopyStream = nullptr;
    plt::DevBuf hostOut[kLanes][2][3];
    cudaEvent_t hostFrameDone[kLanes][2] = {}, hostCopyDone[kLanes][2] = {};
    uint32_t hostSetNext[kLanes] = {};
    uint32_t hostAsyncNext = 0;
    uint64_t hostSeq = 0;                  // frames queued so far; the copy stream finishes them in this order
    cudaEvent_t hostSeqEvent[8] = {};      // copy-done event of frame (seq % 8), borrowed from hostCopyDone
    bool hostAsyncPending = false;
    // waits until at most `keep` queued frames are still on their way to the host
    int waitHostAsyncKeep(uint32_t keep) {
        if (keep == 0 || !hostAsyncPending) return waitHostAsync();
        if (hostSeq <= keep) return 0;
        cudaEvent_t e = hostSeqEvent[(hostSeq - keep - 1) % 8];  // a re-recorded event only makes the wait longer
        return e && cudaEventSynchronize(e) != cudaSuccess ? 1 : 0;
    }
    int waitHostAsync() {
        if (!hostAsyncPending) return 0;
        hostAsyncPending = false;
        int bad = 0;
        for (cudaStream_t s : hostStream) if (s && cudaStreamSynchronize(s) != cudaSuccess) bad = 1;
        if (hostCopyStream && cudaStreamSynchronize(hostCopyStream) != cudaSuccess) bad = 1;
        return bad;
    }
    PltptStats stats{};
    cudaEvent_t evA = nullptr, evB = nullptr;
    // per-kernel timers (pltpt_set_profiling): event after every launch, tagged with the kernel class that precedes it (0xff = frame start)
    bool profiling = false;
    std::vector<cudaEvent_t> kernelEvents;
    std::vector<uint8_t> kernelEventClass;
    uint32_t kernelEventCount = 0;
    void markKernel(uint8_t cls, cudaStream_t s) {
        if (!profiling) return;
        if (kernelEventCount == kernelEvents.size()) { cudaEvent_t e; cudaEventCreate(&e); kernelEvents.push_back(e); kernelEventClass.push_back(0); }
        kernelEventClass[kernelEventCount] = cls;
        cudaEventRecord(kernelEvents[kernelEventCount++], s);
    }
};

namespace plt {
void setError(const std::string& s);
int cudaFail(cudaError_t e, const char* what);
#define PLT_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return plt::cudaFail(_e, #call); } while (0)

// bvh.cu
int buildBvh(PltptContext* c, cudaStream_t s);
int refitBvh(PltptContext* c, cudaStream_t s);
int traceClosest(PltptContext* c, uint32_t n, const float* rays, uint32_t* hits, float* hitT, cudaStream_t s);
int traceAny(PltptContext* c, uint32_t n, const float* rays, uint32_t* occluded, cudaStream_t s);
// integrator.cu
int renderFrame(PltptContext* c, int lane, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t h, float* color, float* albedo, float* normal,
                double* accum, bool solve, cudaStream_t s, uint32_t rowBegin = 0, uint32_t rowEnd = 0);
int renderFrameSplit(PltptContext* c, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t h, float* color, float* albedo, float* normal, cudaStream_t s);
int sampleStageToHost(PltptContext* c, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t h, float* albedo, float* normal, PltptBounce* host);
int syncFrameStats(PltptContext* c);
// stage_unwind.cu: the BSDF query of pltpt_bsdf_eval_mueller evaluated with the replay step's arithmetic (test hook)
void launchBsdfEvalReplayMath(const DevScene& sc, uint32_t materialID, uint32_t lobe, uint32_t n, uint32_t groups, const float* wi, const float* wo, const float* hero, float* out, int hwss,
                              float maxBeamOmega, uint32_t grid, cudaStream_t s);
}  // namespace plt
