// plt_integrator.cuh — device functions of the PLTPT sample and solve stages: light sampling, NEE / MNEE,
// forward replay of the generalized ray, measurement.  Semantics follow
//   Source/RenderPasses/PLTPT/{Path.slang, pltpt.slangh, pltpt_sample.rt.slang, pltpt_solve.rt.slang,
//   LightSampling.slangh, ShadowRays.slangh, Manifolds.slangh} and Source/Falcor/Rendering/Lights/*.slang (cited per function).
// Layout is ours: the bounce buffer is three SoA planes indexed [bounce][pixel] so that a warp of neighbouring
// pixels reads/writes full 128 B lines (the reference's 40 B AoS record is re-assembled only for parity dumps).
#pragma once
#include "bvh_traverse.cuh"
#include "plt_scene.cuh"
#include "plt_debug.cuh"

namespace plt {

// Path.slang:47-109
enum : uint32_t { RF_LastBounceWasDelta = 0x1, RF_HasDispersiveDeltaSegment = 0x2, RF_OnlyDelta = 0x4, RF_Miss = 0x10, RF_Terminated = 0x20, RF_MSMask = 0xffff0000u };

struct MneeItem;
struct FrameParams {
    DevScene sc;
    PltptOptions opt;
    float3 camPos, camU, camV, camW;
    float jitterX, jitterY, polarizer, polarizerTheta;
    float viewRow0[3], viewRow1[3], viewRow2[3];
    uint32_t frameCount, width, height, pixelCount;
    // bounce planes: [ (maxBounces+1) * pixelCount ]
    uint4* bHit;       // PackedHitInfo
    float4* bDir;      // encodedWi.xy, asfloat(lobe), asfloat(flags)
    float2* bPdf;      // pathContribution, lastNonDeltaPdf
    // per-vertex shading cache written by the sample stage so that the O(L^2) forward replay never re-fetches geometry: one
    // 64 B record (two full 32 B sectors) per vertex at [(bounce * pixelCount + pixel) * 4]:
    // [0] = posW.xyz | materialID, [1] = normalW.xyz | uv.x, [2] = faceNormalW.xyz | uv.y, [3] = tangentW
    float4* vc;
    uint8_t* vMatType;  // MaterialType of every recorded vertex (queue binning key)
    // connection jobs (one per NEE / emissive light connection): one 160 B record (five full 32 B sectors) per job, [job][10]
    float4* job;
    uint32_t jobCapacity;
    uint32_t* jobCount;      // [0] allocated jobs, [1] dropped (capacity overflow; must stay 0)
    // replay queues: the step kernel appends surviving job ids (+ the MaterialType of their next vertex) to `pending`, a
    // counting-sort scatter then writes them to `sorted` grouped by type, which is what the next step kernel walks.
    uint32_t* pending;       // [jobCapacity] job ids in completion order
    uint32_t* pendingAux;    // [jobCapacity] pixel | next vertex index << 26 of the job: the step kernel can fetch the job record and
                             //               the vertex record at once instead of one after the other
    uint8_t* pendingType;    // [jobCapacity]
    uint32_t* sorted;        // [jobCapacity] job ids grouped by material type
    uint32_t* sortedAux;     // [jobCapacity]
    uint32_t* queueCount;    // [0..kQueueTypes) per-type counts of `pending`, [kQueueTypes] total, [kQueueTypes+1 ..] scatter cursors,
                             // [2*kQueueTypes+2] number of entries of `sorted`
    uint32_t tileY0, tileRows, outOffset;  // row tile of the frame processed by this launch sequence (work buffers are tile-local)
    // sample-stage wavefront: per-path state (SoA, one float4 plane each) and the closest hit of the current bounce
    uint4* psRng;            // xoshiro128** state of the path
    float4* psOrigin;        // origin.xyz | hero wavelength
    float4* psDir;           // dir.xyz | weight
    float4* psPdf;           // pdf, thp_modifier, lastNonDeltaPdf, asfloat(flags | length << 8)   (flags use bits 0-5 and 16-31)
    uint4* psHit;            // instanceID, primitiveIndex, asuint(u), asuint(v)
    uint32_t* traceQueue;    // paths whose next ray is to be traced (appended by the shade kernel)
    uint32_t* vertexList;    // [(maxBounces+1)][pixelCount] paths with a recorded hit at bounce b, grouped by MaterialType
    uint32_t* levelCount;    // [0 .. maxBounces] entries of vertexList per bounce; [64] trace-queue count
    // solve-stage per-path state
    uint4* ssRng;            // random-number state of the solve stage (continues from vertex to vertex)
    uint4* ssTileRng;        // per-tile generator (NEEUsePerTileSG)
    uint32_t level;          // bounce / vertex index the current launch works on
    MneeItem* mneeItems;     // connections parked for the manifold-walk kernel
    struct MneeWalk* mneeWalks;  // their walk state between rounds
    uint32_t* mneeLists;     // [3][mneeCapacity]: two ping-pong lists of live walks + the list of converged walks
    struct ConnItem* connItems;  // [pixelCount] hand-over between the two connect halves, indexed like the level's work list
    float4* shadowRays;      // [2 * pixelCount] origin|tmin, dir|tmax of the NEE shadow rays
    uint4* shadowHits;       // [pixelCount] instanceID (0xffffffff = nothing hit), primitiveIndex, asuint(u), asuint(v)
    float* shadowT;          // [pixelCount]
    uint32_t* mneeCount;
    uint32_t mneeCapacity;
    uint32_t mneeGrid;       // cooperative grid of k_mnee_walk on this device (every CTA resident)
    float4* contrib;         // [(2 * bounce + kind)][pixel] measured RGB of each connection (w = 1: measure() ran), summed in fixed order by the resolve kernel
    float4* debug;           // same shape: the debug-view event of that connection slot (value.rgb, 1); allocated and zeroed only when debugView != 0
    uint32_t debugView;      // DebugViewType (0 = off), pltpt_debug.slangh
    float debugIntensity;
    float4* color;     // RGBA32F (may be null when accum is used)
    float4* albedo;
    float4* normal;
    double* accum;     // RGBA64F running sums (may be null)
    unsigned long long* counters;  // [0] bsdf evals x wavelengths, [1] recorded vertices, [2] traversal stack overflows, [3] mnee iterations,
                                   // [4] sum L(L-1)/2, [5] BVH nodes visited, [6] triangles tested, [7] unwind job steps
    // derived switches: pltpt.slangh:100-122 + PLTPT::getDefines (PLTPT.cpp:216-256)
    bool useEmissive, useAnalytic, useEnv, doNEE, doMS, doMIS, perTileSG;
    bool tinyScene;          // BVH of a few nodes: the trace kernels skip the persistent-lane refill logic
    // environment-lit scenes: paths that left the scene, in the order the trace kernels found them (each pixel misses at most once)
    uint32_t* missList;      // [pixelCount] pixel | vertex index << 26
    uint32_t* missCount;
};

constexpr int kQueueTypes = 32;  // MaterialType values 0..11, | kFinishingBin for replay jobs whose next step is their last (measure)
constexpr uint32_t kFinishingBin = 16;
// Slots of the per-lane counter block `queueCount` (256 x uint32, zeroed at the start of every tile).
enum : uint32_t {
    QC_Hist = 0,                            // [kQueueTypes] per-type counts of `pending`
    QC_Total = kQueueTypes,                 // entries of `pending`
    QC_Cursor = kQueueTypes + 1,            // [kQueueTypes] scatter cursors of k_sort_queue
    QC_SortedCount = 2 * kQueueTypes + 2,   // entries of `sorted`
    QC_JobCount = 70,                       // allocated connection jobs
    QC_Dropped = 71,                        // connections dropped for lack of capacity (must stay 0)
    QC_MneeCount = 72,                      // parked manifold walks
    QC_MissCount = 73,                      // entries of missList
    QC_MneeRound = 74,                      // [3] rotating live-walk counts of k_mnee_walk
    QC_MneeDone = 77,                       // converged walks
    QC_TraceCursor = 78,                    // fetch cursor of k_trace (re-zeroed by k_flip_queue)
    QC_ShadowCursor = 79,                   // fetch cursor of k_shadow_trace (re-zeroed by k_connect_a)
    QC_MneeTraceCursor = 80,                // [3] rotating fetch cursors of the trial-ray trace in k_mnee_walk
    QC_LevelCount = 96,                     // [maxBounces + 1 <= 64] entries of vertexList per bounce
    QC_TraceCount = QC_LevelCount + 64,     // entries of traceQueue (sample stage) / deferred entries of half B (solve stage)
};
static_assert(QC_SortedCount < QC_JobCount && QC_MneeTraceCursor + 3 <= QC_LevelCount && QC_TraceCount < 256, "queueCount slots overlap");
constexpr uint32_t kTraceCountSlot = 64;  // levelCount[kTraceCountSlot] == queueCount[QC_TraceCount]
constexpr int kJobPlanes = 10;

struct Bounce {
    uint4 hit;
    float3 wi;
    uint32_t lobe, flags;
    float pathContribution, lastNonDeltaPdf;
};
PLT_D void writeBounce(const FrameParams& P, uint32_t pixel, int p, uint4 hit, float3 wi, uint32_t lobe, uint32_t flags, float contribution, float lastNonDeltaPdf) {
    const size_t o = size_t(p) * P.pixelCount + pixel;
    const float2 e = ndir_to_oct_equal_area_snorm(wi);
    P.bHit[o] = hit;
    P.bDir[o] = mk4(e.x, e.y, __uint_as_float(lobe), __uint_as_float(flags));
    P.bPdf[o] = mk2(contribution, lastNonDeltaPdf);
}
PLT_D Bounce readBounce(const FrameParams& P, uint32_t pixel, int p) {
    const size_t o = size_t(p) * P.pixelCount + pixel;
    Bounce b;
    b.hit = P.bHit[o];
    const float4 d = P.bDir[o];
    const float2 q = P.bPdf[o];
    b.wi = oct_to_ndir_equal_area_snorm(mk2(d.x, d.y));
    b.lobe = __float_as_uint(d.z); b.flags = __float_as_uint(d.w);
    b.pathContribution = q.x; b.lastNonDeltaPdf = q.y;
    return b;
}

PLT_D void writeVertexCache(const FrameParams& P, uint32_t pixel, int p, const VertexData& v, uint32_t materialID, uint32_t materialType) {
    const size_t o = size_t(p) * P.pixelCount + pixel;
    float4* r = P.vc + 4 * o;
    r[0] = mk4(v.posW.x, v.posW.y, v.posW.z, __uint_as_float(materialID));
    r[1] = mk4(v.normalW.x, v.normalW.y, v.normalW.z, v.texC.x);
    r[2] = mk4(v.faceNormalW.x, v.faceNormalW.y, v.faceNormalW.z, v.texC.y);
    r[3] = v.tangentW;
    P.vMatType[o] = uint8_t(materialType);
}
PLT_D VertexData readVertexCache(const float4* r, uint32_t& materialID);
PLT_D VertexData readVertexCache(const FrameParams& P, uint32_t pixel, int p, uint32_t& materialID) { return readVertexCache(P.vc + 4 * (size_t(p) * P.pixelCount + pixel), materialID); }
PLT_D VertexData readVertexCache(const float4* r, uint32_t& materialID) {
    const float4 a = r[0], b = r[1], c = r[2], d = r[3];
    VertexData v;
    v.posW = mk3(a.x, a.y, a.z); materialID = __float_as_uint(a.w);
    v.normalW = mk3(b.x, b.y, b.z); v.faceNormalW = mk3(c.x, c.y, c.z);
    v.texC = mk2(b.w, c.w);
    v.tangentW = d;
    return v;
}

// ---- rays ------------------------------------------------------------------------------------------------------
constexpr float kShadowRayTMin = 1e-4f;
PLT_D bool traceClosestHit(const FrameParams& P, TraversalCounters& tc, float3 o, float3 d, float tmin, float tmax, TriHit& th, float& t) {
    Hit h;
    const bool hit = traverseBvh<false>(P.sc.bvhNodes, P.sc.bvhTris, o, d, tmin, tmax, h, reinterpret_cast<uint32_t*>(P.counters + 2), &tc);
    th.instanceID = h.instanceID; th.primitiveIndex = h.primitiveIndex; th.u = h.u; th.v = h.v;
    t = h.t;
    return hit;
}
PLT_D bool occluded(const FrameParams& P, TraversalCounters& tc, float3 o, float3 d, float tmin, float tmax) {
    Hit h;
    return traverseBvh<true>(P.sc.bvhNodes, P.sc.bvhTris, o, d, tmin, tmax, h, reinterpret_cast<uint32_t*>(P.counters + 2), &tc);
}
// ShadowRays.slangh:50-52
PLT_D float3 shadowRayOrigin(const SD& sd, float3 rayDir) { return computeRayOrigin(sd.posW, dot(sd.faceN, rayDir) >= .0f ? sd.faceN : -sd.faceN); }

// ---- sample stage ----------------------------------------------------------------------------------------------
struct PathState {  // Path.slang:146-196
    Rng sg;
    float3 origin, dir;
    float wavelength, weight, pdf, thp;
    uint32_t flags;
    int length;
    float lastNonDeltaPdf;
};
// Camera.slang:46-78 (computeRayPinhole)
PLT_D void cameraRay(const FrameParams& P, uint32_t px, uint32_t py, float3& origin, float3& dir) {
    float2 p = mk2((float(px) + 0.5f) / float(P.width), (float(py) + 0.5f) / float(P.height));
    p = p + mk2(-P.jitterX, P.jitterY);
    const float2 ndc = mk2(2.f * p.x - 1.f, -2.f * p.y + 1.f);
    origin = P.camPos;
    dir = normalize(ndc.x * P.camU + ndc.y * P.camV + P.camW);
}
// pltpt.slangh:170-176
PLT_D float russianRoulette(float rrVal, float u) {
    const float p = rrVal * rrVal;  // pow(rrVal, 2)
    return u < p ? .0f : 1.f - p;
}
// pltpt_sample.rt.slang:63-141 (scatterTriangleMeshClosestHit)
// Returns true when the hit was alpha-masked: the origin moved behind the surface and the same ray has to be re-traced.
template <bool kTex = true>
PLT_D bool scatterClosestHit(const FrameParams& P, uint32_t pixel, PathState& path, const TriHit& th) {
    const uint4 hitInfo = packHit(P.sc, th);
    uint32_t materialID;
    const VertexData v = getVertexData(P.sc, th, materialID);
    const DevMaterial m = loadMaterialAt<kTex>(P.sc, materialID, v.texC);
    const float3 wi = -path.dir;
    const SD sd = prepareShadingData<kTex>(P.sc, v, materialID, m, wi);
    if (P.opt.alphaMasking && m.opacity == .0f) {
        path.origin = sd.computeNewRayOrigin(false);
        return true;
    }
    const float rcp_thp = P.opt.doRussianRoulette ? russianRoulette(1.f - path.weight, path.sg.next1D()) : 1.f;
    bool shouldEndPath = path.length == int(P.opt.maxBounces) || (rcp_thp == .0f && path.length > 0);
    Sample bs;
    bs.wo = mk3(0.f); bs.pdf = 0.f; bs.lobe = 0; bs.weight = 0.f;
    if (!shouldEndPath) {
        const bool valid = bsdfSample(m, sd, path.sg, bs, path.wavelength, L_All, mk3(0.f));
        shouldEndPath = !(valid && bs.pdf != .0f);
        shouldEndPath |= path.pdf * bs.pdf * path.thp * rcp_thp < 1e-6f;
    }
    const bool isDelta = (bs.lobe & L_Delta) != 0, isDispersive = (bs.lobe & L_DispersiveDelta) != 0, isReflection = (bs.lobe & L_Reflection) != 0;
    const bool isMSCompatible = (isReflection && (m.propFlags & P_CanMSReflective)) || (!isReflection && (m.propFlags & P_CanMSTransmissive));
    if (!isDelta && (path.flags & RF_OnlyDelta)) {
        if (P.albedo) P.albedo[P.outOffset + pixel] = mk4(m.albx, m.alby, m.albz, 1.f);
        if (P.normal)
            P.normal[P.outOffset + pixel] = mk4(P.viewRow0[0] * sd.N.x + P.viewRow0[1] * sd.N.y + P.viewRow0[2] * sd.N.z, P.viewRow1[0] * sd.N.x + P.viewRow1[1] * sd.N.y + P.viewRow1[2] * sd.N.z,
                                  P.viewRow2[0] * sd.N.x + P.viewRow2[1] * sd.N.y + P.viewRow2[2] * sd.N.z, 1.f);
    }
    writeVertexCache(P, pixel, path.length, v, materialID, m.type);
    if (shouldEndPath) {
        path.flags |= RF_Terminated;
        writeBounce(P, pixel, path.length, hitInfo, wi, 0u, path.flags, 1.f / (path.pdf * path.thp), path.lastNonDeltaPdf);
        return false;
    }
    path.weight *= bs.weight;
    path.thp *= rcp_thp;
    path.pdf *= bs.pdf;
    if (isDispersive) path.flags |= RF_HasDispersiveDeltaSegment;
    writeBounce(P, pixel, path.length, hitInfo, wi, bs.lobe, path.flags, 1.f / (path.pdf * path.thp), path.lastNonDeltaPdf);
    ++path.length;
    path.lastNonDeltaPdf = !isDelta ? bs.pdf : path.lastNonDeltaPdf;
    const uint32_t ms = isMSCompatible ? (path.flags >> 16) + 1u : 0u;
    path.flags = (path.flags & ~RF_MSMask) | (ms << 16);
    path.flags = isDelta ? (path.flags | RF_LastBounceWasDelta) : (path.flags & ~RF_LastBounceWasDelta);
    if (!isDelta) path.flags &= ~RF_OnlyDelta;
    path.origin = sd.computeNewRayOrigin(isReflection);
    path.dir = bs.wo;
    return false;
}
// ---- sample-stage wavefront ---------------------------------------------------------------------------------------------
// raygen -> [trace -> sort by hit material -> shade] per bounce.  A path lives in the ps* planes between kernels.
PLT_D void storePath(const FrameParams& P, uint32_t pixel, const PathState& path) {
    P.psRng[pixel] = make_uint4(path.sg.s0, path.sg.s1, path.sg.s2, path.sg.s3);
    P.psOrigin[pixel] = mk4(path.origin.x, path.origin.y, path.origin.z, path.wavelength);
    P.psDir[pixel] = mk4(path.dir.x, path.dir.y, path.dir.z, path.weight);
    P.psPdf[pixel] = mk4(path.pdf, path.thp, path.lastNonDeltaPdf, __uint_as_float(path.flags | (uint32_t(path.length) << 8)));
}
PLT_D void loadPath(const FrameParams& P, uint32_t pixel, PathState& path) {
    const uint4 r = P.psRng[pixel];
    const float4 o = P.psOrigin[pixel], d = P.psDir[pixel], q = P.psPdf[pixel];
    path.sg = Rng{r.x, r.y, r.z, r.w};
    path.origin = mk3(o.x, o.y, o.z); path.wavelength = o.w;
    path.dir = mk3(d.x, d.y, d.z); path.weight = d.w;
    path.pdf = q.x; path.thp = q.y; path.lastNonDeltaPdf = q.z;
    const uint32_t fl = __float_as_uint(q.w);
    path.flags = fl & ~0xff00u; path.length = int((fl >> 8) & 0xffu);
}
PLT_D void pixelCoords(const FrameParams& P, uint32_t pixel, uint32_t& px, uint32_t& py) { px = pixel % P.width; py = P.tileY0 + pixel / P.width; }
// i-th entry of the implicit bounce-0 trace queue: pixels in 8x4 tiles so that a warp's primary rays stay coherent
PLT_D bool primaryPixel(const FrameParams& P, uint32_t i, uint32_t& pixel) {
    const uint32_t tilesX = (P.width + 7) / 8;
    const uint32_t warp = i >> 5, lane = i & 31;
    const uint32_t lx = (warp % tilesX) * 8 + (lane & 7), ly = (warp / tilesX) * 4 + (lane >> 3);
    pixel = ly * P.width + lx;
    return lx < P.width && ly < P.tileRows;
}
// pltpt_sample.rt.slang:146-166 (raygen up to the first TraceRay)
PLT_D void raygenPixel(const FrameParams& P, uint32_t pixel) {
    uint32_t px, py;
    pixelCoords(P, pixel, px, py);
    PathState path;
    path.sg = Rng::seed(px, py, P.frameCount);
    path.wavelength = sampleHeroWavelength(path.sg.next1D());
    if (P.normal) P.normal[P.outOffset + pixel] = mk4(0.f);
    if (P.albedo) P.albedo[P.outOffset + pixel] = mk4(0.f);
    cameraRay(P, px, py, path.origin, path.dir);
    path.weight = path.pdf = path.thp = 1.f;
    path.flags = RF_OnlyDelta;
    path.lastNonDeltaPdf = 0.f;
    path.length = 0;
    storePath(P, pixel, path);
}
// Records a path that left the scene for the solve stage's onMiss (warp-aggregated append).
PLT_D void appendMiss(const FrameParams& P, uint32_t pixel, uint32_t vertexIndex) {
    const uint32_t active = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(active) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(P.missCount, uint32_t(__popc(active)));
    base = __shfl_sync(active, base, leader);
    P.missList[base + __popc(active & ((1u << lane) - 1u))] = pixel | (vertexIndex << 26);
}
// TraceRay + scatterMiss (pltpt_sample.rt.slang:55-61,174-183), after the traversal: records the closest hit of the path's ray, or
// ends the path with a miss record.  Returns the MaterialType of the hit (0 = miss, path ended).
PLT_D uint32_t tracePathCommit(const FrameParams& P, uint32_t pixel, bool hit, const Hit& h) {
    if (!hit) {
        const float4 d = P.psDir[pixel];
        const float4 q = P.psPdf[pixel];
        const uint32_t fl = __float_as_uint(q.w);
        const uint32_t flags = (fl & ~0xff00u) | RF_Miss;
        writeBounce(P, pixel, int((fl >> 8) & 0xffu), make_uint4(0, 0, 0, 0), mk3(-d.x, -d.y, -d.z), 0u, flags, 1.f / (q.x * q.y), q.z);
        if (P.useEnv) appendMiss(P, pixel, (fl >> 8) & 0xffu);
        return 0u;
    }
    P.psHit[pixel] = make_uint4(h.instanceID, h.primitiveIndex, __float_as_uint(h.u), __float_as_uint(h.v));
    const uint32_t materialID = __ldg(&P.sc.instances[h.instanceID].materialID);
    return __ldg(&P.sc.materials[materialID].type);
}
// scatterTriangleMeshClosestHit on the stored hit; returns true when the path continues with a new ray.
template <bool kTex = true>
PLT_D bool shadePath(const FrameParams& P, uint32_t pixel, TraversalCounters& tc) {
    PathState path;
    loadPath(P, pixel, path);
    const uint4 h = P.psHit[pixel];
    TriHit th;
    th.instanceID = h.x; th.primitiveIndex = h.y; th.u = __uint_as_float(h.z); th.v = __uint_as_float(h.w);
    // Alpha-masked hits do not count as a bounce (pltpt_sample.rt.slang:83-87): the ray is re-traced at once from behind the
    // surface.  (Rare: kept inline instead of a round trip through the queues; bounded like the oracle's pass-through guard.)
    for (int guard = 0; scatterClosestHit<kTex>(P, pixel, path, th); ++guard) {
        float t;
        if (guard >= 4096 || !traceClosestHit(P, tc, path.origin, path.dir, 0.f, kFltMax, th, t)) {
            path.flags |= RF_Miss;
            writeBounce(P, pixel, path.length, make_uint4(0, 0, 0, 0), -path.dir, 0u, path.flags, 1.f / (path.pdf * path.thp), path.lastNonDeltaPdf);
            if (P.useEnv) appendMiss(P, pixel, uint32_t(path.length));
            return false;
        }
    }
    if (path.flags & (RF_Miss | RF_Terminated)) return false;
    storePath(P, pixel, path);
    return true;
}

// ---- light sampling (LightSampling.slangh, Rendering/Lights/*.slang) ---------------------------------------------
enum { LT_EnvMap = 0, LT_Emissive = 1, LT_Analytic = 2 };
struct LightSample {
    float pdf;
    float3 origin;
    float distance;
    float3 dir;
    Beam beam;
    uint32_t lightType;
};
PLT_D bool anyPositive(float4 L, int hwss) { return L.x > 0.f || (hwss > 1 && L.y > 0.f) || (hwss > 2 && L.z > 0.f) || (hwss > 3 && L.w > 0.f); }
// LightSampling.slangh:40-76
PLT_D void lightTypeProbabilities(const FrameParams& P, float p[3]) {
    p[0] = P.useEnv ? 1.f : 0.f; p[1] = P.useEmissive ? 1.f : 0.f; p[2] = P.useAnalytic ? 1.f : 0.f;
    const float sum = p[0] + p[1] + p[2];
    if (sum == 0.f) return;
    const float invSum = 1.f / sum;
    p[0] *= invSum; p[1] *= invSum; p[2] *= invSum;
}
PLT_D bool selectLightType(const FrameParams& P, uint32_t& lightType, float& pdf, Rng& sg) {
    float p[3];
    lightTypeProbabilities(P, p);
    float u = sg.next1D();
    for (lightType = 0; lightType < 3; ++lightType) {
        if (u < p[lightType]) { pdf = p[lightType]; return true; }
        u -= p[lightType];
    }
    return false;
}
struct EmissiveTri { float3 p0, p1, p2, normal; float area; uint32_t materialID; uint32_t uv0, uv1, uv2; };  // uv: 2 x fp16 per vertex
// LightCollectionShared.slang:118-135
PLT_D EmissiveTri loadEmissiveTriangle(const DevScene& sc, uint32_t idx) {
    const float4* p = reinterpret_cast<const float4*>(sc.etris + idx);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    EmissiveTri t;
    t.p0 = mk3(a.x, a.y, a.z); t.p1 = mk3(b.x, b.y, b.z); t.p2 = mk3(c.x, c.y, c.z);
    t.uv0 = __float_as_uint(a.w); t.uv1 = __float_as_uint(b.w); t.uv2 = __float_as_uint(c.w);
    t.normal = decodeNormal2x16(__float_as_uint(d.x));
    t.area = d.y;  // stored as fp32 bits
    t.materialID = __float_as_uint(d.z);
    return t;
}
// EmissiveLightSamplerHelpers.slang:56-74: emission = profile spectrum x rgbToSpectrum(emissionRGB), emissionRGB = texEmissive at `uv`
// (uniform value 1)
template <bool kTex = true>
PLT_D void materialEvalEmissive(const FrameParams& P, const EmissiveTri& tri, float3 b, float wavelength, bool importanceSample, Rng& sg, float4& Le, float4& wl) {
    const uint32_t materialID = tri.materialID;
    const uint32_t id = __ldg(&P.sc.materials[materialID].emissionId);
    if (id != PLTPT_NO_SPECTRAL_PROFILE) {
        float3 emissionRGB = mk3(1.f);
        if (kTex && P.sc.tex.blobs && (__ldg(&P.sc.materials[materialID].hdrFlags) & HF_Textured)) {
            // EmissiveTriangle::getTexCoord (LightCollectionShared.slang:65-70)
            const float2 uv = mk2(h2f(tri.uv0 & 0xffffu), h2f(tri.uv0 >> 16)) * b.x + mk2(h2f(tri.uv1 & 0xffffu), h2f(tri.uv1 >> 16)) * b.y + mk2(h2f(tri.uv2 & 0xffffu), h2f(tri.uv2 >> 16)) * b.z;
            const DevMaterial m = loadMaterialAt<true>(P.sc, materialID, uv);
            emissionRGB = mk3(m.c1x, m.c1y, m.c1z);
        }
        Le = sampleSpectrum(P.sc.profiles, id, wavelength, sg, importanceSample, int(P.opt.HWSS), wl);
        Le *= rgbToSpectrum4(emissionRGB, wl);
    }
}
struct TriangleLightSample { float3 posW, normalW, dir; float distance; float4 Le, wavelengths; float pdf; };
// EmissiveLightSamplerHelpers.slang:86-128
template <bool kTex = true>
PLT_D bool sampleTriangle(const FrameParams& P, float3 posW, uint32_t triangleIndex, Rng& sg, float wavelength, bool importanceSample, TriangleLightSample& ls) {
    ls.Le = mk4(0.f); ls.wavelengths = mk4(0.f); ls.pdf = 0.f;
    const EmissiveTri tri = loadEmissiveTriangle(P.sc, triangleIndex);
    const float2 u = sg.next2D();
    const float3 b = sample_triangle(u);
    ls.posW = tri.p0 * b.x + tri.p1 * b.y + tri.p2 * b.z;
    const float3 toLight = ls.posW - posW;
    const float distSqr = fmaxf(kFltMin, dot(toLight, toLight));
    ls.distance = sqrtf(distSqr);
    ls.dir = toLight / ls.distance;
    ls.normalW = tri.normal;
    const float cosTheta = dot(ls.normalW, -ls.dir);
    if (cosTheta <= 0.f) return false;
    materialEvalEmissive<kTex>(P, tri, b, wavelength, importanceSample, sg, ls.Le, ls.wavelengths);
    ls.pdf = distSqr / fmaxf(kFltMin, cosTheta * tri.area);
    return true;
}
// EmissivePowerSampler.slang:64-90 / EmissiveUniformSampler.slang:56-72
template <bool kTex = true>
PLT_D bool emissiveSampleLight(const FrameParams& P, float3 posW, Rng& sg, float wavelength, bool importanceSample, TriangleLightSample& ls) {
    const DevScene& sc = P.sc;
    uint32_t triangleIndex;
    float selectionPdf;
    if (P.opt.emissiveSampler == PLTPT_EMISSIVE_SAMPLER_POWER) {
        if (sc.etriCount == 0) return false;
        const float uLight = sg.next1D();
        triangleIndex = min(uint32_t(uLight * sc.etriCount), sc.etriCount - 1);
        const uint2 packed = __ldg(sc.alias + triangleIndex);
        const float threshold = h2f(packed.x >> 16u);
        const uint32_t selectAbove = ((packed.x & 0xFFFFu) << 8u) | ((packed.y >> 24u) & 0xFFu);
        const uint32_t selectBelow = packed.y & 0xFFFFFFu;
        triangleIndex = (sg.next1D() >= threshold) ? selectAbove : selectBelow;
        selectionPdf = __ldg(&sc.eflux[triangleIndex].flux) * sc.invWeightsSum;
    } else {
        if (sc.activeTriCount == 0) return false;
        const float uLight = sg.next1D();
        const uint32_t idx = min(uint32_t(uLight * sc.activeTriCount), sc.activeTriCount - 1);
        triangleIndex = __ldg(sc.activeTris + idx);
        selectionPdf = 1.f / float(sc.activeTriCount);
    }
    if (!sampleTriangle<kTex>(P, posW, triangleIndex, sg, wavelength, importanceSample, ls)) return false;
    ls.pdf *= selectionPdf;
    return true;
}
// EmissivePowerSampler.slang:98-126, EmissiveLightSamplerHelpers.slang:135-151
PLT_D float emissiveEvalPdf(const FrameParams& P, float3 posW, uint32_t triangleIndex, float3 hitPosW, float3 hitNormalW) {
    const DevScene& sc = P.sc;
    if (triangleIndex == 0xffffffffu) return 0.f;
    const float sel = P.opt.emissiveSampler == PLTPT_EMISSIVE_SAMPLER_POWER ? __ldg(&sc.eflux[triangleIndex].flux) * sc.invWeightsSum : (sc.activeTriCount > 0 ? 1.f / float(sc.activeTriCount) : 0.f);
    const float3 toLight = hitPosW - posW;
    const float distSqr = dot(toLight, toLight);
    if (distSqr <= kFltMin) return 0.f;
    const float3 L = toLight / sqrtf(distSqr);
    const float cosTheta = dot(hitNormalW, -L);
    if (cosTheta <= 0.f) return 0.f;
    const float area = __uint_as_float(__ldg(&sc.etris[triangleIndex].area));
    return sel * (distSqr / fmaxf(kFltMin, cosTheta * area));
}
// LightSampling.slangh:86-102
PLT_D bool generateEnvMapSample(const FrameParams& P, const SD& sd, Rng& sg, float wavelength, LightSample& ls) {
    if (!P.useEnv) return false;
    EnvSample s;
    envSample(P.sc.env, sg.next2D(), s);
    const float4 wl = hwssWavelengths(wavelength, int(P.opt.HWSS));
    const float4 Le = rgbToSpectrum4(s.Le, wl);
    const float4 L = s.pdf > 0.f ? Le / s.pdf : mk4(.0f);
    ls.beam = sourceBeamDistant(-s.dir, P.opt.sourcingMaxBeamOmega, L, wl, P.opt.sourcingMaxBeamOmega, int(P.opt.HWSS));
    ls.pdf = s.pdf;
    ls.origin = shadowRayOrigin(sd, s.dir);
    ls.distance = kFltMax;  // kRayTMax
    ls.dir = s.dir;
    return anyPositive(L, int(P.opt.HWSS));
}
// LightSampling.slangh:104-126
template <bool kTex = true>
PLT_D bool generateEmissiveSample(const FrameParams& P, const SD& sd, Rng& sg, float wavelength, bool importanceSample, LightSample& ls) {
    if (!P.useEmissive) return false;
    TriangleLightSample tls;
    if (!emissiveSampleLight<kTex>(P, sd.posW, sg, wavelength, importanceSample, tls)) return false;
    const float4 L = tls.pdf > 0.f ? tls.Le / tls.pdf : mk4(.0f);
    ls.pdf = tls.pdf;
    const float3 lightPos = computeRayOrigin(tls.posW, tls.normalW);
    ls.origin = shadowRayOrigin(sd, lightPos - sd.posW);
    const float3 toLight = lightPos - ls.origin;
    ls.distance = length(toLight);
    ls.dir = normalize(toLight);
    ls.beam = sourceBeamArea(tls.posW, -ls.dir, P.opt.sourcingAreaFromEmissiveGeometry, ls.distance, L, tls.wavelengths, P.opt.sourcingMaxBeamOmega, int(P.opt.HWSS));
    return anyPositive(L, int(P.opt.HWSS));
}
// LightHelpers.slang:70-330 (analytic lights)
struct AnalyticSample { float3 posW, normalW, dir; float distance; float4 Li, wavelengths; float pdf; };
PLT_D bool sampleAnalyticLight(const FrameParams& P, float3 shadingPosW, const PltptLightData* light, Rng& sg, float wavelength, bool importanceSample, AnalyticSample& ls) {
    ls.Li = mk4(0.f); ls.wavelengths = mk4(0.f); ls.pdf = 0.f; ls.distance = 0.f;
    const float3 lightPos = mk3(__ldg(&light->posW[0]), __ldg(&light->posW[1]), __ldg(&light->posW[2]));
    const float3 lightDir = mk3(__ldg(&light->dirW[0]), __ldg(&light->dirW[1]), __ldg(&light->dirW[2]));
    const uint32_t type = __ldg(&light->type), specId = __ldg(&light->intensitySpectrumId);
    const int hwss = int(P.opt.HWSS);
    switch (type) {
        case 0: {  // Point (pdf 0: later rejected by ls.pdf > 1e-6, reference quirk SURVEY §8 a10)
            ls.posW = lightPos;
            ls.normalW = lightDir;
            const float3 toLight = ls.posW - shadingPosW;
            const float distSqr = fmaxf(dot(toLight, toLight), 1e-9f);
            ls.distance = sqrtf(distSqr);
            ls.dir = toLight / ls.distance;
            const float cosTheta = -dot(ls.dir, lightDir);
            float falloff = 1.f;
            const float cosOpening = __ldg(&light->cosOpeningAngle), penumbra = __ldg(&light->penumbraAngle);
            if (cosTheta < cosOpening) falloff = 0.f;
            else if (penumbra > 0.f) {
                const float deltaAngle = __ldg(&light->openingAngle) - acosf(cosTheta);
                const float tt = clampf(deltaAngle / penumbra, 0.f, 1.f);
                falloff = tt * tt * (3.f - 2.f * tt);
            }
            ls.Li = sampleSpectrum(P.sc.profiles, specId, wavelength, sg, importanceSample, hwss, ls.wavelengths);
            ls.Li *= falloff / distSqr;
            ls.pdf = 0.f;
            return true;
        }
        case 1:  // Directional
            ls.posW = mk3(INFINITY);
            ls.normalW = lightDir;
            ls.distance = kFltMax;
            ls.dir = -lightDir;
            ls.Li = sampleSpectrum(P.sc.profiles, specId, wavelength, sg, importanceSample, hwss, ls.wavelengths);
            ls.pdf = 0.f;
            return true;
        case 2: {  // Distant
            ls.posW = mk3(INFINITY);
            const float2 u = sg.next2D();
            const float cosSub = __ldg(&light->cosSubtendedAngle);
            const float3 dir = sample_cone(u, cosSub);
            ls.dir = normalize(mulVec(light->transMat, dir));
            ls.normalW = -ls.dir;
            ls.distance = kFltMax;
            const float solidAngle = k2Pi * (1.f - cosSub);
            ls.Li = sampleSpectrum(P.sc.profiles, specId, wavelength, sg, importanceSample, hwss, ls.wavelengths);
            ls.pdf = 1.f / solidAngle;
            return true;
        }
        case 3: case 4: case 5: {  // Rect, Disc, Sphere
            const float2 u = sg.next2D();
            float3 pos, nrm = mk3(0.f, 0.f, 1.f);
            if (type == 3) pos = mk3(u.x * 2.f - 1.f, u.y * 2.f - 1.f, 0.f);
            else if (type == 4) { const float2 d = sample_disk(u); pos = mk3(d.x, d.y, 0.f); }
            else { pos = sample_sphere(u); nrm = pos; }
            ls.posW = mulPoint(light->transMat, pos);
            ls.normalW = normalize(mulVec(light->transMatIT, nrm));
            // finalizeAreaLightSample
            const float3 toLight = ls.posW - shadingPosW;
            const float distSqr = fmaxf(dot(toLight, toLight), 1e-9f);
            ls.distance = sqrtf(distSqr);
            ls.dir = toLight / ls.distance;
            const float cosTheta = dot(ls.normalW, -ls.dir);
            if (cosTheta <= 0.f) return false;
            const float area = __ldg(&light->surfaceArea);
            ls.Li = sampleSpectrum(P.sc.profiles, specId, wavelength, sg, importanceSample, hwss, ls.wavelengths);
            ls.Li *= area * cosTheta / distSqr;
            ls.pdf = distSqr / (cosTheta * area);
            return true;
        }
        default: return false;
    }
}
// LightSampling.slangh:129-166
PLT_D bool generateAnalyticLightSample(const FrameParams& P, const SD& sd, Rng& sg, float wavelength, bool importanceSample, LightSample& ls) {
    const uint32_t lightCount = P.sc.lightCount;
    if (!P.useAnalytic || lightCount == 0) return false;
    const uint32_t lightIndex = min(uint32_t(sg.next1D() * lightCount), lightCount - 1);
    const PltptLightData* light = P.sc.lights + lightIndex;
    AnalyticSample s;
    if (!sampleAnalyticLight(P, sd.posW, light, sg, wavelength, importanceSample, s)) return false;
    const uint32_t type = __ldg(&light->type);
    const bool isDistantEmitter = type == 1 || type == 2;
    const float A = __ldg(&light->surfaceArea);
    const float Omega = 4.f * kPi * (1.f - sqr(__ldg(&light->cosSubtendedAngle)));
    const float4 L = s.Li * float(lightCount);
    ls.pdf = s.pdf / lightCount;
    ls.origin = shadowRayOrigin(sd, s.dir);
    ls.distance = s.distance;
    ls.dir = s.dir;
    if (isDistantEmitter) ls.beam = sourceBeamDistant(-ls.dir, Omega, L, s.wavelengths, P.opt.sourcingMaxBeamOmega, int(P.opt.HWSS));
    else ls.beam = sourceBeamArea(mk3(__ldg(&light->posW[0]), __ldg(&light->posW[1]), __ldg(&light->posW[2])), -ls.dir, A, ls.distance, L, s.wavelengths, P.opt.sourcingMaxBeamOmega, int(P.opt.HWSS));
    return anyPositive(L, int(P.opt.HWSS));
}
PLT_D void clearLightSample(LightSample& ls) {
    ls.pdf = 0.f; ls.origin = mk3(0.f); ls.distance = 0.f; ls.dir = mk3(0.f); ls.lightType = 0;
    ls.beam = makeBeam(mk4(0.f), 0.f, mk3(0.f), mk3(0.f, 0.f, 1.f), mk4(0.f), false, 4);
    ls.beam.r = mk3(0.f); ls.beam.t = mk3(0.f);
}
// LightSampling.slangh:168-206
template <bool kTex = true>
PLT_D bool generateLightSample(const FrameParams& P, const SD& sd, bool sampleReflection, bool sampleTransmission, Rng& sg, Rng& tilesg, float wavelength, bool importanceSample, LightSample& ls) {
    clearLightSample(ls);
    uint32_t lightType;
    float selectionPdf;
    const bool success = P.perTileSG ? selectLightType(P, lightType, selectionPdf, tilesg) : selectLightType(P, lightType, selectionPdf, sg);
    if (!success) return false;
    bool valid = false;
    if (P.useEnv && lightType == LT_EnvMap) valid = generateEnvMapSample(P, sd, sg, wavelength, ls);
    if (P.useEmissive && lightType == LT_Emissive) valid = generateEmissiveSample<kTex>(P, sd, sg, wavelength, importanceSample, ls);
    if (P.useAnalytic && lightType == LT_Analytic) valid = generateAnalyticLightSample(P, sd, sg, wavelength, importanceSample, ls);
    ls.lightType = lightType;
    ls.pdf *= selectionPdf;
    ls.beam.scale(1.f / selectionPdf);
    const bool isReflection = dot(sd.faceN, ls.dir) * dot(sd.faceN, sd.V) > .0f;
    return valid && ls.pdf > 1e-6f && !(!sampleReflection && isReflection) && !(!sampleTransmission && !isReflection);
}

// ---- solve stage ---------------------------------------------------------------------------------------------
// The solve stage is a wavefront (DESIGN.md "Solve stage"):
//   connect : one thread per path walks its recorded vertices in order (the random-number order of pltpt_solve.rt.slang:420-464
//             is sequential per pixel), samples the light, traces the shadow ray / manifold walk, applies MIS and EMITS one
//             connection job per surviving light connection instead of replaying it;
//   unwind  : every job is a generalized ray that still has to be pushed through vertices q, q-1, ... 0 of its path.  Jobs are
//             binned by the MaterialType of the vertex they evaluate next, so a warp runs ONE BSDF; one launch = one replay step;
//   resolve : sums the measured RGB of a pixel's connections in the reference's order (deterministic, no float atomics).
struct SolveState {
    Rng sg, tilesg;
    float3 pathOrigin;
    float wavelength;
    uint32_t evals;
    TraversalCounters tc;
};
enum : uint32_t { JOB_Q_MASK = 0xffu, JOB_COMPS_SHIFT = 8, JOB_DISTANT = 1u << 11, JOB_FIRST_IS_CURRENT = 1u << 12, JOB_BOUNCE_SHIFT = 16, JOB_KIND_SHIFT = 24 };
struct JobHeader { uint32_t pixel, q, bounceIdx, kind; bool firstIsCurrent; float pathContribution; };

PLT_D float4* jobField(const FrameParams& P, uint32_t id, int k) { return P.job + size_t(id) * kJobPlanes + k; }
PLT_D uint32_t jobAux(uint32_t pixel, uint32_t q) { return pixel | (q << 26); }
PLT_D void storeJob(const FrameParams& P, uint32_t id, const Beam& b, const JobHeader& h) {
    float4* J = P.job + size_t(id) * kJobPlanes;
    constexpr size_t st = 1;
    const uint32_t packed = (h.q & JOB_Q_MASK) | (uint32_t(b.comps) << JOB_COMPS_SHIFT) | (b.distant ? JOB_DISTANT : 0u) | (h.firstIsCurrent ? JOB_FIRST_IS_CURRENT : 0u) |
                            (h.bounceIdx << JOB_BOUNCE_SHIFT) | (h.kind << JOB_KIND_SHIFT);
    J[0 * st] = b.L; J[1 * st] = b.L1; J[2 * st] = b.L2; J[3 * st] = b.L3; J[4 * st] = b.wl; J[5 * st] = b.spdf;
    J[6 * st] = mk4(b.Dxx, b.Dxy, b.Dyy, b.r_m);
    J[7 * st] = mk4(b.posW.x, b.posW.y, b.posW.z, b.r.x);
    J[8 * st] = mk4(b.r.y, b.r.z, b.t.x, b.t.y);
    J[9 * st] = mk4(b.t.z, h.pathContribution, __uint_as_float(h.pixel), __uint_as_float(packed));
}
PLT_D void loadJob(const float4* J, Beam& b, JobHeader& h);
PLT_D void loadJob(const FrameParams& P, uint32_t id, Beam& b, JobHeader& h) { loadJob(P.job + size_t(id) * kJobPlanes, b, h); }
// `J`: the job's ten float4 (in global memory, or staged in shared memory by k_unwind_step)
PLT_D void loadJob(const float4* J, Beam& b, JobHeader& h) {
    constexpr size_t st = 1;
    b.L = J[0 * st]; b.L1 = J[1 * st]; b.L2 = J[2 * st]; b.L3 = J[3 * st]; b.wl = J[4 * st]; b.spdf = J[5 * st];
    const float4 c = J[6 * st], d = J[7 * st], e = J[8 * st], f = J[9 * st];
    b.Dxx = c.x; b.Dxy = c.y; b.Dyy = c.z; b.r_m = c.w;
    b.posW = mk3(d.x, d.y, d.z); b.r = mk3(d.w, e.x, e.y); b.t = mk3(e.z, e.w, f.x);
    const uint32_t packed = __float_as_uint(f.w);
    h.pathContribution = f.y; h.pixel = __float_as_uint(f.z);
    h.q = packed & JOB_Q_MASK; b.comps = int((packed >> JOB_COMPS_SHIFT) & 7u); b.distant = (packed & JOB_DISTANT) != 0;
    h.firstIsCurrent = (packed & JOB_FIRST_IS_CURRENT) != 0; h.bounceIdx = (packed >> JOB_BOUNCE_SHIFT) & 0xffu; h.kind = packed >> JOB_KIND_SHIFT;
}
// Appends `id` to the pending list of the next replay iteration (warp-aggregated: one atomic per warp for the slot, one per
// material type for the histogram).
template <bool kAux = true>
PLT_D void enqueueJob(const FrameParams& P, uint32_t type, uint32_t id, uint32_t aux = 0) {
    const uint32_t active = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(active) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(P.queueCount + QC_Total, uint32_t(__popc(active)));
    base = __shfl_sync(active, base, leader);
    const uint32_t slot = base + __popc(active & ((1u << lane) - 1u));
    P.pending[slot] = id;
    if (kAux) P.pendingAux[slot] = aux;
    P.pendingType[slot] = uint8_t(type);
    const uint32_t peers = __match_any_sync(active, type);
    if (lane == __ffs(peers) - 1) atomicAdd(P.queueCount + QC_Hist + type, uint32_t(__popc(peers)));
}
// Queue bin of a replay job that evaluates vertex q of its path next: the vertex's MaterialType, and a separate set of bins for
// q == 0, where the step ends in measure() instead of storing the job back: a warp then does one or the other, not both in turn.
PLT_D uint32_t replayBin(const FrameParams& P, uint32_t pixel, uint32_t q) {
    return uint32_t(P.vMatType[size_t(q) * P.pixelCount + pixel]) | (q == 0u ? kFinishingBin : 0u);
}
PLT_D uint32_t allocJob(const FrameParams& P) {
    const uint32_t active = __activemask();
    const int leader = __ffs(active) - 1, lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(P.jobCount, uint32_t(__popc(active)));
    base = __shfl_sync(active, base, leader);
    return base + __popc(active & ((1u << lane) - 1u));
}

// pltpt_solve.rt.slang:48-66 (forwardUnwindStep)
template <bool kTex = true>
PLT_D void forwardUnwindStep(const FrameParams& P, uint32_t& evals, const VertexData& v, uint32_t materialID, const DevMaterial& m, float3 wo, uint32_t lobe, Beam& beam) {
    const SD sd = prepareShadingData<kTex>(P.sc, v, materialID, m, -beam.r);
    beam.propagate(sd.posW);
    const BsdfCtx cx = {P.sc.profiles, int(P.opt.HWSS)};
    bsdfEval(m, cx, sd, wo, lobe, beam);
    evals += uint32_t(beam.comps);
    if (P.opt.HWSSDoMIS) beam.spdf *= bsdfLambdaPdf(m, cx, sd, wo, lobe, beam, beam.wl);
}
PLT_D bool finite3(float3 v) { return isfinite(v.x) && isfinite(v.y) && isfinite(v.z); }
PLT_D void debugEvent(float4* dbg, float3 val) { if (dbg && finite3(val)) *dbg = mk4(val.x, val.y, val.z, 1.f); }  // writeDebugOutput, pltpt_debug.slangh:20-23
// pltpt_solve.rt.slang:96-147: returns the RGB contribution of one connection (w = 1 unless the direct-light switch skipped it).
// `dbg` (debug views on): receives the event of the selected view that measure() / markDebug() raise for this connection.
PLT_D float4 measure(const FrameParams& P, int bounceIdx, float pathContribution, Beam& beam, float4* dbg = nullptr) {
    if (dbg) *dbg = mk4(0.f);
    if (!P.opt.useDirectLights && bounceIdx <= 1) return mk4(0.f);
    if (dbg && (P.debugView == DV_CoherenceArea || P.debugView == DV_CoherenceAnisotropy)) {  // :100-106, before the propagation to the camera
        const float k = k2Pi / (1e-3f * beam.wl.x);
        const M22 iT = beam.invThetaK(k);
        if (P.debugView == DV_CoherenceArea) debugEvent(dbg, ccMagma(sqrtf(kPi * sqrtf(1.f / (iT.a * iT.d - iT.b * iT.c))) / 150.f));
        else {  // 1 - S00 / S11 of svd_symmetric (MathHelpers.slang:498-520)
            const float a2 = sqr(iT.a), b2 = sqr(iT.c), c2 = sqr(iT.b), d2 = sqr(iT.d);
            const float y = 2 * iT.a * iT.b + 2 * iT.c * iT.d, x = a2 + b2 - c2 - d2;
            const float s1 = a2 + b2 + c2 + d2, s2 = sqrtf(sqr(x) + sqr(y));
            debugEvent(dbg, ccInferno(1.f - sqrtf((s1 + s2) / 2.f) / sqrtf((s1 - s2) / 2.f)));
        }
    }
    beam.propagate(P.camPos);
    if (P.polarizer > .0f) {
        beam.rotateFrame(normalize(P.camU), -normalize(P.camW));
        float s, c;
        sincosf(P.polarizerTheta, &s, &c);
        const float a = 1.f - P.polarizer, h = .5f * P.polarizer;
        PLT_FOR_LAMBDA(j, beam.comps) {
            const float4 SP = beam.getSP(j);
            // (1-p) I + p * MuellerPolarizer(theta), Mueller.slang linear polariser
            beam.setSP(j, mk4(a * SP.x + h * (SP.x + c * SP.y + s * SP.z), a * SP.y + h * (c * SP.x + c * c * SP.y + c * s * SP.z),
                              a * SP.z + h * (s * SP.x + c * s * SP.y + s * s * SP.z), a * SP.w));
        }
    }
    if (P.opt.HWSSDoMIS) {
        float sum = 0.f;
        PLT_FOR_LAMBDA(j, int(P.opt.HWSS)) sum += lane(beam.spdf, j);
        pathContribution *= sum > kFltEps ? beam.spdf.x / sum : .0f;
    } else pathContribution *= 1.f / float(beam.comps);
    const float3 Lrgb = vmax(mk3(.0f), spectrumToRgb(beam.L, beam.wl, beam.comps) * pathContribution);
    if (dbg) {  // markDebug, pltpt_debug.slangh:31-51 (the NaN / inf views need the running pixel sum: k_resolve raises them)
        if (P.debugView == DV_PathLength) debugEvent(dbg, ccMagma(float(bounceIdx) / float(int(P.opt.maxBounces) - 1)));
        else if (P.debugView == DV_HwssComps) debugEvent(dbg, ccMagma(float(beam.comps) / float(P.opt.HWSS)));
        else if (P.debugView == DV_Polarization) {
            const float4 S = mk4(beam.L.x, beam.L1.x, beam.L2.x, beam.L3.x);
            const float cp = sqrtf(sqr(S.w)) / S.x, lp = sqrtf(sqr(S.y) + sqr(S.z)) / S.x, lpd = .5f * atan2f(S.z, S.y);
            debugEvent(dbg, mk3(lp * sqr(cosf(lpd)), cp, lp * sqr(sinf(lpd))));
        }
    }
    return finite3(Lrgb) ? mk4(Lrgb.x, Lrgb.y, Lrgb.z, 1.f) : mk4(0.f, 0.f, 0.f, 1.f);
}
PLT_D void writeContribution(const FrameParams& P, uint32_t pixel, uint32_t bounceIdx, uint32_t kind, float3 rgb) {
    P.contrib[size_t(2 * bounceIdx + kind) * P.pixelCount + pixel] = mk4(rgb.x, rgb.y, rgb.z, 0.f);
}
PLT_D void writeDebug(const FrameParams& P, uint32_t pixel, uint32_t bounceIdx, uint32_t kind, float4 ev) {
    if (ev.w != 0.f) P.debug[size_t(2 * bounceIdx + kind) * P.pixelCount + pixel] = ev;
}
// measure() of a finished connection into its contribution (and debug) slot
// The debug-view variant is kept out of line: with `dbg == nullptr` known at compile time the inlined measure() of the hot path
// carries none of the view code (inlined both ways it cost k_unwind_step 5 %).
static __device__ __noinline__ void measureIntoDebug(const FrameParams& P, uint32_t pixel, uint32_t bounceIdx, uint32_t kind, float pathContribution, Beam beam) {  // beam by value: the caller's copy never has its address taken
    float4 ev;
    const float4 c = measure(P, int(bounceIdx), pathContribution, beam, &ev);
    P.contrib[size_t(2 * bounceIdx + kind) * P.pixelCount + pixel] = c;
    writeDebug(P, pixel, bounceIdx, kind, ev);
}
PLT_D void measureInto(const FrameParams& P, uint32_t pixel, uint32_t bounceIdx, uint32_t kind, float pathContribution, Beam& beam) {
    if (P.debugView) { measureIntoDebug(P, pixel, bounceIdx, kind, pathContribution, beam); return; }
    P.contrib[size_t(2 * bounceIdx + kind) * P.pixelCount + pixel] = measure(P, int(bounceIdx), pathContribution, beam, nullptr);
}
// Hands a sourced beam over to the unwind kernels (or measures it at once when nothing is left to replay).
PLT_D void emitConnection(const FrameParams& P, uint32_t pixel, int bounceIdx, uint32_t kind, bool evalCurrentVertex, float pathContribution, Beam& beam) {
    if (!evalCurrentVertex && bounceIdx == 0) { measureInto(P, pixel, uint32_t(bounceIdx), kind, pathContribution, beam); return; }
    const uint32_t id = allocJob(P);
    if (id >= P.jobCapacity) { atomicAdd(P.queueCount + QC_Dropped, 1u); writeContribution(P, pixel, bounceIdx, kind, mk3(0.f)); return; }
    JobHeader h;
    h.pixel = pixel; h.bounceIdx = uint32_t(bounceIdx); h.kind = kind; h.firstIsCurrent = evalCurrentVertex; h.pathContribution = pathContribution;
    h.q = uint32_t(evalCurrentVertex ? bounceIdx : bounceIdx - 1);
    storeJob(P, id, beam, h);
    enqueueJob(P, replayBin(P, pixel, h.q), id, jobAux(pixel, h.q));
}
// pltpt_solve.rt.slang:205-217
PLT_D bool wasNEEAvailable(const FrameParams& P, int bounceIdx, uint32_t flags) {
    return P.doNEE && bounceIdx > 0 && (!(flags & RF_LastBounceWasDelta) || (P.doMS && (flags >> 16) <= P.opt.MNEEMaxOccluders && bounceIdx > 1));
}
PLT_D float misWeight(float p0, float p1) { return p0 / (p0 + p1); }

// pltpt_solve.rt.slang:152-200
PLT_D void evalEmissive(const FrameParams& P, SolveState& st, uint32_t pixel, int bounceIdx, const Bounce& bounce, const DevMaterial& m, const SD& sd, const TriHit& th) {
    const bool didNEE = wasNEEAvailable(P, bounceIdx, bounce.flags);
    if (!(P.useEmissive && (m.propFlags & P_IsEmissive) && (P.doMIS || !didNEE))) { writeContribution(P, pixel, bounceIdx, 1, mk3(0.f)); return; }
    const bool needMIS = P.doMIS && didNEE;
    const uint32_t triIdx = getEmissiveTriangleIndex(P.sc, th.instanceID, th.primitiveIndex);
    const float3 hitN = sd.frontFacing ? sd.faceN : -sd.faceN;
    float p[3];
    lightTypeProbabilities(P, p);
    const float lightPdf = p[1] * emissiveEvalPdf(P, st.pathOrigin, triIdx, sd.posW, hitN);
    const float w = needMIS ? misWeight(bounce.lastNonDeltaPdf, lightPdf) : 1.f;
    float4 wl = mk4(.0f);
    float4 Le = sampleSpectrum(P.sc.profiles, m.emissionId, st.wavelength, st.sg, P.opt.doImportanceSampleEmitters != 0, int(P.opt.HWSS), wl);
    Le *= w * rgbToSpectrum4(mk3(m.c1x, m.c1y, m.c1z), wl);  // bsdfProp.emissionRGB (PLTDiffuseBSDF.slang:143), 1 without an emissive texture
    if (!anyPositive(Le, int(P.opt.HWSS))) { writeContribution(P, pixel, bounceIdx, 1, mk3(0.f)); return; }
    const float dist = length(st.pathOrigin - sd.posW);
    Beam beam = sourceBeamArea(sd.posW, bounce.wi, P.opt.sourcingAreaFromEmissiveGeometry, dist, Le, wl, P.opt.sourcingMaxBeamOmega, int(P.opt.HWSS));
    if (!P.opt.HWSSDoMIS && (bounce.flags & RF_HasDispersiveDeltaSegment)) beam.comps = 1;
    emitConnection(P, pixel, bounceIdx, 1, false, bounce.pathContribution, beam);
}

// pltpt_solve.rt.slang:371-396 (onMiss): the environment seen by a path that left the scene at vertex index `p`
PLT_D void onMiss(const FrameParams& P, uint32_t pixel, int p) {
    const Bounce bounce = readBounce(P, pixel, p);
    const float3 wo = bounce.wi;
    const bool didNEE = wasNEEAvailable(P, p, bounce.flags);
    if (!(P.useEnv && (P.doMIS || !didNEE))) { writeContribution(P, pixel, uint32_t(p), 1, mk3(0.f)); return; }
    const bool needMIS = P.doMIS && didNEE;
    float prob[3];
    lightTypeProbabilities(P, prob);
    const float lightPdf = prob[0] * envEvalPdf(P.sc.env, -wo);
    const float w = needMIS ? misWeight(bounce.lastNonDeltaPdf, lightPdf) : 1.f;
    const float4 wl = hwssWavelengths(P.psOrigin[pixel].w, int(P.opt.HWSS));  // hero wavelength of the path
    const float4 Le = rgbToSpectrum4(envEval(P.sc.env, -wo), wl) * w;
    Beam beam = sourceBeamDistant(wo, P.opt.sourcingMaxBeamOmega, Le, wl, P.opt.sourcingMaxBeamOmega, int(P.opt.HWSS));
    if (!P.opt.HWSSDoMIS && (bounce.flags & RF_HasDispersiveDeltaSegment)) beam.comps = 1;
    emitConnection(P, pixel, p, 1, false, bounce.pathContribution, beam);
}

// ---- MNEE (Manifolds.slangh) ------------------------------------------------------------------------------------
constexpr int kMaxMSHits = 5;  // MNEEMaxOccluders (<= 4, validated in pltpt_set_options) + the current vertex
struct MSHit { uint4 hit; float3 posW; uint32_t lobe; };
struct ManifoldSampleData {
    MSHit hits[kMaxMSHits];
    float3 woLastHitToLight;
    int hitCount;
    uint32_t iterations;
    float wavelength, pdf;
    bool dispersive;
};
// Manifolds.slangh:95-108
PLT_D bool manifoldSampleVertex(uint4 hitInfo, const SD& sd, const DevMaterial& m, float3 lightDir, Rng& sg, ManifoldSampleData& data) {
    Sample s;
    if (!bsdfSample(m, sd, sg, s, data.wavelength, L_Delta, lightDir)) return false;
    if (data.hitCount >= kMaxMSHits) return false;
    data.pdf *= s.pdf;
    data.hits[data.hitCount].lobe = s.lobe;
    data.hits[data.hitCount].hit = hitInfo;
    data.hits[data.hitCount].posW = sd.posW;
    data.dispersive |= (s.lobe & L_DispersiveDelta) != 0;
    ++data.hitCount;
    return true;
}
// Manifolds.slangh:110-160
// `first` (optional): the closest hit of the first segment, already found by the shadow-ray trace kernel.
template <bool kTex = true>
PLT_D bool traceMNEEOccluders(const FrameParams& P, TraversalCounters& tc, float3 rayOrigin, float3 rayDir, float rayTMin, float rayTMax, Rng& sg, int maxMSBounces, ManifoldSampleData& data,
                              const TriHit* first = nullptr, float firstT = 0.f) {
    const float3 wi = -rayDir;
    for (bool useFirst = first != nullptr;; useFirst = false) {
        TriHit th;
        float t;
        if (useFirst) { th = *first; t = firstT; }
        else if (!traceClosestHit(P, tc, rayOrigin, rayDir, rayTMin, rayTMax, th, t)) return true;
        if (data.hitCount >= int(P.opt.MNEEMaxOccluders) || data.hitCount >= maxMSBounces) return false;
        uint32_t materialID;
        VertexData v = getVertexData(P.sc, th, materialID);
        v.posW = rayOrigin + t * rayDir;
        const DevMaterial m = loadMaterialAt<kTex>(P.sc, materialID, v.texC);
        const SD sd = prepareShadingData<kTex>(P.sc, v, materialID, m, wi);
        if (!(m.propFlags & P_CanMSTransmissive)) return false;
        if (!manifoldSampleVertex(packHit(P.sc, th), sd, m, -wi, sg, data) || (data.hits[data.hitCount - 1].lobe & L_Transmission) == 0) return false;
        rayOrigin = shadowRayOrigin(sd, rayDir);
        rayTMax -= t;
    }
}
// Manifolds.slangh:284-347
PLT_D float2 sphcoords(float3 w) {
    const float theta = acosf(clampf(w.z, -1.f, 1.f));
    float phi = atan2f(w.y, w.x);
    phi += phi < .0f ? k2Pi : .0f;
    return mk2(theta, phi);
}
PLT_D M22 sphDeriv(float3 v, float3 ddu, float3 ddv) {
    const float dAcos = -1.f / sqrtf(fmaxf(.0f, 1.f - sqr(v.z)));
    const float invNorm = 1.f / (v.x * v.x + v.y * v.y);
    return {dAcos * ddu.z, dAcos * ddv.z, invNorm * (v.x * ddu.y - v.y * ddu.x), invNorm * (v.x * ddv.y - v.y * ddv.x)};
}
PLT_D bool manifoldOneStep(const V3pdv& w, const V3kpdv& ws, const V3pdv& p, float& lenC, float3& positionStepDir) {
    const float2 spho = sphcoords(w.v), sphio = sphcoords(ws.v);
    const float dt = spho.x - sphio.x;
    float dp = spho.y - sphio.y;
    dp = fabsf(dp) > kPi ? dp - signf(dp) * k2Pi : dp;
    const float2 C = mk2(dt, dp);
    lenC = sqrtf(dot(C, C));
    const M22 dO = sphDeriv(w.v, w.ddu, w.ddv), dIO = sphDeriv(ws.v, ws.ddu, ws.ddv);
    const M22 dCdX = {dO.a - dIO.a, dO.b - dIO.b, dO.c - dIO.c, dO.d - dIO.d};
    M22 inv;
    if (!tryInvert(dCdX, inv)) return false;
    const float2 dX = mul(inv, C);
    positionStepDir = p.ddu * dX.x + p.ddv * dX.y;
    return true;
}
// Manifolds.slangh:80-93
template <bool kTex = true>
PLT_D void manifoldForwardUnwind(const FrameParams& P, SolveState& st, const ManifoldSampleData& r, float3 origin, Beam& beam) {
    for (int idx = r.hitCount - 1; idx >= 0; --idx) {
        const TriHit th = unpackHit(P.sc, r.hits[idx].hit);
        uint32_t materialID;
        VertexData v = getVertexData(P.sc, th, materialID);
        v.posW = r.hits[idx].posW;
        const float3 p2 = idx == 0 ? origin : r.hits[idx - 1].posW;
        const DevMaterial m = loadMaterialAt<kTex>(P.sc, materialID, v.texC);
        forwardUnwindStep<kTex>(P, st.evals, v, materialID, m, normalize(p2 - v.posW), r.hits[idx].lobe, beam);
    }
}

// Deferred manifold walk: everything k_mnee needs to finish one NEE connection whose shadow ray crossed delta occluders.
struct alignas(16) MneeItem {
    uint4 hit[kMaxMSHits - 1];
    float4 posLobe[kMaxMSHits - 1];  // posW.xyz | asfloat(lobe)
    float4 originPdf;                // manifold origin.xyz | ms.pdf
    float4 lightPosWl;               // light position.xyz | hero wavelength of the sampled beam
    float4 lightDirLsPdf;            // direction towards the light.xyz | ls.pdf
    uint32_t pixel, bounceIdx, jobId, flags;  // flags: hitCount[0:3] dispersive 4 distant 5 currentVertexMS 6 doMis 7
};
enum : uint32_t { MI_DISPERSIVE = 1u << 4, MI_DISTANT = 1u << 5, MI_CURRENT_VERTEX_MS = 1u << 6, MI_DO_MIS = 1u << 7 };

enum : uint32_t { MI_WALK_OK = 1u << 8 };
PLT_D void mneeUnpack(const MneeItem& it, ManifoldSampleData& ms) {
    ms.hitCount = int(it.flags & 0xfu); ms.iterations = 0; ms.pdf = it.originPdf.w; ms.dispersive = (it.flags & MI_DISPERSIVE) != 0;
    ms.wavelength = it.lightPosWl.w; ms.woLastHitToLight = mk3(0.f);
    for (int i = 0; i < kMaxMSHits - 1; ++i) {
        ms.hits[i].hit = it.hit[i];
        ms.hits[i].posW = mk3(it.posLobe[i].x, it.posLobe[i].y, it.posLobe[i].z);
        ms.hits[i].lobe = __float_as_uint(it.posLobe[i].w);
    }
}
// ---- the manifold walk as a wavefront (k_mnee_walk) -------------------------------------------------------------------------
// Manifolds.slangh:165-281 after the occluder chain has been collected (traceMNEEOccluders).  It consumes no random numbers, so
// it runs in its own kernel.  The reference's nested loops
//     while (!success) { success = true; for (idx ...) { newton step; if (|C| < eps) continue; success = false; for (;;) { trial ray ... } } }
// are cut at the trial ray: mneeWalkAdvance() consumes the answer of the previous trial ray (accept / shrink the step), runs Newton
// steps until the next trial ray is due, and posts it; the rays of all live walks are then traced together.  One round = one
// trial ray per live walk, so every warp executes either Newton steps or traversal, never a mix, and walks that ended drop
// out of the list (walk lengths range from 1 to MNEEMaxIterations trial rays).  Arithmetic and its order are the reference's.
struct alignas(16) MneeWalk {
    float4 beta;        // line-search step scale per chain vertex
    float4 stepInst;    // positionStepDir.xyz | asfloat(instance of the vertex the line search moves)
    float4 ray0, ray1;  // posted trial ray: p0 | tmin, direction | tmax
    uint4 hit;          // its closest hit: instanceID (0xffffffff = none), primitiveIndex, asuint(u), asuint(v)
    float t;
    uint32_t state;     // idx[0:3] | passSuccess << 4
    uint32_t iterations, pad;
};
enum { WALK_CONTINUE = 0, WALK_CONVERGED = 1, WALK_FAILED = -1 };
PLT_D float betaGet(const float4& b, int i) { return lane(b, i); }
template <bool kTex = true>
PLT_D int mneeWalkAdvance(const FrameParams& P, MneeItem& it, MneeWalk& w, bool first) {
    const int hitCount = int(it.flags & 0xfu);
    const float3 origin = mk3(it.originPdf.x, it.originPdf.y, it.originPdf.z), lightPos = mk3(it.lightPosWl.x, it.lightPosWl.y, it.lightPosWl.z);
    const float3 lightDir = mk3(it.lightDirLsPdf.x, it.lightDirLsPdf.y, it.lightDirLsPdf.z);
    const bool distant = (it.flags & MI_DISTANT) != 0;
    const float wavelength = it.lightPosWl.w;
    int idx;
    bool passSuccess, lineSearch = false;
    float4 beta;
    float3 stepDir = mk3(0.f);
    uint32_t hitInstance = 0, iterations;
    if (first) { idx = 0; passSuccess = true; beta = mk4(1.f); iterations = 0; }
    else {
        idx = int(w.state & 0xfu); passSuccess = (w.state & 16u) != 0; beta = w.beta; iterations = w.iterations;
        stepDir = mk3(w.stepInst.x, w.stepInst.y, w.stepInst.z); hitInstance = __float_as_uint(w.stepInst.w);
        // answer of the posted trial ray (Manifolds.slangh:236-250)
        const uint4 ph = w.hit;
        if (ph.x == 0xffffffffu || hitInstance != ph.x) { setLane(beta, idx, betaGet(beta, idx) / 4.f); lineSearch = true; }
        else {
            TriHit th;
            th.instanceID = ph.x; th.primitiveIndex = ph.y; th.u = __uint_as_float(ph.z); th.v = __uint_as_float(ph.w);
            const float3 p0 = mk3(w.ray0.x, w.ray0.y, w.ray0.z), dir = mk3(w.ray1.x, w.ray1.y, w.ray1.z);
            const float3 np = p0 + dir * w.t;
            it.hit[idx] = packHit(P.sc, th);
            it.posLobe[idx] = mk4(np.x, np.y, np.z, it.posLobe[idx].w);
            setLane(beta, idx, fminf(1.0f, 2.f * betaGet(beta, idx)));
            idx += 1;
        }
    }
    while (!lineSearch) {  // Newton steps until a vertex has to move
        if (idx == hitCount) {  // a pass over the chain is complete
            if (passSuccess) { w.iterations = iterations; return WALK_CONVERGED; }
            passSuccess = true; idx = 0;
        }
        const TriHit th = unpackHit(P.sc, it.hit[idx]);
        uint32_t materialID;
        const VertexData v = getVertexData(P.sc, th, materialID);
        const DiffGeom dsg = differentialShadingGeometry(P.sc, th);
        const float4 prevP = it.posLobe[idx > 0 ? idx - 1 : 0], nextP = it.posLobe[idx + 1 < hitCount ? idx + 1 : idx];
        const float3 p0 = idx == 0 ? origin : mk3(prevP.x, prevP.y, prevP.z);
        const V3pdv pdvp0 = dsg.pdw(p0);
        const float3 p2 = idx + 1 == hitCount ? lightPos : mk3(nextP.x, nextP.y, nextP.z);
        V3pdv pdvp2;
        pdvp2.v = lightDir; pdvp2.ddu = pdvp2.ddv = mk3(0.f);
        if (!distant) pdvp2 = dsg.pdw(p2);
        const bool woIsIn = dot(pdvp2.v, dsg.n.v) < .0f;
        const V3pdv wiP = woIsIn ? pdvp0 : pdvp2;
        const V3pdv woP = woIsIn ? pdvp2 : pdvp0;
        const DevMaterial m = loadMaterialAt<kTex>(P.sc, materialID, v.texC);
        const SD sd = prepareShadingData<kTex>(P.sc, v, materialID, m, wiP.v);
        const V3kpdv wos = bsdfPartialDerivatives(m, sd, wiP, dsg, __float_as_uint(it.posLobe[idx].w), wavelength);
        float lenC;
        float3 positionStepDir;
        if (allZero(wos.v) || !manifoldOneStep(woP, wos, dsg.p, lenC, positionStepDir)) { w.iterations = iterations; return WALK_FAILED; }
        if (lenC < P.opt.MNEESolverThreshold) { idx += 1; continue; }
        passSuccess = false;
        lineSearch = true;
        stepDir = positionStepDir;
        hitInstance = th.instanceID;
    }
    // next trial ray of the line search at vertex idx (Manifolds.slangh:226-235)
    if (++iterations > P.opt.MNEEMaxIterations) { w.iterations = iterations; return WALK_FAILED; }
    const float4 prevP = it.posLobe[idx > 0 ? idx - 1 : 0], curP = it.posLobe[idx];
    const float3 p0 = idx == 0 ? origin : mk3(prevP.x, prevP.y, prevP.z);
    const float3 proposal = mk3(curP.x, curP.y, curP.z) - betaGet(beta, idx) * stepDir;
    float3 dirToProposal = proposal - p0;
    const float dist = length(dirToProposal);
    dirToProposal /= dist;
    w.beta = beta;
    w.stepInst = mk4(stepDir.x, stepDir.y, stepDir.z, __uint_as_float(hitInstance));
    w.ray0 = mk4(p0.x, p0.y, p0.z, fmaxf(kShadowRayTMin, .99f * dist));
    w.ray1 = mk4(dirToProposal.x, dirToProposal.y, dirToProposal.z, 1.01f * dist);
    w.state = uint32_t(idx) | (passSuccess ? 16u : 0u);
    w.iterations = iterations;
    return WALK_CONTINUE;
}

// k_mnee_finish (Manifolds.slangh:270-281, pltpt_solve.rt.slang:320-361): final visibility of the converged chain, replay of the
// beam through the found specular vertices, MIS; then the connection joins the unwind queues like any other.
template <bool kTex = true>
PLT_D void mneeFinish(const FrameParams& P, const MneeItem& it, uint32_t& evals, TraversalCounters& tc) {
    ManifoldSampleData ms;
    mneeUnpack(it, ms);
    const float3 origin = mk3(it.originPdf.x, it.originPdf.y, it.originPdf.z), lightPos = mk3(it.lightPosWl.x, it.lightPosWl.y, it.lightPosWl.z);
    const float3 lightDir = mk3(it.lightDirLsPdf.x, it.lightDirLsPdf.y, it.lightDirLsPdf.z);
    const bool distant = (it.flags & MI_DISTANT) != 0, currentVertexMS = (it.flags & MI_CURRENT_VERTEX_MS) != 0;
    {
        const float3 p = ms.hits[ms.hitCount - 1].posW;
        float3 dir;
        float dist;
        if (!distant) { dir = lightPos - p; dist = length(dir); dir /= dist; }
        else { dir = lightDir; dist = kFltMax; }
        ms.woLastHitToLight = dir;
        if (occluded(P, tc, p, dir, kShadowRayTMin, dist)) { writeContribution(P, it.pixel, it.bounceIdx, 0, mk3(0.f)); return; }
        if (P.debugView == DV_MneeBounces) { float4 ev = mk4(0.f); debugEvent(&ev, ms.hitCount == 1 ? mk3(0.f, 0.3f, 1.f) : mk3(1.f, 0.f, 0.f)); writeDebug(P, it.pixel, it.bounceIdx, 0, ev); }
    }
    Beam beam;
    JobHeader h;
    loadJob(P, it.jobId, beam, h);
    float3 lsDir = -beam.r;
    if (!P.opt.HWSSDoMIS && ms.dispersive) beam.comps = 1;
    if (!beam.distant) {
        lsDir = ms.woLastHitToLight;
        beam.r = -ms.woLastHitToLight;
        beam.t = perp_stark(beam.r);
    }
    beam.scale(1.f / ms.pdf);
    SolveState st;
    st.evals = 0;
    manifoldForwardUnwind<kTex>(P, st, ms, origin, beam);
    evals += st.evals;
    if (it.flags & MI_DO_MIS) {
        float misBsdfPdf;
        if (currentVertexMS) misBsdfPdf = P.bPdf[size_t(it.bounceIdx) * P.pixelCount + it.pixel].y;  // bounce.lastNonDeltaPdf
        else {
            uint32_t materialID;
            const VertexData v = readVertexCache(P, it.pixel, int(it.bounceIdx), materialID);
            const DevMaterial m = loadMaterialAt<kTex>(P.sc, materialID, v.texC);
            const float4 d = P.bDir[size_t(it.bounceIdx) * P.pixelCount + it.pixel];
            const SD sd = prepareShadingData<kTex>(P.sc, v, materialID, m, oct_to_ndir_equal_area_snorm(mk2(d.x, d.y)));
            misBsdfPdf = bsdfPdf(m, sd, lsDir, beam.wl.x);
        }
        beam.scale(misWeight(it.lightDirLsPdf.w, misBsdfPdf * ms.pdf));
    }
    h.firstIsCurrent = !currentVertexMS;
    h.q = h.firstIsCurrent ? it.bounceIdx : it.bounceIdx - 1u;
    storeJob(P, it.jobId, beam, h);
    enqueueJob(P, replayBin(P, it.pixel, h.q), it.jobId, jobAux(it.pixel, h.q));
}

// ---- connect wavefront: A (light sampling) -> shadow-ray trace kernel -> B (visibility, MIS, emit) per vertex index ------------
// Per work-list entry (one recorded vertex of one path) the two halves hand over through a ConnItem, a ray and a job slot that
// already holds the sourced beam.
// ls.dir.xyz | ls.pdf ; jobId, flags, asuint(ms.pdf), CI_CURRENT_MS ? lobe of the current-vertex delta sample : asuint(MIS weight of
// the unoccluded connection) — half A already knows that weight (it does not depend on visibility), so for the common outcome
// "shadow ray answered, nothing special about the vertex" half B finishes from this item alone, without rebuilding the vertex.
struct ConnItem { float4 dirPdf; uint4 info; };
enum : uint32_t { CI_TRACE = 1u, CI_ANY_HIT = 2u, CI_DO_MIS = 4u, CI_CURRENT_MS = 8u, CI_DISPERSIVE = 16u, CI_EMISSIVE = 32u };

struct ConnectCtx {  // what both halves rebuild for a vertex
    SolveState st;
    VertexData v;
    Bounce bounce;
    DevMaterial m;
    SD sd;
    uint32_t materialID;
};
template <bool kTex = true>
PLT_D void connectSetup(const FrameParams& P, uint32_t pixel, int p, bool freshRng, ConnectCtx& c) {
    uint32_t px, py;
    pixelCoords(P, pixel, px, py);
    c.st.evals = 0;
    c.st.tc.nodes = c.st.tc.tris = 0;
    c.v = readVertexCache(P, pixel, p, c.materialID);
    if (p == 0) c.st.pathOrigin = P.camPos;
    else { const float4 prev = P.vc[4 * (size_t(p - 1) * P.pixelCount + pixel)]; c.st.pathOrigin = mk3(prev.x, prev.y, prev.z); }
    if (freshRng) {  // pltpt_solve.rt.slang:410-415
        c.st.sg = Rng::seed(px, py, P.frameCount);
        c.st.wavelength = sampleHeroWavelength(c.st.sg.next1D());
        if (P.perTileSG) c.st.tilesg = Rng::seed(px / P.opt.tileSize, py / P.opt.tileSize, P.frameCount << 2);
    } else {
        const uint4 r = P.ssRng[pixel];
        c.st.sg = Rng{r.x, r.y, r.z, r.w};
        c.st.wavelength = P.psOrigin[pixel].w;  // hero wavelength: the first draw of both stages' generators
        if (P.perTileSG) { const uint4 tr = P.ssTileRng[pixel]; c.st.tilesg = Rng{tr.x, tr.y, tr.z, tr.w}; }
    }
    if (!P.perTileSG) c.st.tilesg = c.st.sg;
    c.bounce = readBounce(P, pixel, p);
    c.m = loadMaterialAt<kTex>(P.sc, c.materialID, c.v.texC);
    c.sd = prepareShadingData<kTex>(P.sc, c.v, c.materialID, c.m, c.bounce.wi);
}
PLT_D void storeSolveRng(const FrameParams& P, uint32_t pixel, const SolveState& st) {
    P.ssRng[pixel] = make_uint4(st.sg.s0, st.sg.s1, st.sg.s2, st.sg.s3);
    if (P.perTileSG) P.ssTileRng[pixel] = make_uint4(st.tilesg.s0, st.tilesg.s1, st.tilesg.s2, st.tilesg.s3);
}
// tail of the per-vertex loop body of pltpt_solve.rt.slang:420-464: emissive hit, then hand the generator to the next vertex
PLT_D void connectTail(const FrameParams& P, uint32_t pixel, int p, ConnectCtx& c) {
    evalEmissive(P, c.st, pixel, p, c.bounce, c.m, c.sd, unpackHit(P.sc, c.bounce.hit));
    if (!(c.bounce.flags & RF_Terminated) && p < int(P.opt.maxBounces)) storeSolveRng(P, pixel, c.st);
}

// geometry views of the first vertex, pltpt_solve.rt.slang:450-457
static __device__ __noinline__ void firstVertexDebugViews(const FrameParams& P, uint32_t pixel, float3 N, float3 albedo, float roughness, float2 uv) {
    float4 ev = mk4(0.f);
    if (P.debugView == DV_Normals) debugEvent(&ev, N);
    else if (P.debugView == DV_Albedo) debugEvent(&ev, albedo);
    else if (P.debugView == DV_Roughness) debugEvent(&ev, mk3(roughness));
    else if (P.debugView == DV_UVs) debugEvent(&ev, mk3(uv.x, 0.f, uv.y));
    writeDebug(P, pixel, 0, 1, ev);
}
// Half A (pltpt_solve.rt.slang:245-300): light sample, beam sourcing, delta sample of the current vertex; posts the shadow ray.
template <bool kTex = true>
PLT_D void connectA(const FrameParams& P, uint32_t slot, uint32_t pixel, uint32_t& evals, TraversalCounters& tc) {
    const int p = int(P.level);
    ConnectCtx c;
    connectSetup<kTex>(P, pixel, p, p == 0, c);
    if (P.debugView && p == 0) firstVertexDebugViews(P, pixel, c.sd.N, mk3(c.m.albx, c.m.alby, c.m.albz), c.m.roughnessProp, c.sd.uv);
    ConnItem item;
    item.dirPdf = mk4(0.f); item.info = make_uint4(0, 0, 0, 0);
    bool pending = false;
    if (P.doNEE) {
        const uint32_t lobes = c.m.lobes;
        const bool allowMSOnCurrentVertex = P.doMS && p > 0 && !(c.bounce.flags & RF_LastBounceWasDelta) && (c.m.propFlags & P_CanMSReflective);
        const bool sampleReflection = (lobes & L_NonDeltaReflection) != 0 || allowMSOnCurrentVertex;
        const bool sampleTransmission = (lobes & L_NonDeltaTransmission) != 0;
        if (sampleReflection || sampleTransmission) {
            LightSample ls;
            const bool valid = generateLightSample<kTex>(P, c.sd, sampleReflection, sampleTransmission, c.st.sg, c.st.tilesg, c.st.wavelength, P.opt.doImportanceSampleEmitters != 0, ls);
            if (valid && anyPositive(ls.beam.L, int(P.opt.HWSS))) {
                if (!P.opt.HWSSDoMIS && (c.bounce.flags & RF_HasDispersiveDeltaSegment)) ls.beam.comps = 1;
                uint32_t flags = CI_TRACE | ((P.doMIS && ls.lightType != LT_Analytic) ? CI_DO_MIS : 0u);
                float msPdf = 1.f;
                uint32_t handover = 0;  // ConnItem::info.w: lobe of the current-vertex delta sample, or the MIS weight of the unoccluded connection
                bool alive = true;
                if (!P.doMS || int(P.opt.maxBounces) - p <= 0) flags |= CI_ANY_HIT;
                else {
                    const float3 lightDir = -ls.beam.r;
                    const bool isReflection = dot(c.sd.N, lightDir) * dot(c.sd.N, c.sd.V) > .0f;
                    if (isReflection && allowMSOnCurrentVertex) {
                        ManifoldSampleData ms;
                        ms.hitCount = 0; ms.iterations = 0; ms.pdf = 1.f; ms.dispersive = false; ms.wavelength = ls.beam.wl.x;
                        alive = manifoldSampleVertex(c.bounce.hit, c.sd, c.m, lightDir, c.st.sg, ms) && (ms.hits[0].lobe & L_Reflection) != 0;
                        flags |= CI_CURRENT_MS | (ms.dispersive ? CI_DISPERSIVE : 0u);
                        msPdf = ms.pdf; handover = ms.hits[0].lobe;
                    }
                }
                if (alive) {
                    if ((flags & CI_DO_MIS) && !(flags & CI_CURRENT_MS)) handover = __float_as_uint(misWeight(ls.pdf, bsdfPdf(c.m, c.sd, ls.dir, ls.beam.wl.x)));
                    if (P.useEmissive && (c.m.propFlags & P_IsEmissive)) flags |= CI_EMISSIVE;
                    const uint32_t id = allocJob(P);
                    if (id < P.jobCapacity) {
                        JobHeader h;
                        h.pixel = pixel; h.bounceIdx = uint32_t(p); h.kind = 0; h.firstIsCurrent = true; h.pathContribution = c.bounce.pathContribution; h.q = uint32_t(p);
                        storeJob(P, id, ls.beam, h);
                        item.dirPdf = mk4(ls.dir.x, ls.dir.y, ls.dir.z, ls.pdf);
                        item.info = make_uint4(id, flags, __float_as_uint(msPdf), handover);
                        P.shadowRays[2 * size_t(slot)] = mk4(ls.origin.x, ls.origin.y, ls.origin.z, 0.f);
                        P.shadowRays[2 * size_t(slot) + 1] = mk4(ls.dir.x, ls.dir.y, ls.dir.z, ls.distance);
                        pending = true;
                    } else atomicAdd(P.queueCount + QC_Dropped, 1u);
                }
            }
        }
    }
    P.connItems[slot] = item;
    if (pending) storeSolveRng(P, pixel, c.st);  // half B continues the stream (occluder samples, emissive hit)
    else {
        writeContribution(P, pixel, uint32_t(p), 0, mk3(0.f));
        connectTail(P, pixel, p, c);
    }
    evals += c.st.evals;
    tc.nodes += c.st.tc.nodes; tc.tris += c.st.tc.tris;
}

// Half B (pltpt_solve.rt.slang:283-361): visibility / occluder chain, MIS weight, hand-over to the unwind queues (or to k_mnee).
// Returns true when the entry needs the full half B (connectB): it is then queued for k_connect_b_slow, so that the few entries
// with an occluder chain / delta sample / emissive hit run densely packed instead of stalling 31 finished lanes each.
// Called by the shadow-ray kernels right after the ray of `slot` has been traced (`blocked`, `occluderInstance`: its answer).
PLT_D bool connectBFast(const FrameParams& P, uint32_t pixel, const ConnItem& item, bool blocked, uint32_t occluderInstance) {
    const uint32_t flags = item.info.y;
    const int p = int(P.level);
    const uint32_t jobId = item.info.x;
    // Fast path (pltpt_solve.rt.slang:283-300 with nothing else to do at this vertex): plain visibility answer, no delta sample
    // of the current vertex, no emissive hit.  MIS weight from half A; the emissive slot of a non-emissive vertex is zero; the
    // generator state half A stored is already the one the next vertex continues from (this half would draw nothing).
    if (!(flags & (CI_CURRENT_MS | CI_EMISSIVE)) && ((flags & CI_ANY_HIT) || !blocked)) {
        if (!blocked) {
            if (flags & CI_DO_MIS) {
                const float w = __uint_as_float(item.info.w);
                for (int k = 0; k < 4; ++k) { float4* q = jobField(P, jobId, k); *q = *q * w; }
            }
            enqueueJob(P, replayBin(P, pixel, uint32_t(p)), jobId, jobAux(pixel, uint32_t(p)));
        } else writeContribution(P, pixel, uint32_t(p), 0, mk3(0.f));
        writeContribution(P, pixel, uint32_t(p), 1, mk3(0.f));
        // manifoldNEESample ran with nothing in the way: 0 iterations (pltpt_solve.rt.slang:318-322)
        if (P.debugView == DV_MneeIterations && !(flags & CI_ANY_HIT)) { float4 ev = mk4(0.f); debugEvent(&ev, ccInferno(0.f)); writeDebug(P, pixel, uint32_t(p), 0, ev); }
        return false;
    }
    // Blocked in manifold mode: the occluder chain can only start at a vertex that transmits through a delta lobe
    // (Manifolds.slangh:128-141: the first test after the hit, before any random number is drawn).  An opaque first occluder --
    // the usual case -- ends the connection here, from two table reads, without rebuilding the current vertex.
    if (!(flags & (CI_CURRENT_MS | CI_EMISSIVE)) && blocked) {
        const uint32_t occluderMaterial = __ldg(&P.sc.instances[occluderInstance].materialID);
        if (P.opt.MNEEMaxOccluders == 0 || !(__ldg(&P.sc.materials[occluderMaterial].propFlags) & P_CanMSTransmissive)) {
            writeContribution(P, pixel, uint32_t(p), 0, mk3(0.f));
            writeContribution(P, pixel, uint32_t(p), 1, mk3(0.f));
            if (P.debugView == DV_MneeIterations) { float4 ev = mk4(0.f); debugEvent(&ev, ccInferno(0.f)); writeDebug(P, pixel, uint32_t(p), 0, ev); }
            return false;
        }
    }
    return true;
}
template <bool kTex = true>
PLT_D void connectB(const FrameParams& P, uint32_t slot, uint32_t pixel, uint32_t& evals, TraversalCounters& tc) {
    const ConnItem item = P.connItems[slot];
    const uint32_t flags = item.info.y;
    const int p = int(P.level);
    const uint32_t jobId = item.info.x;
    const uint4 hit = P.shadowHits[slot];
    const bool blocked = hit.x != 0xffffffffu;
    ConnectCtx c;
    connectSetup<kTex>(P, pixel, p, false, c);
    const float3 lsDir = mk3(item.dirPdf.x, item.dirPdf.y, item.dirPdf.z);
    const float lsPdf = item.dirPdf.w;
    bool connected = false;
    float misScale = 1.f;
    if (flags & CI_ANY_HIT) {
        if (!blocked) {
            connected = true;
            if (flags & CI_DO_MIS) misScale = misWeight(lsPdf, bsdfPdf(c.m, c.sd, lsDir, jobField(P, jobId, 4)->x));
        }
    } else {
        ManifoldSampleData ms;
        ms.hitCount = 0; ms.iterations = 0; ms.pdf = __uint_as_float(item.info.z); ms.dispersive = (flags & CI_DISPERSIVE) != 0; ms.woLastHitToLight = mk3(0.f);
        const float heroWl = jobField(P, jobId, 4)->x;  // beam.wl.x
        ms.wavelength = heroWl;
        const bool currentVertexMS = (flags & CI_CURRENT_MS) != 0;
        if (currentVertexMS) { ms.hits[0].hit = c.bounce.hit; ms.hits[0].posW = c.sd.posW; ms.hits[0].lobe = item.info.w; ms.hitCount = 1; }
        const float3 origin = !currentVertexMS ? c.v.posW : c.st.pathOrigin;
        const float4 r0 = P.shadowRays[2 * size_t(slot)], r1 = P.shadowRays[2 * size_t(slot) + 1];
        bool ok = true;
        if (blocked) {
            TriHit th;
            th.instanceID = hit.x; th.primitiveIndex = hit.y; th.u = __uint_as_float(hit.z); th.v = __uint_as_float(hit.w);
            ok = traceMNEEOccluders<kTex>(P, c.st.tc, mk3(r0.x, r0.y, r0.z), mk3(r1.x, r1.y, r1.z), 0.f, r1.w, c.st.sg, int(P.opt.maxBounces) - p, ms, &th, P.shadowT[slot]);
        }
        // debug view: a chain that needs no walk (or cannot be walked) reports 0 iterations; parked walks report at their end
        if (P.debugView == DV_MneeIterations && !(ok && ms.hitCount > 0 && ms.pdf >= 1e-6f)) { float4 ev = mk4(0.f); debugEvent(&ev, ccInferno(0.f)); writeDebug(P, pixel, uint32_t(p), 0, ev); }
        if (ok && ms.hitCount > 0) {
            // a manifold walk is needed: park the connection for k_mnee (it owns the contribution slot from here on)
            connected = true;
            if (ms.pdf < 1e-6f) connected = false;
            else {
                const uint32_t mslot = atomicAdd(P.mneeCount, 1u);
                if (mslot >= P.mneeCapacity) { atomicAdd(P.queueCount + QC_Dropped, 1u); connected = false; }
                else {
                    MneeItem& it = P.mneeItems[mslot];
                    for (int i = 0; i < kMaxMSHits - 1; ++i) {
                        it.hit[i] = i < ms.hitCount ? ms.hits[i].hit : make_uint4(0, 0, 0, 0);
                        it.posLobe[i] = i < ms.hitCount ? mk4(ms.hits[i].posW.x, ms.hits[i].posW.y, ms.hits[i].posW.z, __uint_as_float(ms.hits[i].lobe)) : mk4(0.f);
                    }
                    const float4 bp = *jobField(P, jobId, 7);  // beam.posW
                    it.originPdf = mk4(origin.x, origin.y, origin.z, ms.pdf);
                    it.lightPosWl = mk4(bp.x, bp.y, bp.z, heroWl);
                    it.lightDirLsPdf = mk4(lsDir.x, lsDir.y, lsDir.z, lsPdf);
                    it.pixel = pixel; it.bounceIdx = uint32_t(p); it.jobId = jobId;
                    const bool distant = (__float_as_uint(jobField(P, jobId, 9)->w) & JOB_DISTANT) != 0;
                    it.flags = uint32_t(ms.hitCount) | (ms.dispersive ? MI_DISPERSIVE : 0u) | (distant ? MI_DISTANT : 0u) | (currentVertexMS ? MI_CURRENT_VERTEX_MS : 0u) | ((flags & CI_DO_MIS) ? MI_DO_MIS : 0u);
                }
            }
            if (!connected) writeContribution(P, pixel, uint32_t(p), 0, mk3(0.f));
            connectTail(P, pixel, p, c);
            evals += c.st.evals;
            tc.nodes += c.st.tc.nodes; tc.tris += c.st.tc.tris;
            return;
        }
        if (ok) {
            connected = true;
            if (flags & CI_DO_MIS) misScale = misWeight(lsPdf, bsdfPdf(c.m, c.sd, lsDir, heroWl) * ms.pdf);
        }
    }
    if (connected) {
        if (flags & CI_DO_MIS) {  // beam.scale(w) on the parked beam
            for (int k = 0; k < 4; ++k) { float4* q = jobField(P, jobId, k); *q = *q * misScale; }
        }
        enqueueJob(P, replayBin(P, pixel, uint32_t(p)), jobId, jobAux(pixel, uint32_t(p)));
    } else writeContribution(P, pixel, uint32_t(p), 0, mk3(0.f));
    connectTail(P, pixel, p, c);
    evals += c.st.evals;
    tc.nodes += c.st.tc.nodes; tc.tris += c.st.tc.tris;
}

// One replay step of one connection job (pltpt_solve.rt.slang:68-79 forwardUnwind body / :219-243 evalNEE head).
// Returns true while the job has further vertices to go through.
// One staged replay item: the job record, the vertex record and the bounce direction of the vertex it evaluates next (240 B).
struct alignas(16) UnwindItem { float4 job[kJobPlanes]; float4 vtx[4]; float4 dir; };
// One replay step of one connection job (pltpt_solve.rt.slang:68-79 forwardUnwind body / :219-243 evalNEE head) on its loaded
// inputs: `v`, `materialID`, `d` = vertex record and bounce direction of the vertex the job evaluates next, `beam`, `h` = the job.
// Returns true while the job has further vertices to go through.
template <bool kTex = true>
PLT_D bool unwindStepLoaded(const FrameParams& P, uint32_t id, const VertexData& v, uint32_t materialID, float4 d, Beam& beam, JobHeader& h, uint32_t& evals, uint32_t& nextType, uint32_t& nextAux) {
    const DevMaterial m = loadMaterialAt<kTex>(P.sc, materialID, v.texC);
    const float3 wo = oct_to_ndir_equal_area_snorm(mk2(d.x, d.y));
    const uint32_t lobe = h.firstIsCurrent ? uint32_t(L_NonDelta) : __float_as_uint(d.z);
    forwardUnwindStep<kTex>(P, evals, v, materialID, m, wo, lobe, beam);
    if (h.q == 0) {
        measureInto(P, h.pixel, h.bounceIdx, h.kind, h.pathContribution, beam);
        return false;
    }
    h.q -= 1; h.firstIsCurrent = false;
    nextType = replayBin(P, h.pixel, h.q);
    nextAux = jobAux(h.pixel, h.q);
    storeJob(P, id, beam, h);
    return true;
}
template <bool kTex = true>
PLT_D bool unwindStep(const FrameParams& P, uint32_t id, uint32_t aux, uint32_t& evals, uint32_t& nextType, uint32_t& nextAux) {
    // the queue entry names the vertex: the job record, the vertex record and the bounce direction are independent loads
    const uint32_t pixel = aux & 0x3ffffffu, q = aux >> 26;
    uint32_t materialID;
    const VertexData v = readVertexCache(P, pixel, int(q), materialID);
    const float4 d = P.bDir[size_t(q) * P.pixelCount + pixel];
    Beam beam;
    JobHeader h;
    loadJob(P, id, beam, h);
    return unwindStepLoaded<kTex>(P, id, v, materialID, d, beam, h, evals, nextType, nextAux);
}

// Sums a pixel's connections in the order the reference adds them (bounce ascending, NEE before emissive).
// debug views on: the pixel's events are averaged in the reference's order; the NaN / inf views are raised on the running sum after
// every connection measure() ran for (markDebug, pltpt_debug.slangh:34-38); final mix pltpt_solve.rt.slang:470-475
static __device__ __noinline__ float3 resolvePixelDebug(const FrameParams& P, uint32_t pixel) {
    float3 L = mk3(0.f);
    float4 dbg = mk4(0.f);
    auto slot = [&](int p, int kind) {
        const size_t o = size_t(2 * p + kind) * P.pixelCount + pixel;
        const float4 c = P.contrib[o];
        L += mk3(c.x, c.y, c.z);
        if (P.debugView == DV_NaNs) { if (c.w != 0.f && (isnan(L.x) || isnan(L.y) || isnan(L.z))) dbg += mk4(5.f, 0.f, 3.f, 1.f); }
        else if (P.debugView == DV_Inf) { if (c.w != 0.f && (isinf(L.x) || isinf(L.y) || isinf(L.z))) dbg += mk4(2.f, 0.f, 5.f, 1.f); }
        else dbg += P.debug[o];
    };
    for (int p = 0; p <= int(P.opt.maxBounces); ++p) {
        const uint32_t flags = __float_as_uint(P.bDir[size_t(p) * P.pixelCount + pixel].w);
        if (flags & RF_Miss) {
            if (P.useEnv) slot(p, 1);
            break;
        }
        slot(p, 0);
        slot(p, 1);
        if (flags & RF_Terminated) break;
    }
    const float3 dv = dbg.w > 0.f ? mk3(dbg.x, dbg.y, dbg.z) / dbg.w : mk3(0.f);
    return L + (dv - L) * P.debugIntensity;
}
PLT_D void resolvePixel(const FrameParams& P, uint32_t pixel) {
    float3 L = mk3(0.f);
    if (P.debugView) L = resolvePixelDebug(P, pixel);
    else
        for (int p = 0; p <= int(P.opt.maxBounces); ++p) {
            const uint32_t flags = __float_as_uint(P.bDir[size_t(p) * P.pixelCount + pixel].w);
            if (flags & RF_Miss) {
                if (P.useEnv) { const float4 e = __ldcs(P.contrib + size_t(2 * p + 1) * P.pixelCount + pixel); L += mk3(e.x, e.y, e.z); }
                break;
            }
            const float4 a = __ldcs(P.contrib + size_t(2 * p) * P.pixelCount + pixel), b = __ldcs(P.contrib + size_t(2 * p + 1) * P.pixelCount + pixel);
            L += mk3(a.x, a.y, a.z);
            L += mk3(b.x, b.y, b.z);
            if (flags & RF_Terminated) break;
        }
    if (P.color) P.color[P.outOffset + pixel] = mk4(L.x, L.y, L.z, 1.f);
    if (P.accum) {
        double* a = P.accum + 4 * size_t(P.outOffset + pixel);
        a[0] += double(L.x); a[1] += double(L.y); a[2] += double(L.z); a[3] += 1.0;
    }
}

// ---- kernel-side helpers shared by the stage translation units --------------------------------------------------------
constexpr int kStageThreads = 128;
// Pixels are assigned to threads in 8x4 tiles per warp so that a warp's primary rays (and its bounce-record lines) stay coherent.
PLT_D bool pixelOfThread(const FrameParams& P, uint32_t& px, uint32_t& py) {
    const uint32_t tilesX = (P.width + 7) / 8;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    px = (warp % tilesX) * 8 + (lane & 7);
    const uint32_t ly = (warp / tilesX) * 4 + (lane >> 3);
    py = P.tileY0 + ly;
    return px < P.width && ly < P.tileRows;
}
PLT_D void warpAdd(unsigned long long* dst, uint32_t v) {  // all 32 lanes must call
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, (unsigned long long)v);
}

// Counting-sort view: `pending`/`pendingType`/`hist` are filled by a producer kernel; k_sort_queue scatters the ids into
// `dst` grouped by type and k_flip_queue publishes the count in *dstCount and clears the histogram (stage_unwind.cu).
struct SortView { uint32_t* dst; uint32_t* dstCount; uint32_t* dstAux; };  // dstAux: null when the producer wrote no aux words

// stage launchers (stage_sample.cu, stage_connect.cu, stage_unwind.cu)
void launchRaygen(const FrameParams& P, cudaStream_t s);
void launchTrace(const FrameParams& P, uint32_t grid, cudaStream_t s);
void launchShade(const FrameParams& P, uint32_t grid, cudaStream_t s);
void launchConnectA(const FrameParams& P, uint32_t grid, cudaStream_t s);
void launchConnectMiss(const FrameParams& P, uint32_t grid, cudaStream_t s);
void launchShadowTrace(const FrameParams& P, uint32_t grid, cudaStream_t s);
void launchConnectB(const FrameParams& P, uint32_t grid, cudaStream_t s);  // fast pass + dense pass over the deferred entries
void launchMnee(const FrameParams& P, uint32_t grid, cudaStream_t s);
uint32_t mneeCooperativeGrid(int device);  // resident CTAs of k_mnee_walk on `device` (queried once per handle)
void launchSortFlip(const FrameParams& P, SortView v, bool resetTraceQueue, uint32_t grid, cudaStream_t s);
void launchUnwindStep(const FrameParams& P, uint32_t grid, cudaStream_t s);
void launchResolve(const FrameParams& P, cudaStream_t s);
void launchGatherBounces(const FrameParams& P, PltptBounce* out, uint32_t records, cudaStream_t s);

}  // namespace plt
