// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library; built only where /root/reference lies).
//
// ref_frames.cpp — renders frames with the REFERENCE'S OWN shader text: the two ray-generation programs of the PLTPT pass
// (pltpt_sample.rt.slang / pltpt_solve.rt.slang with everything they import, rewritten textually into C++ by oracle/ref_transpile.py ->
// oracle/_ref/ref_gen.h) driven tile by tile as PLTPT::execute does (PLTPT.cpp:487-495), over a scene in the C ABI's Falcor layouts
// (include/pltpt.h).  Hand-written here is only what the engine / the hardware does around the shaders: buffer binding, the tile loop, and
// the intersection search behind TraceRay (the oracle's ray caster: DXR hardware in the reference, no source text to follow), the texture
// unit and the environment importance map (a compute shader of the reference outside this path's scope; the oracle's).
// Used by tools/make_golden_ref.py to produce tests/golden/ref_frames.npz -- frames "of the reference itself, run here" -- which pin the
// oracle (tests/test_oracle_ref_frames.py) wherever the tests run, also where /root/reference is absent.
// The generated program state is global (shader globals are): one frame at a time, single-threaded.
// Build: make -C oracle ref   ->  oracle/_ref/libref_frames.so and libref_frames_dbg.so (-DREF_DEBUG_VIEW=1: the pass's DO_DEBUG_VIEW variant)
#include <cstdio>
#include <vector>

#include "plt_integrator.h"   // orc:: -- ray casting, texture filtering, importance map, layouts
#include "_ref/ref_gen.h"     // ref:: -- the transpiled reference
#include "ref/ref_shims_end.h"

namespace {
struct RefScene {
    orc::Scene osc;   // the scene services (ray casting, textures, env importance map)
    std::vector<ref::SpectralProfile> profiles;
    std::vector<ref::PackedStaticVertexData> verts;
    std::vector<ref::PackedEmissiveTriangle> etris;
    std::vector<ref::EmissiveFlux> eflux;
    std::vector<ref::uint2> alias;
    std::vector<ref::LightData> lights;
};
const RefScene* g = nullptr;   // the scene bound to the shader globals

void bind(const RefScene& rs) {
    g = &rs;
    const PltptSceneDesc& d = rs.osc.d;
    ref::gScene.spectralProfiles = const_cast<ref::SpectralProfile*>(rs.profiles.data());
    ref::gScene.vertices.p = const_cast<ref::PackedStaticVertexData*>(rs.verts.data());
    static_assert(sizeof(ref::MaterialDataBlob) == sizeof(PltptMaterialBlob), "both sides read the same 128-byte material records");
    ref::gScene.materials.blobs = reinterpret_cast<const ref::MaterialDataBlob*>(d.materialData);
    ref::gTextureService = &rs.osc.textures;
    ref::gScene.materials.textureFn = d.textureCount ? ref::textureServiceFn : nullptr;
    // index-buffer fetch, instance -> material, instance -> matrices (Scene.slang getIndices / getMaterialID / getVertexData:576-585)
    ref::gScene.indexFn = [](ref::GeometryInstanceID id, ref::uint tri) { uint32_t i[3]; g->osc.getIndices(id.index, tri, i); return ref::uint3(i[0], i[1], i[2]); };
    ref::gScene.materialIDFn = [](ref::GeometryInstanceID id) { return ref::uint(g->osc.d.geometryInstances[id.index].materialID); };
    ref::gScene.vertexDataFn = [](ref::TriangleHit hit) {
        const PltptSceneDesc& d = g->osc.d;
        const PltptGeometryInstance& inst = d.geometryInstances[hit.instanceID.index];
        const float* wm = d.worldMatrices + 16 * size_t(inst.globalMatrixID);
        const float* wit = d.inverseTransposeWorldMatrices + 16 * size_t(inst.globalMatrixID);
        ref::float4x4 worldMat; ref::float3x3 worldInvTransposeMat;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) worldMat[r][c] = wm[4 * r + c];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) worldInvTransposeMat[r][c] = wit[4 * r + c];
        const bool isFrontFaceCW = (inst.flags & PLTPT_INSTANCE_FLAG_OBJECT_FRONT_FACE_CW) != 0;
        const ref::uint3 vtxIndices = ref::gScene.getIndices(hit.instanceID, hit.primitiveIndex);
        ref::StaticVertexData vertices[3] = {ref::gScene.getVertex(vtxIndices[0]), ref::gScene.getVertex(vtxIndices[1]), ref::gScene.getVertex(vtxIndices[2])};
        return ref::gScene.computeVertexData(vertices, worldMat, worldInvTransposeMat, isFrontFaceCW, hit.getBarycentricWeights()); };
    ref::gScene.worldMatrixFn = [](ref::GeometryInstanceID id) {
        const float* wm = g->osc.d.worldMatrices + 16 * size_t(g->osc.d.geometryInstances[id.index].globalMatrixID);
        ref::float4x4 m; for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m[r][c] = wm[4 * r + c]; return m; };
    ref::gScene.worldFrontFaceCWFn = [](ref::GeometryInstanceID id) { return (g->osc.d.geometryInstances[id.index].flags & PLTPT_INSTANCE_FLAG_WORLD_FRONT_FACE_CW) != 0; };
    // the packed hit word (HitInfo.slang:292-312: bit allocation by scene defines)
    ref::TriangleHit::packFn = [](const ref::TriangleHit& h) {
        orc::TriangleHit oh; oh.instanceID = h.instanceID.index; oh.primitiveIndex = h.primitiveIndex; oh.barycentrics = orc::float2(h.barycentrics.x, h.barycentrics.y);
        uint32_t w[4]; g->osc.packHit(oh, w); return ref::PackedHitInfo(w[0], w[1], w[2], w[3]); };
    ref::TriangleHit::unpackFn = [](ref::PackedHitInfo p, ref::TriangleHit& h) {
        const uint32_t w[4] = {p.x, p.y, p.z, p.w}; const orc::TriangleHit oh = g->osc.unpackHit(w);
        h.instanceID.index = oh.instanceID; h.primitiveIndex = oh.primitiveIndex; h.barycentrics = ref::float2(oh.barycentrics.x, oh.barycentrics.y); };
    // TraceRay's intersection search
    ref::gCastClosest = [](ref::float3 o, ref::float3 dir, float tMin, float tMax) {
        const orc::RayHit h = g->osc.traceClosest(orc::float3(o.x, o.y, o.z), orc::float3(dir.x, dir.y, dir.z), tMin, tMax);
        ref::RayCastResult res; res.hit = h.hit.valid();
        if (res.hit) { res.instanceIndex = h.hit.instanceID; res.primitiveIndex = h.hit.primitiveIndex; res.barycentrics = ref::float2(h.hit.barycentrics.x, h.hit.barycentrics.y); res.t = h.t; }
        return res; };
    ref::gCastAny = [](ref::float3 o, ref::float3 dir, float tMin, float tMax) { return g->osc.traceAny(orc::float3(o.x, o.y, o.z), orc::float3(dir.x, dir.y, dir.z), tMin, tMax); };
    // light collection, power sampler, analytic lights
    ref::LightCollection& lc = ref::gScene.lightCollection;
    lc.triangleCount = d.emissiveTriangleCount; lc.activeTriangleCount = d.activeTriangleCount;
    lc.triangleData.p = const_cast<ref::PackedEmissiveTriangle*>(rs.etris.data()); lc.fluxData.p = const_cast<ref::EmissiveFlux*>(rs.eflux.data());
    lc.activeTriangles.p = const_cast<uint32_t*>(d.activeTriangles); lc.perMeshInstanceOffset.p = const_cast<uint32_t*>(d.perMeshInstanceOffset);
    ref::gPowerSampler._emissivePower.invWeightsSum = d.invWeightsSum;
    ref::gPowerSampler._emissivePower.triangleAliasTable.p = const_cast<ref::uint2*>(rs.alias.data());
    ref::gScene.lights = const_cast<ref::LightData*>(rs.lights.data()); ref::gScene.lightCount = d.lightCount;
    // environment map: texel loads, the bilinear filter and the importance mips are texture-unit / setup-pass services
    if (rs.osc.env.valid()) {
        ref::EnvMap& renv = ref::gScene.envMap;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) { renv.data.transform[r][c] = d.envMap.transform[4 * r + c]; renv.data.invTransform[r][c] = d.envMap.invTransform[4 * r + c]; }
        renv.data.tint = ref::float3(d.envMap.tint[0], d.envMap.tint[1], d.envMap.tint[2]); renv.data.intensity = d.envMap.intensity;
        renv.envMap.ctx = &rs.osc.env; renv.envMap.width = d.envMapWidth; renv.envMap.height = d.envMapHeight;
        renv.envMap.loadFn = [](const void* c, int x, int y, int) { const orc::float3 t = static_cast<const orc::EnvMap*>(c)->texel(x, y); return ref::float4(t.x, t.y, t.z, 1.f); };
        renv.envMap.sampleFn = [](const void* c, ref::float2 uv, float) { const orc::float3 t = static_cast<const orc::EnvMap*>(c)->sampleBilinear(orc::float2(uv.x, uv.y)); return ref::float4(t.x, t.y, t.z, 1.f); };
        ref::EnvMapSampler& smp = ref::envMapSampler;
        smp.importanceMap.ctx = &rs.osc.env; smp.importanceMap.width = smp.importanceMap.height = orc::EnvMap::kDimension;
        smp.importanceMap.loadFn = [](const void* c, int x, int y, int mip) { return static_cast<const orc::EnvMap*>(c)->load(uint32_t(x), uint32_t(y), uint32_t(mip)); };
        smp.importanceMap.sampleFn = [](const void* c, ref::float2 uv, float) {   // point sampling, clamp to edge (EnvMapSampler.cpp:70-76)
            const int dim = int(orc::EnvMap::kDimension);
            const int x = std::min(std::max(int(std::floor(uv.x * float(dim))), 0), dim - 1), y = std::min(std::max(int(std::floor(uv.y * float(dim))), 0), dim - 1);
            return static_cast<const orc::EnvMap*>(c)->load(uint32_t(x), uint32_t(y), 0u); };
        smp.importanceInvDim = ref::float2(1.f / float(orc::EnvMap::kDimension));
        smp.importanceBaseMip = rs.osc.env.baseMip;
    }
}
}  // namespace

extern "C" {

void* ref_scene_create(const PltptSceneDesc* desc) {
    RefScene* rs = new RefScene;
    rs->osc.init(*desc);
    const PltptSceneDesc& d = rs->osc.d;
    rs->profiles.resize(d.profileCount);
    for (uint32_t i = 0; i < d.profileCount; ++i) {   // Scene/SpectralProfile.slang: the same fields, copied one by one
        const PltptSpectralProfile& p = d.spectralProfiles[i];
        ref::SpectralProfile& q = rs->profiles[i];
        q.rgb = ref::float3(p.rgb[0], p.rgb[1], p.rgb[2]); q.intensity = p.intensity; q.bins = p.bins; q.minWavelength = p.minWavelength; q.maxWavelength = p.maxWavelength;
        for (int b = 0; b < 128; ++b) { q.pdf[b] = p.pdf[b]; q.icdf[b] = p.icdf[b]; }
    }
    rs->verts.resize(d.vertexCount);
    for (uint32_t i = 0; i < d.vertexCount; ++i) {    // SceneTypes.slang PackedStaticVertexData
        const PltptPackedVertex& v = d.vertices[i];
        rs->verts[i].position = ref::float3(v.position[0], v.position[1], v.position[2]);
        rs->verts[i].texCrd = ref::float2(v.texCrd[0], v.texCrd[1]);
        rs->verts[i].packedNormalTangentCurveRadius = ref::float3(ref::asfloat(ref::uint(v.packedNormalTangentCurveRadius[0])), ref::asfloat(ref::uint(v.packedNormalTangentCurveRadius[1])),
                                                                  ref::asfloat(ref::uint(v.packedNormalTangentCurveRadius[2])));
    }
    rs->etris.resize(d.emissiveTriangleCount); rs->eflux.resize(d.emissiveTriangleCount); rs->alias.resize(d.emissiveTriangleCount);
    for (uint32_t t = 0; t < d.emissiveTriangleCount; ++t) {   // LightCollectionShared.slang PackedEmissiveTriangle / EmissiveFlux
        const PltptEmissiveTriangle& e = d.emissiveTriangles[t];
        for (int k = 0; k < 3; ++k) rs->etris[t].posAndTexCoords[k] = ref::float4(e.posAndTexCoords[k][0], e.posAndTexCoords[k][1], e.posAndTexCoords[k][2], e.posAndTexCoords[k][3]);
        rs->etris[t].normal = e.normal; rs->etris[t].area = e.area; rs->etris[t].materialID = e.materialID; rs->etris[t].lightIdx = e.lightIdx;
        rs->eflux[t].flux = d.emissiveFlux[t].flux;
        if (d.triangleAliasTable) rs->alias[t] = ref::uint2(d.triangleAliasTable[2 * t], d.triangleAliasTable[2 * t + 1]);
    }
    rs->lights.resize(d.lightCount);
    for (uint32_t l = 0; l < d.lightCount; ++l) {     // LightData.slang
        const PltptLightData& L = d.lights[l];
        ref::LightData& Lr = rs->lights[l];
        Lr = ref::LightData();
        Lr.posW = ref::float3(L.posW[0], L.posW[1], L.posW[2]); Lr.type = L.type; Lr.dirW = ref::float3(L.dirW[0], L.dirW[1], L.dirW[2]);
        Lr.openingAngle = L.openingAngle; Lr.cosOpeningAngle = L.cosOpeningAngle; Lr.cosSubtendedAngle = L.cosSubtendedAngle; Lr.penumbraAngle = L.penumbraAngle;
        Lr.intensitySpectrumId = L.intensitySpectrumId; Lr.surfaceArea = L.surfaceArea;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { Lr.transMat[r][c] = L.transMat[4 * r + c]; Lr.transMatIT[r][c] = L.transMatIT[4 * r + c]; }
    }
    return rs;
}
void ref_scene_destroy(void* s) { if (g == s) g = nullptr; delete static_cast<RefScene*>(s); }

// One frame = the tile loop of PLTPT::execute (PLTPT.cpp:487-495): per tile the sample stage over the tile, then the solve stage.
// Returns 0, or -1 for an option the transpiled build fixes at compile time (HWSS = REF_HWSS -- 4 unless built with -DREF_HWSS=n, libref_frames_h<n>.so --, MNEE_MAX_MS_OCCLUDERS = 2; a debug view needs the
// DO_DEBUG_VIEW = 1 build, libref_frames_dbg.so, which in turn renders nothing else).
int ref_render_frame(void* sp, const PltptOptions* opt, const PltptCameraData* cam, uint32_t frameCount, uint32_t width, uint32_t height,
                     float* color, float* albedo, float* normal, PltptBounce* bounces) {
    const RefScene& rs = *static_cast<RefScene*>(sp);
    // the pass compiles the debug overlay in only when a view is selected (PLTPT.cpp:230): this translation unit is built once without and once with it
    if (opt->HWSS != ref::hwss_samples || opt->MNEEMaxOccluders != ref::kMNEEMaxMSOccluders || (opt->debugView != 0) != (REF_DEBUG_VIEW != 0)) return -1;
    ref::kDebugViewType = opt->debugView; ref::kDebugViewIntensity = opt->debugViewIntensity;
    orc::hwss_cfg().samples = REF_HWSS;
    bind(rs);
    const PltptSceneDesc& d = rs.osc.d;
    // the pass's defines (PLTPT.cpp:216-256) and constant buffer (PLTPT.cpp:440-470)
    ref::mMaxBounces = opt->maxBounces; ref::mTileSize = opt->tileSize; ref::mHWSSDoMIS = opt->HWSSDoMIS != 0;
    ref::kUseDirectLights = opt->useDirectLights != 0;
    ref::kUseEnvLight = opt->useEnvLights && rs.osc.env.valid();
    ref::kUseEmissiveLights = opt->useEmissiveLights && d.emissiveTriangleCount > 0;
    ref::kUseAnalyticLights = opt->useAnalyticLights && d.lightCount > 0;
    ref::kAlphaMasking = opt->alphaMasking != 0; ref::kImportanceSampleEmissiveSpectra = opt->doImportanceSampleEmitters != 0;
    ref::kDoNEE = opt->doNEE != 0; ref::kDoMSonNEE = opt->doNEE && opt->doMNEE; ref::kDoMIS = opt->doNEE && opt->doMIS;
    ref::kNEEUsePerTileSelectorSG = opt->doNEE && opt->NEEUsePerTileSG;
    ref::kDoRussianRoulette = opt->doRussianRoulette != 0;
    ref::kMNEEMaxIterations = int(opt->MNEEMaxIterations); ref::kMNEESolverThreshold = opt->MNEESolverThreshold;
    ref::kSourcingAreaFromEmissiveGeometry = opt->sourcingAreaFromEmissiveGeometry; ref::kSourcingMaxBeamOmega = opt->sourcingMaxBeamOmega;
    ref::gEmissiveKind = opt->emissiveSampler == PLTPT_EMISSIVE_SAMPLER_POWER ? 2 : 1;
    ref::kOutputSize = ref::uint2(width, height); ref::gFrameCount = frameCount;
    std::memcpy(static_cast<void*>(&ref::gScene.camera.data), cam, sizeof *cam);
    static_assert(sizeof(ref::gScene.camera.data) == sizeof(PltptCameraData), "both sides read the same 656-byte camera block");
    std::vector<ref::float4> c(size_t(width) * height), a(size_t(width) * height), n(size_t(width) * height);
    ref::gOutputColor.p = c.data(); ref::gOutputColor.width = width; ref::gOutputAlbedo.p = a.data(); ref::gOutputAlbedo.width = width; ref::gOutputNormal.p = n.data(); ref::gOutputNormal.width = width;
    const uint32_t T = opt->tileSize, perPixel = opt->maxBounces + 1;
    std::vector<ref::BounceData> tile(size_t(T) * T * perPixel);
    ref::bounceBuffer.p = tile.data();
    const uint32_t tilesX = (width + T - 1) / T, tilesY = (height + T - 1) / T;
    for (uint32_t tx = 0; tx < tilesX; ++tx)
        for (uint32_t ty = 0; ty < tilesY; ++ty) {
            ref::kTile = ref::uint2(tx, ty);
            std::memset(static_cast<void*>(tile.data()), 0, tile.size() * sizeof(ref::BounceData));
            for (uint32_t y = 0; y < T; ++y) for (uint32_t x = 0; x < T; ++x) { ref::gDispatchRaysIndex = ref::uint3(x, y, 0u); ref::sampleRayGen(); }
            for (uint32_t y = 0; y < T; ++y) for (uint32_t x = 0; x < T; ++x) { ref::gDispatchRaysIndex = ref::uint3(x, y, 0u); ref::solveRayGen(); }
            if (bounces)
                for (uint32_t y = 0; y < T && ty * T + y < height; ++y) for (uint32_t x = 0; x < T && tx * T + x < width; ++x)
                    for (uint32_t p = 0; p < perPixel; ++p) {
                        const ref::BounceData& b = tile[(size_t(y) * T + x) * perPixel + p];
                        PltptBounce& o = bounces[(size_t(ty * T + y) * width + (tx * T + x)) * perPixel + p];
                        o.hit[0] = b.hit.x; o.hit[1] = b.hit.y; o.hit[2] = b.hit.z; o.hit[3] = b.hit.w;
                        o.encodedWi[0] = b.encodedWi.x; o.encodedWi[1] = b.encodedWi.y; o.sampledBsdfLobe = b.sampledBsdfLobe; o.flags = b.flags.flags;
                        o.pathContribution = b.pathContribution; o.lastNonDeltaPdf = b.lastNonDeltaPdf;
                    }
        }
    for (size_t i = 0; i < size_t(width) * height; ++i)
        for (int k = 0; k < 4; ++k) { if (color) color[4 * i + k] = c[i][k]; if (albedo) albedo[4 * i + k] = a[i][k]; if (normal) normal[4 * i + k] = n[i][k]; }
    return 0;
}

// AccumulatePass, one frame: the reference's accumulateSingle / accumulateSingleCompensated / accumulateDouble (Accumulate.cs.slang:73-198) over
// plain arrays.  precision 0 / 1 / 2 as AccumulatePass::Precision; lastSum is float4 per pixel (0, 1) or four doubles per pixel (2: the pass keeps
// their bit halves in two uint4 textures, split and joined here); lastCorr only for 1.
int ref_accumulate_frame(uint32_t precision, uint32_t accumCount, int movingAverage, uint32_t w, uint32_t h, const float* cur, void* lastSum, float* lastCorr, float* out) {
    if (precision > 2) return -1;
    const size_t n = size_t(w) * h;
    ref::gResolution = ref::uint2(w, h); ref::gAccumCount = accumCount; ref::gAccumulate = true; ref::gMovingAverageMode = movingAverage != 0;
    struct Ctx { const float* p; uint32_t w; } ctx{cur, w};
    ref::gCurFrame.ctx = &ctx; ref::gCurFrame.width = w; ref::gCurFrame.height = h;
    ref::gCurFrame.loadFn = [](const void* c, int x, int y, int) { const Ctx* k = static_cast<const Ctx*>(c); const float* t = k->p + 4 * (size_t(y) * k->w + size_t(x)); return ref::float4(t[0], t[1], t[2], t[3]); };
    std::vector<ref::float4> o(n), sum(n), corr(n);
    std::vector<ref::uint4> lo(n), hi(n);
    ref::gOutputFrame.p = o.data(); ref::gOutputFrame.width = w;
    ref::gLastFrameSum.p = sum.data(); ref::gLastFrameSum.width = w; ref::gLastFrameCorr.p = corr.data(); ref::gLastFrameCorr.width = w;
    ref::gLastFrameSumLo.p = lo.data(); ref::gLastFrameSumLo.width = w; ref::gLastFrameSumHi.p = hi.data(); ref::gLastFrameSumHi.width = w;
    float* fs = static_cast<float*>(lastSum); double* ds = static_cast<double*>(lastSum);
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 4; ++k) {
            if (precision == 2) { uint64_t b; std::memcpy(&b, &ds[4 * i + k], 8); lo[i][k] = uint32_t(b); hi[i][k] = uint32_t(b >> 32); }
            else { sum[i][k] = fs[4 * i + k]; if (precision == 1) corr[i][k] = lastCorr[4 * i + k]; }
        }
    for (uint32_t y = 0; y < h; ++y) for (uint32_t x = 0; x < w; ++x) {
        const ref::uint3 id(x, y, 0u);
        if (precision == 0) ref::accumulateSingle(id); else if (precision == 1) ref::accumulateSingleCompensated(id); else ref::accumulateDouble(id);
    }
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 4; ++k) {
            out[4 * i + k] = o[i][k];
            if (precision == 2) { const uint64_t b = (uint64_t(hi[i][k]) << 32) | lo[i][k]; std::memcpy(&ds[4 * i + k], &b, 8); }
            else { fs[4 * i + k] = sum[i][k]; if (precision == 1) lastCorr[4 * i + k] = corr[i][k]; }
        }
    return 0;
}

// The light collection's two build passes by the reference's text (BuildTriangleList.cs.slang buildTriangleList, FinalizeIntegration.cs.slang
// finalizeIntegration), dispatched as LightCollection.cpp:481-500 / :520-540 dispatches them: one buildTriangleList dispatch per mesh light (= emissive
// instance, in instance order, gLightIdx counting them), one finalizeIntegration over all triangles.  Untextured emitters (the texel sums of textured
// ones come from a rasterisation pass outside this path).  tris / flux: emissiveTriangleCount entries of the scene's own layout.
int ref_build_light_collection(void* sp, PltptEmissiveTriangle* tris, PltptEmissiveFlux* flux) {
    const RefScene& rs = *static_cast<RefScene*>(sp);
    bind(rs);
    const PltptSceneDesc& d = rs.osc.d;
    std::vector<ref::PackedEmissiveTriangle> packed(d.emissiveTriangleCount);
    std::vector<ref::EmissiveFlux> fl(d.emissiveTriangleCount);
    ref::gTriangleData.p = packed.data(); ref::gFluxData.p = fl.data();
    uint32_t lightIdx = 0;
    for (uint32_t i = 0; i < d.instanceCount; ++i) {
        if (!d.perMeshInstanceOffset || d.perMeshInstanceOffset[i] == 0xffffffffu) continue;
        const PltptMeshDesc& m = d.meshes[d.geometryInstances[i].geometryID];
        ref::gLightIdx = lightIdx++; ref::gMaterialID = d.geometryInstances[i].materialID; ref::gInstanceID = i;
        ref::gTriangleCount = (m.indexCount > 0 ? m.indexCount : m.vertexCount) / 3; ref::gTriangleOffset = d.perMeshInstanceOffset[i];
        if (ref::gTriangleOffset + ref::gTriangleCount > d.emissiveTriangleCount) return -1;
        for (uint32_t t = 0; t < ref::gTriangleCount; ++t) ref::buildTriangleList(ref::uint3(t, 0u, 0u));
    }
    ref::gTriangleCount = d.emissiveTriangleCount;
    for (uint32_t t = 0; t < d.emissiveTriangleCount; ++t) ref::finalizeIntegration(ref::uint3(t % 256u, t / 256u, 0u));
    for (uint32_t t = 0; t < d.emissiveTriangleCount; ++t) {
        for (int k = 0; k < 3; ++k) for (int c = 0; c < 4; ++c) tris[t].posAndTexCoords[k][c] = packed[t].posAndTexCoords[k][c];
        tris[t].normal = packed[t].normal; tris[t].area = packed[t].area; tris[t].materialID = packed[t].materialID; tris[t].lightIdx = packed[t].lightIdx;
        flux[t].flux = fl[t].flux;
        for (int c = 0; c < 3; ++c) flux[t].averageRadiance[c] = fl[t].averageRadiance[c];
    }
    return 0;
}

}  // extern "C"
