// api.cu — C ABI (include/pltpt.h): lifetime, options, scene upload + table repacking, batched BSDF queries.
#include "context.h"
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

using namespace plt;

namespace plt {
static thread_local std::string g_lastError;
void setError(const std::string& s) { g_lastError = s; }
int cudaFail(cudaError_t e, const char* what) {
    setError(std::string("CUDA error '") + cudaGetErrorString(e) + "' in " + what);
    return e == cudaErrorMemoryAllocation ? PLTPT_ERR_OUT_OF_MEMORY : PLTPT_ERR_CUDA;
}
}  // namespace plt

static int fail(int code, const std::string& msg) { setError(msg); return code; }

// ------------------------------------------------------------------------------------------------ options
extern "C" void pltpt_default_options(PltptOptions* o) {  // PLTPT.h:74-108
    if (!o) return;
    *o = PltptOptions{32, 512, 4, 1, 1, 0, 1.f, 1.f, .025f, 1, 1, 1, 1, 1, 1, 1, PLTPT_EMISSIVE_SAMPLER_POWER, 1, 0, 1, 1, 2, 60, 5e-5f};
}
static int validateOptions(const PltptOptions& o) {
    if (o.HWSS < 1 || o.HWSS > 4) return fail(PLTPT_ERR_INVALID_ARGUMENT, "HWSS must be in 1..4");
    if (o.tileSize == 0) return fail(PLTPT_ERR_INVALID_ARGUMENT, "tileSize must be > 0");
    if (o.maxBounces > 63) return fail(PLTPT_ERR_INVALID_ARGUMENT, "maxBounces must be <= 63 (the wavefront queues index recorded vertices with 6 bits)");
    if (o.emissiveSampler != PLTPT_EMISSIVE_SAMPLER_UNIFORM && o.emissiveSampler != PLTPT_EMISSIVE_SAMPLER_POWER) return fail(PLTPT_ERR_INVALID_ARGUMENT, "Unknown emissive light sampler type");  // PLTPT.cpp:409
    if (o.sampleGenerator != 1) return fail(PLTPT_ERR_UNSUPPORTED, "only SAMPLE_GENERATOR_UNIFORM (1) is supported");
    if (o.MNEEMaxOccluders > 4) return fail(PLTPT_ERR_INVALID_ARGUMENT, "MNEEMaxOccluders must be <= 4");
    if (o.debugView >= 14) return fail(PLTPT_ERR_INVALID_ARGUMENT, "debugView must be a DebugViewType (0..13)");  // DebugViewType.slangh:5-20
    return PLTPT_OK;
}

extern "C" const char* pltpt_last_error(void) { return g_lastError.c_str(); }
extern "C" const char* pltpt_version(void) { return "pltpt-b200 0.1 (sm_100a)"; }

extern "C" int pltpt_create(const PltptOptions* opts, int cudaDevice, PltptHandle* out) {
    if (!out) return fail(PLTPT_ERR_INVALID_ARGUMENT, "out handle is null");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return fail(PLTPT_ERR_NO_DEVICE, "no CUDA device: the PLTPT path has no CPU fallback");
    if (cudaDevice < 0 || cudaDevice >= n) return fail(PLTPT_ERR_INVALID_ARGUMENT, "invalid CUDA device index");
    PltptOptions o;
    pltpt_default_options(&o);
    if (opts) o = *opts;
    if (int r = validateOptions(o)) return r;
    PLT_CUDA(cudaSetDevice(cudaDevice));
    PltptContext* c = new PltptContext();
    c->device = cudaDevice;
    c->opt = o;
    cudaEventCreate(&c->evA);
    cudaEventCreate(&c->evB);
    *out = c;
    return PLTPT_OK;
}
extern "C" int pltpt_destroy(PltptHandle h) {
    if (!h) return PLTPT_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto& ln : h->lanes) {
        for (auto e : ln.timerEvents) cudaEventDestroy(e);
        if (ln.done) cudaEventDestroy(ln.done);
        if (ln.lastUse) cudaEventDestroy(ln.lastUse);
        if (ln.guidesReady) cudaEventDestroy(ln.guidesReady);
        if (ln.stream) cudaStreamDestroy(ln.stream);
    }
    if (h->evFork) cudaEventDestroy(h->evFork);
    if (h->guidesReady) cudaEventDestroy(h->guidesReady);
    if (h->copyStream) cudaStreamDestroy(h->copyStream);
    for (cudaStream_t hs : h->hostStream) if (hs) cudaStreamDestroy(hs);
    if (h->hostCopyStream) cudaStreamDestroy(h->hostCopyStream);
    for (auto& lane : h->hostFrameDone) for (cudaEvent_t e : lane) if (e) cudaEventDestroy(e);
    for (auto& lane : h->hostCopyDone) for (cudaEvent_t e : lane) if (e) cudaEventDestroy(e);
    for (auto e : h->kernelEvents) cudaEventDestroy(e);
    if (h->evA) cudaEventDestroy(h->evA);
    if (h->evB) cudaEventDestroy(h->evB);
    delete h;
    return PLTPT_OK;
}
extern "C" int pltpt_set_options(PltptHandle h, const PltptOptions* opts) {
    if (!h || !opts) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    if (int r = validateOptions(*opts)) return r;
    // the scene was checked against the options in force when it was uploaded (pltpt_set_scene); PLTPT::execute re-runs setScene on an
    // option change (PLTPT.cpp:373-383), here the same checks are repeated against the resident scene instead
    if (h->hasScene && h->scene.etriCount > 0 && opts->emissiveSampler == PLTPT_EMISSIVE_SAMPLER_POWER && !h->scene.alias)
        return fail(PLTPT_ERR_INVALID_ARGUMENT, "Power emissive sampler needs the triangle alias table (the resident scene was uploaded without one)");
    if (h->waitHostAsync()) return fail(PLTPT_ERR_CUDA, "a frame queued by pltpt_render_frame_host_async failed");
    h->opt = *opts;
    return PLTPT_OK;
}
extern "C" int pltpt_get_stats(PltptHandle h, PltptStats* out) {
    if (!h || !out) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    if (int r = syncFrameStats(h)) return r;
    *out = h->stats;
    return PLTPT_OK;
}
/* Per-lane work-buffer budget in bytes (0 = default): frames whose buffers would exceed it are rendered in row tiles. */
extern "C" int pltpt_set_work_budget(PltptHandle h, uint64_t bytes) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    if (int r = syncFrameStats(h)) return r;
    h->workBudgetBytes = size_t(bytes);
    return PLTPT_OK;
}
/* Frames kept in flight by pltpt_render_accumulate: 1 = strictly one after the other, 0 = the default (four, all lanes built in). */
extern "C" int pltpt_set_frame_overlap(PltptHandle h, int frames) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    if (frames < 0 || frames > PltptContext::kLanes) return fail(PLTPT_ERR_INVALID_ARGUMENT, "frame overlap out of range (0 = default, 1 .. lanes built in)");
    h->frameOverlap = frames == 0 ? PltptContext::kDefaultOverlap : frames;
    return PLTPT_OK;
}
extern "C" int pltpt_set_bvh_update_mode(PltptHandle h, int mode) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    if (mode != PLTPT_BVH_UPDATE_REBUILD && mode != PLTPT_BVH_UPDATE_REFIT) return fail(PLTPT_ERR_INVALID_ARGUMENT, "unknown BVH update mode");
    h->bvhUpdateMode = mode;
    return PLTPT_OK;
}
extern "C" int pltpt_set_profiling(PltptHandle h, int enable) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    if (int r = syncFrameStats(h)) return r;
    h->profiling = enable != 0;
    return PLTPT_OK;
}
extern "C" int pltpt_reset_stats(PltptHandle h) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    if (int r = syncFrameStats(h)) return r;
    const double b = h->stats.bvhBuildMs, sahCost = h->stats.bvhSahCost, refitMs = h->stats.bvhRefitMs;
    const uint64_t n = h->stats.bvhNodes, l = h->stats.bvhLeafTriangles, t = h->stats.triangleCount, ov = h->stats.traversalStackOverflows;
    h->stats = PltptStats{};
    h->stats.traversalStackOverflows = ov;
    h->stats.bvhSahCost = sahCost; h->stats.bvhRefitMs = refitMs;
    h->stats.bvhBuildMs = b; h->stats.bvhNodes = n; h->stats.bvhLeafTriangles = l; h->stats.triangleCount = t;
    return PLTPT_OK;
}

// ------------------------------------------------------------------------------------------------ scene upload
namespace {
float hostEvalSpectrum(const std::vector<PltptSpectralProfile>& profiles, uint32_t id, float lambda) {  // Scene.slang:251-270
    if (id == PLTPT_NO_SPECTRAL_PROFILE || id >= profiles.size()) return 0.f;
    const PltptSpectralProfile& sp = profiles[id];
    const float x = (lambda - sp.minWavelength) / (sp.maxWavelength - sp.minWavelength) * (sp.bins - 1.0f);
    const float i = std::floor(x), w = x - i;
    if (lambda < sp.minWavelength || lambda > sp.maxWavelength) return 0.f;
    const int i0 = std::max(0, std::min(int(sp.bins) - 1, int(i))), i1 = std::max(0, std::min(int(sp.bins) - 1, int(i + 1)));
    return sp.intensity * (sp.pdf[i0] + (sp.pdf[i1] - sp.pdf[i0]) * w);
}
float hostFresnelConductor(float n, float k) {  // evalFresnelConductor(eta, 1) — Fresnel.slang:60-90
    if (n == 0 && k == 0) return 0.f;
    const std::complex<float> eta(n, k), one(1.f, 0.f);
    const std::complex<float> ie = one / eta;
    std::complex<float> t = std::sqrt(one - 0.f * ie * ie);
    const std::complex<float> rs = (one - eta * t) / (one + eta * t), rp = (eta - t) / (eta + t);
    return .5f * (std::norm(rs) + std::norm(rp));
}
// Derives one MaterialDataBlob into the device record (plt_material.cuh: the same function the textured path runs per hit).  A material
// with a Texture handle keeps its uniform-value record as a placeholder and is flagged for the per-hit path.
int unpackMaterial(const PltptMaterialBlob& m, const std::vector<PltptSpectralProfile>& profiles, uint32_t textureCount, DevMaterial& d) {
    std::memset(&d, 0, sizeof(d));
    bool textured = false;
    for (uint32_t t : {m.texBaseColor, m.texSpecular, m.texEmissive, m.texNormalMap, m.texTransmission, m.texData1, m.texData2}) {
        if (t >> 31) return fail(PLTPT_ERR_UNSUPPORTED, "UDIM texture handles are not supported (MATERIAL_SYSTEM_UDIM_INDIRECTION_ENABLED == 0 builds only)");
        const uint32_t mode = texHandleMode(t);
        if (mode > 1) return fail(PLTPT_ERR_INVALID_ARGUMENT, "unknown TextureHandle mode");
        if (mode == 1 && texHandleID(t) >= textureCount) return fail(PLTPT_ERR_INVALID_ARGUMENT, "material references texture " + std::to_string(texHandleID(t)) + " but the scene has " + std::to_string(textureCount));
    }
    for (uint32_t t : {m.texBaseColor, m.texSpecular, m.texEmissive, m.texTransmission, m.texData1, m.texData2}) textured |= texHandleMode(t) == 1;
    const uint32_t type = m.header[0] & 0xffffu;
    const bool cond = type == PLTPT_MATERIAL_CONDUCTOR || type == PLTPT_MATERIAL_GRATED_CONDUCTOR;
    const float F = cond ? hostFresnelConductor(hostEvalSpectrum(profiles, m.iorNSpectralId, kMeanWavelength), hostEvalSpectrum(profiles, m.iorKSpectralId, kMeanWavelength)) : 0.f;
    if (!deriveMaterial(m, uniformSlots(m), F, d))
        return fail(PLTPT_ERR_UNSUPPORTED, "material type " + std::to_string(type) + " is not a PLT material (only MaterialType 2..11 are reachable from PLTPT)");
    d.hdrFlags = ((m.header[0] >> 28) & 1u ? HF_DoubleSided : 0u) | (textured ? HF_Textured : 0u) |
                 (texHandleMode(m.texNormalMap) == 1 && blobNormalMapType(m) != 0 ? HF_NormalMap : 0u);
    return PLTPT_OK;
}

void repackProfile(const PltptSpectralProfile& sp, DevProfile& d) {
    std::memset(&d, 0, sizeof(d));
    d.intensity = sp.intensity; d.binsF = float(sp.bins); d.minWavelength = sp.minWavelength; d.maxWavelength = sp.maxWavelength;
    const int bins = int(std::min<uint32_t>(sp.bins, 128));
    for (int i = 0; i < 128; ++i) {
        const int i0 = std::max(0, std::min(bins - 1, i)), i1 = std::max(0, std::min(bins - 1, i + 1));
        d.pair[i] = bins > 0 ? make_float2(sp.pdf[i0], sp.pdf[i1]) : make_float2(0.f, 0.f);
        d.icdf[i] = sp.icdf[i];
    }
}

template <typename T>
int fetchHost(std::vector<T>& dst, const T* src, size_t count, bool device) {
    dst.resize(count);
    if (count == 0) return PLTPT_OK;
    if (!src) return fail(PLTPT_ERR_INVALID_ARGUMENT, "scene buffer pointer is null but its count is non-zero");
    if (device) PLT_CUDA(cudaMemcpy(dst.data(), src, count * sizeof(T), cudaMemcpyDeviceToHost));
    else std::memcpy(dst.data(), src, count * sizeof(T));
    return PLTPT_OK;
}
int upload(DevBuf& buf, const void* src, size_t bytes, bool device) {
    if (bytes == 0) { buf.release(); return PLTPT_OK; }
    if (!src) return fail(PLTPT_ERR_INVALID_ARGUMENT, "scene buffer pointer is null but its size is non-zero");
    PLT_CUDA(buf.ensure(bytes));
    PLT_CUDA(cudaMemcpy(buf.p, src, bytes, device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    return PLTPT_OK;
}
}  // namespace

// ------------------------------------------------------------------------------------------------ environment map
namespace {
// EnvMapSamplerSetup.cs.slang:47-75: one thread per importance-map texel, 8x8 bilinear luminance samples of the lat-long map
__global__ void __launch_bounds__(256) k_env_importance(DevEnvMap e, float* __restrict__ out) {
    const uint32_t px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (px >= kEnvImportanceDim || py >= kEnvImportanceDim) return;
    const float invSamples = 1.f / float(kEnvImportanceSamples * kEnvImportanceSamples);
    float L = 0.f;
    for (uint32_t y = 0; y < kEnvImportanceSamples; ++y)
        for (uint32_t x = 0; x < kEnvImportanceSamples; ++x) {
            const float2 p = mk2((float(px * kEnvImportanceSamples + x) + 0.5f) / float(kEnvImportanceDim * kEnvImportanceSamples),
                                 (float(py * kEnvImportanceSamples + y) + 0.5f) / float(kEnvImportanceDim * kEnvImportanceSamples));
            L += luminance709(envSampleBilinear(e, world_to_latlong_map(oct_to_ndir_equal_area_unorm(p))));
        }
    out[size_t(py) * kEnvImportanceDim + px] = L * invSamples;
}
// Texture::generateMips: 2x2 box filter, one launch per mip level
__global__ void __launch_bounds__(256) k_env_mip(const float* __restrict__ src, float* __restrict__ dst, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n) return;
    const uint32_t x = i % n, y = i / n;
    const size_t o = size_t(2 * y) * (2 * n) + 2 * x;
    dst[i] = ((src[o] + src[o + 1]) + (src[o + 2 * n] + src[o + 2 * n + 1])) * 0.25f;
}
__global__ void k_env_eval(DevEnvMap e, uint32_t n, const float* __restrict__ dirs, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 d = mk3(dirs[3 * size_t(i)], dirs[3 * size_t(i) + 1], dirs[3 * size_t(i) + 2]);
    const float3 c = envEval(e, d);
    out[i] = mk4(c.x, c.y, c.z, envEvalPdf(e, d));
}
__global__ void k_env_sample(DevEnvMap e, uint32_t n, const float2* __restrict__ rnd, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvSample s;
    envSample(e, rnd[i], s);
    out[2 * size_t(i)] = mk4(s.dir.x, s.dir.y, s.dir.z, s.pdf);
    out[2 * size_t(i) + 1] = mk4(s.Le.x, s.Le.y, s.Le.z, 0.f);
}

// Replaces EnvMap::setShaderData + EnvMapSampler::createImportanceMap (EnvMapSampler.cpp:82-121)
int buildEnvMap(PltptContext* h, const PltptSceneDesc* s, bool dev) {
    DevEnvMap& e = h->scene.env;
    e = DevEnvMap{};
    h->envTexels.release();
    h->envImportance.release();
    if (!s->envMapTexels) return PLTPT_OK;
    if (s->envMapWidth == 0 || s->envMapHeight == 0) return fail(PLTPT_ERR_INVALID_ARGUMENT, "environment map has zero extent");
    if (int r = upload(h->envTexels, s->envMapTexels, size_t(s->envMapWidth) * s->envMapHeight * 16, dev)) return r;
    PLT_CUDA(h->envImportance.ensure(size_t(kEnvImportanceFloats) * sizeof(float)));
    e.texels = h->envTexels.as<float4>();
    e.importance = h->envImportance.as<float>();
    e.width = s->envMapWidth; e.height = s->envMapHeight;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) { e.toWorld[3 * r + c] = s->envMap.transform[4 * r + c]; e.toLocal[3 * r + c] = s->envMap.invTransform[4 * r + c]; }
    e.intensity = mk3(s->envMap.tint[0] * s->envMap.intensity, s->envMap.tint[1] * s->envMap.intensity, s->envMap.tint[2] * s->envMap.intensity);
    float* imp = h->envImportance.as<float>();
    k_env_importance<<<dim3(kEnvImportanceDim / 16, kEnvImportanceDim / 16), 256>>>(e, imp);
    for (uint32_t mip = 1; mip <= kEnvImportanceBaseMip; ++mip) {
        const uint32_t n = kEnvImportanceDim >> mip;
        k_env_mip<<<(n * n + 255) / 256, 256>>>(imp + envMipOffset(mip - 1), imp + envMipOffset(mip), n);
    }
    h->stats.kernelLaunches += 1 + kEnvImportanceBaseMip;
    PLT_CUDA(cudaDeviceSynchronize());
    return PLTPT_OK;
}
}  // namespace

/* Environment-map queries for the parity checks (EnvMap.eval + EnvMapSampler.evalPdf; EnvMapSampler.sample). Device pointers. */
extern "C" int pltpt_envmap_eval(PltptHandle h, uint32_t n, const float* dirs, float* rgbPdf, void* stream) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null handle");
    if (!h->hasScene) return fail(PLTPT_ERR_NO_SCENE, "pltpt_set_scene has not been called");
    if (!h->scene.env.texels) return fail(PLTPT_ERR_INVALID_ARGUMENT, "the scene has no environment map");
    if (n == 0) return PLTPT_OK;
    if (!dirs || !rgbPdf) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null buffer");
    PLT_CUDA(cudaSetDevice(h->device));
    k_env_eval<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->scene.env, n, dirs, reinterpret_cast<float4*>(rgbPdf));
    h->stats.kernelLaunches++;
    PLT_CUDA(cudaGetLastError());
    return PLTPT_OK;
}
namespace {
__global__ void k_texture_sample(DevTextures tx, uint32_t handle, uint32_t samplerID, float4 uniformValue, uint32_t n, const float2* __restrict__ uv, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sampleTexture(tx, handle, samplerID, uv[i], uniformValue);
}
}  // namespace
extern "C" int pltpt_texture_sample(PltptHandle h, uint32_t handle, uint32_t samplerID, const float uniformValue[4], uint32_t n, const float* uv, float* out, void* stream) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null handle");
    if (!h->hasScene) return fail(PLTPT_ERR_NO_SCENE, "pltpt_set_scene has not been called");
    if (n == 0) return PLTPT_OK;
    if (!uv || !out || !uniformValue) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null buffer");
    if (texHandleMode(handle) == 1 && texHandleID(handle) >= h->scene.tex.textureCount) return fail(PLTPT_ERR_INVALID_ARGUMENT, "texture id out of range");
    PLT_CUDA(cudaSetDevice(h->device));
    k_texture_sample<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->scene.tex, handle, samplerID, make_float4(uniformValue[0], uniformValue[1], uniformValue[2], uniformValue[3]), n,
                                                                      reinterpret_cast<const float2*>(uv), reinterpret_cast<float4*>(out));
    h->stats.kernelLaunches++;
    PLT_CUDA(cudaGetLastError());
    return PLTPT_OK;
}
extern "C" int pltpt_envmap_sample(PltptHandle h, uint32_t n, const float* rnd, float* out, void* stream) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null handle");
    if (!h->hasScene) return fail(PLTPT_ERR_NO_SCENE, "pltpt_set_scene has not been called");
    if (!h->scene.env.texels) return fail(PLTPT_ERR_INVALID_ARGUMENT, "the scene has no environment map");
    if (n == 0) return PLTPT_OK;
    if (!rnd || !out) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null buffer");
    PLT_CUDA(cudaSetDevice(h->device));
    k_env_sample<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->scene.env, n, reinterpret_cast<const float2*>(rnd), reinterpret_cast<float4*>(out));
    h->stats.kernelLaunches++;
    PLT_CUDA(cudaGetLastError());
    return PLTPT_OK;
}

namespace {
// Material textures: every texture's mip 0 is copied into ONE device blob (256 B aligned), next to the descriptor and sampler tables and the
// raw material blobs the per-hit path reads.  The descriptor arrays of the scene description are host memory (include/pltpt.h).
int uploadTextures(PltptContext* h, const PltptSceneDesc* s, const std::vector<PltptMaterialBlob>& mats, bool dev) {
    DevScene& d = h->scene;
    d.tex = DevTextures{};
    bool any = false;
    for (const PltptMaterialBlob& m : mats)
        for (uint32_t t : {m.texBaseColor, m.texSpecular, m.texEmissive, m.texNormalMap, m.texTransmission, m.texData1, m.texData2}) any |= texHandleMode(t) == 1;
    if (!any) return PLTPT_OK;
    if (!s->textures) return fail(PLTPT_ERR_INVALID_ARGUMENT, "textured material but PltptSceneDesc::textures is null");
    static const uint32_t texelBytes[PLTPT_TEXTURE_FORMAT_COUNT] = {16, 4, 4, 8};
    std::vector<DevTexture> descs(s->textureCount);
    std::vector<size_t> offset(s->textureCount);
    size_t total = 0;
    for (uint32_t i = 0; i < s->textureCount; ++i) {
        const PltptTextureDesc& t = s->textures[i];
        if (!t.texels || t.width == 0 || t.height == 0 || t.width > 65536 || t.height > 65536) return fail(PLTPT_ERR_INVALID_ARGUMENT, "texture " + std::to_string(i) + ": null texels or size out of range");
        if (t.format >= PLTPT_TEXTURE_FORMAT_COUNT) return fail(PLTPT_ERR_UNSUPPORTED, "texture " + std::to_string(i) + ": format " + std::to_string(t.format) + " is not supported (RGBA32F, RGBA16F, RGBA8 UNORM / sRGB)");
        offset[i] = total;
        total += (size_t(t.width) * t.height * texelBytes[t.format] + 255) & ~size_t(255);
    }
    PLT_CUDA(h->texBlob.ensure(total));
    for (uint32_t i = 0; i < s->textureCount; ++i) {
        const PltptTextureDesc& t = s->textures[i];
        PLT_CUDA(cudaMemcpy(h->texBlob.as<uint8_t>() + offset[i], t.texels, size_t(t.width) * t.height * texelBytes[t.format], dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
        descs[i] = DevTexture{h->texBlob.as<uint8_t>() + offset[i], t.width, t.height, t.format, 0};
    }
    std::vector<DevSampler> smp(s->samplerCount);
    for (uint32_t i = 0; i < s->samplerCount; ++i) {
        const PltptSamplerDesc& q = s->samplers[i];
        if (q.magFilter > PLTPT_FILTER_LINEAR || q.addressU > PLTPT_ADDRESS_MIRROR_ONCE || q.addressV > PLTPT_ADDRESS_MIRROR_ONCE) return fail(PLTPT_ERR_INVALID_ARGUMENT, "sampler " + std::to_string(i) + ": filter or address mode out of range");
        smp[i] = DevSampler{q.magFilter, q.addressU, q.addressV, 0, make_float4(q.borderColor[0], q.borderColor[1], q.borderColor[2], q.borderColor[3])};
    }
    if (int r = upload(h->texDescs, descs.data(), descs.size() * sizeof(DevTexture), false)) return r;
    if (int r = upload(h->texSamplers, smp.data(), smp.size() * sizeof(DevSampler), false)) return r;
    if (int r = upload(h->materialBlobs, mats.data(), mats.size() * sizeof(PltptMaterialBlob), false)) return r;
    d.tex.textures = h->texDescs.as<DevTexture>();
    d.tex.samplers = h->texSamplers.as<DevSampler>();
    d.tex.blobs = h->materialBlobs.as<PltptMaterialBlob>();
    d.tex.textureCount = s->textureCount;
    d.tex.samplerCount = s->samplerCount;
    return PLTPT_OK;
}
}  // namespace

extern "C" int pltpt_set_scene(PltptHandle h, const PltptSceneDesc* s) {
    if (!h || !s) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    PLT_CUDA(cudaSetDevice(h->device));
    if (h->waitHostAsync()) return fail(PLTPT_ERR_CUDA, "a frame queued by pltpt_render_frame_host_async failed");
    // frames still running on the caller's stream or on a lane's stream read the scene buffers that are about to be rewritten in place
    if (int r = syncFrameStats(h)) return r;
    PLT_CUDA(cudaDeviceSynchronize());
    h->hasScene = false;
    h->ownLightCollection = false;
    const bool dev = s->pointersAreDevice != 0;
    if (s->instanceCount == 0 || s->vertexCount == 0) return fail(PLTPT_ERR_INVALID_ARGUMENT, "scene has no geometry");
    if (s->hitInfoTypeBits == 0 || s->hitInfoTypeBits > 8) return fail(PLTPT_ERR_INVALID_ARGUMENT, "hitInfoTypeBits out of range");

    // host mirrors of the small tables
    std::vector<PltptMaterialBlob> mats;
    std::vector<PltptSpectralProfile> profiles;
    if (int r = fetchHost(h->hInstances, s->geometryInstances, s->instanceCount, dev)) return r;
    if (int r = fetchHost(h->hMeshes, s->meshes, s->meshCount, dev)) return r;
    if (int r = fetchHost(mats, s->materialData, s->materialCount, dev)) return r;
    if (int r = fetchHost(profiles, s->spectralProfiles, s->profileCount, dev)) return r;

    h->hInstTriOffset.assign(s->instanceCount + 1, 0);
    for (uint32_t i = 0; i < s->instanceCount; ++i) {
        const PltptGeometryInstance& gi = h->hInstances[i];
        if ((gi.flags >> 29) != PLTPT_GEOMETRY_TYPE_TRIANGLE_MESH) return fail(PLTPT_ERR_UNSUPPORTED, "only GeometryType::TriangleMesh instances are supported (PLTPT.cpp:527 binds triangle hit groups only)");
        if (gi.geometryID >= s->meshCount) return fail(PLTPT_ERR_INVALID_ARGUMENT, "geometryID out of range");
        if (gi.materialID >= s->materialCount) return fail(PLTPT_ERR_INVALID_ARGUMENT, "materialID out of range");
        if (gi.globalMatrixID >= s->matrixCount) return fail(PLTPT_ERR_INVALID_ARGUMENT, "globalMatrixID out of range");
        const PltptMeshDesc& m = h->hMeshes[gi.geometryID];
        h->hInstTriOffset[i + 1] = h->hInstTriOffset[i] + (m.indexCount > 0 ? m.indexCount : m.vertexCount) / 3;
    }
    std::vector<DevMaterial> dmats(mats.size());
    for (size_t i = 0; i < mats.size(); ++i) {
        if (int r = unpackMaterial(mats[i], profiles, s->textureCount, dmats[i])) return r;
        for (uint32_t id : {dmats[i].emissionId, dmats[i].etaN, dmats[i].etaK})
            if (id != PLTPT_NO_SPECTRAL_PROFILE && id >= s->profileCount) return fail(PLTPT_ERR_INVALID_ARGUMENT, "spectral profile id out of range");
    }
    h->hMaterialEmission.assign(mats.size(), PLTPT_NO_SPECTRAL_PROFILE);
    h->hMaterialRgb.assign(mats.size() * 3, 0.f);
    for (size_t i = 0; i < mats.size(); ++i) {
        const uint32_t id = mats[i].emissionSpectralId;
        if (id == PLTPT_NO_SPECTRAL_PROFILE || id >= profiles.size()) continue;
        h->hMaterialEmission[i] = id;
        for (int c = 0; c < 3; ++c) h->hMaterialRgb[3 * i + c] = profiles[id].rgb[c];
    }
    for (size_t i = 0; i < profiles.size(); ++i)
        if (profiles[i].bins == 0 || profiles[i].bins > 128) return fail(PLTPT_ERR_INVALID_ARGUMENT, "spectral profile " + std::to_string(i) + ": bins must be in 1..128 (BasicMaterialData.slang:43-54)");
    {
        std::vector<PltptLightData> hl;
        if (int r = fetchHost(hl, s->lights, s->lightCount, dev)) return r;
        for (size_t i = 0; i < hl.size(); ++i)
            if (hl[i].intensitySpectrumId != PLTPT_NO_SPECTRAL_PROFILE && hl[i].intensitySpectrumId >= s->profileCount)
                return fail(PLTPT_ERR_INVALID_ARGUMENT, "light " + std::to_string(i) + ": intensitySpectrumId out of range");
    }
    std::vector<DevProfile> dprofiles(profiles.size());
    for (size_t i = 0; i < profiles.size(); ++i) repackProfile(profiles[i], dprofiles[i]);

    if (int r = upload(h->vertices, s->vertices, size_t(s->vertexCount) * sizeof(PltptPackedVertex), dev)) return r;
    if (int r = upload(h->indexData, s->indexData, s->indexDataBytes, dev)) return r;
    if (int r = upload(h->instances, h->hInstances.data(), h->hInstances.size() * sizeof(PltptGeometryInstance), false)) return r;
    if (int r = upload(h->meshes, h->hMeshes.data(), h->hMeshes.size() * sizeof(PltptMeshDesc), false)) return r;
    if (int r = upload(h->world, s->worldMatrices, size_t(s->matrixCount) * 64, dev)) return r;
    if (int r = upload(h->worldIT, s->inverseTransposeWorldMatrices, size_t(s->matrixCount) * 64, dev)) return r;
    if (int r = upload(h->materials, dmats.data(), dmats.size() * sizeof(DevMaterial), false)) return r;
    if (int r = upload(h->profiles, dprofiles.data(), dprofiles.size() * sizeof(DevProfile), false)) return r;
    if (int r = upload(h->lights, s->lights, size_t(s->lightCount) * sizeof(PltptLightData), dev)) return r;
    if (int r = upload(h->etris, s->emissiveTriangles, size_t(s->emissiveTriangleCount) * sizeof(PltptEmissiveTriangle), dev)) return r;
    if (int r = upload(h->eflux, s->emissiveFlux, size_t(s->emissiveTriangleCount) * sizeof(PltptEmissiveFlux), dev)) return r;
    if (int r = upload(h->activeTris, s->activeTriangles, size_t(s->activeTriangleCount) * 4, dev)) return r;
    if (int r = upload(h->perMeshOffset, s->perMeshInstanceOffset, s->perMeshInstanceOffset ? size_t(s->instanceCount) * 4 : 0, dev)) return r;
    if (int r = upload(h->alias, s->triangleAliasTable, s->triangleAliasTable ? size_t(s->emissiveTriangleCount) * 8 : 0, dev)) return r;
    if (int r = upload(h->instTriOffset, h->hInstTriOffset.data(), h->hInstTriOffset.size() * 4, false)) return r;

    DevScene& d = h->scene;
    d = DevScene{};
    d.vertices = h->vertices.as<PltptPackedVertex>();
    d.indexData = h->indexData.as<uint8_t>();
    d.instances = h->instances.as<PltptGeometryInstance>();
    d.meshes = h->meshes.as<PltptMeshDesc>();
    d.world = h->world.as<float>();
    d.worldIT = h->worldIT.as<float>();
    d.materials = h->materials.as<DevMaterial>();
    d.profiles = h->profiles.as<DevProfile>();
    d.lights = h->lights.as<PltptLightData>();
    d.etris = h->etris.as<PltptEmissiveTriangle>();
    d.eflux = h->eflux.as<PltptEmissiveFlux>();
    d.activeTris = h->activeTris.as<uint32_t>();
    d.perMeshInstanceOffset = h->perMeshOffset.as<uint32_t>();
    d.alias = h->alias.as<uint2>();
    d.instTriOffset = h->instTriOffset.as<uint32_t>();
    d.instanceCount = s->instanceCount;
    d.lightCount = s->lightCount;
    d.etriCount = s->emissiveTriangleCount;
    d.activeTriCount = s->activeTriangleCount;
    d.triCount = h->hInstTriOffset.back();
    d.invWeightsSum = s->invWeightsSum;
    d.hitTypeBits = s->hitInfoTypeBits; d.hitInstBits = s->hitInfoInstanceIDBits; d.hitPrimBits = s->hitInfoPrimitiveIndexBits;
    h->emissiveTriangleCount = s->emissiveTriangleCount;
    h->lightCount = s->lightCount;
    h->matrixCount = s->matrixCount;
    if (d.etriCount > 0 && (!d.etris || !d.eflux || !d.perMeshInstanceOffset)) return fail(PLTPT_ERR_INVALID_ARGUMENT, "emissive triangles present but light-collection buffers missing");
    if (d.etriCount > 0 && h->opt.emissiveSampler == PLTPT_EMISSIVE_SAMPLER_POWER && !d.alias) return fail(PLTPT_ERR_INVALID_ARGUMENT, "Power emissive sampler needs the triangle alias table");
    if (d.triCount == 0) return fail(PLTPT_ERR_INVALID_ARGUMENT, "scene has no triangles");

    if (int r = uploadTextures(h, s, mats, dev)) return r;
    if (int r = buildBvh(h, 0)) return r;
    if (int r = buildEnvMap(h, s, dev)) return r;
    h->hasScene = true;
    return PLTPT_OK;
}

// ------------------------------------------------------------------------------------------------ BSDF sweeps
namespace {
__device__ __forceinline__ SD canonicalSd(uint32_t materialID, float3 V) {
    SD sd;
    sd.posW = mk3(0.f); sd.V = V; sd.N = mk3(0.f, 0.f, 1.f); sd.T = mk3(1.f, 0.f, 0.f); sd.B = mk3(0.f, 1.f, 0.f);
    sd.faceN = sd.N; sd.uv = mk2(0.f, 0.f); sd.materialID = materialID; sd.frontFacing = true;
    return sd;
}

#ifndef PLT_MINB_SWEEP
#define PLT_MINB_SWEEP 4  // 128 registers, 16 resident warps per SM: 5.7 -> 8.2 G evals/s (profiles/r1_variants16_17_sweep_occupancy.log); 5, 6, 8 are slower again
#endif
#ifndef PLT_SWEEP_UNROLL
#define PLT_SWEEP_UNROLL 1
#endif
constexpr int kSweepThreads = 128, kSweepUnroll = PLT_SWEEP_UNROLL;
// One thread per (query, hero group): evaluates the BSDF on the four basis Stokes vectors; results are staged in
// shared memory (stride 65 words: conflict-free) and written out as fully coalesced 128 B lines.
__global__ void __launch_bounds__(kSweepThreads, PLT_MINB_SWEEP) k_bsdf_eval_mueller(DevScene sc, uint32_t materialID, uint32_t lobe, uint32_t n, uint32_t groups,
                                                                     const float* __restrict__ wi, const float* __restrict__ wo, const float* __restrict__ hero,
                                                                     float* __restrict__ out, int hwss, float maxBeamOmega) {
    extern __shared__ float stage[];  // kSweepThreads * (hwss*16 + 1)
    const int perThread = hwss * 16, stride = perThread + 1;
    const uint64_t total = uint64_t(n) * groups;
    const DevMaterial m = loadMaterial(sc.materials, materialID);
    const BsdfCtx cx = {sc.profiles, hwss};
    for (uint64_t base = uint64_t(blockIdx.x) * kSweepThreads; base < total; base += uint64_t(gridDim.x) * kSweepThreads) {
        const uint64_t item = base + threadIdx.x;
        if (item < total) {
            const uint32_t i = uint32_t(item / groups), g = uint32_t(item % groups);
            const float3 Wi = mk3(__ldg(wi + 3 * size_t(i)), __ldg(wi + 3 * size_t(i) + 1), __ldg(wi + 3 * size_t(i) + 2));
            const float3 Wo = mk3(__ldg(wo + 3 * size_t(i)), __ldg(wo + 3 * size_t(i) + 1), __ldg(wo + 3 * size_t(i) + 2));
            const float4 wl = hwssWavelengths(__ldg(hero + size_t(g) * n + i), hwss);
            const SD sd = canonicalSd(materialID, Wi);
            float* my = stage + threadIdx.x * stride;
            // the four basis Stokes vectors travel through ONE evaluation as the columns of a Beam4: frame rotation, Fresnel terms, surface
            // statistics and coherence are computed once, only the final matrix-vector products run per column (it was four whole evaluations:
            // 87 warp-instructions per matrix, profiles/r1_sweep_metrics.json)
            Beam4 beam;
            static_cast<Beam&>(beam) = sourceBeamArea(Wi, -Wi, 1.f, 1.f, mk4(1.f), wl, maxBeamOmega, hwss);
            beam.L = mk4(1.f); beam.L1 = beam.L2 = beam.L3 = mk4(0.f);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                beam.xL[c] = mk4(0.f); beam.xL1[c] = mk4(c == 0 ? 1.f : 0.f); beam.xL2[c] = mk4(c == 1 ? 1.f : 0.f); beam.xL3[c] = mk4(c == 2 ? 1.f : 0.f);
            }
            beam.propagate(sd.posW);
            bsdfEval(m, cx, sd, Wo, lobe, beam);
            PLT_FOR_LAMBDA(j, hwss) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 SP = getSPc(beam, j, c);
                    my[j * 16 + 0 + c] = SP.x; my[j * 16 + 4 + c] = SP.y; my[j * 16 + 8 + c] = SP.z; my[j * 16 + 12 + c] = SP.w;
                }
            }
        }
        __syncthreads();
        const uint64_t remaining = total - base;
        const uint32_t valid = uint32_t(remaining < kSweepThreads ? remaining : kSweepThreads) * perThread;
        float* dst = out + base * perThread;
        for (uint32_t idx = threadIdx.x; idx < valid; idx += kSweepThreads) dst[idx] = stage[(idx / perThread) * stride + (idx % perThread)];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128) k_bsdf_sample(DevScene sc, uint32_t materialID, uint32_t lobeMask, uint32_t n, const float* __restrict__ wi,
                                                     const float* __restrict__ wavelengths, uint32_t seed, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevMaterial m = loadMaterial(sc.materials, materialID);
    const float3 Wi = mk3(wi[3 * size_t(i)], wi[3 * size_t(i) + 1], wi[3 * size_t(i) + 2]);
    const SD sd = canonicalSd(materialID, Wi);
    Rng sg = Rng::seed(i & 0xffff, i >> 16, seed);
    Sample s;
    const float wl = wavelengths[i];
    const bool valid = bsdfSample(m, sd, sg, s, wl, lobeMask, mk3(0.f));
    float4* o = reinterpret_cast<float4*>(out + 8 * size_t(i));
    o[0] = mk4(s.wo.x, s.wo.y, s.wo.z, s.pdf);
    o[1] = mk4(asfloat(s.lobe), s.weight, valid ? 1.f : 0.f, valid ? bsdfPdf(m, sd, s.wo, wl) : 0.f);
}
}  // namespace

extern "C" int pltpt_bsdf_eval_mueller(PltptHandle h, uint32_t materialID, uint32_t lobe, uint32_t n, uint32_t groups, const float* wi, const float* wo,
                                       const float* hero, float* mueller, void* stream) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null handle");
    if (!h->hasScene) return fail(PLTPT_ERR_NO_SCENE, "pltpt_set_scene has not been called");
    if (materialID >= h->materials.bytes / sizeof(DevMaterial)) return fail(PLTPT_ERR_INVALID_ARGUMENT, "materialID out of range");
    if (n == 0 || groups == 0) return PLTPT_OK;
    if (!wi || !wo || !hero || !mueller) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null buffer");
    PLT_CUDA(cudaSetDevice(h->device));
    const int hwss = int(h->opt.HWSS);
    const uint64_t total = uint64_t(n) * groups;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const uint64_t blocksNeeded = (total + kSweepThreads - 1) / kSweepThreads;
    const uint32_t grid = uint32_t(std::min<uint64_t>(blocksNeeded, uint64_t(sms) * 8));
    const size_t smem = size_t(kSweepThreads) * (hwss * 16 + 1) * sizeof(float);
    k_bsdf_eval_mueller<<<grid, kSweepThreads, smem, (cudaStream_t)stream>>>(h->scene, materialID, lobe, n, groups, wi, wo, hero, mueller, hwss, h->opt.sourcingMaxBeamOmega);
    h->stats.kernelLaunches++;
    h->stats.bsdfEvals += total * hwss * 4;
    PLT_CUDA(cudaGetLastError());
    return PLTPT_OK;
}
extern "C" int pltpt_bsdf_eval_mueller_replay_math(PltptHandle h, uint32_t materialID, uint32_t lobe, uint32_t n, uint32_t groups, const float* wi, const float* wo,
                                                   const float* hero, float* mueller, void* stream) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null handle");
    if (!h->hasScene) return fail(PLTPT_ERR_NO_SCENE, "pltpt_set_scene has not been called");
    if (materialID >= h->materials.bytes / sizeof(DevMaterial)) return fail(PLTPT_ERR_INVALID_ARGUMENT, "materialID out of range");
    if (n == 0 || groups == 0) return PLTPT_OK;
    if (!wi || !wo || !hero || !mueller) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null buffer");
    PLT_CUDA(cudaSetDevice(h->device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const uint64_t total = uint64_t(n) * groups;
    const uint32_t grid = uint32_t(std::min<uint64_t>((total + 127) / 128, uint64_t(sms) * 16));
    launchBsdfEvalReplayMath(h->scene, materialID, lobe, n, groups, wi, wo, hero, mueller, int(h->opt.HWSS), h->opt.sourcingMaxBeamOmega, grid, (cudaStream_t)stream);
    h->stats.kernelLaunches++;
    PLT_CUDA(cudaGetLastError());
    return PLTPT_OK;
}
extern "C" int pltpt_bsdf_sample(PltptHandle h, uint32_t materialID, uint32_t lobeMask, uint32_t n, const float* wi, const float* wavelengths, uint32_t seed, float* out, void* stream) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null handle");
    if (!h->hasScene) return fail(PLTPT_ERR_NO_SCENE, "pltpt_set_scene has not been called");
    if (materialID >= h->materials.bytes / sizeof(DevMaterial)) return fail(PLTPT_ERR_INVALID_ARGUMENT, "materialID out of range");
    if (n == 0) return PLTPT_OK;
    if (!wi || !wavelengths || !out) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null buffer");
    PLT_CUDA(cudaSetDevice(h->device));
    k_bsdf_sample<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->scene, materialID, lobeMask, n, wi, wavelengths, seed, out);
    h->stats.kernelLaunches++;
    PLT_CUDA(cudaGetLastError());
    return PLTPT_OK;
}

// ------------------------------------------------------------------------------------------------ ray queries
extern "C" int pltpt_trace_closest(PltptHandle h, uint32_t n, const float* rays, uint32_t* hits, float* hitT, void* stream) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null handle");
    if (!h->hasScene) return fail(PLTPT_ERR_NO_SCENE, "pltpt_set_scene has not been called");
    if (n == 0) return PLTPT_OK;
    if (!rays || !hits) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null buffer");
    PLT_CUDA(cudaSetDevice(h->device));
    return traceClosest(h, n, rays, hits, hitT, (cudaStream_t)stream);
}
extern "C" int pltpt_trace_any(PltptHandle h, uint32_t n, const float* rays, uint32_t* occluded, void* stream) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null handle");
    if (!h->hasScene) return fail(PLTPT_ERR_NO_SCENE, "pltpt_set_scene has not been called");
    if (n == 0) return PLTPT_OK;
    if (!rays || !occluded) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null buffer");
    PLT_CUDA(cudaSetDevice(h->device));
    return traceAny(h, n, rays, occluded, (cudaStream_t)stream);
}

extern "C" int pltpt_set_trace_mode(PltptHandle h, int mode) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null handle");
    if (mode < PLTPT_TRACE_AUTO || mode > PLTPT_TRACE_ONE_RAY_PER_THREAD_NEAREST_FIRST) return fail(PLTPT_ERR_INVALID_ARGUMENT, "unknown trace mode");
    h->traceMode = mode;
    return PLTPT_OK;
}

// ------------------------------------------------------------------------------------------------ rendering
static int checkRenderArgs(PltptHandle h, const PltptCameraData* cam, uint32_t w, uint32_t hgt) {
    if (!h || !cam) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    if (!h->hasScene) return fail(PLTPT_ERR_NO_SCENE, "pltpt_set_scene has not been called");
    if (w == 0 || hgt == 0 || w > 65535 || hgt > 65535) return fail(PLTPT_ERR_INVALID_ARGUMENT, "frame dimensions must be in 1..65535 (path id packs the pixel in 16+16 bits, Path.slang:166)");
    if (cudaSetDevice(h->device) != cudaSuccess) return fail(PLTPT_ERR_CUDA, "cudaSetDevice failed");
    return PLTPT_OK;
}
// every entry point that renders outside the host-frame queue: the lanes' work buffers are about to be reused, so the queue drains first
static int checkRender(PltptHandle h, const PltptCameraData* cam, uint32_t w, uint32_t hgt) {
    if (int r = checkRenderArgs(h, cam, w, hgt)) return r;
    if (h->waitHostAsync()) return fail(PLTPT_ERR_CUDA, "a frame queued by pltpt_render_frame_host_async failed");
    return PLTPT_OK;
}
extern "C" int pltpt_render_frame(PltptHandle h, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t hgt, float* color, float* albedo, float* normal, void* stream) {
    if (int r = checkRender(h, cam, w, hgt)) return r;
    if (!color) return fail(PLTPT_ERR_INVALID_ARGUMENT, "color output is null");
    h->splitGuides = false;
    return renderFrameSplit(h, frameCount, cam, w, hgt, color, albedo, normal, (cudaStream_t)stream);
}
namespace {
__global__ void __launch_bounds__(256) k_accum_add(double* __restrict__ dst, const double* __restrict__ src, size_t n) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}
}  // namespace
namespace {
// AccumulatePass output of the Double mode (Accumulate.cs.slang:185-197): running sum / accumulated frame count
__global__ void __launch_bounds__(256) k_resolve_accum(const double* __restrict__ accum, float4* __restrict__ color, size_t pixels) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    const double n = accum[4 * i + 3];
    const double inv = n > 0.0 ? 1.0 / n : 0.0;
    color[i] = make_float4(float(accum[4 * i] * inv), float(accum[4 * i + 1] * inv), float(accum[4 * i + 2] * inv), 1.f);
}
// Sum of the peers' accumulators into this device's one, read over NVLink peer mappings (no staging copy, no host round trip).
constexpr int kMaxPeers = 15;
struct PeerList { const double* p[kMaxPeers]; int n; };
__global__ void __launch_bounds__(256) k_accum_add_peers(double* __restrict__ dst, PeerList peers, size_t n) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        double s = dst[i];
        for (int k = 0; k < peers.n; ++k) s += peers.p[k][i];
        dst[i] = s;
    }
}

// Frames {firstFrame + i * frameStep} restricted to image rows [rowBegin, rowEnd) (rows must start on a multiple of 4: primary rays are
// traced in 8x4 pixel tiles), summed into `accum`.
int renderAccumulateRows(PltptHandle h, uint32_t firstFrame, uint32_t frameStep, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t hgt, uint32_t rowBegin,
                         uint32_t rowEnd, double* accum, cudaStream_t user) {
    if (rowBegin >= rowEnd || rowEnd > hgt || (rowBegin & 3u)) return fail(PLTPT_ERR_INVALID_ARGUMENT, "row range must be non-empty, inside the image and start on a multiple of 4");
    // frameOverlap frames in flight (context.h "Frame lanes") unless per-kernel timers are on (they need one stream) or there are fewer frames.
    const int lanes = h->profiling ? 1 : int(std::min<uint32_t>(frameCount, uint32_t(h->frameOverlap)));
    const size_t accumBytes = size_t(w) * hgt * 4 * sizeof(double);
    if (lanes > 1) {
        if (!h->evFork) PLT_CUDA(cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming));
        PLT_CUDA(cudaEventRecord(h->evFork, user));
        for (int l = 1; l < lanes; ++l) {
            PltptContext::Lane& ln = h->lanes[l];
            if (!ln.stream) PLT_CUDA(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
            if (!ln.done) PLT_CUDA(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
            PLT_CUDA(ln.accum.ensure(accumBytes));
            PLT_CUDA(cudaStreamWaitEvent(ln.stream, h->evFork, 0));
            PLT_CUDA(cudaMemsetAsync(ln.accum.p, 0, accumBytes, ln.stream));
        }
    }
    for (uint32_t i = 0; i < frameCount; ++i) {
        const uint32_t f = firstFrame + i * frameStep;
        const int l = int(i % uint32_t(lanes));
        PltptCameraData c = *cam;
        // GBufferRT Halton-16 jitter (HaltonSamplePattern.cpp:37-67, GBufferBase.cpp:212-228)
        auto halton = [](uint32_t index, uint32_t base) { float result = 0.f, factor = 1.f; for (; index > 0; index /= base) { factor /= base; result += factor * (index % base); } return result; };
        const uint32_t k = f % 16u;
        float jx = halton(k, 2) + 0.5f, jy = halton(k, 3) + 0.5f;
        jx = (jx - std::floor(jx)) - 0.5f; jy = (jy - std::floor(jy)) - 0.5f;
        c.jitterX = jx / float(w); c.jitterY = jy / float(hgt);
        if (int r = renderFrame(h, l, f, &c, w, hgt, nullptr, nullptr, nullptr, l == 0 ? accum : h->lanes[l].accum.as<double>(), true, l == 0 ? user : h->lanes[l].stream, rowBegin, rowEnd)) return r;
    }
    for (int l = 1; l < lanes; ++l) {  // join: the caller's stream waits for the lane, then folds the lane's sums into `accum`
        PltptContext::Lane& ln = h->lanes[l];
        PLT_CUDA(cudaEventRecord(ln.done, ln.stream));
        PLT_CUDA(cudaStreamWaitEvent(user, ln.done, 0));
        const size_t n = size_t(w) * hgt * 4;
        k_accum_add<<<uint32_t((n + 255) / 256), 256, 0, user>>>(accum, ln.accum.as<double>(), n);
        h->stats.kernelLaunches++;
    }
    PLT_CUDA(cudaGetLastError());
    return PLTPT_OK;
}
}  // namespace
extern "C" int pltpt_render_accumulate(PltptHandle h, uint32_t firstFrame, uint32_t frameStep, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t hgt,
                                       double* accum, void* stream) {
    return pltpt_render_accumulate_rows(h, firstFrame, frameStep, frameCount, cam, w, hgt, 0, hgt, accum, stream);
}
extern "C" int pltpt_render_accumulate_rows(PltptHandle h, uint32_t firstFrame, uint32_t frameStep, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t hgt,
                                            uint32_t rowBegin, uint32_t rowEnd, double* accum, void* stream) {
    if (int r = checkRender(h, cam, w, hgt)) return r;
    if (!accum) return fail(PLTPT_ERR_INVALID_ARGUMENT, "accum output is null");
    if (frameStep == 0) return fail(PLTPT_ERR_INVALID_ARGUMENT, "frameStep must be > 0");
    return renderAccumulateRows(h, firstFrame, frameStep, frameCount, cam, w, hgt, rowBegin, rowEnd, accum, (cudaStream_t)stream);
}
namespace {
// AccumulatePass, one frame (Source/RenderPasses/AccumulatePass/Accumulate.cs.slang:73-198): the three precision modes with the
// reference's operations in the reference's order, so a host that steps the pass frame by frame gets the pass's own numbers.
__global__ void __launch_bounds__(256) k_accumulate_frame(uint32_t precision, uint32_t accumCount, bool movingAverage, size_t pixels, const float4* __restrict__ cur,
                                                          void* __restrict__ lastSum, float4* __restrict__ lastCorr, float4* __restrict__ out) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    const float4 c = cur[i];
    if (precision == PLTPT_ACCUMULATE_SINGLE) {  // accumulateSingle, :76-112
        float4* sumBuf = reinterpret_cast<float4*>(lastSum);
        const float w = __fdiv_rn(1.f, float(accumCount + 1u));
        const float4 last = sumBuf[i];
        float4 o;
        if (movingAverage) {  // lerp(last, cur, w) = last + (cur - last) * w
            o = mk4(__fadd_rn(last.x, __fmul_rn(__fsub_rn(c.x, last.x), w)), __fadd_rn(last.y, __fmul_rn(__fsub_rn(c.y, last.y), w)),
                    __fadd_rn(last.z, __fmul_rn(__fsub_rn(c.z, last.z), w)), __fadd_rn(last.w, __fmul_rn(__fsub_rn(c.w, last.w), w)));
            sumBuf[i] = o;
        } else {
            const float4 sum = mk4(__fadd_rn(last.x, c.x), __fadd_rn(last.y, c.y), __fadd_rn(last.z, c.z), __fadd_rn(last.w, c.w));
            o = mk4(__fmul_rn(sum.x, w), __fmul_rn(sum.y, w), __fmul_rn(sum.z, w), __fmul_rn(sum.w, w));
            sumBuf[i] = sum;
        }
        out[i] = o;
    } else if (precision == PLTPT_ACCUMULATE_SINGLE_COMPENSATED) {  // accumulateSingleCompensated, :116-145 (Kahan)
        float4* sumBuf = reinterpret_cast<float4*>(lastSum);
        const float4 sum = sumBuf[i], corr = lastCorr[i];
        const float n = float(accumCount + 1u);
        float s[4] = {sum.x, sum.y, sum.z, sum.w}, k[4] = {corr.x, corr.y, corr.z, corr.w}, v[4] = {c.x, c.y, c.z, c.w}, o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float y = __fsub_rn(v[j], k[j]);
            const float next = __fadd_rn(s[j], y);
            o[j] = __fdiv_rn(next, n);
            k[j] = __fsub_rn(__fsub_rn(next, s[j]), y);
            s[j] = next;
        }
        sumBuf[i] = mk4(s[0], s[1], s[2], s[3]);
        lastCorr[i] = mk4(k[0], k[1], k[2], k[3]);
        out[i] = mk4(o[0], o[1], o[2], o[3]);
    } else {  // accumulateDouble, :149-198
        double* sum = reinterpret_cast<double*>(lastSum) + 4 * i;
        const double w = __ddiv_rn(1.0, double(accumCount + 1u));
        const float v[4] = {c.x, c.y, c.z, c.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double a = sum[j];
            if (movingAverage) { a = __dadd_rn(__dmul_rn(a, __dsub_rn(1.0, w)), __dmul_rn(double(v[j]), w)); o[j] = float(a); }
            else { a = __dadd_rn(a, double(v[j])); o[j] = float(__dmul_rn(a, w)); }
            sum[j] = a;
        }
        out[i] = mk4(o[0], o[1], o[2], o[3]);
    }
}
}  // namespace
extern "C" int pltpt_accumulate_frame(PltptHandle h, uint32_t precision, uint32_t accumCount, int movingAverage, uint32_t w, uint32_t hgt, const float* curFrame,
                                      void* lastFrameSum, float* lastFrameCorr, float* outputFrame, void* stream) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null handle");
    if (precision > PLTPT_ACCUMULATE_DOUBLE) return fail(PLTPT_ERR_INVALID_ARGUMENT, "unknown accumulation precision");
    if (!curFrame || !lastFrameSum || !outputFrame || (precision == PLTPT_ACCUMULATE_SINGLE_COMPENSATED && !lastFrameCorr)) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null buffer");
    if (movingAverage && precision == PLTPT_ACCUMULATE_SINGLE_COMPENSATED) return fail(PLTPT_ERR_INVALID_ARGUMENT, "the compensated mode has no moving-average form (Accumulate.cs.slang:116-145)");
    const size_t pixels = size_t(w) * hgt;
    if (pixels == 0) return PLTPT_OK;
    PLT_CUDA(cudaSetDevice(h->device));
    k_accumulate_frame<<<uint32_t((pixels + 255) / 256), 256, 0, (cudaStream_t)stream>>>(precision, accumCount, movingAverage != 0, pixels, reinterpret_cast<const float4*>(curFrame), lastFrameSum,
                                                                                      reinterpret_cast<float4*>(lastFrameCorr), reinterpret_cast<float4*>(outputFrame));
    h->stats.kernelLaunches++;
    PLT_CUDA(cudaGetLastError());
    return PLTPT_OK;
}
extern "C" int pltpt_resolve_accumulated(PltptHandle h, const double* accum, uint32_t w, uint32_t hgt, float* color, void* stream) {
    if (!h || !accum || !color) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    if (w == 0 || hgt == 0) return fail(PLTPT_ERR_INVALID_ARGUMENT, "empty image");
    PLT_CUDA(cudaSetDevice(h->device));
    const size_t pixels = size_t(w) * hgt;
    k_resolve_accum<<<uint32_t((pixels + 255) / 256), 256, 0, (cudaStream_t)stream>>>(accum, reinterpret_cast<float4*>(color), pixels);
    h->stats.kernelLaunches++;
    PLT_CUDA(cudaGetLastError());
    return PLTPT_OK;
}
// One converged image for a caller with host memory only: `frameCount` frames accumulated on the device (fp64), resolved, one RGBA32F
// image copied to the host.
extern "C" int pltpt_render_accumulate_host(PltptHandle h, uint32_t firstFrame, uint32_t frameStep, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t hgt,
                                            float* colorHost) {
    if (int r = checkRender(h, cam, w, hgt)) return r;
    if (!colorHost) return fail(PLTPT_ERR_INVALID_ARGUMENT, "color output is null");
    if (frameStep == 0 || frameCount == 0) return fail(PLTPT_ERR_INVALID_ARGUMENT, "frameStep and frameCount must be > 0");
    const size_t pixels = size_t(w) * hgt;
    PLT_CUDA(h->hostAccum.ensure(pixels * 4 * sizeof(double)));
    PLT_CUDA(h->outColor.ensure(pixels * 16));
    PLT_CUDA(cudaMemsetAsync(h->hostAccum.p, 0, pixels * 4 * sizeof(double), 0));
    if (int r = renderAccumulateRows(h, firstFrame, frameStep, frameCount, cam, w, hgt, 0, hgt, h->hostAccum.as<double>(), 0)) return r;
    k_resolve_accum<<<uint32_t((pixels + 255) / 256), 256, 0, 0>>>(h->hostAccum.as<double>(), h->outColor.as<float4>(), pixels);
    h->stats.kernelLaunches++;
    PLT_CUDA(cudaMemcpyAsync(colorHost, h->outColor.p, pixels * 16, cudaMemcpyDeviceToHost, 0));
    PLT_CUDA(cudaStreamSynchronize(0));
    return PLTPT_OK;
}
// The same on several GPUs of one process (north-star (e)): handle i, on its own device with the scene replicated, renders its share --
// frame indices {firstFrame + i, + n, ...} of the whole image (partition 0), or all frames of its band of rows (partition 1) -- into an
// accumulator on its device; device 0 then sums the peers' accumulators with one kernel that reads them over NVLink peer mappings,
// resolves and copies ONE image to the host.  Summation order: handle 0, 1, ... per pixel (deterministic).
extern "C" int pltpt_render_accumulate_multi(const PltptHandle* handles, uint32_t n, uint32_t firstFrame, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t hgt,
                                             uint32_t partition, float* colorHost, double* accumOut) {
    if (!handles || n == 0 || n > uint32_t(kMaxPeers) + 1u) return fail(PLTPT_ERR_INVALID_ARGUMENT, "1 .. 16 handles are needed");
    if (!colorHost && !accumOut) return fail(PLTPT_ERR_INVALID_ARGUMENT, "no output requested");
    if (partition > PLTPT_PARTITION_ROW_BANDS) return fail(PLTPT_ERR_INVALID_ARGUMENT, "unknown partition");
    for (uint32_t i = 0; i < n; ++i) {
        if (int r = checkRender(handles[i], cam, w, hgt)) return r;
        for (uint32_t j = 0; j < i; ++j) if (handles[j]->device == handles[i]->device) return fail(PLTPT_ERR_INVALID_ARGUMENT, "handles must live on distinct devices");
    }
    const size_t pixels = size_t(w) * hgt, accumBytes = pixels * 4 * sizeof(double);
    PltptHandle h0 = handles[0];
    // peer access from device 0 to every other device (idempotent)
    PLT_CUDA(cudaSetDevice(h0->device));
    for (uint32_t i = 1; i < n; ++i) {
        int can = 0;
        PLT_CUDA(cudaDeviceCanAccessPeer(&can, h0->device, handles[i]->device));
        if (!can) return fail(PLTPT_ERR_UNSUPPORTED, "device " + std::to_string(h0->device) + " cannot map the memory of device " + std::to_string(handles[i]->device));
        const cudaError_t e = cudaDeviceEnablePeerAccess(handles[i]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cudaFail(e, "cudaDeviceEnablePeerAccess");
        cudaGetLastError();
    }
    // every device renders its share on its own (legacy default) stream, fed by its own host thread (a frame is ~100 launches: one
    // thread would fill the first device's launch queue before the second device saw any work)
    const uint32_t band = ((hgt + n - 1) / n + 3u) & ~3u;
    std::vector<int> status(n, PLTPT_OK);
    std::vector<std::string> message(n);
    auto work = [&](uint32_t i) {
        PltptHandle h = handles[i];
        auto run = [&]() -> int {
            PLT_CUDA(cudaSetDevice(h->device));
            PLT_CUDA(h->hostAccum.ensure(accumBytes));
            PLT_CUDA(cudaMemsetAsync(h->hostAccum.p, 0, accumBytes, 0));
            if (partition == PLTPT_PARTITION_SAMPLE_INDEX) {
                const uint32_t mine = frameCount / n + (i < frameCount % n ? 1u : 0u);
                if (mine) if (int r = renderAccumulateRows(h, firstFrame + i, n, mine, cam, w, hgt, 0, hgt, h->hostAccum.as<double>(), 0)) return r;
            } else {
                const uint32_t r0 = std::min(hgt, i * band), r1 = std::min(hgt, (i + 1) * band);
                if (r0 < r1) if (int r = renderAccumulateRows(h, firstFrame, 1, frameCount, cam, w, hgt, r0, r1, h->hostAccum.as<double>(), 0)) return r;
            }
            PLT_CUDA(cudaStreamSynchronize(0));   // this device's accumulator is final
            return PLTPT_OK;
        };
        status[i] = run();
        if (status[i]) message[i] = pltpt_last_error();  // (the error text is thread-local)
    };
    {
        std::vector<std::thread> threads;
        for (uint32_t i = 1; i < n; ++i) threads.emplace_back(work, i);
        work(0);
        for (auto& t : threads) t.join();
    }
    for (uint32_t i = 0; i < n; ++i) if (status[i]) return fail(status[i], "device " + std::to_string(handles[i]->device) + ": " + message[i]);
    PeerList peers{};
    for (uint32_t i = 1; i < n; ++i) peers.p[peers.n++] = handles[i]->hostAccum.as<double>();
    PLT_CUDA(cudaSetDevice(h0->device));
    if (peers.n) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h0->device);
        k_accum_add_peers<<<uint32_t(sms) * 8, 256, 0, 0>>>(h0->hostAccum.as<double>(), peers, pixels * 4);
        h0->stats.kernelLaunches++;
    }
    if (accumOut) PLT_CUDA(cudaMemcpyAsync(accumOut, h0->hostAccum.p, accumBytes, cudaMemcpyDeviceToDevice, 0));
    if (colorHost) {
        PLT_CUDA(h0->outColor.ensure(pixels * 16));
        k_resolve_accum<<<uint32_t((pixels + 255) / 256), 256, 0, 0>>>(h0->hostAccum.as<double>(), h0->outColor.as<float4>(), pixels);
        h0->stats.kernelLaunches++;
        PLT_CUDA(cudaMemcpyAsync(colorHost, h0->outColor.p, pixels * 16, cudaMemcpyDeviceToHost, 0));
    }
    PLT_CUDA(cudaStreamSynchronize(0));
    PLT_CUDA(cudaGetLastError());
    return PLTPT_OK;
}
extern "C" int pltpt_render_frame_host(PltptHandle h, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t hgt, float* colorHost, float* albedoHost, float* normalHost) {
    if (int r = checkRender(h, cam, w, hgt)) return r;
    const size_t bytes = size_t(w) * hgt * 16;
    PLT_CUDA(h->outColor.ensure(bytes));
    PLT_CUDA(h->outAlbedo.ensure(bytes));
    PLT_CUDA(h->outNormal.ensure(bytes));
    // albedo and normal leave for the host on a copy stream as soon as the sample stage has written them (under the solve stage);
    // color follows on the render stream.  Host buffers should be pinned for the copies to overlap.
    const bool guides = albedoHost || normalHost;
    if (guides) {
        if (!h->copyStream) PLT_CUDA(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
        if (!h->guidesReady) PLT_CUDA(cudaEventCreateWithFlags(&h->guidesReady, cudaEventDisableTiming));
    }
    h->armGuides = guides;  // renderTiles records guidesReady after the last tile's sample stage
    h->splitGuides = false;
    const int rr = renderFrameSplit(h, frameCount, cam, w, hgt, h->outColor.as<float>(), h->outAlbedo.as<float>(), h->outNormal.as<float>(), 0);
    h->armGuides = false;
    if (rr) return rr;
    if (guides) {
        PLT_CUDA(cudaStreamWaitEvent(h->copyStream, h->guidesReady, 0));
        if (h->splitGuides) PLT_CUDA(cudaStreamWaitEvent(h->copyStream, h->lanes[1].guidesReady, 0));
        if (albedoHost) PLT_CUDA(cudaMemcpyAsync(albedoHost, h->outAlbedo.p, bytes, cudaMemcpyDeviceToHost, h->copyStream));
        if (normalHost) PLT_CUDA(cudaMemcpyAsync(normalHost, h->outNormal.p, bytes, cudaMemcpyDeviceToHost, h->copyStream));
    }
    if (colorHost) PLT_CUDA(cudaMemcpyAsync(colorHost, h->outColor.p, bytes, cudaMemcpyDeviceToHost, 0));
    PLT_CUDA(cudaStreamSynchronize(0));
    if (guides) PLT_CUDA(cudaStreamSynchronize(h->copyStream));
    return PLTPT_OK;
}
extern "C" int pltpt_render_frame_host_async(PltptHandle h, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t hgt, float* colorHost, float* albedoHost, float* normalHost) {
    if (int r = checkRenderArgs(h, cam, w, hgt)) return r;
    if (!colorHost) return fail(PLTPT_ERR_INVALID_ARGUMENT, "color output is null");
    const int lanes = (h->profiling || h->frameOverlap < 2) ? 1 : std::min(h->frameOverlap, int(PltptContext::kLanes));
    const int l = int(h->hostAsyncNext++ % uint32_t(lanes));
    if (!h->hostStream[l]) PLT_CUDA(cudaStreamCreateWithFlags(&h->hostStream[l], cudaStreamNonBlocking));
    if (!h->hostCopyStream) PLT_CUDA(cudaStreamCreateWithFlags(&h->hostCopyStream, cudaStreamNonBlocking));
    const cudaStream_t s = h->hostStream[l], cs = h->hostCopyStream;
    const uint32_t k = h->hostSetNext[l]++ & 1u;
    if (!h->hostFrameDone[l][k]) {
        PLT_CUDA(cudaEventCreateWithFlags(&h->hostFrameDone[l][k], cudaEventDisableTiming));
        PLT_CUDA(cudaEventCreateWithFlags(&h->hostCopyDone[l][k], cudaEventDisableTiming));
    }
    const size_t bytes = size_t(w) * hgt * 16;
    plt::DevBuf* out = h->hostOut[l][k];
    for (int i = 0; i < 3; ++i) PLT_CUDA(out[i].ensure(bytes));
    h->hostAsyncPending = true;
    PLT_CUDA(cudaStreamWaitEvent(s, h->hostCopyDone[l][k], 0));  // the set's previous frame must have left for the host (no-op the first time)
    // work queued on this lane by another entry point (on the caller's stream) must finish before the lane's buffers are reused
    if (h->lanes[l].lastUse) PLT_CUDA(cudaStreamWaitEvent(s, h->lanes[l].lastUse, 0));
    if (int r = renderFrame(h, l, frameCount, cam, w, hgt, out[0].as<float>(), out[1].as<float>(), out[2].as<float>(), nullptr, true, s)) return r;
    PLT_CUDA(cudaEventRecord(h->hostFrameDone[l][k], s));
    PLT_CUDA(cudaStreamWaitEvent(cs, h->hostFrameDone[l][k], 0));
    PLT_CUDA(cudaMemcpyAsync(colorHost, out[0].p, bytes, cudaMemcpyDeviceToHost, cs));
    if (albedoHost) PLT_CUDA(cudaMemcpyAsync(albedoHost, out[1].p, bytes, cudaMemcpyDeviceToHost, cs));
    if (normalHost) PLT_CUDA(cudaMemcpyAsync(normalHost, out[2].p, bytes, cudaMemcpyDeviceToHost, cs));
    PLT_CUDA(cudaEventRecord(h->hostCopyDone[l][k], cs));
    h->hostSeqEvent[h->hostSeq++ % 8] = h->hostCopyDone[l][k];
    return PLTPT_OK;
}
extern "C" int pltpt_host_wait(PltptHandle h, uint32_t keepInFlight) {
    if (!h) return fail(PLTPT_ERR_INVALID_ARGUMENT, "null handle");
    if (cudaSetDevice(h->device) != cudaSuccess) return fail(PLTPT_ERR_CUDA, "cudaSetDevice failed");
    if (h->waitHostAsyncKeep(keepInFlight)) return fail(PLTPT_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
    return PLTPT_OK;
}
extern "C" int pltpt_sample_stage_host(PltptHandle h, uint32_t frameCount, const PltptCameraData* cam, uint32_t w, uint32_t hgt, PltptBounce* bouncesHost) {
    if (int r = checkRender(h, cam, w, hgt)) return r;
    if (!bouncesHost) return fail(PLTPT_ERR_INVALID_ARGUMENT, "bounce output is null");
    const size_t bytes = size_t(w) * hgt * 16;
    PLT_CUDA(h->outColor.ensure(bytes));
    PLT_CUDA(h->outAlbedo.ensure(bytes));
    PLT_CUDA(h->outNormal.ensure(bytes));
    return sampleStageToHost(h, frameCount, cam, w, hgt, h->outAlbedo.as<float>(), h->outNormal.as<float>(), bouncesHost);
}
