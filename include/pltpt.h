/* pltpt.h — C ABI of the B200-native PLTPT hot path.
 *
 * This is the drop-in boundary: everything a Falcor-side `PLTPT` RenderPass
 * (reference: Source/RenderPasses/PLTPT/PLTPT.{h,cpp}) needs to hand its work to
 * hand-written CUDA for sm_100a.  Plain pointers and sizes only; no C++ or torch
 * types cross it.  All entry points return a PltptStatus (0 = ok) and never throw;
 * the text of the last failure of the calling thread is in pltpt_last_error().
 *
 * Input data layouts are Falcor's own GPU-side scene buffers (the `gScene`
 * parameter block, reference Source/Falcor/Scene/Scene.slang:61-120), byte for byte.
 * The structs below restate those layouts so a maintainer can memcpy / share the
 * Falcor buffers without conversion.  Each struct cites the reference definition.
 *
 * Threading: a handle is thread-compatible, not thread-safe (the reference runs one
 * execute() per frame on the render thread, PLTPT.cpp:357).  One CUDA stream per call.
 */
#ifndef PLTPT_H_
#define PLTPT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PLTPT_API __attribute__((visibility("default")))
#else
#define PLTPT_API
#endif

/* ------------------------------------------------------------------------------------------
 * Status codes.  The host shell converts non-zero codes to Falcor::RuntimeError
 * (reference error behaviour: Source/Falcor/Core/Errors.h:80-120, PLTPT.cpp:409).
 * ---------------------------------------------------------------------------------------- */
typedef enum PltptStatus {
    PLTPT_OK = 0,
    PLTPT_ERR_INVALID_ARGUMENT = 1,
    PLTPT_ERR_CUDA = 2,
    PLTPT_ERR_UNSUPPORTED = 3, /* e.g. textured material handle, non-triangle geometry */
    PLTPT_ERR_NO_SCENE = 4,
    PLTPT_ERR_OUT_OF_MEMORY = 5,
    PLTPT_ERR_NO_DEVICE = 6
} PltptStatus;

typedef struct PltptContext* PltptHandle;

/* ------------------------------------------------------------------------------------------
 * Pass properties — the 24 keys of PLTPT::parseDictionary (PLTPT.cpp:147-175),
 * defaults from PLTPT.h:74-108.  Booleans are uint32 0/1.
 * ---------------------------------------------------------------------------------------- */
typedef struct PltptOptions {
    uint32_t maxBounces;                        /* 32  */
    uint32_t tileSize;                          /* 512 */
    uint32_t HWSS;                              /* 4   (1..4) */
    uint32_t HWSSDoMIS;                         /* 1   */
    uint32_t sampleGenerator;                   /* 1 = SAMPLE_GENERATOR_UNIFORM (only one supported) */
    uint32_t debugView;                         /* 0   */
    float    debugViewIntensity;                /* 1   */
    float    sourcingAreaFromEmissiveGeometry;  /* 1 mm^2 */
    float    sourcingMaxBeamOmega;              /* 0.025 sr */
    uint32_t useDirectLights;                   /* 1 */
    uint32_t useEnvLights;                      /* 1 */
    uint32_t useEmissiveLights;                 /* 1 */
    uint32_t useAnalyticLights;                 /* 1 */
    uint32_t alphaMasking;                      /* 1 */
    uint32_t doNEE;                             /* 1 */
    uint32_t doMIS;                             /* 1 */
    uint32_t emissiveSampler;                   /* EmissiveLightSamplerType (EmissiveLightSamplerType.slangh:37-44): 2 = Power (default), 0 = Uniform; others rejected */
    uint32_t doRussianRoulette;                 /* 1 */
    uint32_t NEEUsePerTileSG;                   /* 0 */
    uint32_t doImportanceSampleEmitters;        /* 1 */
    uint32_t doMNEE;                            /* 1 */
    uint32_t MNEEMaxOccluders;                  /* 2 */
    uint32_t MNEEMaxIterations;                 /* 60 */
    float    MNEESolverThreshold;               /* 5e-5 */
} PltptOptions;

/* EmissiveLightSamplerType values the pass offers (PLTPT.cpp:80-83; LightBVH = 1 is not offered and is rejected
 * with "Unknown emissive light sampler type" like PLTPT.cpp:409). */
#define PLTPT_EMISSIVE_SAMPLER_UNIFORM 0u
#define PLTPT_EMISSIVE_SAMPLER_POWER 2u

/* Fills *opts with the reference defaults (PLTPT.h:74-108). */
PLTPT_API void pltpt_default_options(PltptOptions* opts);

/* ------------------------------------------------------------------------------------------
 * Falcor scene buffer layouts (read-only inputs).
 * ---------------------------------------------------------------------------------------- */

/* PackedStaticVertexData, 32 B — Source/Falcor/Scene/SceneTypes.slang:246-296.
 * packed[0] = fp16 normal.x | fp16 normal.y << 16
 * packed[1] = fp16 normal.z | fp16 (tangent.w * curveRadius) << 16
 * packed[2] = tangent.xyz as 2x16 snorm octahedral (encodeNormal2x16) */
typedef struct PltptPackedVertex {
    float    position[3];
    uint32_t packedNormalTangentCurveRadius[3];
    float    texCrd[2];
} PltptPackedVertex;

/* GeometryInstanceData, 32 B — SceneTypes.slang:102-137. Upper 3 bits of flags = GeometryType. */
typedef struct PltptGeometryInstance {
    uint32_t flags;
    uint32_t globalMatrixID;
    uint32_t materialID;
    uint32_t geometryID;
    uint32_t vbOffset;
    uint32_t ibOffset; /* in units of 4 bytes into indexData (Scene.slang getLocalIndices: baseIndex = ibOffset*4) */
    uint32_t instanceIndex;
    uint32_t geometryIndex;
} PltptGeometryInstance;
#define PLTPT_GEOMETRY_TYPE_TRIANGLE_MESH 1u
#define PLTPT_INSTANCE_FLAG_USE_16BIT_INDICES 0x1u
#define PLTPT_INSTANCE_FLAG_OBJECT_FRONT_FACE_CW 0x8u
#define PLTPT_INSTANCE_FLAG_WORLD_FRONT_FACE_CW 0x10u

/* MeshDesc, 32 B — SceneTypes.slang:152-210 (indexed by GeometryInstanceData::geometryID). */
typedef struct PltptMeshDesc {
    uint32_t vbOffset, ibOffset, vertexCount, indexCount;
    uint32_t skinningVbOffset, prevVbOffset, materialID, flags;
} PltptMeshDesc;

/* MaterialDataBlob, 128 B = MaterialHeader (8 B, MaterialData.slang:38-160) + BasicMaterialData
 * (BasicMaterialData.slang:61-100).  All PLT parameters are fp16. */
typedef struct PltptMaterialBlob {
    uint32_t header[2];           /* packedData.x: type[0:16) nestedPriority[16:20) lobes[20:28) doubleSided 28 thin 29 emissive 30 basic 31 */
    uint32_t flags;
    uint16_t baseColor[4];
    uint16_t specular[4];
    uint16_t data1[4];
    uint16_t data2[4];
    uint16_t specularTransmission;
    uint16_t transmission[3];
    uint16_t diffuseTransmission;
    uint16_t volumeScattering[3];
    uint16_t IoR;
    uint16_t volumeAbsorption[3];
    uint16_t volumeAnisotropy;
    uint16_t _pad0;
    uint32_t emissionSpectralId;
    uint32_t iorNSpectralId;
    uint32_t iorKSpectralId;
    float    displacementScale;
    float    displacementOffset;
    uint32_t texBaseColor, texSpecular, texEmissive, texNormalMap;
    uint32_t texTransmission, texDisplacementMap, texData1, texData2; /* TextureHandle (Scene/Material/TextureHandle.slang:42-89): textureID = bits [0,29), mode = bits [29,31) (0 Uniform, 1 Texture), udim = bit 31 (rejected) */
    uint32_t _pad1;
} PltptMaterialBlob;
#define PLTPT_NO_SPECTRAL_PROFILE 0xffffffffu

/* MaterialType — Source/Falcor/Scene/Material/MaterialTypes.slang:37-54 */
enum {
    PLTPT_MATERIAL_STANDARD = 1,
    PLTPT_MATERIAL_DIFFUSE = 2,
    PLTPT_MATERIAL_CONDUCTOR = 3,
    PLTPT_MATERIAL_DIELECTRIC = 4,
    PLTPT_MATERIAL_OPAQUE_DIELECTRIC = 5,
    PLTPT_MATERIAL_THIN_DIELECTRIC = 6,
    PLTPT_MATERIAL_COATED_CONDUCTOR = 7,
    PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC = 8,
    PLTPT_MATERIAL_GRATED_CONDUCTOR = 9,
    PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC = 10,
    PLTPT_MATERIAL_MULTILAYERED_STACK = 11
};

/* LobeType — Source/Falcor/Rendering/Materials/LobeType.slang:36-68 */
enum {
    PLTPT_LOBE_DIFFUSE_REFLECTION = 0x01,
    PLTPT_LOBE_SCATTER_REFLECTION = 0x02,
    PLTPT_LOBE_DELTA_REFLECTION = 0x04,
    PLTPT_LOBE_DISPERSIVE_DELTA_REFLECTION = 0x08,
    PLTPT_LOBE_DIFFUSE_TRANSMISSION = 0x10,
    PLTPT_LOBE_SCATTER_TRANSMISSION = 0x20,
    PLTPT_LOBE_DELTA_TRANSMISSION = 0x40,
    PLTPT_LOBE_DISPERSIVE_DELTA_TRANSMISSION = 0x80,
    PLTPT_LOBE_DISPERSIVE_DELTA = 0x88,
    PLTPT_LOBE_DELTA = 0xCC,
    PLTPT_LOBE_NON_DELTA = 0x33,
    PLTPT_LOBE_REFLECTION = 0x0f,
    PLTPT_LOBE_TRANSMISSION = 0xf0,
    PLTPT_LOBE_ALL = 0xffff
};

/* SpectralProfile, 1052 B — BasicMaterialData.slang:43-54; built by genSpectralProfile, SceneBuilder.cpp:911-960. */
typedef struct PltptSpectralProfile {
    float    rgb[3];
    float    intensity;
    uint32_t bins;
    float    minWavelength;
    float    maxWavelength;
    float    pdf[128];
    float    icdf[128];
} PltptSpectralProfile;

/* LightData, 208 B — Source/Falcor/Scene/Lights/LightData.slang:49-72. Matrices row-major. */
typedef struct PltptLightData {
    float    posW[3];
    uint32_t type; /* 0 Point, 1 Directional, 2 Distant, 3 Rect, 4 Disc, 5 Sphere */
    float    dirW[3];
    float    openingAngle;
    float    cosOpeningAngle;
    float    cosSubtendedAngle;
    float    penumbraAngle;
    uint32_t intensitySpectrumId;
    float    tangent[3];
    float    surfaceArea;
    float    bitangent[3];
    float    _pad;
    float    transMat[16];
    float    transMatIT[16];
} PltptLightData;

/* PackedEmissiveTriangle, 64 B — Source/Falcor/Scene/Lights/LightCollectionShared.slang:76-135. */
typedef struct PltptEmissiveTriangle {
    float    posAndTexCoords[3][4]; /* xyz = world position, w = 2 x fp16 uv bits */
    uint32_t normal;                /* 2x16 snorm octahedral */
    uint32_t area;                  /* fp32 bits */
    uint32_t materialID;
    uint32_t lightIdx;
} PltptEmissiveTriangle;

/* EmissiveFlux, 16 B — LightCollectionShared.slang:140-144. */
typedef struct PltptEmissiveFlux {
    float flux;
    float averageRadiance[3];
} PltptEmissiveFlux;

/* EnvMapData, 112 B — Source/Falcor/Scene/Lights/EnvMapData.slang:36-43. float3x4 row-major (3 rows of 4). */
typedef struct PltptEnvMapData {
    float transform[12];    /* local to world */
    float invTransform[12]; /* world to local */
    float tint[3];
    float intensity;
} PltptEnvMapData;

/* CameraData, 656 B — Source/Falcor/Scene/Camera/CameraData.slang:36-69. Matrices row-major. */
typedef struct PltptCameraData {
    float viewMat[16], prevViewMat[16], projMat[16], viewProjMat[16];
    float invViewProj[16], viewProjMatNoJitter[16], prevViewProjMatNoJitter[16], projMatNoJitter[16];
    float posW[3];     float focalLength;
    float prevPosW[3]; float _padding0;
    float up[3];       float aspectRatio;
    float target[3];   float nearZ;
    float cameraU[3];  float farZ;
    float cameraV[3];  float jitterX;
    float cameraW[3];  float jitterY;
    float frameHeight, frameWidth, focalDistance, apertureRadius;
    float shutterSpeed, ISOSpeed, polarizerTheta, polarizer;
} PltptCameraData;

/* BounceData, 40 B — Source/RenderPasses/PLTPT/Path.slang:113-142 (the record appendBounce writes,
 * pltpt.slangh:138-154).  Exposed for parity checks of the sample stage. */
typedef struct PltptBounce {
    uint32_t hit[4];        /* PackedHitInfo, HitInfo.slang:459-487 */
    float    encodedWi[2];  /* equal-area octahedral, MathHelpers.slang:199-240 */
    uint32_t sampledBsdfLobe;
    uint32_t flags;         /* RayFlags, Path.slang:47-109 */
    float    pathContribution;
    float    lastNonDeltaPdf;
} PltptBounce;

/* One material texture: mip level 0 only (the pass samples at explicit LOD 0), linear layout, row 0 = v 0, row pitch = width * texel size.
 * Formats: the uncompressed ones a Falcor TextureManager hands out for PLT material parameters. */
enum {
    PLTPT_TEXTURE_FORMAT_RGBA32_FLOAT = 0,
    PLTPT_TEXTURE_FORMAT_RGBA8_UNORM = 1,
    PLTPT_TEXTURE_FORMAT_RGBA8_UNORM_SRGB = 2, /* rgb decoded to linear before filtering, alpha linear (D3D12 sRGB view) */
    PLTPT_TEXTURE_FORMAT_RGBA16_FLOAT = 3,
    PLTPT_TEXTURE_FORMAT_COUNT = 4,
};
typedef struct PltptTextureDesc {
    const void* texels;
    uint32_t width, height;
    uint32_t format;
    uint32_t _pad;
} PltptTextureDesc;
/* Sampler state (Core/API/Sampler.h:50-66): magnification filter 0 = Point, 1 = Linear; address modes 0 Wrap, 1 Mirror, 2 Clamp,
 * 3 Border, 4 MirrorOnce.  Bilinear weights are exact fp32 (the texture unit's 8-bit fixed-point weights are not reproduced). */
enum { PLTPT_FILTER_POINT = 0, PLTPT_FILTER_LINEAR = 1 };
enum { PLTPT_ADDRESS_WRAP = 0, PLTPT_ADDRESS_MIRROR = 1, PLTPT_ADDRESS_CLAMP = 2, PLTPT_ADDRESS_BORDER = 3, PLTPT_ADDRESS_MIRROR_ONCE = 4 };
typedef struct PltptSamplerDesc {
    uint32_t magFilter, addressU, addressV, _pad;
    float borderColor[4];
} PltptSamplerDesc;
/* BasicMaterialData::flags (BasicMaterialData.slang:102-126): NormalMapType = bits [1,3) (0 None, 1 RGB, 2 RG);
 * MaterialHeader packedData.y (MaterialData.slang:48-67,118-122): default texture sampler ID = bits [17,25). */

/* Scene description: the subset of `gScene` + light-sampler data the pass reads.
 * If pointersAreDevice != 0 all pointers are CUDA device pointers (the Falcor interop case:
 * Resource::getSharedApiHandle -> cudaImportExternalMemory); otherwise they are host pointers.
 * pltpt_set_scene copies everything it needs; the caller keeps ownership. */
typedef struct PltptSceneDesc {
    const PltptPackedVertex*     vertices;            uint32_t vertexCount;
    const void*                  indexData;           uint64_t indexDataBytes;
    const PltptGeometryInstance* geometryInstances;   uint32_t instanceCount;
    const PltptMeshDesc*         meshes;              uint32_t meshCount;
    const float*                 worldMatrices;       /* float4x4 row-major per matrixID */
    const float*                 inverseTransposeWorldMatrices;
    uint32_t                     matrixCount;
    const PltptMaterialBlob*     materialData;        uint32_t materialCount;
    const PltptSpectralProfile*  spectralProfiles;    uint32_t profileCount;
    const PltptLightData*        lights;              uint32_t lightCount;
    /* LightCollection (Source/Falcor/Scene/Lights/LightCollection.slang:40-60) */
    const PltptEmissiveTriangle* emissiveTriangles;   uint32_t emissiveTriangleCount;
    const PltptEmissiveFlux*     emissiveFlux;
    const uint32_t*              activeTriangles;     uint32_t activeTriangleCount;
    const uint32_t*              perMeshInstanceOffset; /* [instanceCount], 0xffffffff = not emissive */
    /* EmissivePowerSampler (Source/Falcor/Rendering/Lights/EmissivePowerSampler.slang:36-40) */
    const uint32_t*              triangleAliasTable;  /* uint2 per emissive triangle */
    float                        invWeightsSum;
    /* HitInfo bit allocation (scene defines, Source/Falcor/Scene/HitInfo.cpp:52-90) */
    uint32_t hitInfoTypeBits, hitInfoInstanceIDBits, hitInfoPrimitiveIndexBits;
    uint32_t pointersAreDevice;
    /* Environment map (Source/Falcor/Scene/Lights/EnvMap.slang:36-52): lat-long RGBA32F texels, row-major, sampled with
     * bilinear filtering, wrap in u / clamp in v (EnvMap.cpp:121-126).  NULL = the scene has no env map
     * (Scene::useEnvLight() false).  pltpt_set_scene builds the hierarchical importance map itself
     * (replaces EnvMapSampler::createImportanceMap, EnvMapSampler.cpp:82-121: 512x512 texels, 64 samples per texel). */
    const float*                 envMapTexels;        uint32_t envMapWidth, envMapHeight;
    PltptEnvMapData              envMap;
    /* Material textures and samplers (MaterialSystem.slang:283-291: `materialTextures[handle.getTextureID()]` sampled with
     * `materialSamplers[header.getDefaultTextureSamplerID()]` at explicit LOD 0, pltpt_sample.rt.slang:76, pltpt_solve.rt.slang:53).
     * The two descriptor arrays are always HOST pointers; every `texels` pointer is host or device like the other scene buffers.
     * textureCount == 0: every TextureHandle must be Uniform.  samplerCount == 0: the material system's default sampler
     * (linear filter, wrap addressing; MaterialSystem.cpp:76-79) is used for every material. */
    const PltptTextureDesc*      textures;            uint32_t textureCount;
    const PltptSamplerDesc*      samplers;            uint32_t samplerCount;
} PltptSceneDesc;

/* ------------------------------------------------------------------------------------------
 * Lifetime
 * ---------------------------------------------------------------------------------------- */
PLTPT_API int pltpt_create(const PltptOptions* opts, int cudaDevice, PltptHandle* out);
PLTPT_API int pltpt_destroy(PltptHandle h);
PLTPT_API const char* pltpt_last_error(void);
PLTPT_API const char* pltpt_version(void);

/* Replaces PLTPT::parseDictionary + the recompile on `mOptionsChanged` (PLTPT.cpp:147-175,361-372). */
PLTPT_API int pltpt_set_options(PltptHandle h, const PltptOptions* opts);

/* Replaces PLTPT::setScene (PLTPT.cpp:501-552) and the driver BLAS/TLAS build
 * (Source/Falcor/Scene/Scene.cpp:3092-3760): uploads the scene, builds the 8-wide BVH. */
PLTPT_API int pltpt_set_scene(PltptHandle h, const PltptSceneDesc* scene);

/* ------------------------------------------------------------------------------------------
 * Rendering — replaces the tile loop of PLTPT::execute (PLTPT.cpp:487-495):
 * one call = one sample per pixel = sample stage + solve stage over the whole frame.
 * color/albedo/normal are RGBA32F device buffers of width*height texels (row-major, like the
 * RenderData textures "color","albedo","normal", PLTPT.cpp:65-78); albedo/normal may be NULL.
 * `camera` is a HOST pointer.  `stream` is a cudaStream_t (0 = default).
 * ---------------------------------------------------------------------------------------- */
PLTPT_API int pltpt_render_frame(PltptHandle h, uint32_t frameCount, const PltptCameraData* camera,
                                 uint32_t width, uint32_t height,
                                 float* color, float* albedo, float* normal, void* stream);

/* AccumulatePass equivalent (Source/RenderPasses/AccumulatePass/Accumulate.cs.slang:148-198, Double mode):
 * renders frames [firstFrame, firstFrame+frameStep*frameCount) with stride frameStep, applying the
 * GBufferRT Halton-16 jitter per frame (GBufferBase.cpp:212-228), summing color into `accum`
 * (width*height*4 doubles, device).  Used for sample-index partitioning across GPUs. */
PLTPT_API int pltpt_render_accumulate(PltptHandle h, uint32_t firstFrame, uint32_t frameStep, uint32_t frameCount,
                                      const PltptCameraData* camera, uint32_t width, uint32_t height,
                                      double* accum, void* stream);

/* The same restricted to image rows [rowBegin, rowEnd) of every frame (rowBegin a multiple of 4): the image-tile half of the multi-GPU
 * partition (north-star (e)); pixels outside the rows are not touched, so accumulators of disjoint bands add up to the whole image. */
PLTPT_API int pltpt_render_accumulate_rows(PltptHandle h, uint32_t firstFrame, uint32_t frameStep, uint32_t frameCount,
                                           const PltptCameraData* camera, uint32_t width, uint32_t height,
                                           uint32_t rowBegin, uint32_t rowEnd, double* accum, void* stream);
/* AccumulatePass stepped frame by frame, all three precision modes (Accumulate.cs.slang:73-198; the PLTPathTracer.py graph runs the
 * colour through Double and the albedo / normal guides through Single): folds `curFrame` (RGBA32F) into the running state and writes
 * the pass's output.  `accumCount` = frames accumulated before this one (gAccumCount); state buffers start zeroed: `lastFrameSum` is
 * RGBA32F for the single-precision modes and RGBA64F (4 doubles per pixel) for Double, `lastFrameCorr` (RGBA32F) only for
 * SingleCompensated.  movingAverage != 0 selects the exponential form the pass switches to at its frame limit (OverflowMode::EMA).
 * Device pointers.  (pltpt_render_accumulate is the fused form of Double for a whole accumulation.) */
enum { PLTPT_ACCUMULATE_SINGLE = 0, PLTPT_ACCUMULATE_SINGLE_COMPENSATED = 1, PLTPT_ACCUMULATE_DOUBLE = 2 };
PLTPT_API int pltpt_accumulate_frame(PltptHandle h, uint32_t precision, uint32_t accumCount, int movingAverage, uint32_t width, uint32_t height,
                                     const float* curFrame, void* lastFrameSum, float* lastFrameCorr, float* outputFrame, void* stream);
/* AccumulatePass output (Accumulate.cs.slang:185-197): color = sum / accumulated frame count, alpha 1.  Device pointers. */
PLTPT_API int pltpt_resolve_accumulated(PltptHandle h, const double* accum, uint32_t width, uint32_t height, float* color, void* stream);
/* One converged image for a caller with host memory only: `frameCount` frames accumulated on the device in fp64, resolved, ONE
 * RGBA32F image copied to `colorHost` (pinned memory recommended).  Returns when the image is on the host. */
PLTPT_API int pltpt_render_accumulate_host(PltptHandle h, uint32_t firstFrame, uint32_t frameStep, uint32_t frameCount,
                                           const PltptCameraData* camera, uint32_t width, uint32_t height, float* colorHost);
/* Several GPUs of one process (north-star (e); one process per GPU uses pltpt_render_accumulate + an NCCL reduce instead):
 * handles[i] lives on its own device with the same scene.  SAMPLE_INDEX: handle i renders frames firstFrame + i, + n, ... of the whole
 * image; ROW_BANDS: every handle renders all frames of its band of rows.  Device 0 sums the peers' fp64 accumulators with one kernel
 * that reads them through NVLink peer mappings, resolves, and copies one image to `colorHost` (may be NULL); `accumOut` (may be NULL)
 * receives the summed RGBA64F accumulator on handles[0]'s device.  The result equals the single-GPU accumulation of the same frames
 * up to fp64 summation order. */
enum { PLTPT_PARTITION_SAMPLE_INDEX = 0, PLTPT_PARTITION_ROW_BANDS = 1 };
PLTPT_API int pltpt_render_accumulate_multi(const PltptHandle* handles, uint32_t handleCount, uint32_t firstFrame, uint32_t frameCount,
                                            const PltptCameraData* camera, uint32_t width, uint32_t height, uint32_t partition,
                                            float* colorHost, double* accumOut);

/* Frames pltpt_render_accumulate keeps in flight (ours; frames are independent, pltpt_sample.rt.slang:154): 0 = the default
 * (every lane the library was built with, each a CUDA stream with private work buffers), 1 = strictly sequential on the
 * caller's stream, n = n frames.
 * The sum is the same set of frames either way; only the fp64 summation order differs. */
PLTPT_API int pltpt_set_frame_overlap(PltptHandle h, int frames);

/* Work-buffer budget per lane in bytes (ours; 0 = the default 24 GB).  The reference bounds its bounce buffer with `tileSize`
 * (512 x 512 pixel tiles, PLTPT.cpp:426-428); this library tiles by rows so that one tile's work buffers fit the budget
 * (1080p and 4K at depth 8 need 9 GB and 37 GB untiled).  Results do not depend on the tiling. */
PLTPT_API int pltpt_set_work_budget(PltptHandle h, uint64_t bytes);

/* Host-buffer convenience used by the Falcor shell when no shared CUDA surface exists:
 * renders one frame and copies RGBA32F results to host memory (any of the three may be NULL). */
PLTPT_API int pltpt_render_frame_host(PltptHandle h, uint32_t frameCount, const PltptCameraData* camera,
                                      uint32_t width, uint32_t height,
                                      float* colorHost, float* albedoHost, float* normalHost);

/* The same call, queued: returns once the frame and its device-to-host copies are enqueued, the way PLTPT::execute records
 * its dispatches and returns (PLTPT.cpp:357-499; the frame is complete only after the render context's flush).  Consecutive
 * calls go to alternating frame lanes (pltpt_set_frame_overlap), each with its own stream and output images, so the copies of
 * one frame run under the kernels of the next.  The host buffers of a queued frame (pinned, or the copies serialise) must stay
 * untouched until pltpt_host_wait says the frame has arrived.  colorHost is required.
 * Every other entry point that renders, changes the scene or destroys the handle waits for the queue first.  Pixel values are
 * those of pltpt_render_frame_host bit for bit. */
PLTPT_API int pltpt_render_frame_host_async(PltptHandle h, uint32_t frameCount, const PltptCameraData* camera,
                                            uint32_t width, uint32_t height,
                                            float* colorHost, float* albedoHost, float* normalHost);
/* Blocks until at most `keepInFlight` of the frames queued by pltpt_render_frame_host_async are still on their way (frames
 * complete in the order they were queued); 0 = every queued frame is in its host buffers.  A consumer that works on frame i-2
 * while frames i-1 and i render calls pltpt_host_wait(h, 2) after queueing frame i and cycles three host image sets.  Reports a
 * failed frame. */
PLTPT_API int pltpt_host_wait(PltptHandle h, uint32_t keepInFlight);

/* Sample-stage only; copies the bounce buffer ((maxBounces+1) records per pixel, pixel-major,
 * pltpt.slangh:132-136 with the tile spanning the whole frame) to host for parity checks. */
PLTPT_API int pltpt_sample_stage_host(PltptHandle h, uint32_t frameCount, const PltptCameraData* camera,
                                      uint32_t width, uint32_t height, PltptBounce* bouncesHost);

/* ------------------------------------------------------------------------------------------
 * Scene ingestion without Falcor (ours): what SceneBuilder / LightCollection / EmissivePowerSampler do on the
 * host and in compute passes, so that a shell can go from `.spd` tables, PLT material parameters and emissive meshes to the
 * buffers pltpt_set_scene takes.  Host pointers.
 * ---------------------------------------------------------------------------------------- */
/* PiecewiseLinearSpectrum::fromFile (Source/Falcor/Utils/Color/Spectrum.cpp:89-112): `lambda value` rows, '#' comments, stops at
 * the first malformed row.  *count receives the number of rows (call with capacity 0 to size the buffers). */
PLTPT_API int pltpt_spd_load(const char* path, float* wavelengths, float* values, uint32_t capacity, uint32_t* count);
/* SampledSpectrum(PiecewiseLinearSpectrum) + genSpectralProfile (SampledSpectrum.h:57-68, SceneBuilder.cpp:911-960);
 * computeICdf = 1 for emitters (normalised pdf + inverse cdf), 0 for IOR tables (raw values). */
PLTPT_API int pltpt_gen_spectral_profile(uint32_t n, const float* wavelengths, const float* values, float scale, int computeICdf,
                                         PltptSpectralProfile* out);
/* Parameters of the ten PLT materials (Source/Falcor/Scene/Material/PLT/ *Material.h setters); unused fields are ignored. */
typedef struct PltptMaterialParams {
    uint32_t type;                 /* PLTPT_MATERIAL_* (DIFFUSE .. MULTILAYERED_STACK) */
    uint32_t doubleSided;
    float    color[3];             /* albedo (diffuse, opaque, coated opaque, grated opaque, stack) or specular tint (conductors, dielectric) */
    uint32_t emissionSpectralId;   /* diffuse emitters; PLTPT_NO_SPECTRAL_PROFILE otherwise */
    uint32_t iorNSpectralId, iorKSpectralId; /* conductors */
    float    roughness, gamma;     /* surface roughness; K-correlation exponent (conductor) */
    float    ior, extEta, dispersionB;
    float    transmission[3];      /* dielectric */
    float    tau, coatIor;         /* coating thickness in um (coated) / plate thickness in mm (thin dielectric); coat IOR */
    float    birefringence, birefringenceScale, opticAxis[2]; /* thin dielectric */
    float    gratingPitch[2], gratingDirection, gratingQ, gratingQScale, gratingPowerMultiplier; /* pitch in um */
    uint32_t gratingType, gratingLobes;
    float    layerThickness[2], layerIor[2]; /* multilayered stack, thickness in um */
    uint32_t layers;
} PltptMaterialParams;
PLTPT_API int pltpt_pack_material(const PltptMaterialParams* params, PltptMaterialBlob* out);
/* LightCollection (BuildTriangleList.cs.slang, FinalizeIntegration.cs.slang:64-110; on the GPU) + EmissivePowerSampler alias table
 * (EmissivePowerSampler.cpp:86-170, without the final random permutation) of the scene held by `h`; replaces whatever tables
 * pltpt_set_scene was given (they may have been NULL).  pltpt_get_light_collection copies the installed tables to host memory
 * (any pointer may be NULL; *count = emissive triangles; perMeshInstanceOffset has instanceCount entries).
 * Triangles are world-space, one mesh light per emissive instance in instance order (lightIdx); the face normal follows the winding in
 * world space and is flipped for instances flagged PLTPT_INSTANCE_FLAG_WORLD_FRONT_FACE_CW (Scene.slang:482-497), so a mirrored
 * instance emits from the same side as the reference's.  Bit-equal to the reference's passes on untextured emitters
 * (tests/golden/ref_light_collection.npz). */
PLTPT_API int pltpt_build_light_collection(PltptHandle h);
PLTPT_API int pltpt_get_light_collection(PltptHandle h, PltptEmissiveTriangle* triangles, PltptEmissiveFlux* flux, uint32_t* perMeshInstanceOffset,
                                         uint32_t* aliasTable, float* invWeightsSum, uint32_t capacity, uint32_t* count);

/* Animated instance transforms: new world / inverse-transpose matrices (float4x4 row-major per matrixID, like PltptSceneDesc) for
 * the scene held by `h`.  Replaces the TLAS refit / rebuild Falcor does on demand (Source/Falcor/Scene/Scene.cpp:3624-3700): the
 * single world-space BVH is rebuilt on the GPU, and a library-built light collection is re-derived. */
PLTPT_API int pltpt_update_transforms(PltptHandle h, const float* worldMatrices, const float* inverseTransposeWorldMatrices, uint32_t matrixCount,
                                      uint32_t pointersAreDevice);
/* How pltpt_update_transforms brings the BVH up to date: Scene::UpdateMode (Source/Falcor/Scene/Scene.h:169-173; the TLAS mode a host
 * sets with Scene::setTlasUpdateMode, Scene.h:935, applied at Scene.cpp:3642-3650).  REBUILD (default, like Falcor's TLAS default,
 * Scene.h:1371): a new tree.  REFIT: the tree keeps its topology, the leaf triangles are re-derived and every node box re-fitted
 * bottom-up (bvh.cu k_refit_nodes) -- a fraction of the rebuild time, hits stay exact (the triangle test decides them), traversal
 * cost grows with the distance the instances have moved since the last build, as it does for a refitted TLAS. */
enum PltptBvhUpdateMode { PLTPT_BVH_UPDATE_REBUILD = 0, PLTPT_BVH_UPDATE_REFIT = 1 };
PLTPT_API int pltpt_set_bvh_update_mode(PltptHandle h, int mode);

/* ------------------------------------------------------------------------------------------
 * Environment-map queries (parity checks of EnvMap.eval, EnvMap.slang:57-74, and EnvMapSampler.sample / evalPdf,
 * EnvMapSampler.slang:71-162).  Device pointers.  eval: n world directions float3 -> float4 {radiance rgb, pdf};
 * sample: n random pairs float2 -> 2 x float4 {dir xyz, pdf}, {Le rgb, 0}.  Fails when the scene has no env map.
 * ---------------------------------------------------------------------------------------- */
PLTPT_API int pltpt_envmap_eval(PltptHandle h, uint32_t n, const float* dirs, float* rgbPdf, void* stream);
PLTPT_API int pltpt_envmap_sample(PltptHandle h, uint32_t n, const float* rnd, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Material-texture query (parity check of MaterialSystem::sampleTexture at LOD 0, MaterialSystem.slang:283-291): n texture
 * coordinates float2 -> float4 texture values through TextureHandle `handle` (Uniform handle: `uniformValue`) and sampler
 * `samplerID` (>= samplerCount: the default linear / wrap sampler).  Device pointers.
 * ---------------------------------------------------------------------------------------- */
PLTPT_API int pltpt_texture_sample(PltptHandle h, uint32_t handle, uint32_t samplerID, const float uniformValue[4], uint32_t n, const float* uv, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Ray queries against the scene BVH (replaces TraceRay, pltpt_sample.rt.slang:174-183,
 * ShadowRays.slangh:68-77,115-124).  Rays are float[8] {ox,oy,oz,tmin,dx,dy,dz,tmax}.
 * Closest hit writes uint4 {instanceID, primitiveIndex, asuint(u), asuint(v)} + t;
 * a miss writes instanceID = 0xffffffff.  Any-hit writes 1 (occluded) / 0 (visible).
 * Pointers are device pointers.
 * ---------------------------------------------------------------------------------------- */
PLTPT_API int pltpt_trace_closest(PltptHandle h, uint32_t rayCount, const float* rays,
                                  uint32_t* hits, float* hitT, void* stream);
PLTPT_API int pltpt_trace_any(PltptHandle h, uint32_t rayCount, const float* rays,
                              uint32_t* occluded, void* stream);
/* Which traversal driver the two ray queries above run through.  The rendering kernels pick by scene size: one ray per thread to
 * completion (children pushed in order of distance) for trees of a few nodes, persistent lanes that refill from a queue otherwise
 * (nearest child first, leaf triangles of a round tested by all lanes together); the rays the connect and manifold kernels trace inline
 * run one ray per thread with the nearest-first push.  AUTO (default) makes the queries follow the trace kernels' rule, the other three
 * force one driver so that the hit-ID parity checks cover every combination on every scene. */
enum { PLTPT_TRACE_AUTO = 0, PLTPT_TRACE_ONE_RAY_PER_THREAD = 1, PLTPT_TRACE_PERSISTENT_LANES = 2, PLTPT_TRACE_ONE_RAY_PER_THREAD_NEAREST_FIRST = 3 };
PLTPT_API int pltpt_set_trace_mode(PltptHandle h, int mode);

/* ------------------------------------------------------------------------------------------
 * Batched PLT BSDF queries (config 1 of the benchmark; IMaterialInstance contract,
 * Source/Falcor/Rendering/Materials/IMaterialInstance.slang:121-183).
 *
 * Queries use the canonical shading frame T=(1,0,0) B=(0,1,0) N=(0,0,1) at the origin.
 * wi = direction to the light (ShadingData.V in the solve stage), wo = direction to the eye.
 * For every query and each of `groups` hero wavelengths the HWSS set (SpectrumSampler.slang:44-54)
 * is formed, an area beam (A = 1 mm^2, 1 m away, PLT.slang:273-290) is sourced along -wi,
 * propagated to the surface and the BSDF is evaluated on the four basis Stokes vectors.
 * Output: mueller[query][group][hwss][row][col], i.e. column c is the Stokes vector that
 * eval() produces from basis vector e_c.  All pointers are device pointers.
 * ---------------------------------------------------------------------------------------- */
PLTPT_API int pltpt_bsdf_eval_mueller(PltptHandle h, uint32_t materialID, uint32_t lobe,
                                      uint32_t queryCount, uint32_t groups,
                                      const float* wi /*3n*/, const float* wo /*3n*/,
                                      const float* heroWavelengths /*groups*n*/,
                                      float* mueller /*n*groups*HWSS*16*/, void* stream);

/* The same query evaluated with the arithmetic of the forward replay (k_unwind_step is compiled with approximate division / square root and
 * contracted FMAs): lets the BSDF tolerance tests state how far the renderer's replay is from the oracle on the same
 * queries.  Same arguments and output layout; a test hook, not a benchmark path. */
PLTPT_API int pltpt_bsdf_eval_mueller_replay_math(PltptHandle h, uint32_t materialID, uint32_t lobe, uint32_t queryCount, uint32_t groups,
                                                  const float* wi, const float* wo, const float* heroWavelengths, float* mueller, void* stream);

/* bsdf.sample() + bsdf.pdf(): per query a sample generator seeded SampleGenerator(pixel=(i&0xffff,i>>16), seed)
 * is used.  out = {wo.x, wo.y, wo.z, pdf, asfloat(lobe), weight, valid(0/1), pdf(wi,wo) re-evaluated} per query. */
PLTPT_API int pltpt_bsdf_sample(PltptHandle h, uint32_t materialID, uint32_t lobeMask, uint32_t queryCount,
                                const float* wi /*3n*/, const float* wavelengths /*n*/, uint32_t seed,
                                float* out /*8n*/, void* stream);

/* ------------------------------------------------------------------------------------------
 * Counters (kernel launches, rays traced, BVH nodes visited ...) for bench.py / DESIGN.md.
 * ---------------------------------------------------------------------------------------- */
enum PltptKernelClass {
    PLTPT_K_RAYGEN = 0, PLTPT_K_TRACE, PLTPT_K_SORT, PLTPT_K_SHADE, PLTPT_K_CONNECT_A, PLTPT_K_SHADOW_TRACE, PLTPT_K_CONNECT_B,
    PLTPT_K_MNEE, PLTPT_K_UNWIND_STEP, PLTPT_K_RESOLVE, PLTPT_K_OTHER, PLTPT_KERNEL_CLASS_COUNT
};
typedef struct PltptStats {
    uint64_t kernelLaunches;     /* launches of our kernels since the last reset */
    uint64_t raysClosest, raysAny;
    uint64_t bvhNodes, bvhLeafTriangles, triangleCount;
    uint64_t pathVertices;       /* recorded bounce vertices of the last frame */
    uint64_t bsdfEvals;          /* IMaterialInstance.eval calls x wavelengths of the last frame (if counted) */
    double   bvhBuildMs;
    uint64_t traversalStackOverflows; /* rays whose traversal stack overflowed since create (must stay 0) */
    uint64_t mneeIterations;          /* manifold-walk proposal rays of the last frame */
    /* accumulated since the last pltpt_reset_stats (device timers: CUDA events on the launching stream) */
    uint64_t framesRendered;
    double   sampleStageMs, solveStageMs;      /* summed kernel durations of the two stages */
    uint64_t totalPathVertices;                /* sum over paths of L (recorded vertices) */
    uint64_t totalReplaySteps;                 /* sum over paths of L(L-1)/2: forward-replay vertex fetches of the NEE connections */
    uint64_t totalBsdfEvals;                   /* eval calls x wavelengths */
    uint64_t traversalNodesVisited, traversalTrianglesTested; /* BVH work of both stages */
    uint64_t droppedConnections;               /* connection jobs that did not fit the job buffer (must stay 0) */
    /* per-kernel device time, filled only while pltpt_set_profiling(h, 1) is in effect: one CUDA event after every launch on the
     * launching stream; the time between consecutive events is booked to the kernel that ran between them. */
    double   kernelMs[PLTPT_KERNEL_CLASS_COUNT];
    uint64_t kernelLaunchCount[PLTPT_KERNEL_CLASS_COUNT];
    uint64_t unwindJobSteps;                   /* work items of k_unwind_step: one generalized ray pushed through one recorded vertex */
    uint64_t mneeWalks;                        /* manifold walks parked for k_mnee in the last tile of the last frame */
    double   bvhSahCost;                       /* SAH cost of the wide BVH relative to the root area (c_node = 1, c_prim = 0.3): the
                                                  minimum the collapse could reach on the binary tree (bvh.cu k_sah_costs) */
    double   bvhRefitMs;                       /* device time of the last refit (pltpt_update_transforms in PLTPT_BVH_UPDATE_REFIT mode) */
} PltptStats;
PLTPT_API int pltpt_get_stats(PltptHandle h, PltptStats* out);
PLTPT_API int pltpt_reset_stats(PltptHandle h);
/* Per-kernel timers on/off (default off).  Costs one cudaEventRecord per launch; no effect on results. */
PLTPT_API int pltpt_set_profiling(PltptHandle h, int enable);

#ifdef __cplusplus
}
#endif
#endif /* PLTPT_H_ */
