"""Falcor GPU-buffer layouts and C-ABI structs of include/pltpt.h as numpy dtypes / ctypes structures.

Every dtype mirrors a struct in include/pltpt.h (which cites the reference definition); sizes are asserted.
"""
import ctypes as C

import numpy as np

NO_SPECTRAL_PROFILE = 0xFFFFFFFF

# MaterialType — Source/Falcor/Scene/Material/MaterialTypes.slang:37-54
MATERIAL_DIFFUSE = 2
MATERIAL_CONDUCTOR = 3
MATERIAL_DIELECTRIC = 4
MATERIAL_OPAQUE_DIELECTRIC = 5
MATERIAL_THIN_DIELECTRIC = 6
MATERIAL_COATED_CONDUCTOR = 7
MATERIAL_COATED_OPAQUE_DIELECTRIC = 8
MATERIAL_GRATED_CONDUCTOR = 9
MATERIAL_GRATED_OPAQUE_DIELECTRIC = 10
MATERIAL_MULTILAYERED_STACK = 11

# LobeType — Source/Falcor/Rendering/Materials/LobeType.slang:36-68
LOBE_DIFFUSE_REFLECTION = 0x01
LOBE_SCATTER_REFLECTION = 0x02
LOBE_DELTA_REFLECTION = 0x04
LOBE_DISPERSIVE_DELTA_REFLECTION = 0x08
LOBE_SCATTER_TRANSMISSION = 0x20
LOBE_DELTA_TRANSMISSION = 0x40
LOBE_DISPERSIVE_DELTA_TRANSMISSION = 0x80
LOBE_DELTA = 0xCC
LOBE_NON_DELTA = 0x33
LOBE_SCATTER = 0x22
LOBE_REFLECTION = 0x0F
LOBE_TRANSMISSION = 0xF0
LOBE_ALL = 0xFFFF

packed_vertex_dtype = np.dtype([("position", "<f4", 3), ("packed", "<u4", 3), ("texCrd", "<f4", 2)])
geometry_instance_dtype = np.dtype([(n, "<u4") for n in ("flags", "globalMatrixID", "materialID", "geometryID", "vbOffset", "ibOffset", "instanceIndex", "geometryIndex")])
mesh_desc_dtype = np.dtype([(n, "<u4") for n in ("vbOffset", "ibOffset", "vertexCount", "indexCount", "skinningVbOffset", "prevVbOffset", "materialID", "flags")])
material_blob_dtype = np.dtype([
    ("header", "<u4", 2), ("flags", "<u4"),
    ("baseColor", "<f2", 4), ("specular", "<f2", 4), ("data1", "<u2", 4), ("data2", "<f2", 4),
    ("specularTransmission", "<f2"), ("transmission", "<f2", 3), ("diffuseTransmission", "<f2"),
    ("volumeScattering", "<f2", 3), ("IoR", "<f2"), ("volumeAbsorption", "<f2", 3), ("volumeAnisotropy", "<f2"),
    ("_pad0", "<u2"),
    ("emissionSpectralId", "<u4"), ("iorNSpectralId", "<u4"), ("iorKSpectralId", "<u4"),
    ("displacementScale", "<f4"), ("displacementOffset", "<f4"),
    ("tex", "<u4", 8), ("_pad1", "<u4"),
])
spectral_profile_dtype = np.dtype([("rgb", "<f4", 3), ("intensity", "<f4"), ("bins", "<u4"), ("minWavelength", "<f4"), ("maxWavelength", "<f4"), ("pdf", "<f4", 128), ("icdf", "<f4", 128)])
light_data_dtype = np.dtype([
    ("posW", "<f4", 3), ("type", "<u4"), ("dirW", "<f4", 3), ("openingAngle", "<f4"), ("cosOpeningAngle", "<f4"),
    ("cosSubtendedAngle", "<f4"), ("penumbraAngle", "<f4"), ("intensitySpectrumId", "<u4"),
    ("tangent", "<f4", 3), ("surfaceArea", "<f4"), ("bitangent", "<f4", 3), ("_pad", "<f4"),
    ("transMat", "<f4", 16), ("transMatIT", "<f4", 16),
])
emissive_triangle_dtype = np.dtype([("posAndTexCoords", "<f4", (3, 4)), ("normal", "<u4"), ("area", "<u4"), ("materialID", "<u4"), ("lightIdx", "<u4")])
emissive_flux_dtype = np.dtype([("flux", "<f4"), ("averageRadiance", "<f4", 3)])
bounce_dtype = np.dtype([("hit", "<u4", 4), ("encodedWi", "<f4", 2), ("sampledBsdfLobe", "<u4"), ("flags", "<u4"), ("pathContribution", "<f4"), ("lastNonDeltaPdf", "<f4")])

assert packed_vertex_dtype.itemsize == 32
assert geometry_instance_dtype.itemsize == 32
assert mesh_desc_dtype.itemsize == 32
assert material_blob_dtype.itemsize == 128, material_blob_dtype.itemsize
assert spectral_profile_dtype.itemsize == 1052
assert light_data_dtype.itemsize == 208
assert emissive_triangle_dtype.itemsize == 64
assert emissive_flux_dtype.itemsize == 16
assert bounce_dtype.itemsize == 40


class PltptOptions(C.Structure):
    """The 24 pass properties, PLTPT.cpp:102-136; defaults PLTPT.h:74-108."""
    _fields_ = [
        ("maxBounces", C.c_uint32), ("tileSize", C.c_uint32), ("HWSS", C.c_uint32), ("HWSSDoMIS", C.c_uint32),
        ("sampleGenerator", C.c_uint32), ("debugView", C.c_uint32), ("debugViewIntensity", C.c_float),
        ("sourcingAreaFromEmissiveGeometry", C.c_float), ("sourcingMaxBeamOmega", C.c_float),
        ("useDirectLights", C.c_uint32), ("useEnvLights", C.c_uint32), ("useEmissiveLights", C.c_uint32), ("useAnalyticLights", C.c_uint32),
        ("alphaMasking", C.c_uint32), ("doNEE", C.c_uint32), ("doMIS", C.c_uint32), ("emissiveSampler", C.c_uint32),
        ("doRussianRoulette", C.c_uint32), ("NEEUsePerTileSG", C.c_uint32), ("doImportanceSampleEmitters", C.c_uint32),
        ("doMNEE", C.c_uint32), ("MNEEMaxOccluders", C.c_uint32), ("MNEEMaxIterations", C.c_uint32), ("MNEESolverThreshold", C.c_float),
    ]


OPTION_DEFAULTS = dict(
    maxBounces=32, tileSize=512, HWSS=4, HWSSDoMIS=True, sampleGenerator=1, debugView=0, debugViewIntensity=1.0,
    sourcingAreaFromEmissiveGeometry=1.0, sourcingMaxBeamOmega=0.025, useDirectLights=True, useEnvLights=True,
    useEmissiveLights=True, useAnalyticLights=True, alphaMasking=True, doNEE=True, doMIS=True, emissiveSampler=2,
    doRussianRoulette=True, NEEUsePerTileSG=False, doImportanceSampleEmitters=True, doMNEE=True, MNEEMaxOccluders=2,
    MNEEMaxIterations=60, MNEESolverThreshold=5e-5,
)


def make_options(**kw) -> PltptOptions:
    """Builds a PltptOptions from the reference defaults plus overrides. Unknown keys raise (the reference logs an error, PLTPT.cpp:173)."""
    o = PltptOptions()
    vals = dict(OPTION_DEFAULTS)
    for k, v in kw.items():
        if k not in vals:
            raise KeyError(f"Unknown field '{k}' in PLTPathTracer dictionary.")
        vals[k] = v
    for k, v in vals.items():
        setattr(o, k, type(getattr(o, k))(v))
    return o


class PltptCameraData(C.Structure):
    """CameraData, Source/Falcor/Scene/Camera/CameraData.slang:36-69 (656 B)."""
    _fields_ = [
        ("viewMat", C.c_float * 16), ("prevViewMat", C.c_float * 16), ("projMat", C.c_float * 16), ("viewProjMat", C.c_float * 16),
        ("invViewProj", C.c_float * 16), ("viewProjMatNoJitter", C.c_float * 16), ("prevViewProjMatNoJitter", C.c_float * 16), ("projMatNoJitter", C.c_float * 16),
        ("posW", C.c_float * 3), ("focalLength", C.c_float),
        ("prevPosW", C.c_float * 3), ("_padding0", C.c_float),
        ("up", C.c_float * 3), ("aspectRatio", C.c_float),
        ("target", C.c_float * 3), ("nearZ", C.c_float),
        ("cameraU", C.c_float * 3), ("farZ", C.c_float),
        ("cameraV", C.c_float * 3), ("jitterX", C.c_float),
        ("cameraW", C.c_float * 3), ("jitterY", C.c_float),
        ("frameHeight", C.c_float), ("frameWidth", C.c_float), ("focalDistance", C.c_float), ("apertureRadius", C.c_float),
        ("shutterSpeed", C.c_float), ("ISOSpeed", C.c_float), ("polarizerTheta", C.c_float), ("polarizer", C.c_float),
    ]


assert C.sizeof(PltptCameraData) == 656
assert C.sizeof(PltptOptions) == 96


class PltptMaterialParams(C.Structure):
    """Parameters of the ten PLT materials for pltpt_pack_material (include/pltpt.h)."""
    _fields_ = [
        ("type", C.c_uint32), ("doubleSided", C.c_uint32), ("color", C.c_float * 3), ("emissionSpectralId", C.c_uint32),
        ("iorNSpectralId", C.c_uint32), ("iorKSpectralId", C.c_uint32), ("roughness", C.c_float), ("gamma", C.c_float),
        ("ior", C.c_float), ("extEta", C.c_float), ("dispersionB", C.c_float), ("transmission", C.c_float * 3),
        ("tau", C.c_float), ("coatIor", C.c_float), ("birefringence", C.c_float), ("birefringenceScale", C.c_float), ("opticAxis", C.c_float * 2),
        ("gratingPitch", C.c_float * 2), ("gratingDirection", C.c_float), ("gratingQ", C.c_float), ("gratingQScale", C.c_float), ("gratingPowerMultiplier", C.c_float),
        ("gratingType", C.c_uint32), ("gratingLobes", C.c_uint32), ("layerThickness", C.c_float * 2), ("layerIor", C.c_float * 2), ("layers", C.c_uint32),
    ]


class PltptEnvMapData(C.Structure):
    """EnvMapData, Source/Falcor/Scene/Lights/EnvMapData.slang:36-43 (112 B); float3x4 row-major."""
    _fields_ = [("transform", C.c_float * 12), ("invTransform", C.c_float * 12), ("tint", C.c_float * 3), ("intensity", C.c_float)]


assert C.sizeof(PltptEnvMapData) == 112


class PltptTextureDesc(C.Structure):
    """One material texture (mip 0, linear layout); include/pltpt.h."""
    _fields_ = [("texels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32), ("_pad", C.c_uint32)]


class PltptSamplerDesc(C.Structure):
    """Sampler state (Core/API/Sampler.h:50-66); include/pltpt.h."""
    _fields_ = [("magFilter", C.c_uint32), ("addressU", C.c_uint32), ("addressV", C.c_uint32), ("_pad", C.c_uint32), ("borderColor", C.c_float * 4)]


TEXTURE_FORMATS = {"rgba32f": 0, "rgba8": 1, "rgba8_srgb": 2, "rgba16f": 3}
assert C.sizeof(PltptTextureDesc) == 24 and C.sizeof(PltptSamplerDesc) == 32


class PltptSceneDesc(C.Structure):
    _fields_ = [
        ("vertices", C.c_void_p), ("vertexCount", C.c_uint32),
        ("indexData", C.c_void_p), ("indexDataBytes", C.c_uint64),
        ("geometryInstances", C.c_void_p), ("instanceCount", C.c_uint32),
        ("meshes", C.c_void_p), ("meshCount", C.c_uint32),
        ("worldMatrices", C.c_void_p), ("inverseTransposeWorldMatrices", C.c_void_p), ("matrixCount", C.c_uint32),
        ("materialData", C.c_void_p), ("materialCount", C.c_uint32),
        ("spectralProfiles", C.c_void_p), ("profileCount", C.c_uint32),
        ("lights", C.c_void_p), ("lightCount", C.c_uint32),
        ("emissiveTriangles", C.c_void_p), ("emissiveTriangleCount", C.c_uint32),
        ("emissiveFlux", C.c_void_p),
        ("activeTriangles", C.c_void_p), ("activeTriangleCount", C.c_uint32),
        ("perMeshInstanceOffset", C.c_void_p),
        ("triangleAliasTable", C.c_void_p), ("invWeightsSum", C.c_float),
        ("hitInfoTypeBits", C.c_uint32), ("hitInfoInstanceIDBits", C.c_uint32), ("hitInfoPrimitiveIndexBits", C.c_uint32),
        ("pointersAreDevice", C.c_uint32),
        ("envMapTexels", C.c_void_p), ("envMapWidth", C.c_uint32), ("envMapHeight", C.c_uint32),
        ("envMap", PltptEnvMapData),
        ("textures", C.c_void_p), ("textureCount", C.c_uint32),
        ("samplers", C.c_void_p), ("samplerCount", C.c_uint32),
    ]


# enum PltptKernelClass (include/pltpt.h)
KERNEL_CLASSES = ("raygen", "trace", "sort", "shade", "connect_a", "shadow_trace", "connect_b", "mnee", "unwind_step", "resolve", "other")
KERNEL_CLASS_COUNT = len(KERNEL_CLASSES)


class PltptStats(C.Structure):
    _fields_ = [
        ("kernelLaunches", C.c_uint64), ("raysClosest", C.c_uint64), ("raysAny", C.c_uint64),
        ("bvhNodes", C.c_uint64), ("bvhLeafTriangles", C.c_uint64), ("triangleCount", C.c_uint64),
        ("pathVertices", C.c_uint64), ("bsdfEvals", C.c_uint64), ("bvhBuildMs", C.c_double),
        ("traversalStackOverflows", C.c_uint64), ("mneeIterations", C.c_uint64),
        ("framesRendered", C.c_uint64), ("sampleStageMs", C.c_double), ("solveStageMs", C.c_double),
        ("totalPathVertices", C.c_uint64), ("totalReplaySteps", C.c_uint64), ("totalBsdfEvals", C.c_uint64),
        ("traversalNodesVisited", C.c_uint64), ("traversalTrianglesTested", C.c_uint64), ("droppedConnections", C.c_uint64),
        ("kernelMs", C.c_double * KERNEL_CLASS_COUNT), ("kernelLaunchCount", C.c_uint64 * KERNEL_CLASS_COUNT), ("unwindJobSteps", C.c_uint64), ("mneeWalks", C.c_uint64),
        ("bvhSahCost", C.c_double), ("bvhRefitMs", C.c_double),
    ]
