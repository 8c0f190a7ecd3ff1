// PLTPT.cpp — host side of the drop-in `PLTPT` render pass: dictionary parsing, reflection, scene hand-over and the per-frame
// call into the CUDA library.  Behaviour follows the reference pass (Source/RenderPasses/PLTPT/PLTPT.cpp, cited per function);
// the implementation is ours: the reference builds two DXR programs and loops over tiles (PLTPT.cpp:487-495), this file makes
// one pltpt_render_frame() call per execute().
#include "PLTPT.h"

#include <string>

namespace {
// reflected channels: PLTPT.cpp:65-78 (names and formats are what scripts/PLTPathTracer.py:37-43 wires)
const char kInputVBuffer[] = "vbuffer";
const char kInputViewDir[] = "viewW";
const char kOutputNormal[] = "normal";
const char kOutputAlbedo[] = "albedo";
const char kOutputColor[] = "color";
const ChannelList kInputChannels = {
    {kInputVBuffer, "gVBuffer", "Visibility buffer in packed format"},
    {kInputViewDir, "gViewW", "World-space view direction (xyz float format)", true /* optional */},
};
const ChannelList kOutputChannels = {
    {kOutputNormal, "gOutputNormal", "Output normal", false, ResourceFormat::RGBA32Float},
    {kOutputAlbedo, "gOutputAlbedo", "Output albedo", false, ResourceFormat::RGBA32Float},
    {kOutputColor, "gOutputColor", "Output color (sum of direct and indirect)", false, ResourceFormat::RGBA32Float},
};

// One row per pass property (PLTPT.cpp:102-136): the dictionary key and how it maps onto PltptOptions.
enum class Kind { U32, F32, Bool, Sampler };
struct Property { const char* key; Kind kind; size_t offset; };
#define PROP(name, kind) {#name, Kind::kind, offsetof(PltptOptions, name)}
const Property kProperties[] = {
    PROP(maxBounces, U32), PROP(tileSize, U32), PROP(debugView, U32), PROP(debugViewIntensity, F32), PROP(sourcingAreaFromEmissiveGeometry, F32),
    PROP(sourcingMaxBeamOmega, F32), PROP(sampleGenerator, U32), PROP(HWSS, U32), PROP(HWSSDoMIS, Bool), PROP(useDirectLights, Bool),
    PROP(useEnvLights, Bool), PROP(useEmissiveLights, Bool), PROP(useAnalyticLights, Bool), PROP(alphaMasking, Bool), PROP(doNEE, Bool),
    PROP(NEEUsePerTileSG, Bool), PROP(doImportanceSampleEmitters, Bool), PROP(doMNEE, Bool), PROP(doMIS, Bool), PROP(emissiveSampler, Sampler),
    PROP(doRussianRoulette, Bool), PROP(MNEEMaxOccluders, U32), PROP(MNEEMaxIterations, U32), PROP(MNEESolverThreshold, F32),
};
#undef PROP

void check(int status, const char* what) {
    if (status != PLTPT_OK) throw RuntimeError(std::string(what) + ": " + pltpt_last_error());
}

#ifndef PLTPT_WITH_FALCOR
// Shim versions of the three Falcor glue points (INTEGRATION.md gives the Falcor versions).
float* cudaPointer(const Texture::SharedPtr& t) { return t ? static_cast<float*>(t->getCudaDevicePtr()) : nullptr; }
void* cudaStreamOf(RenderContext* ctx) { return ctx ? static_cast<void*>(ctx->getCudaStream()) : nullptr; }
int cudaDeviceOf(const std::shared_ptr<Device>& d) { return d ? d->getCudaDevice() : 0; }
#endif
}  // namespace

extern "C" FALCOR_API_EXPORT void registerPlugin(Falcor::PluginRegistry& registry) {  // PLTPT.cpp:43-46
    registry.registerClass<RenderPass, PLTPT>();
#ifdef PLTPT_WITH_FALCOR
    ScriptBindings::registerBinding([](pybind11::module& m) {
        pybind11::class_<PLTPT, RenderPass, PLTPT::SharedPtr> pass(m, "PLTPT");
        pass.def_property("sourcingAreaFromEmissiveGeometry", &PLTPT::getSourcingAreaFromEmissiveGeometry, &PLTPT::setSourcingAreaFromEmissiveGeometry);
    });
#endif
}

PLTPT::PLTPT(std::shared_ptr<Device> pDevice) : RenderPass(std::move(pDevice)) { pltpt_default_options(&mOptions); }

PLTPT::~PLTPT() {
    if (mHandle) pltpt_destroy(mHandle);
}

PLTPT::SharedPtr PLTPT::create(std::shared_ptr<Device> pDevice, const Dictionary& dict) {  // PLTPT.cpp:141-145
    SharedPtr p = SharedPtr(new PLTPT(std::move(pDevice)));
    p->parseDictionary(dict);
    return p;
}

// PLTPT.cpp:147-175: every known key is taken, an unknown key is logged as an error and otherwise ignored.
void PLTPT::parseDictionary(const Dictionary& dict) {
    for (const auto& [key, value] : dict) {
        const Property* prop = nullptr;
        for (const Property& p : kProperties)
            if (key == p.key) { prop = &p; break; }
        if (!prop) { logError("Unknown field '" + key + "' in PLTPathTracer dictionary."); continue; }
        char* field = reinterpret_cast<char*>(&mOptions) + prop->offset;
        switch (prop->kind) {
            case Kind::U32: *reinterpret_cast<uint32_t*>(field) = uint32_t(value); break;
            case Kind::F32: *reinterpret_cast<float*>(field) = float(value); break;
            case Kind::Bool: *reinterpret_cast<uint32_t*>(field) = bool(value) ? 1u : 0u; break;
            case Kind::Sampler: *reinterpret_cast<uint32_t*>(field) = uint32_t(EmissiveLightSamplerType(value)); break;
        }
    }
}

Dictionary PLTPT::getScriptingDictionary() {  // PLTPT.cpp:177-204
    Dictionary d;
    for (const Property& p : kProperties) {
        const char* field = reinterpret_cast<const char*>(&mOptions) + p.offset;
        switch (p.kind) {
            case Kind::U32: d[p.key] = *reinterpret_cast<const uint32_t*>(field); break;
            case Kind::F32: d[p.key] = *reinterpret_cast<const float*>(field); break;
            case Kind::Bool: d[p.key] = *reinterpret_cast<const uint32_t*>(field) != 0; break;
            case Kind::Sampler: d[p.key] = EmissiveLightSamplerType(*reinterpret_cast<const uint32_t*>(field)); break;
        }
    }
    return d;
}

RenderPassReflection PLTPT::reflect(const CompileData&) {  // PLTPT.cpp:206-214
    RenderPassReflection reflector;
    addRenderPassInputs(reflector, kInputChannels);
    addRenderPassOutputs(reflector, kOutputChannels);
    return reflector;
}

void PLTPT::renderUI(Gui::Widgets&) {
    // The reference exposes every property as a widget and sets mOptionsChanged when one is edited (PLTPT.cpp:258-355).  The
    // shim has no GUI; with Falcor the same widget list applies, writing into mOptions.
}

void PLTPT::ensureHandle() {
    if (!mHandle) check(pltpt_create(&mOptions, cudaDeviceOf(mpDevice), &mHandle), "pltpt_create");
}

// Replaces the program / binding-table construction of PLTPT::setScene (PLTPT.cpp:501-552): the scene's GPU buffers are handed
// to the CUDA library, which builds its own BVH (there is no driver acceleration structure on B200).
void PLTPT::setScene(RenderContext* pRenderContext, const std::shared_ptr<Scene>& pScene) {
    mpScene = pScene;
    mFrameCount = 0;
    mSceneUploaded = false;
    if (mpScene) {
        if (pRenderContext) pRenderContext->flush(true);  // Falcor -> CUDA hand-over point (OptixDenoiser.cpp:311-312)
        uploadScene();
    }
}

void PLTPT::onSceneUpdates(RenderContext* pRenderContext, Scene::UpdateFlags sceneUpdates) {
    if (!mpScene || !mSceneUploaded) return;
    const uint32_t u = uint32_t(sceneUpdates);
    const uint32_t structural = uint32_t(Scene::UpdateFlags::GeometryChanged) | uint32_t(Scene::UpdateFlags::MaterialsChanged) | uint32_t(Scene::UpdateFlags::EnvMapChanged) |
                                uint32_t(Scene::UpdateFlags::LightCollectionChanged);
    const uint32_t moved = uint32_t(Scene::UpdateFlags::GeometryMoved) | uint32_t(Scene::UpdateFlags::SceneGraphChanged);
    if (!(u & (structural | moved))) return;
    if (pRenderContext) pRenderContext->flush(true);
    if (u & structural) { uploadScene(); return; }
#ifdef PLTPT_WITH_FALCOR
    const PltptSceneDesc desc = pltptDescribeFalcorScene(*mpScene);
#else
    const PltptSceneDesc& desc = mpScene->getPltptSceneDesc();
#endif
    // the scene's TLAS update mode (Scene.h:935-939) decides between a rebuild and a refit of the library's BVH
    check(pltpt_set_bvh_update_mode(mHandle, mpScene->getTlasUpdateMode() == Scene::UpdateMode::Refit ? PLTPT_BVH_UPDATE_REFIT : PLTPT_BVH_UPDATE_REBUILD), "pltpt_set_bvh_update_mode");
    check(pltpt_update_transforms(mHandle, desc.worldMatrices, desc.inverseTransposeWorldMatrices, desc.matrixCount, desc.pointersAreDevice), "pltpt_update_transforms");
}

void PLTPT::uploadScene() {
    ensureHandle();
    check(pltpt_set_options(mHandle, &mOptions), "pltpt_set_options");
#ifdef PLTPT_WITH_FALCOR
    const PltptSceneDesc desc = pltptDescribeFalcorScene(*mpScene);  // INTEGRATION.md "Scene hand-over"
#else
    const PltptSceneDesc& desc = mpScene->getPltptSceneDesc();
#endif
    check(pltpt_set_scene(mHandle, &desc), "pltpt_set_scene");
    mSceneUploaded = true;
}

void PLTPT::execute(RenderContext* pRenderContext, const RenderData& renderData) {  // PLTPT.cpp:357-499
    auto& dict = renderData.getDictionary();
    const uint2& targetDim = renderData.getDefaultTextureDims();
    if (mOptionsChanged) {  // PLTPT.cpp:361-372: tell downstream passes (the accumulator) to reset
        const auto flags = dict.getValue(kRenderPassRefreshFlags, RenderPassRefreshFlags::None);
        dict[kRenderPassRefreshFlags] = flags | RenderPassRefreshFlags::RenderOptionsChanged;
        if (mHandle) check(pltpt_set_options(mHandle, &mOptions), "pltpt_set_options");
        if (mpScene) {  // the reference re-runs setScene() here, which also restarts the frame counter (PLTPT.cpp:368-369,510)
            mFrameCount = 0;
            if (!mSceneUploaded) uploadScene();
        }
        mOptionsChanged = false;
    }

    if (!mpScene) {  // PLTPT.cpp:374-384: no scene -> clear the outputs
        for (const auto& c : kOutputChannels)
            if (auto pDst = renderData.getTexture(c.name)) pRenderContext->clearTexture(pDst.get());
        return;
    }
    if (!mSceneUploaded) uploadScene();

    const bool useDOF = mpScene->getCamera()->getApertureRadius() > 0.f;  // PLTPT.cpp:419-423
    if (useDOF && renderData[kInputViewDir] == nullptr) logWarning(std::string("Depth-of-field requires the '") + kInputViewDir + "' input. Expect incorrect shading.");

    FALCOR_ASSERT(targetDim.x > 0 && targetDim.y > 0);
    const auto pColor = renderData.getTexture(kOutputColor);
    if (!pColor) throw RuntimeError("PLTPT: output 'color' is not bound");
    // optional outputs: the reference passes is_valid_<name> defines (RenderPassHelpers.h:99-118); here a null pointer
    float* color = cudaPointer(pColor);
    float* albedo = cudaPointer(renderData.getTexture(kOutputAlbedo));
    float* normal = cudaPointer(renderData.getTexture(kOutputNormal));

    pRenderContext->flush(true);  // graphics work on the shared surfaces must be done before CUDA writes them
#ifdef PLTPT_WITH_FALCOR
    const PltptCameraData camera = pltptCameraData(*mpScene->getCamera());  // CameraData is layout-identical (CameraData.slang:36-69)
#else
    const PltptCameraData& camera = mpScene->getCamera()->getData();
#endif
    check(pltpt_render_frame(mHandle, mFrameCount, &camera, targetDim.x, targetDim.y, color, albedo, normal, cudaStreamOf(pRenderContext)), "pltpt_render_frame");
    mFrameCount++;  // PLTPT.cpp:498
}
