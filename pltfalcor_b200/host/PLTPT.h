// PLTPT.h — the drop-in `PLTPT` Falcor RenderPass, host side.  Same class name, plugin type string, public interface,
// 24 dictionary properties, reflected channels and python property as the reference pass
// (Source/RenderPasses/PLTPT/PLTPT.h:44-125, PLTPT.cpp:37-214); the two DXR programs and the bounce buffer are replaced by
// one PltptHandle of the CUDA library (include/pltpt.h).  Not a copy of the reference class: there are no Slang programs,
// defines or binding tables here — options are run-time fields of PltptOptions.
#pragma once

#ifdef PLTPT_WITH_FALCOR
#include "Falcor.h"
#include "Rendering/Lights/EmissiveLightSampler.h"
#else
#include "FalcorShim.h"
#endif
#include "pltpt.h"

using namespace Falcor;

class PLTPT : public RenderPass {
public:
    FALCOR_PLUGIN_CLASS(PLTPT, "PLTPT", PluginInfo{"Physical Light Transport Path Tracer."});

    using SharedPtr = std::shared_ptr<PLTPT>;

    static SharedPtr create(std::shared_ptr<Device> pDevice, const Dictionary& dict);
    ~PLTPT() override;

    Dictionary getScriptingDictionary() override;
    RenderPassReflection reflect(const CompileData& compileData) override;
    void compile(RenderContext*, const CompileData&) override {}
    void execute(RenderContext* pRenderContext, const RenderData& renderData) override;
    void renderUI(Gui::Widgets& widget) override;
    void setScene(RenderContext* pRenderContext, const std::shared_ptr<Scene>& pScene) override;
    // The reference leaves scene changes to the driver (Scene::raytrace refits / rebuilds the TLAS on demand, Scene.cpp:3624-3700); the
    // CUDA library owns its BVH, so the pass forwards them: moved instances -> pltpt_update_transforms, anything structural -> re-upload.
    void onSceneUpdates(RenderContext* pRenderContext, Scene::UpdateFlags sceneUpdates) override;
    bool onMouseEvent(const MouseEvent&) override { return false; }
    bool onKeyEvent(const KeyboardEvent&) override { return false; }

    // python property `sourcingAreaFromEmissiveGeometry` (PLTPT.cpp:37-41)
    void setSourcingAreaFromEmissiveGeometry(float area) { mOptions.sourcingAreaFromEmissiveGeometry = area; mOptionsChanged = true; }
    float getSourcingAreaFromEmissiveGeometry() const { return mOptions.sourcingAreaFromEmissiveGeometry; }

    // not part of the reference interface: introspection for the host harness / tests
    uint32_t getFrameCount() const { return mFrameCount; }
    const PltptOptions& getOptions() const { return mOptions; }
    PltptHandle getHandle() const { return mHandle; }

private:
    explicit PLTPT(std::shared_ptr<Device> pDevice);
    void parseDictionary(const Dictionary& dict);
    void ensureHandle();
    void uploadScene();

    std::shared_ptr<Scene> mpScene;
    PltptOptions mOptions{};
    PltptHandle mHandle = nullptr;
    bool mSceneUploaded = false;
    uint32_t mFrameCount = 0;      // frames since the scene was set (PLTPT.cpp:498,510): the random-number seed of a frame
    bool mOptionsChanged = false;
};
