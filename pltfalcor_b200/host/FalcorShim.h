// FalcorShim.h — the slice of Falcor's host API that the PLTPT render pass touches, restated as a header-only stand-in so
// that host/PLTPT.{h,cpp} compile and run in this repository, where Falcor itself cannot be built (it needs packman
// downloads, a windowing system and D3D12/Vulkan; DESIGN.md "Boundary").  Built against real Falcor (-DPLTPT_WITH_FALCOR)
// this header is not included: PLTPT.h includes "Falcor.h" instead and INTEGRATION.md lists the three glue functions
// (texture -> CUDA pointer, scene -> PltptSceneDesc, render-context flush) that replace the shim's trivial versions.
//
// Every type cites the Falcor declaration it stands in for (paths relative to the reference's Source/Falcor).  Only the
// members PLTPT uses are present; semantics of those members follow Falcor (e.g. Dictionary values convert implicitly,
// PluginRegistry::registerClass keys on T::kPluginType, logError does not throw).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <variant>
#include <vector>

#include "pltpt.h"

#define FALCOR_API_EXPORT __attribute__((visibility("default")))
#define FALCOR_ASSERT(x) do { if (!(x)) throw ::Falcor::RuntimeError(std::string("assertion failed: ") + #x); } while (0)

namespace Falcor {
using uint = uint32_t;
using uint2 = ::uint2;  // Falcor's uint2 is glm's; the shim borrows CUDA's vector type (same x,y members) to avoid two global names
using uint3 = ::uint3;

// Core/Errors.h:80-120
class RuntimeError : public std::runtime_error { public: using std::runtime_error::runtime_error; };
class ArgumentError : public std::runtime_error { public: using std::runtime_error::runtime_error; };

// Utils/Logger.h:47-57 — messages go to stderr; counts are kept so that tests can observe "logError, not throw" (PLTPT.cpp:173)
struct Logger {
    static int& errorCount() { static int n = 0; return n; }
    static int& warningCount() { static int n = 0; return n; }
    static std::string& lastMessage() { static std::string s; return s; }
};
inline void logMessage(const char* level, const std::string& msg) { Logger::lastMessage() = msg; std::fprintf(stderr, "(%s) %s\n", level, msg.c_str()); }
inline void logError(const std::string& msg) { ++Logger::errorCount(); logMessage("Error", msg); }
inline void logWarning(const std::string& msg) { ++Logger::warningCount(); logMessage("Warning", msg); }

// Utils/Scripting/Dictionary.h — string keys, dynamically typed values with implicit conversion on read
class Dictionary {
public:
    class Value {
    public:
        Value() = default;
        Value(bool v) : mV(v) {}
        Value(int v) : mV(int64_t(v)) {}
        Value(uint32_t v) : mV(int64_t(v)) {}
        Value(int64_t v) : mV(v) {}
        Value(float v) : mV(double(v)) {}
        Value(double v) : mV(v) {}
        Value(const char* v) : mV(std::string(v)) {}
        Value(std::string v) : mV(std::move(v)) {}
        template <typename E, std::enable_if_t<std::is_enum_v<E>, int> = 0> Value(E v) : mV(int64_t(v)) {}
        operator bool() const { return asNumber() != 0.0; }
        operator uint32_t() const { return uint32_t(asNumber()); }
        operator int() const { return int(asNumber()); }
        operator float() const { return float(asNumber()); }
        operator double() const { return asNumber(); }
        operator std::string() const { return toString(); }
        template <typename E, std::enable_if_t<std::is_enum_v<E>, int> = 0> operator E() const { return E(int64_t(asNumber())); }
        bool isBool() const { return std::holds_alternative<bool>(mV); }
        bool isInt() const { return std::holds_alternative<int64_t>(mV); }
        bool isFloat() const { return std::holds_alternative<double>(mV); }
        std::string toString() const {
            if (auto p = std::get_if<bool>(&mV)) return *p ? "true" : "false";
            if (auto p = std::get_if<int64_t>(&mV)) return std::to_string(*p);
            if (auto p = std::get_if<double>(&mV)) { char b[64]; std::snprintf(b, sizeof b, "%.9g", *p); return b; }
            if (auto p = std::get_if<std::string>(&mV)) return *p;
            return "";
        }
    private:
        double asNumber() const {
            if (auto p = std::get_if<bool>(&mV)) return *p ? 1.0 : 0.0;
            if (auto p = std::get_if<int64_t>(&mV)) return double(*p);
            if (auto p = std::get_if<double>(&mV)) return *p;
            if (auto p = std::get_if<std::string>(&mV)) {
                if (*p == "true" || *p == "True") return 1.0;
                if (*p == "false" || *p == "False") return 0.0;
                try { return std::stod(*p); } catch (...) { throw ArgumentError("Dictionary value '" + *p + "' is not a number"); }
            }
            return 0.0;
        }
        std::variant<std::monostate, bool, int64_t, double, std::string> mV;
    };
    using Container = std::vector<std::pair<std::string, Value>>;  // insertion-ordered like the python dict it is built from
    Value& operator[](const std::string& key) {
        for (auto& kv : mItems) if (kv.first == key) return kv.second;
        mItems.emplace_back(key, Value());
        return mItems.back().second;
    }
    bool keyExists(const std::string& key) const { for (auto& kv : mItems) if (kv.first == key) return true; return false; }
    template <typename T> T getValue(const std::string& key, const T& def) const { for (auto& kv : mItems) if (kv.first == key) return T(kv.second); return def; }
    Container::const_iterator begin() const { return mItems.begin(); }
    Container::const_iterator end() const { return mItems.end(); }
    size_t size() const { return mItems.size(); }
private:
    Container mItems;
};

// Core/API/Formats.h
enum class ResourceFormat : uint32_t { Unknown = 0, RGBA32Float = 1, RGBA32Uint = 2 };

// RenderGraph/RenderPassReflection.h
class RenderPassReflection {
public:
    struct Field {
        enum class Visibility { Input = 1, Output = 2 };
        std::string name, desc;
        Visibility visibility = Visibility::Input;
        ResourceFormat format = ResourceFormat::Unknown;
        bool optional = false;
    };
    Field& addInput(const std::string& name, const std::string& desc) { mFields.push_back({name, desc, Field::Visibility::Input}); return mFields.back(); }
    Field& addOutput(const std::string& name, const std::string& desc) { mFields.push_back({name, desc, Field::Visibility::Output}); return mFields.back(); }
    size_t getFieldCount() const { return mFields.size(); }
    const Field* getField(size_t i) const { return &mFields[i]; }
    const Field* getField(const std::string& name) const { for (auto& f : mFields) if (f.name == name) return &f; return nullptr; }
private:
    std::vector<Field> mFields;
};

// RenderGraph/RenderPassHelpers.h:40-118
struct ChannelDesc {
    std::string name, texname, desc;
    bool optional = false;
    ResourceFormat format = ResourceFormat::Unknown;
};
using ChannelList = std::vector<ChannelDesc>;
inline void addRenderPassInputs(RenderPassReflection& r, const ChannelList& channels) {
    for (const auto& c : channels) { auto& f = r.addInput(c.name, c.desc); f.format = c.format; f.optional = c.optional; }
}
inline void addRenderPassOutputs(RenderPassReflection& r, const ChannelList& channels) {
    for (const auto& c : channels) { auto& f = r.addOutput(c.name, c.desc); f.format = c.format; f.optional = c.optional; }
}

// RenderGraph/RenderPassStandardFlags.h:37-47
enum class RenderPassRefreshFlags : uint32_t { None = 0x0, LightingChanged = 0x1, RenderOptionsChanged = 0x2 };
inline RenderPassRefreshFlags operator|(RenderPassRefreshFlags a, RenderPassRefreshFlags b) { return RenderPassRefreshFlags(uint32_t(a) | uint32_t(b)); }
static const char kRenderPassRefreshFlags[] = "_refreshFlags";

// Core/API/Device.h — the shim "device" is a CUDA device ordinal
class Device {
public:
    explicit Device(int cudaDevice = 0) : mCudaDevice(cudaDevice) {}
    int getCudaDevice() const { return mCudaDevice; }
private:
    int mCudaDevice;
};

// Core/API/Texture.h — stand-in: a linear RGBA32F surface in CUDA device memory.  With real Falcor this is what a render-graph
// texture becomes after the shared-handle import + texture->buffer conversion the OptixDenoiser pass already does
// (Source/RenderPasses/OptixDenoiser/OptixDenoiser.cpp:432-436,487-513).
class Texture {
public:
    using SharedPtr = std::shared_ptr<Texture>;
    static SharedPtr create2D(uint32_t w, uint32_t h, ResourceFormat fmt) {
        auto t = SharedPtr(new Texture());
        t->mW = w; t->mH = h; t->mFormat = fmt;
        if (cudaMalloc(&t->mpDevice, size_t(w) * h * 16) != cudaSuccess) throw RuntimeError("cudaMalloc failed for a shim texture");
        return t;
    }
    ~Texture() { if (mpDevice) cudaFree(mpDevice); }
    uint32_t getWidth() const { return mW; }
    uint32_t getHeight() const { return mH; }
    ResourceFormat getFormat() const { return mFormat; }
    void* getCudaDevicePtr() const { return mpDevice; }
private:
    Texture() = default;
    void* mpDevice = nullptr;
    uint32_t mW = 0, mH = 0;
    ResourceFormat mFormat = ResourceFormat::Unknown;
};

// Core/API/RenderContext.h — flush(true) is the Falcor->CUDA synchronisation point (OptixDenoiser.cpp:311-312); the shim has no
// graphics queue, so flush only synchronises the CUDA stream the pass submits to.
class RenderContext {
public:
    explicit RenderContext(cudaStream_t s = nullptr) : mStream(s) {}
    void flush(bool wait = false) { if (wait) cudaStreamSynchronize(mStream); }
    void clearTexture(Texture* t) { if (t && t->getCudaDevicePtr()) cudaMemsetAsync(t->getCudaDevicePtr(), 0, size_t(t->getWidth()) * t->getHeight() * 16, mStream); }
    cudaStream_t getCudaStream() const { return mStream; }
private:
    cudaStream_t mStream;
};

// RenderGraph/RenderData.h
class RenderData {
public:
    RenderData(uint2 dims, Dictionary& dict) : mDims(dims), mDict(dict) {}
    void setTexture(const std::string& name, Texture::SharedPtr t) { mTextures[name] = std::move(t); }
    Texture::SharedPtr getTexture(const std::string& name) const { auto it = mTextures.find(name); return it == mTextures.end() ? nullptr : it->second; }
    Texture::SharedPtr operator[](const std::string& name) const { return getTexture(name); }
    Dictionary& getDictionary() const { return mDict; }
    const uint2& getDefaultTextureDims() const { return mDims; }
private:
    uint2 mDims;
    Dictionary& mDict;
    std::map<std::string, Texture::SharedPtr> mTextures;
};

// Scene/Camera/Camera.h — carries the CameraData block (Scene/Camera/CameraData.slang:36-69) GBufferRT has already jittered
class Camera {
public:
    using SharedPtr = std::shared_ptr<Camera>;
    PltptCameraData& getData() { return mData; }
    const PltptCameraData& getData() const { return mData; }
    float getApertureRadius() const { return mData.apertureRadius; }
    void setJitter(float jx, float jy) { mData.jitterX = jx; mData.jitterY = jy; }
private:
    PltptCameraData mData{};
};

// Scene/Scene.h — stand-in: the gScene buffers the pass reads, already in the C-ABI's description (host or device pointers)
class Scene {
public:
    using SharedPtr = std::shared_ptr<Scene>;
    enum class UpdateFlags : uint32_t { None = 0, GeometryMoved = 0x1, SceneGraphChanged = 0x80, LightCollectionChanged = 0x100, MaterialsChanged = 0x200, EnvMapChanged = 0x400,
                                        GeometryChanged = 0x100000 };  // Scene/Scene.h:127-165 (the values the pass looks at)
    Scene(const PltptSceneDesc& d) : mDesc(d), mpCamera(std::make_shared<Camera>()), mHasEnvMap(d.envMapTexels != nullptr) {}
    const PltptSceneDesc& getPltptSceneDesc() const { return mDesc; }
    const Camera::SharedPtr& getCamera() const { return mpCamera; }
    bool useEmissiveLights() const { return mDesc.emissiveTriangleCount > 0; }  // Scene.h: render settings && light collection non-empty
    bool useAnalyticLights() const { return mDesc.lightCount > 0; }
    bool useEnvLight() const { return mHasEnvMap; }
    enum class UpdateMode { Rebuild, Refit };  // Scene/Scene.h:169-173
    void setTlasUpdateMode(UpdateMode mode) { mTlasUpdateMode = mode; }  // Scene.h:935
    UpdateMode getTlasUpdateMode() { return mTlasUpdateMode; }           // Scene.h:939
private:
    UpdateMode mTlasUpdateMode = UpdateMode::Rebuild;                    // Scene.h:1371
    PltptSceneDesc mDesc;
    Camera::SharedPtr mpCamera;
    bool mHasEnvMap;
};

namespace Gui { class Widgets {}; }
struct MouseEvent {};
struct KeyboardEvent {};
enum class HotReloadFlags : uint32_t { None = 0 };

class PluginRegistry;

// RenderGraph/RenderPass.h:109-221
class RenderPass {
public:
    using SharedPtr = std::shared_ptr<RenderPass>;
    using PluginCreate = std::function<std::shared_ptr<RenderPass>(std::shared_ptr<Device> pDevice, const Dictionary& dict)>;
    struct PluginInfo { std::string desc; };
    struct CompileData { uint2 defaultTexDims{0, 0}; ResourceFormat defaultTexFormat = ResourceFormat::Unknown; RenderPassReflection connectedResources; };
    static const std::string& getPluginBaseType() { static std::string type{"RenderPass"}; return type; }
    virtual const std::string& getPluginType() const = 0;
    virtual const PluginInfo& getPluginInfo() const = 0;
    virtual ~RenderPass() = default;
    const std::string& getType() const { return getPluginType(); }
    const std::string& getDesc() const { return getPluginInfo().desc; }
    virtual RenderPassReflection reflect(const CompileData& compileData) = 0;
    virtual void compile(RenderContext*, const CompileData&) {}
    virtual void execute(RenderContext* pRenderContext, const RenderData& renderData) = 0;
    virtual Dictionary getScriptingDictionary() { return {}; }
    virtual void renderUI(Gui::Widgets&) {}
    virtual void setScene(RenderContext*, const std::shared_ptr<Scene>&) {}
    virtual void onSceneUpdates(RenderContext*, Scene::UpdateFlags) {}
    virtual bool onMouseEvent(const MouseEvent&) { return false; }
    virtual bool onKeyEvent(const KeyboardEvent&) { return false; }
    virtual void onHotReload(HotReloadFlags) {}
    const std::string& getName() const { return mName; }
protected:
    explicit RenderPass(std::shared_ptr<Device> pDevice) : mpDevice(std::move(pDevice)) {}
    std::shared_ptr<Device> mpDevice;
    std::string mName;
};

// Core/Plugin.h:305-335
#define FALCOR_PLUGIN_CLASS(cls, type, info)                                              \
public:                                                                                   \
    static inline const std::string kPluginType{type};                                    \
    static inline const PluginInfo kPluginInfo{info};                                     \
    virtual const std::string& getPluginType() const final { return kPluginType; }        \
    virtual const PluginInfo& getPluginInfo() const final { return kPluginInfo; }

// Core/Plugin.h:86-97,200-250 — registry handed to a plugin library's registerPlugin()
class PluginRegistry {
public:
    template <typename BaseT, typename T> void registerClass() {
        mCreators[T::kPluginType] = [](std::shared_ptr<Device> d, const Dictionary& dict) -> std::shared_ptr<BaseT> { return T::create(std::move(d), dict); };
        mInfos[T::kPluginType] = T::kPluginInfo.desc;
    }
    bool hasClass(const std::string& type) const { return mCreators.count(type) != 0; }
    std::shared_ptr<RenderPass> createClass(const std::string& type, std::shared_ptr<Device> d, const Dictionary& dict) const {
        auto it = mCreators.find(type);
        if (it == mCreators.end()) throw RuntimeError("no plugin class '" + type + "' registered");
        return it->second(std::move(d), dict);
    }
    const std::map<std::string, std::string>& getInfos() const { return mInfos; }
private:
    std::map<std::string, RenderPass::PluginCreate> mCreators;
    std::map<std::string, std::string> mInfos;
};

// Rendering/Lights/EmissiveLightSampler.h / EmissiveLightSamplerType.slangh
enum class EmissiveLightSamplerType : uint32_t { Uniform = 0, LightBVH = 1, Power = 2, Null = 0xff };

}  // namespace Falcor
