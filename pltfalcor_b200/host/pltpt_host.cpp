// pltpt_host.cpp — a minimal stand-in for the Mogwai side of scripts/PLTPathTracer.py: loads the PLTPT plugin library with
// dlopen, calls its registerPlugin(), creates the pass by type name from a property dictionary, wires the three RGBA32F
// outputs, plays GBufferRT's role for the camera jitter (Halton, 16 samples: GBufferBase.cpp:212-228) and AccumulatePass's
// role for the running sums (Double mode for colour, Single for the guides: scripts/PLTPathTracer.py:20-24).
//
//   pltpt_host --plugin libPLTPT.so --describe [--prop key=value ...]
//       prints one JSON object: plugin type/description, reflected channels, scripting dictionary, logged errors.  No GPU needed.
//   pltpt_host --plugin libPLTPT.so --scene scene.pltscn --width W --height H --frames N --out PREFIX [--prop key=value ...]
//       renders N frames and writes PREFIX.color.f64 (W*H*4 doubles: RGB sums + sample count), PREFIX.albedo.f32, PREFIX.normal.f32
//       (means) and a JSON line with the frame count and the pass's refresh flags.
// The scene file is the dump written by pltfalcor_b200.scene.SceneBuffers.save(): Falcor-layout buffers, byte for byte.
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "FalcorShim.h"

using namespace Falcor;

namespace {
struct SceneFile {
    std::vector<std::vector<char>> blobs;
    std::vector<PltptTextureDesc> textures;   // descriptor arrays the scene description points at (host memory, include/pltpt.h)
    std::vector<PltptSamplerDesc> samplers;
    PltptSceneDesc desc{};
    PltptCameraData camera{};
};
constexpr int kBlobCount = 14;

bool readExact(std::ifstream& f, void* dst, size_t n) { f.read(static_cast<char*>(dst), std::streamsize(n)); return bool(f); }

// layout: "PLTSCN1\0", uint64 scalars[14], float invWeightsSum + pad, PltptCameraData, then 14 x { uint64 bytes, data }
bool loadScene(const std::string& path, SceneFile& s) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    char magic[8];
    uint64_t sc[14];
    float fw[2];
    if (!readExact(f, magic, 8) || std::memcmp(magic, "PLTSCN1", 8) != 0) return false;
    if (!readExact(f, sc, sizeof sc) || !readExact(f, fw, sizeof fw) || !readExact(f, &s.camera, sizeof s.camera)) return false;
    s.blobs.resize(kBlobCount);
    for (auto& b : s.blobs) {
        uint64_t n = 0;
        if (!readExact(f, &n, 8)) return false;
        b.resize(n);
        if (n && !readExact(f, b.data(), n)) return false;
    }
    auto p = [&](int i) -> const void* { return s.blobs[i].empty() ? nullptr : s.blobs[i].data(); };
    PltptSceneDesc& d = s.desc;
    d.vertices = static_cast<const PltptPackedVertex*>(p(0)); d.vertexCount = uint32_t(sc[0]);
    d.indexData = p(1); d.indexDataBytes = sc[1];
    d.geometryInstances = static_cast<const PltptGeometryInstance*>(p(2)); d.instanceCount = uint32_t(sc[2]);
    d.meshes = static_cast<const PltptMeshDesc*>(p(3)); d.meshCount = uint32_t(sc[3]);
    d.worldMatrices = static_cast<const float*>(p(4)); d.inverseTransposeWorldMatrices = static_cast<const float*>(p(5)); d.matrixCount = uint32_t(sc[4]);
    d.materialData = static_cast<const PltptMaterialBlob*>(p(6)); d.materialCount = uint32_t(sc[5]);
    d.spectralProfiles = static_cast<const PltptSpectralProfile*>(p(7)); d.profileCount = uint32_t(sc[6]);
    d.lights = static_cast<const PltptLightData*>(p(8)); d.lightCount = uint32_t(sc[7]);
    d.emissiveTriangles = static_cast<const PltptEmissiveTriangle*>(p(9)); d.emissiveTriangleCount = uint32_t(sc[8]);
    d.emissiveFlux = static_cast<const PltptEmissiveFlux*>(p(10));
    d.activeTriangles = static_cast<const uint32_t*>(p(11)); d.activeTriangleCount = uint32_t(sc[9]);
    d.perMeshInstanceOffset = static_cast<const uint32_t*>(p(12));
    d.triangleAliasTable = static_cast<const uint32_t*>(p(13));
    d.invWeightsSum = fw[0];
    d.hitInfoTypeBits = uint32_t(sc[10]); d.hitInfoInstanceIDBits = uint32_t(sc[11]); d.hitInfoPrimitiveIndexBits = uint32_t(sc[12]);
    d.pointersAreDevice = 0;
    // optional trailing section: uint64 texel bytes, then width, height (uint32), EnvMapData (112 B), RGBA32F texels
    uint64_t envBytes = 0;
    if (readExact(f, &envBytes, 8) && envBytes) {
        uint32_t wh[2];
        s.blobs.emplace_back(envBytes);
        if (!readExact(f, wh, sizeof wh) || !readExact(f, &d.envMap, sizeof d.envMap) || !readExact(f, s.blobs.back().data(), envBytes)) return false;
        d.envMapTexels = reinterpret_cast<const float*>(s.blobs.back().data());
        d.envMapWidth = wh[0]; d.envMapHeight = wh[1];
    }
    // optional second trailing section: material textures and samplers (pltfalcor_b200/scene.py SceneBuffers.save)
    uint32_t counts[2] = {0, 0};
    if (readExact(f, counts, sizeof counts) && (counts[0] || counts[1])) {
        s.textures.resize(counts[0]);
        s.samplers.resize(counts[1]);
        s.blobs.reserve(s.blobs.size() + counts[0]);  // (the descriptors point into the blobs: no reallocation after this)
        for (auto& t : s.textures) {
            uint32_t hdr[4];
            uint64_t bytes = 0;
            if (!readExact(f, hdr, sizeof hdr) || !readExact(f, &bytes, 8)) return false;
            s.blobs.emplace_back(bytes);
            if (bytes && !readExact(f, s.blobs.back().data(), bytes)) return false;
            t = PltptTextureDesc{s.blobs.back().data(), hdr[0], hdr[1], hdr[2], 0};
        }
        for (auto& q : s.samplers) if (!readExact(f, &q, sizeof q)) return false;
        d.textures = s.textures.data(); d.textureCount = counts[0];
        d.samplers = s.samplers.data(); d.samplerCount = counts[1];
    }
    return true;
}

Dictionary::Value parseValue(const std::string& v) {
    if (v == "true" || v == "True") return true;
    if (v == "false" || v == "False") return false;
    char* end = nullptr;
    const long long i = std::strtoll(v.c_str(), &end, 10);
    if (end && *end == 0 && !v.empty()) return int64_t(i);
    const double d = std::strtod(v.c_str(), &end);
    if (end && *end == 0 && !v.empty()) return d;
    return v;
}

std::string jsonEscape(const std::string& s) {
    std::string o;
    for (char c : s) { if (c == '"' || c == '\\') o += '\\'; o += c; }
    return o;
}

float halton(uint32_t index, uint32_t base) {  // HaltonSamplePattern.cpp:37-67
    float result = 0.f, factor = 1.f;
    for (; index > 0; index /= base) { factor /= float(base); result += factor * float(index % base); }
    return result;
}

bool writeFile(const std::string& path, const void* data, size_t bytes) {
    std::ofstream f(path, std::ios::binary);
    f.write(static_cast<const char*>(data), std::streamsize(bytes));
    return bool(f);
}
}  // namespace

int main(int argc, char** argv) {
    std::string plugin, scenePath, movedScenePath, out = "pltpt_host_out";
    uint32_t moveAfter = 0;
    uint32_t width = 0, height = 0, frames = 1;
    int device = 0;
    bool describe = false;
    Dictionary props;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { std::fprintf(stderr, "missing value after %s\n", a.c_str()); std::exit(2); } return argv[++i]; };
        if (a == "--plugin") plugin = next();
        else if (a == "--scene") scenePath = next();
        else if (a == "--moved-scene") movedScenePath = next();   // same scene with other instance transforms: animation keyframe
        else if (a == "--move-after") moveAfter = uint32_t(std::stoul(next()));
        else if (a == "--width") width = uint32_t(std::stoul(next()));
        else if (a == "--height") height = uint32_t(std::stoul(next()));
        else if (a == "--frames") frames = uint32_t(std::stoul(next()));
        else if (a == "--device") device = std::stoi(next());
        else if (a == "--out") out = next();
        else if (a == "--describe") describe = true;
        else if (a == "--prop") {
            const std::string kv = next();
            const size_t eq = kv.find('=');
            if (eq == std::string::npos) { std::fprintf(stderr, "--prop expects key=value\n"); return 2; }
            props[kv.substr(0, eq)] = parseValue(kv.substr(eq + 1));
        } else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (plugin.empty()) { std::fprintf(stderr, "--plugin is required\n"); return 2; }

    // loadRenderPassLibrary("PLTPT") — PluginManager::loadPlugin (Core/Plugin.h:200-250)
    void* lib = dlopen(plugin.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!lib) { std::fprintf(stderr, "dlopen failed: %s\n", dlerror()); return 3; }
    using RegisterFn = void (*)(PluginRegistry&);
    auto reg = reinterpret_cast<RegisterFn>(dlsym(lib, "registerPlugin"));
    if (!reg) { std::fprintf(stderr, "plugin exports no registerPlugin\n"); return 3; }
    PluginRegistry registry;
    reg(registry);

    try {
        auto pDevice = std::make_shared<Device>(device);
        RenderPass::SharedPtr pass = registry.createClass("PLTPT", pDevice, props);  // createPass("PLTPT", {...})
        if (describe) {
            const RenderPassReflection r = pass->reflect(RenderPass::CompileData{});
            std::string j = "{\"type\": \"" + pass->getType() + "\", \"desc\": \"" + jsonEscape(pass->getDesc()) + "\", \"registered\": [";
            bool first = true;
            for (auto& kv : registry.getInfos()) { j += std::string(first ? "" : ", ") + "\"" + kv.first + "\""; first = false; }
            j += "], \"channels\": [";
            for (size_t i = 0; i < r.getFieldCount(); ++i) {
                const auto* f = r.getField(i);
                j += std::string(i ? ", " : "") + "{\"name\": \"" + f->name + "\", \"output\": " + (f->visibility == RenderPassReflection::Field::Visibility::Output ? "true" : "false") +
                     ", \"optional\": " + (f->optional ? "true" : "false") + ", \"format\": " + std::to_string(uint32_t(f->format)) + "}";
            }
            j += "], \"dictionary\": {";
            const Dictionary d = pass->getScriptingDictionary();
            first = true;
            for (const auto& [k, v] : d) {
                const std::string s = v.toString();
                j += std::string(first ? "" : ", ") + "\"" + k + "\": " + ((v.isBool() || v.isInt() || v.isFloat()) ? s : "\"" + jsonEscape(s) + "\"");
                first = false;
            }
            j += "}, \"logged_errors\": " + std::to_string(Logger::errorCount()) + ", \"last_message\": \"" + jsonEscape(Logger::lastMessage()) + "\"}";
            std::printf("%s\n", j.c_str());
            return 0;
        }

        if (scenePath.empty() || !width || !height) { std::fprintf(stderr, "--scene, --width and --height are required to render\n"); return 2; }
        SceneFile sf;
        if (!loadScene(scenePath, sf)) { std::fprintf(stderr, "cannot read scene file %s\n", scenePath.c_str()); return 3; }
        if (cudaSetDevice(device) != cudaSuccess) { std::fprintf(stderr, "no CUDA device: the PLTPT path has no CPU fallback\n"); return 4; }
        cudaStream_t stream;
        cudaStreamCreate(&stream);
        RenderContext ctx(stream);
        Dictionary graphDict;
        RenderData rd({width, height}, graphDict);
        auto color = Texture::create2D(width, height, ResourceFormat::RGBA32Float);
        auto albedo = Texture::create2D(width, height, ResourceFormat::RGBA32Float);
        auto normal = Texture::create2D(width, height, ResourceFormat::RGBA32Float);
        rd.setTexture("color", color); rd.setTexture("albedo", albedo); rd.setTexture("normal", normal);

        // before a scene is loaded the pass clears its outputs (PLTPT.cpp:374-384)
        cudaMemsetAsync(color->getCudaDevicePtr(), 0xff, size_t(width) * height * 16, stream);
        pass->execute(&ctx, rd);
        std::vector<float> probe(4);
        cudaMemcpyAsync(probe.data(), color->getCudaDevicePtr(), 16, cudaMemcpyDeviceToHost, stream);
        cudaStreamSynchronize(stream);
        const bool clearedWithoutScene = probe[0] == 0.f && probe[3] == 0.f;

        auto scene = std::make_shared<Scene>(sf.desc);
        scene->getCamera()->getData() = sf.camera;
        pass->setScene(&ctx, scene);  // m.loadScene -> RenderGraph::setScene

        const size_t texels = size_t(width) * height;
        std::vector<double> sum(texels * 4, 0.0);
        std::vector<float> guideA(texels * 4, 0.f), guideN(texels * 4, 0.f), host(texels * 4);
        SceneFile moved;
        if (!movedScenePath.empty() && !loadScene(movedScenePath, moved)) { std::fprintf(stderr, "cannot read scene file %s\n", movedScenePath.c_str()); return 3; }
        for (uint32_t f = 0; f < frames; ++f) {
            if (!movedScenePath.empty() && f == moveAfter) {
                // the animation controller moved instances: Falcor rewrites the matrix buffers in place and raises SceneGraphChanged
                // (Scene::update -> RenderGraph::onSceneUpdates); the accumulation restarts like AccumulatePass does on scene changes
                if (moved.blobs[4].size() != sf.blobs[4].size()) { std::fprintf(stderr, "--moved-scene has another matrix count\n"); return 3; }
                std::memcpy(sf.blobs[4].data(), moved.blobs[4].data(), sf.blobs[4].size());
                std::memcpy(sf.blobs[5].data(), moved.blobs[5].data(), sf.blobs[5].size());
                pass->onSceneUpdates(&ctx, Scene::UpdateFlags::SceneGraphChanged);
                std::fill(sum.begin(), sum.end(), 0.0);
            }
            // GBufferRT: Halton(2,3), 16 samples, offset by (0.5, 0.5) and wrapped into [-0.5, 0.5), in pixels
            const uint32_t k = f % 16u;
            float jx = halton(k, 2) + 0.5f, jy = halton(k, 3) + 0.5f;
            jx = (jx - std::floor(jx)) - 0.5f; jy = (jy - std::floor(jy)) - 0.5f;
            scene->getCamera()->setJitter(jx / float(width), jy / float(height));
            pass->execute(&ctx, rd);
            cudaMemcpyAsync(host.data(), color->getCudaDevicePtr(), texels * 16, cudaMemcpyDeviceToHost, stream);
            cudaStreamSynchronize(stream);
            for (size_t i = 0; i < texels * 4; ++i) sum[i] += double(host[i]);
            cudaMemcpy(host.data(), albedo->getCudaDevicePtr(), texels * 16, cudaMemcpyDeviceToHost);
            for (size_t i = 0; i < texels * 4; ++i) guideA[i] += host[i] / float(frames);
            cudaMemcpy(host.data(), normal->getCudaDevicePtr(), texels * 16, cudaMemcpyDeviceToHost);
            for (size_t i = 0; i < texels * 4; ++i) guideN[i] += host[i] / float(frames);
        }
        if (!writeFile(out + ".color.f64", sum.data(), sum.size() * 8) || !writeFile(out + ".albedo.f32", guideA.data(), guideA.size() * 4) ||
            !writeFile(out + ".normal.f32", guideN.data(), guideN.size() * 4)) { std::fprintf(stderr, "cannot write outputs\n"); return 5; }
        std::printf("{\"frames\": %u, \"width\": %u, \"height\": %u, \"cleared_without_scene\": %s, \"logged_errors\": %d}\n", frames, width, height,
                    clearedWithoutScene ? "true" : "false", Logger::errorCount());
        pass.reset();
        cudaStreamDestroy(stream);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
