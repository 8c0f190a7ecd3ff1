// ORACLE — TEST INFRASTRUCTURE ONLY (see plt_math.h header).
// plt_scene.h — Falcor scene-buffer access (vertex fetch, shading frame, hit packing) and the ray caster.
// Vertex unpacking, interpolated vertex data and the shading data at a hit (tangent space, normal maps, double-sided flip) are pinned to the
// reference's shader text (SceneTypes.slang, Scene.slang computeVertexData, ShadingUtils.slang, MaterialFactory.slang: oracle/ref_compare.cpp).
//
// Ray casting: the reference delegates to the D3D12/Vulkan driver + RT cores (Scene.cpp:3092-3760,
// TraceRay at pltpt_sample.rt.slang:174-183) — no source in the reference, so nothing to pin it to: the oracle DEFINES the
// intersection semantics the CUDA path must reproduce bit-for-bit (DESIGN.md "Intersection contract"):
//   * triangles are taken in world space: p_w = ((m0*x + m1*y) + m2*z) + m3 per row, each op rounded to fp32;
//   * watertight test of Woop/Benthin/Wald 2013 with the double-precision fallback on zero edge functions;
//   * a hit needs tmin < t < tmax (DXR triangle rule), no face culling;
//   * closest hit = minimum t; exact ties resolved towards the lower (instanceID, primitiveIndex);
//   * barycentrics (u,v) = (V/det, W/det), weights (1-u-v, u, v) as HitInfo.slang:76-79.
// A brute-force loop and a binary BVH (same leaf test; validated against brute force) are provided.
#pragma once
#include "plt_bsdf.h"
#include "plt_envmap.h"
#include <vector>
#include <numeric>

namespace orc {

struct TriangleHit {
    uint32_t instanceID = 0xffffffffu, primitiveIndex = 0;
    float2 barycentrics;
    float3 getBarycentricWeights() const { return float3(1.f - barycentrics.x - barycentrics.y, barycentrics.x, barycentrics.y); }
    bool valid() const { return instanceID != 0xffffffffu; }
};

struct StaticVertexData {
    float3 position, normal;
    float4 tangent;
    float2 texCrd;
};
struct VertexData {  // SceneTypes.slang:311-320
    float3 posW, normalW, faceNormalW;
    float4 tangentW;
    float2 texC;
};

struct WorldTri { float3 v0, v1, v2; };

struct RayHit { float t = 0; TriangleHit hit; };

struct Scene {
    PltptSceneDesc d{};
    ProfileTable profiles;
    TextureSystem textures;
    EnvMap env;
    // derived
    std::vector<uint32_t> instTriOffset;  // first global triangle of each instance
    std::vector<WorldTri> tris;           // world-space triangles, instance-major
    std::vector<uint32_t> triInst, triPrim;
    // BVH (binary, built over `tris`)
    struct Node { float lo[3], hi[3]; uint32_t left, right, first, count; };
    std::vector<Node> nodes;
    std::vector<uint32_t> order;

    void init(const PltptSceneDesc& desc) {
        d = desc;
        profiles.profiles = d.spectralProfiles;
        profiles.count = d.profileCount;
        textures.textures = d.textures; textures.textureCount = d.textureCount;
        textures.samplers = d.samplers; textures.samplerCount = d.samplerCount;
        instTriOffset.assign(d.instanceCount + 1, 0);
        for (uint32_t i = 0; i < d.instanceCount; ++i) {
            const PltptMeshDesc& m = d.meshes[d.geometryInstances[i].geometryID];
            const uint32_t n = (m.indexCount > 0 ? m.indexCount : m.vertexCount) / 3;
            instTriOffset[i + 1] = instTriOffset[i] + n;
        }
        tris.resize(instTriOffset[d.instanceCount]);
        triInst.resize(tris.size());
        triPrim.resize(tris.size());
        for (uint32_t i = 0; i < d.instanceCount; ++i) {
            const float* M = d.worldMatrices + 16 * size_t(d.geometryInstances[i].globalMatrixID);
            for (uint32_t p = 0; p < instTriOffset[i + 1] - instTriOffset[i]; ++p) {
                uint32_t idx[3];
                getIndices(i, p, idx);
                WorldTri w;
                float3* dst[3] = {&w.v0, &w.v1, &w.v2};
                for (int k = 0; k < 3; ++k) {
                    const float* P = d.vertices[idx[k]].position;
                    for (int r = 0; r < 3; ++r) (*dst[k])[r] = ((M[4 * r + 0] * P[0] + M[4 * r + 1] * P[1]) + M[4 * r + 2] * P[2]) + M[4 * r + 3];
                }
                const uint32_t g = instTriOffset[i] + p;
                tris[g] = w; triInst[g] = i; triPrim[g] = p;
            }
        }
        buildBVH();
        env.init(d);
    }

    // Scene.slang getLocalIndices / getIndices
    void getIndices(uint32_t instanceID, uint32_t triangleIndex, uint32_t out[3]) const {
        const PltptGeometryInstance& inst = d.geometryInstances[instanceID];
        const PltptMeshDesc& mesh = d.meshes[inst.geometryID];
        if (mesh.indexCount == 0) {
            for (int k = 0; k < 3; ++k) out[k] = triangleIndex * 3 + k + inst.vbOffset;
            return;
        }
        const uint8_t* base = (const uint8_t*)d.indexData + size_t(inst.ibOffset) * 4;
        if (inst.flags & PLTPT_INSTANCE_FLAG_USE_16BIT_INDICES) {
            const uint16_t* p = (const uint16_t*)(base + size_t(triangleIndex) * 6);
            for (int k = 0; k < 3; ++k) out[k] = uint32_t(p[k]) + inst.vbOffset;
        } else {
            const uint32_t* p = (const uint32_t*)(base + size_t(triangleIndex) * 12);
            for (int k = 0; k < 3; ++k) out[k] = p[k] + inst.vbOffset;
        }
    }
    // SceneTypes.slang:282-296
    StaticVertexData getVertex(uint32_t index) const {
        const PltptPackedVertex& pv = d.vertices[index];
        StaticVertexData v;
        v.position = float3(pv.position[0], pv.position[1], pv.position[2]);
        v.texCrd = float2(pv.texCrd[0], pv.texCrd[1]);
        const uint32_t* pk = pv.packedNormalTangentCurveRadius;
        v.normal = normalize(float3(f16tof32(pk[0] & 0xffff), f16tof32(pk[0] >> 16), f16tof32(pk[1] & 0xffff)));
        float3 t = decodeNormal2x16(pk[2]);
        v.tangent = float4(t, signf(f16tof32(pk[1] >> 16)));
        return v;
    }
    static float3 mulPoint(const float* M, float3 p) {
        float3 r;
        for (int i = 0; i < 3; ++i) r[i] = M[4 * i] * p.x + M[4 * i + 1] * p.y + M[4 * i + 2] * p.z + M[4 * i + 3];
        return r;
    }
    static float3 mulVec(const float* M, float3 p) {
        float3 r;
        for (int i = 0; i < 3; ++i) r[i] = M[4 * i] * p.x + M[4 * i + 1] * p.y + M[4 * i + 2] * p.z;
        return r;
    }
    // Scene.slang computeVertexData / getVertexData (triangle meshes)
    VertexData getVertexData(const TriangleHit& hit, StaticVertexData* outVerts = nullptr) const {
        uint32_t idx[3];
        getIndices(hit.instanceID, hit.primitiveIndex, idx);
        StaticVertexData vs[3] = {getVertex(idx[0]), getVertex(idx[1]), getVertex(idx[2])};
        if (outVerts) for (int k = 0; k < 3; ++k) outVerts[k] = vs[k];
        const PltptGeometryInstance& inst = d.geometryInstances[hit.instanceID];
        const float* worldMat = d.worldMatrices + 16 * size_t(inst.globalMatrixID);
        const float* worldIT = d.inverseTransposeWorldMatrices + 16 * size_t(inst.globalMatrixID);
        const bool isFrontFaceCW = (inst.flags & PLTPT_INSTANCE_FLAG_OBJECT_FRONT_FACE_CW) != 0;
        const float3 b = hit.getBarycentricWeights();
        VertexData v;
        v.posW = vs[0].position * b[0]; v.posW += vs[1].position * b[1]; v.posW += vs[2].position * b[2];
        v.normalW = vs[0].normal * b[0]; v.normalW += vs[1].normal * b[1]; v.normalW += vs[2].normal * b[2];
        v.tangentW = vs[0].tangent * b[0]; v.tangentW += vs[1].tangent * b[1]; v.tangentW += vs[2].tangent * b[2];
        v.tangentW.w = signf(v.tangentW.w);
        v.texC = vs[0].texCrd * b[0]; v.texC += vs[1].texCrd * b[1]; v.texC += vs[2].texCrd * b[2];
        float3 N = normalize(cross(vs[1].position - vs[0].position, vs[2].position - vs[0].position));
        v.faceNormalW = isFrontFaceCW ? -N : N;
        v.posW = mulPoint(worldMat, v.posW);
        v.normalW = mulVec(worldIT, v.normalW);
        v.faceNormalW = mulVec(worldIT, v.faceNormalW);
        float3 t = mulVec(worldMat, v.tangentW.xyz());
        v.normalW = normalize(v.normalW);
        v.faceNormalW = normalize(v.faceNormalW);
        t = normalize(t);
        v.tangentW = float4(t, v.tangentW.w);
        return v;
    }
    uint32_t getMaterialID(uint32_t instanceID) const { return d.geometryInstances[instanceID].materialID; }

    // MaterialFactory.slang:79-131 (prepareShadingData) + ShadingUtils.slang:126-152 (orthonormalizeSafe), :199-246 (normal mapping)
    ShadingData prepareShadingData(const VertexData& v, uint32_t materialID, float3 viewDir) const {
        ShadingData sd;
        sd.posW = v.posW;
        sd.uv = v.texC;
        sd.V = viewDir;
        sd.N = v.normalW;
        const float3 tw = v.tangentW.xyz();
        const float NdotT = dot(tw, sd.N);
        const bool nonParallel = std::fabs(NdotT) < 0.9999f;
        const bool nonZero = dot(tw, tw) > 0.f;
        const bool valid = std::fabs(v.tangentW.w) == 1.0f && nonZero && nonParallel;
        if (valid) {
            sd.T = normalize(tw - sd.N * NdotT);
            sd.B = cross(sd.N, sd.T) * v.tangentW.w;
        } else {
            sd.T = perp_stark(sd.N);
            sd.B = cross(sd.N, sd.T);
        }
        sd.tangentW = valid ? v.tangentW : float4(sd.T, 1.f);
        sd.faceN = v.faceNormalW;
        sd.frontFacing = dot(sd.V, sd.faceN) >= 0.f;
        const PltptMaterialBlob& m = d.materialData[materialID];
        sd.mtlHeader[0] = m.header[0];
        sd.mtlHeader[1] = m.header[1];
        sd.materialID = materialID;
        sd.opacity = MaterialInstance::evalOpacity(m, textures, v.texC);
        if (valid) {  // material.modifyTangentSpace (every PLT material: sample texNormalMap, applyNormalMap), only with a valid tangent space
            const float3 encodedNormal = textures.sampleTexture(m.texNormalMap, TextureSystem::samplerIDOf(m), sd.uv, float4(0.f)).xyz();
            applyNormalMap(sd, TextureSystem::normalMapTypeOf(m), encodedNormal);
        }
        if (!sd.frontFacing && sd.isDoubleSided()) sd.N = -sd.N;
        return sd;
    }
    // ShadingUtils.slang:199-246: applyNormalMap / computeShadingFrameFromNormalMap, ShadingFrame(N, tangentW) -> orthonormalizeSafe (:126-152)
    static void applyNormalMap(ShadingData& sd, uint32_t type, float3 encodedNormal) {
        float3 mapN;
        if (type == 1) mapN = encodedNormal * 2.f - float3(1.f);                 // NormalMapType::RGB, rgbToNormal
        else if (type == 2) {                                                    // NormalMapType::RG, rgToNormal
            const float2 rg(encodedNormal.x, encodedNormal.y);
            mapN = float3(rg.x * 2.f - 1.f, rg.y * 2.f - 1.f, 0.f);
            mapN.z = std::sqrt(1.f - hlsl_min(hlsl_max(dot(rg, rg), 0.f), 1.f));
        } else return;                                                           // NormalMapType::None: ShadingFrame(sd) unchanged
        const float3 n = sd.T * mapN.x + sd.B * mapN.y + sd.N * mapN.z;
        // ShadingFrame(n, sd.tangentW) (ShadingUtils.slang:61-65): N = normalize(n), then the UNCHECKED orthonormalize (:112-116) --
        // a perturbed normal parallel to the tangent is the asset's problem there (:236-238) and stays so here
        sd.N = normalize(n);
        const float3 tw = sd.tangentW.xyz();
        sd.T = normalize(tw - sd.N * dot(tw, sd.N));
        sd.B = cross(sd.N, sd.T) * sd.tangentW.w;
    }
    MaterialInstance getMaterialInstance(const ShadingData& sd) const { return MaterialInstance(d.materialData[sd.materialID], sd, &profiles, textures); }

    // HitInfo.slang:459-487 / 96-130 (uncompressed 128-bit format, Triangle type = 1)
    void packHit(const TriangleHit& h, uint32_t out[4]) const {
        const uint32_t typeOffset = 32u - d.hitInfoTypeBits;
        const uint32_t headerBits = d.hitInfoTypeBits + d.hitInfoInstanceIDBits + d.hitInfoPrimitiveIndexBits;
        out[0] = out[1] = out[2] = out[3] = 0;
        if (headerBits <= 32) out[0] = (1u << typeOffset) | (h.instanceID << d.hitInfoPrimitiveIndexBits) | h.primitiveIndex;
        else { out[0] = (1u << typeOffset) | h.instanceID; out[1] = h.primitiveIndex; }
        out[2] = asuint(h.barycentrics.x);
        out[3] = asuint(h.barycentrics.y);
    }
    TriangleHit unpackHit(const uint32_t in[4]) const {
        TriangleHit h;
        const uint32_t headerBits = d.hitInfoTypeBits + d.hitInfoInstanceIDBits + d.hitInfoPrimitiveIndexBits;
        if (headerBits <= 32) {
            h.instanceID = (in[0] >> d.hitInfoPrimitiveIndexBits) & ((1u << d.hitInfoInstanceIDBits) - 1);
            h.primitiveIndex = in[0] & ((1u << d.hitInfoPrimitiveIndexBits) - 1);
        } else {
            h.instanceID = in[0] & ((1u << d.hitInfoInstanceIDBits) - 1);
            h.primitiveIndex = in[1];
        }
        h.barycentrics = float2(asfloat(in[2]), asfloat(in[3]));
        return h;
    }
    // LightCollection.slang:97-101
    uint32_t getEmissiveTriangleIndex(uint32_t instanceID, uint32_t primitiveIndex) const {
        if (!d.perMeshInstanceOffset) return 0xffffffffu;
        uint32_t offset = d.perMeshInstanceOffset[instanceID];
        return offset != 0xffffffffu ? offset + primitiveIndex : 0xffffffffu;
    }

    // ------------------------------------------------------------------ ray casting -----
    struct RayPrep {
        int kx, ky, kz;
        float Sx, Sy, Sz;
        float3 org;
    };
    static RayPrep prepRay(float3 org, float3 dir) {
        RayPrep r;
        r.org = org;
        const float ax = std::fabs(dir.x), ay = std::fabs(dir.y), az = std::fabs(dir.z);
        r.kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
        r.kx = r.kz + 1; if (r.kx == 3) r.kx = 0;
        r.ky = r.kx + 1; if (r.ky == 3) r.ky = 0;
        if (dir[r.kz] < 0.0f) std::swap(r.kx, r.ky);
        r.Sx = dir[r.kx] / dir[r.kz];
        r.Sy = dir[r.ky] / dir[r.kz];
        r.Sz = 1.0f / dir[r.kz];
        return r;
    }
    // returns true and (t,u,v) if the triangle is hit with tmin < t < tmax
    static bool intersectTri(const RayPrep& r, const WorldTri& tri, float tmin, float tmax, float& t, float& u, float& v) {
        const float3 A = tri.v0 - r.org, B = tri.v1 - r.org, C = tri.v2 - r.org;
        const float Ax = A[r.kx] - r.Sx * A[r.kz], Ay = A[r.ky] - r.Sy * A[r.kz];
        const float Bx = B[r.kx] - r.Sx * B[r.kz], By = B[r.ky] - r.Sy * B[r.kz];
        const float Cx = C[r.kx] - r.Sx * C[r.kz], Cy = C[r.ky] - r.Sy * C[r.kz];
        float U = Cx * By - Cy * Bx;
        float V = Ax * Cy - Ay * Cx;
        float W = Bx * Ay - By * Ax;
        if (U == 0.0f || V == 0.0f || W == 0.0f) {
            U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
            V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
            W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
        }
        if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
        const float det = U + V + W;
        if (det == 0.0f) return false;
        const float Az = r.Sz * A[r.kz], Bz = r.Sz * B[r.kz], Cz = r.Sz * C[r.kz];
        const float T = U * Az + V * Bz + W * Cz;
        const float rcpDet = 1.0f / det;
        t = T * rcpDet;
        if (!(t > tmin && t < tmax)) return false;
        u = V * rcpDet;
        v = W * rcpDet;
        return true;
    }
    void considerTri(uint32_t g, const RayPrep& r, float tmin, float tmax, RayHit& best) const {
        float t, u, v;
        if (!intersectTri(r, tris[g], tmin, tmax, t, u, v)) return;
        if (best.hit.valid()) {
            if (t > best.t) return;
            if (t == best.t) {
                const uint32_t gb = instTriOffset[best.hit.instanceID] + best.hit.primitiveIndex;
                if (g > gb) return;
            }
        }
        best.t = t;
        best.hit.instanceID = triInst[g];
        best.hit.primitiveIndex = triPrim[g];
        best.hit.barycentrics = float2(u, v);
    }
    RayHit traceClosestBrute(float3 org, float3 dir, float tmin, float tmax) const {
        RayHit best;
        const RayPrep r = prepRay(org, dir);
        for (uint32_t g = 0; g < tris.size(); ++g) considerTri(g, r, tmin, tmax, best);
        return best;
    }
    bool traceAnyBrute(float3 org, float3 dir, float tmin, float tmax) const {
        const RayPrep r = prepRay(org, dir);
        float t, u, v;
        for (uint32_t g = 0; g < tris.size(); ++g) if (intersectTri(r, tris[g], tmin, tmax, t, u, v)) return true;
        return false;
    }

    void buildBVH() {
        const uint32_t n = uint32_t(tris.size());
        order.resize(n);
        std::iota(order.begin(), order.end(), 0u);
        nodes.clear();
        if (n == 0) return;
        nodes.reserve(2 * n);
        std::vector<float3> cent(n), lo(n), hi(n);
        for (uint32_t i = 0; i < n; ++i) {
            lo[i] = vmin(tris[i].v0, vmin(tris[i].v1, tris[i].v2));
            hi[i] = vmax(tris[i].v0, vmax(tris[i].v1, tris[i].v2));
            cent[i] = (lo[i] + hi[i]) * 0.5f;
        }
        struct Job { uint32_t node, first, count; };
        std::vector<Job> stack;
        nodes.push_back(Node{});
        stack.push_back({0, 0, n});
        while (!stack.empty()) {
            Job j = stack.back();
            stack.pop_back();
            float3 blo(FLT_MAXf), bhi(-FLT_MAXf), clo(FLT_MAXf), chi(-FLT_MAXf);
            for (uint32_t i = j.first; i < j.first + j.count; ++i) {
                const uint32_t g = order[i];
                blo = vmin(blo, lo[g]); bhi = vmax(bhi, hi[g]);
                clo = vmin(clo, cent[g]); chi = vmax(chi, cent[g]);
            }
            Node nd{};
            for (int k = 0; k < 3; ++k) { nd.lo[k] = blo[k]; nd.hi[k] = bhi[k]; }
            nd.first = j.first; nd.count = j.count; nd.left = nd.right = 0;
            const float3 ext = chi - clo;
            const int axis = ext.x > ext.y ? (ext.x > ext.z ? 0 : 2) : (ext.y > ext.z ? 1 : 2);
            if (j.count <= 4 || !(ext[axis] > 0.f)) { nodes[j.node] = nd; continue; }
            const uint32_t mid = j.first + j.count / 2;
            std::nth_element(order.begin() + j.first, order.begin() + mid, order.begin() + j.first + j.count,
                             [&](uint32_t a, uint32_t b) { return cent[a][axis] < cent[b][axis]; });
            nd.count = 0;
            nd.left = uint32_t(nodes.size());
            nodes.push_back(Node{});
            nd.right = uint32_t(nodes.size());
            nodes.push_back(Node{});
            nodes[j.node] = nd;
            stack.push_back({nd.left, j.first, mid - j.first});
            stack.push_back({nd.right, mid, j.first + j.count - mid});
        }
    }
    // conservative slab test (double precision, padded) — only has to never reject a box containing a hit
    static bool hitBox(const Node& nd, const float3& org, const double inv[3], double tmin, double tmax) {
        double t0 = tmin, t1 = tmax;
        for (int k = 0; k < 3; ++k) {
            double a = (double(nd.lo[k]) - double(org[k])) * inv[k], b = (double(nd.hi[k]) - double(org[k])) * inv[k];
            if (std::isnan(a) || std::isnan(b)) continue;  // 0 * inf: ray lies in the slab plane
            if (a > b) std::swap(a, b);
            a -= 1e-5 * std::fabs(a) + 1e-9;
            b += 1e-5 * std::fabs(b) + 1e-9;
            t0 = hlsl_max(t0, a); t1 = hlsl_min(t1, b);
        }
        return t0 <= t1;
    }
    template <bool ANY>
    bool traverse(float3 org, float3 dir, float tmin, float tmax, RayHit& best) const {
        if (nodes.empty()) return false;
        const RayPrep r = prepRay(org, dir);
        const double inv[3] = {1.0 / double(dir.x), 1.0 / double(dir.y), 1.0 / double(dir.z)};
        uint32_t stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            const Node& nd = nodes[stack[--sp]];
            // widen by the ray-extent scale so equal-t ties are still visited
            const double far = best.hit.valid() ? double(best.t) * (1.0 + 1e-5) + 1e-9 : double(tmax);
            if (!hitBox(nd, org, inv, double(tmin), far)) continue;
            if (nd.count > 0) {
                for (uint32_t i = nd.first; i < nd.first + nd.count; ++i) {
                    if (ANY) {
                        float t, u, v;
                        if (intersectTri(r, tris[order[i]], tmin, tmax, t, u, v)) return true;
                    } else considerTri(order[i], r, tmin, tmax, best);
                }
            } else {
                stack[sp++] = nd.left;
                stack[sp++] = nd.right;
            }
        }
        return best.hit.valid();
    }
    RayHit traceClosest(float3 org, float3 dir, float tmin, float tmax) const {
        RayHit best;
        traverse<false>(org, dir, tmin, tmax, best);
        return best;
    }
    bool traceAny(float3 org, float3 dir, float tmin, float tmax) const {
        RayHit best;
        return traverse<true>(org, dir, tmin, tmax, best);
    }
};

}  // namespace orc
