"""Procedural scene harness: fabricates Falcor-layout scene buffers (the `gScene` inputs of the PLTPT pass).

The reference gets these buffers from Falcor's SceneBuilder / MaterialSystem / LightCollection; neither Falcor nor
any scene asset is available offline, so the named benchmark scenes are generated here in exactly the same byte
layouts (include/pltpt.h).  Host-side restatements (each cites the reference):
  * .spd loader and SpectralProfile builder      — Spectrum.cpp:89-112, SampledSpectrum.h:57-68, SceneBuilder.cpp:911-960
  * PLT material parameter packing (fp16 slots)  — Rendering/Materials/PLT/*Material.{cpp,slang}
  * LightCollection triangle list + flux         — LightCollectionShared.slang:76-135, FinalizeIntegration.cs.slang:64-110
  * EmissivePowerSampler alias table             — EmissivePowerSampler.cpp:86-170
  * Camera vectors and Halton jitter             — Camera.cpp:111-196, HaltonSamplePattern.cpp:37-67, GBufferBase.cpp:212-228
"""
import ctypes as C
import json
import math
import os

import numpy as np

from . import layouts as L

f32 = np.float32
_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "spd_tables.json")  # input data of the harness (measured tables), not a test fixture


# ----------------------------------------------------------------------------------------------- spectral tables
def load_spd(path):
    """PiecewiseLinearSpectrum::fromFile (Spectrum.cpp:89-112): `lambda value` rows, '#' comments, stop at first bad row."""
    w, v = [], []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line[0] == "#":
                continue
            parts = line.split()
            try:
                a, b = float(parts[0]), float(parts[1])
            except (ValueError, IndexError):
                break
            w.append(a)
            v.append(b)
    if not w:
        raise RuntimeError(f"Path could not be read '{path}'.")
    return np.asarray(w, f32), np.asarray(v, f32)


_tables_cache = None


def bundled_tables():
    """The small set of measured IOR / emission tables bundled with the harness (pltfalcor_b200/data, written by tools/make_golden_tables.py)."""
    global _tables_cache
    if _tables_cache is None:
        with open(_GOLDEN) as f:
            _tables_cache = json.load(f)
    return _tables_cache


def _pls_eval(w, v, lam):
    """PiecewiseLinearSpectrum::eval (Spectrum.h:88-106), fp32."""
    lam = f32(lam)
    if len(w) == 0 or lam < w[0] or lam > w[-1]:
        return f32(0)
    it = int(np.searchsorted(w, lam, side="left"))  # std::lower_bound
    if it == 0:
        return f32(v[0])
    i = it - 1
    t = f32(lam - w[i]) / f32(w[i + 1] - w[i])
    return f32(v[i] + f32(v[i + 1] - v[i]) * t)


def sampled_spectrum_from_pls(w, v, bin_min=5.0):
    """SampledSpectrum(PiecewiseLinearSpectrum) (SampledSpectrum.h:57-68) -> (lambdaStart, lambdaEnd, samples)."""
    w = np.asarray(w, f32)
    v = np.asarray(v, f32)
    start, end = f32(w[0]), f32(w[-1])
    least = f32(end - start)
    for i in range(1, len(w)):
        least = min(least, f32(w[i] - w[i - 1]))
    bin_size = max(f32(bin_min), f32(least / f32(2)))
    n = max(1, int(f32(f32(end - start) / bin_size) + f32(0.5)))
    samples = np.zeros(n, f32)
    for i in range(n):
        lam = f32(start + f32(f32(end - start) / f32(n - 1)) * f32(i)) if n > 1 else start
        samples[i] = _pls_eval(w, v, lam)
    return start, end, samples


def _ss_eval(start, end, samples, lam):
    """SampledSpectrum::eval (SampledSpectrum.h:303-317), fp32."""
    lam = f32(lam)
    if lam < start or lam > end:
        return f32(0)
    x = f32(f32(lam - start) / f32(end - start)) * f32(len(samples) - 1.0)
    i = int(math.floor(x))
    if i + 1 >= len(samples):
        return f32(samples[-1])
    wgt = f32(x - f32(i))
    return f32(samples[i] + f32(samples[i + 1] - samples[i]) * wgt)


def _wyman_xyz(lam):
    def G(x, mu, s1, s2):
        t = (x - mu) * (s1 if x < mu else s2)
        return math.exp(-0.5 * t * t)
    x = 0.362 * G(lam, 442.0, 0.0624, 0.0374) + 1.056 * G(lam, 599.8, 0.0264, 0.0323) - 0.065 * G(lam, 501.1, 0.0490, 0.0382)
    y = 0.821 * G(lam, 568.8, 0.0213, 0.0247) + 0.286 * G(lam, 530.9, 0.0613, 0.0322)
    z = 1.217 * G(lam, 437.0, 0.0845, 0.0278) + 0.681 * G(lam, 459.0, 0.0385, 0.0725)
    return np.array([x, y, z])


_D65 = [49.9755, 52.3118, 54.6482, 68.7015, 82.7549, 87.1204, 91.486, 92.4589, 93.4318, 90.057, 86.6823, 95.7736, 104.865, 110.936,
        117.008, 117.41, 117.812, 116.336, 114.861, 115.392, 115.923, 112.367, 108.811, 109.082, 109.354, 108.578, 107.802, 106.296,
        104.79, 106.239, 107.689, 106.047, 104.405, 104.225, 104.046, 102.023, 100., 98.1671, 96.3342, 96.0611, 95.788, 92.2368, 88.6856,
        89.3459, 90.0062, 89.8026, 89.5991, 88.6489, 87.6987, 85.4936, 83.2886, 83.4939, 83.6992, 81.863, 80.0268, 80.1207, 80.2146,
        81.2462, 82.2778, 80.281, 78.2842, 74.0027, 69.7213, 70.6652, 71.6091, 72.979, 74.349, 67.9765, 61.604]
_XYZ2RGB = np.array([[3.240969941904522, -1.537383177570094, -0.4986107602930032],
                     [-0.9692436362808803, 1.875967501507721, 0.04155505740717569],
                     [0.05563007969699373, -0.2039769588889765, 1.056971514242878]])


def _profile_rgb(start, end, samples):
    """SpectrumUtils::toRGB_D65 (SpectrumUtils.h:78-151): trapezoid sum of XYZ(λ)·D65(λ)·s(λ) -> Rec.709 / Y_D65.

    DEVIATION: the reference integrates against the CIE-1931 1 nm tables (external/include/xyzcurves); this harness
    uses the Wyman analytic fit the device code uses (SpectrumUtils.slang:43-49).  `rgb` only feeds the per-triangle
    flux (light-selection pdf), which the estimator divides out again, so images are unaffected in expectation.
    """
    n = len(samples)
    if n < 2:
        return np.zeros(3, f32)
    delta = (float(end) - float(start)) / (n - 1.0)
    acc = np.zeros(3)
    for q in range(n):
        lam = min(float(start) + delta * q, float(end))
        if lam < 380.0 or lam > 720.0:
            continue
        d65 = np.interp((lam - 380.0) / 5.0, np.arange(69), _D65)
        acc += _wyman_xyz(lam) * d65 * float(samples[q]) * delta * (0.5 if q in (0, n - 1) else 1.0)
    return (_XYZ2RGB @ acc / 10567.0762).astype(f32)


def gen_spectral_profile(wavelengths, values, compute_icdf=True, scale=1.0):
    """genSpectralProfile (SceneBuilder.cpp:911-960) for a `.spd` table; returns a spectral_profile_dtype record."""
    v = (np.asarray(values, f32) * f32(scale)).astype(f32)
    start, end, samples = sampled_spectrum_from_pls(wavelengths, v)
    sp = np.zeros((), L.spectral_profile_dtype)
    sp["rgb"] = _profile_rgb(start, end, samples)
    sp["intensity"] = 1.0
    mn, mx = max(start, f32(380.0)), min(end, f32(720.0))
    sp["minWavelength"], sp["maxWavelength"] = mn, mx
    bins = min(128, int(f32(f32(len(samples)) * f32(mx - mn) / f32(end - start)) + f32(0.5)))
    sp["bins"] = bins
    if bins == 0:
        return sp
    r = f32(1.0) / f32(bins - 1) if bins > 1 else f32(1.0)
    pdf = np.zeros(128, f32)
    total = f32(0)
    for i in range(bins):
        lam = f32(mn + f32(mx - mn) * f32(f32(i) * r))
        pdf[i] = _ss_eval(start, end, samples, lam)
        total = f32(total + pdf[i])
    if total < f32(1.192092896e-07) or not compute_icdf:
        sp["pdf"] = pdf
        return sp
    for i in range(bins):
        pdf[i] = f32(pdf[i] / total)
    sp["intensity"] = total
    icdf = np.zeros(128, f32)
    cdfa, cdfb, pdfi = f32(0), f32(pdf[0]), 0
    for icdfi in range(bins):
        cdf = f32(f32(icdfi) * r)
        while cdf >= cdfb and pdfi + 1 < bins:
            cdfa = cdfb
            pdfi += 1
            cdfb = f32(cdfb + pdf[pdfi])
        l = min(f32(1.0), f32(f32(cdf - cdfa) / f32(cdfb - cdfa)))
        a = f32(max(1, pdfi)) - f32(1.0)
        icdf[icdfi] = f32(f32(a + f32(f32(pdfi) - a) * l) * r)
    sp["pdf"], sp["icdf"] = pdf, icdf
    return sp


# ----------------------------------------------------------------------------------------------- packing helpers
def _half_bits(x):
    return np.asarray(x, np.float16).view(np.uint16)


def _oct_snorm(n):
    n = np.asarray(n, np.float64)
    p = n[:2] / (abs(n[0]) + abs(n[1]) + abs(n[2]))
    if n[2] < 0:
        p = np.array([(1 - abs(p[1])) * (1 if p[0] >= 0 else -1), (1 - abs(p[0])) * (1 if p[1] >= 0 else -1)])
    return p


def _pack_snorm16(v):
    v = min(max(float(v), -1.0), 1.0)
    return int(math.trunc(v * 32767.0 + (0.5 if v >= 0 else -0.5))) & 0xFFFF


def encode_normal_2x16(n):
    """encodeNormal2x16 (Utils/Math/PackedFormats.slang:50-54)."""
    o = _oct_snorm(n)
    return _pack_snorm16(o[0]) | (_pack_snorm16(o[1]) << 16)


def pack_vertices(positions, normals, uvs=None, tangents=None):
    """PackedStaticVertexData::pack (SceneTypes.slang:252-266). tangents: optional (n,4) with w = handedness."""
    n = len(positions)
    out = np.zeros(n, L.packed_vertex_dtype)
    out["position"] = np.asarray(positions, f32)
    if uvs is not None:
        out["texCrd"] = np.asarray(uvs, f32)
    nh = _half_bits(np.asarray(normals, f32)).astype(np.uint32)
    tw = np.zeros(n, f32) if tangents is None else np.asarray(tangents, f32)[:, 3]
    twh = _half_bits(tw).astype(np.uint32)
    out["packed"][:, 0] = nh[:, 0] | (nh[:, 1] << 16)
    out["packed"][:, 1] = nh[:, 2] | (twh << 16)
    if tangents is not None:
        out["packed"][:, 2] = [encode_normal_2x16(t[:3]) for t in np.asarray(tangents, f32)]
    else:
        out["packed"][:, 2] = encode_normal_2x16([1.0, 0.0, 0.0])
    return out


def _header(mtype, double_sided=False, emissive=False):
    x = (mtype & 0xFFFF) | (0xFF << 20) | (int(double_sided) << 28) | (int(emissive) << 30) | (1 << 31)
    y = int(_half_bits(0.5))  # alpha threshold 0.5, AlphaMode::Opaque (Material.cpp:63-65)
    return x, y


def _blank_material(mtype, double_sided=False):
    m = np.zeros((), L.material_blob_dtype)
    m["header"] = _header(mtype, double_sided)
    m["baseColor"] = [1, 1, 1, 1]
    m["transmission"] = [1, 1, 1]
    m["IoR"] = 1.5
    m["emissionSpectralId"] = m["iorNSpectralId"] = m["iorKSpectralId"] = L.NO_SPECTRAL_PROFILE
    return m


def _set_data1(m, vals=(0, 0, 0, 0), w_bits=None):
    d = _half_bits(np.asarray(vals, f32)).copy()
    if w_bits is not None:
        d[3] = w_bits
    m["data1"] = d


# ----------------------------------------------------------------------------------------------- meshes
def quad(p0, p1, p2, p3):
    """Two triangles (p0,p1,p2),(p0,p2,p3); normal = normalize((p1-p0) x (p3-p0))."""
    P = np.asarray([p0, p1, p2, p3], f32)
    n = np.cross(P[1] - P[0], P[3] - P[0])
    n = n / np.linalg.norm(n)
    return P, np.tile(n, (4, 1)).astype(f32), np.asarray([[0, 0], [1, 0], [1, 1], [0, 1]], f32), np.asarray([0, 1, 2, 0, 2, 3], np.uint32)


def box(center, half, rot_y=0.0):
    c, s = math.cos(rot_y), math.sin(rot_y)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    hx, hy, hz = half
    faces = [  # outward-facing quads
        ([-hx, -hy, hz], [hx, -hy, hz], [hx, hy, hz], [-hx, hy, hz]),
        ([hx, -hy, -hz], [-hx, -hy, -hz], [-hx, hy, -hz], [hx, hy, -hz]),
        ([hx, -hy, hz], [hx, -hy, -hz], [hx, hy, -hz], [hx, hy, hz]),
        ([-hx, -hy, -hz], [-hx, -hy, hz], [-hx, hy, hz], [-hx, hy, -hz]),
        ([-hx, hy, hz], [hx, hy, hz], [hx, hy, -hz], [-hx, hy, -hz]),
        ([-hx, -hy, -hz], [hx, -hy, -hz], [hx, -hy, hz], [-hx, -hy, hz]),
    ]
    Ps, Ns, UVs, Is = [], [], [], []
    for f in faces:
        pts = [R @ np.asarray(p, np.float64) + np.asarray(center, np.float64) for p in f]
        P, N, UV, I = quad(*pts)
        Is.append(I + 4 * len(Ps))
        Ps.append(P)
        Ns.append(N)
        UVs.append(UV)
    return np.concatenate(Ps), np.concatenate(Ns), np.concatenate(UVs), np.concatenate(Is).astype(np.uint32)


def uv_sphere(center, radius, n_lat=64, n_lon=128):
    """Tessellated sphere with smooth normals (config 3 needs >= 20k triangles for MNEE)."""
    th = np.linspace(0.0, math.pi, n_lat + 1)
    ph = np.linspace(0.0, 2 * math.pi, n_lon + 1)
    T, Pm = np.meshgrid(th, ph, indexing="ij")
    N = np.stack([np.sin(T) * np.cos(Pm), np.cos(T), np.sin(T) * np.sin(Pm)], -1).reshape(-1, 3)
    P = N * radius + np.asarray(center)
    UV = np.stack([Pm / (2 * math.pi), T / math.pi], -1).reshape(-1, 2)
    idx = []
    for i in range(n_lat):
        for j in range(n_lon):
            a = i * (n_lon + 1) + j
            b = a + n_lon + 1
            if i != 0:
                idx += [a, a + 1, b]
            if i != n_lat - 1:
                idx += [a + 1, b + 1, b]
    return P.astype(f32), N.astype(f32), UV.astype(f32), np.asarray(idx, np.uint32)


def grid_plane(y, x0, x1, z0, z1, nx, nz, height_fn=None):
    xs = np.linspace(x0, x1, nx + 1)
    zs = np.linspace(z0, z1, nz + 1)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    Y = np.full_like(X, y) if height_fn is None else height_fn(X, Z)
    P = np.stack([X, Y, Z], -1).reshape(-1, 3)
    if height_fn is None:
        N = np.tile([0.0, 1.0, 0.0], (len(P), 1))
    else:
        dYdx = np.gradient(Y, xs, axis=0)
        dYdz = np.gradient(Y, zs, axis=1)
        N = np.stack([-dYdx, np.ones_like(Y), -dYdz], -1).reshape(-1, 3)
        N /= np.linalg.norm(N, axis=1, keepdims=True)
    UV = np.stack([(X - x0) / (x1 - x0), (Z - z0) / (z1 - z0)], -1).reshape(-1, 2)
    i, j = np.meshgrid(np.arange(nx), np.arange(nz), indexing="ij")
    a = (i * (nz + 1) + j).reshape(-1)
    b = a + nz + 1
    idx = np.stack([a, a + 1, b + 1, a, b + 1, b], -1).reshape(-1)
    return P.astype(f32), N.astype(f32), UV.astype(f32), idx.astype(np.uint32)


# ----------------------------------------------------------------------------------------------- scene builder
class SceneBuffers:
    """Finalised Falcor-layout buffers (numpy, host) plus the PltptSceneDesc that points at them."""

    def __init__(self, arrays, scalars):
        self.arrays = arrays
        self.scalars = scalars
        self._keep = None
        self._keep_desc = []  # ctypes descriptor arrays handed out by desc()

    def desc(self, arrays=None, device=False) -> L.PltptSceneDesc:
        a = arrays or self.arrays

        def ptr(name):
            x = a.get(name)
            if x is None or (hasattr(x, "size") and x.size == 0) or (hasattr(x, "numel") and x.numel() == 0):
                return None
            return x.data_ptr() if hasattr(x, "data_ptr") else x.ctypes.data

        s = self.scalars
        d = L.PltptSceneDesc()
        d.vertices, d.vertexCount = ptr("vertices"), s["vertexCount"]
        d.indexData, d.indexDataBytes = ptr("indexData"), s["indexDataBytes"]
        d.geometryInstances, d.instanceCount = ptr("geometryInstances"), s["instanceCount"]
        d.meshes, d.meshCount = ptr("meshes"), s["meshCount"]
        d.worldMatrices, d.inverseTransposeWorldMatrices, d.matrixCount = ptr("worldMatrices"), ptr("inverseTransposeWorldMatrices"), s["matrixCount"]
        d.materialData, d.materialCount = ptr("materialData"), s["materialCount"]
        d.spectralProfiles, d.profileCount = ptr("spectralProfiles"), s["profileCount"]
        d.lights, d.lightCount = ptr("lights"), s["lightCount"]
        d.emissiveTriangles, d.emissiveTriangleCount = ptr("emissiveTriangles"), s["emissiveTriangleCount"]
        d.emissiveFlux = ptr("emissiveFlux")
        d.activeTriangles, d.activeTriangleCount = ptr("activeTriangles"), s["emissiveTriangleCount"]
        d.perMeshInstanceOffset = ptr("perMeshInstanceOffset")
        d.triangleAliasTable, d.invWeightsSum = ptr("triangleAliasTable"), s["invWeightsSum"]
        d.hitInfoTypeBits, d.hitInfoInstanceIDBits, d.hitInfoPrimitiveIndexBits = 3, s["instanceIDBits"], s["primitiveIndexBits"]
        d.pointersAreDevice = 1 if device else 0
        d.envMapTexels = ptr("envMapTexels")
        if d.envMapTexels:
            d.envMapWidth, d.envMapHeight = s["envMapWidth"], s["envMapHeight"]
            e = s["envMap"]
            d.envMap.transform[:] = [float(x) for x in np.asarray(e["transform"], np.float32).reshape(-1)]
            d.envMap.invTransform[:] = [float(x) for x in np.asarray(e["invTransform"], np.float32).reshape(-1)]
            d.envMap.tint[:] = [float(x) for x in e["tint"]]
            d.envMap.intensity = float(e["intensity"])
        tex = s.get("textures") or []
        if tex:  # descriptor arrays are host memory in both cases; they must outlive the call that reads `d`
            td = (L.PltptTextureDesc * len(tex))()
            for k, (w, h, fmt) in enumerate(tex):
                td[k].texels, td[k].width, td[k].height, td[k].format = ptr(f"texture{k}"), w, h, fmt
            d.textures, d.textureCount = C.cast(td, C.c_void_p), len(tex)
            self._keep_desc.append(td)
        smp = s.get("samplers") or []
        if smp:
            sd = (L.PltptSamplerDesc * len(smp))()
            for k, (flt, au, av, border) in enumerate(smp):
                sd[k].magFilter, sd[k].addressU, sd[k].addressV = flt, au, av
                sd[k].borderColor[:] = border
            d.samplers, d.samplerCount = C.cast(sd, C.c_void_p), len(smp)
            self._keep_desc.append(sd)
        return d

    def to_device(self, torch, device="cuda:0"):
        """Uploads every buffer as a uint8 torch tensor; returns (tensors, desc) — keep `tensors` alive while in use."""
        t = {}
        for k, v in self.arrays.items():
            if v is None or v.size == 0:
                continue
            t[k] = torch.from_numpy(np.ascontiguousarray(v).view(np.uint8).reshape(-1).copy()).to(device)
        return t, self.desc(t, device=True)

    @property
    def triangle_count(self):
        return self.scalars["triangleCount"]

    BLOB_ORDER = ("vertices", "indexData", "geometryInstances", "meshes", "worldMatrices", "inverseTransposeWorldMatrices", "materialData", "spectralProfiles",
                  "lights", "emissiveTriangles", "emissiveFlux", "activeTriangles", "perMeshInstanceOffset", "triangleAliasTable")

    def save(self, path, cam):
        """Writes the buffers, byte for byte, plus the camera block to the scene file host/pltpt_host.cpp reads."""
        s = self.scalars
        head = np.array([s["vertexCount"], s["indexDataBytes"], s["instanceCount"], s["meshCount"], s["matrixCount"], s["materialCount"], s["profileCount"],
                         s["lightCount"], s["emissiveTriangleCount"], s["emissiveTriangleCount"], 3, s["instanceIDBits"], s["primitiveIndexBits"], 0], np.uint64)
        with open(path, "wb") as f:
            f.write(b"PLTSCN1\0")
            f.write(head.tobytes())
            f.write(np.array([s["invWeightsSum"], 0.0], np.float32).tobytes())
            f.write(bytes(cam))
            for name in self.BLOB_ORDER:
                a = self.arrays.get(name)
                raw = b"" if a is None or a.size == 0 else np.ascontiguousarray(a).tobytes()
                f.write(np.uint64(len(raw)).tobytes())
                f.write(raw)
            env = self.arrays.get("envMapTexels")  # optional trailing section: u64 texel bytes, then w, h (u32), EnvMapData (112 B), texels
            f.write(np.uint64(0 if env is None else env.nbytes).tobytes())
            if env is not None:
                f.write(np.array([s["envMapWidth"], s["envMapHeight"]], np.uint32).tobytes())
                f.write(bytes(self.desc().envMap))
                f.write(np.ascontiguousarray(env).tobytes())
            # optional second trailing section: material textures and samplers -- u32 texture count, u32 sampler count, per texture
            # {u32 width, height, format, 0; u64 texel bytes; texels}, per sampler the 32 B PltptSamplerDesc
            tex, smp = s.get("textures") or [], s.get("samplers") or []
            f.write(np.array([len(tex), len(smp)], np.uint32).tobytes())
            for k, (w, h, fmt) in enumerate(tex):
                raw = np.ascontiguousarray(self.arrays[f"texture{k}"]).tobytes()
                f.write(np.array([w, h, fmt, 0], np.uint32).tobytes())
                f.write(np.uint64(len(raw)).tobytes())
                f.write(raw)
            for flt, au, av, border in smp:
                f.write(np.array([flt, au, av, 0], np.uint32).tobytes())
                f.write(np.array(border, np.float32).tobytes())


def _alloc_bits(count):
    return 0 if count <= 1 else int(count - 1).bit_length()  # HitInfo.cpp allocateBits


def generate_alias_table(weights):
    """EmissivePowerSampler::generateAliasTable (EmissivePowerSampler.cpp:86-170) without the final random
    permutation (the table is *a* valid alias table either way; parity is on the induced pdf flux/sum)."""
    w = np.asarray(weights, f32).copy()
    N = len(w)
    total = float(np.sum(w.astype(np.float64)))
    w = (w * f32(N / f32(total))).astype(f32)
    perm = sorted(range(N), key=lambda i: w[i])
    thresholds = np.zeros(N, f32)
    redirect = np.zeros(N, np.uint32)
    head, tail = 0, N - 1
    if N == 1:
        thresholds[0], redirect[0] = 1.0, 0
    while head != tail:
        i, j = perm[head], perm[tail]
        thresholds[i] = w[i]
        redirect[i] = j
        w[j] = f32(w[j] - f32(f32(1.0) - w[i]))
        if head == tail - 1:
            thresholds[j] = 1.0
            redirect[j] = j
            break
        elif w[j] < 1.0:
            perm[head], perm[tail] = perm[tail], perm[head]
            tail -= 1
        else:
            head += 1
    table = np.zeros((N, 2), np.uint32)
    prob = _half_bits(thresholds).astype(np.uint32) << 16
    lo_x = redirect & 0xFFFFFF
    lo_y = np.arange(N, dtype=np.uint32) & 0xFFFFFF
    table[:, 0] = prob | ((lo_x >> 8) & 0xFFFF)
    table[:, 1] = ((lo_x & 0xFF) << 24) | lo_y
    return table, f32(total)


class SceneBuilder:
    def __init__(self):
        self.meshes = []  # (P, N, UV, I, material, matrix)
        self.extra_instances = []  # (mesh index, material, matrix): further instances of an already added mesh
        self.materials = []
        self.profiles = []
        self.lights = []
        self.env = None  # (texels HxWx4 f32, EnvMapData dict)
        self.tangents = {}  # mesh index -> (n, 4) tangents or None
        self.textures = []  # (texels array, format name)
        self.samplers = []  # (magFilter, addressU, addressV, borderColor)

    # -- material textures (MaterialSystem.slang:283-291; BasicMaterialData.slang:92-100 slot order)
    TEX_SLOTS = {"baseColor": 0, "specular": 1, "emissive": 2, "normalMap": 3, "transmission": 4, "displacement": 5, "data1": 6, "data2": 7}

    def add_texture(self, texels, fmt="rgba32f"):
        """texels: (H, W, 4) float32 for rgba32f, float16 for rgba16f, uint8 for rgba8 / rgba8_srgb.  Returns the texture ID."""
        dt = {"rgba32f": np.float32, "rgba16f": np.float16, "rgba8": np.uint8, "rgba8_srgb": np.uint8}[fmt]
        t = np.ascontiguousarray(texels, dt)
        assert t.ndim == 3 and t.shape[2] == 4
        self.textures.append((t, fmt))
        return len(self.textures) - 1

    def add_sampler(self, linear=True, address_u=0, address_v=0, border=(0, 0, 0, 0)):
        """Sampler IDs index `samplers`; without any sampler every material uses the default (linear, wrap)."""
        self.samplers.append((1 if linear else 0, int(address_u), int(address_v), tuple(float(x) for x in border)))
        return len(self.samplers) - 1

    def set_texture(self, material, slot, texture_id, normal_map_type=1):
        """Points a material's TextureHandle at a texture (TextureHandle::setTextureID: mode = Texture).  For the normal-map slot the
        BasicMaterialData flags also get the NormalMapType (1 = RGB, 2 = RG)."""
        m = self.materials[material]
        m["tex"][self.TEX_SLOTS[slot]] = (1 << 29) | int(texture_id)
        if slot == "normalMap":
            m["flags"] = (int(m["flags"]) & ~(3 << 1)) | (int(normal_map_type) << 1)

    def set_sampler(self, material, sampler_id):
        m = self.materials[material]
        m["header"][1] = (int(m["header"][1]) & ~(0xFF << 17)) | (int(sampler_id) << 17)

    def set_alpha_mode(self, material, mask=True, threshold=0.5):
        m = self.materials[material]
        m["header"][1] = (int(m["header"][1]) & ~0x1FFFF) | int(_half_bits(threshold)) | (int(mask) << 16)

    def set_env_map(self, texels, intensity=1.0, tint=(1.0, 1.0, 1.0), rot_y=0.0):
        """EnvMap (Scene/Lights/EnvMap.cpp:60-110): lat-long RGBA32F texels; `rot_y` (radians) is the env map's rotation about +y
        (EnvMap::setRotation builds transform = rotation, invTransform = its inverse)."""
        c, s_ = math.cos(rot_y), math.sin(rot_y)
        R = np.array([[c, 0, s_, 0], [0, 1, 0, 0], [-s_, 0, c, 0]], f32)
        Ri = np.array([[c, 0, -s_, 0], [0, 1, 0, 0], [s_, 0, c, 0]], f32)
        self.env = (np.ascontiguousarray(texels, f32), dict(transform=R, invTransform=Ri, tint=tuple(tint), intensity=float(intensity)))

    # -- spectral profiles (SceneBuilder::addSpectralProfile*, SceneBuilder.cpp:962-999)
    def add_profile(self, wavelengths, values, compute_icdf=True, scale=1.0):
        self.profiles.append(gen_spectral_profile(wavelengths, values, compute_icdf, scale))
        return len(self.profiles) - 1

    def add_ior_profiles(self, name):
        """addSpectralProfileFromMaterial: (n, k) profile ids, no icdf."""
        t = bundled_tables()["ior"][name]
        return (self.add_profile(t["eta"]["wavelengths"], t["eta"]["values"], False),
                self.add_profile(t["k"]["wavelengths"], t["k"]["values"], False))

    def add_emission_profile(self, name, scale=1.0):
        t = bundled_tables()["emission"][name]
        return self.add_profile(t["wavelengths"], t["values"], True, scale)

    # -- materials: slot assignments follow the *Material.slang unpackers (see oracle/plt_bsdf.h MaterialInstance)
    def _add(self, m):
        self.materials.append(m)
        return len(self.materials) - 1

    def add_diffuse(self, base_color, emission_profile=None, double_sided=False):
        m = _blank_material(L.MATERIAL_DIFFUSE, double_sided)
        m["baseColor"] = list(base_color) + [1.0]
        if emission_profile is not None:
            m["emissionSpectralId"] = emission_profile
            m["header"] = _header(L.MATERIAL_DIFFUSE, double_sided, True)
        return self._add(m)

    def add_conductor(self, ior_profiles, roughness, gamma=3.0, specular=(1, 1, 1), ext_eta=1.0):
        m = _blank_material(L.MATERIAL_CONDUCTOR)
        m["baseColor"] = list(specular) + [1.0]
        m["specular"] = [gamma, roughness, 1.0, 1.0]
        m["IoR"] = ext_eta
        m["iorNSpectralId"], m["iorKSpectralId"] = ior_profiles
        return self._add(m)

    def add_dielectric(self, ior, roughness, ext_eta=1.0, dispersion_b=0.0, specular=(1, 1, 1), transmission=(1, 1, 1)):
        m = _blank_material(L.MATERIAL_DIELECTRIC)  # single-sided: N must not flip inside the medium (MitsubaImporter.cpp:330)
        m["baseColor"] = list(specular) + [1.0]
        m["specular"] = [dispersion_b, roughness, 0.0, ext_eta]
        m["transmission"] = transmission
        m["IoR"] = ior
        return self._add(m)

    def add_opaque_dielectric(self, base_color, roughness):
        m = _blank_material(L.MATERIAL_OPAQUE_DIELECTRIC)
        m["baseColor"] = list(base_color) + [1.0]
        m["specular"] = [0.0, roughness, 0.0, 0.0]
        return self._add(m)

    def add_thin_dielectric(self, ior, tau_mm, birefringence=0.0, optic_axis_xy=(0.15, 0.04), birefringence_scale=1.0):
        m = _blank_material(L.MATERIAL_THIN_DIELECTRIC)
        m["baseColor"] = [birefringence, optic_axis_xy[0], optic_axis_xy[1], 1.0]
        m["specular"] = [birefringence_scale, 0.0, 0.0, tau_mm]
        m["IoR"] = ior
        return self._add(m)

    def add_coated_conductor(self, ior_profiles, roughness, tau_um, coat_ior, specular=(1, 1, 1), ext_eta=1.0):
        m = _blank_material(L.MATERIAL_COATED_CONDUCTOR)
        m["baseColor"] = list(specular) + [1.0]
        m["specular"] = [0.0, roughness, 1.0, 0.0]
        _set_data1(m, (tau_um, 0, 0, coat_ior))
        m["IoR"] = ext_eta
        m["iorNSpectralId"], m["iorKSpectralId"] = ior_profiles
        return self._add(m)

    def add_coated_opaque_dielectric(self, base_color, roughness, tau_um, coat_ior, ext_eta=1.0):
        m = _blank_material(L.MATERIAL_COATED_OPAQUE_DIELECTRIC)
        m["baseColor"] = list(base_color) + [1.0]
        m["specular"] = [0.0, roughness, 1.0, 0.0]
        _set_data1(m, (tau_um, 0, 0, coat_ior))
        m["IoR"] = ext_eta
        return self._add(m)

    def _grating(self, m, pitch_um, direction, q, q_scale, power_mult, roughness, gtype, lobes):
        m["specular"] = [q_scale, power_mult, 0.0, 0.0]
        _set_data1(m, (pitch_um[0], pitch_um[1], direction, 0), w_bits=((gtype & 0xFF) << 8) | (lobes & 0xFF))
        m["data2"] = [q, roughness, 0.0, 0.0]

    def add_grated_conductor(self, ior_profiles, pitch_um, q, roughness=0.0, direction=0.0, q_scale=1.0, power_mult=1.0, gtype=0, lobes=3, specular=(1, 1, 1), ext_eta=1.0):
        m = _blank_material(L.MATERIAL_GRATED_CONDUCTOR)
        m["baseColor"] = list(specular) + [1.0]
        self._grating(m, pitch_um, direction, q, q_scale, power_mult, roughness, gtype, lobes)
        m["IoR"] = ext_eta
        m["iorNSpectralId"], m["iorKSpectralId"] = ior_profiles
        return self._add(m)

    def add_grated_opaque_dielectric(self, base_color, pitch_um, q, roughness=0.0, direction=0.0, q_scale=1.0, power_mult=1.0, gtype=0, lobes=3):
        m = _blank_material(L.MATERIAL_GRATED_OPAQUE_DIELECTRIC)
        m["baseColor"] = list(base_color) + [1.0]
        self._grating(m, pitch_um, direction, q, q_scale, power_mult, roughness, gtype, lobes)
        return self._add(m)

    def add_multilayered_stack(self, base_color, layers, thickness_um=(0.1, 0.15), iors=(1.38, 2.3), ext_eta=1.0):
        m = _blank_material(L.MATERIAL_MULTILAYERED_STACK)
        m["baseColor"] = list(base_color) + [1.0]
        _set_data1(m, (thickness_um[0], thickness_um[1], 0, 0), w_bits=layers & 0xFFFF)
        m["data2"] = [iors[0], iors[1], 0.0, 0.0]
        m["IoR"] = ext_eta
        return self._add(m)

    # -- geometry
    def add_mesh(self, mesh, material, matrix=None, tangents=None):
        """`tangents`: optional (n, 4) per-vertex tangent with handedness in w (needed for normal maps: a mesh without tangents gets an
        invented frame and modifyTangentSpace is skipped, MaterialFactory.slang:112-120)."""
        P, N, UV, I = mesh
        self.meshes.append((np.asarray(P, f32), np.asarray(N, f32), np.asarray(UV, f32), np.asarray(I, np.uint32),
                            material, np.eye(4, dtype=f32) if matrix is None else np.asarray(matrix, f32)))
        self.tangents[len(self.meshes) - 1] = None if tangents is None else np.asarray(tangents, f32)
        return len(self.meshes) - 1

    def add_instance(self, mesh_index, material, matrix):
        """Another GeometryInstanceData of an existing mesh (same geometryID, own matrix and material)."""
        self.extra_instances.append((mesh_index, material, np.asarray(matrix, f32)))

    # -- analytic lights (Light.cpp:366-401 for areas; LightData.slang:49-72)
    def add_rect_light(self, profile, center, normal, half_x, half_y):
        n = np.asarray(normal, np.float64) / np.linalg.norm(normal)
        t = np.cross([0, 1, 0] if abs(n[1]) < 0.9 else [1, 0, 0], n)
        t /= np.linalg.norm(t)
        b = np.cross(n, t)
        M = np.eye(4)
        M[:3, 0], M[:3, 1], M[:3, 2], M[:3, 3] = t * half_x, b * half_y, n, center
        l = np.zeros((), L.light_data_dtype)
        l["posW"], l["type"], l["dirW"] = center, 3, n
        l["openingAngle"], l["cosOpeningAngle"], l["cosSubtendedAngle"] = math.pi, -1.0, 0.9999893
        l["intensitySpectrumId"] = profile
        l["surfaceArea"] = 4.0 * half_x * half_y
        l["transMat"] = M.astype(f32).reshape(-1)
        l["transMatIT"] = np.linalg.inv(M).T.astype(f32).reshape(-1)
        self.lights.append(l)

    def _area_light(self, ltype, profile, center, normal, scale, emission_solid_angle=None):
        """AnalyticAreaLight::update (Light.cpp:341-346): transMat = transform * scale, transMatIT = inverse(transpose(transMat));
        object-space +z is the emitting side of the rect / disc."""
        n = np.asarray(normal, np.float64) / np.linalg.norm(normal)
        t = np.cross([0, 1, 0] if abs(n[1]) < 0.9 else [1, 0, 0], n)
        t /= np.linalg.norm(t)
        b = np.cross(n, t)
        M = np.eye(4)
        M[:3, 0], M[:3, 1], M[:3, 2], M[:3, 3] = t * scale[0], b * scale[1], n * scale[2], center
        M = M.astype(f32)
        l = np.zeros((), L.light_data_dtype)
        l["posW"], l["type"], l["dirW"] = center, ltype, n
        l["openingAngle"], l["cosOpeningAngle"] = math.pi, -1.0
        # AnalyticAreaLight::setEmissionSolidAngle (Light.cpp:348-351); LightData default cosSubtendedAngle = 0.9999893
        l["cosSubtendedAngle"] = 0.9999893 if emission_solid_angle is None else math.sqrt(1.0 - emission_solid_angle / 4.0 / math.pi)
        l["intensitySpectrumId"] = profile
        rx, ry, rz = (float(np.linalg.norm(M[:3, k])) for k in range(3))
        if ltype == 3:
            area = 4.0 * rx * ry                                                   # RectLight::update, Light.cpp:366
        elif ltype == 4:
            area = math.pi * rx * ry                                               # DiscLight::update, Light.cpp:383
        else:  # SphereLight::update, Light.cpp:401 — the reference's operator precedence (only the last term is divided by 3) is kept
            area = 4.0 * math.pi * ((rx * ry) ** 1.6 + (ry * rz) ** 1.6 + (rx * rz) ** 1.6 / 3.0) ** (1.0 / 1.6)
        l["surfaceArea"] = f32(area)
        l["transMat"] = M.reshape(-1)
        l["transMatIT"] = np.linalg.inv(M.astype(np.float64).T).astype(f32).reshape(-1)
        self.lights.append(l)

    def add_disc_light(self, profile, center, normal, radius_x, radius_y=None, emission_solid_angle=None):
        self._area_light(4, profile, center, normal, (radius_x, radius_x if radius_y is None else radius_y, 1.0), emission_solid_angle)

    def add_sphere_light(self, profile, center, radius, emission_solid_angle=None):
        r = (radius,) * 3 if np.isscalar(radius) else tuple(radius)
        self._area_light(5, profile, center, (0, 0, 1), r, emission_solid_angle)

    def add_distant_light(self, profile, direction, half_angle):
        d = np.asarray(direction, np.float64) / np.linalg.norm(direction)
        z = -d  # cone axis points towards the light
        t = np.cross([0, 1, 0] if abs(z[1]) < 0.9 else [1, 0, 0], z)
        t /= np.linalg.norm(t)
        b = np.cross(z, t)
        M = np.eye(4)
        M[:3, 0], M[:3, 1], M[:3, 2] = t, b, z
        l = np.zeros((), L.light_data_dtype)
        l["type"], l["dirW"] = 2, d
        l["openingAngle"], l["cosOpeningAngle"], l["cosSubtendedAngle"] = math.pi, -1.0, math.cos(half_angle)
        l["intensitySpectrumId"] = profile
        l["transMat"] = M.astype(f32).reshape(-1)
        l["transMatIT"] = np.linalg.inv(M).T.astype(f32).reshape(-1)
        self.lights.append(l)

    def finalize(self) -> SceneBuffers:
        verts, idx, insts, mds, mats = [], [], [], [], []
        vb, ib, max_prims = 0, 0, 0
        mesh_offsets = []
        for i, (P, N, UV, I, material, M) in enumerate(self.meshes):
            verts.append(pack_vertices(P, N, UV, self.tangents.get(i)))
            idx.append(I)
            md = np.zeros((), L.mesh_desc_dtype)
            md["vbOffset"], md["ibOffset"], md["vertexCount"], md["indexCount"], md["materialID"] = vb, ib, len(P), len(I), material
            mds.append(md)
            mesh_offsets.append((vb, ib))
            vb += len(P)
            ib += len(I)
            max_prims = max(max_prims, len(I) // 3)
        # instances: one per mesh (in mesh order), then the extra instances
        inst_list = [(i, m[4], m[5]) for i, m in enumerate(self.meshes)] + list(self.extra_instances)
        tri_total = 0
        for k, (mi, material, M) in enumerate(inst_list):
            gi = np.zeros((), L.geometry_instance_dtype)
            gi["flags"] = 1 << 29  # GeometryType::TriangleMesh in the top 3 bits, 32-bit indices, CCW front faces
            gi["globalMatrixID"], gi["materialID"], gi["geometryID"] = k, material, mi
            gi["vbOffset"], gi["ibOffset"] = mesh_offsets[mi]
            gi["instanceIndex"], gi["geometryIndex"] = k, 0
            insts.append(gi)
            mats.append(M)
            tri_total += len(self.meshes[mi][3]) // 3
        world = np.stack(mats).astype(f32) if mats else np.zeros((0, 4, 4), f32)
        invT = np.stack([np.linalg.inv(m.astype(np.float64)).T for m in mats]).astype(f32) if mats else world
        materials = np.stack(self.materials) if self.materials else np.zeros(0, L.material_blob_dtype)
        profiles = np.stack(self.profiles) if self.profiles else np.zeros(0, L.spectral_profile_dtype)

        # LightCollection (BuildTriangleList.cs.slang, FinalizeIntegration.cs.slang:64-110)
        etris, flux, offsets = [], [], np.full(len(inst_list), 0xFFFFFFFF, np.uint32)
        for i, (mi, material, M) in enumerate(inst_list):
            P, N, UV, I = self.meshes[mi][:4]
            if materials[material]["emissionSpectralId"] == L.NO_SPECTRAL_PROFILE:
                continue
            offsets[i] = len(etris)
            Pw = (P.astype(np.float64) @ M[:3, :3].T.astype(np.float64) + M[:3, 3]).astype(f32)
            rgb = np.maximum(profiles[materials[material]["emissionSpectralId"]]["rgb"], 0)
            for t in range(len(I) // 3):
                a, b, c = Pw[I[3 * t]], Pw[I[3 * t + 1]], Pw[I[3 * t + 2]]
                n = np.cross((b - a).astype(f32), (c - a).astype(f32)).astype(f32)
                area = f32(0.5) * f32(np.linalg.norm(n))
                n = n / np.linalg.norm(n)
                e = np.zeros((), L.emissive_triangle_dtype)
                for k, v in enumerate((I[3 * t], I[3 * t + 1], I[3 * t + 2])):
                    e["posAndTexCoords"][k, :3] = Pw[v]
                    uvb = _half_bits(UV[v]).astype(np.uint32)
                    e["posAndTexCoords"][k, 3] = np.array([uvb[0] | (uvb[1] << 16)], np.uint32).view(f32)[0]
                e["normal"] = encode_normal_2x16(n)
                e["area"] = np.array([area], f32).view(np.uint32)[0]
                e["materialID"], e["lightIdx"] = material, 0
                etris.append(e)
                fl = np.zeros((), L.emissive_flux_dtype)
                lum = f32(np.dot(rgb, np.array([0.2126, 0.7152, 0.0722], f32)))
                fl["flux"] = f32(lum * area * f32(math.pi))
                fl["averageRadiance"] = rgb
                flux.append(fl)
        etris = np.stack(etris) if etris else np.zeros(0, L.emissive_triangle_dtype)
        flux = np.stack(flux) if flux else np.zeros(0, L.emissive_flux_dtype)
        if len(etris):
            alias, wsum = generate_alias_table(flux["flux"])
            inv_sum = float(f32(1.0) / f32(wsum))
        else:
            alias, inv_sum = np.zeros((0, 2), np.uint32), 0.0
        arrays = dict(
            vertices=np.concatenate(verts) if verts else np.zeros(0, L.packed_vertex_dtype),
            indexData=np.concatenate(idx).astype(np.uint32) if idx else np.zeros(0, np.uint32),
            geometryInstances=np.stack(insts) if insts else np.zeros(0, L.geometry_instance_dtype),
            meshes=np.stack(mds) if mds else np.zeros(0, L.mesh_desc_dtype),
            worldMatrices=np.ascontiguousarray(world), inverseTransposeWorldMatrices=np.ascontiguousarray(invT),
            materialData=materials, spectralProfiles=profiles,
            lights=np.stack(self.lights) if self.lights else np.zeros(0, L.light_data_dtype),
            emissiveTriangles=etris, emissiveFlux=flux, activeTriangles=np.arange(len(etris), dtype=np.uint32),
            perMeshInstanceOffset=offsets, triangleAliasTable=alias,
        )
        scalars = dict(
            vertexCount=vb, indexDataBytes=ib * 4, instanceCount=len(insts), meshCount=len(mds), matrixCount=len(mats),
            materialCount=len(materials), profileCount=len(profiles), lightCount=len(self.lights),
            emissiveTriangleCount=len(etris), invWeightsSum=inv_sum, triangleCount=tri_total,
            instanceIDBits=_alloc_bits(len(insts)), primitiveIndexBits=_alloc_bits(max_prims),
        )
        if self.env is not None:
            arrays["envMapTexels"] = self.env[0]
            scalars.update(envMapWidth=self.env[0].shape[1], envMapHeight=self.env[0].shape[0], envMap=self.env[1])
        for k, (t, fmt) in enumerate(self.textures):
            arrays[f"texture{k}"] = t
        scalars["textures"] = [(t.shape[1], t.shape[0], L.TEXTURE_FORMATS[fmt]) for t, fmt in self.textures]
        scalars["samplers"] = list(self.samplers)
        return SceneBuffers(arrays, scalars)


def procedural_sky(width=256, height=128, sun_dir=(0.4, 0.6, 0.5), sun_radiance=60.0, seed=3):
    """Synthetic lat-long environment map (no asset files offline): horizon-to-zenith gradient, a dim ground, a soft sun lobe
    and low-amplitude seeded noise so that neighbouring texels differ (exercises the bilinear filter and the importance map)."""
    v = (np.arange(height, dtype=np.float64) + 0.5) / height
    u = (np.arange(width, dtype=np.float64) + 0.5) / width
    U, V = np.meshgrid(u, v)
    phi, theta = math.pi * (2.0 * U - 1.0), math.pi * V            # latlong_map_to_world (MathHelpers.slang:120-129)
    d = np.stack([np.sin(theta) * np.sin(phi), np.cos(theta), -np.sin(theta) * np.cos(phi)], -1)
    up = np.clip(d[..., 1], 0.0, 1.0)
    sky = np.stack([0.25 + 0.35 * (1 - up), 0.35 + 0.35 * (1 - up), 0.55 + 0.3 * (1 - up)], -1)
    ground = np.array([0.12, 0.10, 0.08])
    rgb = np.where(d[..., 1:2] >= 0.0, sky, ground)
    sd = np.asarray(sun_dir, np.float64)
    sd /= np.linalg.norm(sd)
    rgb = rgb + sun_radiance * np.array([1.0, 0.9, 0.75]) * np.exp(-(1.0 - np.clip(d @ sd, -1, 1)) / 0.004)[..., None]
    rng = np.random.default_rng(seed)
    rgb = rgb * (1.0 + 0.1 * (rng.random((height, width, 1)) - 0.5))
    return np.concatenate([rgb, np.ones((height, width, 1))], -1).astype(f32)


# ----------------------------------------------------------------------------------------------- camera
def halton(index, base):
    result, factor = f32(0.0), f32(1.0)
    while index > 0:
        factor = f32(factor / f32(base))
        result = f32(result + factor * f32(index % base))
        index //= base
    return result


def halton_jitter(frame, width, height, sample_count=16):
    """Camera jitter GBufferRT sets for frame `frame` (HaltonSamplePattern.cpp:59-67 * 1/frameDim)."""
    i = frame % sample_count if sample_count else frame
    v = np.array([halton(i, 2), halton(i, 3)], f32)
    v = (v + f32(0.5)) - np.floor(v + f32(0.5)) - f32(0.5)
    return float(f32(v[0]) / f32(width)), float(f32(v[1]) / f32(height))


def make_camera(pos, target, up=(0, 1, 0), focal_length=21.0, frame_height=24.0, aspect=16.0 / 9.0, focal_distance=10000.0,
                jitter=(0.0, 0.0), polarizer=0.0, polarizer_theta=0.0) -> L.PltptCameraData:
    """Camera::calculateCameraParameters (Camera.cpp:111-196): cameraU/V/W, view matrix (glm::lookAt, RH)."""
    cam = L.PltptCameraData()
    pos, target, up = (np.asarray(v, np.float64) for v in (pos, target, up))
    fovY = 2.0 * math.atan(0.5 * frame_height / focal_length)
    W = (target - pos) / np.linalg.norm(target - pos) * focal_distance
    U = np.cross(W, up)
    U /= np.linalg.norm(U)
    V = np.cross(U, W)
    V /= np.linalg.norm(V)
    U *= focal_distance * math.tan(fovY * 0.5) * aspect
    V *= focal_distance * math.tan(fovY * 0.5)
    f = (target - pos) / np.linalg.norm(target - pos)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    view = np.eye(4)
    view[0, :3], view[1, :3], view[2, :3] = s, u, -f
    view[0, 3], view[1, 3], view[2, 3] = -s @ pos, -u @ pos, f @ pos
    for i, x in enumerate(view.astype(f32).reshape(-1)):
        cam.viewMat[i] = x
        cam.prevViewMat[i] = x
    for name, v in (("posW", pos), ("prevPosW", pos), ("up", up), ("target", target), ("cameraU", U), ("cameraV", V), ("cameraW", W)):
        for i in range(3):
            getattr(cam, name)[i] = float(f32(v[i]))
    cam.focalLength, cam.aspectRatio, cam.nearZ, cam.farZ = focal_length, aspect, 0.1, 1000.0
    cam.jitterX, cam.jitterY = jitter
    cam.frameHeight, cam.frameWidth, cam.focalDistance = frame_height, frame_height * aspect, focal_distance
    cam.shutterSpeed, cam.ISOSpeed = 0.004, 100.0
    cam.polarizer, cam.polarizerTheta = polarizer, polarizer_theta
    return cam


def with_jitter(cam: L.PltptCameraData, frame, width, height) -> L.PltptCameraData:
    c = L.PltptCameraData.from_buffer_copy(bytes(cam))
    c.jitterX, c.jitterY = halton_jitter(frame, width, height)
    return c


# ----------------------------------------------------------------------------------------------- named scenes
LAMP = "2475_INC_A19_domestic-use_softwhite-color_Rona-brand_100W_130V_2450K"


def cornell_box(aspect=16.0 / 9.0, dielectric_roughness=0.1, conductor_roughness=0.3, emission_scale=1.0):
    """BASELINE config 2: Cornell box, PLTDiffuse walls, rough Dielectric box, rough Au Conductor box, ceiling area
    emitter with an LSPDD spectral profile."""
    b = SceneBuilder()
    au = b.add_ior_profiles("Au")
    lamp = b.add_emission_profile(LAMP, emission_scale)
    white = b.add_diffuse((0.73, 0.73, 0.73))
    red = b.add_diffuse((0.65, 0.05, 0.05))
    green = b.add_diffuse((0.12, 0.45, 0.15))
    light = b.add_diffuse((0.0, 0.0, 0.0), emission_profile=lamp)
    glass = b.add_dielectric(ior=1.5, roughness=dielectric_roughness)
    gold = b.add_conductor(au, roughness=conductor_roughness, gamma=3.0)
    # room: x in [-X,X] (16:9-shaped so that every pixel sees the interior), y in [0,2], z in [-1,1]; camera looks down -z
    X = 1.6
    b.add_mesh(quad([-X, 0, 1], [X, 0, 1], [X, 0, -1], [-X, 0, -1]), white)       # floor (normal +y)
    b.add_mesh(quad([-X, 2, -1], [X, 2, -1], [X, 2, 1], [-X, 2, 1]), white)       # ceiling (normal -y)
    b.add_mesh(quad([-X, 0, -1], [X, 0, -1], [X, 2, -1], [-X, 2, -1]), white)     # back wall (normal +z)
    b.add_mesh(quad([-X, 0, 1], [-X, 0, -1], [-X, 2, -1], [-X, 2, 1]), red)       # left wall (normal +x)
    b.add_mesh(quad([X, 0, -1], [X, 0, 1], [X, 2, 1], [X, 2, -1]), green)         # right wall (normal -x)
    b.add_mesh(quad([-0.3, 1.995, -0.3], [0.3, 1.995, -0.3], [0.3, 1.995, 0.3], [-0.3, 1.995, 0.3]), light)  # emitter (normal -y)
    b.add_mesh(box([-0.6, 0.6, -0.3], [0.3, 0.6, 0.3], rot_y=0.3), gold)
    b.add_mesh(box([0.6, 0.3, 0.3], [0.3, 0.3, 0.3], rot_y=-0.25), glass)
    cam = make_camera([0.0, 1.0, 3.5], [0.0, 1.0, 0.0], focal_length=35.0, aspect=aspect)
    return b.finalize(), cam


def cornell_box_closed(aspect=16.0 / 9.0, albedo=0.8):
    """Deep-path variant of the Cornell box: the same room closed by a front wall, every wall at albedo 0.8, camera
    inside.  No path can escape, so paths run to the bounce limit unless Russian roulette ends them: the O(L^2) forward replay of the
    solve stage is what this workload stresses."""
    b = SceneBuilder()
    au = b.add_ior_profiles("Au")
    lamp = b.add_emission_profile(LAMP, 1.0)
    white = b.add_diffuse((albedo, albedo, albedo))
    red = b.add_diffuse((albedo, 0.1 * albedo, 0.1 * albedo))
    green = b.add_diffuse((0.15 * albedo, albedo, 0.2 * albedo))
    light = b.add_diffuse((0.0, 0.0, 0.0), emission_profile=lamp)
    glass = b.add_dielectric(ior=1.5, roughness=0.1)
    gold = b.add_conductor(au, roughness=0.3, gamma=3.0)
    X = 1.6
    b.add_mesh(quad([-X, 0, 1], [X, 0, 1], [X, 0, -1], [-X, 0, -1]), white)       # floor
    b.add_mesh(quad([-X, 2, -1], [X, 2, -1], [X, 2, 1], [-X, 2, 1]), white)       # ceiling
    b.add_mesh(quad([-X, 0, -1], [X, 0, -1], [X, 2, -1], [-X, 2, -1]), white)     # back wall (normal +z)
    b.add_mesh(quad([X, 0, 1], [-X, 0, 1], [-X, 2, 1], [X, 2, 1]), white)         # front wall (normal -z), behind the camera
    b.add_mesh(quad([-X, 0, 1], [-X, 0, -1], [-X, 2, -1], [-X, 2, 1]), red)
    b.add_mesh(quad([X, 0, -1], [X, 0, 1], [X, 2, 1], [X, 2, -1]), green)
    b.add_mesh(quad([-0.3, 1.995, -0.3], [0.3, 1.995, -0.3], [0.3, 1.995, 0.3], [-0.3, 1.995, 0.3]), light)
    b.add_mesh(box([-0.6, 0.6, -0.3], [0.3, 0.6, 0.3], rot_y=0.3), gold)
    b.add_mesh(box([0.6, 0.3, 0.3], [0.3, 0.3, 0.3], rot_y=-0.25), glass)
    cam = make_camera([0.0, 1.0, 0.95], [0.0, 0.95, -1.0], focal_length=12.0, aspect=aspect)
    return b.finalize(), cam


def quad_tangents(mesh):
    """Per-vertex tangents of a quad() mesh: T along increasing u (p1 - p0), handedness so that cross(N, T) * w follows increasing v."""
    P, N, UV, I = mesh
    t = (P[1] - P[0]) / np.linalg.norm(P[1] - P[0])
    w = 1.0 if np.dot(np.cross(N[0], t), P[3] - P[0]) > 0 else -1.0
    return np.tile(np.append(t, w).astype(f32), (len(P), 1))


def procedural_textures(seed=5):
    """Synthetic material textures (no asset files offline), one per supported format: an sRGB colour checker with per-texel noise
    (RGBA8 sRGB, alpha 1), a tangent-space normal map of a sinusoidal bump field (RGBA8 UNORM), a specular-slot map that varies the
    roughness (RGBA32F; .x = gamma, .y = roughness), a base-colour map in RGBA16F, and a cut-out whose alpha is exactly 0 or 1."""
    rng = np.random.default_rng(seed)
    H, W = 64, 96
    y, x = np.mgrid[0:H, 0:W]
    checker = ((x // 12 + y // 8) % 2).astype(np.float64)
    col = np.stack([0.25 + 0.6 * checker, 0.55 - 0.3 * checker, 0.2 + 0.5 * (x / W)], -1) + rng.uniform(-0.05, 0.05, (H, W, 3))
    albedo = np.concatenate([np.clip(col, 0, 1), np.ones((H, W, 1))], -1)
    albedo8 = np.round(albedo * 255).astype(np.uint8)
    u, v = (x + 0.5) / W, (y + 0.5) / H
    dhdu, dhdv = 0.35 * np.cos(6 * np.pi * u) * np.cos(4 * np.pi * v), -0.35 * np.sin(6 * np.pi * u) * np.sin(4 * np.pi * v)
    n = np.stack([-dhdu, -dhdv, np.ones_like(u)], -1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    normal8 = np.round(np.concatenate([n * 0.5 + 0.5, np.ones((H, W, 1))], -1) * 255).astype(np.uint8)
    spec = np.zeros((32, 32, 4), f32)
    yy, xx = np.mgrid[0:32, 0:32]
    spec[..., 0] = 3.0                                              # gamma
    spec[..., 1] = 0.15 + 0.35 * (0.5 + 0.5 * np.sin(xx * 0.7) * np.cos(yy * 0.5))  # roughness
    spec[..., 2] = spec[..., 3] = 1.0
    base16 = np.concatenate([0.2 + 0.7 * rng.uniform(size=(16, 16, 3)), np.ones((16, 16, 1))], -1).astype(np.float16)
    cut = np.ones((16, 16, 4), f32)
    cut[..., :3] = [0.7, 0.7, 0.2]
    cut[..., 3] = ((xx[:16, :16] + yy[:16, :16]) % 2).astype(f32)   # alternating opaque / fully transparent texels
    emis = np.ones((8, 8, 4), f32)
    emis[..., 0], emis[..., 1], emis[..., 2] = 0.2 + 0.8 * (xx[:8, :8] / 7.0), 0.9, 1.0 - 0.7 * (yy[:8, :8] / 7.0)
    return dict(albedo8=albedo8, normal8=normal8, spec=spec, base16=base16, cut=cut, emis=emis)


def textured_scene(aspect=16.0 / 9.0):
    """Cornell-shaped room whose materials read their parameter slots from textures (MaterialSystem.slang:283-291): an
    sRGB base-colour map on the floor, a normal-mapped Au conductor with a roughness map on the back wall, an RGBA16F base-colour map on
    an opaque dielectric box (clamp sampler), an emissive texture on the ceiling emitter, and an alpha cut-out (point sampler, alpha
    exactly 0 or 1) in front of the camera that the alpha-masking of the sample stage passes through."""
    b = SceneBuilder()
    tex = procedural_textures()
    au = b.add_ior_profiles("Au")
    lamp = b.add_emission_profile(LAMP, 1.0)
    t_albedo = b.add_texture(tex["albedo8"], "rgba8_srgb")
    t_normal = b.add_texture(tex["normal8"], "rgba8")
    t_spec = b.add_texture(tex["spec"], "rgba32f")
    t_base16 = b.add_texture(tex["base16"], "rgba16f")
    t_cut = b.add_texture(tex["cut"], "rgba32f")
    t_emis = b.add_texture(tex["emis"], "rgba32f")
    s_wrap = b.add_sampler(True, 0, 0)
    s_clamp = b.add_sampler(True, 2, 2)
    s_point = b.add_sampler(False, 1, 4)          # point filter, mirror in u, mirror-once in v
    s_border = b.add_sampler(True, 3, 3, (0.1, 0.6, 0.1, 1.0))
    white = b.add_diffuse((0.73, 0.73, 0.73))
    floor = b.add_diffuse((1.0, 1.0, 1.0))
    b.set_texture(floor, "baseColor", t_albedo); b.set_sampler(floor, s_wrap)
    red = b.add_diffuse((0.65, 0.05, 0.05))
    green = b.add_diffuse((1.0, 1.0, 1.0))
    b.set_texture(green, "baseColor", t_base16); b.set_sampler(green, s_border)
    light = b.add_diffuse((0.0, 0.0, 0.0), emission_profile=lamp)
    b.set_texture(light, "emissive", t_emis); b.set_sampler(light, s_clamp)
    gold = b.add_conductor(au, roughness=0.3, gamma=3.0)
    b.set_texture(gold, "specular", t_spec); b.set_texture(gold, "normalMap", t_normal, 1); b.set_sampler(gold, s_wrap)
    coat = b.add_opaque_dielectric((1.0, 1.0, 1.0), roughness=0.2)
    b.set_texture(coat, "baseColor", t_base16); b.set_texture(coat, "normalMap", t_normal, 2); b.set_sampler(coat, s_clamp)
    cutout = b.add_diffuse((1.0, 1.0, 1.0), double_sided=True)
    b.set_texture(cutout, "baseColor", t_cut); b.set_sampler(cutout, s_point)
    X = 1.6
    m = quad([-X, 0, 1], [X, 0, 1], [X, 0, -1], [-X, 0, -1]); m = (m[0], m[1], m[2] * np.array([3.0, 2.0], f32), m[3])
    b.add_mesh(m, floor)                                                                                  # floor, uv tiled 3 x 2 (wrap)
    b.add_mesh(quad([-X, 2, -1], [X, 2, -1], [X, 2, 1], [-X, 2, 1]), white)
    back = quad([-X, 0, -1], [X, 0, -1], [X, 2, -1], [-X, 2, -1])
    b.add_mesh(back, gold, tangents=quad_tangents(back))                                                  # normal-mapped conductor wall
    b.add_mesh(quad([-X, 0, 1], [-X, 0, -1], [-X, 2, -1], [-X, 2, 1]), red)
    rw = quad([X, 0, -1], [X, 0, 1], [X, 2, 1], [X, 2, -1]); rw = (rw[0], rw[1], rw[2] * np.array([1.5, 1.5], f32) - np.array([0.25, 0.25], f32), rw[3])
    b.add_mesh(rw, green)                                                                                 # uv beyond [0,1]: border colour
    b.add_mesh(quad([-0.4, 1.995, -0.4], [0.4, 1.995, -0.4], [0.4, 1.995, 0.4], [-0.4, 1.995, 0.4]), light)
    bx = box([0.5, 0.35, 0.2], [0.35, 0.35, 0.35], rot_y=-0.3)
    P, N, UV, I = bx
    T = np.zeros((len(P), 4), f32)                # per-face tangents: any unit vector orthogonal to the face normal, w = +1
    for i in range(len(P)):
        a = np.array([0.0, 1.0, 0.0]) if abs(N[i][1]) < 0.9 else np.array([1.0, 0.0, 0.0])
        t = np.cross(a, N[i]); t /= np.linalg.norm(t)
        T[i] = [t[0], t[1], t[2], 1.0]
    b.add_mesh(bx, coat, tangents=T)
    cq = quad([-1.2, 0.3, 0.6], [-0.2, 0.3, 0.6], [-0.2, 1.3, 0.6], [-1.2, 1.3, 0.6])
    cq = (cq[0], cq[1], cq[2] * np.array([1.0, 2.5], f32) - np.array([0.0, 1.0], f32), cq[3])               # v in [-1, 1.5]: mirror-once
    b.add_mesh(cq, cutout)
    cam = make_camera([0.0, 1.0, 3.5], [0.0, 1.0, 0.0], focal_length=35.0, aspect=aspect)
    return b.finalize(), cam


def env_scene(aspect=16.0 / 9.0, with_emitter=True):
    """Environment-lit scene (EnvMapSampler NEE + onMiss): open ground plane, a rough Au conductor sphere, a smooth
    Dielectric sphere (MNEE towards the distant env light), a coated sphere and a diffuse box under a procedural sky rotated by
    0.7 rad; optionally a small area emitter so that the light-type selection mixes env and emissive sampling."""
    b = SceneBuilder()
    au, cu = b.add_ior_profiles("Au"), b.add_ior_profiles("Cu")
    lamp = b.add_emission_profile(LAMP, 2.0)
    ground = b.add_diffuse((0.5, 0.5, 0.45))
    gold = b.add_conductor(au, roughness=0.2, gamma=3.0)
    glass = b.add_dielectric(ior=1.5, roughness=0.0)
    coated = b.add_coated_conductor(cu, 0.05, tau_um=2.0, coat_ior=1.5)
    red = b.add_diffuse((0.65, 0.1, 0.1))
    b.add_mesh(grid_plane(0.0, -6.0, 6.0, -6.0, 6.0, 4, 4), ground)
    b.add_mesh(uv_sphere([-1.3, 0.6, 0.0], 0.6, 32, 64), gold)
    b.add_mesh(uv_sphere([0.0, 0.6, 0.4], 0.6, 32, 64), glass)
    b.add_mesh(uv_sphere([1.3, 0.6, 0.0], 0.6, 32, 64), coated)
    b.add_mesh(box([0.0, 0.4, -1.6], [0.9, 0.4, 0.3], rot_y=0.2), red)
    if with_emitter:
        light = b.add_diffuse((0.0, 0.0, 0.0), emission_profile=lamp)
        b.add_mesh(quad([-0.4, 2.8, -0.4], [0.4, 2.8, -0.4], [0.4, 2.8, 0.4], [-0.4, 2.8, 0.4]), light)
    b.set_env_map(procedural_sky(), intensity=1.5, tint=(1.0, 0.95, 0.9), rot_y=0.7)
    cam = make_camera([0.0, 1.5, 4.5], [0.0, 0.6, 0.0], focal_length=30.0, aspect=aspect)
    return b.finalize(), cam


def analytic_lights_scene(aspect=16.0 / 9.0, kinds=("rect", "disc", "sphere", "distant")):
    """Scene lit by the analytic area lights of LightHelpers.slang:100-168 (rect, disc, sphere; plus one distant light and a small
    emissive quad so that the light-type selection mixes emissive and analytic sampling): diffuse room, a rough conductor and a
    smooth dielectric sphere (MNEE towards analytic lights)."""
    b = SceneBuilder()
    au = b.add_ior_profiles("Au")
    lamp = b.add_emission_profile(LAMP, 3.0)
    led = b.add_emission_profile("2499_LED_Encaster_Calculite_domestic-use_softwhite-color_Philips_brand_22W_1500lm_7inch_3000K", 1.0)
    white = b.add_diffuse((0.7, 0.7, 0.7))
    blue = b.add_diffuse((0.2, 0.3, 0.65))
    gold = b.add_conductor(au, roughness=0.25)
    glass = b.add_dielectric(ior=1.5, roughness=0.0)
    light = b.add_diffuse((0.0, 0.0, 0.0), emission_profile=lamp)
    b.add_mesh(grid_plane(0.0, -3.0, 3.0, -3.0, 3.0, 2, 2), white)
    b.add_mesh(quad([-3, 0, -2], [3, 0, -2], [3, 4, -2], [-3, 4, -2]), blue)
    b.add_mesh(uv_sphere([-0.9, 0.5, 0.2], 0.5, 24, 48), gold)
    b.add_mesh(uv_sphere([0.9, 0.5, 0.4], 0.5, 24, 48), glass)
    b.add_mesh(box([0.0, 0.3, -1.0], [0.5, 0.3, 0.3], rot_y=0.4), white)
    b.add_mesh(quad([-0.2, 3.0, -0.2], [0.2, 3.0, -0.2], [0.2, 3.0, 0.2], [-0.2, 3.0, 0.2]), light)
    if "rect" in kinds:
        b.add_rect_light(lamp, center=(-1.8, 2.2, 0.5), normal=(0.6, -1.0, -0.1), half_x=0.4, half_y=0.25)
    if "disc" in kinds:
        b.add_disc_light(led, center=(1.6, 2.4, 0.8), normal=(-0.5, -1.0, -0.2), radius_x=0.35, radius_y=0.2)
    if "sphere" in kinds:
        b.add_sphere_light(lamp, center=(0.0, 1.8, 1.2), radius=(0.15, 0.2, 0.1))
    if "distant" in kinds:
        b.add_distant_light(led, direction=(0.3, -1.0, -0.4), half_angle=math.radians(1.0))
    cam = make_camera([0.0, 1.5, 4.5], [0.0, 0.7, 0.0], focal_length=30.0, aspect=aspect)
    return b.finalize(), cam


def sweep_scene(dielectric_roughness=0.1, conductor_roughness=0.3):
    """BASELINE config 1 materials: 0 = rough Dielectric (IoR 1.5), 1 = rough Au Conductor (γ = 3); plus one of each
    remaining PLT material for the wider BSDF parity tests."""
    b = SceneBuilder()
    au, cu = b.add_ior_profiles("Au"), b.add_ior_profiles("Cu")
    lamp = b.add_emission_profile(LAMP)
    ids = dict(
        dielectric=b.add_dielectric(1.5, dielectric_roughness),
        conductor=b.add_conductor(au, conductor_roughness, 3.0),
        diffuse=b.add_diffuse((0.6, 0.4, 0.2)),
        opaque=b.add_opaque_dielectric((0.3, 0.5, 0.7), 0.05),
        thin=b.add_thin_dielectric(1.5, 0.3, birefringence=0.1),
        coated_conductor=b.add_coated_conductor(cu, 0.05, tau_um=3.0, coat_ior=1.5),
        coated_opaque=b.add_coated_opaque_dielectric((0.7, 0.2, 0.2), 0.05, tau_um=3.0, coat_ior=1.5),
        grated_conductor=b.add_grated_conductor(au, (3.0, 0.0), q=0.5 * 0.3, lobes=3),
        grated_opaque=b.add_grated_opaque_dielectric((0.5, 0.5, 0.5), (2.0, 2.0), q=0.2, gtype=1, lobes=4),
        stack=b.add_multilayered_stack((1, 1, 1), layers=6),
        smooth_dielectric=b.add_dielectric(1.5, 0.0, dispersion_b=0.01),
        smooth_conductor=b.add_conductor(au, 0.0),
        emitter=b.add_diffuse((0, 0, 0), emission_profile=lamp),
    )
    b.add_mesh(quad([-1, 0, 1], [1, 0, 1], [1, 0, -1], [-1, 0, -1]), ids["diffuse"])
    return b.finalize(), ids


def grating_sphere_scene(aspect=16.0 / 9.0, n_lat=64, n_lon=128):
    """BASELINE config 3: ground plane = sinusoidal diffraction-grated Au conductor (pitch 3 um, 3 lobes, q-scale 0.5),
    a tessellated smooth Dielectric sphere (MNEE refracts through it), a small area emitter and one distant light."""
    b = SceneBuilder()
    au = b.add_ior_profiles("Au")
    lamp = b.add_emission_profile(LAMP, 4.0)
    led = b.add_emission_profile("2499_LED_Encaster_Calculite_domestic-use_softwhite-color_Philips_brand_22W_1500lm_7inch_3000K", 0.02)
    grated = b.add_grated_conductor(au, (3.0, 0.0), q=0.3, q_scale=0.5, roughness=0.05, lobes=3)
    glass = b.add_dielectric(ior=1.5, roughness=0.0)
    wall = b.add_diffuse((0.6, 0.6, 0.6))
    light = b.add_diffuse((0.0, 0.0, 0.0), emission_profile=lamp)
    b.add_mesh(grid_plane(0.0, -3.0, 3.0, -3.0, 3.0, 8, 8), grated)
    b.add_mesh(uv_sphere([0.0, 0.75, 0.0], 0.6, n_lat, n_lon), glass)
    b.add_mesh(quad([-3, 0, -3], [3, 0, -3], [3, 4, -3], [-3, 4, -3]), wall)                                  # back wall (normal +z)
    b.add_mesh(quad([-0.25, 2.6, -0.25], [0.25, 2.6, -0.25], [0.25, 2.6, 0.25], [-0.25, 2.6, 0.25]), light)   # emitter facing down
    b.add_distant_light(led, direction=(-0.4, -1.0, -0.3), half_angle=math.radians(0.5))
    cam = make_camera([0.0, 1.6, 4.0], [0.0, 0.6, 0.0], focal_length=35.0, aspect=aspect)
    return b.finalize(), cam


def layered_stress_scene(aspect=16.0 / 9.0):
    """BASELINE config 4: coated conductors, coated opaque dielectric, birefringent thin dielectric and alternating
    multilayer stacks (2..20 layers) as spheres / plates over a diffuse floor, under one area emitter."""
    b = SceneBuilder()
    au, cu, al = b.add_ior_profiles("Au"), b.add_ior_profiles("Cu"), b.add_ior_profiles("Al")
    lamp = b.add_emission_profile(LAMP, 6.0)
    floor = b.add_diffuse((0.55, 0.55, 0.55))
    light = b.add_diffuse((0.0, 0.0, 0.0), emission_profile=lamp)
    mats = [
        b.add_coated_conductor(cu, 0.05, tau_um=3.0, coat_ior=1.5),
        b.add_coated_conductor(au, 0.15, tau_um=1.0, coat_ior=1.38),
        b.add_coated_opaque_dielectric((0.7, 0.2, 0.2), 0.05, tau_um=3.0, coat_ior=1.5),
        b.add_coated_opaque_dielectric((0.2, 0.3, 0.7), 0.1, tau_um=0.5, coat_ior=1.6),
        b.add_thin_dielectric(1.5, 0.3, birefringence=0.1),
        b.add_multilayered_stack((1, 1, 1), layers=2),
        b.add_multilayered_stack((1, 1, 1), layers=8),
        b.add_multilayered_stack((1, 1, 1), layers=20),
        b.add_conductor(al, 0.1),
        b.add_opaque_dielectric((0.3, 0.5, 0.7), 0.05),
    ]
    b.add_mesh(grid_plane(0.0, -4.0, 4.0, -3.0, 3.0, 4, 4), floor)
    b.add_mesh(quad([-4, 0, -3], [4, 0, -3], [4, 4, -3], [-4, 4, -3]), floor)
    k = 0
    for row in range(2):
        for col in range(8):
            m = mats[k % len(mats)]
            k += 1
            c = [-3.15 + col * 0.9, 0.36 + row * 0.0, -1.2 + row * 1.6]
            if m == mats[4]:  # thin dielectric as an upright plate
                b.add_mesh(quad([c[0] - 0.35, 0.02, c[2]], [c[0] + 0.35, 0.02, c[2]], [c[0] + 0.35, 0.9, c[2]], [c[0] - 0.35, 0.9, c[2]]), m)
            else:
                b.add_mesh(uv_sphere(c, 0.34, 24, 48), m)
    b.add_mesh(quad([-1.0, 3.2, -1.0], [1.0, 3.2, -1.0], [1.0, 3.2, 1.0], [-1.0, 3.2, 1.0]), light)
    cam = make_camera([0.0, 2.2, 5.5], [0.0, 0.4, 0.0], focal_length=35.0, aspect=aspect)
    return b.finalize(), cam


def _value_noise(X, Z, seed, octaves=5):
    """Seeded value noise (bilinear interpolation of a hashed lattice, summed over octaves), vectorised."""
    def hash01(ix, iz, k):
        h = (ix.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) ^ (iz.astype(np.uint64) + np.uint64(seed * 7919 + k)) * np.uint64(0xC2B2AE3D27D4EB4F))
        h ^= h >> np.uint64(29)
        h *= np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(32)
        return (h & np.uint64(0xFFFFFF)).astype(np.float64) / float(1 << 24)
    out = np.zeros_like(X, dtype=np.float64)
    amp, freq = 1.0, 1.0
    for k in range(octaves):
        x, z = X * freq, Z * freq
        ix, iz = np.floor(x).astype(np.int64), np.floor(z).astype(np.int64)
        fx, fz = x - ix, z - iz
        fx, fz = fx * fx * (3 - 2 * fx), fz * fz * (3 - 2 * fz)
        ix, iz = ix + (1 << 20), iz + (1 << 20)
        v00, v10, v01, v11 = hash01(ix, iz, k), hash01(ix + 1, iz, k), hash01(ix, iz + 1, k), hash01(ix + 1, iz + 1, k)
        out += amp * ((v00 * (1 - fx) + v10 * fx) * (1 - fz) + (v01 * (1 - fx) + v11 * fx) * fz)
        amp *= 0.5
        freq *= 2.0
    return out


def _boxes_mesh(centers, halves):
    """Many axis-aligned boxes merged into one mesh (24 vertices / 12 triangles each), vectorised."""
    c, h = np.asarray(centers, np.float64), np.asarray(halves, np.float64)
    sgn = np.array([[-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1],      # +z
                    [1, -1, -1], [-1, -1, -1], [-1, 1, -1], [1, 1, -1],  # -z
                    [1, -1, 1], [1, -1, -1], [1, 1, -1], [1, 1, 1],      # +x
                    [-1, -1, -1], [-1, -1, 1], [-1, 1, 1], [-1, 1, -1],  # -x
                    [-1, 1, 1], [1, 1, 1], [1, 1, -1], [-1, 1, -1],      # +y
                    [-1, -1, -1], [1, -1, -1], [1, -1, 1], [-1, -1, 1]], np.float64)  # -y
    nrm = np.repeat(np.array([[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0]], np.float64), 4, axis=0)
    P = (c[:, None, :] + sgn[None] * h[:, None, :]).reshape(-1, 3)
    N = np.tile(nrm, (len(c), 1))
    UV = np.tile(np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float64), (6 * len(c), 1))
    quad_idx = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    I = (np.arange(6 * len(c), dtype=np.uint32)[:, None] * 4 + quad_idx[None]).reshape(-1)
    return P.astype(f32), N.astype(f32), UV.astype(f32), I


def city_scene(target_triangles=10_000_000, aspect=16.0 / 9.0, seed=1):
    """BASELINE config 5: large synthetic scene with mixed PLT materials — a seeded value-noise terrain (80 % of the triangles),
    box "buildings" merged into one unique mesh per material (12 %) and an instanced tessellated sphere (8 %), lit by an
    area emitter and a distant light.  Material of a mesh = hash(mesh index) % 10 over the ten PLT types."""
    b = SceneBuilder()
    au, cu, al = b.add_ior_profiles("Au"), b.add_ior_profiles("Cu"), b.add_ior_profiles("Al")
    lamp = b.add_emission_profile(LAMP, 40.0)
    sun = b.add_emission_profile("2499_LED_Encaster_Calculite_domestic-use_softwhite-color_Philips_brand_22W_1500lm_7inch_3000K", 0.05)
    mats = [
        b.add_diffuse((0.55, 0.5, 0.45)),
        b.add_conductor(al, 0.2),
        b.add_dielectric(1.5, 0.1),
        b.add_opaque_dielectric((0.3, 0.5, 0.7), 0.1),
        b.add_thin_dielectric(1.5, 0.3, birefringence=0.1),
        b.add_coated_conductor(cu, 0.05, tau_um=3.0, coat_ior=1.5),
        b.add_coated_opaque_dielectric((0.7, 0.2, 0.2), 0.05, tau_um=3.0, coat_ior=1.5),
        b.add_grated_conductor(au, (3.0, 0.0), q=0.3, q_scale=0.5, roughness=0.05, lobes=3),
        b.add_grated_opaque_dielectric((0.5, 0.5, 0.5), (2.0, 2.0), q=0.2, gtype=1, lobes=4),
        b.add_multilayered_stack((1, 1, 1), layers=6),
    ]
    light = b.add_diffuse((0.0, 0.0, 0.0), emission_profile=lamp)
    rng = np.random.default_rng(seed)
    E = 60.0  # half extent of the terrain in metres
    n = max(4, int(math.sqrt(0.8 * target_triangles / 2)))
    height = lambda X, Z: 2.5 * _value_noise((X + E) * 0.05, (Z + E) * 0.05, seed) - 2.0
    b.add_mesh(grid_plane(0.0, -E, E, -E, E, n, n, height_fn=height), mats[0])
    n_boxes = max(10, int(0.12 * target_triangles / 12))
    centers = np.stack([rng.uniform(-E * 0.9, E * 0.9, n_boxes), np.zeros(n_boxes), rng.uniform(-E * 0.9, E * 0.9, n_boxes)], -1)
    halves = np.stack([rng.uniform(0.05, 0.4, n_boxes), rng.uniform(0.2, 3.0, n_boxes), rng.uniform(0.05, 0.4, n_boxes)], -1)
    centers[:, 1] = height(centers[:, 0], centers[:, 2]) + halves[:, 1] - 0.1
    group = (np.arange(n_boxes) * 2654435761 >> 7) % 10
    for g in range(10):
        sel = group == g
        if sel.any():
            b.add_mesh(_boxes_mesh(centers[sel], halves[sel]), mats[(g * 7 + 3) % 10])
    sphere = uv_sphere([0, 0, 0], 1.0, 48, 96) if target_triangles >= 2_000_000 else uv_sphere([0, 0, 0], 1.0, 12, 24)
    n_inst = max(1, int(0.08 * target_triangles / (len(sphere[3]) // 3)))
    first = None
    for k in range(n_inst):
        r = rng.uniform(0.3, 1.2)
        x, z = rng.uniform(-E * 0.8, E * 0.8, 2)
        y = float(height(np.array([x]), np.array([z]))[0]) + r + rng.uniform(0.0, 2.0)
        M = np.eye(4, dtype=f32)
        M[0, 0] = M[1, 1] = M[2, 2] = r
        M[:3, 3] = [x, y, z]
        m = mats[(k * 2654435761 >> 5) % 10]
        if first is None:
            first = b.add_mesh(sphere, m, M)
        else:
            b.add_instance(first, m, M)
    b.add_mesh(quad([-6.0, 25.0, -6.0], [6.0, 25.0, -6.0], [6.0, 25.0, 6.0], [-6.0, 25.0, 6.0]), light)
    b.add_distant_light(sun, direction=(-0.4, -1.0, -0.3), half_angle=math.radians(0.5))
    cam = make_camera([0.0, 14.0, 58.0], [0.0, 0.0, 0.0], focal_length=28.0, aspect=aspect)
    return b.finalize(), cam
