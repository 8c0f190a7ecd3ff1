"""ctypes binding of libpltpt.so (include/pltpt.h).  Used by tests/, bench.py and __graft_entry__: the product is the
C-ABI library, this file only moves pointers.  There is no fallback: a missing library or a failing call raises."""
import ctypes as C
import os

import numpy as np

from . import layouts as L
from .build import LIB

LIB = os.environ.get("PLTPT_LIB") or LIB  # development: a variant built by tools/variants.py

_lib = None

# every symbol include/pltpt.h declares (tests check that the library exports all of them)
EXPORTS = [
    "pltpt_default_options", "pltpt_create", "pltpt_destroy", "pltpt_last_error", "pltpt_version", "pltpt_set_options", "pltpt_set_scene",
    "pltpt_render_frame", "pltpt_render_accumulate", "pltpt_render_frame_host", "pltpt_sample_stage_host", "pltpt_trace_closest", "pltpt_trace_any",
    "pltpt_bsdf_eval_mueller", "pltpt_bsdf_sample", "pltpt_get_stats", "pltpt_reset_stats", "pltpt_set_profiling",
    "pltpt_envmap_eval", "pltpt_envmap_sample", "pltpt_set_frame_overlap",
    "pltpt_spd_load", "pltpt_gen_spectral_profile", "pltpt_pack_material", "pltpt_build_light_collection", "pltpt_get_light_collection", "pltpt_update_transforms", "pltpt_set_bvh_update_mode", "pltpt_set_work_budget",
    "pltpt_render_frame_host_async", "pltpt_host_wait", "pltpt_set_trace_mode",
    "pltpt_render_accumulate_rows", "pltpt_resolve_accumulated", "pltpt_render_accumulate_host", "pltpt_render_accumulate_multi",
    "pltpt_texture_sample", "pltpt_accumulate_frame", "pltpt_bsdf_eval_mueller_replay_math",
]


class PltptError(RuntimeError):
    """Non-zero PltptStatus (the Falcor shell throws Falcor::RuntimeError for the same codes, PLTPT.cpp:409)."""

    def __init__(self, code, msg):
        super().__init__(f"pltpt status {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} is missing: build it with `python -m pltfalcor_b200.build` (there is no CPU fallback for the PLTPT path)")
        _lib = C.CDLL(LIB)
        _lib.pltpt_last_error.restype = C.c_char_p
        _lib.pltpt_version.restype = C.c_char_p
        vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
        _lib.pltpt_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
        _lib.pltpt_destroy.argtypes = [vp]
        _lib.pltpt_set_options.argtypes = [vp, vp]
        _lib.pltpt_set_scene.argtypes = [vp, vp]
        _lib.pltpt_render_frame.argtypes = [vp, u32, vp, u32, u32, vp, vp, vp, vp]
        _lib.pltpt_render_accumulate.argtypes = [vp, u32, u32, u32, vp, u32, u32, vp, vp]
        _lib.pltpt_render_accumulate_rows.argtypes = [vp, u32, u32, u32, vp, u32, u32, u32, u32, vp, vp]
        _lib.pltpt_resolve_accumulated.argtypes = [vp, vp, u32, u32, vp, vp]
        _lib.pltpt_render_accumulate_host.argtypes = [vp, u32, u32, u32, vp, u32, u32, vp]
        _lib.pltpt_render_accumulate_multi.argtypes = [vp, u32, u32, u32, vp, u32, u32, u32, vp, vp]
        _lib.pltpt_render_frame_host.argtypes = [vp, u32, vp, u32, u32, vp, vp, vp]
        _lib.pltpt_render_frame_host_async.argtypes = [vp, u32, vp, u32, u32, vp, vp, vp]
        _lib.pltpt_host_wait.argtypes = [vp, u32]
        _lib.pltpt_sample_stage_host.argtypes = [vp, u32, vp, u32, u32, vp]
        _lib.pltpt_trace_closest.argtypes = [vp, u32, vp, vp, vp, vp]
        _lib.pltpt_trace_any.argtypes = [vp, u32, vp, vp, vp]
        _lib.pltpt_set_trace_mode.argtypes = [vp, C.c_int]
        _lib.pltpt_bsdf_eval_mueller.argtypes = [vp, u32, u32, u32, u32, vp, vp, vp, vp, vp]
        _lib.pltpt_bsdf_eval_mueller_replay_math.argtypes = [vp, u32, u32, u32, u32, vp, vp, vp, vp, vp]
        _lib.pltpt_bsdf_sample.argtypes = [vp, u32, u32, u32, vp, vp, u32, vp, vp]
        _lib.pltpt_get_stats.argtypes = [vp, vp]
        _lib.pltpt_reset_stats.argtypes = [vp]
        _lib.pltpt_set_profiling.argtypes = [vp, C.c_int]
        _lib.pltpt_set_frame_overlap.argtypes = [vp, C.c_int]
        _lib.pltpt_set_work_budget.argtypes = [vp, C.c_uint64]
        _lib.pltpt_spd_load.argtypes = [C.c_char_p, vp, vp, u32, vp]
        _lib.pltpt_gen_spectral_profile.argtypes = [u32, vp, vp, f32, C.c_int, vp]
        _lib.pltpt_pack_material.argtypes = [vp, vp]
        _lib.pltpt_build_light_collection.argtypes = [vp]
        _lib.pltpt_update_transforms.argtypes = [vp, vp, vp, u32, u32]
        _lib.pltpt_set_bvh_update_mode.argtypes = [vp, C.c_int]
        _lib.pltpt_get_light_collection.argtypes = [vp, vp, vp, vp, vp, vp, u32, vp]
        _lib.pltpt_envmap_eval.argtypes = [vp, u32, vp, vp, vp]
        _lib.pltpt_envmap_sample.argtypes = [vp, u32, vp, vp, vp]
        _lib.pltpt_accumulate_frame.argtypes = [vp, u32, u32, C.c_int, u32, u32, vp, vp, vp, vp, vp]
        _lib.pltpt_texture_sample.argtypes = [vp, u32, u32, vp, u32, vp, vp, vp]
    return _lib


def render_accumulate_multi(handles, cam, first_frame, frame_count, width, height, partition=0, color_host=None, accum_out=None):
    """Single-process multi-GPU accumulation (pltpt_render_accumulate_multi): `handles` = Pltpt objects on distinct devices with the same scene."""
    arr = (C.c_void_p * len(handles))(*[h.h for h in handles])
    ch = None if color_host is None else (color_host.data_ptr() if hasattr(color_host, "data_ptr") else _hp(color_host))
    _check(lib().pltpt_render_accumulate_multi(arr, len(handles), first_frame, frame_count, C.byref(cam), width, height, partition, ch, _dp(accum_out)))


def _check(code):
    if code != 0:
        raise PltptError(code, lib().pltpt_last_error().decode())


def _hp(a):
    return None if a is None else a.ctypes.data


def _dp(t):
    return None if t is None else t.data_ptr()


# ---- scene ingestion without Falcor (host-side entry points; no handle, no GPU) -------------------------------------------------
def spd_load(path):
    n = C.c_uint32()
    _check(lib().pltpt_spd_load(str(path).encode(), None, None, 0, C.byref(n)))
    w, v = np.zeros(n.value, np.float32), np.zeros(n.value, np.float32)
    _check(lib().pltpt_spd_load(str(path).encode(), _hp(w), _hp(v), n.value, C.byref(n)))
    return w, v


def gen_spectral_profile(wavelengths, values, compute_icdf=True, scale=1.0):
    w, v = np.ascontiguousarray(wavelengths, np.float32), np.ascontiguousarray(values, np.float32)
    out = np.zeros((), L.spectral_profile_dtype)
    _check(lib().pltpt_gen_spectral_profile(len(w), _hp(w), _hp(v), scale, int(compute_icdf), out.ctypes.data))
    return out


def pack_material(**params):
    p = L.PltptMaterialParams()
    p.emissionSpectralId = p.iorNSpectralId = p.iorKSpectralId = L.NO_SPECTRAL_PROFILE
    p.color[:] = [1.0, 1.0, 1.0]
    p.transmission[:] = [1.0, 1.0, 1.0]
    p.extEta, p.ior, p.gamma, p.gratingQScale, p.gratingPowerMultiplier, p.birefringenceScale = 1.0, 1.5, 3.0, 1.0, 1.0, 1.0
    for k, v in params.items():
        cur = getattr(p, k)
        if hasattr(cur, "__len__"):
            cur[:] = [float(x) for x in v]
        else:
            setattr(p, k, type(cur)(v))
    out = np.zeros((), L.material_blob_dtype)
    _check(lib().pltpt_pack_material(C.byref(p), out.ctypes.data))
    return out


class Pltpt:
    """One PltptHandle. `options` are the 24 PLTPT pass properties (PLTPT.cpp:102-136)."""

    def __init__(self, device=0, **options):
        self.h = C.c_void_p()
        self.options = L.make_options(**options)
        _check(lib().pltpt_create(C.byref(self.options), int(device), C.byref(self.h)))
        self.device = device
        self._scene_keep = None

    def close(self):
        if getattr(self, "h", None):
            lib().pltpt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, **options):
        vals = {k: getattr(self.options, k) for k, _ in L.PltptOptions._fields_}
        vals.update(options)
        self.options = L.make_options(**vals)
        _check(lib().pltpt_set_options(self.h, C.byref(self.options)))

    def set_scene(self, buffers, device_tensors=None):
        """buffers: scene.SceneBuffers (host numpy). If device_tensors = buffers.to_device(torch) is given the
        interop path (device pointers) is exercised instead."""
        desc = device_tensors[1] if device_tensors else buffers.desc()
        self._scene_keep = (buffers, device_tensors, desc)
        _check(lib().pltpt_set_scene(self.h, C.byref(desc)))

    def stats(self):
        s = L.PltptStats()
        _check(lib().pltpt_get_stats(self.h, C.byref(s)))
        out = {k: getattr(s, k) for k, _ in L.PltptStats._fields_}
        out["kernelMs"] = dict(zip(L.KERNEL_CLASSES, list(s.kernelMs)))
        out["kernelLaunchCount"] = dict(zip(L.KERNEL_CLASSES, list(s.kernelLaunchCount)))
        return out

    def set_profiling(self, enable):
        """Per-kernel device timers (one CUDA event per launch) on/off."""
        _check(lib().pltpt_set_profiling(self.h, 1 if enable else 0))

    def set_frame_overlap(self, frames):
        """Frames render_accumulate keeps in flight: 0 = default (4), 1 .. 4."""
        _check(lib().pltpt_set_frame_overlap(self.h, int(frames)))

    def set_work_budget(self, nbytes):
        """Per-lane work-buffer budget (bytes; 0 = default): smaller budgets render a frame in more row tiles."""
        _check(lib().pltpt_set_work_budget(self.h, int(nbytes)))

    def reset_stats(self):
        _check(lib().pltpt_reset_stats(self.h))

    # ---- rendering -------------------------------------------------------------------------------------------
    def render_frame_host(self, cam, frame, width, height, want_guides=True):
        color = np.zeros((height, width, 4), np.float32)
        albedo = np.zeros((height, width, 4), np.float32) if want_guides else None
        normal = np.zeros((height, width, 4), np.float32) if want_guides else None
        _check(lib().pltpt_render_frame_host(self.h, frame, C.byref(cam), width, height, _hp(color), _hp(albedo), _hp(normal)))
        return dict(color=color, albedo=albedo, normal=normal)

    def render_frames_host_queued(self, cams_and_frames, width, height):
        """pltpt_render_frame_host_async for every (camera, frame) pair, then pltpt_host_wait; returns the color images."""
        out = [np.zeros((height, width, 4), np.float32) for _ in cams_and_frames]
        for (cam, frame), color in zip(cams_and_frames, out):
            _check(lib().pltpt_render_frame_host_async(self.h, frame, C.byref(cam), width, height, _hp(color), None, None))
        _check(lib().pltpt_host_wait(self.h, 0))
        return out

    def sample_stage_host(self, cam, frame, width, height):
        bounces = np.zeros((height, width, self.options.maxBounces + 1), L.bounce_dtype)
        _check(lib().pltpt_sample_stage_host(self.h, frame, C.byref(cam), width, height, _hp(bounces)))
        return bounces

    def render_frame(self, cam, frame, width, height, color, albedo=None, normal=None, stream=None):
        """color/albedo/normal: torch float32 CUDA tensors of width*height*4."""
        _check(lib().pltpt_render_frame(self.h, frame, C.byref(cam), width, height, _dp(color), _dp(albedo), _dp(normal), stream))

    def render_accumulate(self, cam, first_frame, frame_step, frame_count, width, height, accum, stream=None):
        """accum: torch float64 CUDA tensor of width*height*4 (RGB sums + sample count)."""
        _check(lib().pltpt_render_accumulate(self.h, first_frame, frame_step, frame_count, C.byref(cam), width, height, _dp(accum), stream))

    def render_accumulate_rows(self, cam, first_frame, frame_step, frame_count, width, height, row_begin, row_end, accum, stream=None):
        _check(lib().pltpt_render_accumulate_rows(self.h, first_frame, frame_step, frame_count, C.byref(cam), width, height, row_begin, row_end, _dp(accum), stream))

    def resolve_accumulated(self, accum, width, height, color, stream=None):
        _check(lib().pltpt_resolve_accumulated(self.h, _dp(accum), width, height, _dp(color), stream))

    def render_accumulate_host(self, cam, first_frame, frame_step, frame_count, width, height, color_host):
        """color_host: pinned torch float32 CPU tensor (h, w, 4) or a numpy array."""
        ptr = color_host.data_ptr() if hasattr(color_host, "data_ptr") else _hp(color_host)
        _check(lib().pltpt_render_accumulate_host(self.h, first_frame, frame_step, frame_count, C.byref(cam), width, height, ptr))

    # ---- ray queries ------------------------------------------------------------------------------------------
    def trace_closest(self, rays, hits, hit_t=None, stream=None):
        _check(lib().pltpt_trace_closest(self.h, rays.shape[0], _dp(rays), _dp(hits), _dp(hit_t), stream))

    def trace_any(self, rays, occluded, stream=None):
        _check(lib().pltpt_trace_any(self.h, rays.shape[0], _dp(rays), _dp(occluded), stream))

    def set_trace_mode(self, mode):
        """0 = follow the renderer's choice, 1 = one ray per thread, 2 = persistent lanes (include/pltpt.h PLTPT_TRACE_*)."""
        _check(lib().pltpt_set_trace_mode(self.h, mode))

    def update_transforms(self, world, world_it):
        """world, world_it: (n,4,4) float32 host arrays (row-major), like SceneBuffers.arrays['worldMatrices']."""
        world, world_it = np.ascontiguousarray(world, np.float32), np.ascontiguousarray(world_it, np.float32)
        _check(lib().pltpt_update_transforms(self.h, _hp(world), _hp(world_it), len(world), 0))

    def set_bvh_update_mode(self, mode):
        """0 = rebuild (default), 1 = refit (include/pltpt.h PltptBvhUpdateMode; Scene::UpdateMode, Scene.h:169-173)."""
        _check(lib().pltpt_set_bvh_update_mode(self.h, int(mode)))

    # ---- light collection built by the library -------------------------------------------------------------------
    def build_light_collection(self):
        _check(lib().pltpt_build_light_collection(self.h))

    def get_light_collection(self, instance_count):
        n, inv = C.c_uint32(), C.c_float()
        _check(lib().pltpt_get_light_collection(self.h, None, None, None, None, None, 0, C.byref(n)))
        tris = np.zeros(n.value, L.emissive_triangle_dtype)
        flux = np.zeros(n.value, L.emissive_flux_dtype)
        offs = np.zeros(instance_count, np.uint32)
        alias = np.zeros((n.value, 2), np.uint32)
        _check(lib().pltpt_get_light_collection(self.h, _hp(tris), _hp(flux), _hp(offs), _hp(alias), C.byref(inv), n.value, C.byref(n)))
        return dict(emissiveTriangles=tris, emissiveFlux=flux, perMeshInstanceOffset=offs, triangleAliasTable=alias, invWeightsSum=inv.value)

    # ---- environment map ------------------------------------------------------------------------------------
    def envmap_eval(self, dirs, out, stream=None):
        """dirs: (n,3) world directions; out: (n,4) {radiance rgb, pdf} — torch CUDA float32 tensors."""
        _check(lib().pltpt_envmap_eval(self.h, dirs.shape[0], _dp(dirs), _dp(out), stream))

    def envmap_sample(self, rnd, out, stream=None):
        """rnd: (n,2) in [0,1); out: (n,8) {dir xyz, pdf, Le rgb, 0}."""
        _check(lib().pltpt_envmap_sample(self.h, rnd.shape[0], _dp(rnd), _dp(out), stream))

    def accumulate_frame(self, precision, accum_count, cur, last_sum, out, last_corr=None, moving_average=False, stream=None):
        """AccumulatePass, one frame (precision 0 Single, 1 SingleCompensated, 2 Double); torch CUDA tensors, cur / out (H, W, 4) float32."""
        h, w = cur.shape[:2]
        _check(lib().pltpt_accumulate_frame(self.h, precision, accum_count, int(moving_average), w, h, _dp(cur), _dp(last_sum), _dp(last_corr) if last_corr is not None else None, _dp(out), stream))

    def texture_sample(self, handle, sampler_id, uv, out, uniform=(0.0, 0.0, 0.0, 0.0), stream=None):
        """uv: (n,2); out: (n,4) — torch CUDA float32 tensors; `handle` is a TextureHandle word (mode << 29 | textureID)."""
        u = (C.c_float * 4)(*[float(x) for x in uniform])
        _check(lib().pltpt_texture_sample(self.h, handle, sampler_id, C.cast(u, C.c_void_p), uv.shape[0], _dp(uv), _dp(out), stream))

    # ---- BSDF sweeps ------------------------------------------------------------------------------------------
    def bsdf_eval_mueller(self, material_id, lobe, wi, wo, hero, out, stream=None):
        """wi, wo: (n,3); hero: (groups,n); out: (n,groups,HWSS,4,4) — torch CUDA float32 tensors."""
        groups, n = hero.shape
        _check(lib().pltpt_bsdf_eval_mueller(self.h, material_id, lobe, n, groups, _dp(wi), _dp(wo), _dp(hero), _dp(out), stream))

    def bsdf_eval_mueller_replay_math(self, material_id, lobe, wi, wo, hero, out, stream=None):
        """bsdf_eval_mueller with the forward replay's arithmetic (test hook)."""
        groups, n = hero.shape
        _check(lib().pltpt_bsdf_eval_mueller_replay_math(self.h, material_id, lobe, n, groups, _dp(wi), _dp(wo), _dp(hero), _dp(out), stream))

    def bsdf_sample(self, material_id, lobe_mask, wi, wavelengths, seed, out, stream=None):
        _check(lib().pltpt_bsdf_sample(self.h, material_id, lobe_mask, wi.shape[0], _dp(wi), _dp(wavelengths), seed, _dp(out), stream))
