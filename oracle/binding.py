"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY — see oracle/plt_math.h).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".h", ".cpp"))] + [os.path.join(_HERE, "..", "include", "pltpt.h")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_scene_create.restype = C.c_void_p
        _lib.oracle_scene_create.argtypes = [C.c_void_p]
        _lib.oracle_scene_destroy.argtypes = [C.c_void_p]
        _lib.oracle_scene_triangle_count.argtypes = [C.c_void_p]
        _lib.oracle_render_frame.restype = C.c_uint64
        _lib.oracle_sampled_spectrum_eval.restype = C.c_float
        _lib.oracle_eval_spectrum.restype = C.c_float
        _lib.oracle_count_ties.restype = C.c_uint32
        _lib.oracle_count_ties.argtypes = [C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleScene:
    """Wraps an orc::Scene built from host-side Falcor-layout buffers (pltfalcor_b200.scene.SceneBuffers)."""

    def __init__(self, buffers):
        self.buffers = buffers
        self._desc = buffers.desc()
        self.h = C.c_void_p(lib().oracle_scene_create(C.byref(self._desc)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_scene_destroy(self.h)
            self.h = None

    def trace_closest(self, rays, brute=False, threads=8):
        rays = np.ascontiguousarray(rays, np.float32)
        n = len(rays)
        hits = np.zeros((n, 4), np.uint32)
        t = np.zeros(n, np.float32)
        lib().oracle_trace_closest(self.h, C.c_uint32(n), _p(rays), _p(hits), _p(t), C.c_int(int(brute)), C.c_int(threads))
        return hits, t

    def trace_any(self, rays, brute=False, threads=8):
        rays = np.ascontiguousarray(rays, np.float32)
        n = len(rays)
        occ = np.zeros(n, np.uint32)
        lib().oracle_trace_any(self.h, C.c_uint32(n), _p(rays), _p(occ), C.c_int(int(brute)), C.c_int(threads))
        return occ

    def count_ties(self, ray):
        ray = np.ascontiguousarray(ray, np.float32)
        return int(lib().oracle_count_ties(self.h, _p(ray)))

    def bsdf_eval_mueller(self, material_id, lobe, wi, wo, hero, hwss=4, max_beam_omega=0.025, threads=8):
        wi = np.ascontiguousarray(wi, np.float32)
        wo = np.ascontiguousarray(wo, np.float32)
        hero = np.ascontiguousarray(hero, np.float32)  # (groups, n)
        groups, n = hero.shape
        out = np.zeros((n, groups, hwss, 4, 4), np.float32)
        lib().oracle_bsdf_eval_mueller(self.h, C.c_uint32(material_id), C.c_uint32(lobe), C.c_uint32(n), C.c_uint32(groups), _p(wi), _p(wo), _p(hero), _p(out),
                                       C.c_int(hwss), C.c_float(max_beam_omega), C.c_int(threads))
        return out

    def bsdf_sample(self, material_id, lobe_mask, wi, wavelengths, seed, hwss=4, threads=8):
        wi = np.ascontiguousarray(wi, np.float32)
        wavelengths = np.ascontiguousarray(wavelengths, np.float32)
        n = len(wi)
        out = np.zeros((n, 8), np.float32)
        lib().oracle_bsdf_sample(self.h, C.c_uint32(material_id), C.c_uint32(lobe_mask), C.c_uint32(n), _p(wi), _p(wavelengths), C.c_uint32(seed), _p(out), C.c_int(hwss), C.c_int(threads))
        return out

    def envmap_eval(self, dirs):
        dirs = np.ascontiguousarray(dirs, np.float32)
        out = np.zeros((len(dirs), 4), np.float32)
        lib().oracle_envmap_eval(self.h, C.c_uint32(len(dirs)), _p(dirs), _p(out))
        return out

    def envmap_sample(self, rnd):
        rnd = np.ascontiguousarray(rnd, np.float32)
        out = np.zeros((len(rnd), 8), np.float32)
        lib().oracle_envmap_sample(self.h, C.c_uint32(len(rnd)), _p(rnd), _p(out))
        return out

    def texture_sample(self, handle, sampler_id, uv, uniform=(0, 0, 0, 0)):
        uv = np.ascontiguousarray(uv, np.float32)
        out = np.zeros((len(uv), 4), np.float32)
        u = np.asarray(uniform, np.float32)
        lib().oracle_texture_sample(self.h, C.c_uint32(handle), C.c_uint32(sampler_id), _p(u), C.c_uint32(len(uv)), _p(uv), _p(out))
        return out

    def envmap_importance(self):
        out = np.zeros((512, 512), np.float32)
        lib().oracle_envmap_importance(self.h, _p(out))
        return out

    def render_frame(self, opts, cam, frame, width, height, solve=True, want_bounces=False, threads=8):
        from pltfalcor_b200 import layouts as L
        color = np.zeros((height, width, 4), np.float32)
        albedo = np.zeros((height, width, 4), np.float32)
        normal = np.zeros((height, width, 4), np.float32)
        bounces = np.zeros((height, width, opts.maxBounces + 1), L.bounce_dtype) if want_bounces else None
        evals = lib().oracle_render_frame(self.h, C.byref(opts), C.byref(cam), C.c_uint32(frame), C.c_uint32(width), C.c_uint32(height),
                                          _p(color), _p(albedo), _p(normal), _p(bounces), C.c_int(int(solve)), C.c_int(threads))
        return dict(color=color, albedo=albedo, normal=normal, bounces=bounces, evals=int(evals))


def hardware_threads():
    return int(lib().oracle_hardware_threads())
