"""ctypes binding of oracle/_ref/libref_frames.so (TEST INFRASTRUCTURE ONLY): frames rendered by the reference's OWN shader text
(oracle/ref_frames.cpp over the transpile of oracle/ref_transpile.py).  Exists only where /root/reference lies -- it is the
generator of tests/golden/ref_frames.npz (tools/make_golden_ref.py), not something tests on the GPU box can call."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")
_libs = {}


def available():
    return os.path.isdir("/root/reference") or os.path.exists(os.path.join(_DIR, "libref_frames.so"))


def lib(hwss=4, debug=False, occluders=2):
    """The build of the transpiled pass for these shader defines: HWSS_SAMPLES (PLTPT.cpp:226), DO_DEBUG_VIEW (PLTPT.cpp:230), MNEE_MAX_MS_OCCLUDERS (:248)."""
    if (debug) + (hwss != 4) + (occluders != 2) > 1:
        raise ValueError("builds exist for one non-default define at a time")
    name = "libref_frames_dbg.so" if debug else (f"libref_frames_m{occluders}.so" if occluders != 2 else ("libref_frames.so" if hwss == 4 else f"libref_frames_h{hwss}.so"))
    if name not in _libs:
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-s", "-C", _HERE, "_ref/" + name])
        l = C.CDLL(os.path.join(_DIR, name))
        l.ref_scene_create.restype = C.c_void_p
        l.ref_scene_create.argtypes = [C.c_void_p]
        l.ref_scene_destroy.argtypes = [C.c_void_p]
        _libs[name] = l
    return _libs[name]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefScene:
    """The reference's shader text bound to host-side Falcor-layout buffers (pltfalcor_b200.scene.SceneBuffers)."""

    def __init__(self, buffers, debug=False, hwss=4, occluders=2):
        self.buffers = buffers
        self._desc = buffers.desc()
        self._lib = lib(hwss, debug, occluders)
        self.h = C.c_void_p(self._lib.ref_scene_create(C.byref(self._desc)))

    def __del__(self):
        if getattr(self, "h", None):
            self._lib.ref_scene_destroy(self.h)
            self.h = None

    def build_light_collection(self):
        """(emissiveTriangles, emissiveFlux) as the reference's BuildTriangleList / FinalizeIntegration passes build them for this scene's emissive instances."""
        from pltfalcor_b200 import layouts as L
        n = int(self._desc.emissiveTriangleCount)
        tris, flux = np.zeros(n, L.emissive_triangle_dtype), np.zeros(n, L.emissive_flux_dtype)
        if self._lib.ref_build_light_collection(self.h, _p(tris), _p(flux)) != 0:
            raise ValueError("perMeshInstanceOffset does not fit emissiveTriangleCount")
        return tris, flux

    def render_frame(self, opts, cam, frame, width, height):
        from pltfalcor_b200 import layouts as L
        color = np.zeros((height, width, 4), np.float32)
        albedo = np.zeros((height, width, 4), np.float32)
        normal = np.zeros((height, width, 4), np.float32)
        bounces = np.zeros((height, width, opts.maxBounces + 1), L.bounce_dtype)
        rc = self._lib.ref_render_frame(self.h, C.byref(opts), C.byref(cam), C.c_uint32(frame), C.c_uint32(width), C.c_uint32(height), _p(color), _p(albedo), _p(normal), _p(bounces))
        if rc != 0:
            raise ValueError("option fixed at compile time in the transpiled reference (MNEEMaxOccluders = 2; HWSS other than 4 needs RefScene(hwss=n), debug views RefScene(debug=True))")
        return dict(color=color, albedo=albedo, normal=normal, bounces=bounces)


def accumulate_frames(mode, frames, moving_average):
    """The reference's AccumulatePass (Accumulate.cs.slang) stepped over `frames` ((h, w, 4) float32 each): the output image after every frame."""
    h, w, _ = frames[0].shape
    last = np.zeros((h, w, 4), np.float64 if mode == 2 else np.float32)
    corr = np.zeros((h, w, 4), np.float32)
    outs = []
    for k, c in enumerate(frames):
        out = np.zeros((h, w, 4), np.float32)
        rc = lib().ref_accumulate_frame(C.c_uint32(mode), C.c_uint32(k), C.c_int(int(moving_average)), C.c_uint32(w), C.c_uint32(h), _p(np.ascontiguousarray(c, np.float32)), _p(last), _p(corr), _p(out))
        assert rc == 0
        outs.append(out)
    return outs


_spectrum = None


def spectrum_lib():
    """oracle/_ref/libref_spectrum.so: the reference's own C++ (Spectrum.cpp, SpectrumUtils.cpp, SampledSpectrum.h, genSpectralProfile), compiled as it lies."""
    global _spectrum
    if _spectrum is None:
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-s", "-C", _HERE, "_ref/libref_spectrum.so"])
        _spectrum = C.CDLL(os.path.join(_DIR, "libref_spectrum.so"))
    return _spectrum


def profile_from_spd_file(path, scale=1.0, compute_icdf=True):
    """SpectralProfile (the C ABI's 1052-byte record) as the reference's scene builder makes it from a table on disk: PiecewiseLinearSpectrum::fromFile ->
    scale -> SampledSpectrum -> genSpectralProfile, run in oracle/_ref/ref_spectrum_tool (a C++ main around the same code: the reader faults inside the
    interpreter process)."""
    from pltfalcor_b200 import layouts as L
    tool = os.path.join(_DIR, "ref_spectrum_tool")
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_ref/ref_spectrum_tool"])
    raw = subprocess.run([tool, str(path), repr(float(scale)), str(int(compute_icdf))], capture_output=True, check=True).stdout
    return np.frombuffer(raw, L.spectral_profile_dtype)[0]


def gen_spectral_profile(wavelengths, values, scale=1.0, compute_icdf=True):
    from pltfalcor_b200 import layouts as L
    w, v = np.ascontiguousarray(wavelengths, np.float32), np.ascontiguousarray(values, np.float32)
    out = np.zeros((), L.spectral_profile_dtype)
    if spectrum_lib().ref_gen_spectral_profile(C.c_uint32(len(w)), _p(w), _p(v), C.c_float(scale), C.c_int(int(compute_icdf)), out.ctypes.data_as(C.c_void_p)) != 0:
        raise ValueError("the reference rejected the spectrum")
    return out
