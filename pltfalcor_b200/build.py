"""Builds libpltpt.so (the C-ABI library of include/pltpt.h) in-tree with nvcc for sm_100a. No JIT cache, no torch extension."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libpltpt.so")
SOURCES = ["api.cu", "bvh.cu", "integrator.cu", "stage_sample.cu", "stage_connect.cu", "stage_unwind.cu", "ingest.cu"]
# per-source flags.  api.cu holds the BSDF sweep kernel: there the wavelength loops are unrolled (PLT_FOR_LAMBDA, plt_math.cuh) -- the four
# Stokes columns of its beam stay in registers and the sweep runs 1.45x faster; in the rendering kernels the same unrolling costs
# 6 % on the replay step (profiles/r2f_variants.log), so they keep the runtime-bounded loops.
# stage_unwind.cu (the forward replay) evaluates BSDFs along paths that are already fixed: nothing it computes feeds a ray or a branch of
# the path, so it is held to the 1e-4 BSDF tolerance, not to the operation-by-operation rounding the other stages keep for identical
# paths.  There division and square root are the MUFU approximations (<= 2 ulp), FMAs are contracted, and PLT_FAST_MATH selects the
# rsqrt / exp2-log2 / rotation-without-atan2 forms in plt_math.cuh and plt_bsdf.cuh: k_unwind_step 1.54 -> 1.08 ms per Cornell frame with
# every parity figure unchanged (profiles/r2h_variants.log, r2i_variants.log).  The same flags on stage_connect.cu move light-sampling
# decisions and were rejected (profiles/r2h_variants.log: fm_uc).
SOURCE_FLAGS = {"api.cu": ["-DPLT_LAMBDA_UNROLL=1"],
                "stage_unwind.cu": ["-prec-div=false", "-prec-sqrt=false", "-fmad=true", "-DPLT_FAST_MATH=1"]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _deps():
    out = [os.path.join(_HERE, "..", "include", "pltpt.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False):
    """Compiles csrc/*.cu -> libpltpt.so. Returns the library path."""
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(_HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        deps = _deps()
        if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
            return obj
        extra = SOURCE_FLAGS.get(src, [])
        cmd = [nvcc, *[f for f in NVCC_FLAGS if not (f == "-fmad=false" and "-fmad=true" in extra)], *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-lcudart_static", "-Xlinker", "--no-undefined"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


HOST = os.path.join(_HERE, "host")
PLUGIN = os.path.join(_HERE, "libPLTPT.so")       # the drop-in render-pass plugin (host/PLTPT.cpp), exports registerPlugin
HOST_EXE = os.path.join(_HERE, "pltpt_host")      # Mogwai stand-in that loads the plugin (host/pltpt_host.cpp)


def _cuda_home():
    nv = _nvcc()
    return os.path.dirname(os.path.dirname(nv)) if os.path.isabs(nv) else "/usr/local/cuda"


def build_host(force=False):
    """Compiles the C++ host shell: libPLTPT.so (plugin, links libpltpt.so) and pltpt_host.  g++ only; no GPU code here."""
    srcs = [os.path.join(HOST, f) for f in os.listdir(HOST)] + [os.path.join(_HERE, "..", "include", "pltpt.h"), LIB]
    stale = lambda out: force or not os.path.exists(out) or any(os.path.getmtime(x) > os.path.getmtime(out) for x in srcs)
    cuda = _cuda_home()
    common = ["-std=c++17", "-O2", "-Wall", "-I" + os.path.join(_HERE, "..", "include"), "-I" + os.path.join(cuda, "include")]
    libs = ["-L" + os.path.join(cuda, "lib64"), "-lcudart_static", "-ldl", "-lpthread", "-lrt"]
    cmds = []
    if stale(PLUGIN):
        cmds.append(["g++", *common, "-fPIC", "-shared", os.path.join(HOST, "PLTPT.cpp"), "-o", PLUGIN, "-L" + _HERE, "-lpltpt", *libs, "-Wl,-rpath,$ORIGIN"])
    if stale(HOST_EXE):
        cmds.append(["g++", *common, "-rdynamic", os.path.join(HOST, "pltpt_host.cpp"), "-o", HOST_EXE, *libs])
    for cmd in cmds:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host build failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return PLUGIN, HOST_EXE


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
