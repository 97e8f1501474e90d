"""In-tree build of the two shared libraries (explicit nvcc / g++; no JIT cache).

  lib/librtb_host.so   host-side producers (scene, BVH, view) — plain C++
  lib/librtb.so        the sm_100a plugin: CUDA kernels + the C ABI of include/rtb.h

nvcc cross-compiles sm_100a without a GPU, so this runs in the authoring container; the
built .so files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib")
INCLUDE = os.path.join(ROOT, "include")

HOST_SOURCES = [os.path.join(CSRC, "host", "rtb_host.cpp")]
CUDA_SOURCES = [os.path.join(CSRC, "plugin.cu")]
FAST_SOURCES = [os.path.join(CSRC, "fast_kernels.cu")]       # the opt-in fast-arithmetic build of the sphere megakernel
CUDA_DEPS = [
    os.path.join(CSRC, f)
    for f in ("kernel_common.cuh", "media.cuh", "sample_kernels.cuh", "volume_kernel.cuh", "aux_kernels.cuh", "retree.hpp")
] + [os.path.join(INCLUDE, "rtb.h"), os.path.join(INCLUDE, "rtb", "umath.h")]
HOST_DEPS = [os.path.join(INCLUDE, "rtb_host.h"), os.path.join(INCLUDE, "rtb.h"), os.path.join(INCLUDE, "rtb", "umath.h")]

NVCC_COMMON = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-I", INCLUDE, "-I", CSRC,
]
# -fmad=false: the parity contract of include/rtb/umath.h (FMAs only where written).  fast_kernels.cu is the one
# translation unit compiled with contraction on (RTB_OPT_MATH = 1).
NVCC_PARITY = ["-fmad=false"]
NVCC_FAST = ["-fmad=true"]
NVCC_LINK = ["-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "static"]
GXX_FLAGS = [
    "-std=c++17", "-O2", "-ffp-contract=off", "-march=x86-64-v3",
    "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-Wextra", "-pthread",
    "-I", INCLUDE,
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr, flush=True)


def host_lib_path():
    return os.path.join(LIB, "librtb_host.so")


def plugin_lib_path():
    return os.path.join(LIB, "librtb.so")


def build_host(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    out = host_lib_path()
    if force or _stale(out, HOST_SOURCES + HOST_DEPS):
        gxx = shutil.which("g++") or "g++"
        _run([gxx] + GXX_FLAGS + ["-o", out] + HOST_SOURCES, verbose)
    return out


def _compile_and_link(out, objdir, extra_flags, verbose):
    """plugin.cu (parity flags) and fast_kernels.cu (fast flags) -> objects, in parallel, then one shared library."""
    from concurrent.futures import ThreadPoolExecutor

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(objdir, exist_ok=True)
    jobs = [(CUDA_SOURCES[0], NVCC_PARITY, os.path.join(objdir, "plugin.o")),
            (FAST_SOURCES[0], NVCC_FAST, os.path.join(objdir, "fast_kernels.o"))]
    with ThreadPoolExecutor(2) as ex:
        list(ex.map(lambda j: _run([nvcc] + NVCC_COMMON + j[1] + list(extra_flags) + ["-c", j[0], "-o", j[2]], verbose), jobs))
    _run([nvcc] + NVCC_LINK + ["-o", out] + [j[2] for j in jobs], verbose)


def build_plugin(force=False, verbose=False, extra_flags=()):
    os.makedirs(LIB, exist_ok=True)
    out = plugin_lib_path()
    if force or _stale(out, CUDA_SOURCES + FAST_SOURCES + CUDA_DEPS):
        _compile_and_link(out, os.path.join(PKG, "build"), extra_flags, verbose)
    return out


def build_variant(tag, defines, verbose=False):
    """Experiment builds: lib/variants/librtb_<tag>.so with extra -D flags (selected at run time through
    the RTB_PLUGIN_LIB environment variable, see plugin.lib())."""
    d = os.path.join(LIB, "variants")
    os.makedirs(d, exist_ok=True)
    out = os.path.join(d, f"librtb_{tag}.so")
    _compile_and_link(out, os.path.join(PKG, "build", "variant_" + tag), [f"-D{x}" for x in defines], verbose)
    return out


def build_all(force=False, verbose=False):
    return build_host(force, verbose), build_plugin(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("built:", host_lib_path(), plugin_lib_path())
