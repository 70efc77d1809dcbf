"""In-tree build of the sm_100a library (libhitl_gpu.so) and the host mirror (libhitl_host.so).

nvcc cross-compiles without a GPU; the built .so files are git-ignored but travel to the GPU
box with the gpurun snapshot.  Flags:
  * -gencode arch=compute_100a,code=sm_100a -lineinfo            (B200 only, source-mapped SASS)
  * --fmad=false + host -ffp-contract=off for search.cu / em.cu  (bit-exact float geometry)
  * eval.cu keeps FMA contraction (tolerance-level FP64 arithmetic)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "lib")
OBJ = os.path.join(HERE, "lib", "obj")
NVCC = os.environ.get("HITL_NVCC", "/usr/local/cuda/bin/nvcc")
GXX = "/usr/bin/g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-ccbin", GXX] + ARCH

GPU_SOURCES = [
    ("ctx.cu", ["--fmad=false"]),
    ("search.cu", ["--fmad=false"]),
    ("em.cu", ["--fmad=false"]),
    ("kdtree_gpu.cu", ["--fmad=false"]),
    ("backprop.cu", ["--fmad=false"]),
    ("eval.cu", []),
    ("comm.cu", []),
    ("kdtree_build.cpp", []),
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build failed: " + os.path.basename(cmd[-1]))
    return r.stdout


def build_gpu(force=False, verbose=False):
    """Compile every CUDA translation unit for sm_100a and link libhitl_gpu.so."""
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in ("hitl_internal.h", "hitl_math.h", "stdsort_exact.h")] + [os.path.join(HERE, "..", "include", "hitl_gpu.h")]
    objs = []
    for name, extra in GPU_SOURCES:
        src = os.path.join(CSRC, name)
        obj = os.path.join(OBJ, name + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            cmd = [NVCC] + COMMON + extra + ["-Xcompiler", "-fPIC,-ffp-contract=off,-pthread", "-c", src, "-o", obj]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            out = _run(cmd)
            if verbose:
                print(out)
    so = os.path.join(LIB, "libhitl_gpu.so")
    if force or _newer(so, objs):
        _run([NVCC, "-shared", "-ccbin", GXX] + ARCH + ["-o", so] + objs + ["-Xcompiler", "-pthread", "-ldl"])   # NCCL is bound at run time (comm.cu)
    return so


def build_host(force=False):
    """Compile the C++ host mirror (Ceres-shaped cost functions, JointOpt / EMInput mirrors, I/O,
    synthetic generator) against libhitl_gpu.so."""
    srcs = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp")) if os.path.isdir(HOST) else []
    if not srcs:
        return None
    so = os.path.join(LIB, "libhitl_host.so")
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")] + [os.path.join(CSRC, "hitl_math.h")]
    if force or _newer(so, deps + [os.path.join(LIB, "libhitl_gpu.so")]):
        _run([GXX, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I" + os.path.join(HERE, "..", "include"), "-I" + CSRC,
              "-o", so] + srcs + ["-L" + LIB, "-lhitl_gpu", "-Wl,-rpath,$ORIGIN"])
    return so


def build_all(force=False, verbose=False):
    build_gpu(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", os.listdir(LIB))
