"""Builds csrc/*.cu into lib/libpic_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python optimal-control-1d-electrostatic-plasma_b200/build.py [--force]

One object per translation unit, compiled in parallel; objects live in build/ (git-ignored), the shared library
in-tree under lib/ so that it travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(ROOT, "build", "pic_b200")
LIB = os.path.join(LIBDIR, "libpic_b200.so")

SOURCES = ["pic_stream_f64_a.cu", "pic_stream_f64_b.cu", "pic_stream_f64_c.cu", "pic_stream_f32.cu",
           "pic_stream_tex.cu", "pic_coop.cu", "pic_resident.cu", "pic_cluster.cu", "pic_tsc.cu", "pic_capi.cu", "pic_variants.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]
NVCC_FLAGS += os.environ.get("PIC_EXTRA_NVCC_FLAGS", "").split()     # experiments, e.g. -DPIC_ST_FLAVOR=1


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(d, f), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    stamp = LIB + ".stamp"
    return not (os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == _digest())


def build_library(force=False, verbose=True):
    """Compile if sources changed since the last build. Returns the path of the shared library."""
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(LIB + ".stamp", "w") as f:
        f.write(_digest())
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv)
