"""Build an experimental variant of the library into build/variants/<name>.so (used through PIC_LIB_PATH):

    python tools/build_variant.py <name> [extra nvcc flags ...]      e.g.  noclobber -DPIC_ATOM_NOCLOBBER
"""
import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optimal-control-1d-electrostatic-plasma_b200"))
import build as B  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
objdir = os.path.join(ROOT, "build", "variants", name)
os.makedirs(objdir, exist_ok=True)
only = os.environ.get("PIC_VARIANT_SOURCES")
sources = only.split(",") if only else B.SOURCES


def one(src):
    obj = os.path.join(objdir, src.replace(".cu", ".o"))
    r = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + extra + ["-c", os.path.join(B.CSRC, src), "-o", obj],
                       capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr)
    return obj


with cf.ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(one, sources))
lib = os.path.join(ROOT, "build", "variants", name + ".so")
r = subprocess.run([B._nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"],
                   capture_output=True, text=True)
if r.returncode:
    raise RuntimeError(r.stderr)
print("built", lib)
