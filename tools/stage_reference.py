"""Stage the UNMODIFIED reference under baseline/_ref/ so that it travels to the GPU box.

    python tools/stage_reference.py [--src /root/reference]

`baseline/_ref/` is git-ignored (the reference's sources never enter this repository's history) but NOT
gpurun-ignored, so the copy ships with the snapshot exactly like the built `.so`.  On the box it is used for
  * the `-m gpu` tests that run the reference's own `run_wo_oc.py` / `run_ddpg.py` UNCHANGED against `pic_b200.PIC`
    (tests/test_runners.py) and, beside them, against the reference's own CPU `PIC` for a live comparison;
  * `bench.py --impl reference` and the `cpu_baseline` legs, which time the reference's `PIC.update_state`
    (src/env/pic.py:131-146) itself when it is present (`kind: "reference"`).
Only the Python sources are copied (runners + src/); result PDFs, notebooks and byte-code caches are left behind.
`__graft_entry__.build()` calls this when /root/reference exists.
"""
import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")


def find_reference():
    """Where the unmodified reference lives: $PIC_REFERENCE, /root/reference, or the staged copy."""
    for p in (os.environ.get("PIC_REFERENCE"), "/root/reference", DEST):
        if p and os.path.exists(os.path.join(p, "src", "env", "pic.py")):
            return p
    return None


def stage(src="/root/reference", dest=DEST, verbose=True):
    if not os.path.exists(os.path.join(src, "src", "env", "pic.py")):
        raise FileNotFoundError("no reference tree at %s" % src)
    if os.path.abspath(src) == os.path.abspath(dest):
        return dest
    n = 0
    for base, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "result", "analysis", ".git")]
        rel = os.path.relpath(base, src)
        for f in files:
            if not f.endswith((".py", ".md", ".cff")):
                continue
            out_dir = os.path.join(dest, rel) if rel != "." else dest
            os.makedirs(out_dir, exist_ok=True)
            s, d = os.path.join(base, f), os.path.join(out_dir, f)
            if not os.path.exists(d) or open(s, "rb").read() != open(d, "rb").read():
                shutil.copyfile(s, d)
                os.chmod(d, 0o644)
            n += 1
    if verbose:
        print("staged %d reference files from %s into %s" % (n, src, dest))
    return dest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    stage(a.src)
    sys.exit(0)
