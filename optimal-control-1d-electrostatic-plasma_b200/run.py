"""Launcher that runs the reference's own scripts UNCHANGED on the B200 PIC:

    python -m pic_b200.run /path/to/reference/run_wo_oc.py --simcase bump-on-tail
    python -m pic_b200.run /path/to/reference/run_ddpg.py --simcase bump-on-tail

What it does:
  1. registers a module object as ``sys.modules["src.env.pic"]`` whose ``PIC`` is `pic_b200.PIC` -- every
     ``from src.env.pic import PIC`` in the runners and trainers (run_wo_oc.py:5, ddpg.py:9, ...) then resolves to
     the CUDA env.  ``src`` is a namespace package without ``__init__.py``, so pre-seeding ``sys.modules`` suffices;
  2. reproduces the import side effect of the reference module: ``np.random.seed(42)`` (src/env/pic.py:12) at the
     moment the module would have been imported, so the host sampler draws the same particles;
  3. installs a do-nothing ``matplotlib`` when the real one is missing (``src/plot.py:2`` imports it at import time
     and the runners cannot start without it); figures are then skipped, everything else runs;
  4. ``runpy``-executes the script with the reference directory first on ``sys.path``.
"""
import os
import runpy
import sys
import types

import numpy as np


class _Anything:
    """Absorbs any matplotlib call chain: attributes, calls, indexing, iteration over axes arrays."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __getitem__(self, k):
        return _Anything()

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter([_Anything() for _ in range(8)])

    def __len__(self):
        return 8

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _PltStub(types.ModuleType):
    def subplots(self, *a, **k):
        return _Anything(), _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


def install_matplotlib_stub(force=False):
    """Returns True when a stub was installed (the real matplotlib is absent or `force`)."""
    if not force:
        try:
            import matplotlib  # noqa: F401
            return False
        except Exception:
            pass
    root = types.ModuleType("matplotlib")
    root.__path__ = []
    root.use = lambda *a, **k: None
    root.rcParams = {}
    plt = _PltStub("matplotlib.pyplot")
    root.pyplot = plt
    sys.modules["matplotlib"] = root
    sys.modules["matplotlib.pyplot"] = plt
    for sub in ("gridspec", "colors", "cm", "animation", "ticker", "patches"):
        m = _PltStub("matplotlib." + sub)
        setattr(root, sub, m)
        sys.modules["matplotlib." + sub] = m
    return True


def make_pic_module(pic_class=None, seed=42):
    """The module object that stands in for the reference's ``src/env/pic.py``."""
    if pic_class is None:
        from .pic import PIC as pic_class          # the CUDA env; raises if the library cannot be loaded
    from .dist import BumpOnTail, TwoStream
    mod = types.ModuleType("src.env.pic")
    mod.PIC = pic_class
    mod.TwoStream, mod.BumpOnTail = TwoStream, BumpOnTail
    mod.__file__ = __file__
    if seed is not None:
        np.random.seed(seed)                        # src/env/pic.py:12 runs this in the class body at import
    return mod


def run_script(script, argv=(), pic_class=None, reference_dir=None, stub_matplotlib=None):
    """Execute `script` (a reference runner) as __main__ with our PIC injected.  Returns the script's globals."""
    script = os.path.abspath(script)
    ref = os.path.abspath(reference_dir or os.path.dirname(script))
    saved_argv, saved_path = sys.argv[:], sys.path[:]
    saved_mods = {k: sys.modules.get(k) for k in ("src.env.pic", "matplotlib", "matplotlib.pyplot")}
    try:
        if stub_matplotlib is None:
            install_matplotlib_stub()
        elif stub_matplotlib:
            install_matplotlib_stub(force=True)
        sys.path.insert(0, ref)
        sys.dont_write_bytecode = True              # the reference tree may be read-only
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]                      # a previous run's reference modules hold the previous PIC
        sys.modules["src.env.pic"] = make_pic_module(pic_class)
        sys.argv = [script] + list(argv)
        return runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv, sys.path[:] = saved_argv, saved_path
        for k, m in saved_mods.items():
            if m is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = m


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        raise SystemExit(2)
    run_script(sys.argv[1], sys.argv[2:])


if __name__ == "__main__":
    main()
