"""Import alias: the package directory is named `optimal-control-1d-electrostatic-plasma_b200` (not a valid Python
identifier), so `import pic_b200` resolves to it through this shim, which turns itself into that package."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "optimal-control-1d-electrostatic-plasma_b200")]
__package__ = "pic_b200"
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
