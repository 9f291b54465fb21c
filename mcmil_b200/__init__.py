"""Importable alias of the package directory `montecarlo-gated-mil_b200/` (a hyphen is not a
valid Python identifier): submodules resolve inside that directory."""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "montecarlo-gated-mil_b200")
__path__ = [_pkg_dir]
with open(_os.path.join(_pkg_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
