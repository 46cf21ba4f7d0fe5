"""Importable alias of the package directory `unified-unlearning-w-remain-geometry_b200/`
(whose name, mandated by the repo layout, is not a valid Python identifier)."""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "unified-unlearning-w-remain-geometry_b200")
__path__ = [_pkg_dir]
with open(_os.path.join(_pkg_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
