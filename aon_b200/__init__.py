"""Importable alias of the package directory ``articulated-object-nerf_b200/`` (a hyphen cannot
appear in a Python identifier).  ``import aon_b200`` executes that directory's ``__init__.py`` with
this module's ``__path__`` pointing there, so ``aon_b200.lib``, ``aon_b200.nerf`` ... resolve to
``articulated-object-nerf_b200/lib.py`` etc."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "articulated-object-nerf_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
