"""Importable alias of the `marie-icr_b200/` package directory (a hyphen is not a valid module name).

All code lives in `marie-icr_b200/`; this shim only redirects the package search path there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "marie-icr_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
