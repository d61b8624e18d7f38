"""Importable alias of the `meta-tts_b200/` package directory (a hyphen is not a valid module name).

All code lives in `meta-tts_b200/`; this shim only redirects the package search path.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "meta-tts_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
