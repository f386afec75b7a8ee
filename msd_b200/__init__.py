"""Importable alias of ``multi-spectral-satellite-image-denoising_b200/`` (a hyphenated directory cannot be named in
an import statement): this package's search path IS that directory, so ``msd_b200.models`` etc. resolve there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "multi-spectral-satellite-image-denoising_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
