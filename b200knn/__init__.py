"""Import shim: the package lives in ``image-retrieval---thesis-2026_b200/`` (not a valid Python identifier),
so ``import b200knn`` re-roots this package's search path there and executes its ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "image-retrieval---thesis-2026_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
