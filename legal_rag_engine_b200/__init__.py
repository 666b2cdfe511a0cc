"""Import shim: the package directory is ``legal-rag-engine_b200/`` (the name
the build contract fixes, not a valid Python identifier); this module makes it
importable as ``legal_rag_engine_b200`` by pointing ``__path__`` at it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "legal-rag-engine_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
