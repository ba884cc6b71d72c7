"""Importable name for the package that lives in ``disentangled-multimodal-learning_b200/``.

The directory name required by the repo layout contains hyphens and cannot be written
in an ``import`` statement; this shim points ``dml_b200``'s package path at it, so
``import dml_b200`` / ``from dml_b200.mil import TransMIL`` load the real sources.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "disentangled-multimodal-learning_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
