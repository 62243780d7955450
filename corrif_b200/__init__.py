"""Importable alias of the product package, whose directory name
(``corrifnet-correlation-aware-interactive-fusion-multimodal-learning-for-multispectral-images_b200``)
is mandated by the project layout but is not a valid Python identifier."""
import os as _os

_REAL = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "corrifnet-correlation-aware-interactive-fusion-multimodal-learning-for-multispectral-images_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
