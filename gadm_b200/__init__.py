"""Importable alias of the product package.

The package directory is named after the reference repository
(``group-attribution-for-diffusion-models_b200/``), which is not a valid Python identifier;
this shim extends ``__path__`` so that it is importable as ``gadm_b200``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "group-attribution-for-diffusion-models_b200")
__path__.append(_real)

from gadm_b200._api import *  # noqa: E402,F401,F403
from gadm_b200._api import __all__  # noqa: E402,F401
