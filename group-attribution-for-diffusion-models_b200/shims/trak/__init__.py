"""`trak` import shim: put `<repo>/group-attribution-for-diffusion-models_b200/shims` (and the repo root) on
PYTHONPATH and the reference's featuriser scripts run unmodified:

    from trak.projectors import CudaProjector, ProjectionType   # d_trak_grad.py:14, grad_text_to_image_lora.py:68
    from trak.utils import is_not_buffer                         # d_trak_grad.py:15
"""
from . import projectors, utils  # noqa: F401

__version__ = "0.1.3+gadm_b200"
