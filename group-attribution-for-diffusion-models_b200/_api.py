"""Public names of the package (import as ``gadm_b200``)."""
from .projectors import BasicProjector, CudaProjector, DeferredProjection, ProjectionType, is_not_buffer  # noqa: F401
from ._lib import GadmError, load_library  # noqa: F401

__all__ = [
    "BasicProjector", "CudaProjector", "DeferredProjection", "ProjectionType", "is_not_buffer",
    "GadmError", "load_library",
]
