from gadm_b200.projectors import BasicProjector, CudaProjector, ProjectionType  # noqa: F401
