from gadm_b200.projectors import is_not_buffer  # noqa: F401
