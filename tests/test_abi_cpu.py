"""CPU checks: the C-ABI library loads and exports every symbol include/gadm.h declares, the ctypes table
covers the header, and the host-side mirror of the reference interface behaves like the reference on bad
input.  No kernel is launched (there is no GPU in the build container)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "gadm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gadm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import gadm_b200
    from gadm_b200 import _lib

    lib = gadm_b200.load_library()
    declared = _header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gadm.h but not exported"
    assert sorted(_lib.declared_symbols()) == declared, "ctypes signature table and header disagree"
    assert lib.gadm_version() >= 100
    assert isinstance(lib.gadm_last_error(), bytes)


def test_no_gpu_fails_loudly():
    from gadm_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = _lib.load_library().gadm_create(ctypes.byref(h), 0)
    assert rc < 0 and h.value is None
    assert len(_lib.last_error()) > 0
    with pytest.raises((RuntimeError, ValueError)):
        _lib.get_handle("cuda:0")


def test_missing_library_is_an_import_error(tmp_path):
    from gadm_b200 import _lib

    saved = _lib._lib
    _lib._lib = None
    try:
        with pytest.raises(ImportError):
            _lib.load_library(str(tmp_path / "nope.so"))
    finally:
        _lib._lib = saved


def test_projector_host_side_errors_mirror_trak():
    from gadm_b200 import CudaProjector, ProjectionType, is_not_buffer
    from gadm_b200.projectors import _as_blocks

    assert ProjectionType("normal") is ProjectionType.normal and ProjectionType.rademacher.value == "rademacher"
    with pytest.raises(ValueError):  # trak: "CudaProjector only works on cuda device!"
        CudaProjector(grad_dim=10, proj_dim=512, seed=0, proj_type=ProjectionType.normal, device="cpu", max_batch_size=8)
    with pytest.raises(KeyError):
        CudaProjector(10, 512, 0, "gaussian", "cpu", 8)
    names = ["conv.weight", "bn.running_mean", "bn.running_var", "bn.num_batches_tracked", "fc.bias"]
    assert [is_not_buffer(i, names) for i in range(5)] == [True, False, False, False, True]
    # dict of per-parameter gradients (vmap(grad) output) is flattened block by block
    g = {"a": torch.zeros(3, 2, 5), "b": torch.zeros(3, 7)}
    blocks = _as_blocks(g)
    assert [tuple(b.shape) for b in blocks] == [(3, 10), (3, 7)]
    with pytest.raises(ValueError):
        _as_blocks({"a": torch.zeros(3, 2), "b": torch.zeros(4, 2)})
    with pytest.raises(ValueError):
        _as_blocks(torch.zeros(3))


def test_trak_shim_imports():
    import sys

    shim = os.path.join(ROOT, "group-attribution-for-diffusion-models_b200", "shims")
    sys.path.insert(0, shim)
    try:
        for m in [m for m in sys.modules if m == "trak" or m.startswith("trak.")]:
            del sys.modules[m]
        from trak.projectors import BasicProjector, CudaProjector, ProjectionType  # noqa: F401  (d_trak_grad.py:14)
        from trak.utils import is_not_buffer  # noqa: F401  (d_trak_grad.py:15)
        import gadm_b200

        assert CudaProjector is gadm_b200.CudaProjector
    finally:
        sys.path.remove(shim)


def test_aggregation_host_validation():
    from gadm_b200 import aggregation as agg

    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):  # no CPU fallback
            agg.data_banzhaf(np.zeros((4, 3)), np.zeros(4))
    with pytest.raises(ValueError):
        agg._device("cpu")


def test_shard_range_partitions():
    from gadm_b200.distributed import shard_range

    for n, w in ((50000, 8), (5000, 3), (7, 8), (0, 2)):
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
