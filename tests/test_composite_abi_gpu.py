"""The one-call-per-reference-function entry points of include/gadm.h (gadm_project, gadm_gram, gadm_score,
gadm_shapley, gadm_banzhaf, gadm_lds) through ctypes, against the oracle and against the host-composed path
(same kernels in the same order => bitwise equal)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import aggregation as oagg
from oracle import scorer as oscore
from oracle.projector import project_explicit

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


class Block(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("numel_per_example", C.c_int64), ("example_stride", C.c_int64),
                ("row_offset", C.c_int64)]


def _lib():
    from gadm_b200 import _lib as L

    h = L.get_handle(DEV)
    return L, h, L.stream_ptr(DEV)


def test_gadm_project_blocks():
    import gadm_b200 as G

    L, h, st = _lib()
    B, k = 12, 512
    sizes = [100, 3000, 7, 4093]
    D = sum(sizes)
    d_pad = -(-D // 64) * 64
    g = torch.Generator().manual_seed(3)
    parts = [(torch.randn(B, n, generator=g) * 1e-2).to(DEV) for n in sizes]
    staged = torch.empty(d_pad // 64, 32, 64, dtype=torch.float16, device=DEV)  # GADM_STAGE_F16G, no pre-zeroing needed
    inv_scale = torch.empty(32, int(h.lib.gadm_stage_scale_count(d_pad)), device=DEV)
    ws = torch.empty(int(h.lib.gadm_project_workspace_bytes(h.ptr, B, d_pad, k, 2)), dtype=torch.uint8, device=DEV)
    out = torch.empty(B, k, device=DEV)
    blocks = (Block * len(parts))()
    off = 0
    for i, p in enumerate(parts):
        blocks[i] = Block(p.data_ptr(), p.shape[1], p.stride(0), off)
        off += p.shape[1]
    L.check(h.lib.gadm_project(h.ptr, C.cast(blocks, C.c_void_p), len(parts), 0, B, 0.5, staged.data_ptr(), 1,
                               inv_scale.data_ptr(), d_pad, 32, k, 1234, 1, out.data_ptr(), out.stride(0), 0, ws.data_ptr(),
                               ws.numel(), 2, st))
    torch.cuda.synchronize()
    full = torch.cat(parts, dim=1)
    want = project_explicit((full * 0.5).cpu().numpy(), seed=1234, model_id=0, proj_type="rademacher", proj_dim=k)
    gn = np.linalg.norm(full.cpu().numpy().astype(np.float64) * 0.5, axis=1, keepdims=True)
    assert np.all(np.abs(out.cpu().numpy() - want) <= 2e-4 * gn)
    ref = G.CudaProjector(D, k, 1234, G.ProjectionType.rademacher, DEV, 32).project(full * 0.5, 0)
    assert torch.equal(out, ref)


def test_gadm_gram_and_score():
    import gadm_b200 as G

    L, h, st = _lib()
    N, k, T = 1500, 512, 24
    g = torch.Generator().manual_seed(5)
    train, gen = torch.randn(N, k, generator=g).to(DEV), torch.randn(T, k, generator=g).to(DEV)
    ld_t = -(-N // 4) * 4
    phi_t = torch.empty(k, ld_t, device=DEV)
    gram = torch.zeros(k, k, device=DEV)
    L.check(h.lib.gadm_gram(h.ptr, train.data_ptr(), N, k, train.stride(0), phi_t.data_ptr(), ld_t, gram.data_ptr(), k, 0.5, 0, st))
    sc = G.TrakScorer(0.5).factor_(gram)
    z = torch.empty(2 * T, k, device=DEV)
    S = torch.empty(T, N, device=DEV)
    mean = torch.empty(N, device=DEV)
    L.check(h.lib.gadm_score(h.ptr, gen.data_ptr(), T, k, sc.X.data_ptr(), sc.X.stride(0), sc.Xt.data_ptr(), sc.Xt.stride(0), k,
                             train.data_ptr(), N, k, z.data_ptr(), k, S.data_ptr(), N, None, mean.data_ptr(), st))
    torch.cuda.synchronize()
    want = oscore.score_fp64(train.cpu().numpy(), gen.cpu().numpy(), 0.5)
    assert np.abs(S.cpu().numpy() - want["scores"]).max() < 2e-4 * np.abs(want["scores"]).max()
    assert np.abs(mean.cpu().numpy() - want["trak"]).max() < 2e-4 * np.abs(want["trak"]).max()
    # host-composed [T, N] path: the same kernels in the same order => bitwise equal
    ref_s = G.TrakScorer(0.5).fit(train, dual=False).score_matrix(gen, train)
    assert torch.equal(S, ref_s) and torch.equal(mean, G.col_mean_scaled(ref_s))
    # trak_scores takes the mean over generated images FIRST (one solved row + a matvec): same numbers to fp32 rounding
    ref = G.trak_scores(train, gen, lam=0.5, variants=("trak",), dual=False)["trak"]
    assert float((mean - ref).abs().max()) < 2e-5 * float(ref.abs().max())


@pytest.mark.parametrize("n,d,K", [(200, 30, 7), (20, 30, 4)])
def test_gadm_shapley_banzhaf_lds(n, d, K):
    import gadm_b200 as G

    L, h, st = _lib()
    rng = np.random.RandomState(n)
    X = oagg.shapley_masks(d, list(range(n)))
    w = rng.normal(size=(d, K))
    Y = X @ w + 0.1 * rng.normal(size=(n, K))
    v0 = 0.05 * rng.normal(size=K)
    v1 = w.sum(axis=0) + v0
    masks = G.PackedMasks(X, DEV)
    y = torch.as_tensor(Y).to(DEV)
    ws = torch.empty(int(h.lib.gadm_shapley_workspace_bytes(d, K)), dtype=torch.uint8, device=DEV)
    phi = torch.empty(d, K, dtype=torch.float64, device=DEV)
    tv1, tv0 = torch.as_tensor(v1).to(DEV), torch.as_tensor(v0).to(DEV)
    L.check(h.lib.gadm_shapley(h.ptr, masks.rowbits.data_ptr(), masks.colbits.data_ptr(), n, d, y.data_ptr(), K, tv1.data_ptr(),
                               tv0.data_ptr(), ws.data_ptr(), ws.numel(), phi.data_ptr(), st))
    want = np.stack([oagg.data_shapley(d, X, Y[:, i], v1[i], v0[i])[:, 0] for i in range(K)], axis=1)
    np.testing.assert_allclose(phi.cpu().numpy(), want, rtol=1e-7, atol=1e-9)
    assert np.array_equal(phi.cpu().numpy(), G.data_shapley_batched(X, Y, v1, v0))
    Xu = oagg.uniform_masks(d, list(range(n)))
    mu = G.PackedMasks(Xu, DEV)
    phib = torch.empty(d, K, dtype=torch.float64, device=DEV)
    L.check(h.lib.gadm_banzhaf(h.ptr, mu.rowbits.data_ptr(), mu.colbits.data_ptr(), n, d, y.data_ptr(), K, ws.data_ptr(),
                               ws.numel(), phib.data_ptr(), st))
    assert np.array_equal(phib.cpu().numpy(), G.data_banzhaf_batched(Xu, Y))
    # LDS of one test set, identity evaluation and two resamples
    m = 25
    Xt = oagg.datamodel_masks(d, list(range(500, 500 + m)))
    Yt = Xt @ w + 0.3 * rng.normal(size=(m, K))
    tm = G.PackedMasks(Xt, DEV)
    yt = torch.as_tensor(Yt).to(DEV)
    ws2 = torch.empty((m + 3) * K * 8, dtype=torch.uint8, device=DEV)
    out = torch.empty(1, dtype=torch.float64, device=DEV)
    L.check(h.lib.gadm_lds(h.ptr, tm.colbits.data_ptr(), m, d, yt.data_ptr(), phi.data_ptr(), K, None, 1, m, ws2.data_ptr(),
                           ws2.numel(), out.data_ptr(), st))
    want_lds = np.mean(oagg.spearman_matrix(Xt, Yt, phi.cpu().numpy())) * 100
    assert abs(float(out.item()) - want_lds) < 1e-9
    idx = torch.as_tensor(rng.randint(0, m, size=(3, m)).astype(np.int32)).to(DEV)
    out3 = torch.empty(3, dtype=torch.float64, device=DEV)
    L.check(h.lib.gadm_lds(h.ptr, tm.colbits.data_ptr(), m, d, yt.data_ptr(), phi.data_ptr(), K, idx.data_ptr(), 3, m,
                           ws2.data_ptr(), ws2.numel(), out3.data_ptr(), st))
    ref3 = G.lds_per_test_set(tm, yt, phi, idx.cpu().numpy())
    assert np.array_equal(out3.cpu().numpy(), ref3)
    # workspace too small -> error code, message, no launch
    rc = h.lib.gadm_shapley(h.ptr, masks.rowbits.data_ptr(), masks.colbits.data_ptr(), n, d, y.data_ptr(), K, tv1.data_ptr(),
                            tv0.data_ptr(), ws.data_ptr(), 16, phi.data_ptr(), st)
    assert rc == -3 and "workspace" in L.last_error()
