"""Edge cases at the boundaries of the hot path: tiny / ragged / maximal shapes, rank-deficient and degenerate inputs."""
import numpy as np
import pytest
import torch

from oracle import aggregation as oagg
from oracle import philox
from oracle.projector import project_explicit

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,D", [(1, 1), (1, 63), (3, 64), (2, 65), (33, 127)])
def test_projection_tiny_and_ragged_shapes(B, D):
    from gadm_b200 import CudaProjector, ProjectionType

    g = torch.Generator().manual_seed(D)
    grads = torch.randn(B, D, generator=g).to(DEV)
    p = CudaProjector(D, 512, 3, ProjectionType.rademacher, DEV, 8)
    got = p.project(grads, 0).cpu().numpy().astype(np.float64)
    want = project_explicit(grads.cpu().numpy(), seed=3, model_id=0, proj_type="rademacher", proj_dim=512)
    assert np.abs(got - want).max() <= 1e-5 * max(1.0, np.abs(want).max())


def test_projection_max_proj_dim_32768():
    """BASELINE config 4 uses proj_dim 32768 (traks.py:31-35): 128 column tiles."""
    from gadm_b200 import CudaProjector, ProjectionType

    D, k, B = 3000, 32768, 5
    grads = torch.randn(B, D, generator=torch.Generator().manual_seed(0)).to(DEV)
    p = CudaProjector(D, k, 11, ProjectionType.rademacher, DEV, 8)
    got = p.project(grads, 0).cpu().numpy().astype(np.float64)
    want = project_explicit(grads.cpu().numpy(), seed=11, model_id=0, proj_type="rademacher", proj_dim=k)
    assert np.abs(got - want).max() <= 2e-4 * np.linalg.norm(grads.cpu().numpy(), axis=1).max()


def test_projection_mixed_dtype_blocks_and_empty_block():
    from gadm_b200 import CudaProjector, ProjectionType

    B = 6
    g = torch.Generator().manual_seed(2)
    a = torch.randn(B, 10, 7, generator=g).to(DEV)
    b = torch.randn(B, 33, generator=g).to(DEV).to(torch.bfloat16)
    c = torch.zeros(B, 0, device=DEV)  # parameter with no elements
    d = torch.randn(B, 50, generator=g).to(DEV).half()
    p = CudaProjector(70 + 33 + 50, 512, 1, ProjectionType.normal, DEV, 8)
    got = p.project({"a": a, "b": b, "c": c, "d": d}, 0)
    flat = torch.cat([a.reshape(B, -1), b.float(), d.float()], dim=1)
    assert torch.equal(got, p.project(flat, 0))


def test_shapley_single_behaviour_single_row_and_degenerate_columns():
    import gadm_b200 as G

    rng = np.random.RandomState(0)
    n, d = 40, 6
    X = (rng.rand(n, d) > 0.5).astype(float)
    X[:, 2] = 1.0  # always-present player
    X[:, 4] = 0.0  # never-present player  -> singular normal equations, pinv path
    y = rng.normal(size=n)
    got = G.data_shapley(d, X, y, 1.3, -0.2)
    want = oagg.data_shapley(d, X, y, 1.3, -0.2)
    assert got.shape == (d, 1)
    assert np.abs(got - want).max() <= 1e-8 * max(1.0, np.abs(want).max())
    gb = G.data_banzhaf(X, y)
    wb = oagg.data_banzhaf(X, y)
    assert np.abs(gb - wb).max() <= 1e-8 * max(1.0, np.abs(wb).max())
    # fewer subsets than players (lds.py sweeps start at 10 subsets)
    Xs, ys = X[:4], y[:4]
    assert np.abs(G.data_shapley(d, Xs, ys, 1.0, 0.0) - oagg.data_shapley(d, Xs, ys, 1.0, 0.0)).max() < 1e-8


def test_lds_degenerate_sizes():
    import warnings
    import gadm_b200 as G
    from scipy.stats import spearmanr

    # one test subset: spearmanr of a single point is NaN
    X1 = np.array([[1.0, 0.0, 1.0]])
    rho = G.spearman_matrix(X1, np.array([[0.3, 0.1]]), np.ones((3, 2)))
    assert np.isnan(rho).all()
    # two subsets, perfectly (anti)correlated
    X2 = np.array([[1.0, 0.0], [0.0, 1.0]])
    attrs = np.array([[1.0, 1.0], [2.0, 2.0]])
    Y2 = np.array([[0.0, 1.0], [1.0, 0.0]])
    rho = G.spearman_matrix(X2, Y2, attrs)[0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = [spearmanr(X2 @ attrs[:, k], Y2[:, k]).statistic for k in range(2)]
    np.testing.assert_allclose(rho, want, atol=1e-12)
    with pytest.raises(ValueError):
        G.PackedMasks(np.array([[0.5, 1.0]]))  # masks must be 0/1
    with pytest.raises(ValueError):
        G.data_shapley(3, np.zeros((4, 2)), np.zeros(4), 0.0, 0.0)  # dataset_size != mask width


def test_group_reduce_and_rank_edge_cases():
    import gadm_b200 as G

    vals = np.array([1.0, -2.0, 3.5, 0.25], dtype=np.float32)
    groups = np.array([0, 2, 2, 0])
    s = G.group_reduce(vals, groups, 3, "sum")
    np.testing.assert_allclose(s, [1.25, 0.0, 1.5])  # empty group sums to 0 like numpy
    m = G.group_reduce(vals, groups, 3, "mean")
    assert np.isnan(m[1]) and m[0] == 0.625
    assert G.stable_rank(np.array([2.0])).tolist() == [0]
    x = np.array([[0.0, 0.0], [1.0, -1.0], [0.0, 0.0]])  # all row means equal -> index order
    assert G.stable_rank(x).tolist() == [0, 1, 2]


def test_scorer_more_features_than_examples():
    """N < k (SD-LoRA config: 5000 examples, 32768 features): K = Phi^T Phi + 0.5 I is N-rank plus a ridge."""
    import gadm_b200 as G
    from oracle import scorer as oscore

    N, k, T = 40, 256, 5
    train = torch.randn(N, k, generator=torch.Generator().manual_seed(0)).to(DEV)
    gen = torch.randn(T, k, generator=torch.Generator().manual_seed(1)).to(DEV)
    got = G.trak_scores(train, gen, lam=0.5)
    want = oscore.score_fp64(train.cpu().numpy(), gen.cpu().numpy(), 0.5)
    ref32 = oscore.score_torch(train.cpu().numpy(), gen.cpu().numpy(), 0.5)
    for name in ("trak", "relative_influence", "renorm_influence", "grad_sim"):
        g = got[name].cpu().numpy().astype(np.float64)
        scale = np.abs(want[name]).max()
        ours = np.abs(g - want[name]).max() / scale
        theirs = np.abs(ref32[name].astype(np.float64) - want[name]).max() / scale
        assert ours <= max(8 * theirs, 2e-4), (name, ours, theirs)
