"""The oracle restatements against vectors produced by the reference's own code
(tests/golden/make_golden.py) and against published known-answer vectors."""
import os

import numpy as np
import pytest

from oracle import aggregation as oagg
from oracle import philox
from oracle import scorer as oscore

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = tuple(int(x) for x in philox.philox4x32_10(*ctr, *key))
        assert got == want


def test_matrix_is_tiling_independent():
    R = philox.rademacher_matrix(42, 0, 300, 96)
    assert set(np.unique(R)) == {-1, 1}
    assert (philox.rademacher_matrix(42, 37, 100, 64) == R[37:137, :64]).all()
    N = philox.normal_matrix(42, 0, 300, 96)
    assert (philox.normal_matrix(42, 37, 100, 64) == N[37:137, :64]).all()
    assert abs(N.std() - 1) < 0.02 and abs(N.mean()) < 0.02
    assert not (philox.rademacher_matrix(43, 0, 64, 64) == R[:64, :64]).all()


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_shapley_banzhaf_vs_reference(case):
    g = np.load(os.path.join(GOLDEN, "aggregation_golden.npz"))
    Xs, Ys, Xu, Yu = g[f"{case}_Xs"], g[f"{case}_Ys"], g[f"{case}_Xu"], g[f"{case}_Yu"]
    v0, v1 = g[f"{case}_v0"], g[f"{case}_v1"]
    d, K = Xs.shape[1], Ys.shape[1]
    phi_s = np.stack([oagg.data_shapley(d, Xs, Ys[:, k], v1[k], v0[k]).flatten() for k in range(K)], axis=1)
    phi_b = np.stack([oagg.data_banzhaf(Xu, Yu[:, k]) for k in range(K)], axis=1)
    np.testing.assert_array_equal(phi_s, g[f"{case}_phi_shapley"])  # same arithmetic, same library
    np.testing.assert_array_equal(phi_b, g[f"{case}_phi_banzhaf"])
    tests = [(g[f"{case}_Xt{t}"], g[f"{case}_Yt{t}"]) for t in range(3)]
    np.testing.assert_allclose(oagg.evaluate_lds(phi_s, tests, K), g[f"{case}_lds_shapley"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(oagg.evaluate_lds(phi_b, tests, K), g[f"{case}_lds_banzhaf"], rtol=0, atol=1e-12)


def test_mask_samplers_vs_reference():
    g = np.load(os.path.join(GOLDEN, "aggregation_golden.npz"))
    Xs, Xu = g["b_Xs"], g["b_Xu"]
    n, d = Xs.shape
    np.testing.assert_array_equal(oagg.shapley_masks(d, list(range(n))), Xs)
    np.testing.assert_array_equal(oagg.uniform_masks(d, list(range(n))), Xu)
    m = g["b_Xt0"].shape[0]
    np.testing.assert_array_equal(oagg.datamodel_masks(d, list(range(1000, 1000 + m))), g["b_Xt0"])


def test_lds_py_convention():
    g = np.load(os.path.join(GOLDEN, "lds_py_golden.npz"))
    attrs = g["attrs"]
    tests = [(g[f"Xt{t}"], g[f"Yt{t}"]) for t in range(3)]
    got = oagg.evaluate_lds(list(attrs), tests, attrs.shape[0], index_first=True)
    np.testing.assert_allclose(got, g["lds"], rtol=0, atol=1e-12)


def test_score_numpy_vs_compute_gradient_scores():
    g = np.load(os.path.join(GOLDEN, "gradient_scores_golden.npz"))
    tr, va, labels = g["in_train"], g["in_val"], g["in_labels"]
    for gtype in ("trak", "d_trak", "relative_if", "renormalized_if", "vanilla_gradient"):
        scores, kernel = oscore.score_numpy(tr, va, gradient_type=gtype, average=False)
        np.testing.assert_allclose(kernel, g[f"out_{gtype}_kernel"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(scores, g[f"out_{gtype}_byclass=0"], rtol=1e-10, atol=1e-13)
        coeff = oscore.aggregate_by_class(scores.mean(axis=0), labels, "mean")
        np.testing.assert_allclose(coeff, g[f"out_{gtype}_byclass=1"], rtol=1e-10, atol=1e-13)


def test_score_torch_vs_traks_main():
    g = np.load(os.path.join(GOLDEN, "traks_golden.npz"))
    groups = g["in_groups"]
    gid = {i: np.where(groups == i)[0] for i in range(int(groups.max()) + 1)}
    out = oscore.score_torch(g["in_train_loss"], g["in_gen_loss"], journey_grads=g["in_journey"])
    out["dtrak"] = oscore.score_torch(g["in_train_dtrak"], g["in_gen_dtrak"])["trak"]
    agg = oscore.group_aggregate(out, gid)
    for name, val in agg.items():
        np.testing.assert_allclose(val, g[f"out_artist_{name}"], rtol=2e-5, atol=1e-7)
        np.testing.assert_array_equal(oscore.stable_rank(g[f"out_artist_{name}"]),
                                      g[f"out_all_generated_images_artist_rank_{name}"])


def test_shapley_properties():
    rng = np.random.RandomState(0)
    d, n = 10, 400
    X = oagg.shapley_masks(d, list(range(n)))
    w = rng.normal(size=d)
    w[3] = 0.0  # null player
    w[5] = w[6]  # symmetric players
    y = X @ w
    phi = oagg.data_shapley(d, X, y, w.sum(), 0.0).flatten()
    assert abs(phi.sum() - w.sum()) < 1e-9  # efficiency
    np.testing.assert_allclose(phi, w, atol=1e-8)  # linear game -> exact recovery
    Xu = oagg.uniform_masks(d, list(range(n)))
    np.testing.assert_allclose(oagg.data_banzhaf(Xu, (Xu - 0.5) @ w), w, atol=1e-8)  # no intercept in the model


def test_numpy_pairwise_sum_restatement_is_numpys_order():
    """The summation order csrc/aggregate.cuh reproduces (rank hygiene: np.argsort(-x.mean(-1)) near-ties) is numpy's
    own: 0 + pairwise_sum, for every n up to 300 and sizes around numpy's buffer / block limits, fp32 and fp64."""
    import warnings

    from oracle.aggregation import numpy_pairwise_sum

    rng = np.random.RandomState(0)
    for dtype in (np.float32, np.float64):
        for n in list(range(1, 300)) + [1000, 1001, 4096, 8191, 8192, 8193, 20000]:
            a = (rng.normal(size=n) * rng.choice([1, 1e3, 1e-3], size=n)).astype(dtype)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                assert dtype(dtype(0) + numpy_pairwise_sum(a, dtype)) == a.sum(), (dtype, n)
    x = rng.normal(size=(40, 53))
    want = x.mean(axis=-1)
    got = np.array([(0.0 + numpy_pairwise_sum(r)) / 53 for r in x])
    np.testing.assert_array_equal(got, want)


def test_by_class_mask_samplers_match_reference_functions(golden_dir):
    """src/datasets.py:603-617,651-673 (by_class branches), golden produced by the reference's own functions."""
    import importlib.util

    root = os.path.dirname(os.path.dirname(golden_dir))
    spec = importlib.util.spec_from_file_location(
        "gadm_masks", os.path.join(root, "group-attribution-for-diffusion-models_b200", "masks.py"))
    masks = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(masks)
    g = np.load(os.path.join(golden_dir, "masks_by_class_golden.npz"))
    labels = g["labels"]
    for seed in range(6):
        rem, removed = masks.remove_data_by_datamodel(len(labels), 0.5, seed, by_class=True, labels=labels)
        np.testing.assert_array_equal(rem, g[f"datamodel_rem_{seed}"])
        np.testing.assert_array_equal(removed, g[f"datamodel_removed_{seed}"])
        rem, removed = masks.remove_data_by_shapley(len(labels), seed, by_class=True, labels=labels)
        np.testing.assert_array_equal(rem, g[f"shapley_rem_{seed}"])
        np.testing.assert_array_equal(removed, g[f"shapley_removed_{seed}"])


def test_f16_projection_matrix_keeps_eight_significant_bits():
    """P for F16G staging is an fp16 number whose three low mantissa bits are zero (csrc/philox.cuh::box_muller_pair:
    half-ulp bias, then mask), i.e. it carries the 8 significant bits of the bf16 path; it never differs from the
    bf16-rounded value by more than one bf16 ulp, and sign / magnitude statistics are untouched."""
    from oracle import philox

    s = philox.seed64_of(42, 0)
    f16 = philox.normal_matrix(s, 0, 512, 96, "f16")
    bf16 = philox.normal_matrix(s, 0, 512, 96, "bf16")
    raw = philox.normal_matrix(s, 0, 512, 96, None)
    bits = f16.astype(np.float16).view(np.uint16)
    assert (f16.astype(np.float16).astype(np.float32) == f16).all()       # representable in fp16
    assert ((bits & 7) == 0).all()                                        # 8 significant bits
    normal = np.abs(raw) >= 2.0 ** -14                                    # fp16 subnormals have a coarser grid
    # round to fp16 (11 bits), then half-up to 8 bits: at most half an 8-bit ulp plus half an 11-bit ulp
    assert (np.abs(f16 - raw)[normal] <= np.abs(raw)[normal] * 2.0 ** -8 * 1.13).all()
    assert (np.abs(f16 - bf16)[normal] <= np.abs(raw)[normal] * 2.0 ** -7).all()
    assert (np.sign(f16) == np.sign(raw))[normal].all()
    assert abs(float(f16.std()) - float(raw.std())) < 2e-3 and abs(float(f16.mean()) - float(raw.mean())) < 1e-3
