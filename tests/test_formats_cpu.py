"""Host-side format handlers against what the reference's own collect_data produced (tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_collect_data_matches_reference(tmp_path):
    from gadm_b200.formats import collect_data, load_lds_test_sets, read_behavior_db

    g = np.load(os.path.join(GOLDEN, "aggregation_golden.npz"))
    db = tmp_path / "db.jsonl"
    db.write_bytes(g["collect_db"].tobytes())
    recs = read_behavior_db(str(db))
    assert [r["subset_seed"] for r in recs] == list(range(6))
    X, Y = collect_data(recs, num_groups=9, model_behavior_key="ssim", n_samples=3)
    assert X.dtype == np.float64 and Y.dtype == np.float64
    np.testing.assert_array_equal(X, g["collect_X"])
    np.testing.assert_array_equal(Y, g["collect_Y"])
    # the same through a pandas DataFrame (the reference's call style) and without masks
    import pandas as pd

    df = pd.read_json(str(db), lines=True)
    X2, Y2 = collect_data(df=df, num_groups=9, model_behavior_key="ssim", n_samples=3)
    np.testing.assert_array_equal(X2, X)
    np.testing.assert_array_equal(collect_data(df, 9, "ssim", 3, collect_remaining_masks=False), Y)
    # test-set selection by subset seed (shapley_lds.py:163-170) and the size assertion
    sel = read_behavior_db(str(db), subset_seeds=range(4))
    assert len(sel) == 4
    tests = load_lds_test_sets([str(db)] * 3, 9, "ssim", 3, test_size=5)
    assert len(tests) == 3 and tests[0][0].shape == (5, 9) and tests[0][1].shape == (5, 3)
    try:
        load_lds_test_sets([str(db)], 9, "ssim", 3, test_size=7)
        raise RuntimeError("expected an AssertionError")
    except AssertionError:
        pass


def test_records_out_of_order_and_global_key(tmp_path):
    import json

    from gadm_b200.formats import collect_data, read_behavior_db

    db = tmp_path / "g.jsonl"
    with open(db, "w") as f:
        for s in (3, 0, 2, 1):
            f.write(json.dumps({"exp_name": f"retrain_seed_{s}", "remaining_idx": [s], "fid": float(s) / 2}) + "\n")
    recs = read_behavior_db(str(db))
    x, y = collect_data(recs, 4, "fid", None)
    np.testing.assert_array_equal(x, np.eye(4))
    np.testing.assert_array_equal(y[:, 0], np.arange(4) / 2)
    null = tmp_path / "null.jsonl"
    null.write_text(json.dumps({"exp_name": "null_model", "fid": 1.5}) + "\n")
    assert collect_data(read_behavior_db(str(null)), 4, "fid", None, collect_remaining_masks=False).tolist() == [[1.5]]


def test_journey_points_and_group_csv(tmp_path):
    """Journey-TRAK bookkeeping of grad_text_to_image_lora.py:515-529: which trajectory steps are featurised and the
    group.csv that maps every featurised latent to its generated image."""
    import pandas as pd
    import pytest

    from gadm_b200.formats import journey_point_indices, write_journey_group_csv

    np.testing.assert_array_equal(journey_point_indices(100, 50), np.arange(1, 100, 2))  # the reference's settings
    np.testing.assert_array_equal(journey_point_indices(100, 30), np.arange(1, 100, 3))  # 33 points, as upstream yields
    np.testing.assert_array_equal(journey_point_indices(7, 7), np.arange(1, 7))
    with pytest.raises(ValueError):
        journey_point_indices(10, 0)
    with pytest.raises(ValueError):
        journey_point_indices(10, 11)
    path = write_journey_group_csv(str(tmp_path / "generated_journey"), [0, 0, 1, 1], [1, 3, 1, 3])
    df = pd.read_csv(path, index_col=0)
    assert df.columns.tolist() == ["generated_image_idx", "step_idx"]
    assert df.index.tolist() == [0, 1, 2, 3] and df["step_idx"].tolist() == [1, 3, 1, 3]
