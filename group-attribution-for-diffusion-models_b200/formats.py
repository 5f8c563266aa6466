"""The reference's on-disk formats on either side of the hot path.

Host-side Python like the reference (pandas / numpy / torch.load); everything numeric goes to the device kernels.

* jsonl model-behaviour databases (one record per retrained / unlearned model: ``exp_name``, ``remaining_idx``,
  ``generated_image_{i}_{key}`` or a global ``{key}``): ``read_behavior_db``, ``collect_data``
  (text_to_image/shapley_lds.py:105-136,161-176; lds.py:203-257 builds the same arrays).
* the gradient directory of ``text_to_image/traks.py:64-135`` (``emb_f=..._num_timesteps=..._proj_dim=....pt`` tensors,
  ``group.csv``) and its outputs ``baselines/{group}_{name}.npy`` (float64 [G, 1]) and
  ``baselines/all_generated_images_{group}_rank_{name}.npy`` (int64 [G]): ``run_traks``.
* ``artist_{prefix}_fit_size={n}.npy`` / ``all_generated_images_artist_rank_{prefix}_fit_size={n}.npy`` written by
  shapley_lds.py:284-298 and banzhaf_lds.py: ``save_lds_outputs``.
* raw fp32 memmaps + ``kernel_*.npy`` of the unconditional path: ``scoring.compute_gradient_scores``.
"""
from __future__ import annotations

import os
from typing import Iterable, Sequence

import numpy as np
import torch


def read_behavior_db(path: str, subset_seeds: Iterable[int] | None = None) -> list[dict]:
    """Records of a jsonl db sorted by the subset seed parsed from ``exp_name`` (``..._seed_{s}``), optionally
    restricted to ``subset_seeds`` (shapley_lds.py:163-170: ``test_df[test_df["subset_seed"].isin(range(test_size))]``).
    Records without a ``seed_`` suffix (null / full model dbs) keep their file order."""
    import pandas as pd

    # pandas' json reader on purpose: its default float parser (precise_float=False) differs from json.loads in the
    # last bits, and the reference's arrays come from pd.read_json (shapley_lds.py:162)
    df = pd.read_json(path, lines=True)
    if len(df) and "exp_name" in df.columns and df["exp_name"].astype(str).str.contains("seed_").all():
        df["subset_seed"] = df["exp_name"].str.split("seed_", expand=True)[1].astype(int)
        df = df.sort_values(by="subset_seed")
        if subset_seeds is not None:
            df = df[df["subset_seed"].isin([int(s) for s in subset_seeds])]
    return df.to_dict("records")


def _records(df) -> list[dict]:
    if isinstance(df, list):
        return df
    if hasattr(df, "to_dict"):  # pandas DataFrame, row order preserved like iterrows()
        return df.to_dict("records")
    raise TypeError("expected a list of records or a pandas DataFrame")


def collect_data(df, num_groups: int, model_behavior_key: str, n_samples: int | None, collect_remaining_masks: bool = True):
    """Reference signature and return values (shapley_lds.py:105-136): ``(masks [n, num_groups] float64,
    behaviours [n, n_samples or 1] float64)`` or only the behaviours.  ``df`` is a DataFrame or a list of records."""
    recs = _records(df)
    keys = [model_behavior_key] if n_samples is None else [f"generated_image_{i}_{model_behavior_key}" for i in range(n_samples)]
    y = np.empty((len(recs), len(keys)), dtype=np.float64)
    x = np.zeros((len(recs), num_groups), dtype=np.float64) if collect_remaining_masks else None
    for r, rec in enumerate(recs):
        if collect_remaining_masks:
            x[r, np.asarray(rec["remaining_idx"], dtype=np.int64)] = 1
        for c, key in enumerate(keys):
            y[r, c] = rec[key]
    return (x, y) if collect_remaining_masks else y


def load_lds_test_sets(test_db_list: Sequence[str], num_groups: int, model_behavior_key: str, n_samples: int | None,
                       test_size: int):
    """The three held-out datamodel retraining dbs (seeds 42 / 43 / 44) -> ``test_data_list`` of (x_test, y_test)
    (shapley_lds.py:159-180)."""
    out = []
    for path in test_db_list:
        recs = read_behavior_db(path, range(test_size))
        if len(recs) != test_size:
            raise AssertionError(f"{path}: {len(recs)} records with subset_seed < {test_size}, expected {test_size}")
        out.append(collect_data(recs, num_groups, model_behavior_key, n_samples))
    return out


def save_lds_outputs(output_dir: str, outfile_prefix: str, fit_size: int, attrs_all, group: str = "artist") -> np.ndarray:
    """shapley_lds.py:284-298: ``{group}_{prefix}_fit_size={n}.npy`` (float64 [d, K]) and the stable descending rank
    of the row means (int64 [d]).  Returns the rank."""
    from .aggregation import stable_rank

    attrs_all = np.asarray(attrs_all, dtype=np.float64)
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, f"{group}_{outfile_prefix}_fit_size={fit_size}.npy"), "wb") as handle:
        np.save(handle, attrs_all)
    rank = stable_rank(attrs_all)
    with open(os.path.join(output_dir, f"all_generated_images_{group}_rank_{outfile_prefix}_fit_size={fit_size}.npy"), "wb") as handle:
        np.save(handle, rank)
    return rank


def journey_point_indices(num_inference_steps: int, num_journey_points: int) -> np.ndarray:
    """Sampling-trajectory steps whose latents Journey-TRAK featurises
    (text_to_image/grad_text_to_image_lora.py:515-523): ``np.arange(1, n, n // num_journey_points)``."""
    if num_journey_points < 1 or num_journey_points > num_inference_steps:
        raise ValueError(f"num_journey_points must lie in [1, {num_inference_steps}], got {num_journey_points}")
    return np.arange(start=1, stop=num_inference_steps, step=num_inference_steps // num_journey_points)


def write_journey_group_csv(output_dir: str, generated_image_idx: Sequence[int], step_idx: Sequence[int]) -> str:
    """``group.csv`` of the generated / generated_journey featurisation: one row per featurised latent with the image
    it belongs to and its trajectory step (grad_text_to_image_lora.py:526-529; read back by traks.py)."""
    import pandas as pd

    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, "group.csv")
    pd.DataFrame({"generated_image_idx": list(generated_image_idx), "step_idx": list(step_idx)}).to_csv(path, index=True)
    return path


def _group_ids(group_names: Sequence, train_groups: Sequence) -> np.ndarray:
    """traks.py:92-98: index of each training example's group in the group table (-1: not in any group)."""
    lut = {name: i for i, name in enumerate(group_names)}
    return np.array([lut.get(g, -1) for g in train_groups], dtype=np.int32)


def run_traks(args, dataset_dir: str | None = None, device=None):
    """Drop-in for ``text_to_image/traks.py:main(args)``: same inputs (``args.output_dir/gradients/...``, the group table
    ``{dataset_dir}/artbench-10-imagefolder-split/train/{cls}_{group}s.csv``), same output files.

    ``args``: output_dir, num_timesteps, proj_dim, dataset, cls, group, lam (traks.py:12-63).  Returns
    (output_dict, rank_dict)."""
    import pandas as pd

    from .scoring import group_and_rank, trak_scores

    if getattr(args, "dataset", "artbench") != "artbench":
        raise ValueError  # traks.py:77-78
    if dataset_dir is None:
        try:
            from src.ddpm_config import DATASET_DIR as dataset_dir  # the reference's constant (traks.py:10)
        except Exception:
            from src.constants import DATASET_DIR as dataset_dir
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    gradient_dir = getattr(args, "gradient_dir", None) or os.path.join(args.output_dir, "gradients")
    group_df = pd.read_csv(os.path.join(dataset_dir, "artbench-10-imagefolder-split", "train", f"{args.cls}_{args.group}s.csv"))
    sfx = f"num_timesteps={args.num_timesteps}_proj_dim={args.proj_dim}.pt"

    def load(*parts):
        return torch.load(os.path.join(gradient_dir, *parts), map_location="cpu").to(device=device, dtype=torch.float32)

    train_grads = load("train", f"emb_f=loss_{sfx}")
    train_dtrak = load("train", f"emb_f=mean-squared-l2-norm_{sfx}")
    train_df = pd.read_csv(os.path.join(gradient_dir, "train", "group.csv"))
    group_names = [group_df.iloc[i].item() for i in group_df.index]
    group_ids = _group_ids(group_names, train_df[args.group].tolist())
    gen_grads = load("generated", f"emb_f=loss_{sfx}")
    gen_dtrak = load("generated", f"emb_f=mean-squared-l2-norm_{sfx}")
    journey = load("generated_journey", f"emb_f=loss_num_journey_points=50_num_journey_noises=1_proj_dim={args.proj_dim}.pt")

    sample_output_dict = trak_scores(train_grads, gen_grads, lam=args.lam, journey_phi=journey)
    sample_output_dict["dtrak"] = trak_scores(train_dtrak, gen_dtrak, lam=args.lam, variants=("trak",))["trak"]
    output_dict, rank_dict = group_and_rank(sample_output_dict, group_ids, len(group_names))

    output_dir = os.path.join(args.output_dir, "baselines")
    os.makedirs(output_dir, exist_ok=True)
    for name, output in output_dict.items():
        with open(os.path.join(output_dir, f"{args.group}_{name}.npy"), "wb") as handle:
            np.save(handle, output)
    for name, rank in rank_dict.items():
        with open(os.path.join(output_dir, f"all_generated_images_{args.group}_rank_{name}.npy"), "wb") as handle:
            np.save(handle, rank)
    return output_dict, rank_dict
