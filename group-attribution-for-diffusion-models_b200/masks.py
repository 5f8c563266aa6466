"""Subset-mask samplers and the counterfactual rank consumer -- host-side index arithmetic.

SURVEY.md section 2 item 8: the removal distributions of ``src/datasets.py:559-717`` draw from legacy
``np.random.RandomState(seed)`` streams (``choice`` / ``shuffle`` / ``normal``); they stay on the host and call numpy
verbatim so that masks are bit-exact with the reference's retraining jobs.  These functions return the dense 0/1 mask
rows that ``collect_data`` builds from ``remaining_idx`` (text_to_image/shapley_lds.py:114-119); feed them to
``gadm_b200.PackedMasks`` / ``data_shapley_batched``.
"""
from __future__ import annotations

import math

import numpy as np


def _class_split(labels, kept_classes):
    """Indices whose label is (not) among ``kept_classes``, ascending."""
    labels = np.asarray(labels)
    keep = np.isin(labels, np.asarray(list(kept_classes)))
    idx = np.arange(len(labels))
    return idx[keep], idx[~keep]


def remove_data_by_shapley(dataset_size: int, seed: int = 0, by_class: bool = False, labels=None):
    """src/datasets.py:677-697: size ~ (n-1)/(s(n-s)), then the first s of a shuffle.  Returns (remaining, removed).
    ``by_class`` (src/datasets.py:651-673): the same draw over the distinct class labels -- the classes after the
    sampled size in the shuffled order are removed as a whole; ``labels`` is the per-example label vector (the
    reference iterates ``dataset`` for ``data[1]``)."""
    rng = np.random.RandomState(seed)
    if by_class:
        classes = np.unique(np.asarray(labels))
        sizes = np.arange(1, len(classes))
        probs = (len(classes) - 1) / (sizes * (len(classes) - sizes))
        probs /= probs.sum()
        n_kept = rng.choice(sizes, size=1, p=probs)[0]
        order = np.arange(len(classes))
        rng.shuffle(order)
        removed_idx, remaining_idx = _class_split(labels, classes[order[n_kept:]])
        return remaining_idx, removed_idx
    possible_remaining_sizes = np.arange(1, dataset_size)
    remaining_size_probs = (dataset_size - 1) / (possible_remaining_sizes * (dataset_size - possible_remaining_sizes))
    remaining_size_probs /= remaining_size_probs.sum()
    remaining_size = rng.choice(possible_remaining_sizes, size=1, p=remaining_size_probs)[0]
    all_idx = np.arange(dataset_size)
    rng.shuffle(all_idx)
    return all_idx[:remaining_size], all_idx[remaining_size:]


def remove_data_by_uniform(dataset_size: int, seed: int = 0):
    """src/datasets.py:576-579: each unit kept with probability 1/2 (``rng.normal(size=n) > 0``)."""
    rng = np.random.RandomState(seed)
    selected = rng.normal(size=dataset_size) > 0
    all_idx = np.arange(dataset_size)
    return all_idx[selected], all_idx[~selected]


def remove_data_by_datamodel(dataset_size: int, alpha: float = 0.5, seed: int = 0, by_class: bool = False, labels=None):
    """src/datasets.py:619-626: the first int(alpha * n) of a shuffle.  ``by_class`` (src/datasets.py:603-617): the
    first int(alpha * n_classes) of a shuffle of the distinct labels (a Python list, shuffled in place like the
    reference's) are kept as whole classes."""
    rng = np.random.RandomState(seed)
    if by_class:
        classes = np.unique(np.asarray(labels)).tolist()
        n_kept = int(alpha * len(classes))
        rng.shuffle(classes)
        return _class_split(labels, classes[:n_kept])
    all_idx = np.arange(dataset_size)
    num_selected = int(alpha * dataset_size)
    rng.shuffle(all_idx)
    return all_idx[:num_selected], all_idx[num_selected:]


def remove_data_by_loo(dataset_size: int, loo_idx: int):
    """src/datasets.py:700-706."""
    return np.array([i for i in range(dataset_size) if i != loo_idx]), np.array([loo_idx])


def remove_data_for_aoi(dataset_size: int, aoi_idx: int):
    """src/datasets.py:709-715."""
    return np.array([aoi_idx]), np.array([i for i in range(dataset_size) if i != aoi_idx])


def masks_from_seeds(dataset_size: int, seeds, dist: str = "shapley", alpha: float = 0.5) -> np.ndarray:
    """Dense uint8 mask rows [len(seeds), dataset_size] for one removal distribution."""
    out = np.zeros((len(seeds), dataset_size), dtype=np.uint8)
    for r, seed in enumerate(seeds):
        if dist == "shapley":
            rem, _ = remove_data_by_shapley(dataset_size, seed)
        elif dist == "uniform":
            rem, _ = remove_data_by_uniform(dataset_size, seed)
        elif dist == "datamodel":
            rem, _ = remove_data_by_datamodel(dataset_size, alpha, seed)
        else:
            raise ValueError(f"removal distribution '{dist}' has to be one of shapley / uniform / datamodel")
        out[r, rem] = 1
    return out


def counterfactual_split(removal_rank, removal_rank_proportion: float | None = None,
                         removal_bottom_proportion: float | None = None):
    """text_to_image/train_text_to_image_lora.py:992-1006: split a rank file into (remaining_idx, removed_idx).

    Top-k removal takes the first ``floor(len(rank) * p)`` ranked units, bottom-k removal the last ones."""
    removal_rank = np.asarray(removal_rank)
    if removal_rank_proportion is not None:
        num_removed_units = math.floor(len(removal_rank) * removal_rank_proportion)
        removed_idx = removal_rank[:num_removed_units]
        remaining_idx = removal_rank[num_removed_units:]
    elif removal_bottom_proportion is not None:
        num_removed_units = math.floor(len(removal_rank) * removal_bottom_proportion)
        removed_idx = removal_rank[-num_removed_units:]
        remaining_idx = removal_rank[:-num_removed_units]
    else:
        raise ValueError("give removal_rank_proportion or removal_bottom_proportion")
    return remaining_idx, removed_idx
