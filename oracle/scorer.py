"""TRAK / D-TRAK scorer oracles.  TEST ORACLE (see oracle/__init__.py).

``score_numpy``   restates src/attributions/methods/compute_gradient_score.py:102-132
                  (fp32 Gram -> + 0.5*I in fp64 -> np.linalg.inv -> fp64 scores).
``score_torch``   restates text_to_image/traks.py:141-186 on CPU tensors (fp32
                  throughout, torch.inverse).
``score_fp64``    the same algebra entirely in float64 -- the yardstick the GPU path's
                  error is compared with (SURVEY.md section 7 "Gram/solve in fp32").
``group_aggregate`` restates text_to_image/traks.py:188-204.
``aggregate_by_class`` restates src/attributions/methods/attribution_utils.py:15-48
                  (including the reference's "max over all rows" behaviour at :46).
``stable_rank``   restates text_to_image/traks.py:216-218 and
                  text_to_image/shapley_lds.py:294.
"""
from __future__ import annotations

import numpy as np


def score_numpy(train_phi: np.ndarray, val_phi: np.ndarray, gradient_type: str = "trak", lam: float = 5e-1,
                average: bool = True):
    """compute_gradient_score.py:102-132.  Returns (scores_or_coeff, kernel_inverse)."""
    train_phi = np.asarray(train_phi, dtype=np.float32)
    val_phi = np.asarray(val_phi, dtype=np.float32)
    kernel = train_phi.T @ train_phi  # fp32 (:108)
    kernel = kernel + lam * np.eye(kernel.shape[0])  # -> fp64 (:109)
    kernel = np.linalg.inv(kernel)  # (:110)
    if gradient_type == "vanilla_gradient":  # (:114-117)
        tp = train_phi / np.linalg.norm(train_phi, axis=1, keepdims=True)
        vp = val_phi / np.linalg.norm(val_phi, axis=1, keepdims=True)
        scores = np.dot(vp, tp.T)
    else:
        if gradient_type == "relative_if":  # (:119-120)
            magnitude = np.linalg.norm((train_phi @ kernel).T, axis=0)
        elif gradient_type == "renormalized_if":  # (:121-122)
            magnitude = np.linalg.norm(train_phi.T, axis=0)
        else:
            magnitude = 1.0
        scores = val_phi @ ((train_phi @ kernel).T) / magnitude  # (:126)
    if average:
        return np.mean(scores, axis=0), kernel  # (:130)
    return scores, kernel


def score_fp64(train_phi, val_phi, lam: float = 5e-1):
    """All-float64 version: returns dict with the full score matrix and the variants."""
    tp = np.asarray(train_phi, dtype=np.float64)
    vp = np.asarray(val_phi, dtype=np.float64)
    K = tp.T @ tp + lam * np.eye(tp.shape[1])
    W = np.linalg.solve(K, tp.T)  # [k, N]
    S = vp @ W  # [T, N]
    out = {
        "scores": S,
        "trak": S.mean(axis=0),
        "relative_influence": (S / np.linalg.norm(W, axis=0)).mean(axis=0),
        "renorm_influence": (S / np.linalg.norm(tp, axis=1)).mean(axis=0),
    }
    cos = (vp @ tp.T) / (np.linalg.norm(vp, axis=1, keepdims=True) * np.linalg.norm(tp, axis=1, keepdims=True).T)
    out["grad_sim"] = cos.mean(axis=0)
    out["cosine"] = cos
    return out


def score_torch(train_grads, gen_grads, lam: float = 5e-1, journey_grads=None):
    """text_to_image/traks.py:141-173 with CPU tensors (fp32)."""
    import torch

    train_grads = torch.as_tensor(train_grads, dtype=torch.float32)
    gen_grads = torch.as_tensor(gen_grads, dtype=torch.float32)
    out = {}
    grad_sim = torch.matmul(gen_grads, train_grads.T)
    grad_sim /= torch.matmul(gen_grads.norm(dim=-1, keepdim=True), train_grads.norm(dim=-1, keepdim=True).T)
    out["grad_sim"] = grad_sim.mean(dim=0).numpy()
    ihdp = torch.matmul(train_grads.T, train_grads)
    ihdp += lam * torch.eye(train_grads.shape[1])
    ihdp = torch.inverse(ihdp)
    ihdp = torch.matmul(ihdp, train_grads.T)  # proj_dim x train_size
    trak = torch.matmul(gen_grads, ihdp)
    out["trak"] = trak.mean(dim=0).numpy()
    influence = torch.matmul(gen_grads, ihdp)
    out["relative_influence"] = (influence / ihdp.norm(dim=0)).mean(dim=0).numpy()
    out["renorm_influence"] = (influence / train_grads.norm(dim=-1)).mean(dim=0).numpy()
    if journey_grads is not None:
        journey_grads = torch.as_tensor(journey_grads, dtype=torch.float32)
        out["journey_trak"] = torch.matmul(journey_grads, ihdp).mean(dim=0).numpy()
    return out


def group_aggregate(sample_output_dict: dict, group_indices_dict: dict):
    """text_to_image/traks.py:188-204: per-group sum (TRAK family) or mean/max (grad_sim)."""
    num_groups = len(group_indices_dict.keys())
    output_dict = {}
    for method, attrs in sample_output_dict.items():
        if method in ["grad_sim"]:
            group_avg_attrs = np.zeros(shape=(num_groups, 1))
            group_max_attrs = np.zeros(shape=(num_groups, 1))
            for i, group_indices in group_indices_dict.items():
                group_avg_attrs[i, :] = attrs[group_indices].mean()
                group_max_attrs[i, :] = attrs[group_indices].max()
            output_dict[f"avg_{method}"] = group_avg_attrs
            output_dict[f"max_{method}"] = group_max_attrs
        else:
            group_attrs = np.zeros(shape=(num_groups, 1))
            for i, group_indices in group_indices_dict.items():
                group_attrs[i, :] = attrs[group_indices].sum()
            output_dict[method] = group_attrs
    return output_dict


def aggregate_by_class(scores: np.ndarray, labels: np.ndarray, by: str = "mean"):
    """attribution_utils.py:15-48 with the dataset replaced by its label vector."""
    scores = np.asarray(scores)
    if scores.ndim == 1:
        scores = scores.reshape(1, -1)
    n, _ = scores.shape
    unique_values = sorted(set(labels.tolist()))
    value_to_number = {value: i for i, value in enumerate(unique_values)}
    lab = np.array([value_to_number[v] for v in labels.tolist()])
    num_labels = len(np.unique(lab))
    result = np.zeros((n, num_labels))
    for i in range(num_labels):
        label_mask = lab == i
        if by == "mean":
            result[:, i] = np.divide(scores[:, label_mask].sum(axis=1), np.sum(label_mask))
        elif by == "max":
            result[:, i] = np.max(scores[:, label_mask])  # sic: over all rows (:46)
    return result


def stable_rank(output: np.ndarray) -> np.ndarray:
    """np.argsort(-x.mean(axis=-1), kind='stable') (traks.py:218, shapley_lds.py:294)."""
    return np.argsort(-np.asarray(output).mean(axis=-1), kind="stable")
