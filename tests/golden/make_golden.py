"""Generate golden vectors by running the reference's OWN code (build container only).

Run:  python tests/golden/make_golden.py         (needs /root/reference; writes *.npz here)

The reference has no tests or fixtures, so the pins are made by
executing its functions on seeded synthetic inputs:

* ``data_shapley`` / ``data_banzhaf``  -- imported from
  src/attributions/methods/{datashapley,databanzhaf}.py
* ``evaluate_lds`` / ``collect_data``  -- imported from text_to_image/shapley_lds.py,
  text_to_image/banzhaf_lds.py and lds.py (``attrs_all[k]`` convention)
* ``text_to_image/traks.py:main``      -- executed end to end on synthetic ``.pt``
  feature files (``.to("cuda")`` redirected to CPU: this container has no GPU)
* ``compute_gradient_scores``          -- src/attributions/methods/compute_gradient_score.py
  executed on synthetic memmaps with the dataset classes stubbed

``src/constants.py`` is a user-created file (README.md:19-28); it is injected as a
module with temp-dir paths, nothing under /root/reference is modified or copied.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _inject_env(tmp):
    sys.path.insert(0, REF)
    c = types.ModuleType("src.constants")
    c.DATASET_DIR = os.path.join(tmp, "datasets")
    c.OUTDIR = os.path.join(tmp, "out")
    c.LOGDIR = os.path.join(tmp, "log")
    c.MAX_NUM_SAMPLE_IMAGES_TO_SAVE = 64
    c.DATASET = ["cifar", "cifar2", "celeba", "mnist", "cifar100", "cifar100_f"]
    c.METHOD = ["retrain", "gd", "ga", "iu", "lora"]
    sys.modules["src.constants"] = c
    # src/utils.py pulls in the vendored diffusers patch (needs the absent `diffusers` package); the
    # scripts only take `print_args` from it, which is not on the hot path -> stub that one helper.
    u = types.ModuleType("src.utils")
    u.print_args = lambda args: None
    u.get_max_steps = lambda *a, **k: None
    sys.modules["src.utils"] = u
    for name in ("clip", "lightning", "lightning.pytorch", "matplotlib", "matplotlib.pyplot"):  # unused on this path
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    return c


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def shapley_kernel_masks(d, seeds):
    """src/datasets.py:677-697 executed through the reference function."""
    ds = _load("src/datasets.py", "ref_datasets")
    out = np.zeros((len(seeds), d))
    dummy = list(range(d))
    for r, s in enumerate(seeds):
        rem, _ = ds.remove_data_by_shapley(dummy, seed=s)
        out[r, rem] = 1
    uni = np.zeros((len(seeds), d))
    for r, s in enumerate(seeds):
        rem, _ = ds.remove_data_by_uniform(dummy, seed=s)
        uni[r, rem] = 1
    dm = np.zeros((len(seeds), d))
    for r, s in enumerate(seeds):
        rem, _ = ds.remove_data_by_datamodel(dummy, alpha=0.5, seed=s)
        dm[r, rem] = 1
    return out, uni, dm


def make_aggregation(tmp):
    from src.attributions.methods.databanzhaf import data_banzhaf
    from src.attributions.methods.datashapley import data_shapley

    shap_lds = _load("text_to_image/shapley_lds.py", "ref_shapley_lds")
    banz_lds = _load("text_to_image/banzhaf_lds.py", "ref_banzhaf_lds")

    out = {}
    cases = {"a": (64, 12, 5, 20), "b": (200, 30, 7, 40), "c": (20, 30, 4, 25)}  # c: n < d (rank deficient)
    for name, (n, d, K, m) in cases.items():
        rng = np.random.RandomState(100 + n)
        Xs, Xu, Xd = shapley_kernel_masks(d, list(range(n)))
        w = rng.normal(size=(d, K))
        Ys = Xs @ w + 0.1 * rng.normal(size=(n, K))
        Yu = Xu @ w + 0.1 * rng.normal(size=(n, K))
        v0 = 0.05 * rng.normal(size=K)
        v1 = w.sum(axis=0) + v0
        phi_s = np.stack([data_shapley(d, Xs, Ys[:, k], v1[k], v0[k]).flatten() for k in range(K)], axis=1)
        phi_b = np.stack([data_banzhaf(Xu, Yu[:, k]) for k in range(K)], axis=1)
        tests = []
        for t in range(3):
            _, _, Xt = shapley_kernel_masks(d, list(range(1000 * (t + 1), 1000 * (t + 1) + m)))
            Yt = Xt @ w + 0.3 * rng.normal(size=(m, K))
            if t == 1:  # ties in the behaviours (exercise average ranks)
                Yt = np.round(Yt, 0)
            tests.append((Xt, Yt))
        lds_s = shap_lds.evaluate_lds(phi_s, tests, K)
        lds_b = banz_lds.evaluate_lds(phi_b, tests, K)
        out.update({
            f"{name}_Xs": Xs, f"{name}_Xu": Xu, f"{name}_Ys": Ys, f"{name}_Yu": Yu, f"{name}_v0": v0, f"{name}_v1": v1,
            f"{name}_phi_shapley": phi_s, f"{name}_phi_banzhaf": phi_b,
            f"{name}_lds_shapley": np.array(lds_s), f"{name}_lds_banzhaf": np.array(lds_b),
        })
        for t, (Xt, Yt) in enumerate(tests):
            out[f"{name}_Xt{t}"] = Xt
            out[f"{name}_Yt{t}"] = Yt

    # collect_data (shapley_lds.py:105-135) on a synthetic jsonl db
    import pandas as pd

    d, n_samples = 9, 3
    rng = np.random.RandomState(7)
    recs = []
    for s in range(6):
        rem = sorted(rng.choice(d, size=rng.randint(1, d), replace=False).tolist())
        rec = {"exp_name": f"x_seed_{s}", "remaining_idx": rem}
        for i in range(n_samples):
            rec[f"generated_image_{i}_ssim"] = float(rng.normal())
        recs.append(rec)
    db = os.path.join(tmp, "db.jsonl")
    with open(db, "w") as f:
        for r in recs:
            f.write(json.dumps(r) + "\n")
    df = pd.read_json(db, lines=True)
    X, Y = shap_lds.collect_data(df=df, num_groups=d, model_behavior_key="ssim", n_samples=n_samples)
    out["collect_X"] = X
    out["collect_Y"] = Y
    out["collect_db"] = np.frombuffer(open(db, "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "aggregation_golden.npz"), **out)
    print("aggregation_golden.npz", {k: v.shape for k, v in out.items() if k.startswith("a_")})


def make_traks(tmp, consts):
    """Run text_to_image/traks.py:main on synthetic features (CPU)."""
    import pandas as pd
    import torch

    N, T, J, k, G = 96, 6, 10, 64, 7
    g = torch.Generator().manual_seed(5)
    feats = {
        "train_loss": torch.randn(N, k, generator=g),
        "train_dtrak": torch.randn(N, k, generator=g),
        "gen_loss": torch.randn(T, k, generator=g),
        "gen_dtrak": torch.randn(T, k, generator=g),
        "journey": torch.randn(T * J, k, generator=g),
    }
    out_dir = os.path.join(tmp, "t2i")
    gd = os.path.join(out_dir, "gradients")
    for sub in ("train", "generated", "generated_journey"):
        os.makedirs(os.path.join(gd, sub), exist_ok=True)
    sfx = f"num_timesteps=100_proj_dim={k}.pt"
    torch.save(feats["train_loss"], os.path.join(gd, "train", f"emb_f=loss_{sfx}"))
    torch.save(feats["train_dtrak"], os.path.join(gd, "train", f"emb_f=mean-squared-l2-norm_{sfx}"))
    torch.save(feats["gen_loss"], os.path.join(gd, "generated", f"emb_f=loss_{sfx}"))
    torch.save(feats["gen_dtrak"], os.path.join(gd, "generated", f"emb_f=mean-squared-l2-norm_{sfx}"))
    torch.save(feats["journey"], os.path.join(
        gd, "generated_journey", f"emb_f=loss_num_journey_points=50_num_journey_noises=1_proj_dim={k}.pt"))
    rng = np.random.RandomState(3)
    groups = rng.randint(0, G, size=N)
    groups[:G] = np.arange(G)
    pd.DataFrame({"artist": [f"artist_{a}" for a in groups]}).to_csv(os.path.join(gd, "train", "group.csv"), index=False)
    ddir = os.path.join(consts.DATASET_DIR, "artbench-10-imagefolder-split", "train")
    os.makedirs(ddir, exist_ok=True)
    pd.DataFrame({"artist": [f"artist_{a}" for a in range(G)]}).to_csv(
        os.path.join(ddir, "post_impressionism_artists.csv"), index=False)

    # no GPU here: send every .to("cuda") to the CPU
    orig_to = torch.Tensor.to

    def to_cpu(self, *a, **kw):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        return orig_to(self, *a, **kw)

    torch.Tensor.to = to_cpu
    try:
        traks = _load("text_to_image/traks.py", "ref_traks")
        args = argparse.Namespace(output_dir=out_dir, num_timesteps=100, proj_dim=k, dataset="artbench",
                                  cls="post_impressionism", group="artist", lam=5e-1,
                                  gradient_dir=gd)
        traks.main(args)
    finally:
        torch.Tensor.to = orig_to
    res = {f"in_{n}": v.numpy() for n, v in feats.items()}
    res["in_groups"] = groups
    bdir = os.path.join(out_dir, "baselines")
    for fn in sorted(os.listdir(bdir)):
        res["out_" + fn[:-4]] = np.load(os.path.join(bdir, fn))
    np.savez_compressed(os.path.join(HERE, "traks_golden.npz"), **res)
    print("traks_golden.npz", sorted(k for k in res if k.startswith("out_")))


def make_gradient_scores(tmp, consts):
    """Run compute_gradient_scores (compute_gradient_score.py:13-139) on synthetic memmaps."""
    cgs = _load("src/attributions/methods/compute_gradient_score.py", "ref_cgs")
    N, T, k, C = 120, 9, 32, 5
    rng = np.random.RandomState(11)
    train = rng.normal(size=(N, k)).astype(np.float32)
    val = rng.normal(size=(T, k)).astype(np.float32)
    labels = rng.randint(0, C, size=N)
    labels[:C] = np.arange(C)
    dataset = [(None, int(l)) for l in labels]
    cgs.create_dataset = lambda dataset_name, train: dataset
    cgs.ImageDataset = lambda sample_dir: list(range(T))
    res = {"in_train": train, "in_val": val, "in_labels": labels}
    for gtype in ("trak", "d_trak", "relative_if", "renormalized_if", "vanilla_gradient"):
        behavior = "mean-squared-l2-norm" if gtype == "d_trak" else "loss"
        sample_dir = os.path.join(tmp, "samples")
        vdir = os.path.join(sample_dir, "d_trak")
        tdir = os.path.join(consts.OUTDIR, "cifar100", "d_trak", "full")
        os.makedirs(vdir, exist_ok=True)
        os.makedirs(tdir, exist_ok=True)
        val.tofile(os.path.join(vdir, f"reference_f={behavior}_t=uniform_k=10_d={k}"))
        train.tofile(os.path.join(tdir, f"train_f={behavior}_t=uniform_k=10_d={k}"))
        kp = os.path.join(tdir, f"kernel_train_f={behavior}_t=uniform_k=10_d={k}.npy")
        if os.path.exists(kp):
            os.remove(kp)
        for by_class in (False, True):
            args = argparse.Namespace(dataset="cifar100", sample_dir=sample_dir, gradient_type=gtype, k_partition=10,
                                      projector_dim=k, sample_size=T, model_behavior_key="fid", by_class=by_class,
                                      by="mean")
            res[f"out_{gtype}_byclass={int(by_class)}"] = np.asarray(cgs.compute_gradient_scores(args))
        res[f"out_{gtype}_kernel"] = np.load(kp)
    np.savez_compressed(os.path.join(HERE, "gradient_scores_golden.npz"), **res)
    print("gradient_scores_golden.npz", {k: v.shape for k, v in res.items() if k.startswith("out_")})


def make_lds_py(tmp):
    """lds.py:158-170 evaluate_lds with the ``attrs_all[k]`` convention."""
    import src.attributions.methods.databanzhaf  # noqa: F401

    try:
        lds = _load("lds.py", "ref_lds")
    except Exception as e:  # heavy imports (torchvision datasets etc.)
        print("lds.py not importable here:", type(e).__name__, e)
        return
    rng = np.random.RandomState(21)
    d, K, m = 15, 6, 30
    attrs = [rng.normal(size=d) for _ in range(K)]
    tests = [(rng.binomial(1, 0.5, size=(m, d)).astype(float), rng.normal(size=(m, K))) for _ in range(3)]
    mean, ci = lds.evaluate_lds(attrs, tests, K)
    res = {"attrs": np.stack(attrs), "lds": np.array([mean, ci])}
    for t, (x, y) in enumerate(tests):
        res[f"Xt{t}"], res[f"Yt{t}"] = x, y
    np.savez_compressed(os.path.join(HERE, "lds_py_golden.npz"), **res)
    print("lds_py_golden.npz", mean, ci)


def make_ridge():
    """lds.py:411-421: the datamodel branch is a plain sklearn call per behaviour; executed here as written there
    (`RidgeCV(alphas=np.linspace(0.01, 10, 100)).fit(train_masks_fold, train_targets_fold[:, i])`), on masks drawn
    by the reference's own datamodel sampler (src/datasets.py:619-626 logic via oracle.aggregation.datamodel_masks,
    itself pinned to the reference in aggregation_golden.npz)."""
    import sklearn
    from sklearn.linear_model import RidgeCV

    from oracle.aggregation import datamodel_masks

    res = {"sklearn_version": np.array(sklearn.__version__)}
    rng = np.random.RandomState(7)
    for tag, n, d, K, noise in (("long", 300, 40, 6, 0.3), ("wide", 24, 40, 4, 0.05), ("square", 40, 40, 3, 1.0)):
        X = datamodel_masks(d, list(range(1000, 1000 + n)), alpha=0.5)
        w = rng.normal(size=(d, K))
        Y = X @ w + noise * rng.normal(size=(n, K)) + 3.0
        coefs, alphas, ics = [], [], []
        for i in range(K):
            m = RidgeCV(alphas=np.linspace(0.01, 10, 100)).fit(X, Y[:, i])
            coefs.append(m.coef_); alphas.append(m.alpha_); ics.append(m.intercept_)
        res[f"{tag}_X"] = X.astype(np.uint8)
        res[f"{tag}_Y"] = Y
        res[f"{tag}_coef"] = np.stack(coefs, axis=1)
        res[f"{tag}_alpha"] = np.asarray(alphas)
        res[f"{tag}_intercept"] = np.asarray(ics)
    np.savez_compressed(os.path.join(HERE, "ridge_golden.npz"), **res)
    print("ridge_golden.npz", {k: v.shape for k, v in res.items()})


def make_datamodel():
    """src/attributions/methods/datamodel.py:8-37 executed through the reference's own function (bootstrapped
    RidgeCV(cv=5)); the module's `from src.datasets import create_dataset` is satisfied by the injected environment."""
    dm = _load("src/attributions/methods/datamodel.py", "ref_datamodel")
    from oracle.aggregation import datamodel_masks

    res = {}
    for tag, n, d, runs, noise, seed in (("wide", 60, 200, 4, 0.3, 11), ("tall", 150, 40, 3, 1.0, 12), ("odd", 23, 64, 2, 0.1, 13)):
        rng = np.random.RandomState(seed)
        X = datamodel_masks(d, list(range(3000, 3000 + n)), alpha=0.5)
        Y = X @ rng.normal(size=d) + noise * rng.normal(size=n)
        np.random.seed(seed)
        coeff = dm.datamodel(X, Y, runs)
        res[f"{tag}_X"] = X.astype(np.uint8)
        res[f"{tag}_Y"] = Y
        res[f"{tag}_seed"] = np.array(seed)
        res[f"{tag}_coeff"] = coeff
    np.savez_compressed(os.path.join(HERE, "datamodel_golden.npz"), **res)
    print("datamodel_golden.npz", {k: v.shape for k, v in res.items()})


def make_masks_by_class():
    """src/datasets.py:603-617 (datamodel, by_class) and :651-673 (Shapley kernel, by_class) through the reference's
    own functions, on a dataset of (x, label) tuples with unequal class sizes."""
    ds = _load("src/datasets.py", "ref_datasets")
    rng = np.random.RandomState(123)
    labels = rng.choice(10, size=200, p=np.arange(1, 11) / 55.0)
    dataset = [(None, int(c)) for c in labels]
    res = {"labels": labels}
    for seed in range(6):
        rem, removed = ds.remove_data_by_datamodel(dataset, alpha=0.5, seed=seed, by_class=True)
        res[f"datamodel_rem_{seed}"], res[f"datamodel_removed_{seed}"] = np.asarray(rem), np.asarray(removed)
        rem, removed = ds.remove_data_by_shapley(dataset, seed=seed, by_class=True)
        res[f"shapley_rem_{seed}"], res[f"shapley_removed_{seed}"] = np.asarray(rem), np.asarray(removed)
    np.savez_compressed(os.path.join(HERE, "masks_by_class_golden.npz"), **res)
    print("masks_by_class_golden.npz", len(res))


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))  # repo root (oracle/)
    only = sys.argv[1] if len(sys.argv) > 1 else None
    if only == "masks_by_class":
        with tempfile.TemporaryDirectory() as tmp:
            _inject_env(tmp)
            make_masks_by_class()
        sys.exit(0)
    if only == "ridge":  # regenerate only the sklearn-based fixture
        make_ridge()
        sys.exit(0)
    if only == "datamodel":
        with tempfile.TemporaryDirectory() as tmp:
            _inject_env(tmp)
            make_datamodel()
        sys.exit(0)
    with tempfile.TemporaryDirectory() as tmp:
        consts = _inject_env(tmp)
        make_aggregation(tmp)
        make_traks(tmp, consts)
        make_gradient_scores(tmp, consts)
        make_lds_py(tmp)
        make_datamodel()
        make_masks_by_class()
    make_ridge()
