"""End-to-end slice on the reference's own model shape: per-example diffusion-loss gradients of the 35 746 307-parameter
DDPM-CIFAR U-Net (PyTorch vmap(grad), as in d_trak_grad.py:689-776) -> projector -> TRAK scores."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def test_unet_restatement_has_the_reference_parameter_count():
    from ddpm_unet import DDPMCifarUNet, count_parameters

    assert count_parameters(DDPMCifarUNet()) == 35_746_307  # grad_dim passed to CudaProjector (d_trak_grad.py:505)


def test_featurize_dict_path_equals_flattened_path_and_scores_match_oracle():
    from ddpm_unet import DDPMCifarUNet, DDPMScheduler, count_parameters
    from featurize_and_score import featurize
    from gadm_b200 import CudaProjector, ProjectionType, trak_scores
    from oracle import scorer as oscore

    dev = torch.device("cuda:0")
    free, _ = torch.cuda.mem_get_info()
    if free < 30 * 2**30:
        pytest.skip("needs ~20 GB of free HBM")
    torch.manual_seed(0)
    model = DDPMCifarUNet().to(dev).eval()
    D = count_parameters(model)
    sched = DDPMScheduler(device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    images = torch.rand(12, 3, 32, 32, device=dev, generator=g) * 2 - 1
    proj = CudaProjector(D, 512, 42, ProjectionType.rademacher, dev, 4, stage_rows=64)
    phi = featurize(model, images, proj, sched, k_partition=2, opt_seed=42, behavior="loss", batch=4)
    assert phi.shape == (12, 512) and bool(torch.isfinite(phi).all())

    # reference-style path for the first batch: vectorize_and_ignore_buffers + emb / K + project (d_trak_grad.py:757-776)
    from torch.func import functional_call, grad, vmap
    import torch.nn.functional as F
    params = {k: v.detach() for k, v in model.named_parameters()}

    def compute_f(params, noisy, t, target):
        pred = functional_call(model, params, args=(noisy.unsqueeze(0), t.unsqueeze(0)))
        return F.mse_loss(pred.float(), target.unsqueeze(0).float(), reduction="none").mean()

    sg = vmap(grad(compute_f), in_dims=(None, 0, 0, 0))
    emb = None
    for t in range(0, 1000, 500):
        ts = torch.full((4,), t, device=dev, dtype=torch.long)
        torch.manual_seed(42 * 1000 + t)
        noise = torch.randn_like(images[:4])
        gr = sg(params, sched.add_noise(images[:4], noise, ts), ts, noise)
        flat = torch.stack([torch.cat([x[b].flatten() for x in gr.values()]) for b in range(4)])
        emb = flat if emb is None else emb + flat
    emb = emb / 2
    ref = proj.project(emb, model_id=0)
    # identical random matrix; inputs differ only by fp32 rounding of (a + b) / 2 vs (a + b) * 0.5 -> bitwise equal here
    scale = float(ref.abs().max())
    assert float((ref - phi[:4]).abs().max()) <= 1e-3 * scale
    # JL: ||phi|| / sqrt(k) ~ ||g||
    ratio = phi[:4].double().norm(dim=1) / 512 ** 0.5 / emb.double().norm(dim=1)
    assert float((ratio - 1).abs().max()) < 0.2

    scores = trak_scores(phi[:8], phi[8:], lam=0.5)
    want = oscore.score_fp64(phi[:8].cpu().numpy(), phi[8:].cpu().numpy(), 0.5)
    ref32 = oscore.score_torch(phi[:8].cpu().numpy(), phi[8:].cpu().numpy(), 0.5)  # the reference's own fp32 arithmetic
    # N = 8 < k = 512 with real gradient features: K = Phi^T Phi + 0.5 I has 8 large eigenvalues over a floor of 0.5,
    # so fp32 arithmetic (ours and traks.py's alike) loses digits; require fp32-grade agreement with the fp64 answer
    # and an error no worse than a few times the reference's own.
    for name in ("trak", "grad_sim"):
        got = scores[name].cpu().numpy().astype(np.float64)
        scale = np.abs(want[name]).max()
        ours = np.abs(got - want[name]).max() / scale
        theirs = np.abs(ref32[name].astype(np.float64) - want[name]).max() / scale
        cond = float(np.linalg.cond(phi[:8].double().cpu().numpy().T @ phi[:8].double().cpu().numpy() + 0.5 * np.eye(512)))
        assert ours <= max(8 * theirs, 2e-4), (name, ours, theirs, cond)


def test_journey_trak_featurisation_example():
    """Journey-TRAK end to end on the U-Net (grad_text_to_image_lora.py:485-545,729-770): trajectory latents at the
    journey points, noise-averaged gradients through DeferredProjection.accumulate, group.csv, journey_trak scores."""
    import pandas as pd
    import journey_trak as J
    from gadm_b200 import journey_point_indices

    free, _ = torch.cuda.mem_get_info()
    if free < 30 * 2**30:
        pytest.skip("needs ~20 GB of free HBM")
    np.testing.assert_array_equal(journey_point_indices(100, 50), np.arange(1, 100, 2))   # the reference's defaults
    np.testing.assert_array_equal(journey_point_indices(10, 3), np.array([1, 4, 7]))
    phi, idx, steps, scores, out_dir = J.main(["--n-train", "12", "--n-images", "2", "--num-inference-steps", "8",
                                               "--num-journey-points", "4", "--num-journey-noises", "2",
                                               "--proj-dim", "512", "--batch", "4"])
    want_steps = list(journey_point_indices(8, 4))
    assert steps == want_steps * 2 and idx == [0] * len(want_steps) + [1] * len(want_steps)
    assert phi.shape == (2 * len(want_steps), 512) and bool(torch.isfinite(phi).all())
    df = pd.read_csv(os.path.join(out_dir, "generated_journey", "group.csv"), index_col=0)
    assert df["generated_image_idx"].tolist() == idx and df["step_idx"].tolist() == steps
    assert scores["journey_trak"].shape == (12,) and bool(torch.isfinite(scores["journey_trak"]).all())
