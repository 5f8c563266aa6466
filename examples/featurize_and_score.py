"""End-to-end D-TRAK / TRAK on the reference's DDPM-CIFAR U-Net (random init, synthetic images).

Mirrors the hot loop of src/attributions/methods/d_trak_grad.py:700-794 with the B200 path dropped in:
  * per-example gradients: torch.func.vmap(grad(compute_f)) exactly as the reference (PyTorch, by design);
  * timestep sum and mean inside the projector (`DeferredProjection.accumulate(g, scale=1/K, last=...)`: an fp32 slab
    summed by one kernel per timestep, staged on the last one; d_trak_grad.py:764-770);
  * the per-parameter gradient dict goes straight to `CudaProjector` (no vectorize_and_ignore_buffers copy);
  * features stay on the device; `trak_scores` (Gram -> Cholesky -> solve -> score GEMM) replaces traks.py:141-186.

    python examples/featurize_and_score.py --n-train 64 --n-gen 8 --k-partition 2 --proj-dim 2048
"""
from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch
import torch.nn.functional as F
from torch.func import functional_call, grad, vmap

from ddpm_unet import DDPMCifarUNet, DDPMScheduler, count_parameters
from gadm_b200 import CudaProjector, ProjectionType, trak_scores


def featurize(model, images, projector, scheduler, k_partition: int, opt_seed: int, behavior: str, batch: int,
              sink=None, project: bool = True):
    """[N, 3, 32, 32] images -> [N, proj_dim] features (device resident).  ``project=False`` runs the gradient
    producer alone (bench.py uses it to split the wall time)."""
    params = {k: v.detach() for k, v in model.named_parameters() if v.requires_grad}
    buffers = {k: v.detach() for k, v in model.named_buffers()}

    def compute_f(params, buffers, noisy_latents, timesteps, targets):  # d_trak_grad.py:668-687 / :523-553
        pred = functional_call(model, (params, buffers), args=(noisy_latents.unsqueeze(0), timesteps.unsqueeze(0)))
        if behavior == "loss":
            return F.mse_loss(pred.float(), targets.unsqueeze(0).float(), reduction="none").mean()
        return F.mse_loss(pred.float(), torch.zeros_like(targets).unsqueeze(0).float(), reduction="none").mean()

    sample_grad = vmap(grad(compute_f), in_dims=(None, None, 0, 0, 0))
    selected_timesteps = list(range(0, 1000, 1000 // k_partition))  # t_strategy == "uniform" (d_trak_grad.py:718-719)
    own_sink = sink is None
    if own_sink and project:
        sink = projector.deferred(model_id=0)
    for i in range(0, images.shape[0], batch):
        image = images[i:i + batch]
        for j, t in enumerate(selected_timesteps):
            timesteps = torch.full((image.shape[0],), t, device=image.device, dtype=torch.long)
            torch.manual_seed(opt_seed * 1000 + t)  # seed_everything(args.opt_seed * 1000 + t) (d_trak_grad.py:727)
            noise = torch.randn_like(image)
            noisy = scheduler.add_noise(image, noise, timesteps)
            g = sample_grad(params, buffers, noisy, timesteps, noise)
            if project:  # emb += g; emb / K on the last timestep; stage (d_trak_grad.py:764-770)
                sink.accumulate(g, scale=1.0 / k_partition, last=(j == len(selected_timesteps) - 1))
    return sink.result() if (project and own_sink) else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-train", type=int, default=64)
    ap.add_argument("--n-gen", type=int, default=8)
    ap.add_argument("--k-partition", type=int, default=2)
    ap.add_argument("--proj-dim", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--opt-seed", type=int, default=42)
    ap.add_argument("--proj-type", default="normal")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = DDPMCifarUNet().to(dev).eval()
    n_params = count_parameters(model)
    assert n_params == 35_746_307
    scheduler = DDPMScheduler(device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    train = torch.rand(a.n_train, 3, 32, 32, device=dev, generator=g) * 2 - 1
    gen = torch.rand(a.n_gen, 3, 32, 32, device=dev, generator=g) * 2 - 1
    projector = CudaProjector(grad_dim=n_params, proj_dim=a.proj_dim, seed=a.opt_seed,
                              proj_type=ProjectionType(a.proj_type), device=dev, max_batch_size=a.batch)
    t0 = time.time()
    phi_train = featurize(model, train, projector, scheduler, a.k_partition, a.opt_seed, "loss", a.batch)
    phi_gen = featurize(model, gen, projector, scheduler, a.k_partition, a.opt_seed, "loss", a.batch)
    torch.cuda.synchronize()
    t1 = time.time()
    scores = trak_scores(phi_train, phi_gen, lam=0.5)
    torch.cuda.synchronize()
    t2 = time.time()
    print(f"params {n_params}; features {tuple(phi_train.shape)} / {tuple(phi_gen.shape)} in {t1 - t0:.2f} s; "
          f"scores in {(t2 - t1) * 1e3:.1f} ms; top-5 TRAK contributors {torch.argsort(-scores['trak'])[:5].tolist()}")
    return phi_train, phi_gen, scores


if __name__ == "__main__":
    main()
