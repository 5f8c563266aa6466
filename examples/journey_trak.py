"""Journey-TRAK featurisation on the reference's DDPM-CIFAR U-Net (random init): the generated images' sampling
trajectories are featurised at ``num_journey_points`` steps each, averaged over ``num_journey_noises`` noise draws,
projected, and scored against the training features with the journey-TRAK variant of traks.py:171-173.

Mirrors text_to_image/grad_text_to_image_lora.py:485-545 (trajectory latents through the sampler callback, journey
point selection, group.csv) and :729-770 (the noise loop `emb += grads ... emb / num_journey_noises; project`), with
the B200 path dropped in: the noise sum and mean live in the projector (`DeferredProjection.accumulate`), features
stay on the device.  The Stable-Diffusion pipeline of the reference is not available offline; the sampler here is a
plain ancestral DDPM loop over the same U-Net the unconditional scripts train (src/ddpm_config.py).

    python examples/journey_trak.py --n-train 32 --n-images 2 --num-inference-steps 20 --num-journey-points 5
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch
import torch.nn.functional as F
from torch.func import functional_call, grad, vmap

from ddpm_unet import DDPMCifarUNet, DDPMScheduler, count_parameters
from featurize_and_score import featurize
from gadm_b200 import (CudaProjector, ProjectionType, journey_point_indices, trak_scores, write_journey_group_csv)


@torch.no_grad()
def sample_trajectories(model, scheduler, n_images: int, num_inference_steps: int, seed: int, device):
    """Ancestral sampling; returns per image the lists (step_idx, t, latents) the reference's pipeline callback
    collects (grad_text_to_image_lora.py:492-506)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    timesteps = scheduler.inference_timesteps(num_inference_steps)
    out = []
    for _ in range(n_images):
        x = torch.randn(1, 3, 32, 32, device=device, generator=gen)
        steps, ts, lats = [], [], []
        for step_idx, t in enumerate(timesteps):
            tt = torch.full((1,), t, device=device, dtype=torch.long)
            eps = model(x, tt)
            t_prev = timesteps[step_idx + 1] if step_idx + 1 < len(timesteps) else -1
            x = scheduler.step(eps, t, t_prev, x, generator=gen)
            steps.append(step_idx)
            ts.append(t)
            lats.append(x.clone())
        out.append((steps, ts, lats))
    return out


def journey_features(model, scheduler, projector, trajectories, num_journey_points: int, num_journey_noises: int,
                     batch: int, output_dir: str | None = None, seed: int = 0):
    """[n_images * journey points, proj_dim] features + the (generated_image_idx, step_idx) groups."""
    device = next(model.parameters()).device
    params = {k: v.detach() for k, v in model.named_parameters() if v.requires_grad}
    buffers = {k: v.detach() for k, v in model.named_buffers()}

    def compute_f(params, buffers, noisy_latents, timesteps, targets):  # f = loss (grad_text_to_image_lora.py:668-700)
        pred = functional_call(model, (params, buffers), args=(noisy_latents.unsqueeze(0), timesteps.unsqueeze(0)))
        return F.mse_loss(pred.float(), targets.unsqueeze(0).float(), reduction="none").mean()

    sample_grad = vmap(grad(compute_f), in_dims=(None, None, 0, 0, 0))
    all_idx, all_step, all_t, all_lat = [], [], [], []
    for i, (steps, ts, lats) in enumerate(trajectories):
        for j in journey_point_indices(len(steps), num_journey_points):
            all_idx.append(i)
            all_step.append(steps[j])
            all_t.append(ts[j])
            all_lat.append(lats[j])
    if output_dir is not None:
        write_journey_group_csv(output_dir, all_idx, all_step)
    latents = torch.cat(all_lat)
    t_all = torch.tensor(all_t, device=device, dtype=torch.long)
    gen = torch.Generator(device=device).manual_seed(seed)
    sink = projector.deferred(model_id=0)
    for lo in range(0, latents.shape[0], batch):
        lat, tt = latents[lo:lo + batch], t_all[lo:lo + batch]
        for index_noise in range(num_journey_noises):  # emb += grads; emb / num_journey_noises; project  (:738-765)
            noise = torch.randn(lat.shape, device=device, generator=gen)
            noisy = scheduler.add_noise(lat, noise, tt)
            g = sample_grad(params, buffers, noisy, tt, noise)
            sink.accumulate(g, scale=1.0 / num_journey_noises, last=(index_noise == num_journey_noises - 1))
    return sink.result(), all_idx, all_step


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-train", type=int, default=32)
    ap.add_argument("--n-images", type=int, default=2)
    ap.add_argument("--num-inference-steps", type=int, default=20)
    ap.add_argument("--num-journey-points", type=int, default=5)
    ap.add_argument("--num-journey-noises", type=int, default=2)
    ap.add_argument("--proj-dim", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--output-dir", default=None)
    a = ap.parse_args(argv)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = DDPMCifarUNet().to(dev).eval()
    n_params = count_parameters(model)
    scheduler = DDPMScheduler(device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    train = torch.rand(a.n_train, 3, 32, 32, device=dev, generator=g) * 2 - 1
    projector = CudaProjector(grad_dim=n_params, proj_dim=a.proj_dim, seed=42, proj_type=ProjectionType.normal, device=dev,
                              max_batch_size=a.batch, stage_rows=max(64, a.batch))
    out_dir = a.output_dir or tempfile.mkdtemp(prefix="journey_")
    phi_train = featurize(model, train, projector, scheduler, 2, 42, "loss", a.batch)
    traj = sample_trajectories(model, scheduler, a.n_images, a.num_inference_steps, seed=7, device=dev)
    finals = torch.cat([lats[-1] for _, _, lats in traj]).clamp(-1, 1)  # source == "generated": the final latent only
    phi_gen = featurize(model, finals, projector, scheduler, 2, 42, "loss", a.batch)
    phi_journey, idx, steps = journey_features(model, scheduler, projector, traj, a.num_journey_points, a.num_journey_noises,
                                               a.batch, output_dir=os.path.join(out_dir, "generated_journey"))
    scores = trak_scores(phi_train, phi_gen, lam=0.5, journey_phi=phi_journey)
    print(f"params {n_params}; train {tuple(phi_train.shape)}, generated {tuple(phi_gen.shape)}, journey "
          f"{tuple(phi_journey.shape)} ({a.n_images} images x {len(steps) // a.n_images} points, group.csv in {out_dir}); "
          f"top-5 journey-TRAK contributors {torch.argsort(-scores['journey_trak'])[:5].tolist()}")
    return phi_journey, idx, steps, scores, out_dir


if __name__ == "__main__":
    main()
