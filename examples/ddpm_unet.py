"""PyTorch restatement of the reference's DDPM-CIFAR U-Net and scheduler (gradient producer for the examples).

`diffusers` is not installable here, so this follows the `UNet2DModel` configuration the reference uses
(src/ddpm_config.py:48-82: block_out_channels [128, 256, 256, 256], two ResNet blocks per level, attention at the
second level and in the mid block, GroupNorm(32, eps 1e-6), SiLU, sinusoidal timestep embedding with
flip_sin_to_cos=False / freq_shift=1, stride-2 convolution downsampling with padding 0, nearest + conv upsampling)
and `DDPMScheduler.add_noise` (src/ddpm_config.py:83-100: linear betas 1e-4 .. 0.02, 1000 steps).
Acceptance check: 35 746 307 parameters.  Gradient production stays in PyTorch by design
(BASELINE north_star); only the projection / scoring / aggregation run on the gadm kernels.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def timestep_embedding(t: torch.Tensor, dim: int = 128, freq_shift: float = 1.0) -> torch.Tensor:
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - freq_shift)
    emb = t.float()[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)  # flip_sin_to_cos=False


class ResnetBlock(nn.Module):
    def __init__(self, cin, cout, temb=512, groups=32, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class Attention(nn.Module):
    """Single-head self-attention over the spatial positions (attention_head_dim=None -> one head of width C)."""

    def __init__(self, c, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q, self.to_k, self.to_v, self.to_out = (nn.Linear(c, c) for _ in range(4))

    def forward(self, x):
        b, c, hh, ww = x.shape
        h = self.group_norm(x).view(b, c, hh * ww).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        attn = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c), dim=-1)
        h = self.to_out(attn @ v).transpose(1, 2).reshape(b, c, hh, ww)
        return x + h


class Downsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))  # downsample_padding=0 -> asymmetric pad


class Upsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DDPMCifarUNet(nn.Module):
    def __init__(self, channels=(128, 256, 256, 256), attn_levels=(1,), in_ch=3, layers=2):
        super().__init__()
        c0 = channels[0]
        self.conv_in = nn.Conv2d(in_ch, c0, 3, padding=1)
        self.time_embedding = nn.Sequential(nn.Linear(c0, 4 * c0), nn.SiLU(), nn.Linear(4 * c0, 4 * c0))
        temb = 4 * c0
        self.down = nn.ModuleList()
        cin = c0
        for lvl, cout in enumerate(channels):
            blk = nn.ModuleDict({"resnets": nn.ModuleList(), "attns": nn.ModuleList()})
            for j in range(layers):
                blk["resnets"].append(ResnetBlock(cin if j == 0 else cout, cout, temb))
                if lvl in attn_levels:
                    blk["attns"].append(Attention(cout))
            if lvl != len(channels) - 1:
                blk["down"] = Downsample(cout)
            self.down.append(blk)
            cin = cout
        self.mid = nn.ModuleDict({"res1": ResnetBlock(cin, cin, temb), "attn": Attention(cin), "res2": ResnetBlock(cin, cin, temb)})
        rev = list(reversed(channels))
        self.up = nn.ModuleList()
        prev = rev[0]
        for i, cout in enumerate(rev):
            cin_skip = rev[min(i + 1, len(rev) - 1)]
            blk = nn.ModuleDict({"resnets": nn.ModuleList(), "attns": nn.ModuleList()})
            for j in range(layers + 1):
                skip = cin_skip if j == layers else cout
                rin = prev if j == 0 else cout
                blk["resnets"].append(ResnetBlock(rin + skip, cout, temb))
                if (len(channels) - 1 - i) in attn_levels:
                    blk["attns"].append(Attention(cout))
            if i != len(rev) - 1:
                blk["up"] = Upsample(cout)
            self.up.append(blk)
            prev = cout
        self.conv_norm_out = nn.GroupNorm(32, c0, eps=1e-6)
        self.conv_out = nn.Conv2d(c0, in_ch, 3, padding=1)

    def forward(self, x, t):
        temb = self.time_embedding(timestep_embedding(t, self.conv_in.out_channels))
        h = self.conv_in(x)
        skips = [h]
        for blk in self.down:
            for j, res in enumerate(blk["resnets"]):
                h = res(h, temb)
                if len(blk["attns"]):
                    h = blk["attns"][j](h)
                skips.append(h)
            if "down" in blk:
                h = blk["down"](h)
                skips.append(h)
        h = self.mid["res2"](self.mid["attn"](self.mid["res1"](h, temb)), temb)
        for blk in self.up:
            for j, res in enumerate(blk["resnets"]):
                h = res(torch.cat([h, skips.pop()], dim=1), temb)
                if len(blk["attns"]):
                    h = blk["attns"][j](h)
            if "up" in blk:
                h = blk["up"](h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class DDPMScheduler:
    """add_noise of diffusers' DDPMScheduler (linear beta schedule, src/ddpm_config.py:83-100)."""

    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, device="cpu"):
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32, device=device)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)

    def add_noise(self, x, noise, t):
        a = self.alphas_cumprod.to(x.device)[t]
        return a.sqrt()[:, None, None, None] * x + (1 - a).sqrt()[:, None, None, None] * noise

    def inference_timesteps(self, num_inference_steps: int):
        """Evenly strided, descending (diffusers' DDPMScheduler.set_timesteps, "leading" spacing)."""
        n_train = self.alphas_cumprod.shape[0]
        return list(range(0, n_train, n_train // num_inference_steps))[::-1][:num_inference_steps]

    def step(self, eps, t: int, t_prev: int, x, generator=None):
        """One ancestral DDPM step x_t -> x_{t_prev} for an epsilon-prediction model (fixed-small variance)."""
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[t_prev] if t_prev >= 0 else torch.ones_like(a_t)
        beta_t = 1 - a_t / a_prev
        x0 = ((x - (1 - a_t).sqrt() * eps) / a_t.sqrt()).clamp(-1, 1)
        mean = (a_prev.sqrt() * beta_t / (1 - a_t)) * x0 + ((1 - beta_t).sqrt() * (1 - a_prev) / (1 - a_t)) * x
        if t_prev < 0:
            return mean
        var = (1 - a_prev) / (1 - a_t) * beta_t
        return mean + var.clamp(min=1e-20).sqrt() * torch.randn(x.shape, device=x.device, dtype=x.dtype, generator=generator)


def count_parameters(model) -> int:
    """src/attributions/methods/d_trak_grad.py:183-185."""
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
