"""Synthetic radar frames for measurement and tests.  Shape-and-statistics stand-ins for the
reference's offline generators; values do not affect throughput.

  rayleigh_target_frames : Rayleigh(sigma=1) clutter + Gaussian extended targets at a peak SNR (the recipe of
                           Rayleigh_bg_Gaussian_EOT_generator_20230208.py:219-249, without its 400x400 crop logic)
  k_clutter_frames       : compound-Gaussian K-distributed clutter, Rayleigh speckle x sqrt(Gamma(nu, 1/nu)) texture
                           (the distribution K_distributed_SeaClutter_Simulation_20210919.py:469-526 synthesises
                           with spatial correlation at ~9 s per frame; here uncorrelated, nu = 5)
Every frame is min-max normalised per (image, channel) like utils_20231218.py:673-689.
"""
import numpy as np
import torch


def normalize_per_frame(x: torch.Tensor) -> torch.Tensor:
    nb, nc, h, w = x.shape
    v = x.reshape(nb, nc, h * w)
    mn = v.min(dim=-1, keepdim=True)[0]
    mx = v.max(dim=-1, keepdim=True)[0]
    return ((v - mn) / (mx - mn + np.spacing(1))).reshape(nb, nc, h, w)


def _targets(x, gen, n_targets, amp):
    nb, nc, h, w = x.shape
    yy = torch.arange(h, dtype=torch.float32).view(1, h, 1)
    xx = torch.arange(w, dtype=torch.float32).view(1, 1, w)
    for b in range(nb):
        cy = torch.rand(n_targets, generator=gen).view(-1, 1, 1) * h
        cx = torch.rand(n_targets, generator=gen).view(-1, 1, 1) * w
        sy = 1.5 + 3.5 * torch.rand(n_targets, generator=gen).view(-1, 1, 1)
        sx = 1.5 + 3.5 * torch.rand(n_targets, generator=gen).view(-1, 1, 1)
        blob = torch.exp(-((yy - cy) ** 2 / (2 * sy ** 2) + (xx - cx) ** 2 / (2 * sx ** 2))).sum(0)
        x[b] += amp * blob
    return x


def rayleigh_target_frames(batch, chans=1, h=256, w=256, seed=1981, n_targets=20, snr_db=2.0):
    gen = torch.Generator().manual_seed(seed)
    u = torch.rand(batch, chans, h, w, generator=gen).clamp_min(1e-12)
    x = torch.sqrt(-2.0 * torch.log(u))                                  # Rayleigh(sigma = 1)
    x = _targets(x, gen, n_targets, 3.0 * 10 ** (snr_db / 20.0))
    return normalize_per_frame(x)


def k_clutter_frames(batch, chans=1, h=256, w=256, seed=1981, nu=5.0, n_targets=20, snr_db=2.0):
    gen = torch.Generator().manual_seed(seed)
    u = torch.rand(batch, chans, h, w, generator=gen).clamp_min(1e-12)
    speckle = torch.sqrt(-2.0 * torch.log(u))
    # Gamma(nu, 1/nu) texture as a sum of nu exponentials (integer nu)
    e = -torch.log(torch.rand(int(nu), batch, chans, h, w, generator=gen).clamp_min(1e-12))
    texture = e.sum(0) / nu
    x = speckle * torch.sqrt(texture)
    if n_targets:
        x = _targets(x, gen, n_targets, 3.0 * 10 ** (snr_db / 20.0))
    return normalize_per_frame(x)
