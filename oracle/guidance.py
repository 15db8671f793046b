"""Oracle: APG / ADG guidance on the conditional / unconditional velocity pair.

Restates /root/reference/acestep/models/base/apg_guidance.py (MomentumBuffer :5-13, project
:16-30, apg_forward :33-56, adg_forward :107-180) as used by the base sampler with dims=[1]
(time axis), momentum -0.75, norm_threshold 2.5, eta 0.
"""
from __future__ import annotations

import torch


class Momentum:
    """running = diff + momentum * running   (apg_guidance.py:5-13)."""

    def __init__(self, momentum: float = -0.75):
        self.momentum = momentum
        self.running = None

    def update(self, diff: torch.Tensor) -> torch.Tensor:
        self.running = diff if self.running is None else diff + self.momentum * self.running
        return self.running


def apg(pred_cond: torch.Tensor, pred_uncond: torch.Tensor, guidance_scale: float,
        momentum: Momentum | None, norm_threshold: float = 2.5, eta: float = 0.0, dim: int = 1):
    """apg_forward (:33-56): momentum-averaged difference, L2-clipped along `dim`, projected
    (in fp64, :26) orthogonally to pred_cond; result = cond + (scale-1) * (orth + eta*parallel)."""
    diff = pred_cond - pred_uncond
    if momentum is not None:
        diff = momentum.update(diff)
    if norm_threshold > 0:
        nrm = diff.norm(p=2, dim=dim, keepdim=True)
        diff = diff * torch.minimum(torch.ones_like(diff), norm_threshold / nrm)
    v0, v1 = diff.double(), pred_cond.double()
    v1 = torch.nn.functional.normalize(v1, dim=dim)
    par = (v0 * v1).sum(dim=dim, keepdim=True) * v1
    orth = v0 - par
    upd = orth.to(diff.dtype) + eta * par.to(diff.dtype)
    return pred_cond + (guidance_scale - 1) * upd


def adg(latents, pred_cond, pred_uncond, sigma, guidance_scale: float, angle_clip: float = 3.14 / 6):
    """adg_forward (:107-180), apply_norm=False, apply_clip=True.  `sigma` is the scalar t_curr."""
    n, t, c = pred_cond.shape
    sigma = torch.as_tensor(sigma, dtype=latents.dtype).to(latents.device).view(1, 1, 1).expand(n, 1, 1)
    weight = guidance_scale - 1
    weight = weight * (weight > 0) + 1e-3
    x_text = latents - sigma * pred_cond
    x_unc = latents - sigma * pred_uncond
    diff = x_text - x_unc

    a = x_text.reshape(-1, c).to(torch.float64)
    b = x_unc.reshape(-1, c).to(torch.float64)
    a = a / torch.linalg.norm(a, dim=1, keepdim=True)
    b = b / torch.linalg.norm(b, dim=1, keepdim=True)
    theta = torch.acos((a * b).sum(dim=1, keepdim=True))  # [n*t, 1] fp64
    theta_new = torch.clip(weight * theta, -angle_clip, angle_clip)

    d = diff.reshape(n * t, c).float()
    u = x_unc.reshape(n * t, c).float()
    proj = ((d * u).sum(1, keepdim=True) / ((u * u).sum(1, keepdim=True) + 1e-8)) * u
    perp = (d - proj).reshape(n, t, c)

    # NOTE the reference broadcasts the [n*t,1] angle tensors against [n,t,c] tensors (:169-172),
    # which is only shape-valid for n == 1 (or t == 1); the oracle reproduces exactly that.
    v_new = torch.cos(theta_new) * x_text
    p_new = perp * torch.sin(theta_new) / torch.sin(theta) * (torch.sin(theta) > 1e-3) \
        + perp * weight * (torch.sin(theta) <= 1e-3)
    x_new = v_new + p_new
    out = (latents - x_new) / sigma
    return out.reshape(n, t, c).to(latents.dtype)
