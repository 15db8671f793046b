"""Oracle: the two denoising loops (turbo fixed-table sampler, base/sft CFG sampler).

Restates generate_audio of
  * /root/reference/acestep/models/turbo/modeling_acestep_v15_turbo.py:1780-2001 (turbo), and
  * /root/reference/acestep/models/base/modeling_acestep_v15_base.py:1783-1989 (base; the sft
    variant adds an explicit `timesteps` override, sft/modeling_acestep_v15_base.py:1866-1873),
starting AFTER prepare_condition: the oracle takes encoder_hidden_states / context_latents as
inputs, exactly like the reference's backend seam `_mlx_run_diffusion`
(acestep/core/generation/handler/diffusion.py:18-33).

`velocity(xt, t_vec, ctx, enc, cache)` is the DiT call (oracle.dit.dit_forward bound to weights).
All schedule arithmetic is done with torch ops in `dtype` so bf16 runs round like the reference.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from .guidance import Momentum, adg, apg

# turbo tables (turbo modeling :1811-1823)
VALID_SHIFTS = [1.0, 2.0, 3.0]
VALID_TIMESTEPS = [
    1.0, 0.9545454545454546, 0.9333333333333333, 0.9, 0.875,
    0.8571428571428571, 0.8333333333333334, 0.7692307692307693, 0.75,
    0.6666666666666666, 0.6428571428571429, 0.625, 0.5454545454545454,
    0.5, 0.4, 0.375, 0.3, 0.25, 0.2222222222222222, 0.125,
]
SHIFT_TIMESTEPS = {
    1.0: [1.0, 0.875, 0.75, 0.625, 0.5, 0.375, 0.25, 0.125],
    2.0: [1.0, 0.9333333333333333, 0.8571428571428571, 0.7692307692307693, 0.6666666666666666,
          0.5454545454545454, 0.4, 0.2222222222222222],
    3.0: [1.0, 0.9545454545454546, 0.9, 0.8333333333333334, 0.75, 0.6428571428571429, 0.5, 0.3],
}


def prepare_noise(shape, seed, dtype=torch.float32, device="cpu") -> torch.Tensor:
    """prepare_noise (turbo :1730-1767): int seed -> one generator; list -> one per sample."""
    b, t, c = shape
    if seed is None:
        return torch.randn(shape, dtype=dtype, device=device)
    if isinstance(seed, list):
        parts = []
        for s in seed:
            if s is None or s < 0:
                parts.append(torch.randn(1, t, c, dtype=dtype, device=device))
            else:
                g = torch.Generator(device=device).manual_seed(int(s))
                parts.append(torch.randn(1, t, c, generator=g, dtype=dtype, device=device))
        return torch.cat(parts, dim=0)
    g = torch.Generator(device=device).manual_seed(int(seed))
    return torch.randn(shape, generator=g, dtype=dtype, device=device)


def turbo_schedule(shift: float = 3.0, timesteps: Optional[Sequence[float]] = None) -> List[float]:
    """Schedule selection (turbo :1807-1865): custom timesteps are stripped of trailing zeros,
    truncated to 20 and snapped to the nearest valid value; otherwise the table of the nearest
    valid shift."""
    sched = None
    if timesteps is not None:
        ts = timesteps.tolist() if isinstance(timesteps, torch.Tensor) else list(timesteps)
        while ts and ts[-1] == 0:
            ts.pop()
        if len(ts) >= 1:
            sched = [min(VALID_TIMESTEPS, key=lambda x: abs(x - t)) for t in ts[:20]]
    if sched is None:
        sched = SHIFT_TIMESTEPS[min(VALID_SHIFTS, key=lambda x: abs(x - shift))]
    return list(sched)


def sample_turbo(velocity: Callable, enc, ctx, src_latents, seed, *, shift=3.0, timesteps=None,
                 infer_method="ode", cover_noise_strength=0.0, audio_cover_strength=1.0,
                 enc_non_cover=None, ctx_non_cover=None, sde_noise: Optional[List[torch.Tensor]] = None,
                 new_cache: Callable = lambda: None, noise=None):
    """Turbo loop (:1917-1991).  `sde_noise[i]` replaces torch.randn_like in renoise so the SDE
    branch is reproducible."""
    sched = turbo_schedule(shift, timesteps)
    dtype = ctx.dtype
    bsz = ctx.shape[0]
    if noise is None:
        noise = prepare_noise((bsz, ctx.shape[1], ctx.shape[-1] // 2), seed, dtype)
    if cover_noise_strength > 0.0:
        level = 1.0 - cover_noise_strength
        nearest = min(sched, key=lambda x: abs(x - level))
        xt = nearest * noise + (1 - nearest) * src_latents
        sched = sched[sched.index(nearest):]
    else:
        xt = noise
    t_sched = torch.tensor(sched, dtype=dtype)  # host scalars (.item()) like the reference
    n = len(sched)
    cover_steps = int(n * audio_cover_strength)
    cache = new_cache()
    switched = False
    for i in range(n):
        t_cur = t_sched[i].item()
        t_vec = t_cur * torch.ones((bsz,), dtype=dtype, device=ctx.device)
        if i >= cover_steps and not switched:
            switched = True
            enc, ctx = enc_non_cover, ctx_non_cover
            cache = new_cache()
        vt = velocity(xt, t_vec, ctx, enc, cache)
        if i == n - 1:
            xt = xt - vt * t_vec[:, None, None]
            break
        t_next = t_sched[i + 1].item()
        if infer_method == "sde":
            x0 = xt - vt * t_vec[:, None, None]
            eps = sde_noise[i] if sde_noise is not None else torch.randn_like(x0)
            xt = t_next * eps + (1 - t_next) * x0
        else:
            dt = t_cur - t_next
            xt = xt - vt * (dt * torch.ones((bsz,), dtype=dtype, device=xt.device))[:, None, None]
    return xt


def base_schedule(infer_steps: int, shift: float, dtype, timesteps=None) -> torch.Tensor:
    """linspace(1, 0, N+1) in the model dtype, then t <- s*t / (1 + (s-1)*t) (base :1864-1867);
    sft accepts an explicit tensor that already ends with 0."""
    if timesteps is not None:
        return torch.as_tensor(timesteps).to(dtype)
    t = torch.linspace(1.0, 0.0, infer_steps + 1, dtype=dtype)
    if shift != 1.0:
        t = shift * t / (1 + (shift - 1) * t)
    return t


def sample_base(velocity: Callable, enc, ctx, src_latents, seed, *, null_emb, infer_steps=30,
                guidance_scale=7.0, shift=1.0, cfg_interval_start=0.0, cfg_interval_end=1.0,
                use_adg=False, infer_method="ode", cover_noise_strength=0.0, audio_cover_strength=1.0,
                enc_non_cover=None, ctx_non_cover=None, timesteps=None,
                sde_noise: Optional[List[torch.Tensor]] = None, new_cache: Callable = lambda: None,
                noise=None):
    """Base/sft loop (:1860-1979): CFG by batch doubling with null_condition_emb, APG (or ADG)
    inside [cfg_interval_start, cfg_interval_end], Euler update in `dtype`."""
    dtype = ctx.dtype
    bsz = ctx.shape[0]
    t = base_schedule(infer_steps, shift, dtype, timesteps)
    infer_steps = len(t) - 1
    cover_steps = int(infer_steps * audio_cover_strength)
    if noise is None:
        noise = prepare_noise((bsz, ctx.shape[1], ctx.shape[-1] // 2), seed, dtype)
    momentum = Momentum()
    if cover_noise_strength > 0.0:
        level = 1.0 - cover_noise_strength
        tv = t[:-1].tolist()
        nearest = min(tv, key=lambda x: abs(x - level))
        start = tv.index(nearest)
        xt = nearest * noise + (1 - nearest) * src_latents
        t = t[start:]
        infer_steps = len(t) - 1
        cover_steps = int(infer_steps * audio_cover_strength)
    else:
        xt = noise
    do_cfg = guidance_scale > 1.0
    if do_cfg:
        enc = torch.cat([enc, null_emb.expand_as(enc)], dim=0)
        ctx = torch.cat([ctx, ctx], dim=0)
    cache = new_cache()
    switched = False
    for i, (t_cur, t_prev) in enumerate(zip(t[:-1], t[1:])):
        if i >= cover_steps and not switched:
            switched = True
            if do_cfg:
                enc_non_cover = torch.cat([enc_non_cover, null_emb.expand_as(enc_non_cover)], dim=0)
                ctx_non_cover = torch.cat([ctx_non_cover, ctx_non_cover], dim=0)
            enc, ctx = enc_non_cover, ctx_non_cover
            cache = new_cache()
        x = torch.cat([xt, xt], dim=0) if do_cfg else xt
        t_vec = t_cur * torch.ones((x.shape[0],), dtype=dtype, device=x.device)
        vt = velocity(x, t_vec, ctx, enc, cache)
        in_interval = bool(t_cur >= cfg_interval_start and t_cur <= cfg_interval_end)
        if do_cfg:
            pc, pn = vt.chunk(2)
            if in_interval:
                vt = adg(xt, pc, pn, t_cur, guidance_scale) if use_adg else \
                    apg(pc, pn, guidance_scale, momentum, dim=1)
            else:
                vt = pc
        if infer_method == "sde":
            tb = t_cur * torch.ones((bsz,), dtype=dtype, device=xt.device)
            x0 = xt - vt * tb[:, None, None]
            nxt = 1.0 - float(i + 1) / infer_steps  # NOTE: ignores `shift` (base :1972)
            eps = sde_noise[i] if sde_noise is not None else torch.randn_like(x0)
            xt = nxt * eps + (1 - nxt) * x0
        else:
            dt = t_cur - t_prev
            xt = xt - vt * (dt * torch.ones((bsz,), dtype=dtype, device=xt.device))[:, None, None]
    return xt
