"""Denoising loops on the B200 DiT engine — host-side mirror of the reference samplers.

`generate_turbo` follows AceStepConditionGenerationModel.generate_audio of the turbo model
(acestep/models/turbo/modeling_acestep_v15_turbo.py:1807-2001) and `generate_base` the base/sft one
(acestep/models/base/modeling_acestep_v15_base.py:1860-1989; sft `timesteps` override
sft/modeling_acestep_v15_base.py:1866-1873), both starting after prepare_condition — the same cut as
the handler's backend seam `_mlx_run_diffusion` (handler/diffusion.py:18-33).  Schedules are built
with the same torch ops in the model dtype so timesteps round exactly like the reference's; the
per-step arithmetic (DiT, APG/ADG, Euler/SDE update) runs in CUDA kernels through the C ABI.
"""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional, Sequence, Union

import torch

from . import _lib
from .dit import B200DiT

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

Seed = Union[int, List[Optional[int]], None]


def prepare_noise(shape, seed: Seed, device, dtype=torch.bfloat16) -> torch.Tensor:
    """Same RNG calls as the reference's prepare_noise (turbo modeling :1730-1767) so a given seed
    yields the same starting noise."""
    b, t, c = shape
    if seed is None:
        return torch.randn(shape, device=device, dtype=dtype)
    if isinstance(seed, list):
        parts = []
        for s in seed:
            if s is None or s < 0:
                parts.append(torch.randn(1, t, c, device=device, dtype=dtype))
            else:
                g = torch.Generator(device=device).manual_seed(int(s))
                parts.append(torch.randn(1, t, c, generator=g, device=device, dtype=dtype))
        return torch.cat(parts, dim=0)
    g = torch.Generator(device=device).manual_seed(int(seed))
    return torch.randn(shape, generator=g, device=device, dtype=dtype)


def turbo_schedule(shift: float, timesteps) -> List[float]:
    """turbo :1826-1865 — custom timesteps snapped to the 20 valid values, else the shift table."""
    sched = None
    if timesteps is not None:
        ts = timesteps.tolist() if hasattr(timesteps, "tolist") else list(timesteps)
        while ts and ts[-1] == 0:
            ts.pop()
        if len(ts) >= 1:
            sched = [min(VALID_TIMESTEPS, key=lambda x: abs(x - t)) for t in ts[:20]]
    if sched is None:
        sched = SHIFT_TIMESTEPS[min(VALID_SHIFTS, key=lambda x: abs(x - shift))]
    return list(sched)


def _bf16(x: float) -> float:
    return float(torch.tensor(x, dtype=torch.bfloat16))


class B200Sampler:
    def __init__(self, dit: B200DiT, null_condition_emb: Optional[torch.Tensor] = None):
        self.dit = dit
        self.lib = dit.lib
        self.device = dit.device
        self.null_condition_emb = None
        if null_condition_emb is not None:
            self.null_condition_emb = null_condition_emb.detach().to(self.device, torch.bfloat16)
        self._sched: Dict[Any, torch.Tensor] = {}  # (infer_steps, shift) -> bf16 schedule, host copy

    def _base_schedule(self, infer_steps: int, shift: float, timesteps) -> torch.Tensor:
        """The base / sft timestep schedule as a HOST bf16 tensor.  linspace and the shift formula run on the device
        in the model dtype exactly like the reference (base :1895-1899) — once per (steps, shift); the host copy is
        kept, so a steady-state request reads no device value back (each `.tolist()` on a device tensor is a host
        synchronisation that drains the queue between two songs).  Explicit `timesteps` (sft :1866-1873) are a pure
        dtype conversion, which rounds the same on either side."""
        bf = torch.bfloat16
        if timesteps is not None:
            return torch.as_tensor(timesteps).detach().to("cpu").to(bf)
        key = (int(infer_steps), float(shift))
        t = self._sched.get(key)
        if t is None:
            t = torch.linspace(1.0, 0.0, infer_steps + 1, device=self.device, dtype=bf)
            if shift != 1.0:
                t = shift * t / (1 + (shift - 1) * t)
            t = t.cpu()
            self._sched[key] = t
        return t

    # ------------------------------------------------------------------
    def _stream(self):
        return _lib.stream_handle(self.device)

    def _euler(self, xt, vt, dt: float, dup=None):
        if dup is None:
            _lib.check(self.lib.ace_euler_step(xt.data_ptr(), vt.data_ptr(), dt, xt.numel(), self._stream()),
                       "ace_euler_step")
        else:  # the new state also lands in `dup` (unconditional half of the next step's CFG batch)
            _lib.check(self.lib.ace_euler_step_dup(xt.data_ptr(), vt.data_ptr(), dt, xt.numel(), dup.data_ptr(),
                                                   self._stream()), "ace_euler_step_dup")

    def _sde(self, xt, vt, eps, t_cur: float, t_next: float):
        _lib.check(self.lib.ace_sde_step(xt.data_ptr(), vt.data_ptr(), eps.data_ptr(), t_cur, t_next, xt.numel(),
                                         self._stream()), "ace_sde_step")

    def _validate(self, enc, ctx, src, infer_method, timesteps, enc_nc, ctx_nc):
        # same checks (and exception types) as DiffusionMixin._mlx_run_diffusion, diffusion.py:70-95
        if infer_method not in {"ode", "sde"}:
            raise ValueError(f"Unsupported infer_method '{infer_method}'. Expected 'ode' or 'sde'.")
        if timesteps is not None and not (hasattr(timesteps, "__iter__") or hasattr(timesteps, "tolist")):
            raise TypeError("timesteps must be iterable, tensor-like, or None")
        if enc.shape[0] != ctx.shape[0]:
            raise ValueError("Batch dimension mismatch: encoder_hidden_states and context_latents must share dim 0")
        if enc.shape[0] != src.shape[0]:
            raise ValueError("Batch dimension mismatch: encoder_hidden_states and src_latents must share dim 0")
        if enc_nc is not None and enc_nc.shape[0] != enc.shape[0]:
            raise ValueError("Batch dimension mismatch: encoder_hidden_states_non_cover must share dim 0 with encoder_hidden_states")
        if ctx_nc is not None and ctx_nc.shape[0] != ctx.shape[0]:
            raise ValueError("Batch dimension mismatch: context_latents_non_cover must share dim 0 with context_latents")

    def _to(self, x):
        return None if x is None else x.detach().to(self.device, torch.bfloat16).contiguous()

    # ------------------------------------------------------------------
    def generate_turbo(self, encoder_hidden_states, context_latents, src_latents, seed: Seed = None, *,
                       infer_method: str = "ode", shift: float = 3.0, timesteps=None,
                       audio_cover_strength: float = 1.0, cover_noise_strength: float = 0.0,
                       encoder_hidden_states_non_cover=None, context_latents_non_cover=None,
                       noise: Optional[torch.Tensor] = None, sde_noise: Optional[Sequence[torch.Tensor]] = None,
                       sync: bool = True) -> Dict[str, Any]:
        """`sync=False`: return as soon as the loop is enqueued (no host synchronisation anywhere in the call; the
        time costs are then enqueue times) — for callers that queue the decode right behind it."""
        self._validate(encoder_hidden_states, context_latents, src_latents, infer_method, timesteps,
                       encoder_hidden_states_non_cover, context_latents_non_cover)
        t0 = time.time()
        enc, ctx, src = self._to(encoder_hidden_states), self._to(context_latents), self._to(src_latents)
        enc_nc, ctx_nc = self._to(encoder_hidden_states_non_cover), self._to(context_latents_non_cover)
        B, T, _ = ctx.shape
        sched = turbo_schedule(shift, timesteps)
        if noise is None:
            noise = prepare_noise((B, T, ctx.shape[-1] // 2), seed, self.device)
        noise = self._to(noise)
        if cover_noise_strength > 0.0:
            level = 1.0 - cover_noise_strength
            nearest = min(sched, key=lambda x: abs(x - level))
            xt = (nearest * noise + (1 - nearest) * src).contiguous()  # renoise(), bf16 torch ops
            sched = sched[sched.index(nearest):]
        else:
            xt = noise.clone()
        t_sched = torch.tensor(sched, dtype=torch.bfloat16).tolist()  # the model-dtype rounding of the table (host)
        n = len(t_sched)
        cover_steps = int(n * audio_cover_strength)
        self.dit.bind(B, T, enc.shape[1])
        self.dit.set_condition(enc)
        self.dit.prepare_timesteps(t_sched)  # timestep cache: a no-op once this schedule has been seen
        # loop state in the handle's static I/O slots: no per-step copies (see generate_base)
        xin, ctxin, vt = self.dit.io_views()
        ctxin.copy_(ctx)
        xin.copy_(xt)
        xt = xin
        switched = False
        for i in range(n):
            t_cur = t_sched[i]
            if i >= cover_steps and not switched:
                switched = True
                if enc_nc is None or ctx_nc is None:
                    raise ValueError("audio_cover_strength < 1 needs non-cover conditioning")
                enc, ctx = enc_nc, ctx_nc
                keep = xt.clone()  # a different condition length re-carves the workspace
                self.dit.bind(B, T, enc.shape[1])
                self.dit.set_condition(enc)  # re-creates the cross-KV cache (turbo :1956)
                xin, ctxin, vt = self.dit.io_views()
                ctxin.copy_(ctx)
                xin.copy_(keep)
                xt = xin
            self.dit.step(xt, ctxin, [t_cur] * B, out=vt)
            if i == n - 1:
                self._euler(xt, vt, t_cur)  # x0 = xt - vt * t (:1975-1977)
                break
            t_next = t_sched[i + 1]
            if infer_method == "sde":
                eps = self._to(sde_noise[i]) if sde_noise is not None else torch.randn_like(xt)
                self._sde(xt, vt, eps, t_cur, t_next)
            else:
                self._euler(xt, vt, _bf16(t_cur - t_next))
        xt = xt.clone()  # the slots are reused by the next request
        if sync:
            torch.cuda.synchronize(self.device)
        t1 = time.time()
        return {"target_latents": xt, "steps": n,
                "time_costs": {"diffusion_time_cost": t1 - t0, "diffusion_per_step_time_cost": (t1 - t0) / max(n, 1),
                               "total_time_cost": t1 - t0}}

    # ------------------------------------------------------------------
    def generate_base(self, encoder_hidden_states, context_latents, src_latents, seed: Seed = None, *,
                      infer_method: str = "ode", infer_steps: int = 30, diffusion_guidance_sale: float = 7.0,
                      shift: float = 1.0, timesteps=None, cfg_interval_start: float = 0.0,
                      cfg_interval_end: float = 1.0, use_adg: bool = False, audio_cover_strength: float = 1.0,
                      cover_noise_strength: float = 0.0, encoder_hidden_states_non_cover=None,
                      context_latents_non_cover=None, null_condition_emb: Optional[torch.Tensor] = None,
                      noise: Optional[torch.Tensor] = None, sde_noise: Optional[Sequence[torch.Tensor]] = None,
                      sync: bool = True) -> Dict[str, Any]:
        """`sync=False`: see generate_turbo."""
        self._validate(encoder_hidden_states, context_latents, src_latents, infer_method, timesteps,
                       encoder_hidden_states_non_cover, context_latents_non_cover)
        t0 = time.time()
        dev, bf = self.device, torch.bfloat16
        enc, ctx, src = self._to(encoder_hidden_states), self._to(context_latents), self._to(src_latents)
        enc_nc, ctx_nc = self._to(encoder_hidden_states_non_cover), self._to(context_latents_non_cover)
        B, T, _ = ctx.shape
        t = self._base_schedule(infer_steps, shift, timesteps)  # host bf16; sft: explicit schedule incl. the trailing 0
        if noise is None:
            noise = prepare_noise((B, T, ctx.shape[-1] // 2), seed, dev)
        noise = self._to(noise)
        if cover_noise_strength > 0.0:
            level = 1.0 - cover_noise_strength
            tv = t[:-1].tolist()
            nearest = min(tv, key=lambda x: abs(x - level))
            xt = (nearest * noise + (1 - nearest) * src).contiguous()
            t = t[tv.index(nearest):]
        else:
            xt = noise.clone()
        n = len(t) - 1
        cover_steps = int(n * audio_cover_strength)
        in_interval = ((t[:-1] >= cfg_interval_start) & (t[:-1] <= cfg_interval_end)).tolist()
        dts = (t[:-1] - t[1:]).tolist()  # bf16 subtraction, like `dt = t_curr - t_prev` (:1977)
        ts = t.tolist()

        do_cfg = diffusion_guidance_sale > 1.0
        null = null_condition_emb if null_condition_emb is not None else self.null_condition_emb
        if do_cfg:
            if null is None:
                raise ValueError("guidance > 1 needs null_condition_emb")
            null = null.to(dev, bf)
            enc = torch.cat([enc, null.expand_as(enc)], dim=0)
            ctx = torch.cat([ctx, ctx], dim=0)
        Bc = enc.shape[0]
        self.dit.bind(Bc, T, enc.shape[1])
        self.dit.set_condition(enc)
        self.dit.prepare_timesteps(ts[:-1])  # timestep cache: a no-op once this schedule has been seen
        # The loop state lives in the DiT handle's static I/O slots (xin = [xt | xt copy for the unconditional
        # half], ctx, vt), so a step is: set t -> CUDA graph -> guidance -> Euler, with no copies in between.
        xin, ctxin, vt = self.dit.io_views()
        ctxin.copy_(ctx)
        xin[:B].copy_(xt)
        if do_cfg:
            xin[B:].copy_(xt)
        xt = xin[:B]
        dup = xin[B:] if do_cfg else None
        vg = torch.empty_like(xt)
        momentum = torch.zeros_like(xt)
        first_apg = True
        switched = False
        for i in range(n):
            t_cur = ts[i]
            if i >= cover_steps and not switched:
                switched = True
                if enc_nc is None or ctx_nc is None:
                    raise ValueError("audio_cover_strength < 1 needs non-cover conditioning")
                enc, ctx = enc_nc, ctx_nc
                if do_cfg:
                    enc = torch.cat([enc, null.expand_as(enc)], dim=0)
                    ctx = torch.cat([ctx, ctx], dim=0)
                keep = xt.clone()  # a different condition length re-carves the workspace
                self.dit.bind(Bc, T, enc.shape[1])
                self.dit.set_condition(enc)  # new cross-KV cache (base :1927)
                xin, ctxin, vt = self.dit.io_views()
                ctxin.copy_(ctx)
                xin[:B].copy_(keep)
                if do_cfg:
                    xin[B:].copy_(keep)
                xt = xin[:B]
                dup = xin[B:] if do_cfg else None
            self.dit.step(xin, ctxin, [t_cur] * Bc, out=vt)
            v = vt[:B]
            if do_cfg and in_interval[i]:
                if use_adg:
                    _lib.check(self.lib.ace_adg(xt.data_ptr(), vt[:B].data_ptr(), vt[B:].data_ptr(), t_cur,
                                                float(diffusion_guidance_sale), 3.14 / 6, vg.data_ptr(), B, T,
                                                self._stream()), "ace_adg")
                else:
                    _lib.check(self.lib.ace_apg(vt[:B].data_ptr(), vt[B:].data_ptr(), momentum.data_ptr(),
                                                1 if first_apg else 0, -0.75, 2.5, float(diffusion_guidance_sale),
                                                vg.data_ptr(), B, T, self._stream()), "ace_apg")
                    first_apg = False
                v = vg
            if infer_method == "sde":
                nxt = 1.0 - float(i + 1) / n  # ignores `shift`, like the reference (:1972)
                eps = self._to(sde_noise[i]) if sde_noise is not None else torch.randn_like(xt)
                self._sde(xt, v, eps, t_cur, nxt)
                if do_cfg:
                    dup.copy_(xt)
            else:
                self._euler(xt, v, dts[i], dup)
        xt = xt.clone()  # the slots are reused by the next request
        if sync:
            torch.cuda.synchronize(dev)
        t1 = time.time()
        return {"target_latents": xt, "steps": n,
                "time_costs": {"diffusion_time_cost": t1 - t0, "diffusion_per_step_time_cost": (t1 - t0) / max(n, 1),
                               "total_time_cost": t1 - t0}}
