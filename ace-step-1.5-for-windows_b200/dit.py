"""Host-side DiT engine: owns an AceDit handle, its workspace and the cross-KV conditioning state.

Mirrors how the reference's MLX backend holds `mlx_decoder` (handler/mlx_dit_init.py:9-43): built
once from the loaded PyTorch decoder's state_dict, then called per denoising step.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from .pack import pack_dit


@dataclass
class DiTShape:
    """Decoder hyper-parameters (AceStepConfig fields, configuration_acestep_v15.py:148-260)."""

    hidden_size: int = 2048
    intermediate_size: int = 6144
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    sliding_window: int = 128
    rope_theta: float = 1000000.0
    rms_norm_eps: float = 1e-6
    layer_types: Optional[List[str]] = None

    def __post_init__(self):
        if self.layer_types is None:
            self.layer_types = ["sliding_attention" if (i + 1) % 2 else "full_attention"
                                for i in range(self.num_hidden_layers)]

    @classmethod
    def from_config(cls, cfg) -> "DiTShape":
        """Build from any object with AceStepConfig-like attributes (reference config, oracle DiTConfig)."""
        names = ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                 "num_key_value_heads", "head_dim", "sliding_window", "rope_theta", "rms_norm_eps")
        kw = {n: getattr(cfg, n) for n in names if getattr(cfg, n, None) is not None}
        lt = getattr(cfg, "layer_types", None)
        return cls(layer_types=list(lt) if lt is not None else None, **kw)


class B200DiT:
    """tcgen05 DiT decoder.  All tensors are bf16 CUDA; results stay on the device."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], shape: DiTShape, device="cuda:0", prefix: str = "",
                 lib=None):
        self.lib = lib or _lib.load()  # `lib`: tests / tools hand in the probe build (_lib.load_probe())
        self.device = torch.device(device)
        self.shape = shape
        if self.device.type != "cuda":
            raise _lib.B200Error("B200DiT needs a CUDA device (there is no CPU path)")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ace_init(self.device.index or 0), "ace_init", self.lib)
            cfg = _lib.AceDitConfig()
            cfg.hidden_size, cfg.intermediate_size = shape.hidden_size, shape.intermediate_size
            cfg.num_layers, cfg.num_heads = shape.num_hidden_layers, shape.num_attention_heads
            cfg.num_kv_heads, cfg.head_dim = shape.num_key_value_heads, shape.head_dim
            cfg.sliding_window = int(shape.sliding_window)
            for i, t in enumerate(shape.layer_types):
                cfg.layer_is_sliding[i] = 1 if t == "sliding_attention" else 0
            cfg.rope_theta, cfg.rms_eps = float(shape.rope_theta), float(shape.rms_norm_eps)
            self._cfg = cfg
            blob = pack_dit(state_dict, shape.num_hidden_layers, prefix)
            expect = self.lib.ace_dit_packed_elems(C.byref(cfg))
            if blob.numel() != expect:
                raise _lib.B200Error(f"packed DiT blob has {blob.numel()} elements, library expects {expect}")
            handle = C.c_void_p()
            _lib.check(self.lib.ace_dit_create(C.byref(handle), C.byref(cfg), blob.data_ptr(), blob.numel()), "ace_dit_create", self.lib)
            self.handle = handle
        self.bound = None  # (bc, T, E)
        self._ws = None

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ace_dit_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def bind(self, bc: int, T: int, E: int) -> None:
        """Size + bind the workspace for effective batch `bc`, `T` latent frames, `E` condition tokens."""
        if self.bound == (bc, T, E):
            return
        with torch.cuda.device(self.device):
            need = self.lib.ace_dit_workspace_bytes(self.handle, bc, T, E)
            if need == 0:
                raise _lib.B200Error(f"invalid DiT shape bc={bc} T={T} E={E}")
            torch.cuda.synchronize(self.device)
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.ace_dit_bind(self.handle, bc, T, E, self._ws.data_ptr(), need), "ace_dit_bind", self.lib)
        self.bound = (bc, T, E)

    def io_views(self):
        """(xt [bc,T,64], ctx [bc,T,128], vt [bc,T,64]) bf16 views of the handle's static I/O slots inside the
        bound workspace (ace_dit_io_slots).  A sampler that keeps its state in them passes exactly these to
        step(), which then skips the three device-to-device copies per step."""
        if self.bound is None:
            raise _lib.B200Error("io_views: handle not bound")
        bc, T, _ = self.bound
        px, pc, pv = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(self.lib.ace_dit_io_slots(self.handle, C.byref(px), C.byref(pc), C.byref(pv)), "ace_dit_io_slots", self.lib)
        base = self._ws.data_ptr()

        def view(ptr, ch):
            off = ptr.value - base
            n = bc * T * ch * 2
            if off < 0 or off + n > self._ws.numel():
                raise _lib.B200Error("io slot outside the bound workspace")
            return self._ws[off: off + n].view(torch.bfloat16).view(bc, T, ch)

        return view(px, 64), view(pc, 128), view(pv, 64)

    def set_condition(self, enc: torch.Tensor) -> None:
        """enc [bc, E, hidden]: condition_embedder + all layers' cross K/V (cached until next call)."""
        bc, E, D = enc.shape
        if self.bound is None or self.bound[0] != bc or self.bound[2] != E:
            raise _lib.B200Error(f"set_condition: enc {tuple(enc.shape)} does not match bound shape {self.bound}")
        enc = enc.to(device=self.device, dtype=torch.bfloat16).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ace_dit_set_condition(self.handle, enc.data_ptr(), _lib.stream_handle(self.device)), "ace_dit_set_condition", self.lib)
        self._enc_keepalive = enc

    def prepare_timesteps(self, ts: Sequence[float]) -> None:
        """Fill the handle's timestep cache for a whole schedule in one batched pass (ace_dit_prepare_timesteps);
        values already cached cost nothing.  Optional: step() computes missing entries on first sight."""
        vals = [float(x) for x in ts]
        if not vals:
            return
        tv = (C.c_float * len(vals))(*vals)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ace_dit_prepare_timesteps(self.handle, tv, len(vals), _lib.stream_handle(self.device)),
                       "ace_dit_prepare_timesteps", self.lib)

    def step(self, xt: torch.Tensor, ctx: torch.Tensor, t: Sequence[float], out: Optional[torch.Tensor] = None):
        """One velocity prediction vt = decoder(xt, t, ctx).  xt [bc,T,64], ctx [bc,T,128] bf16."""
        bc, T, _ = xt.shape
        if self.bound is None or self.bound[:2] != (bc, T):
            raise _lib.B200Error(f"step: xt {tuple(xt.shape)} does not match bound shape {self.bound}")
        xt = xt.to(device=self.device, dtype=torch.bfloat16).contiguous()
        ctx = ctx.to(device=self.device, dtype=torch.bfloat16).contiguous()
        if out is None:
            out = torch.empty_like(xt)
        tv = (C.c_float * bc)(*[float(x) for x in t])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ace_dit_step(self.handle, xt.data_ptr(), ctx.data_ptr(), tv, out.data_ptr(),
                                             _lib.stream_handle(self.device)), "ace_dit_step", self.lib)
        return out

    def cross_attentions(self, xt: torch.Tensor, ctx: torch.Tensor, t: Sequence[float], n_layers: int) -> torch.Tensor:
        """Cross-attention probabilities of layers [0, n_layers): bf16 [n_layers, bc, heads, S, E] with
        S = ceil(T / 2) tokens — element [l] is what the reference decoder returns as
        `outputs[2][l]` with output_attentions=True (turbo :1448-1482); the forward stops after layer
        n_layers - 1.  Call set_condition() first, like step()."""
        bc, T, _ = xt.shape
        if self.bound is None or self.bound[:2] != (bc, T):
            raise _lib.B200Error(f"cross_attentions: xt {tuple(xt.shape)} does not match bound shape {self.bound}")
        if not 1 <= n_layers <= self.shape.num_hidden_layers:
            raise ValueError(f"n_layers {n_layers} out of range [1, {self.shape.num_hidden_layers}]")
        xt = xt.to(device=self.device, dtype=torch.bfloat16).contiguous()
        ctx = ctx.to(device=self.device, dtype=torch.bfloat16).contiguous()
        S, E = (T + 1) // 2, self.bound[2]
        probs = torch.empty(n_layers, bc, self.shape.num_attention_heads, S, E, dtype=torch.bfloat16, device=self.device)
        tv = (C.c_float * bc)(*[float(x) for x in t])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ace_dit_cross_attentions(self.handle, xt.data_ptr(), ctx.data_ptr(), tv, n_layers,
                                                         probs.data_ptr(), _lib.stream_handle(self.device)), "ace_dit_cross_attentions", self.lib)
        return probs
