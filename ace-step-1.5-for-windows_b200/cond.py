"""Host-side condition encoder: text projector + lyric / timbre transformer stacks + sequence packing.

Mirrors AceStepConditionEncoder.forward (acestep/models/turbo/modeling_acestep_v15_turbo.py:1524-1552):
same keyword names, same return pair (encoder_hidden_states [B, Ll + Nt + Lt, D], encoder_attention_mask),
so it can stand in for `model.encoder(...)` inside prepare_condition (:1621-1628).  The transformer
stacks and the projection run on libacestep_b200 (csrc/cond.cu); pack_sequences (:135-166) and
unpack_timbre_embeddings (:1020-1071) are index plumbing and stay in PyTorch on the device.
SURVEY §8f row 1.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .pack import pack_encoder


@dataclass
class CondShape:
    """AceStepConfig fields the condition encoder reads (configuration_acestep_v15.py:148-260)."""

    hidden_size: int = 2048
    intermediate_size: int = 6144
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    sliding_window: int = 128
    rope_theta: float = 1000000.0
    rms_norm_eps: float = 1e-6
    text_hidden_dim: int = 1024
    timbre_hidden_dim: int = 64
    num_lyric_encoder_hidden_layers: int = 8
    num_timbre_encoder_hidden_layers: int = 4

    @classmethod
    def from_config(cls, cfg) -> "CondShape":
        kw = {n: getattr(cfg, n) for n in cls.__dataclass_fields__ if getattr(cfg, n, None) is not None}
        return cls(**kw)


class _EncoderStack:
    """One AceEnc handle (embed_tokens -> n layers -> norm) plus its growable workspace."""

    def __init__(self, lib, sd, prefix: str, n_layers: int, in_dim: int, shape: CondShape, device):
        self.lib, self.device, self.hidden = lib, device, shape.hidden_size
        cfg = _lib.AceEncConfig()
        cfg.hidden_size, cfg.intermediate_size = shape.hidden_size, shape.intermediate_size
        cfg.num_layers, cfg.num_heads, cfg.num_kv_heads = n_layers, shape.num_attention_heads, shape.num_key_value_heads
        cfg.head_dim, cfg.sliding_window, cfg.in_dim = shape.head_dim, int(shape.sliding_window), in_dim
        for i in range(n_layers):  # configuration_acestep_v15.py:251-254: even layers slide, odd are full
            cfg.layer_is_sliding[i] = 1 if (i + 1) % 2 else 0
        cfg.rope_theta, cfg.rms_eps = float(shape.rope_theta), float(shape.rms_norm_eps)
        blob = pack_encoder(sd, n_layers, prefix, embed=in_dim != 0)  # in_dim 0: tokens arrive already embedded
        expect = lib.ace_enc_packed_elems(C.byref(cfg))
        if blob.numel() != expect:
            raise _lib.B200Error(f"packed {prefix} blob has {blob.numel()} elements, library expects {expect}")
        handle = C.c_void_p()
        _lib.check(lib.ace_enc_create(C.byref(handle), C.byref(cfg), blob.data_ptr(), blob.numel()), "ace_enc_create")
        self.handle, self.in_dim, self._ws = handle, in_dim, None

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ace_enc_destroy(self.handle)
            self.handle = None

    def __call__(self, x: torch.Tensor, lengths: Optional[torch.Tensor]) -> torch.Tensor:
        """x [B, S, in_dim] -> [B, S, hidden] (bf16); lengths: int32 [B] valid tokens per sample or None."""
        B, S, _ = x.shape
        x = x.to(self.device, torch.bfloat16).contiguous()
        out = torch.empty(B, S, self.hidden, device=self.device, dtype=torch.bfloat16)
        need = self.lib.ace_enc_workspace_bytes(self.handle, B, S)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        if lengths is not None:
            lengths = lengths.to(self.device, torch.int32).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ace_enc_forward(self.handle, x.data_ptr(), _lib.ptr(lengths), out.data_ptr(), B, S,
                                                self._ws.data_ptr(), self._ws.numel(),
                                                _lib.stream_handle(self.device)), "ace_enc_forward")
        self._keep = (x, lengths)
        return out


def pack_sequences(h1, h2, m1, m2) -> Tuple[torch.Tensor, torch.Tensor]:
    """Valid tokens of [h1 | h2] first, order preserved; new mask = position < valid count (:135-166)."""
    hc, mc = torch.cat([h1, h2], dim=1), torch.cat([m1, m2], dim=1)
    B, L, D = hc.shape
    order = mc.argsort(dim=1, descending=True, stable=True)
    packed = torch.gather(hc, 1, order.unsqueeze(-1).expand(B, L, D))
    lengths = mc.sum(dim=1)
    return packed, torch.arange(L, device=hc.device).unsqueeze(0) < lengths.unsqueeze(1)


def unpack_timbre_embeddings(embs: torch.Tensor, order_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Packed [N, d] + sample id per row -> [B, max_count, d] and its 0/1 mask (:1020-1071); rows keep
    their packed order within each sample."""
    N, d = embs.shape
    B = int(order_mask.max().item()) + 1
    counts = torch.bincount(order_mask, minlength=B)
    mx = int(counts.max().item())
    order = torch.argsort(order_mask, stable=True)
    starts = torch.cumsum(counts, 0) - counts
    pos = torch.arange(N, device=embs.device) - starts[order_mask[order]]
    out = torch.zeros(B, mx, d, dtype=embs.dtype, device=embs.device)
    mask = torch.zeros(B, mx, dtype=torch.long, device=embs.device)
    out[order_mask[order], pos] = embs[order]
    mask[order_mask[order], pos] = 1
    return out, mask


class B200ConditionEncoder:
    """tcgen05 condition encoder.  `state_dict` = `model.encoder.state_dict()` of the reference
    AceStepConditionEncoder (`prefix` e.g. "encoder." for the full model dict)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], shape: Optional[CondShape] = None, device="cuda:0",
                 prefix: str = ""):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.shape = shape or CondShape()
        if self.device.type != "cuda":
            raise _lib.B200Error("B200ConditionEncoder needs a CUDA device (there is no CPU path)")
        sd = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)} if prefix else state_dict
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ace_init(self.device.index or 0), "ace_init")
            s = self.shape
            self.lyric = _EncoderStack(self.lib, sd, "lyric_encoder.", s.num_lyric_encoder_hidden_layers,
                                       s.text_hidden_dim, s, self.device)
            self.timbre = _EncoderStack(self.lib, sd, "timbre_encoder.", s.num_timbre_encoder_hidden_layers,
                                        s.timbre_hidden_dim, s, self.device)
            self.text_w = sd["text_projector.weight"].detach().to(self.device, torch.bfloat16).contiguous()

    def close(self):
        self.lyric.close()
        self.timbre.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _project_text(self, text: torch.Tensor) -> torch.Tensor:
        B, L, K = text.shape
        x = text.to(self.device, torch.bfloat16).contiguous()
        out = torch.empty(B, L, self.shape.hidden_size, device=self.device, dtype=torch.bfloat16)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ace_linear(x.data_ptr(), self.text_w.data_ptr(), 0, out.data_ptr(), B * L,
                                           self.shape.hidden_size, K, _lib.stream_handle(self.device)), "ace_linear")
        return out

    def __call__(self, text_hidden_states=None, text_attention_mask=None, lyric_hidden_states=None,
                 lyric_attention_mask=None, refer_audio_acoustic_hidden_states_packed=None,
                 refer_audio_order_mask=None) -> Tuple[torch.Tensor, torch.Tensor]:
        dev = self.device
        text_mask = text_attention_mask.to(dev)
        lyric_mask = lyric_attention_mask.to(dev)
        # the key-padding mask is passed to the kernels as a valid-token count, which is what a
        # right-padded mask is; anything else (holes, left padding) is not representable
        lengths = lyric_mask.long().sum(dim=1)
        expect = torch.arange(lyric_mask.shape[1], device=dev).unsqueeze(0) < lengths.unsqueeze(1)
        if not torch.equal(lyric_mask.bool(), expect):
            raise ValueError("lyric_attention_mask must be right-padded (valid tokens first)")
        text = self._project_text(text_hidden_states)
        all_valid = bool((lengths == lyric_mask.shape[1]).all())
        lyric = self.lyric(lyric_hidden_states, None if all_valid else lengths)
        timbre = self.timbre(refer_audio_acoustic_hidden_states_packed, None)[:, 0, :]
        t_unpack, t_mask = unpack_timbre_embeddings(timbre, refer_audio_order_mask.to(dev).long())
        h, m = pack_sequences(lyric, t_unpack, lyric_mask.long(), t_mask)
        return pack_sequences(h, text, m.long(), text_mask.long())

    forward = __call__
