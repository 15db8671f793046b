"""Host-side audio tokenizer / residual FSQ / detokenizer: the LM-hint branch of `prepare_condition`.

Mirrors AceStepConditionGenerationModel.tokenize / detokenize (acestep/models/turbo/modeling_acestep_v15_turbo.py
:1577-1600) and the modules behind them — AceStepAudioTokenizer (:1178-1218: audio_acoustic_proj -> AttentionPooler
:730-856 -> ResidualFSQ) and AudioTokenDetokenizer (:859-990) — with the same method names, argument order and
return tuples, so `_b200_prepare_condition` can run the cover-song branch (:1630-1646) without leaving the backend.

Device work: the Linear layers (`ace_linear`), the two 2-layer encoder stacks over sequences of pool_window_size (+1)
tokens (`ace_enc_forward` with in_dim = 0: csrc/cond.cu) and the quantizer (`ace_fsq`).  Joining the special
token(s), the silence padding and the 5 Hz mask pooling are index plumbing and stay in PyTorch on the device.
SURVEY §8f row 1.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import _lib
from .cond import CondShape, _EncoderStack


@dataclass
class TokShape(CondShape):
    """AceStepConfig fields of the tokenizer path (configuration_acestep_v15.py:148-260)."""

    audio_acoustic_hidden_dim: int = 64
    pool_window_size: int = 5
    fsq_dim: int = 2048
    fsq_input_levels: List[int] = field(default_factory=lambda: [8, 8, 8, 5, 5, 5])
    fsq_input_num_quantizers: int = 1
    num_attention_pooler_hidden_layers: int = 2


class B200AudioTokenizer:
    """`state_dict`: the reference MODEL's state dict restricted to `tokenizer.*` and `detokenizer.*` keys (or the
    whole `model.state_dict()`; other keys are ignored)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], shape: Optional[TokShape] = None, device="cuda:0"):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.shape = s = shape or TokShape()
        if self.device.type != "cuda":
            raise _lib.B200Error("B200AudioTokenizer needs a CUDA device (there is no CPU path)")
        if s.pool_window_size + 1 > s.sliding_window + 1:
            raise _lib.B200Error("pool_window_size wider than the sliding window is not supported")
        if s.fsq_dim != s.hidden_size:
            raise _lib.B200Error(f"fsq_dim {s.fsq_dim} != hidden_size {s.hidden_size}")
        sd = state_dict
        dev = lambda k: sd[k].detach().to(self.device, torch.bfloat16).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ace_init(self.device.index or 0), "ace_init")
            n = s.num_attention_pooler_hidden_layers
            self.pooler = _EncoderStack(self.lib, sd, "tokenizer.attention_pooler.", n, 0, s, self.device)
            self.detok = _EncoderStack(self.lib, sd, "detokenizer.", n, 0, s, self.device)
        self.w = {k: dev(k) for k in (
            "tokenizer.audio_acoustic_proj.weight", "tokenizer.audio_acoustic_proj.bias",
            "tokenizer.attention_pooler.embed_tokens.weight", "tokenizer.attention_pooler.embed_tokens.bias",
            "tokenizer.attention_pooler.special_token",
            "tokenizer.quantizer.project_in.weight", "tokenizer.quantizer.project_in.bias",
            "tokenizer.quantizer.project_out.weight", "tokenizer.quantizer.project_out.bias",
            "detokenizer.embed_tokens.weight", "detokenizer.embed_tokens.bias", "detokenizer.special_tokens",
            "detokenizer.proj_out.weight", "detokenizer.proj_out.bias")}
        self._levels = (C.c_int * len(s.fsq_input_levels))(*[int(x) for x in s.fsq_input_levels])

    def close(self):
        self.pooler.close()
        self.detok.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def _linear(self, x: torch.Tensor, wkey: str, bkey: Optional[str], k_pad: int = 0) -> torch.Tensor:
        """x [..., K] -> [..., N] through ace_linear (K must be a multiple of 64)."""
        w = self.w[wkey]
        lead, K = x.shape[:-1], x.shape[-1]
        x2 = x.reshape(-1, K).to(self.device, torch.bfloat16).contiguous()
        out = torch.empty(x2.shape[0], w.shape[0], device=self.device, dtype=torch.bfloat16)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ace_linear(x2.data_ptr(), w.data_ptr(), _lib.ptr(self.w[bkey]) if bkey else 0,
                                           out.data_ptr(), x2.shape[0], w.shape[0], K,
                                           _lib.stream_handle(self.device)), "ace_linear")
        return out.reshape(*lead, w.shape[0])

    def quantize(self, h: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """ResidualFSQ.forward: h [B, Tp, D] -> (quantized [B, Tp, D] bf16, indices [B, Tp, num_quantizers] int32)."""
        B, Tp, D = h.shape
        s = self.shape
        x = h.reshape(B * Tp, D).to(self.device, torch.bfloat16).contiguous()
        q = torch.empty_like(x)
        idx = torch.empty(B * Tp, s.fsq_input_num_quantizers, device=self.device, dtype=torch.int32)
        w = self.w
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ace_fsq(x.data_ptr(), w["tokenizer.quantizer.project_in.weight"].data_ptr(),
                                        w["tokenizer.quantizer.project_in.bias"].data_ptr(), self._levels,
                                        len(s.fsq_input_levels), s.fsq_input_num_quantizers,
                                        w["tokenizer.quantizer.project_out.weight"].data_ptr(),
                                        w["tokenizer.quantizer.project_out.bias"].data_ptr(), q.data_ptr(),
                                        idx.data_ptr(), B * Tp, D, _lib.stream_handle(self.device)), "ace_fsq")
        return q.reshape(B, Tp, D), idx.reshape(B, Tp, -1)

    def pool(self, x: torch.Tensor) -> torch.Tensor:
        """audio_acoustic_proj + AttentionPooler (:1206-1209, :755-856): x [B, Tp, P, 64] -> [B, Tp, D]."""
        B, Tp, P, _ = x.shape
        D = self.shape.hidden_size
        h = self._linear(x, "tokenizer.audio_acoustic_proj.weight", "tokenizer.audio_acoustic_proj.bias")
        h = self._linear(h, "tokenizer.attention_pooler.embed_tokens.weight",
                         "tokenizer.attention_pooler.embed_tokens.bias")
        special = self.w["tokenizer.attention_pooler.special_token"].expand(B, Tp, 1, D)
        seq = torch.cat([special, h], dim=2).reshape(B * Tp, P + 1, D)
        return self.pooler(seq, None)[:, 0, :].reshape(B, Tp, D)

    def tokenizer_forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """AceStepAudioTokenizer.forward (:1200-1213): x [B, Tp, P, 64] -> (quantized, indices)."""
        return self.quantize(self.pool(x))

    def tokenize(self, x: torch.Tensor, silence_latent: torch.Tensor, attention_mask: torch.Tensor):
        """model.tokenize (:1577-1588): x [B, T, 64] -> (quantized [B, T/P, D], indices, attention mask at 5 Hz)."""
        P = self.shape.pool_window_size
        x = x.to(self.device, torch.bfloat16)
        attention_mask = attention_mask.to(self.device)
        if x.shape[1] % P != 0:
            pad = P - x.shape[1] % P
            sil = silence_latent.to(self.device, torch.bfloat16)
            x = torch.cat([x, sil[:1, :pad].repeat(x.shape[0], 1, 1)], dim=1)
            attention_mask = F.pad(attention_mask, (0, pad), mode="constant", value=0)
        x = x.reshape(x.shape[0], x.shape[1] // P, P, x.shape[2])
        seq_len = x.shape[1]
        chunk = math.ceil(attention_mask.shape[1] / seq_len)
        m = F.max_pool1d(attention_mask.to(x.dtype).unsqueeze(1), kernel_size=chunk, stride=chunk,
                         ceil_mode=True).squeeze(1)
        q, idx = self.tokenizer_forward(x)
        return q, idx, m

    def detokenize(self, quantized: torch.Tensor) -> torch.Tensor:
        """model.detokenize / AudioTokenDetokenizer.forward (:887-990): [B, Tp, D] -> [B, Tp * P, 64] bf16."""
        B, Tp, D = quantized.shape
        P = self.shape.pool_window_size
        h = self._linear(quantized, "detokenizer.embed_tokens.weight", "detokenizer.embed_tokens.bias")
        h = h.unsqueeze(2).repeat(1, 1, P, 1) + self.w["detokenizer.special_tokens"].expand(B, Tp, -1, -1)
        h = self.detok(h.reshape(B * Tp, P, D), None)
        out = self._linear(h, "detokenizer.proj_out.weight", "detokenizer.proj_out.bias")
        return out.reshape(B, Tp * P, -1)

    def lm_hints(self, hidden_states, silence_latent, attention_mask, n_frames: int) -> torch.Tensor:
        """tokenize -> detokenize -> crop to n_frames: `lm_hints_25Hz` of prepare_condition (:1640-1645)."""
        q, _idx, _m = self.tokenize(hidden_states, silence_latent, attention_mask)
        return self.detokenize(q)[:, :n_frames, :]
