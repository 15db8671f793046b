#!/usr/bin/env python
"""Generate tests/golden/cond_encoder.npz by running the REAL reference AceStepConditionEncoder
(/root/reference/acestep/models/turbo/modeling_acestep_v15_turbo.py:1506-1552, imported unmodified)
on a tiny config with oracle.cond.make_cond_weights tensors and padded synthetic inputs.

Runs only in the build container (the GPU box has no /root/reference); the fixture is committed.
`vector_quantize_pytorch` (absent) is stubbed; the condition encoder never touches it.

Usage: python tools/make_golden_cond.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

stub = types.ModuleType("vector_quantize_pytorch")


class _ResidualFSQ(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


stub.ResidualFSQ = _ResidualFSQ
sys.modules["vector_quantize_pytorch"] = stub

from acestep.models.turbo import modeling_acestep_v15_turbo as T  # noqa: E402
from acestep.models.turbo.configuration_acestep_v15 import AceStepConfig  # noqa: E402

from oracle.cond import CondConfig, make_cond_weights  # noqa: E402

torch.set_grad_enabled(False)


def synth_inputs(cfg: CondConfig, seed: int = 21):
    """B = 2 with right-padded lyrics / text and three packed timbre references (one for sample 0,
    two for sample 1)."""
    g = torch.Generator().manual_seed(seed)
    B, Ll, Lt, Lr = 2, 24, 7, 10  # sample 1: 11 valid lyric tokens -> rows >= 19 have an all-padding band on sliding layers
    lyric = torch.randn(B, Ll, cfg.text_hidden_dim, generator=g)
    lyric_mask = torch.ones(B, Ll, dtype=torch.long)
    lyric_mask[1, 11:] = 0
    text = torch.randn(B, Lt, cfg.text_hidden_dim, generator=g)
    text_mask = torch.ones(B, Lt, dtype=torch.long)
    text_mask[1, 5:] = 0
    refer = torch.randn(3, Lr, cfg.timbre_hidden_dim, generator=g)
    order = torch.tensor([0, 1, 1], dtype=torch.long)
    return dict(text=text, text_mask=text_mask, lyric=lyric, lyric_mask=lyric_mask, refer=refer, order=order)


def main():
    cfg = CondConfig.tiny()
    w = make_cond_weights(cfg, seed=5)
    rc = AceStepConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, num_hidden_layers=4,
                       num_attention_heads=cfg.num_attention_heads, num_key_value_heads=cfg.num_key_value_heads,
                       head_dim=cfg.head_dim, sliding_window=cfg.sliding_window, rope_theta=cfg.rope_theta,
                       rms_norm_eps=cfg.rms_norm_eps, text_hidden_dim=cfg.text_hidden_dim,
                       timbre_hidden_dim=cfg.timbre_hidden_dim,
                       num_lyric_encoder_hidden_layers=cfg.num_lyric_encoder_hidden_layers,
                       num_timbre_encoder_hidden_layers=cfg.num_timbre_encoder_hidden_layers)
    rc._attn_implementation = "sdpa"
    m = T.AceStepConditionEncoder(rc).float().eval()
    missing, unexpected = m.load_state_dict(w, strict=False)
    assert not unexpected, unexpected
    assert all("rotary_emb" in k for k in missing), missing
    x = synth_inputs(cfg)
    hs, mask = m(text_hidden_states=x["text"], text_attention_mask=x["text_mask"], lyric_hidden_states=x["lyric"],
                 lyric_attention_mask=x["lyric_mask"], refer_audio_acoustic_hidden_states_packed=x["refer"],
                 refer_audio_order_mask=x["order"])
    out = os.path.join(ROOT, "tests", "golden", "cond_encoder.npz")
    np.savez_compressed(out, out_hidden=hs.numpy(), out_mask=mask.numpy().astype(np.int64),
                        **{k: v.numpy() for k, v in x.items()})
    print("wrote", out, tuple(hs.shape), mask.long().sum(1).tolist())


if __name__ == "__main__":
    main()
