#!/usr/bin/env python
"""Golden cross-attention probabilities from the REAL reference decoder (build container only):
`AceStepDiTModel(..., output_attentions=True, custom_layers_config=cfg, enable_early_exit=True)[2]`, the call the
lyric aligner makes (handler/lyric_timestamp.py:78-91), tiny config, fp32.  -> tests/golden/dit_cross_attn_tiny.npz"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (stubs vector_quantize_pytorch, puts the reference on sys.path)

from oracle.dit import DiTConfig  # noqa: E402
from oracle.weights import make_dit_weights  # noqa: E402


def main():
    cfg = DiTConfig.tiny()
    w = make_dit_weights(cfg, seed=0)
    dec = mg.ref_decoder(cfg, w, mg.T)
    enc, src, ctx, xt = mg.synth(cfg, 2, 37, 21, seed=31)
    t = torch.tensor([0.125, 0.125])
    layers = {0: [0], cfg.num_hidden_layers - 1: [1]}
    out = dec(hidden_states=xt, timestep=t, timestep_r=t, attention_mask=torch.ones(2, 37),
              encoder_hidden_states=enc, use_cache=False, past_key_values=None,
              encoder_attention_mask=torch.ones(2, 21), context_latents=ctx, output_attentions=True,
              custom_layers_config=layers, enable_early_exit=True)
    attn = out[2]
    assert len(attn) == cfg.num_hidden_layers and all(a is not None for a in attn)
    mg.save("dit_cross_attn_tiny", xt=xt, t=t, ctx=ctx, enc=enc, probs=torch.stack(list(attn)), vt=out[0])


if __name__ == "__main__":
    main()
