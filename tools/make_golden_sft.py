#!/usr/bin/env python
"""Golden for the SFT sampler variant from the REAL reference (build container only):
`acestep.models.sft.modeling_acestep_v15_base.AceStepConditionGenerationModel.generate_audio` with an explicit
`timesteps` tensor (sft :1864-1875: overrides infer_steps and shift), CFG + APG, and the same with a cover-noise
start.  -> tests/golden/sft_timesteps.npz, sft_timesteps_cover.npz"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402

from acestep.models.sft import modeling_acestep_v15_base as S  # noqa: E402

from oracle.dit import DiTConfig  # noqa: E402
from oracle.weights import make_dit_weights, make_null_condition_emb  # noqa: E402


def main():
    cfg = DiTConfig.tiny()
    w = make_dit_weights(cfg, seed=0)
    null_emb = make_null_condition_emb(cfg, seed=1)
    dec = mg.ref_decoder(cfg, w, S)
    enc, src, ctx, _ = mg.synth(cfg, 2, 40, 17, seed=41)
    ts = [1.0, 0.93, 0.8, 0.55, 0.3, 0.12, 0.0]
    fs = mg.FakeSelf(S, dec, null_emb, (enc, ctx))
    out = fs.generate(src, seed=[13, 14], infer_steps=99, diffusion_guidance_sale=6.0, shift=2.0,
                      timesteps=torch.tensor(ts), use_progress_bar=False)
    mg.save("sft_timesteps", enc=enc, ctx=ctx, src=src, seeds=[13, 14], timesteps=ts, null_emb=null_emb, out=out)
    enc2, _, ctx2, _ = mg.synth(cfg, 2, 40, 17, seed=42)
    fs = mg.FakeSelf(S, dec, null_emb, (enc, ctx), (enc2, ctx2))
    out = fs.generate(src, seed=15, infer_steps=99, diffusion_guidance_sale=4.0, shift=3.0,
                      timesteps=torch.tensor(ts), cover_noise_strength=0.25, audio_cover_strength=0.6,
                      use_progress_bar=False, non_cover_text_hidden_states=torch.zeros(1),
                      non_cover_text_attention_mask=torch.zeros(1))
    mg.save("sft_timesteps_cover", enc=enc, ctx=ctx, src=src, enc_nc=enc2, ctx_nc=ctx2, seed=15, timesteps=ts,
            null_emb=null_emb, out=out)


if __name__ == "__main__":
    main()
