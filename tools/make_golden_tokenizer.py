#!/usr/bin/env python
"""Generate tests/golden/tokenizer.npz by running the REAL reference AceStepAudioTokenizer / AudioTokenDetokenizer
(/root/reference/acestep/models/turbo/modeling_acestep_v15_turbo.py:1178-1218, :730-856, :859-990, imported
unmodified) and the model-level tokenize / detokenize / LM-hint substitution (:1577-1600, :1630-1646) on a tiny config
with oracle.tokenizer.make_tokenizer_weights tensors.

`vector_quantize_pytorch` (third-party, absent) is provided by oracle.tokenizer.ResidualFSQ — the restatement of its
published algorithm — so this fixture pins everything AROUND the quantizer against the reference's own code and the
quantizer against itself only (stated in oracle/tokenizer.py and DESIGN.md).

Runs only in the build container (the GPU box has no /root/reference); the fixture is committed.
Usage: python tools/make_golden_tokenizer.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle.tokenizer import ResidualFSQ, TokConfig, make_tokenizer_weights  # noqa: E402

stub = types.ModuleType("vector_quantize_pytorch")
stub.ResidualFSQ = ResidualFSQ
sys.modules["vector_quantize_pytorch"] = stub

from acestep.models.turbo import modeling_acestep_v15_turbo as T  # noqa: E402
from acestep.models.turbo.configuration_acestep_v15 import AceStepConfig  # noqa: E402

torch.set_grad_enabled(False)


def main():
    cfg = TokConfig.tiny()
    w = make_tokenizer_weights(cfg, seed=9)
    rc = AceStepConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, num_hidden_layers=4,
                       num_attention_heads=cfg.num_attention_heads, num_key_value_heads=cfg.num_key_value_heads,
                       head_dim=cfg.head_dim, sliding_window=cfg.sliding_window, rope_theta=cfg.rope_theta,
                       rms_norm_eps=cfg.rms_norm_eps, fsq_dim=cfg.fsq_dim, fsq_input_levels=cfg.fsq_input_levels,
                       fsq_input_num_quantizers=cfg.fsq_input_num_quantizers, pool_window_size=cfg.pool_window_size,
                       num_attention_pooler_hidden_layers=cfg.num_attention_pooler_hidden_layers)
    rc._attn_implementation = "sdpa"
    tok = T.AceStepAudioTokenizer(rc).float().eval()
    det = T.AudioTokenDetokenizer(rc).float().eval()
    for mod, prefix in ((tok, "tokenizer."), (det, "detokenizer.")):
        sd = {k[len(prefix):]: v for k, v in w.items() if k.startswith(prefix)}
        missing, unexpected = mod.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all("rotary_emb" in k for k in missing), missing

    # the model-level methods are plain functions of (self.config, self.tokenizer, self.detokenizer): run them
    # UNBOUND on a minimal host so the 2-billion-parameter model class is not instantiated
    host = types.SimpleNamespace(config=rc, tokenizer=tok, detokenizer=det)
    host.tokenize = types.MethodType(T.AceStepConditionGenerationModel.tokenize, host)
    host.detokenize = types.MethodType(T.AceStepConditionGenerationModel.detokenize, host)

    g = torch.Generator().manual_seed(31)
    B, Tn = 2, 23  # 23 % 5 != 0: exercises the silence-latent padding
    x = torch.randn(B, Tn, 64, generator=g)
    silence = torch.randn(1, 40, 64, generator=g)
    mask = torch.ones(B, Tn)
    mask[1, 17:] = 0
    src = torch.randn(B, Tn, 64, generator=g)
    is_covers = torch.tensor([1, 0])
    q, idx, m5 = host.tokenize(x, silence, mask)
    hints = host.detokenize(q)
    hints_c = hints[:, :Tn, :]
    new_src = torch.where(is_covers.unsqueeze(-1).unsqueeze(-1) > 0, hints_c, src)  # :1646
    out = os.path.join(ROOT, "tests", "golden", "tokenizer.npz")
    np.savez_compressed(out, x=x.numpy(), silence=silence.numpy(), mask=mask.numpy(), src=src.numpy(),
                        is_covers=is_covers.numpy(), quantized=q.numpy(), indices=idx.numpy().astype(np.int64),
                        mask5=m5.numpy(), hints=hints.numpy(), new_src=new_src.numpy())
    print("wrote", out, tuple(q.shape), tuple(idx.shape), tuple(hints.shape), "distinct codes:",
          len(set(idx.reshape(-1).tolist())))


if __name__ == "__main__":
    main()
