#!/usr/bin/env python
"""Condition encoder at the shipped size: CUDA-event time of B200ConditionEncoder vs
the fp32 oracle on the host CPU, same inputs (random-init weights, synthetic embeddings).

    python tools/time_cond.py [--lyric 2048] [--text 256] [--refer 750] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acestep_b200.cond import B200ConditionEncoder, CondShape
from oracle.cond import CondConfig, condition_encoder, make_cond_weights

ap = argparse.ArgumentParser()
ap.add_argument("--lyric", type=int, default=2048)
ap.add_argument("--text", type=int, default=256)
ap.add_argument("--refer", type=int, default=750)
ap.add_argument("--no-cpu", action="store_true")
a = ap.parse_args()

cfg = CondConfig()
dev = torch.device("cuda:0")
w = make_cond_weights(cfg, seed=5)
g = torch.Generator().manual_seed(0)
text = torch.randn(1, a.text, cfg.text_hidden_dim, generator=g)
lyric = torch.randn(1, a.lyric, cfg.text_hidden_dim, generator=g)
refer = torch.randn(1, a.refer, cfg.timbre_hidden_dim, generator=g)
tm, lm = torch.ones(1, a.text, dtype=torch.long), torch.ones(1, a.lyric, dtype=torch.long)
order = torch.zeros(1, dtype=torch.long)
enc = B200ConditionEncoder(w, CondShape.from_config(cfg), dev)
args_d = dict(text_hidden_states=text.to(dev), text_attention_mask=tm.to(dev), lyric_hidden_states=lyric.to(dev),
              lyric_attention_mask=lm.to(dev), refer_audio_acoustic_hidden_states_packed=refer.to(dev),
              refer_audio_order_mask=order.to(dev))
for _ in range(3):
    out, mask = enc(**args_d)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    out, mask = enc(**args_d)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n


def stack_flops(S, layers, in_dim):
    D, I, NQ, NKV, W = cfg.hidden_size, cfg.intermediate_size, 2048, 1024, cfg.sliding_window
    per_tok = 2 * ((NQ + 2 * NKV) * D + NQ * D + 3 * I * D)
    attn = sum(4 * S * (min(S, 2 * W + 1) if (i + 1) % 2 else S) * NQ for i in range(layers))
    return S * (layers * per_tok + 2 * in_dim * D) + attn


flops = (stack_flops(a.lyric, cfg.num_lyric_encoder_hidden_layers, cfg.text_hidden_dim) +
         stack_flops(a.refer, cfg.num_timbre_encoder_hidden_layers, cfg.timbre_hidden_dim) +
         2 * a.text * cfg.text_hidden_dim * cfg.hidden_size)
res = {"lyric_tokens": a.lyric, "text_tokens": a.text, "refer_frames": a.refer, "gpu_ms": ms,
       "algorithmic_tflop": flops / 1e12, "gpu_tflops": flops / (ms * 1e-3) / 1e12, "finite": bool(torch.isfinite(out).all())}
if not a.no_cpu:
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        condition_encoder(w, cfg, text, tm, lyric, lm, refer, order)
        t0 = time.perf_counter()
        want, _ = condition_encoder(w, cfg, text, tm, lyric, lm, refer, order)
        res["cpu_oracle_s"] = time.perf_counter() - t0
    res["cpu_cores"] = os.cpu_count()
    res["rel_l2_vs_fp32_oracle"] = float((out.cpu().float() - want).norm() / want.norm())
print(json.dumps(res))
