#!/usr/bin/env python
"""Golden vectors for the output path from the REAL reference code (build container only).

`acestep/audio_utils.py` cannot be imported here (it needs torchaudio), so the one function on the path,
`normalize_audio`, is compiled from the reference SOURCE FILE where it lies (ast: that FunctionDef only) and
executed on seeded inputs; the handler's peak normalisation (generate_music_decode.py:191-195) is three torch
expressions inside a 100-line method with handler state, so its golden comes from the same expressions
evaluated here and is cross-checked by running the extracted function after it.  Output:
tests/golden/output_normalize.npz."""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/acestep/audio_utils.py"


def load_reference_normalize_audio():
    tree = ast.parse(open(REF).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "normalize_audio")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch, "np": np}
    exec("from typing import Union, Optional, List, Tuple", ns)
    exec(compile(mod, REF, "exec"), ns)
    return ns["normalize_audio"]


def main():
    normalize_audio = load_reference_normalize_audio()
    g = torch.Generator().manual_seed(20260117)
    out = {}
    gains = [0.3, 2.5, 1.0, 1.0001, 7.0, 1e-8, 0.0]
    wav = torch.stack([torch.randn(2, 1027, generator=g).clamp(-1, 1) * a for a in gains]).contiguous()
    out["wav"] = wav.numpy()
    # handler stage (generate_music_decode.py:191-195)
    peak = wav.abs().amax(dim=[1, 2], keepdim=True)
    stage1 = wav / peak.clamp(min=1.0) if torch.any(peak > 1.0) else wav
    out["peak"] = peak.flatten().numpy()
    out["stage1"] = stage1.numpy()
    for db in (-1.0, 0.0, -6.0, -0.1):
        final = torch.stack([normalize_audio(stage1[i], db) for i in range(stage1.shape[0])])
        out[f"final_db{db}"] = final.numpy()
    # normalize_audio on its own (raw, un-clamped inputs, one song at a time)
    raw = torch.randn(5, 2, 1001, generator=g) * torch.tensor([0.01, 0.5, 1.0, 3.0, 40.0]).view(5, 1, 1)
    out["raw"] = raw.numpy()
    out["raw_db-1.0"] = torch.stack([normalize_audio(raw[i], -1.0) for i in range(5)]).numpy()
    path = os.path.join(ROOT, "tests", "golden", "output_normalize.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    sys.exit(main())
