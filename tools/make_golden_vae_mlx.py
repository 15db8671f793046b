#!/usr/bin/env python
"""Golden vectors for the Oobleck codec arithmetic from the reference's OWN in-tree implementation (build
container only): `acestep/models/mlx/vae_model.py:MLXAutoEncoderOobleck` with weights converted by
`acestep/models/mlx/vae_convert.py:convert_vae_weights`, both unmodified, executed through tools/mlx_shim.py
(torch-backed primitives; `mlx` itself is Apple-only).  Two configs: the oracle's tiny one, and a three-stage one
with an odd stride (the shipped ratios contain 6 and 10: ceil(stride / 2) paddings differ for odd strides).
-> tests/golden/vae_mlx_reference.npz"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mlx_shim  # noqa: E402

mlx_shim.install()


def load_ref(name):
    """Import one file of acestep/models/mlx/ without the package __init__ (which pulls the whole model zoo)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(name, f"/root/reference/acestep/models/mlx/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


vae_model = load_ref("vae_model")
vae_convert = load_ref("vae_convert")

from oracle import vae as ovae  # noqa: E402
from oracle.weights import make_vae_weights  # noqa: E402


def run(cfg, seed, frames, tag, out):
    sd = make_vae_weights(cfg, seed=seed)
    fake_pt = types.SimpleNamespace(state_dict=lambda: sd)
    m = vae_model.MLXAutoEncoderOobleck(
        encoder_hidden_size=cfg.encoder_hidden_size, downsampling_ratios=list(cfg.downsampling_ratios),
        channel_multiples=list(cfg.channel_multiples), decoder_channels=cfg.decoder_channels,
        decoder_input_channels=cfg.decoder_input_channels, audio_channels=cfg.audio_channels)
    vae_convert.convert_and_load(fake_pt, m)
    g = torch.Generator().manual_seed(seed + 100)
    z = torch.randn(1, frames, cfg.decoder_input_channels, generator=g)           # NLC
    wav = torch.rand(1, frames * cfg.hop, cfg.audio_channels, generator=g) - 0.5   # NLC
    with torch.no_grad():
        audio = m.decode(z)
        mean = m.encode_mean(wav)
        h = m.encoder(wav)
    out[f"{tag}_z"] = z.numpy()
    out[f"{tag}_wav"] = wav.numpy()
    out[f"{tag}_decoded"] = audio.numpy()
    out[f"{tag}_mean"] = mean.numpy()
    out[f"{tag}_moments"] = h.numpy()
    out[f"{tag}_seed"] = np.asarray(seed)
    print(tag, "decoded", tuple(audio.shape), "mean", tuple(mean.shape))


def main():
    out = {}
    run(ovae.VaeConfig.tiny(), 3, 12, "tiny", out)
    odd = ovae.VaeConfig(encoder_hidden_size=32, downsampling_ratios=[2, 3, 5], channel_multiples=[1, 2, 4],
                         decoder_channels=32, decoder_input_channels=16, audio_channels=2)
    run(odd, 7, 9, "odd", out)
    path = os.path.join(ROOT, "tests", "golden", "vae_mlx_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
