"""CPU tests: the C-ABI library loads and exports every symbol include/acestep_b200.h declares
(no compute calls), and the Python packers produce blobs of exactly the size the C walkers expect."""
import ctypes as C
import os
import re

import pytest
import torch

from acestep_b200 import _lib
from acestep_b200.build import LIB_PATH, build
from acestep_b200.dit import DiTShape
from acestep_b200.pack import fold_weight_norm, pack_dit, pack_vae
from acestep_b200.synthetic import random_dit_state, random_vae_state
from acestep_b200.vae import VaeShape

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB_PATH):
        build()
    return _lib.load()


def _header_functions(probe_section: bool):
    """Function names the header declares outside (False) / inside (True) its `#ifdef ACE_PROBE` block."""
    src = open(os.path.join(ROOT, "include", "acestep_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    m = re.search(r"#ifdef ACE_PROBE(.*?)#endif", src, flags=re.S)
    assert m, "header has no ACE_PROBE section"
    part = m.group(1) if probe_section else src.replace(m.group(0), "")
    return sorted(set(re.findall(r"\b(ace_[a-z0-9_]+)\s*\(", part)))


def test_library_exports_every_declared_symbol(lib):
    names = _header_functions(False)
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libacestep_b200.so does not export {n}"
    # and the ctypes binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_release_library_has_no_debug_switches(lib):
    """Release hygiene: the A/B hooks and every ACE_* environment switch exist only in the probe build
    (libacestep_b200_probe.so, -DACE_PROBE); the release library exports none and never reads the environment."""
    probe_names = _header_functions(True)
    assert sorted(_lib.PROBE_SIGNATURES) == probe_names and len(probe_names) == 3
    for n in probe_names:
        assert not hasattr(lib, n), f"release library exports the probe hook {n}"
    blob = open(LIB_PATH, "rb").read()
    for needle in (b"ACE_SKIP", b"ACE_ATTN", b"ACE_GEMM_BN", b"ACE_NO_GRAPH", b"ACE_NO_PDL", b"ACE_PREFETCH",
                   b"ACE_VAE_FUSED", b"ACE_NO_SPLITK", b"ACE_TMAP_L2", b"gemm_ref_kernel", b"getenv"):
        assert needle not in blob, f"release library still contains {needle!r}"
    probe = _lib.load_probe()
    for n in probe_names + _header_functions(False):
        assert hasattr(probe, n), f"probe library does not export {n}"


def test_abi_version_and_error_string(lib):
    assert lib.ace_abi_version() == 1
    assert isinstance(lib.ace_last_error(), bytes)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.B200Error):
        _lib.load()


def _dit_cfg(shape):
    cfg = _lib.AceDitConfig()
    cfg.hidden_size, cfg.intermediate_size = shape.hidden_size, shape.intermediate_size
    cfg.num_layers, cfg.num_heads = shape.num_hidden_layers, shape.num_attention_heads
    cfg.num_kv_heads, cfg.head_dim = shape.num_key_value_heads, shape.head_dim
    return cfg


@pytest.mark.parametrize("shape", [
    DiTShape(hidden_size=256, intermediate_size=512, num_hidden_layers=4, num_attention_heads=2, num_key_value_heads=1),
    DiTShape(hidden_size=512, intermediate_size=1024, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2),
])
def test_dit_pack_size_matches_library(lib, shape):
    blob = pack_dit(random_dit_state(shape, 0, "cpu", torch.float32), shape.num_hidden_layers)
    assert blob.dtype == torch.bfloat16 and blob.dim() == 1
    assert blob.numel() == lib.ace_dit_packed_elems(C.byref(_dit_cfg(shape)))


def test_full_size_dit_param_count(lib):
    n = lib.ace_dit_packed_elems(C.byref(_dit_cfg(DiTShape())))
    assert 1.55e9 < n < 1.75e9  # the 24-layer decoder (4.7 GB bf16 incl. condition encoders)


def test_dit_pack_layout():
    shape = DiTShape(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                     num_key_value_heads=1)
    sd = random_dit_state(shape, 0, "cpu", torch.float32)
    blob = pack_dit(sd, 2).float()
    D = 256
    # proj_in is first: [D, (k, c)] with k-major columns
    w = sd["proj_in.1.weight"]
    assert torch.equal(blob[:D * 384].view(D, 384)[:, :192], w[:, :, 0].to(torch.bfloat16).float())
    assert torch.equal(blob[:D * 384].view(D, 384)[:, 192:], w[:, :, 1].to(torch.bfloat16).float())
    # gate/up interleave in blocks of 64 rows (tail of layer 1 is [gate_up | down])
    I = 512
    tail = blob[-(2 * I * D + D * I):]
    gu = tail[: 2 * I * D].view(2 * I, D)
    g, u = sd["layers.1.mlp.gate_proj.weight"], sd["layers.1.mlp.up_proj.weight"]
    assert torch.equal(gu[0:64], g[0:64].to(torch.bfloat16).float())
    assert torch.equal(gu[64:128], u[0:64].to(torch.bfloat16).float())
    assert torch.equal(gu[128:192], g[64:128].to(torch.bfloat16).float())


@pytest.mark.parametrize("ratios,mult", [([2, 4], [1, 2]), ([2, 4, 4, 6, 10], [1, 2, 4, 8, 16])])
def test_vae_pack_size_matches_library(lib, ratios, mult):
    shape = VaeShape(downsampling_ratios=ratios, channel_multiples=mult)
    blob = pack_vae(random_vae_state(shape, 0), ratios, mult)
    cfg = _lib.AceVaeConfig()
    cfg.num_stages = len(ratios)
    for i, (r, m) in enumerate(zip(ratios, mult)):
        cfg.ratios[i], cfg.channel_multiples[i] = r, m
    cfg.encoder_hidden, cfg.decoder_channels, cfg.latent_channels, cfg.audio_channels = 128, 128, 64, 2
    assert blob.dtype == torch.uint8
    assert blob.numel() == lib.ace_vae_packed_bytes(C.byref(cfg))


def test_weight_norm_fold_matches_torch():
    conv = torch.nn.utils.weight_norm(torch.nn.Conv1d(8, 16, 7))
    sd = {"c.weight_g": conv.weight_g.detach(), "c.weight_v": conv.weight_v.detach()}
    assert torch.allclose(fold_weight_norm(sd, "c"), conv.weight.detach(), atol=1e-6)
    convt = torch.nn.utils.weight_norm(torch.nn.ConvTranspose1d(8, 16, 4, stride=2))
    sd = {"c.weight_g": convt.weight_g.detach(), "c.weight_v": convt.weight_v.detach()}
    assert torch.allclose(fold_weight_norm(sd, "c"), convt.weight.detach(), atol=1e-6)


def test_engines_refuse_cpu():
    shape = DiTShape(hidden_size=256, intermediate_size=512, num_hidden_layers=1, num_attention_heads=2,
                     num_key_value_heads=1)
    from acestep_b200.dit import B200DiT

    with pytest.raises(_lib.B200Error):
        B200DiT(random_dit_state(shape, 0, "cpu", torch.float32), shape, "cpu")


def test_argument_errors_are_reported_before_any_device_work(lib):
    """Entry points validate their arguments first and report through the int status + ace_last_error(): these
    calls never reach a CUDA API, so they behave the same on a box without a GPU."""
    assert lib.ace_peak_normalize_db(None, 1, 8, None, 0.0, None) != 0
    assert b"target_amp" in lib.ace_last_error()
    assert lib.ace_peak_normalize(None, 1, 8, None, None) != 0
    assert b"null argument" in lib.ace_last_error()
    assert lib.ace_peak_normalize(None, 0, 8, None, None) == 0  # empty batch: nothing to do
    assert lib.ace_latent_guard(None, 8, None, None) != 0
    assert lib.ace_dit_cross_attentions(None, None, None, None, 1, None, None) != 0
    assert b"not bound" in lib.ace_last_error()
