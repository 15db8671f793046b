"""ctypes binding of libacestep_b200.so (the C ABI declared in include/acestep_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, an exception
is raised (`B200Error`); nothing silently routes to PyTorch or the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# ACE_B200_LIB: alternative build of the same ABI (A/B timing of two kernel versions on one box)
LIB_PATH = os.environ.get("ACE_B200_LIB") or os.path.join(PKG_DIR, "libacestep_b200.so")


class B200Error(RuntimeError):
    """Raised when libacestep_b200 is unavailable or an entry point reports a failure."""


class AceDitConfig(C.Structure):
    _fields_ = [
        ("hidden_size", C.c_int), ("intermediate_size", C.c_int), ("num_layers", C.c_int),
        ("num_heads", C.c_int), ("num_kv_heads", C.c_int), ("head_dim", C.c_int),
        ("sliding_window", C.c_int), ("layer_is_sliding", C.c_int * 64),
        ("rope_theta", C.c_float), ("rms_eps", C.c_float),
    ]


class AceVaeConfig(C.Structure):
    _fields_ = [
        ("num_stages", C.c_int), ("ratios", C.c_int * 8), ("channel_multiples", C.c_int * 8),
        ("encoder_hidden", C.c_int), ("decoder_channels", C.c_int), ("latent_channels", C.c_int),
        ("audio_channels", C.c_int),
    ]


class AceEncConfig(C.Structure):
    _fields_ = [
        ("hidden_size", C.c_int), ("intermediate_size", C.c_int), ("num_layers", C.c_int),
        ("num_heads", C.c_int), ("num_kv_heads", C.c_int), ("head_dim", C.c_int),
        ("sliding_window", C.c_int), ("layer_is_sliding", C.c_int * 64), ("in_dim", C.c_int),
        ("rope_theta", C.c_float), ("rms_eps", C.c_float),
    ]


_P = C.c_void_p
# name -> (restype, argtypes); must list every symbol of include/acestep_b200.h
SIGNATURES = {
    "ace_last_error": (C.c_char_p, []),
    "ace_abi_version": (C.c_int, []),
    "ace_init": (C.c_int, [C.c_int]),
    "ace_dit_packed_elems": (C.c_size_t, [C.POINTER(AceDitConfig)]),
    "ace_dit_create": (C.c_int, [C.POINTER(_P), C.POINTER(AceDitConfig), _P, C.c_size_t]),
    "ace_dit_destroy": (None, [_P]),
    "ace_dit_workspace_bytes": (C.c_size_t, [_P, C.c_int, C.c_int, C.c_int]),
    "ace_dit_bind": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_size_t]),
    "ace_dit_set_condition": (C.c_int, [_P, _P, _P]),
    "ace_dit_step": (C.c_int, [_P, _P, _P, C.POINTER(C.c_float), _P, _P]),
    "ace_dit_prepare_timesteps": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int, _P]),
    "ace_dit_cross_attentions": (C.c_int, [_P, _P, _P, C.POINTER(C.c_float), C.c_int, _P, _P]),
    "ace_euler_step": (C.c_int, [_P, _P, C.c_float, C.c_size_t, _P]),
    "ace_euler_step_dup": (C.c_int, [_P, _P, C.c_float, C.c_size_t, _P, _P]),
    "ace_sde_step": (C.c_int, [_P, _P, _P, C.c_float, C.c_float, C.c_size_t, _P]),
    "ace_apg": (C.c_int, [_P, _P, _P, C.c_int, C.c_float, C.c_float, C.c_float, _P, C.c_int, C.c_int, _P]),
    "ace_adg": (C.c_int, [_P, _P, _P, C.c_float, C.c_float, C.c_float, _P, C.c_int, C.c_int, _P]),
    "ace_peak_normalize": (C.c_int, [_P, C.c_int, C.c_size_t, _P, _P]),
    "ace_peak_normalize_db": (C.c_int, [_P, C.c_int, C.c_size_t, _P, C.c_float, _P]),
    "ace_latent_guard": (C.c_int, [_P, C.c_size_t, _P, _P]),
    "ace_vae_packed_bytes": (C.c_size_t, [C.POINTER(AceVaeConfig)]),
    "ace_vae_create": (C.c_int, [C.POINTER(_P), C.POINTER(AceVaeConfig), _P, C.c_size_t]),
    "ace_vae_destroy": (None, [_P]),
    "ace_vae_decode_workspace_bytes": (C.c_size_t, [_P, C.c_int]),
    "ace_vae_encode_workspace_bytes": (C.c_size_t, [_P, C.c_int]),
    "ace_vae_decode": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "ace_vae_encode": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "ace_vae_encode_moments": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "ace_vae_posterior_sample": (C.c_int, [_P, _P, _P, _P, C.c_int, _P]),
    "ace_dit_io_slots": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "ace_enc_packed_elems": (C.c_size_t, [C.POINTER(AceEncConfig)]),
    "ace_enc_create": (C.c_int, [C.POINTER(_P), C.POINTER(AceEncConfig), _P, C.c_size_t]),
    "ace_enc_destroy": (None, [_P]),
    "ace_enc_workspace_bytes": (C.c_size_t, [_P, C.c_int, C.c_int]),
    "ace_enc_forward": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "ace_linear": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "ace_fsq": (C.c_int, [_P, _P, _P, C.POINTER(C.c_int), C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "ace_launch_count": (C.c_uint64, []),
    "ace_profile_start": (None, []),
    "ace_profile_stop": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.POINTER(C.c_int)]),
    "ace_profile_gemm_shapes": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "ace_attention": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
}
# exported only by the probe build (libacestep_b200_probe.so, -DACE_PROBE): A/B hooks for tools/ and the
# two-path equivalence tests; the release library has none of them
PROBE_SIGNATURES = {
    "ace_debug_set_gemm_reference": (None, [C.c_int]),
    "ace_debug_set_vae_fused": (None, [C.c_int]),
    "ace_debug_set_attention_p_in_tmem": (None, [C.c_int]),
}
PROBE_LIB_PATH = os.path.join(PKG_DIR, "libacestep_b200_probe.so")

_lib: Optional[C.CDLL] = None
_probe: Optional[C.CDLL] = None


def _open(path: str, signatures) -> C.CDLL:
    if not os.path.exists(path):
        raise B200Error(
            f"{path} not found — build it with `python -m acestep_b200.build` "
            "(the B200 backend has no CPU/PyTorch fallback)")
    try:
        lib = C.CDLL(path)
    except OSError as exc:
        raise B200Error(f"cannot load {path}: {exc}") from exc
    for name, (res, args) in signatures.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise B200Error(f"{path} does not export {name}") from exc
        fn.restype = res
        fn.argtypes = args
    return lib


def load() -> C.CDLL:
    """Load the shared library (once) and declare every prototype.  Raises B200Error if absent."""
    global _lib
    if _lib is None:
        _lib = _open(LIB_PATH, SIGNATURES)
    return _lib


def load_probe() -> C.CDLL:
    """The probe build (same ABI + the ace_debug_set_* hooks and ACE_* environment switches).  Tests and
    tools only: engines take it through their `lib=` argument; nothing in the product path calls this."""
    global _probe
    if _probe is None:
        _probe = _open(PROBE_LIB_PATH, {**SIGNATURES, **PROBE_SIGNATURES})
    return _probe


def check(status: int, what: str = "", lib: Optional[C.CDLL] = None) -> None:
    if status != 0:
        msg = (lib or load()).ace_last_error()
        raise B200Error(f"{what or 'libacestep_b200 call'} failed (status {status}): "
                        f"{msg.decode(errors='replace') if msg else '?'}")


def ptr(t) -> int:
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def stream_handle(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream
