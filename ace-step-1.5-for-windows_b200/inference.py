"""Drop-in façade for `acestep.inference` with the B200 backend pre-installed.

    from acestep_b200.inference import generate_music, GenerationParams, GenerationConfig

Re-exports the reference's public dataclasses and `generate_music(dit_handler, llm_handler, params,
config, save_dir=None, progress=None) -> GenerationResult` (acestep/inference.py:38-221, 309-776)
unchanged; importing this module grafts `B200BackendMixin` onto `AceStepHandler` (see backend.py),
so front-ends (Gradio UI, FastAPI server, CLI) keep calling the same function with the same
arguments.  The reference package must be importable (it is reused, not re-implemented: SURVEY §7).
"""
from __future__ import annotations

try:
    import acestep.inference as _ref
    from acestep.handler import AceStepHandler
except ImportError as exc:  # pragma: no cover - exercised only without the reference installed
    raise ImportError(
        "acestep_b200.inference is a façade over the reference package `acestep`; put the ACE-Step "
        "checkout on sys.path (the B200 backend replaces its DiT/VAE hot path, not its front-end)") from exc

from .backend import install

install(AceStepHandler)

GenerationParams = _ref.GenerationParams
GenerationConfig = _ref.GenerationConfig
GenerationResult = _ref.GenerationResult
generate_music = _ref.generate_music
understand_music = getattr(_ref, "understand_music", None)
create_sample = getattr(_ref, "create_sample", None)
format_sample = getattr(_ref, "format_sample", None)


def enable_b200(dit_handler, dit: bool = True, vae: bool = True):
    """Call after `dit_handler.initialize_service(...)`: packs the loaded weights and switches the
    handler's DiT / VAE seams to the B200 engines.  Returns (dit_status, vae_status) strings like
    `_initialize_mlx_backends` (handler/init_service_setup.py:116-148)."""
    return dit_handler._init_b200_backends(dit=dit, vae=vae)


__all__ = ["GenerationParams", "GenerationConfig", "GenerationResult", "generate_music", "understand_music",
           "create_sample", "format_sample", "enable_b200", "AceStepHandler"]
