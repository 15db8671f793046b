"""Drop-in façade for `acestep.acestep_v15_pipeline` (Gradio launcher) with the B200 backend.

`create_demo(init_params=None, language='en')` and `main()` are the reference's own functions
(acestep/acestep_v15_pipeline.py:54-84, 87-462); this module only grafts the backend onto
AceStepHandler first and, after the service initialises on a CUDA device, activates it
(ACESTEP_B200=0 keeps the stock PyTorch path).
"""
from __future__ import annotations

import os

from .backend import install

try:
    from acestep.handler import AceStepHandler
except ImportError as exc:  # pragma: no cover
    raise ImportError("acestep_b200.acestep_v15_pipeline needs the reference package `acestep` on sys.path") from exc

install(AceStepHandler)

_orig_initialize = AceStepHandler.initialize_service


def _initialize_service_b200(self, *args, **kwargs):
    status, ok = _orig_initialize(self, *args, **kwargs)
    if ok and os.environ.get("ACESTEP_B200", "1") != "0" and str(self.device).startswith("cuda"):
        dit_status, vae_status = self._init_b200_backends()
        status = f"{status}\nB200 DiT: {dit_status}\nB200 VAE: {vae_status}"
    return status, ok


AceStepHandler.initialize_service = _initialize_service_b200


def create_demo(*args, **kwargs):
    from acestep.acestep_v15_pipeline import create_demo as _create_demo

    return _create_demo(*args, **kwargs)


def main():
    from acestep.acestep_v15_pipeline import main as _main

    return _main()


if __name__ == "__main__":
    main()
