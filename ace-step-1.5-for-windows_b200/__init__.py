"""acestep_b200 — B200-native backend for the ACE-Step 1.5 hot path (DiT denoising loop + Oobleck
VAE decode/encode) behind the reference's backend seam.  See DESIGN.md / INTEGRATION.md.

Layout: csrc/ (CUDA kernels + C ABI, built to libacestep_b200.so), _lib.py (ctypes binding),
pack.py (weight packers), dit.py / sampler.py / vae.py (host-side mirrors of the reference
interfaces), backend.py (handler mixin + install()), multi_gpu.py (song sharding + gather).
"""
__version__ = "0.1.0"
