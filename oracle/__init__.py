"""CPU oracle for the ACE-Step 1.5 hot path (TEST INFRASTRUCTURE — never imported by the product).

Plain-PyTorch fp32 restatements of the reference algorithms on the path named by
BASELINE.json's north_star: the DiT velocity prediction, the turbo and base/sft samplers, APG/ADG
guidance, and the Oobleck VAE decode/encode with the handler's overlap-discard tiling.  Every
function cites the reference file:line it follows.

Pinning status (see DESIGN.md "Oracle"):
  * DiT forward, samplers, APG/ADG: PINNED — checked against outputs of the real reference
    modules imported from /root/reference (tools/make_golden.py -> tests/golden/*.npz).
  * Oobleck VAE: the arithmetic lives in third-party `diffusers` (unpinned version, not installed
    in this image, source absent from /root/reference).  The restatement follows the in-tree MLX
    re-implementation (acestep/models/mlx/vae_model.py, vae_convert.py).  PINNED against that
    implementation, executed unmodified through a torch-backed stand-in for the mlx primitives
    (tools/mlx_shim.py, tools/make_golden_vae_mlx.py -> tests/golden/vae_mlx_reference.npz);
    `diffusers` itself remains unchecked.  The tiling glue is pinned against the reference's own
    handler code driven with this oracle as the `vae` object.
  * Output path (peak / dB normalisation), lyric-alignment attentions, SFT timesteps: PINNED
    (tools/make_golden_output.py, make_golden_attn.py, make_golden_sft.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.
"""
