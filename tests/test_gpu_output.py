"""GPU parity of the output path against the reference's own torch expressions:
per-sample peak normalisation (handler/generate_music_decode.py:191-195) and the latent sanity guard
(:66-77).  The arithmetic is a max and one IEEE division, so the bar is BIT-EXACT."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from acestep_b200 import _lib  # noqa: E402
from acestep_b200.output import check_latents, latent_flags, peak_normalize_  # noqa: E402

DEV = torch.device("cuda:0")


def reference_normalize(pred_wavs):
    """generate_music_decode.py:191-195, verbatim semantics (runs on the CPU copy)."""
    peak = pred_wavs.abs().amax(dim=[1, 2], keepdim=True)
    if torch.any(peak > 1.0):
        pred_wavs = pred_wavs / peak.clamp(min=1.0)
    return pred_wavs, peak.flatten()


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 1920 * 48, 2_880_000])
@pytest.mark.parametrize("gains", [(0.3,), (2.5,), (0.9, 3.0, 1.0, 1.0001), (7.0, 0.5)])
def test_peak_normalize_bit_exact(n, gains):
    g = torch.Generator().manual_seed(n * 31 + len(gains))
    wav = torch.stack([torch.randn(2, n, generator=g).clamp(-1, 1) * a for a in gains]).contiguous()
    want, want_peak = reference_normalize(wav.clone())
    got = wav.to(DEV)
    peak = peak_normalize_(got)
    torch.cuda.synchronize()
    assert torch.equal(peak.cpu(), want_peak)
    assert torch.equal(got.cpu(), want)


def test_peak_normalize_unaligned_view_and_single_sample():
    g = torch.Generator().manual_seed(5)
    base = (torch.randn(2 * 1001 + 1, generator=g) * 3).to(DEV)
    wav = base[1:].view(2, 1001)  # 4-byte aligned only: the scalar path
    want, want_peak = reference_normalize(wav.cpu().clone().unsqueeze(0))
    peak = peak_normalize_(wav)
    assert torch.equal(peak.cpu(), want_peak) and torch.equal(wav.cpu(), want[0])


def test_peak_normalize_edges():
    empty = torch.empty(0, 2, 100, dtype=torch.float32, device=DEV)
    assert peak_normalize_(empty).numel() == 0
    zero_len = torch.empty(3, 2, 0, dtype=torch.float32, device=DEV)
    assert torch.equal(peak_normalize_(zero_len).cpu(), torch.zeros(3))
    silent = torch.zeros(2, 2, 777, dtype=torch.float32, device=DEV)
    assert torch.equal(peak_normalize_(silent).cpu(), torch.zeros(2)) and float(silent.abs().max()) == 0.0
    big = torch.full((1, 2, 64), -float("inf"), device=DEV)
    big[0, 0, 0] = 2.0
    assert float(peak_normalize_(big)[0]) == float("inf")  # x / inf like the reference: finite -> 0, inf -> NaN
    assert float(big[0, 0, 0]) == 0.0 and torch.isnan(big[0, 1, 3])
    with pytest.raises(ValueError):
        peak_normalize_(torch.zeros(2, 2, 8, dtype=torch.bfloat16, device=DEV))
    with pytest.raises(_lib.B200Error):
        peak_normalize_(torch.zeros(1, 2, 8))  # CPU tensor: no fallback


def test_peak_normalize_nan_sample_untouched():
    wav = torch.full((2, 2, 100), 3.0, device=DEV)
    wav[0, 1, 50] = float("nan")
    peak = peak_normalize_(wav)
    assert torch.isnan(peak[0]) and float(peak[1]) == 3.0
    assert float(wav[0, 0, 0]) == 3.0 and float(wav[1, 0, 0]) == 1.0


@pytest.mark.parametrize("shape", [(1, 1, 64), (2, 250, 64), (1, 6000, 64), (3, 77, 64)])
def test_latent_guard_matches_reference_predicates(shape):
    g = torch.Generator().manual_seed(shape[1])
    lat = torch.randn(*shape, generator=g).to(torch.bfloat16).to(DEV)

    def ref(x):
        x = x.cpu()
        return bool(torch.isnan(x).any() or torch.isinf(x).any()), bool(x.abs().sum() != 0)

    assert latent_flags(lat) == ref(lat) == (False, True)
    check_latents(lat)
    for bad in (float("nan"), float("inf"), -float("inf")):
        x = lat.clone()
        x.view(-1)[x.numel() - 1] = bad
        assert latent_flags(x) == (True, True)
        with pytest.raises(RuntimeError, match="NaN or Inf latents"):
            check_latents(x)
    z = torch.zeros_like(lat)
    assert latent_flags(z) == (False, False)
    with pytest.raises(RuntimeError, match="zero latents"):
        check_latents(z)
    z.view(-1)[z.numel() // 2] = -0.0  # negative zero is still zero (abs().sum() == 0)
    assert latent_flags(z) == (False, False)
    z.view(-1)[0] = 1e-30  # a subnormal-range bf16 is non-zero
    assert latent_flags(z)[1] == bool(z.cpu().abs().sum() != 0)
    check_latents(torch.empty(0, 4, 64, dtype=torch.bfloat16, device=DEV))  # empty: nothing to flag
