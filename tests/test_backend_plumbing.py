"""CPU tests of the drop-in boundary: install() wraps exactly the three selection sites, the B200
branch forwards the reference's arguments with the MLX sibling's contract, inactive == neutral.

Engines are stubbed (no GPU here); numerics of the real engines are covered by the -m gpu tests.
When the reference checkout is present (/root/reference, build container only) the graft is also
applied to the REAL AceStepHandler class and the reference's own hot-path plumbing tests are run
against it.
"""
import os
import sys
import types
import unittest

import pytest
import torch

from acestep_b200.backend import B200BackendMixin, install

REF = "/root/reference"


class _Ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class FakeModel:
    """prepare_condition stand-in returning recognisable tensors."""

    null_condition_emb = torch.zeros(1, 1, 8)

    def __init__(self):
        self.calls = []

    def prepare_condition(self, **kw):
        self.calls.append(kw)
        b, t = kw["src_latents"].shape[:2]
        tag = float(len(self.calls))
        return torch.full((b, 5, 8), tag), torch.ones(b, 5), torch.full((b, t, 128), tag)


class FakeHandler:
    """Minimal host with the three reference selection sites."""

    def __init__(self):
        self.device, self.dtype = "cpu", torch.float32
        self.model = FakeModel()
        self.config = types.SimpleNamespace(is_turbo=False)
        self.silence_latent = torch.zeros(1, 100, 64)
        self.ref_calls = []

    def _load_model_context(self, name):
        return _Ctx()

    def _execute_service_generate_diffusion(self, payload, generate_kwargs, seed_param, infer_method, shift,
                                            audio_cover_strength):
        self.ref_calls.append("dit")
        return {"target_latents": torch.zeros(1)}, None, None, None

    def tiled_decode(self, latents, chunk_size=None, overlap=64, offload_wav_to_cpu=None):
        self.ref_calls.append("decode")
        return torch.zeros(latents.shape[0], 2, latents.shape[2] * 1920)

    def tiled_encode(self, audio, chunk_size=None, overlap=None, offload_latent_to_cpu=True):
        self.ref_calls.append("encode")
        return torch.zeros(1, 64, 1)


class StubSampler:
    def __init__(self):
        self.calls = []

    def generate_base(self, enc, ctx, src, seed, **kw):
        self.calls.append(("base", enc, ctx, src, seed, kw))
        return {"target_latents": torch.ones(src.shape, dtype=torch.bfloat16),
                "time_costs": {"diffusion_time_cost": 1.0, "diffusion_per_step_time_cost": 0.1, "total_time_cost": 1.0}}

    def generate_turbo(self, enc, ctx, src, seed, **kw):
        self.calls.append(("turbo", enc, ctx, src, seed, kw))
        return {"target_latents": torch.ones(src.shape, dtype=torch.bfloat16),
                "time_costs": {"diffusion_time_cost": 1.0, "diffusion_per_step_time_cost": 0.1, "total_time_cost": 1.0}}


class StubVae:
    def decode(self, lat):
        return torch.full((lat.shape[0], 2, lat.shape[2] * 1920), 0.5)

    decode_normalized = decode  # the codec seam asks for the peak-normalised waveform

    def encode(self, audio, sample=True):
        return torch.full((audio.shape[0], 64, audio.shape[2] // 1920), 0.25)

    encode_cached = encode  # the encode seam goes through the posterior-moments cache


def _payload(b=2, t=20):
    z = torch.zeros(1)
    return {"text_hidden_states": z, "text_attention_mask": z, "lyric_hidden_states": z, "lyric_attention_mask": z,
            "refer_audio_acoustic_hidden_states_packed": z, "refer_audio_order_mask": z,
            "src_latents": torch.zeros(b, t, 64), "chunk_mask": z, "is_covers": torch.zeros(b),
            "precomputed_lm_hints_25Hz": None, "non_cover_text_hidden_states": z,
            "non_cover_text_attention_masks": z}


def test_install_is_neutral_until_enabled():
    h = install(FakeHandler())
    assert h._b200_installed and not h.use_b200_dit and not h.use_b200_vae
    h._execute_service_generate_diffusion(_payload(), {}, 1, "ode", 3.0, 1.0)
    h.tiled_decode(torch.zeros(1, 64, 4))
    h.tiled_encode(torch.zeros(1, 2, 1920))
    assert h.ref_calls == ["dit", "decode", "encode"]
    assert install(h) is h  # idempotent


def test_dit_seam_forwards_reference_arguments():
    h = install(FakeHandler())
    h.b200_sampler, h.use_b200_dit = StubSampler(), True
    kw = {"infer_steps": 27, "diffusion_guidance_sale": 6.5, "cfg_interval_start": 0.1, "cfg_interval_end": 0.9,
          "use_adg": True, "cover_noise_strength": 0.2, "timesteps": None}
    out, enc, mask, ctx = h._execute_service_generate_diffusion(_payload(), kw, [3, 4], "sde", 2.0, 0.5)
    assert h.ref_calls == []
    kind, enc_a, ctx_a, src_a, seed, skw = h.b200_sampler.calls[0]
    assert kind == "base" and seed == [3, 4]
    assert torch.equal(enc_a, enc) and torch.equal(ctx_a, ctx)
    assert skw["infer_method"] == "sde" and skw["shift"] == 2.0 and skw["infer_steps"] == 27
    assert skw["diffusion_guidance_sale"] == 6.5 and skw["use_adg"] is True and skw["cover_noise_strength"] == 0.2
    assert skw["cfg_interval_start"] == 0.1 and skw["cfg_interval_end"] == 0.9 and skw["audio_cover_strength"] == 0.5
    # audio_cover_strength < 1 -> the non-cover conditioning is prepared (second prepare_condition call)
    assert len(h.model.calls) == 2 and skw["encoder_hidden_states_non_cover"] is not None
    assert float(skw["context_latents_non_cover"][0, 0, 0]) == 2.0
    # same return contract as the MLX sibling: latents on self.device in self.dtype + the three time_costs keys
    assert out["target_latents"].dtype == torch.float32 and out["target_latents"].shape == (2, 20, 64)
    assert {"diffusion_time_cost", "diffusion_per_step_time_cost", "total_time_cost"} <= set(out["time_costs"])


def test_condition_seam_replaces_model_encoder_only():
    """use_b200_cond: the encoder call of prepare_condition (:1621-1628) goes to b200_cond with the
    reference's keywords; context_latents = [src | chunk_mask] (:1651); the stock prepare_condition is
    not called; with is_covers > 0 the precomputed LM hints replace the source latents (:1649)."""
    h = install(FakeHandler())
    h.b200_sampler, h.use_b200_dit = StubSampler(), True
    seen = []

    def fake_cond(**kw):
        seen.append(kw)
        b = kw["text_hidden_states"].shape[0]
        return torch.full((b, 6, 8), 7.0, dtype=torch.bfloat16), torch.ones(b, 6, dtype=torch.bool)

    h.b200_cond, h.use_b200_cond = fake_cond, True
    p = _payload()
    p["text_hidden_states"] = torch.zeros(2, 3, 4)
    p["src_latents"] = torch.full((2, 20, 64), 0.25)
    p["chunk_mask"] = torch.ones(2, 20, 64, dtype=torch.bool)
    p["is_covers"] = torch.tensor([0.0, 1.0])
    p["precomputed_lm_hints_25Hz"] = torch.full((2, 24, 64), 0.5)
    out, enc, mask, ctx = h._execute_service_generate_diffusion(p, {}, 1, "ode", 3.0, 1.0)
    assert h.model.calls == [] and len(seen) == 1
    assert set(seen[0]) == {"text_hidden_states", "text_attention_mask", "lyric_hidden_states", "lyric_attention_mask",
                            "refer_audio_acoustic_hidden_states_packed", "refer_audio_order_mask"}
    assert enc.dtype == torch.float32 and float(enc[0, 0, 0]) == 7.0 and mask.shape == (2, 6)
    assert ctx.shape == (2, 20, 128) and float(ctx[0, 0, 0]) == 0.25 and float(ctx[1, 0, 0]) == 0.5
    assert float(ctx[0, 0, 64]) == 1.0


def test_effective_decoder_state_folds_active_lora():
    """PEFT-shaped LoRA layers (duck-typed: base_layer / lora_A / lora_B / scaling / active_adapters) are
    folded into plain Linear weights under the reference's key names; disabled adapters are ignored."""
    from acestep_b200.pack import UnsupportedAdapterError, effective_decoder_state

    class LoraLinear(torch.nn.Module):
        def __init__(self, i, o, r=2):
            super().__init__()
            self.base_layer = torch.nn.Linear(i, o, bias=False)
            self.lora_A = torch.nn.ModuleDict({"a": torch.nn.Linear(i, r, bias=False), "b": torch.nn.Linear(i, r, bias=False)})
            self.lora_B = torch.nn.ModuleDict({"a": torch.nn.Linear(r, o, bias=False), "b": torch.nn.Linear(r, o, bias=False)})
            self.scaling = {"a": 0.5, "b": 2.0}
            self.active_adapters = ["a"]
            self.disable_adapters = False

    class Inner(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj = LoraLinear(4, 3)
            self.norm = torch.nn.LayerNorm(3)

    class Peft(torch.nn.Module):  # PeftModel.base_model.model.<decoder>
        def __init__(self):
            super().__init__()
            self.base_model = torch.nn.Module()
            self.base_model.model = Inner()

    torch.manual_seed(0)
    m = Peft()
    lin = m.base_model.model.q_proj
    sd = effective_decoder_state(m)
    assert set(sd) == {"q_proj.weight", "norm.weight", "norm.bias"}
    want = lin.base_layer.weight + 0.5 * lin.lora_B["a"].weight @ lin.lora_A["a"].weight
    assert torch.allclose(sd["q_proj.weight"], want.detach())
    lin.active_adapters = ["a", "b"]
    want2 = want + 2.0 * lin.lora_B["b"].weight @ lin.lora_A["b"].weight
    assert torch.allclose(effective_decoder_state(m)["q_proj.weight"], want2.detach())
    lin.disable_adapters = True
    assert torch.equal(effective_decoder_state(m)["q_proj.weight"], lin.base_layer.weight.detach())
    plain = torch.nn.Linear(4, 3)
    assert set(effective_decoder_state(plain)) == {"weight", "bias"}
    plain._lycoris_net = object()  # a net without a `loras` list cannot be folded: loud, not silent
    with pytest.raises(UnsupportedAdapterError):
        effective_decoder_state(plain)


def test_effective_decoder_state_folds_lycoris_lokr():
    """LoKr (handler/lora/lifecycle.py:101-156): `decoder._lycoris_net.loras` wrap the decoder's own Linear modules;
    the fold is W + multiplier * scale * kron(w1, w2) with either factor optionally low-rank (w_a @ w_b), the
    module's own get_diff_weight taking precedence; multiplier 0 (set_use_lora(False), controls.py:13-31) folds
    nothing; DoRA / Tucker variants are rejected."""
    from acestep_b200.pack import UnsupportedAdapterError, effective_decoder_state

    class Dec(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj = torch.nn.Linear(6, 4, bias=False)
            self.o_proj = torch.nn.Linear(4, 6, bias=False)

    class Lokr:  # duck-typed LokrModule
        def __init__(self, org, w1, w2=None, w2a=None, w2b=None, scale=1.0, multiplier=1.0):
            self.org_module, self.lora_name = [org], "lycoris_x"
            self.lokr_w1, self.lokr_w2, self.lokr_w2_a, self.lokr_w2_b = w1, w2, w2a, w2b
            self.scale, self.multiplier = scale, multiplier

    class Net:
        def __init__(self, loras):
            self.loras = loras

    g = torch.Generator().manual_seed(3)
    dec = Dec()
    w1, w2 = torch.randn(2, 3, generator=g), torch.randn(2, 2, generator=g)          # kron -> [4, 6]
    v1, v2a, v2b = torch.randn(3, 2, generator=g), torch.randn(2, 1, generator=g), torch.randn(1, 2, generator=g)  # -> [6, 4]
    a, b = Lokr(dec.q_proj, w1, w2=w2, multiplier=0.5), Lokr(dec.o_proj, v1, w2a=v2a, w2b=v2b, scale=0.25, multiplier=2.0)
    dec._lycoris_net = Net([a, b])
    sd = effective_decoder_state(dec)
    assert set(sd) == {"q_proj.weight", "o_proj.weight"}
    assert torch.allclose(sd["q_proj.weight"], dec.q_proj.weight.detach() + 0.5 * torch.kron(w1, w2))
    assert torch.allclose(sd["o_proj.weight"], dec.o_proj.weight.detach() + 2.0 * 0.25 * torch.kron(v1, v2a @ v2b))
    a.multiplier = b.multiplier = 0.0  # adapter switched off through its multiplier
    sd0 = effective_decoder_state(dec)
    assert torch.equal(sd0["q_proj.weight"], dec.q_proj.weight.detach()) and torch.equal(sd0["o_proj.weight"], dec.o_proj.weight.detach())
    a.multiplier = 1.0
    a.get_diff_weight = lambda m: (torch.full((4, 6), 3.0) * m, None)  # the library's own method wins
    assert torch.allclose(effective_decoder_state(dec)["q_proj.weight"], dec.q_proj.weight.detach() + 3.0)
    a.wd = True
    with pytest.raises(UnsupportedAdapterError):
        effective_decoder_state(dec)


def test_lora_mutators_trigger_a_repack():
    class LoraHost(FakeHandler):
        def load_lora(self, path):
            self.ref_calls.append(("load_lora", path))
            return "✅ loaded"

        def set_lora_scale(self, a, b=None):
            self.ref_calls.append(("scale", a, b))
            return "✅ scale"

    h = install(LoraHost())
    repacks = []
    h._init_b200_backends = lambda dit=True, vae=True, cond=True: repacks.append((dit, vae, cond))
    assert h.load_lora("x") == "✅ loaded" and repacks == []  # inactive backend: nothing to repack
    h.use_b200_dit = True
    assert h.load_lora("y") == "✅ loaded" and h.set_lora_scale("a", 0.5) == "✅ scale"
    assert repacks == [(True, False, False)] * 2
    assert h.ref_calls == [("load_lora", "x"), ("load_lora", "y"), ("scale", "a", 0.5)]

    def boom(dit=True, vae=True, cond=True):
        from acestep_b200.pack import UnsupportedAdapterError
        raise UnsupportedAdapterError("LoKr")

    h._init_b200_backends = boom
    msg = h.load_lora("z")
    assert "B200 DiT disabled" in msg and h.use_b200_dit is False  # loud, not silent


def test_adapter_switch_repacks_and_unload_reenables_after_a_failed_repack(monkeypatch):
    """(a) `set_active_lora_adapter` (handler/lora/controls.py:193-206 -> decoder.set_adapter) changes which LoRA
    the decoder applies, so it must repack like the other mutators; (b) a repack that fails for ANY reason never
    raises out of the reference's method (the model is already mutated), switches the backend off loudly, keeps
    the previous engine alive until a new one exists, and the next successful mutation (unload_lora) brings the
    backend back."""
    import acestep_b200.backend as backend

    class LoraHost(FakeHandler):
        def load_lora(self, path):
            return "✅ loaded"

        def unload_lora(self):
            return "✅ unloaded"

        def set_active_lora_adapter(self, name):
            self.ref_calls.append(("active", name))
            return f"✅ active {name}"

    built, closed = [], []

    class FakeDit:
        def __init__(self, state, shape, device):
            if state == "boom":
                raise KeyError("layers.0.self_attn.q_proj.conv.weight")
            built.append(self)

        def close(self):
            closed.append(self)

    h = install(LoraHost())
    h.model.decoder = object()
    h.model.config = object()
    state = {"v": "ok"}
    monkeypatch.setattr(backend, "B200DiT", FakeDit)
    monkeypatch.setattr(backend, "B200Sampler", lambda dit, null=None: ("sampler", dit))
    monkeypatch.setattr(backend, "effective_decoder_state", lambda dec: state["v"])
    monkeypatch.setattr(backend.DiTShape, "from_config", classmethod(lambda cls, cfg: "shape"))
    h._init_b200_backends(dit=True, vae=False, cond=False)
    first = h.b200_dit
    assert h.use_b200_dit and built == [first] and closed == []
    # (a) two adapters loaded, switch the active one: a repack, the old engine closed only after the new one exists
    assert h.set_active_lora_adapter("voice") == "✅ active voice"
    assert len(built) == 2 and closed == [first] and h.b200_dit is built[1] and h.b200_sampler == ("sampler", built[1])
    # (b) a repack that dies on an unexpected key: no exception, backend off, previous engine NOT destroyed
    state["v"] = "boom"
    msg = h.load_lora("conv-adapter")
    assert "B200 DiT disabled" in msg and "KeyError" in msg
    assert h.use_b200_dit is False and h.b200_dit is built[1] and closed == [first]
    # unload restores a plain decoder: the backend comes back although use_b200_dit was False
    state["v"] = "ok"
    assert h.unload_lora() == "✅ unloaded"
    assert h.use_b200_dit is True and len(built) == 3 and h.b200_dit is built[2] and closed == [first, built[1]]


def test_turbo_models_use_the_turbo_sampler():
    h = install(FakeHandler())
    h.config.is_turbo = True
    h.b200_sampler, h.use_b200_dit = StubSampler(), True
    h._execute_service_generate_diffusion(_payload(), {"timesteps": torch.tensor([1.0, 0.5, 0.0])}, 7, "ode", 3.0, 1.0)
    kind, *_, skw = h.b200_sampler.calls[0]
    assert kind == "turbo" and skw["timesteps"] is not None and len(h.model.calls) == 1


def test_plain_base_model_ignores_timesteps_like_the_reference():
    """base/modeling_acestep_v15_base.py:1812 swallows `timesteps` in **kwargs (only sft and turbo declare it):
    the backend decides from the loaded model's generate_audio signature."""
    enc, ctx, src = torch.zeros(1, 2, 8), torch.zeros(1, 4, 128), torch.zeros(1, 4, 64)

    class BaseModel(FakeModel):
        def generate_audio(self, infer_steps=30, **kwargs):
            return {}

    class SftModel(FakeModel):
        def generate_audio(self, infer_steps=30, timesteps=None, **kwargs):
            return {}

    for model, expect in ((BaseModel(), None), (SftModel(), [1.0, 0.5, 0.0])):
        h = install(FakeHandler())
        h.model = model
        h.b200_sampler, h.use_b200_dit = StubSampler(), True
        h._b200_run_diffusion(enc, None, ctx, src, 0, timesteps=[1.0, 0.5, 0.0])
        kind, *_, skw = h.b200_sampler.calls[0]
        assert kind == "base" and skw["timesteps"] == expect


def test_missing_sampler_raises_attribute_error_like_mlx_sibling():
    h = install(FakeHandler())
    with pytest.raises(AttributeError):
        h._b200_run_diffusion(torch.zeros(1, 2, 8), None, torch.zeros(1, 4, 128), torch.zeros(1, 4, 64), 0)


def test_codec_seams():
    h = install(FakeHandler())
    h.b200_vae, h.use_b200_vae = StubVae(), True
    wav = h.tiled_decode(torch.zeros(3, 64, 10))
    assert wav.shape == (3, 2, 19200) and float(wav[0, 0, 0]) == 0.5
    lat = h.tiled_encode(torch.zeros(2, 3840))  # 2-D input is un-batched again, like the reference
    assert lat.shape == (64, 2)
    lat = h.tiled_encode(torch.zeros(4, 2, 3840))
    assert lat.shape == (4, 64, 2)
    assert h.ref_calls == []


def test_no_silent_fallback_when_active():
    h = install(FakeHandler())

    class Boom:
        def decode_normalized(self, lat):
            raise RuntimeError("kernel failure")

    h.b200_vae, h.use_b200_vae = Boom(), True
    with pytest.raises(RuntimeError):
        h.tiled_decode(torch.zeros(1, 64, 4))
    assert h.ref_calls == []  # did NOT quietly route to the PyTorch path


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_graft_onto_real_handler_and_run_reference_plumbing_tests():
    sys.path.insert(0, REF)
    stub = types.ModuleType("vector_quantize_pytorch")
    stub.ResidualFSQ = type("ResidualFSQ", (torch.nn.Module,), {})
    sys.modules.setdefault("vector_quantize_pytorch", stub)
    try:
        from acestep.handler import AceStepHandler
    except Exception as exc:  # optional deps of the reference missing
        pytest.skip(f"reference handler not importable here: {exc}")
    install(AceStepHandler)
    for name in ("_execute_service_generate_diffusion", "tiled_decode", "tiled_encode"):
        assert hasattr(AceStepHandler, "_ref_" + name.lstrip("_"))
    assert AceStepHandler.use_b200_dit is False and AceStepHandler.use_b200_vae is False
    mods = ["diffusion_test", "vae_decode_chunks_test", "vae_decode_mixin_test", "vae_encode_test",
            "service_generate_execute_test", "service_generate_test", "generate_music_decode_test",
            "generate_music_test", "generate_music_execute_test"]
    suite = unittest.TestSuite()
    for m in mods:
        suite.addTests(unittest.defaultTestLoader.loadTestsFromName(f"acestep.core.generation.handler.{m}"))
    result = unittest.TextTestRunner(verbosity=0, stream=open(os.devnull, "w")).run(suite)
    assert result.testsRun >= 30 and result.wasSuccessful(), (result.failures, result.errors)


class _StubDiT:
    """Records the attention-extraction calls the decoder shim makes."""

    shape = types.SimpleNamespace(num_hidden_layers=24, num_attention_heads=16)

    def __init__(self):
        self.calls = []

    def bind(self, bc, T, E):
        self.calls.append(("bind", bc, T, E))

    def set_condition(self, enc):
        self.calls.append(("cond", tuple(enc.shape)))

    def cross_attentions(self, xt, ctx, t, n_layers):
        self.calls.append(("attn", tuple(xt.shape), tuple(ctx.shape), list(t), n_layers))
        bc, T, _ = xt.shape
        return torch.full((n_layers, bc, 16, (T + 1) // 2, 5), 0.2, dtype=torch.bfloat16)


def test_lyric_alignment_callers_use_the_b200_decoder_shim():
    """get_lyric_timestamp / get_lyric_score call model.decoder(..., output_attentions=True,
    custom_layers_config, enable_early_exit=True) and read [2] (handler/lyric_timestamp.py:78-103): with the
    B200 DiT active the decoder is swapped for the shim during the call and restored afterwards — also when
    the call raises — and an inactive backend leaves everything alone."""

    class RealDecoder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.seen = 0

        def forward(self, **kw):
            self.seen += 1
            return "vt", None, ("ref-attn",) if kw.get("output_attentions") else None

    class Model(torch.nn.Module):  # a real nn.Module: assigning a non-module to .decoder would raise
        def __init__(self):
            super().__init__()
            self.decoder = RealDecoder()

    class Host(FakeHandler):
        custom_layers_config = {2: [6], 6: [8]}

        def get_lyric_timestamp(self, xt, enc, ctx, fail=False, plain=False):
            t = torch.tensor([0.125] * xt.shape[0], dtype=torch.bfloat16)
            if plain:
                return self.model.decoder(hidden_states=xt, timestep=t, timestep_r=t, encoder_hidden_states=enc,
                                          context_latents=ctx)
            out = self.model.decoder(hidden_states=xt, timestep=t, timestep_r=t, attention_mask=None,
                                     encoder_hidden_states=enc, use_cache=False, past_key_values=None,
                                     encoder_attention_mask=None, context_latents=ctx, output_attentions=True,
                                     custom_layers_config=self.custom_layers_config, enable_early_exit=True)
            if fail:
                raise RuntimeError("aligner failure")
            return out

    h = Host()
    h.model = Model()
    real = h.model.decoder
    install(h)
    xt, enc, ctx = torch.zeros(2, 9, 64), torch.zeros(2, 5, 8), torch.zeros(2, 9, 128)
    assert h.get_lyric_timestamp(xt, enc, ctx)[2] == ("ref-attn",) and real.seen == 1  # inactive: stock decoder
    h.b200_dit, h.use_b200_dit = _StubDiT(), True
    out = h.get_lyric_timestamp(xt, enc, ctx)
    assert real.seen == 1 and h.model.decoder is real  # not called; restored
    assert out[0] is None and len(out[2]) == 7 and out[2][6].shape == (2, 16, 5, 5)  # layers 0..max(cfg)
    assert out[2][0].dtype == xt.dtype
    assert h.b200_dit.calls == [("bind", 2, 9, 5), ("cond", (2, 5, 8)),
                                ("attn", (2, 9, 64), (2, 9, 128), [0.125, 0.125], 7)]
    with pytest.raises(RuntimeError):
        h.get_lyric_timestamp(xt, enc, ctx, fail=True)
    assert h.model.decoder is real
    # a call without attention extraction made while the shim is in place goes to the real decoder
    assert h.get_lyric_timestamp(xt, enc, ctx, plain=True)[0] == "vt" and real.seen == 2
    # timestep_r != timestep is not representable by the engine: loud, not silent
    from acestep_b200.backend import B200DecoderShim
    shim = B200DecoderShim(h, real)
    with pytest.raises(ValueError):
        shim(hidden_states=xt, timestep=torch.tensor([0.5, 0.5]), timestep_r=torch.tensor([0.0, 0.0]),
             encoder_hidden_states=enc, context_latents=ctx, output_attentions=True)


def test_output_path_has_no_cpu_fallback():
    """The product output path refuses CPU tensors instead of quietly running torch (no CPU fallback anywhere
    on the hot path); argument validation happens before any library call."""
    from acestep_b200 import _lib
    from acestep_b200.output import check_latents, latent_flags, peak_normalize_

    with pytest.raises(_lib.B200Error):
        peak_normalize_(torch.zeros(1, 2, 8))
    with pytest.raises(_lib.B200Error):
        latent_flags(torch.zeros(1, 4, 64, dtype=torch.bfloat16))
    with pytest.raises(_lib.B200Error):
        check_latents(torch.zeros(1, 4, 64, dtype=torch.bfloat16))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_real_get_lyric_timestamp_runs_through_the_decoder_shim():
    """The reference's OWN `LyricTimestampMixin.get_lyric_timestamp` (handler/lyric_timestamp.py:14-147) and its
    DTW aligner (core/scoring/dit_alignment.py), unmodified, on a host grafted with the B200 backend: the decoder
    call is answered by the shim (the stock decoder would raise), the aligner consumes the [layers, heads,
    tokens, frames] stack built from it, and a near-diagonal attention pattern yields monotone timestamps."""
    sys.path.insert(0, REF)
    stub = types.ModuleType("vector_quantize_pytorch")
    stub.ResidualFSQ = type("ResidualFSQ", (torch.nn.Module,), {})
    sys.modules.setdefault("vector_quantize_pytorch", stub)
    from acestep.core.generation.handler.lyric_timestamp import LyricTimestampMixin

    class Tok:
        def encode(self, s, add_special_tokens=False):
            return [1, 2, 3]

        def decode(self, ids, **k):
            return "".join(chr(97 + (i % 26)) for i in ids)

        def convert_ids_to_tokens(self, ids):
            return [chr(97 + (i % 26)) for i in ids]

    class Dec(torch.nn.Module):
        def forward(self, **kw):
            raise AssertionError("the stock decoder must not run while the B200 DiT is active")

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.decoder = Dec()

    class Host(LyricTimestampMixin, FakeHandler):
        def __init__(self):
            FakeHandler.__init__(self)
            self.model = Model()
            self.text_tokenizer = Tok()
            self.custom_layers_config = {2: [6], 3: [10, 11]}

    class DiagonalDiT(_StubDiT):
        def set_condition(self, enc):
            self.E = enc.shape[1]

        def cross_attentions(self, xt, ctx, t, n_layers):
            self.calls.append(("attn", n_layers, list(t)))
            bc, T, _ = xt.shape
            S = (T + 1) // 2
            g = torch.Generator().manual_seed(0)
            p = torch.rand(n_layers, bc, 16, S, self.E, generator=g) + 5 * torch.eye(S, self.E)[None, None, None]
            return (p / p.sum(-1, keepdim=True)).to(torch.bfloat16)

    h = install(Host())
    h.b200_dit, h.use_b200_dit = DiagonalDiT(), True
    T, E = 40, 20
    ids = torch.tensor([[1, 2, 3] + list(range(10, 22)) + [151643] + [0] * 4])
    out = h.get_lyric_timestamp(pred_latent=torch.randn(1, T, 64), encoder_hidden_states=torch.randn(1, E, 32),
                                encoder_attention_mask=torch.ones(1, E), context_latents=torch.randn(1, T, 128),
                                lyric_token_ids=ids, total_duration_seconds=1.6, inference_steps=8)
    assert out["success"] is True and out["error"] is None, out
    assert len(out["token_timestamps"]) == 12 and out["lrc_text"].startswith("[00:00")
    starts = [t.start for t in out["token_timestamps"]]
    assert starts == sorted(starts) and starts[-1] > starts[0]
    # early exit at max(custom_layers_config) + 1 = 4 layers, t = 1 / inference_steps
    assert h.b200_dit.calls[-1] == ("attn", 4, [0.125])
    assert isinstance(h.model.decoder, Dec)  # restored


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_real_decode_step_is_an_identity_after_the_b200_peak_normalisation():
    """The reference's OWN `_decode_generate_music_pred_latents` (handler/generate_music_decode.py:98-201),
    unmodified, over the wrapped `tiled_decode`: the B200 seam returns waveforms that are already peak-normalised
    (`decode_normalized`; stubbed here with the oracle expression the CUDA kernel is bit-exact against), so the
    reference's own normalisation at :191-195 must leave them untouched and the result must equal the reference
    normalising the raw waveforms itself — bit for bit."""
    sys.path.insert(0, REF)
    stub = types.ModuleType("vector_quantize_pytorch")
    stub.ResidualFSQ = type("ResidualFSQ", (torch.nn.Module,), {})
    sys.modules.setdefault("vector_quantize_pytorch", stub)
    from acestep.core.generation.handler.generate_music_decode import GenerateMusicDecodeMixin
    from oracle import output as oout

    raw = {}

    class Engine:
        def decode(self, lat):
            g = torch.Generator().manual_seed(3)
            gains = torch.tensor([0.2, 1.7, 3.1]).view(-1, 1, 1)[: lat.shape[0]]
            raw["w"] = torch.randn(lat.shape[0], 2, lat.shape[2] * 1920, generator=g) * gains
            return raw["w"].clone()

        def decode_normalized(self, lat):
            return oout.peak_normalize(self.decode(lat))[0]

    class Vae(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

        @property
        def dtype(self):
            return torch.float32

    class Host(GenerateMusicDecodeMixin, FakeHandler):
        def __init__(self):
            FakeHandler.__init__(self)
            self.vae, self.use_mlx_vae, self.mlx_vae, self.current_offload_cost = Vae(), False, None, 0.0

        def _empty_cache(self):
            pass

        def _memory_allocated(self):
            return 0

        def _max_memory_allocated(self):
            return 0

    h = install(Host())
    h.b200_vae, h.use_b200_vae = Engine(), True
    wavs, lat_cpu, costs = h._decode_generate_music_pred_latents(torch.randn(3, 5, 64), None, True,
                                                                {"total_time_cost": 1.0})
    want, _ = oout.peak_normalize(raw["w"])
    assert torch.equal(wavs, want) and h.ref_calls == []
    assert wavs.abs().amax(dim=[1, 2]).tolist()[1:] == [1.0, 1.0] and float(wavs.abs().amax(dim=[1, 2])[0]) < 1.0
    assert lat_cpu.shape == (3, 5, 64) and "vae_decode_time_cost" in costs


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_real_reference_audio_callers_reach_the_b200_encoder():
    """the reference's OWN callers of `tiled_encode` — `infer_refer_latent`
    (handler/conditioning_embed.py:18-69, with its per-request encode cache), `_encode_audio_to_latents`
    (handler/batch_prep.py:63-76) and `_prepare_target_latents_and_wavs` (handler/conditioning_target.py:18-107,
    with its same-audio cache and silence shortcut) — run unmodified on a grafted host and reach the B200 VAE
    through the wrapped seam with the shapes / layouts they expect."""
    sys.path.insert(0, REF)
    stub = types.ModuleType("vector_quantize_pytorch")
    stub.ResidualFSQ = type("ResidualFSQ", (torch.nn.Module,), {})
    sys.modules.setdefault("vector_quantize_pytorch", stub)
    from acestep.core.generation.handler.batch_prep import BatchPrepMixin
    from acestep.core.generation.handler.conditioning_embed import ConditioningEmbedMixin
    from acestep.core.generation.handler.conditioning_target import ConditioningTargetMixin

    class CountingVae(StubVae):
        def __init__(self):
            self.encodes = []

        def encode(self, audio, sample=True):
            self.encodes.append(tuple(audio.shape))
            return torch.full((audio.shape[0], 64, audio.shape[2] // 1920), float(len(self.encodes)))

        encode_cached = encode  # the seam's entry point (posterior-moments cache in the real engine)

    class Host(ConditioningEmbedMixin, ConditioningTargetMixin, BatchPrepMixin, FakeHandler):
        def __init__(self):
            FakeHandler.__init__(self)
            self.silence_latent = torch.full((1, 800, 64), -1.0)

        def _ensure_silence_latent_on_device(self):
            pass

        def is_silence(self, audio):
            return bool(torch.all(audio.abs() < 1e-6))

        def _decode_audio_codes_to_latents(self, code_hint):
            return None

    h = install(Host())
    h.b200_vae, h.use_b200_vae = CountingVae(), True
    # reference clips are 30 s = 750 frames (the silence shortcut returns silence_latent[:, :750], :47)
    ref_a = torch.rand(2, 1920 * 750) - 0.5
    ref_b = torch.rand(1, 1920 * 750) - 0.5  # mono: duplicated to stereo by the caller (:33-35)
    # item 0: two references, the first one twice (cache hit on data_ptr); item 1: all-zero audio -> silence latent
    lat, order = h.infer_refer_latent([[ref_a, ref_b, ref_a], [torch.zeros(2, 1920 * 3)]])
    assert h.b200_vae.encodes == [(1, 2, 1920 * 750), (1, 2, 1920 * 750)] and h.ref_calls == []
    assert order.tolist() == [0, 0, 0, 1] and lat.shape == (4, 750, 64)
    assert float(lat[0, 0, 0]) == 1.0 and float(lat[1, 0, 0]) == 2.0 and float(lat[2, 0, 0]) == 1.0
    assert float(lat[3, 0, 0]) == -1.0
    a = torch.rand(2, 1920 * 6) - 0.5
    z = h._encode_audio_to_latents(a)  # [2, N] -> [T, 64] in the handler dtype on the handler device
    assert z.shape == (6, 64) and z.dtype == h.dtype and float(z[0, 0]) == 3.0
    # batch prep: item 1 repeats item 0's audio (cached, no second encode), item 2 is silence
    h.b200_vae.encodes.clear()
    wavs = torch.stack([a, a, torch.zeros_like(a)])
    tw, tl, masks, max_len, sil = h._prepare_target_latents_and_wavs(3, wavs, [None, None, None])
    assert h.b200_vae.encodes == [(1, 2, 1920 * 6)] and h.ref_calls == []
    assert tl.shape == (3, 128, 64) and max_len == 128 and masks.sum(1).tolist() == [6, 6, 6]
    assert float(tl[0, 0, 0]) == float(tl[1, 0, 0]) == 1.0 and float(tl[2, 0, 0]) == -1.0
    assert float(tl[0, 6, 0]) == -1.0  # padded with the silence latent (:84-89)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_real_get_lyric_score_runs_through_the_decoder_shim():
    """The reference's OWN `get_lyric_score` (handler/lyric_score.py:14-160) + `MusicLyricScorer`: one decoder call
    at batch 2 (pure noise at t = 1 and the regressed latent at t = 1 / steps), answered by the shim."""
    sys.path.insert(0, REF)
    stub = types.ModuleType("vector_quantize_pytorch")
    stub.ResidualFSQ = type("ResidualFSQ", (torch.nn.Module,), {})
    sys.modules.setdefault("vector_quantize_pytorch", stub)
    from acestep.core.generation.handler.lyric_score import LyricScoreMixin

    class Tok:
        def encode(self, s, add_special_tokens=False):
            return [1, 2, 3]

        def decode(self, ids, **k):
            return "".join(chr(97 + (i % 26)) for i in ids)

        def convert_ids_to_tokens(self, ids):
            return [chr(97 + (i % 26)) for i in ids]

    class Dec(torch.nn.Module):
        def forward(self, **kw):
            raise AssertionError("the stock decoder must not run while the B200 DiT is active")

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.decoder = Dec()

    class Host(LyricScoreMixin, FakeHandler):
        def __init__(self):
            FakeHandler.__init__(self)
            self.model, self.text_tokenizer = Model(), Tok()
            self.custom_layers_config = {2: [6], 3: [10, 11]}

    class DiagonalDiT(_StubDiT):
        def set_condition(self, enc):
            self.E = enc.shape[1]

        def cross_attentions(self, xt, ctx, t, n_layers):
            self.calls.append(("attn", tuple(xt.shape), list(t), n_layers))
            bc, T, _ = xt.shape
            S = (T + 1) // 2
            g = torch.Generator().manual_seed(0)
            p = torch.rand(n_layers, bc, 16, S, self.E, generator=g) + 5 * torch.eye(S, self.E)[None, None, None]
            return (p / p.sum(-1, keepdim=True)).to(torch.bfloat16)

    h = install(Host())
    h.b200_dit, h.use_b200_dit = DiagonalDiT(), True
    T, E = 40, 20
    ids = torch.tensor([[1, 2, 3] + list(range(10, 22)) + [151643] + [0] * 4])
    out = h.get_lyric_score(pred_latent=torch.randn(1, T, 64), encoder_hidden_states=torch.randn(1, E, 32),
                            encoder_attention_mask=torch.ones(1, E), context_latents=torch.randn(1, T, 128),
                            lyric_token_ids=ids, inference_steps=8)
    assert out["success"] is True and out["error"] is None, out
    assert isinstance(out["lm_score"], float) and isinstance(out["dit_score"], float)
    assert h.b200_dit.calls[-1] == ("attn", (2, T, 64), [1.0, 0.125], 4)
    assert isinstance(h.model.decoder, Dec)


def test_song_pipeline_ordering_and_deferred_errors():
    """SongPipeline is plain host logic: results come back one submit later, in order; a song whose latent guard
    fired raises at ITS turn (from wait()), and the queue keeps serving the songs behind it."""
    from acestep_b200.pipeline import SongPipeline

    class Pending:
        def __init__(self, value, fail=False):
            self.value, self.fail, self.waited = value, fail, False

        def wait(self):
            self.waited = True
            if self.fail:
                raise RuntimeError("Generation produced NaN or Inf latents.")
            return {"audio": self.value}

    q = SongPipeline(depth=1)
    a, b, c = Pending("a"), Pending("b", fail=True), Pending("c")
    assert q.submit(a) is None and not a.waited           # one song stays in flight
    assert q.submit(b)["audio"] == "a" and not b.waited   # a's result arrives when b is submitted
    with pytest.raises(RuntimeError, match="NaN or Inf"):
        q.submit(c)                                       # b's guard fires at b's turn
    assert q.drain()["audio"] == "c" and q.drain() is None
    q0 = SongPipeline(depth=0)                            # depth 0 = blocking
    assert q0.submit(Pending("x"))["audio"] == "x"
    q2 = SongPipeline(depth=2)
    assert q2.submit(Pending(1)) is None and q2.submit(Pending(2)) is None and q2.submit(Pending(3))["audio"] == 1
    assert q2.drain()["audio"] == 3
