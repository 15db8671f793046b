"""B200 backend for the reference's AceStepHandler — a third sibling of "PyTorch" and "MLX".

The reference selects its DiT / VAE backend at three seams:
  * `_execute_service_generate_diffusion`   handler/service_generate_execute.py:107-196
  * `tiled_decode`                           handler/vae_decode.py:16-85
  * `tiled_encode`                           handler/vae_encode.py:15-82
with the MLX backend installed as mixin methods + flags on the handler (`use_mlx_dit/mlx_decoder`,
`use_mlx_vae/mlx_vae`: handler.py:161-168, init_service_setup.py:116-148).  `B200BackendMixin` adds
the same shape of state (`use_b200_dit/b200_dit`, `use_b200_vae/b200_vae`) and `install()` grafts it
onto an AceStepHandler class or instance, wrapping exactly those three sites; everything else of
the handler (request parsing, conditioning, payloads, LoRA, UI plumbing) is reused unchanged.

Unlike the MLX path there is NO silent fallback: when the B200 backend is active and fails, the
exception propagates (the handler's outer try/except turns it into the usual error payload,
generate_music.py:181-190).
"""
from __future__ import annotations

import inspect
import types
from typing import Any, Dict, Optional, Tuple

import torch

from .cond import B200ConditionEncoder, CondShape
from .dit import B200DiT, DiTShape
from .pack import UnsupportedAdapterError, effective_decoder_state
from .sampler import B200Sampler
from .tokenizer import B200AudioTokenizer, TokShape
from .vae import B200Vae, VaeShape


class B200BackendMixin:
    """Methods + flags grafted onto AceStepHandler (mirrors the Mlx*Mixin classes)."""

    use_b200_dit: bool = False
    use_b200_vae: bool = False
    use_b200_cond: bool = False
    b200_cond: Optional[B200ConditionEncoder] = None
    b200_tok: Optional[B200AudioTokenizer] = None
    b200_dit: Optional[B200DiT] = None
    b200_sampler: Optional[B200Sampler] = None
    b200_vae: Optional[B200Vae] = None

    # ------------------------------------------------------------------ init
    def _init_b200_backends(self, dit: bool = True, vae: bool = True, cond: bool = True) -> Tuple[str, str]:
        """Build the engines from the already-loaded PyTorch modules' state_dict()s — the weight
        source the MLX converters use too (models/mlx/dit_convert.py:33-66).  Call again after a
        LoRA load/unload/scale change (lora/lifecycle.py:212-252 mutates model.decoder) to repack."""
        dit_status = vae_status = "Disabled"
        device = torch.device(self.device if str(self.device).startswith("cuda") else "cuda:0")
        if dit:
            decoder = self.model.decoder
            shape = DiTShape.from_config(self.model.config)
            # adapter-aware: active PEFT LoRA / LoKr deltas are folded into the packed weights.  The new engine is
            # built FIRST and swapped in only once it exists, so a failed repack (unsupported adapter, bad keys,
            # out of memory) leaves the previous engine — and the sampler that points at it — intact.
            new_dit = B200DiT(effective_decoder_state(decoder), shape, device)
            old_dit, self.b200_dit = self.b200_dit, new_dit
            self.b200_sampler = B200Sampler(self.b200_dit, getattr(self.model, "null_condition_emb", None))
            if old_dit is not None:
                old_dit.close()
            self.use_b200_dit = True
            self._b200_dit_wanted = True  # survives a temporary switch-off by the LoRA hooks (see _make_repacking)
            dit_status = "Active (B200 tcgen05)"
            if cond and getattr(self.model, "encoder", None) is not None:
                # condition encoder: same layer kernels, weights from model.encoder
                if self.b200_cond is not None:
                    self.b200_cond.close()
                self.b200_cond = B200ConditionEncoder(self.model.encoder.state_dict(),
                                                      CondShape.from_config(self.model.config), device)
                self.use_b200_cond = True
                dit_status = "Active (B200 tcgen05, condition encoder included)"
                # the cover-song branch of prepare_condition: audio tokenizer -> FSQ -> detokenizer
                if getattr(self.model, "tokenizer", None) is not None and getattr(self.model, "detokenizer", None) is not None:
                    if self.b200_tok is not None:
                        self.b200_tok.close()
                    tsd = {"tokenizer." + k: v for k, v in self.model.tokenizer.state_dict().items()}
                    tsd.update({"detokenizer." + k: v for k, v in self.model.detokenizer.state_dict().items()})
                    self.b200_tok = B200AudioTokenizer(tsd, TokShape.from_config(self.model.config), device)
        if vae:
            if self.b200_vae is not None:
                self.b200_vae.close()
            self.b200_vae = B200Vae(self.vae.state_dict(), VaeShape.from_config(self.vae.config), device)
            self.use_b200_vae = True
            vae_status = "Active (B200 tcgen05)"
        return dit_status, vae_status

    def _b200_is_turbo(self) -> bool:
        cfg = getattr(self, "config", None)
        return bool(getattr(cfg, "is_turbo", False))

    def _b200_honours_timesteps(self) -> bool:
        """The reference's own rule: turbo and sft `generate_audio` take a `timesteps` parameter
        (turbo :1795, sft/modeling_acestep_v15_base.py:1866-1868); the plain base model's does not — the
        keyword is swallowed by **kwargs and linspace(infer_steps) + shift is used
        (base/modeling_acestep_v15_base.py:1812, :1862-1866).  Decided from the loaded model's signature."""
        fn = getattr(getattr(self, "model", None), "generate_audio", None)
        if fn is None:
            return True
        try:
            return "timesteps" in inspect.signature(fn).parameters
        except (TypeError, ValueError):
            return True

    # ------------------------------------------------------------------ DiT
    def _b200_run_diffusion(
        self,
        encoder_hidden_states,
        encoder_attention_mask,
        context_latents,
        src_latents,
        seed,
        infer_method: str = "ode",
        shift: float = 3.0,
        timesteps=None,
        audio_cover_strength: float = 1.0,
        encoder_hidden_states_non_cover=None,
        encoder_attention_mask_non_cover=None,
        context_latents_non_cover=None,
        disable_tqdm: bool = False,
        *,
        infer_steps: int = 30,
        diffusion_guidance_sale: float = 7.0,
        cfg_interval_start: float = 0.0,
        cfg_interval_end: float = 1.0,
        use_adg: bool = False,
        cover_noise_strength: float = 0.0,
    ) -> Dict[str, Any]:
        """Same positional signature and return contract as `_mlx_run_diffusion`
        (handler/diffusion.py:18-56), extended keyword-only with the base-model knobs the MLX path
        lacks.  Attention masks are accepted and unused — the DiT drops them (turbo :1381-1382).

        Returns {"target_latents": Tensor[B,T,64] on self.device in self.dtype, "time_costs": {...}}.
        Raises AttributeError / ValueError / TypeError exactly like the MLX sibling."""
        _ = encoder_attention_mask, encoder_attention_mask_non_cover, disable_tqdm
        for required_attr in ("b200_sampler", "device", "dtype"):
            if not hasattr(self, required_attr) or getattr(self, required_attr) is None:
                raise AttributeError(f"B200BackendMixin host is missing required attribute '{required_attr}'")
        s = self.b200_sampler
        if timesteps is not None and not self._b200_honours_timesteps():
            timesteps = None  # plain base model: ignored, exactly like its generate_audio(**kwargs)
        if self._b200_is_turbo():
            out = s.generate_turbo(
                encoder_hidden_states, context_latents, src_latents, seed, infer_method=infer_method, shift=shift,
                timesteps=timesteps, audio_cover_strength=audio_cover_strength,
                cover_noise_strength=cover_noise_strength,
                encoder_hidden_states_non_cover=encoder_hidden_states_non_cover,
                context_latents_non_cover=context_latents_non_cover)
        else:
            out = s.generate_base(
                encoder_hidden_states, context_latents, src_latents, seed, infer_method=infer_method,
                infer_steps=infer_steps, diffusion_guidance_sale=diffusion_guidance_sale, shift=shift,
                timesteps=timesteps, cfg_interval_start=cfg_interval_start, cfg_interval_end=cfg_interval_end,
                use_adg=use_adg, audio_cover_strength=audio_cover_strength,
                cover_noise_strength=cover_noise_strength,
                encoder_hidden_states_non_cover=encoder_hidden_states_non_cover,
                context_latents_non_cover=context_latents_non_cover)
        out["target_latents"] = out["target_latents"].to(device=self.device, dtype=self.dtype)
        return out

    # ------------------------------------------------------------------ conditioning
    def _b200_prepare_condition(self, *, text_hidden_states, text_attention_mask, lyric_hidden_states,
                                lyric_attention_mask, refer_audio_acoustic_hidden_states_packed,
                                refer_audio_order_mask, hidden_states, attention_mask, silence_latent, src_latents,
                                chunk_masks, is_covers, precomputed_lm_hints_25Hz=None, audio_codes=None):
        """`model.prepare_condition` (modeling_acestep_v15_turbo.py:1604-1649) with the condition encoder
        (`self.encoder(...)`, :1621-1628) and the LM-hint branch (audio tokenizer -> residual FSQ -> detokenizer,
        :1630-1646) on the B200 kernels; same keywords, same 3-tuple.  The LM-hint branch is only consumed where
        is_covers > 0, so it is skipped entirely for plain text2music / repaint batches, whose hints the reference
        computes and then discards (:1646).  `audio_codes` (indices -> codes without the tokenizer) stays on the
        stock quantizer: the reference's own line for it (:1638, `self.tokenize.quantizer`) raises AttributeError."""
        dtype = hidden_states.dtype
        enc, enc_mask = self.b200_cond(
            text_hidden_states=text_hidden_states, text_attention_mask=text_attention_mask,
            lyric_hidden_states=lyric_hidden_states, lyric_attention_mask=lyric_attention_mask,
            refer_audio_acoustic_hidden_states_packed=refer_audio_acoustic_hidden_states_packed,
            refer_audio_order_mask=refer_audio_order_mask)
        if bool((is_covers > 0).any()):
            if precomputed_lm_hints_25Hz is not None:
                hints = precomputed_lm_hints_25Hz[:, : src_latents.shape[1], :]
            else:
                if audio_codes is not None:
                    hints5 = self.model.tokenizer.quantizer.get_output_from_indices(audio_codes)
                    hints = self.model.detokenize(hints5)[:, : src_latents.shape[1], :]
                elif getattr(self, "b200_tok", None) is not None:
                    hints = self.b200_tok.lm_hints(hidden_states, silence_latent, attention_mask,
                                                   src_latents.shape[1]).to(src_latents.dtype)
                else:
                    hints5, _, _ = self.model.tokenize(hidden_states, silence_latent, attention_mask)
                    hints = self.model.detokenize(hints5)[:, : src_latents.shape[1], :]
            src_latents = torch.where(is_covers.unsqueeze(-1).unsqueeze(-1) > 0, hints, src_latents)
        context_latents = torch.cat([src_latents, chunk_masks.to(dtype)], dim=-1)
        return enc.to(dtype), enc_mask, context_latents

    # ------------------------------------------------------------------ codec
    def _b200_vae_decode(self, latents_torch: torch.Tensor) -> torch.Tensor:
        """latents [B, 64, T] -> audio [B, 2, T*1920] (fp32, on the device); cf. _mlx_vae_decode
        (handler/mlx_vae_decode_native.py:31-76).  One pass per sample, no overlap-discard waste."""
        if self.b200_vae is None:
            raise RuntimeError("B200 VAE decode requested but b200_vae is not initialized.")
        # tiled_decode's only caller peak-normalises next (generate_music_decode.py:191-195: |x| temporary,
        # amax, host sync on `any`, divide).  Doing it here in one read pass makes that step an identity —
        # every peak it then sees is <= 1 — without touching the caller.
        return self.b200_vae.decode_normalized(latents_torch)

    def _b200_vae_encode_sample(self, audio_torch: torch.Tensor) -> torch.Tensor:
        """audio [B, 2, N] -> sampled latents [B, 64, N // 1920]; cf. _mlx_vae_encode_sample."""
        if self.b200_vae is None:
            raise RuntimeError("B200 VAE encode requested but b200_vae is not initialized.")
        # Every caller of tiled_encode (infer_refer_latent / _encode_audio_to_latents / _prepare_target_latents_and_wavs,
        # handler/conditioning_embed.py:18-69, batch_prep.py:63-76, conditioning_target.py:18-107) passes
        # offload_latent_to_cpu=True and moves the result straight back to the device; here the latents never
        # leave it, and the posterior moments of a clip seen before (the same reference audio across requests)
        # come from a device-resident cache: only the posterior noise is drawn again.
        return self.b200_vae.encode_cached(audio_torch, sample=True)


# ---------------------------------------------------------------------------------------------
# The three wrapped selection sites
# ---------------------------------------------------------------------------------------------
def _execute_service_generate_diffusion(self, payload, generate_kwargs, seed_param, infer_method, shift,
                                        audio_cover_strength):
    """B200 branch of handler/service_generate_execute.py:107-196 (same return 4-tuple)."""
    if not (getattr(self, "use_b200_dit", False) and self.b200_sampler is not None):
        return self._ref_execute_service_generate_diffusion(payload, generate_kwargs, seed_param, infer_method,
                                                            shift, audio_cover_strength)
    src = payload["src_latents"]
    ones = lambda x: torch.ones(x.shape[0], x.shape[1], device=x.device, dtype=x.dtype)
    with torch.inference_mode():
        with self._load_model_context("model"):
            cond = dict(
                lyric_hidden_states=payload["lyric_hidden_states"], lyric_attention_mask=payload["lyric_attention_mask"],
                refer_audio_acoustic_hidden_states_packed=payload["refer_audio_acoustic_hidden_states_packed"],
                refer_audio_order_mask=payload["refer_audio_order_mask"], silence_latent=self.silence_latent,
                chunk_masks=payload["chunk_mask"])
            use_cond = getattr(self, "use_b200_cond", False) and self.b200_cond is not None
            prepare = self._b200_prepare_condition if use_cond else self.model.prepare_condition
            enc, enc_mask, ctx = prepare(
                text_hidden_states=payload["text_hidden_states"], text_attention_mask=payload["text_attention_mask"],
                hidden_states=src, attention_mask=ones(src), src_latents=src, is_covers=payload["is_covers"],
                precomputed_lm_hints_25Hz=payload["precomputed_lm_hints_25Hz"], **cond)
            enc_nc = mask_nc = ctx_nc = None
            if audio_cover_strength < 1.0 and payload["non_cover_text_hidden_states"] is not None:
                sil = self.silence_latent[:, : src.shape[1], :].expand(src.shape[0], -1, -1)
                enc_nc, mask_nc, ctx_nc = prepare(
                    text_hidden_states=payload["non_cover_text_hidden_states"],
                    text_attention_mask=payload["non_cover_text_attention_masks"], hidden_states=sil,
                    attention_mask=ones(sil), src_latents=sil, is_covers=torch.zeros_like(payload["is_covers"]), **cond)
            outputs = self._b200_run_diffusion(
                enc, enc_mask, ctx, src, seed_param, infer_method=infer_method, shift=shift,
                timesteps=generate_kwargs.get("timesteps"), audio_cover_strength=audio_cover_strength,
                encoder_hidden_states_non_cover=enc_nc, encoder_attention_mask_non_cover=mask_nc,
                context_latents_non_cover=ctx_nc,
                infer_steps=generate_kwargs.get("infer_steps", 30),
                diffusion_guidance_sale=generate_kwargs.get("diffusion_guidance_sale", 7.0),
                cfg_interval_start=generate_kwargs.get("cfg_interval_start", 0.0),
                cfg_interval_end=generate_kwargs.get("cfg_interval_end", 1.0),
                use_adg=generate_kwargs.get("use_adg", False),
                cover_noise_strength=generate_kwargs.get("cover_noise_strength", 0.0))
    return outputs, enc, enc_mask, ctx


def tiled_decode(self, latents, chunk_size: Optional[int] = None, overlap: int = 64,
                 offload_wav_to_cpu: Optional[bool] = None):
    """B200 fast path of handler/vae_decode.py:16-48 (slot of the MLX fast path)."""
    if getattr(self, "use_b200_vae", False) and self.b200_vae is not None:
        return self._b200_vae_decode(latents)
    return self._ref_tiled_decode(latents, chunk_size, overlap, offload_wav_to_cpu)


def tiled_encode(self, audio, chunk_size=None, overlap=None, offload_latent_to_cpu=True):
    """B200 fast path of handler/vae_encode.py:15-43."""
    if getattr(self, "use_b200_vae", False) and self.b200_vae is not None:
        input_was_2d = audio.dim() == 2
        if input_was_2d:
            audio = audio.unsqueeze(0)
        result = self._b200_vae_encode_sample(audio)
        return result.squeeze(0) if input_was_2d else result
    return self._ref_tiled_encode(audio, chunk_size, overlap, offload_latent_to_cpu)


# Lyric alignment: `get_lyric_timestamp` / `get_lyric_score` call `self.model.decoder(...,
# output_attentions=True, custom_layers_config=cfg, enable_early_exit=True)` directly and read only element [2],
# the per-layer cross-attention probabilities (handler/lyric_timestamp.py:78-103, lyric_score.py:94-122).  While
# the B200 DiT is active those two methods run with `model.decoder` swapped for this shim.
class B200DecoderShim(torch.nn.Module):
    """Callable with the reference decoder's keywords (turbo :1300-1317).  Attention-extraction calls go to
    `B200DiT.cross_attentions`; anything else is forwarded to the real decoder."""

    def __init__(self, host, real_decoder):
        super().__init__()
        object.__setattr__(self, "_host", host)          # plain references: neither is a child module
        object.__setattr__(self, "_real", real_decoder)

    def forward(self, hidden_states=None, timestep=None, timestep_r=None, attention_mask=None,
                encoder_hidden_states=None, encoder_attention_mask=None, context_latents=None,
                output_attentions=False, custom_layers_config=None, enable_early_exit=False, **kwargs):
        wants_attn = bool(output_attentions) or (custom_layers_config is not None and enable_early_exit)
        if not wants_attn:
            return self._real(hidden_states=hidden_states, timestep=timestep, timestep_r=timestep_r,
                              attention_mask=attention_mask, encoder_hidden_states=encoder_hidden_states,
                              encoder_attention_mask=encoder_attention_mask, context_latents=context_latents,
                              output_attentions=output_attentions, custom_layers_config=custom_layers_config,
                              enable_early_exit=enable_early_exit, **kwargs)
        dit = self._host.b200_dit
        if dit is None:
            raise RuntimeError("B200 DiT requested for attention extraction but b200_dit is not initialized.")
        if timestep_r is not None and not torch.equal(timestep_r, timestep):
            # the engine precomputes time_embed_r(t - r) for t == r (every inference caller); be loud otherwise
            raise ValueError("B200 decoder shim supports timestep_r == timestep only")
        n_total = dit.shape.num_hidden_layers
        n_layers = n_total
        if custom_layers_config is not None and enable_early_exit:
            n_layers = min(n_total, max(int(k) for k in custom_layers_config.keys()) + 1)
        bc, T, _ = hidden_states.shape
        dit.bind(bc, T, encoder_hidden_states.shape[1])
        dit.set_condition(encoder_hidden_states)
        probs = dit.cross_attentions(hidden_states, context_latents, timestep.float().tolist(), n_layers)
        probs = probs.to(hidden_states.dtype)
        # (hidden_states, past_key_values, all_cross_attentions): the alignment callers read [2] only, and
        # index it by absolute layer number, so layers [0, n_layers) are all present
        return None, None, tuple(probs[i] for i in range(n_layers))


def _make_attention_caller(name):
    def wrapper(self, *args, **kwargs):
        ref = getattr(self, "_ref_" + name)
        model = getattr(self, "model", None)
        if not (getattr(self, "use_b200_dit", False) and self.b200_dit is not None and model is not None):
            return ref(*args, **kwargs)
        real = model.decoder
        model.decoder = B200DecoderShim(self, real)
        try:
            return ref(*args, **kwargs)
        finally:
            model.decoder = real

    wrapper.__name__ = name
    wrapper.__doc__ = f"{name} of the reference handler with the decoder's attention extraction on the B200 DiT."
    return wrapper


_ATTENTION_CALLERS = ("get_lyric_timestamp", "get_lyric_score")


# Weight-mutation hooks.  Every LoRA lifecycle / control entry point of the handler
# (handler/lora/lifecycle.py:164-440 add_lora / load_lora / add_voice_lora / remove_lora / unload_lora,
# handler/lora/controls.py:35-206 set_use_lora / set_lora_scale / set_active_lora_adapter — the last one calls
# model.decoder.set_adapter(name), i.e. changes WHICH adapter the decoder applies) changes what model.decoder
# computes; a backend that owns PACKED weights must repack or it silently keeps generating with the old ones.
_LORA_MUTATORS = ("add_lora", "load_lora", "add_voice_lora", "remove_lora", "unload_lora", "set_use_lora",
                  "set_lora_scale", "set_active_lora_adapter")


def _make_repacking(name):
    def wrapper(self, *args, **kwargs):
        status = getattr(self, "_ref_" + name)(*args, **kwargs)
        # repack whenever the backend was installed AND enabled, also while a previous failure has it switched
        # off: unload_lora / remove_lora after an unsupported adapter must bring it back
        wanted = getattr(self, "_b200_dit_wanted", False) or getattr(self, "use_b200_dit", False)
        if wanted and getattr(self, "model", None) is not None:
            try:
                self._init_b200_backends(dit=True, vae=False, cond=False)
            except Exception as exc:  # noqa: BLE001 — the reference already mutated the model: never raise here
                # not silent: the B200 DiT is switched off, the stock path takes over with the mutated PyTorch
                # decoder, and the status string the UI shows says so
                self.use_b200_dit = False
                self._b200_dit_wanted = True
                why = str(exc) if isinstance(exc, UnsupportedAdapterError) else f"{type(exc).__name__}: {exc}"
                note = f" | ⚠️ B200 DiT disabled ({why}); using the PyTorch decoder"
                status = status + note if isinstance(status, str) else status
        return status

    wrapper.__name__ = name
    wrapper.__doc__ = f"{name} of the reference handler, followed by a repack of the B200 DiT weights."
    return wrapper


_WRAPPED = {
    "_execute_service_generate_diffusion": _execute_service_generate_diffusion,
    "tiled_decode": tiled_decode,
    "tiled_encode": tiled_encode,
}
_MIXIN_ATTRS = ("_init_b200_backends", "_b200_is_turbo", "_b200_honours_timesteps", "_b200_run_diffusion", "_b200_prepare_condition",
                "_b200_vae_decode", "_b200_vae_encode_sample")


def install(target):
    """Graft the B200 backend onto an AceStepHandler CLASS (affects all instances) or INSTANCE.

    The original selection-site methods stay reachable as `_ref_<name>` and are used whenever the
    B200 flags are off, so an installed-but-inactive backend is behaviour-neutral (the reference's
    own plumbing tests keep passing against a grafted handler)."""
    is_class = isinstance(target, type)
    cls = target if is_class else type(target)
    if getattr(target, "_b200_installed", False):
        return target
    bind = (lambda f: f) if is_class else (lambda f: types.MethodType(f, target))
    for name in _MIXIN_ATTRS:
        setattr(target, name, bind(B200BackendMixin.__dict__[name]))
    for name, fn in _WRAPPED.items():
        original = getattr(cls, name, None)
        if original is None:
            raise AttributeError(f"{cls.__name__} has no method '{name}' to wrap")
        setattr(target, "_ref_" + name.lstrip("_"), bind(original))
        setattr(target, name, bind(fn))
    for name in _LORA_MUTATORS:  # present on the real handler; optional on minimal hosts
        original = getattr(cls, name, None)
        if original is None:
            continue
        setattr(target, "_ref_" + name, bind(original))
        setattr(target, name, bind(_make_repacking(name)))
    for name in _ATTENTION_CALLERS:  # present on the real handler; optional on minimal hosts
        original = getattr(cls, name, None)
        if original is None:
            continue
        setattr(target, "_ref_" + name, bind(original))
        setattr(target, name, bind(_make_attention_caller(name)))
    for flag, val in (("use_b200_dit", False), ("use_b200_vae", False), ("use_b200_cond", False), ("b200_dit", None),
                      ("b200_tok", None),
                      ("_b200_dit_wanted", False),
                      ("b200_sampler", None), ("b200_vae", None), ("b200_cond", None)):
        if not hasattr(target, flag):
            setattr(target, flag, val)
    setattr(target, "_b200_installed", True)
    return target
