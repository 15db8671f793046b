"""Weight packers: reference `state_dict()` tensors -> the blobs libacestep_b200 consumes.

Same role as the reference's MLX converters (acestep/models/mlx/dit_convert.py:11-66,
vae_convert.py:18-132): walk the PyTorch module's state dict, fold weight-norm, pre-concatenate
fused projections and re-lay convolution kernels for the target backend.  The element order here
must match the walkers in csrc/dit.cu (ace_dit_create) and csrc/vae.cu (walk_blob).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch


def _get(sd: Dict[str, torch.Tensor], key: str) -> torch.Tensor:
    if key not in sd:
        raise KeyError(f"state_dict is missing '{key}'")
    return sd[key].detach()


# --------------------------------------------------------------------------------------------
# Adapter-aware weight source (LoRA load / unload / scale mutates model.decoder)
# --------------------------------------------------------------------------------------------
class UnsupportedAdapterError(RuntimeError):
    """model.decoder carries a parametrisation this packer cannot fold into plain Linear weights."""


def _lokr_delta(mod) -> torch.Tensor:
    """Weight delta of one LyCORIS LoKr module — what its (non-bypass) forward adds to the wrapped Linear's weight:
    multiplier * scale * kron(w1, w2), with w1 = lokr_w1 or lokr_w1_a @ lokr_w1_b and w2 = lokr_w2 or
    lokr_w2_a @ lokr_w2_b (third-party `lycoris-lora`, unpinned in the reference's requirements and absent here;
    this restates its published LokrModule.get_diff_weight for Linear layers).  The module's own
    `get_diff_weight` is used when it exists, so a newer library version stays authoritative."""
    mult = float(getattr(mod, "multiplier", 1.0))
    if bool(getattr(mod, "wd", False)) or getattr(mod, "dora_scale", None) is not None:
        raise UnsupportedAdapterError("LoKr module with weight decomposition (DoRA) is not a linear fold")
    if getattr(mod, "use_tucker", False) or getattr(mod, "tucker", False) or getattr(mod, "lokr_t2", None) is not None:
        raise UnsupportedAdapterError("LoKr module with Tucker decomposition (conv layers) is not supported")
    fn = getattr(mod, "get_diff_weight", None)
    if callable(fn):
        out = fn(mult)
        d = out[0] if isinstance(out, (tuple, list)) else out
        return d.detach().float()

    def factor(name):
        full = getattr(mod, name, None)
        if full is not None:
            return full.detach().float()
        a, b_ = getattr(mod, name + "_a", None), getattr(mod, name + "_b", None)
        if a is None or b_ is None:
            raise UnsupportedAdapterError(f"LoKr module has neither {name} nor {name}_a / {name}_b")
        return a.detach().float() @ b_.detach().float()

    scale = float(getattr(mod, "scale", 1.0))
    return torch.kron(factor("lokr_w1"), factor("lokr_w2")) * (scale * mult)


def _lycoris_deltas(decoder) -> Dict[str, torch.Tensor]:
    """{decoder module name: weight delta} for the LyCORIS net `_load_lokr_adapter` attaches as
    `decoder._lycoris_net` (handler/lora/lifecycle.py:101-156): `net.loras` wrap the decoder's own Linear modules
    (`org_module[0]`), whose parameters — and state_dict keys — stay untouched; scale / enable go through each
    module's `multiplier` (controls.py:13-31, :121-128)."""
    net = getattr(decoder, "_lycoris_net", None)
    if net is None:
        return {}
    loras = getattr(net, "loras", None)
    if loras is None:
        raise UnsupportedAdapterError("decoder._lycoris_net has no `loras` list to fold")
    by_id = {id(m): n for n, m in decoder.named_modules()}
    deltas: Dict[str, torch.Tensor] = {}
    for lo in loras:
        org = getattr(lo, "org_module", None)
        org = org[0] if isinstance(org, (list, tuple)) and org else org
        name = by_id.get(id(org))
        if name is None:
            raise UnsupportedAdapterError(f"LyCORIS module {getattr(lo, 'lora_name', '?')} wraps a module outside the decoder")
        if not isinstance(org, torch.nn.Linear):
            raise UnsupportedAdapterError(f"LyCORIS module on non-Linear layer {name}")
        d = _lokr_delta(lo)
        if tuple(d.shape) != tuple(org.weight.shape):
            d = d.reshape(org.weight.shape)
        deltas[name] = deltas[name] + d if name in deltas else d
    return deltas


def effective_decoder_state(decoder) -> Dict[str, torch.Tensor]:
    """`state_dict()` of the PLAIN decoder that computes what `decoder` computes right now.

    The reference wraps `model.decoder` in a PEFT `PeftModel` when a LoRA is loaded
    (handler/lora/lifecycle.py:164-287): Linear layers become LoRA layers with `base_layer`,
    `lora_A[name]`, `lora_B[name]`, `scaling[name]`, `active_adapters`, `disable_adapters`, and the state
    dict keys gain a `base_model.model.` prefix and `.base_layer` infixes.  This folds every ACTIVE
    adapter into its base weight, W + sum_a scaling[a] * B_a @ A_a (what PEFT's `get_delta_weight` adds
    on merge), and restores the reference key names so pack_dit can walk it.  A decoder without
    adapters is returned as its own state dict.  LoKr / LyCORIS nets (`_lycoris_net`, lifecycle.py:101-156) are
    folded the same way, W + multiplier * scale * kron(w1, w2) (see _lokr_delta); DoRA-style weight decomposition
    and Tucker (conv) factors are not linear folds: UnsupportedAdapterError."""
    lyco = _lycoris_deltas(decoder)
    deltas: Dict[str, torch.Tensor] = {}
    for name, mod in decoder.named_modules():
        if not (hasattr(mod, "base_layer") and hasattr(mod, "lora_A") and hasattr(mod, "lora_B")):
            continue
        if getattr(mod, "merged", False):
            continue  # already folded into base_layer.weight by PEFT
        if getattr(mod, "disable_adapters", False):
            continue
        active = getattr(mod, "active_adapters", None)
        if active is None:
            active = list(mod.lora_A.keys())
        if isinstance(active, str):
            active = [active]
        delta = None
        for a in active:
            if a not in mod.lora_A or a not in mod.lora_B:
                continue
            wa, wb = mod.lora_A[a].weight.detach().float(), mod.lora_B[a].weight.detach().float()
            d = float(mod.scaling[a]) * (wb @ wa)
            delta = d if delta is None else delta + d
        if delta is not None:
            deltas[name] = delta
    out: Dict[str, torch.Tensor] = {}
    for k, v in decoder.state_dict().items():
        if ".lora_A." in k or ".lora_B." in k or ".lora_embedding_" in k or ".lora_magnitude_vector" in k:
            continue
        key = k
        owner = None
        if ".base_layer." in key:
            owner = key.split(".base_layer.")[0]
            key = key.replace(".base_layer.", ".")
        t = v.detach()
        if owner is not None and key.endswith(".weight") and owner in deltas:
            t = (t.float() + deltas[owner].to(t.device)).to(t.dtype)
        if key.endswith(".weight") and k[: -len(".weight")] in lyco:
            t = (t.float() + lyco[k[: -len(".weight")]].to(t.device)).to(t.dtype)
        for pre in ("base_model.model.", "base_model."):
            if key.startswith(pre):
                key = key[len(pre):]
                break
        out[key] = t
    return out


# --------------------------------------------------------------------------------------------
# DiT
# --------------------------------------------------------------------------------------------
def pack_dit(sd: Dict[str, torch.Tensor], num_layers: int, prefix: str = "") -> torch.Tensor:
    """Returns a 1-D bf16 CPU tensor in the order ace_dit_create walks.

    `sd` is `model.decoder.state_dict()` of the reference AceStepDiTModel
    (modeling_acestep_v15_turbo.py:1237-1298); `prefix` e.g. "decoder." for the full model dict.
    """
    g = lambda k: _get(sd, prefix + k)
    parts: List[torch.Tensor] = []
    w_in = g("proj_in.1.weight")  # [D, 192, 2] -> [D, (k, c)]
    parts += [w_in.permute(0, 2, 1).reshape(w_in.shape[0], -1), g("proj_in.1.bias")]
    for te in ("time_embed.", "time_embed_r."):
        parts += [g(te + "linear_1.weight"), g(te + "linear_1.bias"), g(te + "linear_2.weight"),
                  g(te + "linear_2.bias"), g(te + "time_proj.weight"), g(te + "time_proj.bias")]
    parts += [g("condition_embedder.weight"), g("condition_embedder.bias"), g("norm_out.weight"),
              g("scale_shift_table").reshape(2, -1)]
    w_out = g("proj_out.1.weight")  # [D, 64, 2] -> [(k, o), D]
    parts += [w_out.permute(2, 1, 0).reshape(-1, w_out.shape[0]), g("proj_out.1.bias")]
    parts += [torch.stack([g(f"layers.{l}.scale_shift_table").reshape(6, -1) for l in range(num_layers)])]
    for l in range(num_layers):
        p = f"layers.{l}."
        gate, up = g(p + "mlp.gate_proj.weight"), g(p + "mlp.up_proj.weight")
        inter, d = gate.shape
        gate_up = torch.stack([gate.view(inter // 64, 64, d), up.view(inter // 64, 64, d)], dim=1).reshape(2 * inter, d)
        parts += [
            g(p + "self_attn_norm.weight"), g(p + "cross_attn_norm.weight"), g(p + "mlp_norm.weight"),
            torch.cat([g(p + "self_attn.q_proj.weight"), g(p + "self_attn.k_proj.weight"),
                       g(p + "self_attn.v_proj.weight")], dim=0),
            g(p + "self_attn.q_norm.weight"), g(p + "self_attn.k_norm.weight"), g(p + "self_attn.o_proj.weight"),
            # the cross-attention RMSNorm has no AdaLN modulation, so its weight vector is folded into the q
            # projection (x_hat * w) @ Wq^T = x_hat @ (Wq * w)^T: the q GEMM then reads the residual stream itself
            # and only applies rstd in its epilogue (csrc/epilogues.cuh NormIn, csrc/dit.cu)
            g(p + "cross_attn.q_proj.weight").float() * g(p + "cross_attn_norm.weight").float()[None, :],
            torch.cat([g(p + "cross_attn.k_proj.weight"), g(p + "cross_attn.v_proj.weight")], dim=0),
            g(p + "cross_attn.q_norm.weight"), g(p + "cross_attn.k_norm.weight"), g(p + "cross_attn.o_proj.weight"),
            gate_up, g(p + "mlp.down_proj.weight"),
        ]
    return torch.cat([t.reshape(-1).to(torch.bfloat16) for t in parts]).contiguous()


# --------------------------------------------------------------------------------------------
# Condition encoders
# --------------------------------------------------------------------------------------------
def pack_encoder(sd: Dict[str, torch.Tensor], num_layers: int, prefix: str, embed: bool = True) -> torch.Tensor:
    """One encoder stack (`prefix` = "lyric_encoder." / "timbre_encoder." inside
    AceStepConditionEncoder.state_dict(), modeling_acestep_v15_turbo.py:574-598, 994-1018) as the 1-D
    bf16 blob ace_enc_create walks: embed_tokens, final norm, then per layer the two norms, fused qkv,
    q/k norms, o_proj, gate/up interleaved in 64-row blocks (like pack_dit), down_proj."""
    g = lambda k: _get(sd, prefix + k)
    # embed=False (AceEncConfig.in_dim = 0): the caller feeds already-embedded tokens (audio tokenizer pooler /
    # detokenizer, whose special tokens join after embed_tokens) and applies embed_tokens itself
    parts: List[torch.Tensor] = ([g("embed_tokens.weight"), g("embed_tokens.bias")] if embed else []) + [g("norm.weight")]
    for l in range(num_layers):
        p = f"layers.{l}."
        gate, up = g(p + "mlp.gate_proj.weight"), g(p + "mlp.up_proj.weight")
        inter, d = gate.shape
        gate_up = torch.stack([gate.view(inter // 64, 64, d), up.view(inter // 64, 64, d)], dim=1).reshape(2 * inter, d)
        parts += [
            g(p + "input_layernorm.weight"), g(p + "post_attention_layernorm.weight"),
            torch.cat([g(p + "self_attn.q_proj.weight"), g(p + "self_attn.k_proj.weight"),
                       g(p + "self_attn.v_proj.weight")], dim=0),
            g(p + "self_attn.q_norm.weight"), g(p + "self_attn.k_norm.weight"), g(p + "self_attn.o_proj.weight"),
            gate_up, g(p + "mlp.down_proj.weight"),
        ]
    return torch.cat([t.reshape(-1).to(torch.bfloat16) for t in parts]).contiguous()


# --------------------------------------------------------------------------------------------
# VAE
# --------------------------------------------------------------------------------------------
def fold_weight_norm(sd: Dict[str, torch.Tensor], name: str) -> torch.Tensor:
    """w = g * v / ||v|| over all dims but 0 (torch weight_norm; mlx/vae_convert.py:18-34)."""
    if name + ".weight" in sd:
        return _get(sd, name + ".weight").float()
    gsc, v = _get(sd, name + ".weight_g").float(), _get(sd, name + ".weight_v").float()
    nrm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return gsc * v / nrm


def folded_vae_state(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """State dict with weight-norm folded and every tensor rounded through bf16 (fp32 storage):
    exactly the parameter values the CUDA path computes with (used by the parity tests)."""
    out: Dict[str, torch.Tensor] = {}
    for k in sd:
        if k.endswith(".weight_g"):
            base = k[: -len(".weight_g")]
            out[base + ".weight"] = fold_weight_norm(sd, base).to(torch.bfloat16).float()
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = _get(sd, k).to(torch.bfloat16).float()
    return out


class _Blob:
    def __init__(self):
        self.chunks: List[bytes] = []
        self.off = 0

    def _align(self):
        pad = (-self.off) % 256
        if pad:
            self.chunks.append(b"\0" * pad)
            self.off += pad

    def bf16(self, t: torch.Tensor):
        self._align()
        b = t.detach().cpu().contiguous().to(torch.bfloat16).view(torch.int16).numpy().tobytes()
        self.chunks.append(b)
        self.off += len(b)

    def f32(self, t: torch.Tensor):
        self._align()
        b = t.detach().cpu().contiguous().to(torch.bfloat16).float().numpy().tobytes()  # bf16-representable values
        self.chunks.append(b)
        self.off += len(b)

    def raw_f32(self, t: torch.Tensor):
        self._align()
        b = t.detach().cpu().contiguous().float().numpy().tobytes()
        self.chunks.append(b)
        self.off += len(b)

    def finish(self) -> torch.Tensor:
        self._align()
        import numpy as np

        return torch.from_numpy(np.frombuffer(b"".join(self.chunks), dtype=np.uint8).copy())


def pack_vae(sd: Dict[str, torch.Tensor], ratios: Sequence[int], channel_multiples: Sequence[int],
             encoder_hidden: int = 128, decoder_channels: int = 128) -> torch.Tensor:
    """Returns a uint8 CPU tensor laid out as csrc/vae.cu:walk_blob expects.

    `sd` is `AutoencoderOobleck.state_dict()` (keys encoder./decoder.*, weight_g/weight_v pairs).
    """
    b = _Blob()
    n = len(ratios)

    def conv(name, bias=True):  # Conv1d [Cout, Cin, K] -> tap-major [Cout, K*Cin]
        w = fold_weight_norm(sd, name)
        b.bf16(w.permute(0, 2, 1).reshape(w.shape[0], -1))
        if bias:
            b.f32(_get(sd, name + ".bias"))

    def conv_t(name, s):  # ConvTranspose1d [Cin, Cout, 2s] -> [(p, co), (tap, ci)], k = tap*s + p
        w = fold_weight_norm(sd, name)
        cin, cout, k = w.shape
        assert k == 2 * s
        b.bf16(w.view(cin, cout, 2, s).permute(3, 1, 2, 0).reshape(s * cout, 2 * cin))
        b.f32(_get(sd, name + ".bias"))

    def snake(name):
        alpha = _get(sd, name + ".alpha").to(torch.bfloat16).float().reshape(-1)
        beta = _get(sd, name + ".beta").to(torch.bfloat16).float().reshape(-1)
        b.raw_f32(torch.exp(alpha))
        b.raw_f32(1.0 / (torch.exp(beta) + 1e-9))

    def res_unit(name):
        snake(name + ".snake1")
        conv(name + ".conv1")
        snake(name + ".snake2")
        conv(name + ".conv2")

    # decoder (upsampling ratios = reversed downsampling ratios)
    conv("decoder.conv1")
    for i in range(n):
        s = ratios[n - 1 - i]
        blk = f"decoder.block.{i}"
        snake(blk + ".snake1")
        conv_t(blk + ".conv_t1", s)
        for j in (1, 2, 3):
            res_unit(f"{blk}.res_unit{j}")
    snake("decoder.snake1")
    w2 = fold_weight_norm(sd, "decoder.conv2")  # [2, C, 7] -> [2][7][C]
    b.f32(w2.permute(0, 2, 1))
    # encoder
    b.f32(fold_weight_norm(sd, "encoder.conv1"))  # [C][2][7]
    b.f32(_get(sd, "encoder.conv1.bias"))
    for i in range(n):
        blk = f"encoder.block.{i}"
        for j in (1, 2, 3):
            res_unit(f"{blk}.res_unit{j}")
        snake(blk + ".snake1")
        conv(blk + ".conv1")
    snake("encoder.snake1")
    conv("encoder.conv2")
    return b.finish()
