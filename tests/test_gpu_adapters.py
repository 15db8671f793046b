"""GPU parity of the weight-mutation path: a decoder that carries PEFT-shaped LoRA layers and a
LyCORIS-shaped LoKr net is folded by `pack.effective_decoder_state` (what `_init_b200_backends` packs after every
handler LoRA call, handler/lora/lifecycle.py:165-300, controls.py:35-206), and the CUDA forward of the folded
weights must equal the fp32 oracle evaluated with W + sum_a s_a B_a A_a (LoRA) / W + m * scale * kron(w1, w2) (LoKr).

`peft` and `lycoris` are not installed in this image: the adapter modules are duck-typed with the attribute names
those libraries use (base_layer / lora_A / lora_B / scaling / active_adapters; org_module / lokr_w1 / lokr_w2_a /
lokr_w2_b / scale / multiplier).  Tolerance: one DiT forward, rel-L2 <= 2e-2 vs the fp32 oracle on the same
bf16-rounded folded weights (as in test_gpu_kernels.py)."""
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from acestep_b200.dit import B200DiT, DiTShape  # noqa: E402
from acestep_b200.pack import effective_decoder_state  # noqa: E402
from oracle.dit import DiTConfig, dit_forward  # noqa: E402
from oracle.weights import bf16_round_, make_dit_weights  # noqa: E402

DEV = torch.device("cuda:0")


def module_tree(state):
    """nn.Module tree whose state_dict() is `state` (2-D '*.weight' leaves become real nn.Linear modules)."""
    root = torch.nn.Module()
    groups = {}
    for k, v in state.items():
        path, leaf = k.rsplit(".", 1) if "." in k else ("", k)
        groups.setdefault(path, {})[leaf] = v
    for path, leaves in groups.items():
        parent = root
        parts = path.split(".") if path else []
        if "weight" in leaves and leaves["weight"].dim() == 2 and parts:
            mod = torch.nn.Linear(leaves["weight"].shape[1], leaves["weight"].shape[0], bias="bias" in leaves)
            with torch.no_grad():
                mod.weight.copy_(leaves["weight"])
                if "bias" in leaves:
                    mod.bias.copy_(leaves["bias"])
            for p in parts[:-1]:
                if not hasattr(parent, p):
                    parent.add_module(p, torch.nn.Module())
                parent = getattr(parent, p)
            parent.add_module(parts[-1], mod)
            continue
        for p in parts:
            if not hasattr(parent, p):
                parent.add_module(p, torch.nn.Module())
            parent = getattr(parent, p)
        for leaf, v in leaves.items():
            parent.register_parameter(leaf, torch.nn.Parameter(v.clone(), requires_grad=False))
    return root


class LoraLinear(torch.nn.Module):  # the attributes PEFT's lora.Linear exposes
    def __init__(self, base, rank, gen):
        super().__init__()
        i, o = base.in_features, base.out_features
        self.base_layer = base
        self.lora_A = torch.nn.ModuleDict({n: torch.nn.Linear(i, rank, bias=False) for n in ("style", "voice")})
        self.lora_B = torch.nn.ModuleDict({n: torch.nn.Linear(rank, o, bias=False) for n in ("style", "voice")})
        with torch.no_grad():
            for n in ("style", "voice"):
                self.lora_A[n].weight.copy_(torch.randn(rank, i, generator=gen) * 0.3)
                self.lora_B[n].weight.copy_(torch.randn(o, rank, generator=gen) * 0.3)
        self.scaling = {"style": 4.0, "voice": 2.0}
        self.active_adapters = ["style"]
        self.disable_adapters = False


class Lokr:  # the attributes LyCORIS' LokrModule exposes
    def __init__(self, org, gen, f=4, rank=2):
        o, i = org.weight.shape
        self.org_module, self.lora_name = [org], "lycoris_lokr"
        self.lokr_w1 = torch.randn(f, f, generator=gen) * 0.5
        self.lokr_w2 = None
        self.lokr_w2_a = torch.randn(o // f, rank, generator=gen) * 0.3
        self.lokr_w2_b = torch.randn(rank, i // f, generator=gen) * 0.3
        self.scale, self.multiplier = 2.0, 1.2


class LycorisNet:
    def __init__(self, loras):
        self.loras = loras


def _wrap(root, path, factory):
    parts = path.split(".")
    parent = root
    for p in parts[:-1]:
        parent = getattr(parent, p)
    new = factory(getattr(parent, parts[-1]))
    setattr(parent, parts[-1], new)
    return new


def test_folded_lora_and_lokr_match_the_oracle_with_merged_weights():
    cfg = DiTConfig.tiny()
    w = bf16_round_(make_dit_weights(cfg, seed=0))
    gen = torch.Generator().manual_seed(21)
    dec = module_tree(w)
    assert set(dec.state_dict()) == set(w)
    a = _wrap(dec, "layers.0.self_attn.q_proj", lambda base: LoraLinear(base, 4, gen))
    b = _wrap(dec, "layers.1.mlp.down_proj", lambda base: LoraLinear(base, 4, gen))
    b.active_adapters = ["style", "voice"]
    lokr = Lokr(dec.layers[2].cross_attn.o_proj if hasattr(dec.layers, "__getitem__") else
                getattr(dec.layers, "2").cross_attn.o_proj, gen)
    dec._lycoris_net = LycorisNet([lokr])

    want_w = dict(w)
    d = lambda m, n: m.scaling[n] * (m.lora_B[n].weight @ m.lora_A[n].weight).detach()
    want_w["layers.0.self_attn.q_proj.weight"] = w["layers.0.self_attn.q_proj.weight"] + d(a, "style")
    want_w["layers.1.mlp.down_proj.weight"] = w["layers.1.mlp.down_proj.weight"] + d(b, "style") + d(b, "voice")
    want_w["layers.2.cross_attn.o_proj.weight"] = w["layers.2.cross_attn.o_proj.weight"] + \
        1.2 * 2.0 * torch.kron(lokr.lokr_w1, lokr.lokr_w2_a @ lokr.lokr_w2_b)

    folded = effective_decoder_state(dec)
    assert set(folded) == set(w)
    for k in want_w:
        assert torch.allclose(folded[k].float(), want_w[k], atol=1e-6), k
    want_w = bf16_round_(want_w)

    g = torch.Generator().manual_seed(22)
    B, T, E = 2, 77, 19
    xt = torch.randn(B, T, 64, generator=g).to(torch.bfloat16)
    ctx = torch.randn(B, T, 128, generator=g).to(torch.bfloat16)
    enc = torch.randn(B, E, cfg.hidden_size, generator=g).to(torch.bfloat16)
    t = torch.tensor([0.75, 0.25]).to(torch.bfloat16)
    want = dit_forward(want_w, cfg, xt.float(), t.float(), ctx.float(), enc.float(), bf16_time=True)
    base = dit_forward(w, cfg, xt.float(), t.float(), ctx.float(), enc.float(), bf16_time=True)
    assert rel_l2(base, want) > 4e-2, "the adapters must change the output well beyond the 2e-2 tolerance"

    dit = B200DiT(folded, DiTShape.from_config(cfg), DEV)
    dit.bind(B, T, E)
    dit.set_condition(enc.to(DEV))
    vt = dit.step(xt.to(DEV), ctx.to(DEV), t.float().tolist())
    torch.cuda.synchronize()
    assert rel_l2(vt.cpu().float(), want) <= 2e-2
    dit.close()

    # adapters switched off (set_use_lora(False): disable_adapters / multiplier 0) -> the plain decoder again
    a.disable_adapters = b.disable_adapters = True
    lokr.multiplier = 0.0
    plain = effective_decoder_state(dec)
    for k in w:
        assert torch.equal(plain[k].float(), w[k]), k
