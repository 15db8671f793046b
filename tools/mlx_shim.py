"""A minimal torch-backed stand-in for the `mlx` package (TEST INFRASTRUCTURE, build container only).

Purpose: execute the reference's OWN in-tree Oobleck implementation (`acestep/models/mlx/vae_model.py`,
`vae_convert.py`) on this Linux box, where `mlx` (Apple-silicon only) cannot be installed, so that the oracle's
VAE restatement can be pinned against reference code instead of being "parity unpinned".  Only the primitives
those two files touch are provided, with MLX's documented semantics:

  * arrays are torch tensors (fp32); `mx.array(np)`, `zeros`, `exp`, `sin`, `log`, `power`, `reciprocal`,
    `where`, `split(x, n, axis)`, `random.normal(shape)`, `eval` (no-op: torch is eager);
  * `nn.Conv1d` / `nn.ConvTranspose1d`: data in NLC, weight [C_out, K, C_in] (MLX layout), symmetric padding;
    evaluated with torch's conv1d / conv_transpose1d after moving to NCL and to torch's weight layouts
    ([C_out, C_in, K] / [C_in, C_out, K]) — exactly the inverse of the axis moves `vae_convert.py:86-91` applies;
  * `nn.Module`: attribute tree with lists of sub-modules, `load_weights([(dotted.name, array), ...])`,
    `parameters()`.

Everything structural — layer order, paddings, strides, dilations, the Snake formula, the weight-norm fusion,
the posterior — stays the reference's code.  install() registers the modules in sys.modules."""
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F


def _array(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x.float()
    return torch.from_numpy(np.asarray(x, dtype=np.float32).copy())


class Module:
    def __call__(self, *a, **k):  # subclasses define __call__ themselves (MLX style)
        raise NotImplementedError

    def _resolve(self, path):
        obj = self
        for part in path:
            obj = obj[int(part)] if isinstance(obj, (list, tuple)) else getattr(obj, part)
        return obj

    def load_weights(self, weights, strict=True):
        for name, value in weights:
            *parents, leaf = name.split(".")
            owner = self._resolve(parents)
            cur = getattr(owner, leaf)
            value = _array(value)
            if tuple(cur.shape) != tuple(value.shape):
                raise ValueError(f"{name}: shape {tuple(value.shape)} does not match {tuple(cur.shape)}")
            setattr(owner, leaf, value)
        return self

    def parameters(self):
        out = {}
        for k, v in vars(self).items():
            if isinstance(v, torch.Tensor):
                out[k] = v
            elif isinstance(v, Module):
                out[k] = v.parameters()
            elif isinstance(v, list) and v and isinstance(v[0], Module):
                out[k] = [m.parameters() for m in v]
        return out


class Conv1d(Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
        assert groups == 1
        self.weight = torch.zeros(out_channels, kernel_size, in_channels)
        if bias:
            self.bias = torch.zeros(out_channels)
        self.stride, self.padding, self.dilation = stride, padding, dilation

    def __call__(self, x):  # x [N, L, C_in]
        y = F.conv1d(x.transpose(1, 2), self.weight.permute(0, 2, 1), getattr(self, "bias", None),
                     stride=self.stride, padding=self.padding, dilation=self.dilation)
        return y.transpose(1, 2)


class ConvTranspose1d(Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, output_padding=0,
                 bias=True):
        assert dilation == 1 and output_padding == 0
        self.weight = torch.zeros(out_channels, kernel_size, in_channels)
        if bias:
            self.bias = torch.zeros(out_channels)
        self.stride, self.padding = stride, padding

    def __call__(self, x):  # x [N, L, C_in]
        y = F.conv_transpose1d(x.transpose(1, 2), self.weight.permute(2, 0, 1), getattr(self, "bias", None),
                               stride=self.stride, padding=self.padding)
        return y.transpose(1, 2)


def install():
    mlx = types.ModuleType("mlx")
    core = types.ModuleType("mlx.core")
    nn = types.ModuleType("mlx.nn")
    core.array = _array
    core.float32 = torch.float32
    core.zeros = lambda *shape: torch.zeros(*shape)
    core.exp, core.sin, core.log = torch.exp, torch.sin, torch.log
    core.power = lambda x, p: torch.pow(x, p)
    core.reciprocal = torch.reciprocal
    core.where = lambda c, a, b: torch.where(c, a, b)
    core.split = lambda x, n, axis=0: torch.chunk(x, n, dim=axis)
    core.eval = lambda *a, **k: None
    core.random = types.SimpleNamespace(normal=lambda shape: torch.randn(*shape))
    nn.Module, nn.Conv1d, nn.ConvTranspose1d = Module, Conv1d, ConvTranspose1d
    mlx.core, mlx.nn = core, nn
    sys.modules["mlx"], sys.modules["mlx.core"], sys.modules["mlx.nn"] = mlx, core, nn
    return mlx
