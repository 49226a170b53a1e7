"""Building blocks of the PaiNN drop-in (mirror of /root/reference/Geom3D/models/painn_utils.py: same
names, signatures and buffer/parameter names so that ``state_dict`` keys match)."""
import math
from typing import Callable, Optional, Sequence, Union

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import xavier_uniform_, zeros_


class Dense(nn.Linear):
    """Linear layer with an optional activation, xavier-uniform weight and zero bias (painn_utils.py:9-35)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True,
                 activation: Union[Callable, nn.Module] = None, weight_init: Callable = xavier_uniform_,
                 bias_init: Callable = zeros_):
        self.weight_init = weight_init
        self.bias_init = bias_init
        super().__init__(in_features, out_features, bias)
        self.activation = nn.Identity() if activation is None else activation

    def reset_parameters(self):
        self.weight_init(self.weight)
        if self.bias is not None:
            self.bias_init(self.bias)

    def forward(self, input: torch.Tensor):
        return self.activation(F.linear(input, self.weight, self.bias))


def build_mlp(n_in: int, n_out: int, n_hidden: Optional[Union[int, Sequence[int]]] = None, n_layers: int = 2,
              activation: Callable = F.silu) -> nn.Module:
    """Pyramidal (or fixed-width) MLP of Dense layers (painn_utils.py:38-70)."""
    if n_hidden is None:
        widths, c = [], n_in
        for _ in range(n_layers):
            widths.append(c)
            c = max(n_out, c // 2)
        widths.append(n_out)
    else:
        hidden = [n_hidden] * (n_layers - 1) if type(n_hidden) is int else list(n_hidden)
        widths = [n_in] + hidden + [n_out]
    layers = [Dense(widths[i], widths[i + 1], activation=activation) for i in range(n_layers - 1)]
    layers.append(Dense(widths[-2], widths[-1], activation=None))
    return nn.Sequential(*layers)


def scatter_add(x: torch.Tensor, idx_i: torch.Tensor, dim_size: int, dim: int = 0) -> torch.Tensor:
    shape = list(x.shape)
    shape[dim] = dim_size
    return torch.zeros(shape, dtype=x.dtype, device=x.device).index_add(dim, idx_i, x)


def replicate_module(module_factory: Callable[[], nn.Module], n: int, share_params: bool):
    if share_params:
        return nn.ModuleList([module_factory()] * n)
    return nn.ModuleList([module_factory() for _ in range(n)])


def gaussian_rbf(inputs: torch.Tensor, offsets: torch.Tensor, widths: torch.Tensor):
    coeff = -0.5 / torch.pow(widths, 2)
    diff = inputs[..., None] - offsets
    return torch.exp(coeff * torch.pow(diff, 2))


class GaussianRBF(nn.Module):
    """Gaussian radial basis (painn_utils.py:106-136): buffers ``widths`` and ``offsets``."""

    def __init__(self, n_rbf: int, cutoff: float, start: float = 0.0, trainable: bool = False):
        super().__init__()
        self.n_rbf = n_rbf
        offset = torch.linspace(start, cutoff, n_rbf)
        widths = torch.FloatTensor(torch.abs(offset[1] - offset[0]) * torch.ones_like(offset))
        if trainable:
            self.widths = nn.Parameter(widths)
            self.offsets = nn.Parameter(offset)
        else:
            self.register_buffer("widths", widths)
            self.register_buffer("offsets", offset)

    def forward(self, inputs: torch.Tensor):
        return gaussian_rbf(inputs, self.offsets, self.widths)


def cosine_cutoff(input: torch.Tensor, cutoff: torch.Tensor):
    """Behler cosine cutoff with the hard ``d < rc`` mask (painn_utils.py:139-155)."""
    input_cut = 0.5 * (torch.cos(input * math.pi / cutoff) + 1.0)
    return input_cut * (input < cutoff).float()


class CosineCutoff(nn.Module):
    def __init__(self, cutoff: float):
        super().__init__()
        self.register_buffer("cutoff", torch.FloatTensor([cutoff]))

    def forward(self, input: torch.Tensor):
        return cosine_cutoff(input, self.cutoff)
