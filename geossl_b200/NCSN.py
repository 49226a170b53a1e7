"""DDM score-matching head -- drop-in for ``NCSN_version_03`` and ``MultiLayerPerceptron`` of
/root/reference/examples/NCSN.py (:9-43, :168-220): same constructor, ``forward(data, node_feature,
distance, debug=False)``, parameter creation order and ``state_dict`` keys (``sigmas``,
``input_distance_mlp.layers.{0,1}.*``, ``output_mlp.layers.{0,1,2}.*``).

The forward draws the per-graph noise level and the per-pair N(0,1) noise with the same two torch RNG
calls, in the same order, as the reference (RNG contract); everything after the draws
is one fused CUDA kernel (ops.DDMHead).  ``noise_level`` / ``distance_noise`` can be injected for parity tests.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class MultiLayerPerceptron(nn.Module):
    def __init__(self, input_dim, hidden_dims, activation="relu", dropout=0):
        super().__init__()
        self.dims = [input_dim] + hidden_dims
        self.activation = getattr(F, activation) if isinstance(activation, str) else None
        self.dropout = nn.Dropout(dropout) if dropout else None
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(self.dims[:-1], self.dims[1:]))
        self.reset_parameters()

    def reset_parameters(self):
        for layer in self.layers:
            nn.init.xavier_uniform_(layer.weight)
            nn.init.constant_(layer.bias, 0.)

    def forward(self, input):
        # stand-alone use (library GEMMs); NCSN_version_03.forward below does not go through here
        x = input
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if i < len(self.layers) - 1:
                if self.activation:
                    x = self.activation(x)
                if self.dropout:
                    x = self.dropout(x)
        return x


class NCSN_version_03(torch.nn.Module):
    def __init__(self, emb_dim, sigma_begin, sigma_end, num_noise_level, noise_type, anneal_power):
        super().__init__()
        self.anneal_power = anneal_power
        self.noise_type = noise_type
        self.input_distance_mlp = MultiLayerPerceptron(1, [emb_dim, 1], activation="relu")
        self.output_mlp = MultiLayerPerceptron(1 + emb_dim, [emb_dim, emb_dim // 2, 1])
        # float64 numpy schedule cast to fp32, kept as a frozen Parameter like the reference (NCSN.py:178-179)
        sigmas = torch.tensor(np.exp(np.linspace(np.log(sigma_begin), np.log(sigma_end), num_noise_level)),
                              dtype=torch.float32)
        self.sigmas = nn.Parameter(sigmas, requires_grad=False)

    def _mlp_parameters(self):
        i, o = self.input_distance_mlp.layers, self.output_mlp.layers
        return (i[0].weight, i[0].bias, i[1].weight, i[1].bias,
                o[0].weight, o[0].bias, o[1].weight, o[1].bias, o[2].weight, o[2].bias)

    def forward(self, data, node_feature, distance, debug=False, noise_level=None, distance_noise=None):
        self.device = self.sigmas.device
        if noise_level is None:
            noise_level = torch.randint(0, self.sigmas.size(0), (data.num_graphs,), device=self.device)   # NCSN.py:190
        if distance_noise is None:
            distance_noise = torch.randn_like(distance)                                                   # NCSN.py:194
        # capacity-padded batches (pretrain.pad_batch) carry the live pair count on the device
        n_pairs_live = getattr(data, "extras", {}).get("n_pairs_live") if hasattr(data, "extras") else None
        loss = ops.DDMHead.apply(node_feature, data.super_edge_index, data.batch, distance, distance_noise,
                                 noise_level, self.sigmas, self.anneal_power, n_pairs_live, torch.is_grad_enabled(),
                                 *self._mlp_parameters())
        if debug:
            print("distance", distance[:10].squeeze())
            print("loss", loss)
        return loss
