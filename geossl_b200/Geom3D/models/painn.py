"""PaiNN encoder -- drop-in for /root/reference/Geom3D/models/painn.py (same classes, signatures,
parameter creation order and ``state_dict`` keys).

Edge-wise work (geometry, rbf, cutoff, the 3F-wide filter, the scalar/vector message and its
aggregation over ``idx_i``) runs in the CUDA message kernels of libgeossl_b200 (ops.PaiNNMessage);
node-level Dense layers are library GEMMs through torch.
"""
from typing import Callable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from .painn_utils import CosineCutoff, Dense, GaussianRBF, build_mlp, replicate_module, scatter_add


class PaiNNInteraction(nn.Module):
    """Scalar/vector message block (painn.py:14-66)."""

    def __init__(self, n_atom_basis: int, activation: Callable):
        super().__init__()
        self.n_atom_basis = n_atom_basis
        self.interatomic_context_net = nn.Sequential(
            Dense(n_atom_basis, n_atom_basis, activation=activation),
            Dense(n_atom_basis, 3 * n_atom_basis, activation=None),
        )

    def forward(self, q, mu, Wij, dir_ij, idx_i, idx_j, n_atoms):
        """Reference signature with a materialised filter ``Wij`` (E,1,3F) -- stand-alone API."""
        x = self.interatomic_context_net(q)
        xe = Wij * x[idx_j]
        dq, dmuR, dmumu = torch.split(xe, self.n_atom_basis, dim=-1)
        dq = scatter_add(dq, idx_i, dim_size=n_atoms)
        dmu = scatter_add(dmuR * dir_ij[..., None] + dmumu * mu[idx_j], idx_i, dim_size=n_atoms)
        return q + dq, mu + dmu


class PaiNNMixing(nn.Module):
    """Intra-atomic update block (painn.py:69-114)."""

    def __init__(self, n_atom_basis: int, activation: Callable, epsilon: float = 1e-8):
        super().__init__()
        self.n_atom_basis = n_atom_basis
        self.intraatomic_context_net = nn.Sequential(
            Dense(2 * n_atom_basis, n_atom_basis, activation=activation),
            Dense(n_atom_basis, 3 * n_atom_basis, activation=None),
        )
        self.mu_channel_mix = Dense(n_atom_basis, 2 * n_atom_basis, activation=None, bias=False)
        self.epsilon = epsilon

    def forward_fused(self, q, mu, images=None):
        """q (N,F), mu (N,3,F): the three Dense layers (128->256 without bias on the 3N vector rows, 256->128 + SiLU,
        128->384) run as 128 x 128 tensor-core blocks (ops.DenseTC) when ``images`` holds their packed weights, else as
        library GEMMs; everything between them is two fused kernels."""
        n, _, Fd = mu.shape
        net = self.intraatomic_context_net
        if images and id(self.mu_channel_mix.weight) in images:
            mix, mu_rows = self.mu_channel_mix, mu.reshape(3 * n, Fd)
            if ops.dense_chain2_applies(mu_rows, mix, None, images) and mix.bias is None:
                # both 128-column blocks of mu_channel_mix in one launch (the row tile is staged once)
                mu_mix = ops.DenseChain2.apply(mu_rows, mix.weight, None, None, None, images[id(mix.weight)], None).view(n, 3, 2 * Fd)
            else:
                mu_mix = ops.dense(mu_rows, mix, images).view(n, 3, 2 * Fd)
            ctx, dot = ops.PaiNNMixPre.apply(q, mu_mix, self.epsilon)
            if ops.dense_chain2_applies(ctx, net[0], net[1], images):
                # 256 -> 128 (two K-blocks summed in the accumulator) -> SiLU -> 384 (three blocks) without leaving the SM
                y = ops.DenseChain2.apply(ctx, net[0].weight, net[0].bias, net[1].weight, net[1].bias,
                                          images[id(net[0].weight)], images[id(net[1].weight)])
            else:
                y = ops.dense(ops.dense(ctx, net[0], images), net[1], images, pre_act=ops.ACT_SILU)
        else:
            mu_mix = self.mu_channel_mix(mu)
            ctx, dot = ops.PaiNNMixPre.apply(q, mu_mix, self.epsilon)
            y = net(ctx)
        return ops.PaiNNMixPost.apply(q, mu, y, mu_mix, dot)

    def forward(self, q, mu):
        mu_V, mu_W = torch.split(self.mu_channel_mix(mu), self.n_atom_basis, dim=-1)
        mu_Vn = torch.sqrt(torch.sum(mu_V ** 2, dim=-2, keepdim=True) + self.epsilon)
        x = self.intraatomic_context_net(torch.cat([q, mu_Vn], dim=-1))
        dq_intra, dmu_intra, dqmu_intra = torch.split(x, self.n_atom_basis, dim=-1)
        q = q + dq_intra + dqmu_intra * torch.sum(mu_V * mu_W, dim=1, keepdim=True)
        mu = mu + dmu_intra * mu_W
        return q, mu


class PaiNN(nn.Module):
    def __init__(self, n_atom_basis: int, n_interactions: int, n_rbf: int, cutoff: float, n_out: int, readout: str,
                 n_out_hidden: int = None, n_out_layers: int = 2, activation: Optional[Callable] = F.silu,
                 max_z: int = 100, shared_interactions: bool = False, shared_filters: bool = False,
                 epsilon: float = 1e-8):
        super().__init__()
        self.n_atom_basis = n_atom_basis
        self.n_interactions = n_interactions
        self.n_out = n_out
        self.n_out_hidden = n_out_hidden
        self.n_out_layers = n_out_layers
        self.activation = activation
        self.cutoff = cutoff
        self.cutoff_fn = CosineCutoff(cutoff)
        self.radial_basis = GaussianRBF(n_rbf=n_rbf, cutoff=cutoff)
        self.readout = readout
        self.embedding = nn.Embedding(max_z, n_atom_basis, padding_idx=0)
        self.share_filters = shared_filters
        n_filter_out = 3 * n_atom_basis if shared_filters else self.n_interactions * n_atom_basis * 3
        self.filter_net = Dense(self.radial_basis.n_rbf, n_filter_out, activation=None)
        self.interactions = replicate_module(
            lambda: PaiNNInteraction(n_atom_basis=self.n_atom_basis, activation=activation),
            self.n_interactions, shared_interactions)
        self.mixing = replicate_module(
            lambda: PaiNNMixing(n_atom_basis=self.n_atom_basis, activation=activation, epsilon=epsilon),
            self.n_interactions, shared_interactions)

    def create_output_layers(self):
        return build_mlp(n_in=self.n_atom_basis, n_out=self.n_out, n_hidden=self.n_out_hidden,
                         n_layers=self.n_out_layers, activation=self.activation)

    def forward(self, x, positions, radius_edge_index, batch, return_latent=False, num_graphs=None, assume_sorted=False):
        """``num_graphs`` / ``assume_sorted`` are optional extensions that remove the two host syncs of this forward
        (graph count, sortedness check of the edge list) so that a training step can be captured in a CUDA graph."""
        atomic_numbers = x[:, 0] if x.dim() == 2 else x
        n_atoms = atomic_numbers.size(0)
        Fd = self.n_atom_basis
        q = ops.embedding(self.embedding, atomic_numbers)       # (N,F)
        mu = torch.zeros((n_atoms, 3, Fd), dtype=q.dtype, device=q.device)
        edges = ops.painn_edges(positions, radius_edge_index, n_atoms, batch, self.radial_basis.offsets,
                                self.radial_basis.widths, self.cutoff, num_graphs=num_graphs, assume_sorted=assume_sorted)
        # tensor-core path (F = 128, silu): every Dense layer as 128 x 128 blocks, and the filter Dense(n_rbf -> 3F) of
        # painn.py:241 as a GEMM over the K-padded rbf matrix [phi, 1, 0...] against [W_f | b_f | 0] (bias folded into
        # column n_rbf), materialising the pre-cutoff filter rows per interaction; the slicing / padding below is plain
        # torch so that autograd carries the GEMM's weight gradient back to filter_net.{weight,bias}
        tc = (ops.FILTER_MODE != "simt" and q.is_cuda and Fd == 128 and self.activation is F.silu
              and self.radial_basis.n_rbf < 128 and not self.share_filters)
        images, wf_pad = None, None
        if tc:
            R = self.radial_basis.n_rbf
            buf = self.__dict__.get("_wf_pad_buf")
            if buf is None or buf.device != q.device or buf.size(0) != self.filter_net.weight.size(0):
                buf = self.__dict__["_wf_pad_buf"] = torch.zeros((self.filter_net.weight.size(0), 128), dtype=torch.float32,
                                                                 device=q.device)
            wf_pad = ops.FilterPad.apply(self.filter_net.weight, self.filter_net.bias, buf.detach())
            mats = [wf_pad]
            for interaction, mixing in zip(self.interactions, self.mixing):
                mats += [interaction.interatomic_context_net[0].weight, interaction.interatomic_context_net[1].weight,
                         mixing.mu_channel_mix.weight, mixing.intraatomic_context_net[0].weight,
                         mixing.intraatomic_context_net[1].weight]
            images = ops.prepack_dense_blocks(mats)
        for i, (interaction, mixing) in enumerate(zip(self.interactions, self.mixing)):
            fo = 0 if self.share_filters else i * 3 * Fd
            if tc:
                net = interaction.interatomic_context_net
                if ops.dense_chain2_applies(q, net[0], net[1], images):
                    ctx = ops.DenseChain2.apply(q, net[0].weight, net[0].bias, net[1].weight, net[1].bias,
                                                images[id(net[0].weight)], images[id(net[1].weight)])             # (N,3F), one launch
                else:
                    ctx = ops.dense(ops.dense(q, net[0], images), net[1], images, pre_act=ops.ACT_SILU)      # (N,3F)
                wpre = ops.DenseTC.apply(edges.phi_pad(), wf_pad[fo:fo + 3 * Fd], None, ops.ACT_NONE,
                                         images[id(wf_pad)][3 * i:3 * i + 3])                                   # (E,3F)
                q, mu = ops.PaiNNMessage.apply(q, mu, ctx, None, None, edges, wpre)
            else:
                ctx = interaction.interatomic_context_net(q)        # (N,3F)
                q, mu = ops.PaiNNMessage.apply(q, mu, ctx, self.filter_net.weight[fo:fo + 3 * Fd],
                                               self.filter_net.bias[fo:fo + 3 * Fd], edges)
            q, mu = mixing.forward_fused(q, mu, images)
        if num_graphs is None:
            num_graphs = int(batch[-1].item()) + 1 if batch.numel() else 0
        h = torch.zeros((num_graphs, Fd), dtype=q.dtype, device=q.device).index_add_(0, batch, q)
        if self.readout == "mean":
            cnt = torch.zeros(num_graphs, dtype=q.dtype, device=q.device).index_add_(
                0, batch, torch.ones(n_atoms, dtype=q.dtype, device=q.device)).clamp_(min=1)
            h = h / cnt.view(-1, 1)
        if return_latent:
            return h, q
        return h
