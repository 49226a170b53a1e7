"""Functional CPU restatement (torch fp32/fp64, autograd-differentiable) of the reference encoders
and of the DDM objective.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Every function takes the reference's ``state_dict`` (same key names) so that weights can be moved
between the reference modules, this oracle and the CUDA product without renaming.

Reference lines followed (all under /root/reference):
  SchNet            Geom3D/models/schnet.py:85-125 (forward), :163-167 (InteractionBlock),
                    :185-195 (CFConv), :198-207 (GaussianSmearing), :210-216 (ShiftedSoftplus)
  PaiNN             Geom3D/models/painn.py:32-66 (message), :91-114 (mixing), :216-269 (forward);
                    Geom3D/models/painn_utils.py:99-103 (rbf), :139-155 (cosine cutoff)
  NCSN_version_03   examples/NCSN.py:9-43 (MLP), :168-220
  perturb / do_DDM  examples/pretrain_GeoSSL.py:68-74, :179-212
Third-party pieces (radius_graph, scatter, propagate) are restated from SURVEY.md Appendix B
(parity unpinned, see oracle/radius.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .radius import radius_graph

LOG2 = torch.log(torch.tensor(2.0)).item()  # schnet.py:213 (fp32 log, widened to a python float)


# ----------------------------------------------------------------------------- helpers
def segment_sum(src, index, dim_size):
    """torch_scatter.scatter(..., reduce='sum') along dim 0 (Appendix B.2)."""
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add(0, index, src)


def segment_mean(src, index, dim_size):
    s = segment_sum(src, index, dim_size)
    cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device).index_add(
        0, index, torch.ones_like(index, dtype=src.dtype)).clamp(min=1)
    return s / cnt.view(-1, *([1] * (src.dim() - 1)))


def shifted_softplus(x):
    return F.softplus(x) - LOG2                                    # schnet.py:215-216


def smearing_coeff(offset):
    return -0.5 / (offset[1] - offset[0]).item() ** 2              # schnet.py:202


def gaussian_smearing(dist, offset):
    d = dist.view(-1, 1) - offset.view(1, -1)                      # schnet.py:206
    return torch.exp(smearing_coeff(offset) * torch.pow(d, 2))    # schnet.py:207


# ----------------------------------------------------------------------------- SchNet
def schnet_num_layers(sd):
    n = 0
    while f"interactions.{n}.lin.weight" in sd:
        n += 1
    return n


def schnet_filter(sd, layer, edge_weight, edge_attr, cutoff):
    """W_e of one CFConv (schnet.py:186-187)."""
    p = f"interactions.{layer}.mlp."
    C = 0.5 * (torch.cos(edge_weight * math.pi / cutoff) + 1.0)
    a = F.linear(edge_attr, sd[p + "0.weight"], sd[p + "0.bias"])
    s = shifted_softplus(a)
    return F.linear(s, sd[p + "2.weight"], sd[p + "2.bias"]) * C.view(-1, 1)


def cfconv_aggregate(x, W, edge_index):
    """PyG propagate with aggr='add' (schnet.py:190,194-195; Appendix B.3)."""
    return segment_sum(x.index_select(0, edge_index[0]) * W, edge_index[1], x.size(0))


def schnet_forward(sd, z, pos, batch=None, *, cutoff=10.0, readout="mean", edge_index=None,
                   return_edge_index=False):
    """Returns (out (B,H), h (N,H)) exactly as SchNet.forward(..., return_latent=True)."""
    assert z.dim() == 1 and z.dtype == torch.long
    batch = torch.zeros_like(z) if batch is None else batch
    h = F.embedding(z, sd["embedding.weight"])                                        # :89
    if edge_index is None:
        edge_index = radius_graph(pos, r=cutoff, batch=batch)                         # :91
    row, col = edge_index
    edge_weight = (pos[row] - pos[col]).norm(dim=-1)                                  # :93
    edge_attr = gaussian_smearing(edge_weight, sd["distance_expansion.offset"])      # :94
    for l in range(schnet_num_layers(sd)):                                            # :96-97
        p = f"interactions.{l}."
        W = schnet_filter(sd, l, edge_weight, edge_attr, cutoff)
        x = F.linear(h, sd[p + "conv.lin1.weight"])                                   # :189
        x = cfconv_aggregate(x, W, edge_index)                                        # :190
        x = F.linear(x, sd[p + "conv.lin2.weight"], sd[p + "conv.lin2.bias"])         # :191
        x = shifted_softplus(x)                                                       # :165
        x = F.linear(x, sd[p + "lin.weight"], sd[p + "lin.bias"])                     # :166
        h = h + x
    h = F.linear(h, sd["lin1.weight"], sd["lin1.bias"])                               # :99
    h = shifted_softplus(h)
    h = F.linear(h, sd["lin2.weight"], sd["lin2.bias"])                               # :101
    nb = int(batch.max()) + 1 if batch.numel() else 0
    if readout == "mean":
        out = segment_mean(h, batch, nb)                                              # :115
    else:
        out = segment_sum(h, batch, nb)
    if return_edge_index:
        return out, h, edge_index
    return out, h


# ----------------------------------------------------------------------------- PaiNN
def painn_num_layers(sd):
    n = 0
    while f"interactions.{n}.interatomic_context_net.0.weight" in sd:
        n += 1
    return n


def painn_forward(sd, x, positions, radius_edge_index, batch, *, readout="add", epsilon=1e-8,
                  activation=F.silu):
    """Returns (h (B,F), q (N,F)) as PaiNN.forward(..., return_latent=True) (painn.py:216-269)."""
    z = x[:, 0] if x.dim() == 2 else x
    idx_i, idx_j = radius_edge_index[0], radius_edge_index[1]
    r_ij = positions[idx_i] - positions[idx_j]                                        # :232
    n_atoms = z.size(0)
    d_ij = torch.norm(r_ij, dim=1, keepdim=True)                                      # :236
    dir_ij = r_ij / d_ij
    offsets, widths, rc = sd["radial_basis.offsets"], sd["radial_basis.widths"], sd["cutoff_fn.cutoff"]
    coeff = -0.5 / torch.pow(widths, 2)                                               # utils:100
    phi = torch.exp(coeff * torch.pow(d_ij[..., None] - offsets, 2))                  # (E,1,R)
    fcut = 0.5 * (torch.cos(d_ij * math.pi / rc) + 1.0)                               # utils:152
    fcut = fcut * (d_ij < rc).float()
    filters = F.linear(phi, sd["filter_net.weight"], sd["filter_net.bias"]) * fcut[..., None]
    nf = sd["embedding.weight"].size(1)
    n_int = painn_num_layers(sd)
    filter_list = torch.split(filters, 3 * nf, dim=-1)                                # :245
    q = F.embedding(z, sd["embedding.weight"], padding_idx=0)[:, None]                # :174,247
    mu = torch.zeros((q.shape[0], 3, q.shape[2]), dtype=q.dtype, device=q.device)
    for l in range(n_int):
        p = f"interactions.{l}.interatomic_context_net."
        xx = activation(F.linear(q, sd[p + "0.weight"], sd[p + "0.bias"]))
        xx = F.linear(xx, sd[p + "1.weight"], sd[p + "1.bias"])                       # (N,1,3F)
        xj, muj = xx[idx_j], mu[idx_j]                                                # :54-55
        xe = filter_list[l] * xj
        dq, dmuR, dmumu = torch.split(xe, nf, dim=-1)
        dq = segment_sum(dq, idx_i, n_atoms)                                          # :59
        dmu = dmuR * dir_ij[..., None] + dmumu * muj                                  # :60
        dmu = segment_sum(dmu, idx_i, n_atoms)
        q, mu = q + dq, mu + dmu
        m = f"mixing.{l}."
        mu_mix = F.linear(mu, sd[m + "mu_channel_mix.weight"])                        # :100
        mu_V, mu_W = torch.split(mu_mix, nf, dim=-1)
        mu_Vn = torch.sqrt(torch.sum(mu_V ** 2, dim=-2, keepdim=True) + epsilon)
        ctx = torch.cat([q, mu_Vn], dim=-1)
        c = m + "intraatomic_context_net."
        y = activation(F.linear(ctx, sd[c + "0.weight"], sd[c + "0.bias"]))
        y = F.linear(y, sd[c + "1.weight"], sd[c + "1.bias"])
        dq_intra, dmu_intra, dqmu_intra = torch.split(y, nf, dim=-1)
        dmu_intra = dmu_intra * mu_W
        dqmu_intra = dqmu_intra * torch.sum(mu_V * mu_W, dim=1, keepdim=True)
        q = q + dq_intra + dqmu_intra
        mu = mu + dmu_intra
    q = q.squeeze(1)
    nb = int(batch.max()) + 1 if batch.numel() else 0
    h = segment_mean(q, batch, nb) if readout == "mean" else segment_sum(q, batch, nb)
    return h, q


# ----------------------------------------------------------------------------- DDM
def ncsn_sigmas(sigma_begin, sigma_end, num_noise_level):
    """float64 numpy schedule cast to fp32 (NCSN.py:178)."""
    return torch.tensor(np.exp(np.linspace(np.log(sigma_begin), np.log(sigma_end), num_noise_level)),
                        dtype=torch.float32)


def _mlp(sd, prefix, x, n_layers):
    for i in range(n_layers):                                                         # NCSN.py:33-43
        x = F.linear(x, sd[f"{prefix}.layers.{i}.weight"], sd[f"{prefix}.layers.{i}.bias"])
        if i < n_layers - 1:
            x = F.relu(x)
    return x


def ncsn_forward(sd, batch, super_edge_index, node_feature, distance, noise_level, distance_noise,
                 anneal_power, num_graphs=None):
    """NCSN_version_03.forward (NCSN.py:183-212) with the two random draws (:190, :194) injected."""
    edge2graph = batch[super_edge_index[0]]                                           # :187
    used_sigmas = sd["sigmas"][noise_level]                                           # :191
    used_sigmas = used_sigmas[edge2graph].unsqueeze(-1)                               # :192
    perturbed = distance + distance_noise * used_sigmas                               # :196
    emb = _mlp(sd, "input_distance_mlp", perturbed, 2)                                # :197
    target = -1 / (used_sigmas ** 2) * (perturbed - distance)                         # :199
    h_row, h_col = node_feature[super_edge_index[0]], node_feature[super_edge_index[1]]
    feat = torch.cat([h_row + h_col, emb], dim=-1)                                    # :203
    scores = _mlp(sd, "output_mlp", feat, 3)                                          # :204
    scores = scores * (1. / used_sigmas)                                              # :205
    target, scores = target.view(-1), scores.view(-1)
    loss = 0.5 * ((scores - target) ** 2) * (used_sigmas.squeeze(-1) ** anneal_power)  # :209
    ng = int(edge2graph.max()) + 1 if edge2graph.numel() else 0                       # scatter dim_size
    loss = segment_sum(loss, edge2graph, ng)                                          # :210
    return loss.mean()                                                                # :212


def pair_distance(pos, super_edge_index):
    u = torch.index_select(pos, 0, super_edge_index[0])                               # pretrain:199-201
    v = torch.index_select(pos, 0, super_edge_index[1])
    return torch.sqrt(torch.sum((u - v) ** 2, dim=1)).unsqueeze(1)


def perturb(x, positions, mu, sigma, noise=None):
    """pretrain_GeoSSL.py:68-74; ``noise`` lets a test inject the torch.normal draw."""
    if noise is None:
        noise = torch.normal(mu, sigma, size=positions.size())
    return x, positions + noise.to(positions.device)


def ddm_loss(encoder, sd_head1, sd_head2, x, positions, positions_02, batch, super_edge_index,
             draws1, draws2, anneal_power, normalize=False):
    """do_DDM (pretrain_GeoSSL.py:179-212).  ``encoder(x, pos)`` returns node representations;
    ``draws*`` = (noise_level (B,), distance_noise (P,1))."""
    repr_01 = encoder(x, positions)
    repr_02 = encoder(x, positions_02)
    if normalize:
        repr_01, repr_02 = F.normalize(repr_01, dim=-1), F.normalize(repr_02, dim=-1)
    d01 = pair_distance(positions, super_edge_index)
    d02 = pair_distance(positions_02, super_edge_index)
    l1 = ncsn_forward(sd_head1, batch, super_edge_index, repr_01, d02, draws1[0], draws1[1], anneal_power)
    l2 = ncsn_forward(sd_head2, batch, super_edge_index, repr_02, d01, draws2[0], draws2[1], anneal_power)
    return (l1 + l2) / 2, (repr_01, repr_02, l1, l2)


# ------------------------------------------------------------------------------------------ sibling objectives
# the objectives that share the two-view encoder pass with do_DDM.
def cycle_index(num, shift):
    """examples/util.py:19-22."""
    arr = torch.arange(num) + shift
    arr[-shift:] = torch.arange(shift)
    return arr


def info_nce_loss(repr_01, repr_02, T, normalize=False):
    """do_InfoNCE (pretrain_GeoSSL.py:141-176) on the two (B,H) molecule representations: returns (loss, acc)."""
    if normalize:
        repr_01, repr_02 = F.normalize(repr_01, dim=-1), F.normalize(repr_02, dim=-1)

    def cal_loss(X, Y):                                                               # :159-168
        B = X.size(0)
        logits = torch.div(torch.mm(X, Y.transpose(1, 0)), T)
        labels = torch.arange(B, device=logits.device)
        loss = F.cross_entropy(logits, labels)
        acc = logits.argmax(dim=1).eq(labels).sum().item() * 1. / B
        return loss, acc
    l1, a1 = cal_loss(repr_01, repr_02)
    l2, a2 = cal_loss(repr_02, repr_01)
    return (l1 + l2) / 2, (a1 + a2) / 2                                               # :173-176


def ebm_nce_loss(repr_01, repr_02, num_neg=1, normalize=False):
    """do_EBM_NCE (pretrain_GeoSSL.py:103-138) with criterion = BCEWithLogitsLoss (:344), computed in float64."""
    if normalize:
        repr_01, repr_02 = F.normalize(repr_01, dim=-1), F.normalize(repr_02, dim=-1)
    B = len(repr_01)
    neg_01 = repr_01.repeat((num_neg, 1))                                             # :125
    neg_02 = torch.cat([repr_02[cycle_index(B, i + 1)] for i in range(num_neg)], dim=0)  # :126
    pred_pos = torch.sum(repr_01 * repr_02, dim=1)
    pred_neg = torch.sum(neg_01 * neg_02, dim=1)
    loss_pos = F.binary_cross_entropy_with_logits(pred_pos.double(), torch.ones(B, dtype=torch.float64, device=pred_pos.device))
    loss_neg = F.binary_cross_entropy_with_logits(pred_neg.double(), torch.zeros(B * num_neg, dtype=torch.float64, device=pred_pos.device))
    loss = (loss_pos + num_neg * loss_neg) / (1 + num_neg)                            # :134
    acc = (torch.sum(pred_pos > 0).float() + torch.sum(pred_neg < 0).float()) / (len(pred_pos) + len(pred_neg))
    return loss, acc.item()


def distance_prediction_loss(sd, node_repr, positions, super_edge_index):
    """DistancePredictor.forward + the pair block of train() (pretrain_DistancePrediction.py:15-26,72-79);
    ``sd`` holds ``predictor.weight`` (1,2H) and ``predictor.bias`` (1,)."""
    u = torch.index_select(node_repr, 0, super_edge_index[0])
    v = torch.index_select(node_repr, 0, super_edge_index[1])
    d = torch.sqrt(torch.sum((positions[super_edge_index[0]] - positions[super_edge_index[1]]) ** 2, dim=1))
    pred = F.linear(torch.cat([u, v], dim=1), sd["predictor.weight"], sd["predictor.bias"]).squeeze()
    return F.l1_loss(pred, d)
