"""Test helpers: product modules instantiated from a golden fixture's config + weights."""
import torch

from _golden import Golden


def full_sd(sd):
    """Golden files drop the aliased ``conv.nn.*`` keys (bit-identical to ``mlp.*``); restore them."""
    out = dict(sd)
    for k, v in sd.items():
        if ".mlp." in k:
            out[k.replace(".mlp.", ".conv.nn.")] = v
    return out


def schnet_from(g: Golden, device=None):
    from geossl_b200.Geom3D.models import SchNet
    c = g.cfg
    m = SchNet(hidden_channels=c.get("hidden", c.get("emb")), num_filters=c["filters"], num_interactions=c["layers"],
               num_gaussians=c["gaussians"], cutoff=c["cutoff"], node_class=9, readout=c.get("readout", "mean"))
    m.load_state_dict(full_sd(g.sd()), strict=True)
    return m.to(device) if device else m


def painn_from(g: Golden, device=None):
    from geossl_b200.Geom3D.models import PaiNN
    c = g.cfg
    m = PaiNN(n_atom_basis=c.get("feat", c.get("emb")), n_interactions=c["layers"], n_rbf=c["rbf"], cutoff=c["cutoff"],
              max_z=9, n_out=1, readout=c.get("readout", "add"))
    m.load_state_dict(g.sd(), strict=True)
    return m.to(device) if device else m


def head_from(g: Golden, grp="sd", device=None):
    from geossl_b200.NCSN import NCSN_version_03
    c = g.cfg
    m = NCSN_version_03(c["emb"], sigma_begin=10, sigma_end=0.01, num_noise_level=c["levels"], noise_type="symmetry",
                        anneal_power=c["anneal_power"])
    m.load_state_dict(g.sd(grp), strict=True)
    return m.to(device) if device else m


def grads_of(module):
    return {k: p.grad for k, p in module.named_parameters() if p.grad is not None}
