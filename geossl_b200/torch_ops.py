"""``torch.ops.geossl_b200.*`` -- the kernels registered with PyTorch's dispatcher (the Python-side equivalent of a
``TORCH_LIBRARY(geossl_b200, m)`` block: ``torch.library.Library`` + CUDA-only implementations).

Each op is a thin shim over one C-ABI entry point of libgeossl_b200.so (``include/geossl_b200.h``) through
``geossl_b200.ops``; there is deliberately NO CPU implementation, so a CPU tensor fails in the dispatcher
("no kernel for backend CPU") instead of falling back.  Differentiation lives in ``geossl_b200.ops``'s
``autograd.Function`` wrappers, which call the same entry points; these ops are the inference / composition surface
(``torch.ops.geossl_b200.cfconv(x, filt, rowptr, src)`` etc.) and what ``torch.library.opcheck``-style tooling sees.
"""
import torch

from . import ops

_lib = torch.library.Library("geossl_b200", "DEF")

_lib.define("radius_csr(Tensor pos, Tensor batch, float r, int max_num_neighbors, int num_graphs) -> (Tensor, Tensor, Tensor, Tensor)")
_lib.define("radius_graph(Tensor pos, Tensor batch, float r, int max_num_neighbors) -> Tensor")
_lib.define("cfconv(Tensor x, Tensor filt, Tensor rowptr, Tensor src) -> Tensor")
_lib.define("cfconv_transpose(Tensor filt, Tensor grad_out, Tensor t_rowptr, Tensor t_eid, Tensor t_tgt) -> Tensor")
_lib.define("cfconv_edge_product(Tensor x, Tensor grad_out, Tensor rowptr, Tensor src) -> Tensor")
_lib.define("filter_network(Tensor dist, Tensor n_edges, Tensor offset, float coeff, float cutoff, Tensor w1, Tensor b1, Tensor w2, "
            "Tensor b2) -> Tensor")
_lib.define("linear128(Tensor x, Tensor weight, Tensor? bias, bool pre_ssp, Tensor? residual) -> Tensor")
_lib.define("pair_distance(Tensor pos, Tensor super_edge_index) -> Tensor")


class _G:
    """Minimal stand-in for ops.RadiusCSR built from raw CSR tensors."""

    def __init__(self, n_atoms, capacity, **kw):
        self.n_atoms, self.capacity = n_atoms, capacity
        self.__dict__.update(kw)

    def ensure_transpose(self):
        return self


def _radius_csr(pos, batch, r, max_num_neighbors, num_graphs):
    g = ops.radius_csr(pos, batch, r, max_num_neighbors, num_graphs=num_graphs if num_graphs >= 0 else None, transpose=False)
    return g.rowptr, g.src, g.tgt, g.dist


def _radius_graph(pos, batch, r, max_num_neighbors):
    return ops.radius_graph(pos, r, batch, max_num_neighbors=max_num_neighbors)


def _cfconv(x, filt, rowptr, src):
    n = rowptr.numel() - 1
    return ops._cfconv_fwd(ops._req(x, torch.float32, "x", 2), ops._req(filt, torch.float32, "filt", 2),
                           _G(n, filt.size(0), rowptr=ops._req(rowptr, torch.int32, "rowptr", 1), src=ops._req(src, torch.int32, "src", 1)))


def _cfconv_transpose(filt, grad_out, t_rowptr, t_eid, t_tgt):
    n = t_rowptr.numel() - 1
    return ops._cfconv_bwd_x(ops._req(filt, torch.float32, "filt", 2), ops._req(grad_out, torch.float32, "grad_out", 2),
                             _G(n, filt.size(0), t_rowptr=ops._req(t_rowptr, torch.int32, "t_rowptr", 1),
                                t_eid=ops._req(t_eid, torch.int32, "t_eid", 1), t_tgt=ops._req(t_tgt, torch.int32, "t_tgt", 1)))


def _cfconv_edge_product(x, grad_out, rowptr, src):
    n = rowptr.numel() - 1
    return ops._cfconv_bwd_w(ops._req(x, torch.float32, "x", 2), ops._req(grad_out, torch.float32, "grad_out", 2),
                             _G(n, src.numel(), rowptr=ops._req(rowptr, torch.int32, "rowptr", 1), src=ops._req(src, torch.int32, "src", 1)))


def _filter_network(dist, n_edges, offset, coeff, cutoff, w1, b1, w2, b2):
    g = _G(0, dist.numel(), dist=ops._req(dist, torch.float32, "dist", 1), n_edges_dev=ops._req(n_edges, torch.int32, "n_edges", 1))
    return ops.filter_forward(g, ops._req(offset, torch.float32, "offset", 1), coeff, cutoff, *(ops._req(t, torch.float32, "param") for t in (w1, b1, w2, b2)))


def _linear128(x, weight, bias, pre_ssp, residual):
    x, weight = ops._req(x, torch.float32, "x", 2), ops._req(weight, torch.float32, "weight", 2)
    if x.size(1) != 128 or tuple(weight.shape) != (128, 128):
        raise RuntimeError("geossl_b200::linear128 is the 128 -> 128 tensor-core layer")
    return ops._linear_tc(x, weight, False, bias, pre_ssp, None, residual, False, "linear_fwd")


for _name, _fn in (("radius_csr", _radius_csr), ("radius_graph", _radius_graph), ("cfconv", _cfconv),
                   ("cfconv_transpose", _cfconv_transpose), ("cfconv_edge_product", _cfconv_edge_product),
                   ("filter_network", _filter_network), ("linear128", _linear128),
                   ("pair_distance", lambda pos, sei: ops.pair_distance(pos, sei))):
    _lib.impl(_name, _fn, "CUDA")

OP_NAMES = ("radius_csr", "radius_graph", "cfconv", "cfconv_transpose", "cfconv_edge_product", "filter_network", "linear128",
            "pair_distance")
