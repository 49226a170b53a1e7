"""Host-side operator layer: torch tensors in, C-ABI kernel launches (geossl_b200/_lib.py) out.

Everything here is plumbing -- allocation through torch's caching allocator, the current CUDA stream,
and ``torch.autograd.Function`` wrappers that pair each forward kernel with its backward kernel.  No
numerics happen in Python and there is no CPU path: a non-CUDA tensor raises.
"""
import contextlib
import ctypes
import os
from typing import Optional

import torch

from . import _lib
from ._lib import DdmPtrs, check

MAX_NEIGHBORS_DEFAULT = 32      # torch_cluster.radius_graph default, never overridden (schnet.py:91)


class _KernelTimers:
    """Optional CUDA-event brackets around named C-ABI calls (bench.py's live per-kernel durations).
    Events are recorded on torch's current stream, which is the stream every kernel here is launched on.

    ``enable(names, in_graph=True)``: the events are created ``external=True`` so that, recorded while a stream is being
    captured, they become event-record NODES of the CUDA graph; every replay re-records them and ``collect_replay()``
    reads the durations of the kernels as they ran inside the replayed step (back to back, warm instruction cache, no
    host enqueue gap between the opening record and the launch -- the eager brackets of round 1 included that gap)."""

    def __init__(self):
        self.names, self.events, self.in_graph = frozenset(), {}, False

    def enable(self, names, in_graph=False):
        self.names, self.events, self.in_graph = frozenset(names), {}, in_graph

    def disable(self):
        self.names = frozenset()

    def start(self, name):
        if name not in self.names:
            return None
        kw = {"external": True} if self.in_graph else {}
        a, b = torch.cuda.Event(enable_timing=True, **kw), torch.cuda.Event(enable_timing=True, **kw)
        a.record()
        self.events.setdefault(name, []).append((a, b))
        return b

    def _stats(self, per_name):
        return {name: {"n": len(ts), "total_ms": sum(ts), "mean_ms": sum(ts) / max(len(ts), 1)} for name, ts in per_name.items()}

    def collect(self):
        torch.cuda.synchronize()
        return self._stats({name: [a.elapsed_time(b) for a, b in pairs] for name, pairs in self.events.items()})

    def collect_replay(self, acc=None):
        """After one replay of the instrumented graph (+ synchronize): append this replay's durations to ``acc``."""
        torch.cuda.synchronize()
        acc = {} if acc is None else acc
        for name, pairs in self.events.items():
            acc.setdefault(name, []).extend(a.elapsed_time(b) for a, b in pairs)
        return acc

    def stats(self, acc):
        return self._stats(acc)


KERNEL_TIMERS = _KernelTimers()


# NVTX ranges (the reference has no tracing at all): GEOSSL_NVTX=1 (or ops.NVTX = True) brackets every
# named C-ABI call and the phases of the training step (pretrain.train_step) with torch.cuda.nvtx ranges, which Nsight
# Systems / Compute show next to the kernels.  Off by default: one flag test per call.
NVTX = os.environ.get("GEOSSL_NVTX", "0") not in ("", "0")


@contextlib.contextmanager
def nvtx_range(name):
    if not NVTX:
        yield
        return
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


def _timed(name, rc_fn):
    """Run ``rc_fn()`` (a C-ABI call returning rc) between two events if ``name`` is being timed."""
    end = KERNEL_TIMERS.start(name) if KERNEL_TIMERS.names else None
    if NVTX:
        torch.cuda.nvtx.range_push("geossl/" + name)
    rc = rc_fn()
    if NVTX:
        torch.cuda.nvtx.range_pop()
    if end is not None:
        end.record()
    check(rc, name)


def _stream(device=None):
    """torch's current stream on ``device`` (default: the current device).  Callers run under
    ``torch.cuda.device_of(tensor)`` semantics: one process drives one GPU."""
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _req(t, dtype, name, dim=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"geossl_b200: `{name}` must be a CUDA tensor (no CPU fallback by design)")
    if t.dtype != dtype:
        raise RuntimeError(f"geossl_b200: `{name}` must be {dtype}, got {t.dtype}")
    if dim is not None and t.dim() != dim:
        raise RuntimeError(f"geossl_b200: `{name}` must be {dim}-D, got shape {tuple(t.shape)}")
    return t.contiguous()


# =====================================================================================================
# graph construction
# =====================================================================================================
class RadiusCSR:
    """Destination-sorted CSR of the radius graph (+ its source-sorted transpose), int32, capacity sized.

    ``rowptr[n_atoms]`` is the device-resident edge count E; ``src``/``tgt``/``dist`` have ``capacity``
    slots of which the first E are valid.  ``edge_index`` / ``num_edges`` synchronise with the host (the
    reference's radius_graph does too); nothing on the training path needs them.
    """

    def __init__(self, n_atoms, capacity, rowptr, src, tgt, dist, batch, graph_ptr):
        self.n_atoms, self.capacity = n_atoms, capacity
        self.rowptr, self.src, self.tgt, self.dist = rowptr, src, tgt, dist
        self.batch, self.graph_ptr = batch, graph_ptr
        self.t_rowptr = self.t_eid = self.t_tgt = None
        self.pair_rowptr = self.pair_of_edge = self.pair_e1 = self.pair_e2 = self.pair_atoms = self.pair_dist = None
        self._n_edges = None
        self._exact = None
        # host-known bound on the atoms of any graph that has edges (None = unknown): what lets the pair-centric cfconv
        # kernel (geossl_cfconv_pairs) size its shared-memory window without a device sync
        self.max_graph_atoms = None

    @property
    def n_edges_dev(self):
        return self.rowptr[self.n_atoms:]

    @property
    def num_edges(self):
        if self._n_edges is None:
            self._n_edges = int(self.rowptr[self.n_atoms].item())
            if self._n_edges > self.capacity:
                raise RuntimeError(f"radius graph has {self._n_edges} edges but capacity is {self.capacity}")
        return self._n_edges

    @property
    def edge_index(self):
        """(2,E) int64 ``[source; target]`` exactly as torch_geometric.nn.radius_graph returns it."""
        e = self.num_edges
        out = torch.empty((2, e), dtype=torch.int64, device=self.rowptr.device)
        check(_lib.load().geossl_csr_to_edge_index(_p(self.src), _p(self.tgt), e, _p(out), _stream()), "edge_index")
        return out

    def ensure_transpose(self):
        if self.t_rowptr is None:
            dev = self.rowptr.device
            n = self.n_atoms
            self.t_rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
            self.t_eid = torch.empty(self.capacity, dtype=torch.int32, device=dev)
            self.t_tgt = torch.empty(self.capacity, dtype=torch.int32, device=dev)
            scratch = torch.empty(n + 2, dtype=torch.int32, device=dev)
            check(_lib.load().geossl_csr_transpose(_p(self.rowptr), _p(self.src), _p(self.batch), _p(self.graph_ptr), n,
                                                   _p(scratch), _p(self.t_rowptr), _p(self.t_eid), _p(self.t_tgt),
                                                   _stream()), "csr_transpose")
        return self

    @property
    def n_pairs_dev(self):
        return self.pair_rowptr[self.n_atoms:]

    def ensure_pairs(self):
        """Undirected-pair index (geossl_pair_index): both directions of a pair share one filter row."""
        if self.pair_rowptr is None:
            dev = self.rowptr.device
            n = self.n_atoms
            self.pair_rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
            self.pair_of_edge = torch.empty(self.capacity, dtype=torch.int32, device=dev)
            self.pair_e1 = torch.empty(self.capacity, dtype=torch.int32, device=dev)
            self.pair_e2 = torch.empty(self.capacity, dtype=torch.int32, device=dev)
            self.pair_dist = torch.empty(self.capacity, dtype=torch.float32, device=dev)
            self.pair_atoms = torch.empty((self.capacity, 2), dtype=torch.int32, device=dev)
            scratch = torch.empty(n + 1, dtype=torch.int32, device=dev)
            check(_lib.load().geossl_pair_index(_p(self.rowptr), _p(self.src), _p(self.dist), n, _p(scratch),
                                                1 if PAIR_OWNER_SMALL else 0, _p(self.pair_rowptr),
                                                _p(self.pair_of_edge), _p(self.pair_e1), _p(self.pair_e2), _p(self.pair_atoms),
                                                _p(self.pair_dist), _stream()), "pair_index")
        return self

    @property
    def num_pairs(self):
        """Host-side count of undirected pairs (one sync, cached) -- the composed double-backward path sizes its per-pair
        tensors with it, like ``num_edges`` for the per-edge form."""
        if getattr(self, "_n_pairs", None) is None:
            self.ensure_pairs()
            self._n_pairs = int(self.pair_rowptr[self.n_atoms].item())
        return self._n_pairs

    def sync_counts(self):
        """Edge count (and pair count, if the pair index exists) with ONE host read."""
        if self.pair_rowptr is not None and (self._n_edges is None or getattr(self, "_n_pairs", None) is None):
            e, u = torch.stack([self.rowptr[self.n_atoms], self.pair_rowptr[self.n_atoms]]).tolist()
            if e > self.capacity:
                raise RuntimeError(f"radius graph has {e} edges but capacity is {self.capacity}")
            self._n_edges, self._n_pairs = int(e), int(u)
        return self.num_edges

    def exact(self):
        """Copy trimmed to exactly E edges (host sync) -- used by the general double-backward path.  A pair index built
        before the call is carried over (trimmed to E rows as well)."""
        if self._exact is None:
            e = self.sync_counts()
            self.ensure_transpose()
            g = RadiusCSR(self.n_atoms, e, self.rowptr, self.src[:e].contiguous(), self.tgt[:e].contiguous(),
                          None if self.dist is None else self.dist[:e].contiguous(), self.batch, self.graph_ptr)
            g.t_rowptr, g.t_eid, g.t_tgt = self.t_rowptr, self.t_eid[:e].contiguous(), self.t_tgt[:e].contiguous()
            g._n_edges = e
            if self.pair_rowptr is not None:
                g.pair_rowptr, g._n_pairs = self.pair_rowptr, self._n_pairs
                g.pair_of_edge, g.pair_atoms = self.pair_of_edge[:e].contiguous(), self.pair_atoms[:e].contiguous()
                g.pair_e1, g.pair_e2, g.pair_dist = self.pair_e1[:e], self.pair_e2[:e], self.pair_dist[:e]
            g.max_graph_atoms = self.max_graph_atoms
            g._exact = g
            self._exact = g
        return self._exact


def graph_ptr_from_batch(batch, num_graphs=None):
    batch = _req(batch, torch.int64, "batch", 1)
    n = batch.numel()
    if num_graphs is None:
        num_graphs = int(batch[-1].item()) + 1 if n else 0     # same sync as dataloaders_AtomTuple.py:78
    ptr = torch.empty(num_graphs + 1, dtype=torch.int32, device=batch.device)
    check(_lib.load().geossl_graph_ptr(_p(batch), n, num_graphs, _p(ptr), _stream()), "graph_ptr")
    return ptr


def super_edges(graph_ptr, pair_ptr, n_graphs, n_atoms, n_pairs, permutation=False):
    """(super_edge_index (2,P) int64, batch (N,) int64) from the atom / pair offsets (see geossl_super_edges)."""
    graph_ptr, pair_ptr = _req(graph_ptr, torch.int32, "graph_ptr", 1), _req(pair_ptr, torch.int64, "pair_ptr", 1)
    sei = torch.empty((2, n_pairs), dtype=torch.int64, device=graph_ptr.device)
    batch = torch.empty(n_atoms, dtype=torch.int64, device=graph_ptr.device)
    check(_lib.load().geossl_super_edges(_p(graph_ptr), _p(pair_ptr), n_graphs, n_atoms, 1 if permutation else 0, n_pairs,
                                         _p(sei), _p(batch), _stream()), "super_edges")
    return sei, batch


# Distance rounding of the neighbour search: False = ((dx*dx)+dy*dy)+dz*dz rounded per operation (the survey's reading of the
# torch_cluster kernel), True = the FMA-contracted form nvcc's default -fmad=true would emit for it.  The third-party
# binary the reference ran is not available, so this stays a documented switch (tests enumerate where the two differ).
RADIUS_FMA = False
# Mean atoms per graph from which the neighbour search builds a spatial cell list (bit-identical output; the index-order
# scan is faster for molecules and ~600-atom pockets).  None disables the cell list.
CELL_LIST_MIN_ATOMS = 2048


def radius_csr(pos, batch, r, max_num_neighbors=MAX_NEIGHBORS_DEFAULT, *, graph_ptr=None, num_graphs=None,
               capacity=None, transpose=True, cell_list=None, max_graph_atoms=None):
    """Neighbour search -> RadiusCSR (torch_cluster.radius_graph semantics, see include/geossl_b200.h).
    ``cell_list``: None = automatic (graphs of at least CELL_LIST_MIN_ATOMS atoms on average), True / False force it."""
    pos = _req(pos.detach(), torch.float32, "pos", 2)
    if pos.size(1) != 3:
        raise RuntimeError("geossl_b200: `pos` must be (N,3)")
    n = pos.size(0)
    if batch is None:
        batch = torch.zeros(n, dtype=torch.int64, device=pos.device)
        num_graphs = 1 if n else 0
    batch = _req(batch, torch.int64, "batch", 1)
    if graph_ptr is None:
        graph_ptr = graph_ptr_from_batch(batch, num_graphs)
    graph_ptr = _req(graph_ptr, torch.int32, "graph_ptr", 1)
    if capacity is None:
        capacity = (max_num_neighbors + 1) * n
    dev = pos.device
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    src = torch.empty(capacity, dtype=torch.int32, device=dev)
    tgt = torch.empty(capacity, dtype=torch.int32, device=dev)
    dist = torch.empty(capacity, dtype=torch.float32, device=dev)
    scratch = torch.empty(2 * n + 2, dtype=torch.int32, device=dev)
    n_graphs = graph_ptr.numel() - 1
    if cell_list is None:
        cell_list = CELL_LIST_MIN_ATOMS is not None and n_graphs > 0 and n >= CELL_LIST_MIN_ATOMS * n_graphs
    box = keys = order = None
    min_atoms = 0
    lib = _lib.load()
    if cell_list and n > 0 and n_graphs > 0:
        box = torch.empty((n_graphs, 8), dtype=torch.float32, device=dev)
        keys = torch.empty(n, dtype=torch.int64, device=dev)
        check(lib.geossl_radius_cell_keys(_p(pos), _p(batch), _p(graph_ptr), n, n_graphs, float(r), _p(box), _p(keys), _stream()),
              "radius_cell_keys")
        keys, order = torch.sort(keys, stable=True)          # library radix sort: atoms of one cell stay in ascending index
        min_atoms = 0 if cell_list is True else int(CELL_LIST_MIN_ATOMS or 0)
    _timed("radius_csr", lambda: lib.geossl_radius_csr(
        _p(pos), _p(batch), _p(graph_ptr), n, float(r), int(max_num_neighbors), capacity, _p(scratch), _p(rowptr), _p(src),
        _p(tgt), _p(dist), 1 if RADIUS_FMA else 0, _p(box), _p(keys), _p(order), min_atoms, _stream()))
    g = RadiusCSR(n, capacity, rowptr, src, tgt, dist, batch, graph_ptr)
    g.max_graph_atoms = None if max_graph_atoms is None else int(max_graph_atoms)
    if transpose and not (max_graph_atoms is not None and _pairs_kernel_applies(g)):
        g.ensure_transpose()              # the pair-centric cfconv kernels never read the source-sorted view
    return g


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=MAX_NEIGHBORS_DEFAULT, flow="source_to_target"):
    """Drop-in for ``torch_geometric.nn.radius_graph`` (the signature schnet.py:91 uses)."""
    if loop or flow != "source_to_target":
        raise NotImplementedError("only loop=False, flow='source_to_target' (the reference's call) is built")
    return radius_csr(x, batch, r, max_num_neighbors, transpose=False).edge_index


def csr_from_edge_index(edge_index, n_atoms, batch, graph_ptr=None, num_graphs=None, dist=None):
    """RadiusCSR view of a (2,E) ``[source; target]`` list that is sorted by target with ascending sources
    inside a target (what radius_graph emits).  Used by the stand-alone CFConv / InteractionBlock API and
    by PaiNN's precomputed ``radius_edge_index``."""
    edge_index = _req(edge_index, torch.int64, "edge_index", 2)
    e = edge_index.size(1)
    dev = edge_index.device
    batch = _req(batch, torch.int64, "batch", 1)
    if graph_ptr is None:
        graph_ptr = graph_ptr_from_batch(batch, num_graphs)
    rowptr = torch.empty(n_atoms + 1, dtype=torch.int32, device=dev)
    check(_lib.load().geossl_rowptr_from_sorted(_p(edge_index[1].contiguous()), e, n_atoms, _p(rowptr), _stream()), "rowptr")
    g = RadiusCSR(n_atoms, e, rowptr, edge_index[0].to(torch.int32), edge_index[1].to(torch.int32), dist, batch, graph_ptr)
    g._n_edges = e
    return g


# =====================================================================================================
# cfconv primitives (closed under differentiation: A, A^T and the edge product)
# =====================================================================================================
def _cfconv_fwd(x, filt, g, filt_row=None):
    out = torch.empty((g.n_atoms, x.size(1)), dtype=torch.float32, device=x.device)
    _timed("cfconv_fwd", lambda: _lib.load().geossl_cfconv_fwd(_p(x), _p(filt), _p(filt_row), _p(g.rowptr), _p(g.src), g.n_atoms,
                                                               x.size(1), _p(out), _stream()))
    return out


def _cfconv_bwd_x(filt, grad_out, g, filt_row=None):
    g.ensure_transpose()
    dx = torch.empty((g.n_atoms, grad_out.size(1)), dtype=torch.float32, device=grad_out.device)
    _timed("cfconv_bwd_x", lambda: _lib.load().geossl_cfconv_bwd_x(_p(filt), _p(filt_row), _p(grad_out), _p(g.t_rowptr),
                                                                   _p(g.t_eid), _p(g.t_tgt), g.n_atoms, grad_out.size(1), _p(dx),
                                                                   _stream()))
    return dx


# Pair-centric cfconv (geossl_cfconv_pairs): used by the fused layer when filter rows are shared per atom pair and the batch
# carries a host-known bound on its graph sizes that fits the shared-memory window.
CFCONV_PAIRS = True
CFCONV_PAIRS_MAX_ATOMS = 32         # one CTA walks a graph's n(n-1)/2 pairs with two warps: larger graphs take the row-gather kernels
CFCONV_PAIRS_TUNING = 0           # 0 = library default; groups * 100 + unroll (profiles/bench_cfconv.py sweeps it)


def _pairs_kernel_applies(g):
    return (CFCONV_PAIRS and SHARE_PAIR_FILTERS and FILTER_MODE != "simt" and g.max_graph_atoms is not None
            and 1 <= g.max_graph_atoms <= CFCONV_PAIRS_MAX_ATOMS and g.dist is not None)


def _cfconv_pairs(v, filt, g, transposed):
    g.ensure_pairs()
    out = torch.empty((g.n_atoms, 128), dtype=torch.float32, device=v.device)
    _timed("cfconv_bwd_x" if transposed else "cfconv_fwd", lambda: _lib.load().geossl_cfconv_pairs(
        _p(v), _p(filt), _p(g.pair_atoms), _p(g.pair_rowptr), _p(g.graph_ptr), g.graph_ptr.numel() - 1,
        int(g.max_graph_atoms), 1 if transposed else 0, int(CFCONV_PAIRS_TUNING), _p(out), _stream()))
    return out


def _cfconv_bwd_w(x, grad_out, g):
    dw = torch.empty((g.capacity, x.size(1)), dtype=torch.float32, device=x.device)
    check(_lib.load().geossl_cfconv_bwd_w(_p(x), _p(grad_out), _p(g.rowptr), _p(g.src), g.n_atoms, x.size(1), _p(dw), _stream()),
          "cfconv_bwd_w")
    return dw


class CFConvAggregate(torch.autograd.Function):
    """m_i = sum_{e in row i} x[src_e] * W_e  -- differentiable to any order (schnet.py:190,194-195)."""

    @staticmethod
    def forward(ctx, x, filt, graph):
        x, filt = _req(x, torch.float32, "x", 2), _req(filt, torch.float32, "filt", 2)
        ctx.graph = graph
        ctx.save_for_backward(x, filt)
        return _cfconv_fwd(x, filt, graph)

    @staticmethod
    def backward(ctx, grad_out):
        x, filt = ctx.saved_tensors
        gx = CFConvAggregateT.apply(filt, grad_out, ctx.graph) if ctx.needs_input_grad[0] else None
        gw = CFConvEdgeProduct.apply(x, grad_out, ctx.graph) if ctx.needs_input_grad[1] else None
        return gx, gw, None


class CFConvAggregateT(torch.autograd.Function):
    """dx_j = sum_{e: src_e = j} W_e * g[tgt_e]."""

    @staticmethod
    def forward(ctx, filt, grad_out, graph):
        filt, grad_out = _req(filt, torch.float32, "filt", 2), _req(grad_out, torch.float32, "grad_out", 2)
        ctx.graph = graph
        ctx.save_for_backward(filt, grad_out)
        return _cfconv_bwd_x(filt, grad_out, graph)

    @staticmethod
    def backward(ctx, v):
        filt, grad_out = ctx.saved_tensors
        gw = CFConvEdgeProduct.apply(v, grad_out, ctx.graph) if ctx.needs_input_grad[0] else None
        gg = CFConvAggregate.apply(v, filt, ctx.graph) if ctx.needs_input_grad[1] else None
        return gw, gg, None


class CFConvEdgeProduct(torch.autograd.Function):
    """dW_e = x[src_e] * g[tgt_e]  (E,F)."""

    @staticmethod
    def forward(ctx, x, grad_out, graph):
        x, grad_out = _req(x, torch.float32, "x", 2), _req(grad_out, torch.float32, "grad_out", 2)
        ctx.graph = graph
        ctx.save_for_backward(x, grad_out)
        return _cfconv_bwd_w(x, grad_out, graph)

    @staticmethod
    def backward(ctx, v):
        x, grad_out = ctx.saved_tensors
        gx = CFConvAggregateT.apply(v, grad_out, ctx.graph) if ctx.needs_input_grad[0] else None
        gg = CFConvAggregate.apply(x, v, ctx.graph) if ctx.needs_input_grad[1] else None
        return gx, gg, None


# ---- the same three primitives with ONE filter row per atom pair (rows indexed through graph.pair_of_edge): the composed
# double-backward path then builds and differentiates U = E/2 filter rows instead of E (COMPOSED_PAIRS)
COMPOSED_PAIRS = True


class CFConvAggregateP(torch.autograd.Function):
    """m_i = sum_{e in row i} x[src_e] * W[pair(e)]  with W (U,F), differentiable to any order."""

    @staticmethod
    def forward(ctx, x, filt, graph):
        x, filt = _req(x, torch.float32, "x", 2), _req(filt, torch.float32, "filt", 2)
        ctx.graph = graph
        ctx.save_for_backward(x, filt)
        return _cfconv_fwd(x, filt, graph, graph.pair_of_edge)

    @staticmethod
    def backward(ctx, grad_out):
        x, filt = ctx.saved_tensors
        gx = CFConvAggregateTP.apply(filt, grad_out, ctx.graph) if ctx.needs_input_grad[0] else None
        gw = CFConvPairProduct.apply(x, grad_out, ctx.graph) if ctx.needs_input_grad[1] else None
        return gx, gw, None


class CFConvAggregateTP(torch.autograd.Function):
    """dx_j = sum_{e: src_e = j} W[pair(e)] * g[tgt_e]."""

    @staticmethod
    def forward(ctx, filt, grad_out, graph):
        filt, grad_out = _req(filt, torch.float32, "filt", 2), _req(grad_out, torch.float32, "grad_out", 2)
        ctx.graph = graph
        ctx.save_for_backward(filt, grad_out)
        return _cfconv_bwd_x(filt, grad_out, graph, graph.pair_of_edge)

    @staticmethod
    def backward(ctx, v):
        filt, grad_out = ctx.saved_tensors
        gw = CFConvPairProduct.apply(v, grad_out, ctx.graph) if ctx.needs_input_grad[0] else None
        gg = CFConvAggregateP.apply(v, filt, ctx.graph) if ctx.needs_input_grad[1] else None
        return gw, gg, None


class CFConvPairProduct(torch.autograd.Function):
    """dW_u = x[s] * g[t] + x[t] * g[s]  (U,F)  (second term only where the reverse edge exists)."""

    @staticmethod
    def forward(ctx, x, grad_out, graph):
        x, grad_out = _req(x, torch.float32, "x", 2), _req(grad_out, torch.float32, "grad_out", 2)
        ctx.graph = graph
        ctx.save_for_backward(x, grad_out)
        u = graph.num_pairs
        dw = torch.empty((u, x.size(1)), dtype=torch.float32, device=x.device)
        check(_lib.load().geossl_cfconv_pair_product(_p(x), _p(grad_out), _p(graph.pair_atoms), u, x.size(1), _p(dw), _stream()),
              "cfconv_pair_product")
        return dw

    @staticmethod
    def backward(ctx, v):
        x, grad_out = ctx.saved_tensors
        gx = CFConvAggregateTP.apply(v, grad_out, ctx.graph) if ctx.needs_input_grad[0] else None
        gg = CFConvAggregateP.apply(x, v, ctx.graph) if ctx.needs_input_grad[1] else None
        return gx, gg, None


def pair_endpoints(graph):
    """(s, t) int64 atom ids of the graph's undirected pairs, (U,) each -- the composed path's per-pair geometry."""
    pa = graph.pair_atoms[:graph.num_pairs]
    t = pa[:, 1]
    return pa[:, 0].long(), torch.where(t < 0, torch.bitwise_not(t), t).long()


# =====================================================================================================
# fused interaction filter + aggregate (training fast path, first order)
# =====================================================================================================
# Which kernel computes the filter network forward: "tc_fp16" / "tc_bf16" = tcgen05 tensor cores with fp32
# operands split into two fp16 / bf16 parts (three MMAs per product, fp32 accumulate; needs F = 128),
# "simt" = fp32 CUDA cores (any supported width; also what narrower models use).
FILTER_MODE = "tc_fp16"
# The filter of an edge depends only on its length, so the two directions of an atom pair share one filter row
# (geossl_pair_index): the tensor-core filter kernels then run over U <= E undirected pairs -- E/2 when no neighbour row
# is truncated.  Forward results are bit-identical to the per-edge form; parameter gradients differ by summation order.
# Applies to the tensor-core modes; "simt" stays per edge (the exact fp32 cross-check).
SHARE_PAIR_FILTERS = True
# Which row owns a pair's filter row: the larger atom (False) or the smaller atom (True: rows walked in ascending order
# would read their own block first).  Measured on the bench workload: no difference for the cfconv kernels (47.8 vs
# 48.5 us), so the default keeps the cheaper index (profiles/README.md).
PAIR_OWNER_SMALL = False


def filter_forward(graph, offset, coeff, cutoff, w1, b1, w2, b2, mode=None, pairs=False):
    """W (capacity,F) from the edge distances held by ``graph`` (fused rbf + filter MLP + cutoff); one row per directed
    edge, or per undirected pair with ``pairs=True`` (index rows through ``graph.pair_of_edge``)."""
    F_, G = w1.size(0), w1.size(1)
    filt = torch.empty((graph.capacity, F_), dtype=torch.float32, device=w1.device)
    mode = mode or FILTER_MODE
    dist, count = (graph.ensure_pairs().pair_dist, graph.n_pairs_dev) if pairs else (graph.dist, graph.n_edges_dev)
    if mode != "simt" and F_ == 128:
        _timed("filter_fwd", lambda: _lib.load().geossl_filter_fwd_tc(
            _p(dist), _p(count), graph.capacity, _p(offset), float(coeff), float(cutoff), G, F_, _p(w1),
            _p(b1), _p(w2), _p(b2), _p(filt), 1 if mode == "tc_bf16" else 0, _stream()))
        return filt
    _timed("filter_fwd", lambda: _lib.load().geossl_filter_fwd(_p(dist), _p(count), graph.capacity, _p(offset),
                                                               float(coeff), float(cutoff), G, F_, _p(w1), _p(b1), _p(w2), _p(b2),
                                                               _p(filt), _stream()))
    return filt


class CFConvLayer(torch.autograd.Function):
    """x (N,F), filter-network parameters, graph -> m (N,F):  rbf -> Lin1 -> ssp -> Lin2 -> *cutoff -> cfconv.

    forward  : geossl_filter_fwd + geossl_cfconv_fwd (W_e materialised once, kept for the backward)
    backward : geossl_cfconv_bwd_x + geossl_filter_bwd (dW_e = x[src]*g[tgt] formed on the fly)
    First-order only; force training (positions require grad) takes the composable path in SchNet.
    """

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, offset, graph, coeff, cutoff):
        x = _req(x, torch.float32, "x", 2)
        w1, b1, w2, b2 = (_req(t, torch.float32, n) for t, n in ((w1, "w1"), (b1, "b1"), (w2, "w2"), (b2, "b2")))
        offset = _req(offset, torch.float32, "offset", 1)
        pairs = (SHARE_PAIR_FILTERS and FILTER_MODE != "simt" and w1.size(0) == 128 and w1.size(1) <= 63
                 and graph.dist is not None)
        filt = filter_forward(graph, offset, coeff, cutoff, w1, b1, w2, b2, pairs=pairs)
        ctx.pair_kernel = pairs and x.size(1) == 128 and _pairs_kernel_applies(graph)
        if ctx.pair_kernel:
            out = _cfconv_pairs(x, filt, graph, False)
        else:
            out = _cfconv_fwd(x, filt, graph, graph.pair_of_edge if pairs else None)
        ctx.graph, ctx.coeff, ctx.cutoff, ctx.mode, ctx.pairs = graph, coeff, cutoff, FILTER_MODE, pairs
        ctx.save_for_backward(x, filt, w1, b1, w2, b2, offset)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        x, filt, w1, b1, w2, b2, offset = ctx.saved_tensors
        g = ctx.graph
        grad_out = grad_out.contiguous()
        lib = _lib.load()
        F_, G = w1.size(0), w1.size(1)
        gx = None
        if ctx.needs_input_grad[0]:
            gx = (_cfconv_pairs(grad_out, filt, g, True) if ctx.pair_kernel
                  else _cfconv_bwd_x(filt, grad_out, g, g.pair_of_edge if ctx.pairs else None))
        gw1, gb1, gw2, gb2 = torch.empty_like(w1), torch.empty_like(b1), torch.empty_like(w2), torch.empty_like(b2)
        if ctx.mode != "simt" and F_ == 128 and G <= 63:
            ws = torch.empty(lib.geossl_filter_bwd_tc_workspace(), dtype=torch.float32, device=x.device)
            dist, count = (g.pair_dist, g.n_pairs_dev) if ctx.pairs else (g.dist, g.n_edges_dev)
            pa = g.pair_atoms if ctx.pairs else None
            _timed("filter_bwd", lambda: lib.geossl_filter_bwd_tc(
                _p(dist), _p(count), g.capacity, _p(offset), float(ctx.coeff), float(ctx.cutoff), G, F_, _p(w1), _p(b1),
                _p(w2), _p(x), _p(grad_out), _p(g.src), _p(g.tgt), _p(pa), _p(ws), _p(gw1), _p(gb1), _p(gw2), _p(gb2),
                _stream()))
            return gx, gw1, gb1, gw2, gb2, None, None, None, None
        gw1, gb1, gw2, gb2 = torch.empty_like(w1), torch.empty_like(b1), torch.empty_like(w2), torch.empty_like(b2)
        ws = torch.empty(lib.geossl_filter_bwd_workspace(G, F_), dtype=torch.float32, device=x.device)
        _timed("filter_bwd", lambda: lib.geossl_filter_bwd(
            _p(g.dist), _p(g.n_edges_dev), g.capacity, _p(offset), float(ctx.coeff), float(ctx.cutoff), G, F_, _p(w1), _p(b1),
            _p(w2), _p(b2), _p(x), _p(grad_out), _p(g.src), _p(g.tgt), None, _p(ws), _p(gw1), _p(gb1), _p(gw2), _p(gb2), _stream()))
        return gx, gw1, gb1, gw2, gb2, None, None, None, None


# =====================================================================================================
# atom-wise dense layers on the tensor cores
# =====================================================================================================
def _pack_weight(weight, transpose, bf16_parts):
    lib = _lib.load()
    image = torch.empty(lib.geossl_weight_image_bytes(), dtype=torch.uint8, device=weight.device)
    check(lib.geossl_pack_weight(_p(weight), 1 if transpose else 0, 1 if bf16_parts else 0, _p(image), _stream()), "pack_weight")
    return image


def _linear_tc(x, weight, transpose, bias, pre_ssp, act_grad_input, residual, bf16_parts, name, image=None):
    y = torch.empty((x.size(0), 128), dtype=torch.float32, device=x.device)
    if image is None:
        image = _pack_weight(weight, transpose, bf16_parts)
    _timed(name, lambda: _lib.load().geossl_linear_tc(_p(x), x.size(0), _p(image), _p(bias), 1 if pre_ssp else 0,
                                                      _p(act_grad_input), _p(residual), _p(y), 1 if bf16_parts else 0, _stream()))
    return y


# Weight-gradient kernels are off the critical path of the backward chain (nothing upstream consumes them), and they are
# small, latency-bound launches.  Inside ``side_stream_wgrads()`` they are issued on a second stream so that they fill the
# gaps of the main chain; ``join_side_stream()`` must run before anything reads the gradients (optimizer, all-reduce).
# Autograd knows nothing about that second stream (it would sum two uses of a weight, or add into an existing
# ``p.grad``, on the main stream while the side-stream kernel is still writing), so these gradients never pass through
# autograd: backward returns None for them, the kernels write into private buffers, and ``join_side_stream()`` -- after
# the main stream has waited for the side stream -- assigns / accumulates them into ``p.grad`` in launch order.
# Off by default: a caller that runs loss.backward(); optimizer.step() itself never sees a second stream.
_SIDE = {"on": False, "stream": {}, "dirty": False, "pending": []}


class side_stream_wgrads:
    def __enter__(self):
        self.prev = _SIDE["on"]
        _SIDE["on"] = True

    def __exit__(self, *exc):
        join_side_stream()
        _SIDE["on"] = self.prev


def _side_stream(device):
    st = _SIDE["stream"].get(device)
    if st is None:
        st = _SIDE["stream"][device] = torch.cuda.Stream(device=device)
    return st


def join_side_stream():
    if _SIDE["dirty"]:
        for dev, st in _SIDE["stream"].items():
            torch.cuda.current_stream(dev).wait_stream(st)
        _SIDE["dirty"] = False
    pending, _SIDE["pending"] = _SIDE["pending"], []
    for param, buf in pending:                      # main stream, after the wait: plain stream-ordered accumulation
        if param.grad is None:
            param.grad = buf
        else:
            param.grad.add_(buf)


def _linear_wgrad(gy, x, pre_ssp, wp, bp):
    """Weight / bias gradient of one 128 -> 128 layer (geossl_linear_wgrad_tc): dW = gy^T [ssp](x), db = colsum(gy).
    Inside ``side_stream_wgrads()`` (and for leaf parameters) the kernels run on the side stream and the results are handed to
    ``join_side_stream()`` instead of autograd: returns (None, None) then."""
    lib = _lib.load()
    n = x.size(0)

    def wgrad():
        gw_ = torch.empty((128, 128), dtype=torch.float32, device=x.device)
        gb_ = torch.empty(128, dtype=torch.float32, device=x.device) if bp is not None else None
        ws = torch.empty(lib.geossl_linear_wgrad_tc_workspace(n), dtype=torch.float32, device=x.device)
        _timed("linear_wgrad", lambda: lib.geossl_linear_wgrad_tc(_p(gy), _p(x), n, 1 if pre_ssp else 0, _p(ws), _p(gw_),
                                                                  _p(gb_), _stream()))
        return gw_, gb_

    deferrable = wp.is_leaf and wp.requires_grad and (bp is None or (bp.is_leaf and bp.requires_grad))
    if _SIDE["on"] and deferrable:
        main, side = torch.cuda.current_stream(x.device), _side_stream(x.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            gw, gb = wgrad()
        for t in (gy, x):
            t.record_stream(side)
        for t in (gw, gb):
            if t is not None:
                t.record_stream(main)
        _SIDE["dirty"] = True
        _SIDE["pending"].append((wp, gw))
        if gb is not None:
            _SIDE["pending"].append((bp, gb))
        return None, None                  # handed to join_side_stream(), not to autograd
    return wgrad()


def _wgrad_batch(n, problems, name="linear_wgrad"):
    """One geossl_linear_wgrad_tc_batch launch (+ one reduction) for up to six weight-gradient products over the same ``n``
    rows.  problems: dicts with gy=(tensor, column), x=(tensor, column), pre_act, x_cols, gw=(tensor, element offset, row
    stride), gb=(tensor, element offset) or None."""
    lib = _lib.load()
    ws_floats = lib.geossl_linear_wgrad_tc_workspace(n)
    dev = problems[0]["gy"][0].device
    ws = torch.empty(ws_floats * len(problems), dtype=torch.float32, device=dev)
    for i0 in range(0, len(problems), 6):
        chunk = problems[i0:i0 + 6]
        arr = (_lib.WgradProblem * len(chunk))()
        for k, (a, pr) in enumerate(zip(arr, chunk)):
            gy, gyc = pr["gy"]
            x, xc = pr["x"]
            gw, gwo, ldgw = pr["gw"]
            a.grad_y, a.ld_dy = gy.data_ptr() + 4 * gyc, gy.stride(0)
            a.x, a.ld_x = x.data_ptr() + 4 * xc, x.stride(0)
            a.workspace = ws.data_ptr() + 4 * ws_floats * (i0 + k)
            a.grad_weight, a.ld_gw = gw.data_ptr() + 4 * gwo, ldgw
            gb = pr.get("gb")
            a.grad_bias = None if gb is None else gb[0].data_ptr() + 4 * gb[1]
            a.pre_act, a.x_cols = int(pr.get("pre_act", 0)), int(pr.get("x_cols", 128))
        _timed(name, lambda: lib.geossl_linear_wgrad_tc_batch(arr, len(chunk), n, _stream()))


def _linear_wgrads(items):
    """Weight / bias gradients of several 128 -> 128 layers over the same rows in ONE launch: items = [(gy, x, pre_ssp, weight
    param, bias param or None)].  Returns [(gw, gb)] for autograd -- (None, None) for every layer when the launch went to the
    side stream (see ``side_stream_wgrads``)."""
    n = items[0][1].size(0)
    params = []
    for _, _, _, wp, bp in items:
        params += [wp, bp]

    def compute():
        out, probs = [], []
        for gy, x, pre, wp, bp in items:
            gw = torch.empty((128, 128), dtype=torch.float32, device=x.device)
            gb = torch.empty(128, dtype=torch.float32, device=x.device) if bp is not None else None
            probs.append({"gy": (gy, 0), "x": (x, 0), "pre_act": ACT_SSP if pre else ACT_NONE, "gw": (gw, 0, 128),
                          "gb": None if gb is None else (gb, 0)})
            out += [gw, gb]
        _wgrad_batch(n, probs)
        return out

    inputs = [t for gy, x, _, _, _ in items for t in (gy, x)]
    flat = _deferred_wgrads(compute, params, inputs)
    return [(flat[2 * i], flat[2 * i + 1]) for i in range(len(items))]


class LinearTC(torch.autograd.Function):
    """y = [ssp](x) @ W^T + b [+ residual] for 128 -> 128 atom-wise layers (geossl_linear_tc / _wgrad_tc).
    Forward operands are split into fp16 parts, gradient operands into bf16 parts (see tc.cuh)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, pre_ssp, images=None):
        x, weight = _req(x, torch.float32, "x", 2), _req(weight, torch.float32, "weight", 2)
        if x.size(1) != 128 or tuple(weight.shape) != (128, 128):
            raise RuntimeError("geossl_b200: LinearTC is built for 128 -> 128 layers")
        bias = None if bias is None else _req(bias, torch.float32, "bias", 1)
        residual = None if residual is None else _req(residual, torch.float32, "residual", 2)
        ctx.pre_ssp, ctx.has_bias, ctx.has_res = bool(pre_ssp), bias is not None, residual is not None
        ctx.image_t = None if images is None else images[1]
        ctx.params = (weight, bias)             # the leaf objects themselves (deferred side-stream accumulation)
        ctx.save_for_backward(x, weight)
        if x.size(0) == 0:
            return x.new_zeros((0, 128))
        return _linear_tc(x, weight, False, bias, pre_ssp, None, residual, False, "linear_fwd",
                          None if images is None else images[0])

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        n = x.size(0)
        gx = gw = gb = None
        if n == 0:
            return (torch.zeros_like(x), torch.zeros_like(weight), torch.zeros(128, device=x.device) if ctx.has_bias else None,
                    gy if ctx.has_res else None, None, None)
        if ctx.needs_input_grad[0]:
            gx = _linear_tc(gy, weight, True, None, False, x if ctx.pre_ssp else None, None, True, "linear_dgrad", ctx.image_t)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gw, gb = _linear_wgrad(gy, x, ctx.pre_ssp, *ctx.params)
        return gx, gw, gb, (gy if ctx.has_res else None), None, None


# The three atom-wise layers between two cfconv aggregations (conv.lin2 -> ssp -> lin (+ residual) -> next conv.lin1) and the
# head (lin1 -> ssp -> lin2) run as ONE launch each (geossl_linear_chain_tc): the row tile never leaves the SM between the
# GEMMs.  False restores one launch per layer (ops.linear).
FUSE_DENSE_CHAIN = True


def _chain(x, stages, bf16_parts, name):
    """stages: [(weight_image, bias, act_grad_input, residual, store, act_next)]; runs geossl_linear_chain_tc over x (n,128)."""
    arr = (_lib.ChainStage * len(stages))()
    for a, (img, bias, z, res, store, act_next) in zip(arr, stages):
        a.weight_image, a.bias, a.act_grad_input = img.data_ptr(), (bias.data_ptr() if bias is not None else None), \
            (z.data_ptr() if z is not None else None)
        a.residual, a.store, a.act_next = (res.data_ptr() if res is not None else None), \
            (store.data_ptr() if store is not None else None), act_next
    _timed(name, lambda: _lib.load().geossl_linear_chain_tc(_p(x), x.size(0), arr, len(stages), 1 if bf16_parts else 0, ACT_SSP, _stream()))


class InteractionTail(torch.autograd.Function):
    """(m, h) -> (h', x'):  y1 = m W2^T + b2;  h' = h + ssp(y1) W3^T + b3;  x' = h' W1n^T  (x' only when the next block's
    conv.lin1 weight W1n is given) -- schnet.py:191,165-166,97,189 in one launch.  y1 is kept for the backward, which is one
    launch too: g = g_x' W1n + g_h';  g_y1 = (g W3) * sigmoid(y1);  g_m = g_y1 W2;  the weight gradients are three
    geossl_linear_wgrad_tc launches (side stream inside ``side_stream_wgrads()``)."""

    @staticmethod
    def forward(ctx, m, h, w2, b2, w3, b3, w1n, img2, img3, img1n):
        m, h = _req(m, torch.float32, "m", 2), _req(h, torch.float32, "h", 2)
        n = m.size(0)
        y1, h_next = torch.empty_like(m), torch.empty_like(m)
        x_next = torch.empty_like(m) if w1n is not None else None
        stages = [(img2[0], b2, None, None, y1, ACT_SSP), (img3[0], b3, None, h, h_next, ACT_NONE)]
        if w1n is not None:
            stages.append((img1n[0], None, None, None, x_next, ACT_NONE))
        if n:
            _chain(m, stages, False, "dense_chain_fwd")
        ctx.imgs = (img2, img3, img1n)
        ctx.params = ((w2, b2), (w3, b3), (w1n, None))
        ctx.save_for_backward(m, y1, h_next)
        return h_next, x_next

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_hn, g_xn):
        m, y1, h_next = ctx.saved_tensors
        img2, img3, img1n = ctx.imgs
        (w2, b2), (w3, b3), (w1n, _) = ctx.params
        n = m.size(0)
        if g_hn is None:
            g_hn = torch.zeros_like(m)
        g_hn = g_hn.contiguous()
        g_y1, g_m = torch.empty_like(m), torch.empty_like(m)
        if w1n is not None and g_xn is not None:
            g_xn = g_xn.contiguous()
            g_tot = torch.empty_like(m)
            stages = [(img1n[1], None, None, g_hn, g_tot, ACT_NONE), (img3[1], None, y1, None, g_y1, ACT_NONE),
                      (img2[1], None, None, None, g_m, ACT_NONE)]
            x0 = g_xn
        else:
            g_tot = g_hn
            stages = [(img3[1], None, y1, None, g_y1, ACT_NONE), (img2[1], None, None, None, g_m, ACT_NONE)]
            x0 = g_hn
        if n:
            _chain(x0, stages, True, "dense_chain_bwd")
        gw1n = None
        if not n:
            return (g_m, g_tot, torch.zeros_like(w2), torch.zeros_like(b2), torch.zeros_like(w3), torch.zeros_like(b3),
                    None if w1n is None else torch.zeros_like(w1n), None, None, None)
        # three separate launches on purpose: batched into one (ops._linear_wgrads) the step is slower (2.33-2.37 vs 2.22 ms) --
        # the longer kernel holds SMs against the main chain and lengthens the tail that the final stream join waits for
        if w1n is not None and g_xn is not None:
            gw1n, _ = _linear_wgrad(g_xn, h_next, False, w1n, None)
        gw3, gb3 = _linear_wgrad(g_tot, y1, True, w3, b3)
        gw2, gb2 = _linear_wgrad(g_y1, m, False, w2, b2)
        return g_m, g_tot, gw2, gb2, gw3, gb3, gw1n, None, None, None


class HeadChain(torch.autograd.Function):
    """out = ssp(h W1^T + b1) W2^T + b2  (schnet.py:99-101) in one launch; backward: g_z = (g W2) * sigmoid(z), g_h = g_z W1."""

    @staticmethod
    def forward(ctx, h, w1, b1, w2, b2, img1, img2):
        h = _req(h, torch.float32, "h", 2)
        z, out = torch.empty_like(h), torch.empty_like(h)
        if h.size(0):
            _chain(h, [(img1[0], b1, None, None, z, ACT_SSP), (img2[0], b2, None, None, out, ACT_NONE)], False, "dense_chain_fwd")
        ctx.imgs, ctx.params = (img1, img2), ((w1, b1), (w2, b2))
        ctx.save_for_backward(h, z)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        h, z = ctx.saved_tensors
        img1, img2 = ctx.imgs
        (w1, b1), (w2, b2) = ctx.params
        g = g.contiguous()
        g_z, g_h = torch.empty_like(h), torch.empty_like(h)
        if not h.size(0):
            return g_h, torch.zeros_like(w1), torch.zeros_like(b1), torch.zeros_like(w2), torch.zeros_like(b2), None, None
        _chain(g, [(img2[1], None, z, None, g_z, ACT_NONE), (img1[1], None, None, None, g_h, ACT_NONE)], True, "dense_chain_bwd")
        gw2, gb2 = _linear_wgrad(g, z, True, w2, b2)
        gw1, gb1 = _linear_wgrad(g_z, h, False, w1, b1)
        return g_h, gw1, gb1, gw2, gb2, None, None


def chain_applies(layers, images):
    """True when every layer of the list has packed tensor-core images (128 -> 128, tensor-core mode) and fusion is on."""
    return bool(FUSE_DENSE_CHAIN and images) and all(l in images for l in layers)


def linear_tc_applies(layer):
    w = layer.weight
    return FILTER_MODE != "simt" and w.is_cuda and tuple(w.shape) == (128, 128)


_ptr_tables = {}


def prepack_linear_weights(layers):
    """Pack the operand images of several 128x128 layers (forward orientation in fp16 parts, transposed orientation in
    bf16 parts for the data gradient) with ONE launch (geossl_pack_weights_batched).  The device table of weight
    addresses is built once per set of layers (parameter storage does not move).  Returns {layer: (image, image_t)};
    valid for the current forward/backward only (weights change at the next optimizer step)."""
    layers = [l for l in layers if linear_tc_applies(l)]
    if not layers:
        return {}
    dev = layers[0].weight.device
    ws = [l.weight.detach() for l in layers]
    for w in ws:
        _req(w, torch.float32, "weight", 2)
    key = (dev, tuple(w.data_ptr() for w in ws))
    table = _ptr_tables.get(key)
    if table is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("geossl_b200: run the model once before capturing it in a CUDA graph (weight pointer table)")
        if len(_ptr_tables) > 16:
            _ptr_tables.clear()
        table = _ptr_tables[key] = torch.tensor(key[1], dtype=torch.int64, device=dev)
    lib = _lib.load()
    nbytes = lib.geossl_weight_image_bytes()
    images = torch.empty((len(ws), 2, nbytes), dtype=torch.uint8, device=dev)
    check(lib.geossl_pack_weights_batched(_p(table), None, len(ws), _p(images), _stream()), "pack_weights_batched")
    return {l: (images[i, 0], images[i, 1]) for i, l in enumerate(layers)}


def linear(x, layer, pre_ssp=False, residual=None, images=None):
    """nn.Linear on the tensor-core path when it applies (128 -> 128, tensor-core mode), else library GEMM."""
    w = layer.weight
    if x.dim() == 2 and x.size(1) == 128 and x.is_cuda and linear_tc_applies(layer):
        return LinearTC.apply(x, w, layer.bias, residual, pre_ssp, None if images is None else images.get(layer))
    if pre_ssp:
        x = torch.nn.functional.softplus(x) - 0.6931471824645996
    y = torch.nn.functional.linear(x, w, layer.bias)
    return y if residual is None else residual + y


# =====================================================================================================
# wider atom-wise dense layers (PaiNN: 128->384, 256->128, 128->256) as 128 x 128 blocks of the same kernels
# =====================================================================================================
ACT_NONE, ACT_SSP, ACT_SILU = 0, 1, 2


def dense_tc_applies(layer):
    w = layer.weight
    return FILTER_MODE != "simt" and w.is_cuda and w.size(0) % 128 == 0 and w.size(1) % 128 == 0


def prepack_dense_blocks(weights):
    """Operand images of every 128 x 128 block of the given weight matrices (out, in multiples of 128) with ONE launch:
    returns {id(weight): images (out/128, in/128, 2, nbytes) uint8} -- [..., 0] forward (fp16 parts), [..., 1] transposed
    (bf16 parts, data gradient).  Valid for the current forward/backward only."""
    weights = [w for w in weights if w.is_cuda and w.size(0) % 128 == 0 and w.size(1) % 128 == 0]
    if not weights or FILTER_MODE == "simt":
        return {}
    dev = weights[0].device
    ptrs, lds, shapes = [], [], []
    for w in weights:
        wd = _req(w.detach(), torch.float32, "weight", 2)
        no, ni = wd.size(0) // 128, wd.size(1) // 128
        shapes.append((no, ni))
        for ob in range(no):
            for ib in range(ni):
                ptrs.append(wd.data_ptr() + 4 * (ob * 128 * wd.size(1) + ib * 128))
                lds.append(wd.size(1))
    key = (dev, "blocks", tuple(ptrs))
    table = _ptr_tables.get(key)
    if table is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("geossl_b200: run the model once before capturing it in a CUDA graph (weight pointer table)")
        if len(_ptr_tables) > 16:
            _ptr_tables.clear()
        table = _ptr_tables[key] = (torch.tensor(ptrs, dtype=torch.int64, device=dev), torch.tensor(lds, dtype=torch.int32, device=dev))
    lib = _lib.load()
    nbytes = lib.geossl_weight_image_bytes()
    images = torch.empty((len(ptrs), 2, nbytes), dtype=torch.uint8, device=dev)
    check(lib.geossl_pack_weights_batched(_p(table[0]), _p(table[1]), len(ptrs), _p(images), _stream()), "pack_weights_batched")
    out, off = {}, 0
    for w, (no, ni) in zip(weights, shapes):
        out[id(w)] = images[off:off + no * ni].view(no, ni, 2, nbytes)
        off += no * ni
    return out


def _off(t, cols):
    """Device pointer of column ``cols`` of row 0 of a row-major fp32 matrix."""
    return ctypes.c_void_p(t.data_ptr() + 4 * cols)


class DenseTC(torch.autograd.Function):
    """y (n,N) = act(x (n,K)) @ W (N,K)^T + b, K and N multiples of 128, as N/128 x K/128 launches of the 128 x 128 tensor-core
    layer kernel over column windows (geossl_linear_tc_block); K-blocks accumulate through the residual input.
    ``pre_act``: ACT_NONE / ACT_SSP / ACT_SILU applied to x as it is staged -- PaiNN's ``Dense(activation=silu)`` followed by a
    ``Dense`` is ``DenseTC(DenseTC(x, W1, b1), W2, b2, pre_act=ACT_SILU)`` (painn.py:21-24, painn_utils.py:9-35).
    Backward: data gradient through the transposed images (x act'(x) in the epilogue), weight gradients by
    geossl_linear_wgrad_tc_block per block (side stream inside ``side_stream_wgrads()``)."""

    @staticmethod
    def forward(ctx, x, weight, bias, pre_act, images):
        x, weight = _req(x, torch.float32, "x", 2), _req(weight, torch.float32, "weight", 2)
        N, K = weight.shape
        # K-padded operand (PaiNN's rbf matrix): x may hold only the first 32 / 64 live columns of a K = 128 layer
        k_cols = x.size(1) if (K == 128 and x.size(1) in (32, 64)) else 128
        if (x.size(1) != K and k_cols == 128) or N % 128 or K % 128:
            raise RuntimeError(f"geossl_b200: DenseTC needs in/out features in multiples of 128, got {tuple(weight.shape)}")
        ctx.k_cols = k_cols
        bias = None if bias is None else _req(bias, torch.float32, "bias", 1)
        n = x.size(0)
        ctx.pre_act, ctx.images, ctx.params = int(pre_act), images, (weight, bias)
        ctx.save_for_backward(x, weight)
        y = torch.empty((n, N), dtype=torch.float32, device=x.device)
        if n == 0:
            return y
        lib = _lib.load()
        act = ctx.pre_act or ACT_SSP
        for ob in range(N // 128):
            for ib in range(K // 128):
                _timed("dense_fwd", lambda: lib.geossl_linear_tc_block(
                    _off(x, ib * 128), x.size(1), n, _p(images[ob, ib, 0]), None if (bias is None or ib) else _off(bias, ob * 128), act,
                    1 if ctx.pre_act else 0, None, 128, _off(y, ob * 128) if ib else None, N, _off(y, ob * 128), N, 0, k_cols,
                    _stream()))
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        N, K = weight.shape
        n = x.size(0)
        lib = _lib.load()
        images, act = ctx.images, ctx.pre_act or ACT_SSP
        gx = None
        if n == 0:
            return torch.zeros_like(x), torch.zeros_like(weight), None if ctx.params[1] is None else torch.zeros(N, device=x.device), None, None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            for ib in range(K // 128):
                for ob in range(N // 128):
                    _timed("dense_dgrad", lambda: lib.geossl_linear_tc_block(
                        _off(gy, ob * 128), N, n, _p(images[ob, ib, 1]), None, act, 0,
                        _off(x, ib * 128) if ctx.pre_act else None, K, _off(gx, ib * 128) if ob else None, K, _off(gx, ib * 128), K, 1,
                        128, _stream()))
        gw = gb = None
        wp, bp = ctx.params
        if ctx.needs_input_grad[1] or (bp is not None and ctx.needs_input_grad[2]):
            def wgrad():
                gw_ = torch.empty_like(weight)
                gb_ = torch.empty(N, dtype=torch.float32, device=x.device) if bp is not None else None
                _wgrad_batch(n, [{"gy": (gy, ob * 128), "x": (x, ib * 128), "pre_act": ctx.pre_act, "x_cols": ctx.k_cols,
                                  "gw": (gw_, ob * 128 * K + ib * 128, K), "gb": None if (gb_ is None or ib) else (gb_, ob * 128)}
                                 for ob in range(N // 128) for ib in range(K // 128)], "dense_wgrad")   # all blocks: one launch per six
                return gw_, gb_

            deferrable = wp.is_leaf and wp.requires_grad and (bp is None or (bp.is_leaf and bp.requires_grad))
            if _SIDE["on"] and deferrable:
                main, side = torch.cuda.current_stream(x.device), _side_stream(x.device)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    gw, gb = wgrad()
                for t in (gy, x):
                    t.record_stream(side)
                for t in (gw, gb):
                    if t is not None:
                        t.record_stream(main)
                _SIDE["dirty"] = True
                _SIDE["pending"].append((wp, gw))
                if gb is not None:
                    _SIDE["pending"].append((bp, gb))
                gw = gb = None
            else:
                gw, gb = wgrad()
        return gx, gw, gb, None, None


def _chain_ex(n_rows, stages, bf16_parts, act, name):
    """stages: list of dicts with the fields of geossl_chain_stage_ex (tensors or None; missing = 0 / None); every tensor
    argument is (pointer to its first used element, row stride) given as ``(tensor, column_offset)``."""
    arr = (_lib.ChainStageEx * len(stages))()

    def ptr(spec):
        if spec is None:
            return None, 0
        t, col = spec
        return t.data_ptr() + 4 * col, t.stride(0) if t.dim() == 2 else 0

    for a, st in zip(arr, stages):
        a.weight_image = st["image"].data_ptr()
        b = st.get("bias")
        a.bias = None if b is None else b[0].data_ptr() + 4 * b[1]
        a.act_grad_input, a.ldz = ptr(st.get("z"))
        a.residual, a.ldr = ptr(st.get("residual"))
        a.store, a.ld_store = ptr(st.get("store"))
        a.x, a.ldx = ptr(st.get("x"))
        a.act_next, a.x_act = int(st.get("act_next", 0)), int(st.get("x_act", 0))
        a.keep, a.accumulate, a.partial = int(st.get("keep", 0)), int(st.get("accumulate", 0)), int(st.get("partial", 0))
    _timed(name, lambda: _lib.load().geossl_linear_chain_ex(n_rows, arr, len(stages), 1 if bf16_parts else 0, act, _stream()))


def _deferred_wgrads(compute, params, inputs):
    """Run ``compute() -> [grad tensors]`` (weight-gradient launches) on the side stream inside ``side_stream_wgrads()`` when every
    parameter is a leaf, handing the results to ``join_side_stream()``; otherwise run it in place.  Returns the list of
    gradients for autograd (all None when deferred)."""
    live = [p for p in params if p is not None]
    if _SIDE["on"] and all(p.is_leaf and p.requires_grad for p in live):
        dev = live[0].device
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            grads = compute()
        for t in inputs:
            t.record_stream(side)
        for p, g in zip(params, grads):
            if p is not None and g is not None:
                g.record_stream(main)
                _SIDE["pending"].append((p, g))
        _SIDE["dirty"] = True
        return [None] * len(params)
    return compute()


class DenseChain2(torch.autograd.Function):
    """y (n,N2) = silu(x (n,K) @ Wa (128,K)^T + ba) @ Wb (N2,128)^T + bb -- PaiNN's two-layer context nets (painn.py:21-24: K = 128,
    N2 = 384; :76-79: K = 256, N2 = 384) as ONE launch forward (K-blocks summed in the accumulator, the hidden tile handed to
    the three output blocks through shared memory) and ONE launch for the data gradient (geossl_linear_chain_ex); the hidden
    pre-activation z (n,128) is the only tensor kept besides x.  ``Wb = None``: a single bias-free Dense with N = Wa rows wider
    than 128 (mu_channel_mix, painn.py:83) -- fan-out only."""

    @staticmethod
    def forward(ctx, x, wa, ba, wb, bb, img_a, img_b):
        x = _req(x, torch.float32, "x", 2)
        n, K = x.shape
        ctx.two = wb is not None
        ctx.params = (wa, ba, wb, bb)
        ctx.images = (img_a, img_b)
        stages = []
        if ctx.two:
            nk, no = K // 128, wb.size(0) // 128
            z = torch.empty((n, 128), dtype=torch.float32, device=x.device)
            y = torch.empty((n, wb.size(0)), dtype=torch.float32, device=x.device)
            for kb in range(nk):
                st = {"image": img_a[0, kb, 0], "x": (x, kb * 128), "partial": kb < nk - 1, "accumulate": kb > 0}
                if kb == nk - 1:
                    st.update(bias=None if ba is None else (ba, 0), store=(z, 0), act_next=ACT_SILU)
                stages.append(st)
            for ob in range(no):
                stages.append({"image": img_b[ob, 0, 0], "bias": None if bb is None else (bb, ob * 128), "store": (y, ob * 128),
                               "keep": ob > 0})
            ctx.save_for_backward(x, z)
        else:
            no = wa.size(0) // 128
            y = torch.empty((n, wa.size(0)), dtype=torch.float32, device=x.device)
            for ob in range(no):
                st = {"image": img_a[ob, 0, 0], "bias": None if ba is None else (ba, ob * 128), "store": (y, ob * 128), "keep": ob > 0}
                if ob == 0:
                    st["x"] = (x, 0)
                stages.append(st)
            ctx.save_for_backward(x)
        if n:
            _chain_ex(n, stages, False, ACT_SILU, "dense_chain_fwd")
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        wa, ba, wb, bb = ctx.params
        img_a, img_b = ctx.images
        gy = gy.contiguous()
        lib = _lib.load()
        if ctx.two:
            x, z = ctx.saved_tensors
        else:
            (x,) = ctx.saved_tensors
        n, K = x.shape
        if n == 0:
            return (torch.zeros_like(x), torch.zeros_like(wa), None if ba is None else torch.zeros_like(ba),
                    None if wb is None else torch.zeros_like(wb), None if bb is None else torch.zeros_like(bb), None, None)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        stages = []
        if ctx.two:
            nk, no = K // 128, wb.size(0) // 128
            gz = torch.empty((n, 128), dtype=torch.float32, device=x.device)
            for ob in range(no):
                st = {"image": img_b[ob, 0, 1], "x": (gy, ob * 128), "partial": ob < no - 1, "accumulate": ob > 0}
                if ob == no - 1:
                    st.update(z=(z, 0), store=(gz, 0))
                stages.append(st)
            if gx is not None:
                for kb in range(nk):
                    stages.append({"image": img_a[0, kb, 1], "store": (gx, kb * 128), "keep": kb > 0})
            _chain_ex(n, stages, True, ACT_SILU, "dense_chain_bwd")

            def wgrads():
                gwa, gwb = torch.empty_like(wa), torch.empty_like(wb)
                gba = torch.empty_like(ba) if ba is not None else None
                gbb = torch.empty_like(bb) if bb is not None else None
                probs = [{"gy": (gy, ob * 128), "x": (z, 0), "pre_act": ACT_SILU, "gw": (gwb, ob * 128 * 128, 128),
                          "gb": None if gbb is None else (gbb, ob * 128)} for ob in range(no)]
                probs += [{"gy": (gz, 0), "x": (x, kb * 128), "gw": (gwa, kb * 128, K),
                           "gb": None if (gba is None or kb) else (gba, 0)} for kb in range(nk)]
                _wgrad_batch(n, probs, "dense_wgrad")          # every block of both layers: one launch
                return [gwa, gba, gwb, gbb]

            gwa, gba, gwb, gbb = _deferred_wgrads(wgrads, [wa, ba, wb, bb], [gy, z, gz, x])
            return gx, gwa, gba, gwb, gbb, None, None
        no = wa.size(0) // 128
        if gx is not None:
            for ob in range(no):
                st = {"image": img_a[ob, 0, 1], "x": (gy, ob * 128), "partial": ob < no - 1, "accumulate": ob > 0}
                if ob == no - 1:
                    st["store"] = (gx, 0)
                stages.append(st)
            _chain_ex(n, stages, True, ACT_SILU, "dense_chain_bwd")

        def wgrads1():
            gwa = torch.empty_like(wa)
            gba = torch.empty_like(ba) if ba is not None else None
            _wgrad_batch(n, [{"gy": (gy, ob * 128), "x": (x, 0), "gw": (gwa, ob * 128 * 128, 128),
                              "gb": None if gba is None else (gba, ob * 128)} for ob in range(no)], "dense_wgrad")
            return [gwa, gba, None, None]

        gwa, gba, _, _ = _deferred_wgrads(wgrads1, [wa, ba, None, None], [gy, x])
        return gx, gwa, gba, None, None, None, None


def dense_chain2_applies(x, la, lb, images):
    """Two Dense layers (la with SiLU, then lb) as one chain launch: 128-wide hidden layer, K and N2 multiples of 128, at most
    six 128 x 128 blocks in all; ``lb = None``: one bias-free fan-out layer with 128 inputs."""
    if not images or id(la.weight) not in images or not x.is_cuda or x.dim() != 2:
        return False
    if lb is None:
        return la.weight.size(1) == 128 and 2 <= la.weight.size(0) // 128 <= 6
    return (id(lb.weight) in images and la.weight.size(0) == 128 and lb.weight.size(1) == 128
            and la.weight.size(1) // 128 + lb.weight.size(0) // 128 <= 6)


class FilterPad(torch.autograd.Function):
    """[W_f | b_f | 0] (C,128) written into a PERSISTENT buffer (its address must not change between forwards: the packed
    block pointer table is cached, and a captured CUDA graph replays fixed addresses).  Backward slices the gradient."""

    @staticmethod
    def forward(ctx, weight, bias, buf):
        R = weight.size(1)
        buf[:, :R].copy_(weight)
        buf[:, R].copy_(bias)
        ctx.R = R
        ctx.mark_dirty(buf)
        return buf

    @staticmethod
    def backward(ctx, g):
        return g[:, :ctx.R], g[:, ctx.R], None


def dense(x, layer, images=None, pre_act=ACT_NONE):
    """``layer(act(x))`` for an nn.Linear-like layer with in/out features in multiples of 128: tensor-core blocks when
    ``images`` (from prepack_dense_blocks) has the layer, library GEMM otherwise."""
    w = layer.weight
    if images and id(w) in images and x.is_cuda and x.dim() == 2:
        return DenseTC.apply(x, w, layer.bias, pre_act, images[id(w)])
    if pre_act == ACT_SILU:
        x = torch.nn.functional.silu(x)
    elif pre_act == ACT_SSP:
        x = torch.nn.functional.softplus(x) - 0.6931471824645996
    return torch.nn.functional.linear(x, w, layer.bias)


# =====================================================================================================
# tensor-core products closed under differentiation (the edge-sized filter MLP of the double-backward path)
# =====================================================================================================
# Force training (finetune_md17.py:32-54) differentiates the backward pass again, so the fused once-differentiable filter
# kernels do not apply and the filter MLP (schnet.py:141-145) runs as separate products over ~10^5 edge rows.  The three
# forms below are each other's derivatives, so autograd can differentiate them to any order while every product stays on the
# 128 x 128 tcgen05 block kernels (split-precision operands, fp32 accumulate):
#   MatXWt : y = x[:, :k] @ W[:, :k]^T (+ b)     geossl_linear_tc_block, weight image as stored
#   MatXW  : y = g @ W                          geossl_linear_tc_block, transposed weight image
#   MatTX  : G = a^T @ b[:, :k]                 geossl_linear_wgrad_tc_block (reduction over the rows)
# W and G are (128,128); x / b may be K-padded operands with k in {32, 64, 128} live columns (the 50 gaussians, padded to 64).
# Packed operand images of the composed path, valid for ONE forward and its backward passes: SchNet.forward clears the table
# (``begin_composed_pass``), so an image is never reused across an optimizer step; inside the pass a weight is packed once per
# orientation instead of once per product (118 -> ~30 pack launches per MD17 step).  Entries keep their source tensor alive,
# so a recycled address cannot alias.
_mm_images = {}


def begin_composed_pass():
    _mm_images.clear()


class _param_grads_off:
    """Inside this context the backward of MatXWt / MatXW skips weight and bias gradients.  For ``autograd.grad(energy,
    positions, create_graph=True)`` (finetune_md17.py:40-44): only d/dpos is requested there, torch prunes the parameter
    branches of its own ops in that case, and a custom Function cannot see the pruning -- without the flag every product of
    the force pass would also launch a (discarded) edge-sized weight-gradient GEMM."""
    on = False

    def __enter__(self):
        self.prev = _param_grads_off.on
        _param_grads_off.on = True

    def __exit__(self, *exc):
        _param_grads_off.on = self.prev


def param_grads_disabled():
    return _param_grads_off()


def _mm_image(w, transpose, bf16_parts):
    key = (w.data_ptr(), w._version, bool(transpose), bool(bf16_parts))
    hit = _mm_images.get(key)
    if hit is not None:                       # (the entry holds the source tensor, so this address + version is still that tensor's data)
        return hit[1]
    lib = _lib.load()
    image = torch.empty(lib.geossl_weight_image_bytes(), dtype=torch.uint8, device=w.device)
    check(lib.geossl_pack_weight(_p(w), 1 if transpose else 0, 1 if bf16_parts else 0, _p(image), _stream()), "pack_weight")
    _mm_images[key] = (w, image)
    return image


def _mm_check(t, name, cols=None):
    t = _req(t, torch.float32, name, 2)
    if cols is not None and t.size(1) not in cols:
        raise RuntimeError(f"geossl_b200: `{name}` must have {cols} columns, got {t.size(1)}")
    return t


class MatXWt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, fp16_parts):
        x, w = _mm_check(x, "x", (32, 64, 128)), _mm_check(w, "w", (128,))
        if w.size(0) != 128:
            raise RuntimeError("geossl_b200: MatXWt is built for (128,128) weights")
        bias = None if bias is None else _req(bias, torch.float32, "bias", 1)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, w)
        n, k = x.shape
        y = torch.empty((n, 128), dtype=torch.float32, device=x.device)
        if n:
            image = _mm_image(w, False, not fp16_parts)
            _timed("mm_xwt", lambda: _lib.load().geossl_linear_tc_block(
                _p(x), k, n, _p(image), _p(bias), ACT_SSP, 0, None, 128, None, 128, _p(y), 128, 0 if fp16_parts else 1, k, _stream()))
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = gy.contiguous()
        gx = MatXW.apply(gy, w)[:, :x.size(1)] if ctx.needs_input_grad[0] else None
        want_b = ctx.has_bias and ctx.needs_input_grad[2] and not _param_grads_off.on
        gw = gb = None
        if ctx.needs_input_grad[1] and not _param_grads_off.on:
            gw, colsum = MatTX.apply(gy, x, want_b)          # the bias gradient is the kernel's column sum of gy: no extra pass
            gb = colsum if want_b else None
        elif want_b:
            gb = gy.sum(0)
        return gx, gw, gb, None


class MatXW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, w):
        g, w = _mm_check(g, "g", (128,)), _mm_check(w, "w", (128,))
        ctx.save_for_backward(g, w)
        n = g.size(0)
        y = torch.empty((n, 128), dtype=torch.float32, device=g.device)
        if n:
            image = _mm_image(w, True, True)
            _timed("mm_xw", lambda: _lib.load().geossl_linear_tc_block(
                _p(g), 128, n, _p(image), None, ACT_SSP, 0, None, 128, None, 128, _p(y), 128, 1, 128, _stream()))
        return y

    @staticmethod
    def backward(ctx, go):
        g, w = ctx.saved_tensors
        go = go.contiguous()
        gg = MatXWt.apply(go, w, None, False) if ctx.needs_input_grad[0] else None
        gw = MatTX.apply(g, go, False)[0] if (ctx.needs_input_grad[1] and not _param_grads_off.on) else None
        return gg, gw


class MatTX(torch.autograd.Function):
    """(a^T @ b[:, :k], column sums of a) -- the second output only when ``want_colsum`` (else an empty tensor)."""

    @staticmethod
    def forward(ctx, a, b, want_colsum):
        a, b = _mm_check(a, "a", (128,)), _mm_check(b, "b", (32, 64, 128))
        ctx.save_for_backward(a, b)
        ctx.set_materialize_grads(False)
        n, k = b.shape
        # (the kernel writes all 128 x 128 entries -- columns past k come out zero -- and the 128 column sums)
        out = torch.empty((128, 128), dtype=torch.float32, device=a.device) if n else torch.zeros((128, 128), dtype=torch.float32, device=a.device)
        colsum = (torch.empty if n else torch.zeros)(128 if want_colsum else 0, dtype=torch.float32, device=a.device)
        if n:
            lib = _lib.load()
            ws = torch.empty(lib.geossl_linear_wgrad_tc_workspace(n), dtype=torch.float32, device=a.device)
            _timed("mm_tx", lambda: lib.geossl_linear_wgrad_tc_block(_p(a), 128, _p(b), k, n, 0, _p(ws), _p(out), 128,
                                                                     _p(colsum) if want_colsum else None, k, _stream()))
        return out, colsum

    @staticmethod
    def backward(ctx, gG, gcol):
        a, b = ctx.saved_tensors
        ga = gb = None
        if gG is not None:
            gG = gG.contiguous()
            ga = MatXWt.apply(b, gG, None, False) if ctx.needs_input_grad[0] else None
            gb = MatXW.apply(a, gG)[:, :b.size(1)] if ctx.needs_input_grad[1] else None
        if gcol is not None and gcol.numel() and ctx.needs_input_grad[0]:
            ga = gcol.unsqueeze(0).expand_as(a) if ga is None else ga + gcol.unsqueeze(0)
        return ga, gb, None


class SspFamily(torch.autograd.Function):
    """E_k(a, g) = g * ssp^(k)(a) (geossl_ssp_family), k = 0 the activation itself: each member's derivatives are members,
    so the activation of the double-backward path costs one launch per use instead of torch's expanded formulas."""

    @staticmethod
    def forward(ctx, a, g, k):
        a = _req(a, torch.float32, "a")
        g = None if g is None else _req(g, torch.float32, "g")
        ctx.k = int(k)
        ctx.save_for_backward(a, g)
        out = torch.empty_like(a)
        check(_lib.load().geossl_ssp_family(_p(a), _p(g), a.numel(), ctx.k, _p(out), _stream()), "ssp_family")
        return out

    @staticmethod
    def backward(ctx, v):
        a, g = ctx.saved_tensors
        if ctx.k >= 3 and ctx.needs_input_grad[0]:
            raise RuntimeError("geossl_b200: SspFamily is built up to the third derivative of the activation")
        ga = SspFamily.apply(a, v if g is None else v * g, ctx.k + 1) if ctx.needs_input_grad[0] else None
        gg = SspFamily.apply(a, v.contiguous(), ctx.k) if (g is not None and ctx.needs_input_grad[1]) else None
        return ga, gg, None


def _closed_elementwise_applies(x):
    return FILTER_MODE != "simt" and x.is_cuda and x.dtype == torch.float32 and x.numel() > 0


def ssp_any_order(x, shift=0.6931471824645996):
    """Shifted softplus (schnet.py:215-216), differentiable to any order the force training needs, one launch per use."""
    if _closed_elementwise_applies(x):
        return SspFamily.apply(x.contiguous(), None, 0)
    return torch.nn.functional.softplus(x) - shift


class RowScale(torch.autograd.Function):
    """out[r][f] = x[r][f] * c[r]."""

    @staticmethod
    def forward(ctx, x, c):
        x, c = _req(x, torch.float32, "x", 2), _req(c, torch.float32, "c", 1)
        ctx.save_for_backward(x, c)
        out = torch.empty_like(x)
        check(_lib.load().geossl_row_scale(_p(x), _p(c), x.size(0), x.size(1), _p(out), _stream()), "row_scale")
        return out

    @staticmethod
    def backward(ctx, v):
        x, c = ctx.saved_tensors
        v = v.contiguous()
        gx = RowScale.apply(v, c) if ctx.needs_input_grad[0] else None
        gc = RowDot.apply(v, x) if ctx.needs_input_grad[1] else None
        return gx, gc


class RowDot(torch.autograd.Function):
    """out[r] = sum_f a[r][f] * b[r][f]."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _req(a, torch.float32, "a", 2), _req(b, torch.float32, "b", 2)
        ctx.save_for_backward(a, b)
        out = torch.empty(a.size(0), dtype=torch.float32, device=a.device)
        check(_lib.load().geossl_row_dot(_p(a), _p(b), a.size(0), a.size(1), _p(out), _stream()), "row_dot")
        return out

    @staticmethod
    def backward(ctx, v):
        a, b = ctx.saved_tensors
        v = v.contiguous()
        ga = RowScale.apply(b, v) if ctx.needs_input_grad[0] else None
        gb = RowScale.apply(a, v) if ctx.needs_input_grad[1] else None
        return ga, gb


def row_scale(x, c):
    """``x * c.view(-1, 1)`` (the cutoff product of schnet.py:187) with derivatives that stay single launches."""
    if _closed_elementwise_applies(x) and x.dim() == 2 and x.size(1) in (32, 64, 128):
        return RowScale.apply(x.contiguous(), c.contiguous())
    return x * c.view(-1, 1)


def linear_any_order(x, layer):
    """``layer(x)`` for a 128 -> 128 nn.Linear, differentiable to any order on the tensor cores (MatXWt); other shapes,
    CPU tensors and the exact ``simt`` mode take torch's linear."""
    w = layer.weight
    if (FILTER_MODE != "simt" and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and tuple(w.shape) == (128, 128)
            and x.size(1) == 128):
        return MatXWt.apply(x, w, layer.bias, True)
    return layer(x)


def filter_mlp_applies(lin0, lin2, edge_attr):
    return (FILTER_MODE != "simt" and edge_attr.is_cuda and edge_attr.dim() == 2 and edge_attr.dtype == torch.float32
            and tuple(lin2.weight.shape) == (128, 128) and lin0.weight.size(0) == 128 and lin0.weight.size(1) <= 128
            and lin0.bias is not None and lin2.bias is not None)


def filter_mlp(edge_attr, lin0, lin2):
    """``lin2(ssp(lin0(edge_attr)))`` (the filter-generating network, schnet.py:141-145) differentiable to any order with
    both products on the tensor cores; the activation stays a torch op between them."""
    G = edge_attr.size(1)
    k = 32 if G <= 32 else (64 if G <= 64 else 128)
    x = torch.nn.functional.pad(edge_attr, (0, k - G)) if k != G else edge_attr
    w1 = torch.nn.functional.pad(lin0.weight, (0, 128 - G)) if G != 128 else lin0.weight
    h = MatXWt.apply(x, w1, lin0.bias, True)
    return MatXWt.apply(ssp_any_order(h), lin2.weight, lin2.bias, True)


# =====================================================================================================
# embedding lookup with a short, deterministic backward
# =====================================================================================================
class EmbeddingLookup(torch.autograd.Function):
    """``weight[z]`` for a small vocabulary (atom classes).  torch's embedding backward sorts the indices and runs a
    segmented reduction (about twenty small launches); with a 9..100-row table the gradient is one GEMM
    ``one_hot(z)^T @ grad`` -- four launches, no atomics, fixed summation order."""

    @staticmethod
    def forward(ctx, weight, z, padding_idx):
        ctx.save_for_backward(z)
        ctx.rows, ctx.padding_idx = weight.size(0), padding_idx
        return weight.index_select(0, z)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        (z,) = ctx.saved_tensors
        onehot = torch.nn.functional.one_hot(z, ctx.rows).to(grad.dtype)
        gw = onehot.t() @ grad
        if ctx.padding_idx is not None:
            gw[ctx.padding_idx] = 0                     # nn.Embedding(padding_idx=...) never updates that row (painn.py:174)
        return gw, None, None


def embedding(module, z):
    """``module(z)`` for an nn.Embedding, through EmbeddingLookup when the table is small and on the GPU."""
    w = module.weight
    if w.is_cuda and w.size(0) <= 128 and z.dim() == 1 and module.max_norm is None and not module.sparse:
        return EmbeddingLookup.apply(w, z, module.padding_idx)
    return module(z)


# =====================================================================================================
# DDM head
# =====================================================================================================
def pair_distance(pos, super_edge_index):
    """(P,1) distances over ``super_edge_index`` (pretrain_GeoSSL.py:199-205); no gradient."""
    pos = _req(pos.detach(), torch.float32, "pos", 2)
    sei = _req(super_edge_index, torch.int64, "super_edge_index", 2)
    n_pairs = sei.size(1)
    out = torch.empty((n_pairs, 1), dtype=torch.float32, device=pos.device)
    check(_lib.load().geossl_pair_distance(_p(pos), _p(sei), n_pairs, _p(out), _stream()), "pair_distance")
    return out


def _ddm_ptrs(tensors):
    s = DdmPtrs()
    for name, t in zip([f[0] for f in DdmPtrs._fields_], tensors):
        setattr(s, name, t.data_ptr())
    return s


# Training fast path of the tensor-core head: ONE pass computes the loss and (unscaled) gradients (geossl_ddm_head_fwd_bwd_tc),
# the backward is a multi-tensor scale by grad_loss / n_graphs.  False restores separate forward / backward kernels.
FUSE_DDM_HEAD = True


class DDMHead(torch.autograd.Function):
    """NCSN_version_03.forward with the random draws supplied (NCSN.py:183-212).  Returns the scalar loss.
    Gradients: node_feature and the ten MLP parameters (distance carries none, as in the reference).
    ``n_pairs_live``: optional (1,) int32 device tensor -- the batch is capacity padded and only that many pairs at the
    front are live (CUDA-graph replay over variable-size batches)."""

    @staticmethod
    def forward(ctx, node_feature, sei, batch, dist, noise, noise_level, sigmas, anneal_power, n_pairs_live, train_pass, *params):
        h = _req(node_feature, torch.float32, "node_feature", 2)
        sei = _req(sei, torch.int64, "super_edge_index", 2)
        batch = _req(batch, torch.int64, "batch", 1)
        dist = _req(dist.detach(), torch.float32, "distance").view(-1)
        noise = _req(noise, torch.float32, "distance_noise").view(-1)
        noise_level = _req(noise_level, torch.int64, "noise_level", 1)
        sigmas = _req(sigmas.detach(), torch.float32, "sigmas", 1)
        live = None if n_pairs_live is None else _req(n_pairs_live, torch.int32, "n_pairs_live", 1)
        params = tuple(_req(p, torch.float32, "mlp parameter") for p in params)
        lib = _lib.load()
        H = h.size(1)
        n_pairs = sei.size(1)
        ctx.tc = FILTER_MODE != "simt" and H == 128
        # ``train_pass`` = torch.is_grad_enabled() at the call site (inside forward() grad mode is always off, and
        # needs_input_grad ignores torch.no_grad()): evaluation takes the forward-only kernel
        ctx.fused = bool(ctx.tc and FUSE_DDM_HEAD and n_pairs > 0 and train_pass
                         and any(ctx.needs_input_grad[i] for i in (0, *range(10, 10 + len(params)))))
        ws = torch.empty(max(lib.geossl_ddm_workspace_fused(n_pairs, h.size(0)) if ctx.fused else
                             (lib.geossl_ddm_workspace_tc(n_pairs) if ctx.tc else lib.geossl_ddm_workspace(H)), 1),
                         dtype=torch.float32, device=h.device)
        loss = torch.empty(2, dtype=torch.float32, device=h.device)
        pp = _ddm_ptrs(params)
        ctx.anneal_power = float(anneal_power)
        if ctx.fused:
            grad_h = torch.empty_like(h)
            grads = [torch.empty_like(p) for p in params]
            gp = _ddm_ptrs(grads)
            _timed("ddm_head_fused", lambda: lib.geossl_ddm_head_fwd_bwd_tc(
                _p(h), _p(sei), _p(batch), n_pairs, _p(live), h.size(0), _p(dist), _p(noise), _p(noise_level), _p(sigmas),
                sigmas.numel(), float(anneal_power), H, ctypes.byref(pp), _p(ws), _p(loss), _p(grad_h), ctypes.byref(gp), _stream()))
            ctx.save_for_backward(loss, grad_h, *grads)
            return loss[0]
        fwd = lib.geossl_ddm_head_fwd_tc if ctx.tc else lib.geossl_ddm_head_fwd
        _timed("ddm_head_fwd", lambda: fwd(
            _p(h), _p(sei), _p(batch), n_pairs, _p(live), _p(dist), _p(noise), _p(noise_level), _p(sigmas), sigmas.numel(),
            float(anneal_power), H, ctypes.byref(pp), _p(ws), _p(loss), _stream()))
        ctx.live = live
        ctx.save_for_backward(h, sei, batch, dist, noise, noise_level, sigmas, loss, *params)
        return loss[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss):
        if ctx.fused:
            loss, grad_h, *grads = ctx.saved_tensors
            ng = loss[1]
            scale = torch.where(ng > 0, grad_loss.to(torch.float32) / ng, torch.zeros_like(ng))
            scaled = torch._foreach_mul([grad_h, *grads], scale)              # out of place: a retained graph may run again
            return (scaled[0], None, None, None, None, None, None, None, None, None, *scaled[1:])
        h, sei, batch, dist, noise, noise_level, sigmas, loss, *params = ctx.saved_tensors
        lib = _lib.load()
        H, n_pairs = h.size(1), sei.size(1)
        grad_h = torch.empty_like(h)
        grads = [torch.empty_like(p) for p in params]
        if n_pairs == 0:
            return (torch.zeros_like(h), None, None, None, None, None, None, None, None, None, *[torch.zeros_like(p) for p in params])
        ws = torch.empty(lib.geossl_ddm_workspace_tc(n_pairs) if ctx.tc else lib.geossl_ddm_workspace(H),
                         dtype=torch.float32, device=h.device)
        gl = grad_loss.contiguous().view(1).to(torch.float32)
        pp, gp = _ddm_ptrs(params), _ddm_ptrs(grads)
        bwd = lib.geossl_ddm_head_bwd_tc if ctx.tc else lib.geossl_ddm_head_bwd
        _timed("ddm_head_bwd", lambda: bwd(
            _p(h), _p(sei), _p(batch), n_pairs, _p(ctx.live), h.size(0), _p(dist), _p(noise), _p(noise_level), _p(sigmas),
            sigmas.numel(), ctx.anneal_power, H, ctypes.byref(pp), _p(loss), _p(gl), _p(ws), _p(grad_h), ctypes.byref(gp),
            _stream()))
        return (grad_h, None, None, None, None, None, None, None, None, None, *grads)


# =====================================================================================================
# PaiNN message block
# =====================================================================================================
class PaiNNEdges:
    """Per-edge geometry of one view + the two groupings of a ``radius_edge_index`` = [idx_i; idx_j]."""

    def __init__(self, structure, dist, dir_, fcut, offsets, widths):
        self.s, self.dist, self.dir, self.fcut = structure, dist, dir_, fcut
        self.offsets, self.widths = offsets, widths
        self._phi_pad = None

    def phi_pad(self):
        """(E,32|64|128) = [rbf values, 1, 0...]: K-padded operand of the tensor-core filter GEMM (built once per forward)."""
        if self._phi_pad is None:
            e, R = self.dist.numel(), self.offsets.numel()
            ld = 32 if R < 32 else (64 if R < 64 else 128)
            self._phi_pad = torch.empty((e, ld), dtype=torch.float32, device=self.dist.device)
            check(_lib.load().geossl_painn_rbf_pad(_p(self.dist), _p(self.fcut), e, _p(self.offsets), _p(self.widths),
                                                   R, ld, _p(self._phi_pad), _stream()), "painn_rbf_pad")
        return self._phi_pad


_structure_cache = {}


def _painn_structure(radius_edge_index, n_atoms, batch, num_graphs, assume_sorted=False):
    """CSR by idx_j (row 1) and its grouping by idx_i, cached per edge tensor (both DDM views share it,
    pretrain_GeoSSL.py:190-191).  Falls back to a stable sort if the list is not (idx_j, idx_i)-sorted; the check
    costs one host sync, ``assume_sorted=True`` (lists produced by radius_graph are sorted) skips it so that the step
    can be captured in a CUDA graph."""
    key = (radius_edge_index.data_ptr(), radius_edge_index._version, tuple(radius_edge_index.shape), n_atoms,
           batch.data_ptr(), num_graphs)
    hit = _structure_cache.get(key)
    # the entry keeps its source tensors alive, so a recycled address can only hit with the very same objects
    if hit is not None and hit.key_rei is radius_edge_index and hit.key_batch is batch \
            and not torch.cuda.is_current_stream_capturing():
        return hit
    rei = _req(radius_edge_index, torch.int64, "radius_edge_index", 2)
    if rei.size(1) > 1 and not assume_sorted:
        k = rei[1] * n_atoms + rei[0]
        if not bool((k[1:] > k[:-1]).all().item()):
            perm = torch.sort(k, stable=True).indices
            rei = rei[:, perm].contiguous()
    g = csr_from_edge_index(rei, n_atoms, batch, num_graphs=num_graphs).ensure_transpose()
    g.rei = rei
    g.key_rei, g.key_batch = radius_edge_index, batch
    if len(_structure_cache) > 8:
        _structure_cache.clear()
    _structure_cache[key] = g
    return g


def painn_edges(positions, radius_edge_index, n_atoms, batch, offsets, widths, cutoff, num_graphs=None, assume_sorted=False):
    pos = _req(positions.detach(), torch.float32, "positions", 2)
    s = _painn_structure(radius_edge_index, n_atoms, batch, num_graphs, assume_sorted)
    e = s.rei.size(1)
    dev = pos.device
    dist = torch.empty(e, dtype=torch.float32, device=dev)
    dir_ = torch.empty((e, 3), dtype=torch.float32, device=dev)
    fcut = torch.empty(e, dtype=torch.float32, device=dev)
    check(_lib.load().geossl_painn_edge_geometry(_p(pos), _p(s.rei), e, n_atoms, float(cutoff), _p(dist), _p(dir_), _p(fcut),
                                                 _stream()), "painn_edge_geometry")
    return PaiNNEdges(s, dist, dir_, fcut, _req(offsets, torch.float32, "offsets", 1), _req(widths, torch.float32, "widths", 1))


class PaiNNMessage(torch.autograd.Function):
    """(q, mu, ctx, filter) -> (q + dq, mu + dmu)   (painn.py:53-64).  The filter of painn.py:241 is either rebuilt per edge
    from the ``filter_net`` slice ``(fw, fb)`` inside the kernels (``wpre`` None) or streamed from the materialised
    pre-cutoff rows ``wpre`` (E,3F) of the tensor-core filter GEMM -- then ``wpre`` receives the per-edge gradient and the
    ``filter_net`` gradients flow through that GEMM."""

    @staticmethod
    def forward(ctx, q, mu, x, fw, fb, edges, wpre=None):
        q, mu, x = _req(q, torch.float32, "q", 2), _req(mu, torch.float32, "mu", 3), _req(x, torch.float32, "ctx", 2)
        if wpre is None:
            fw, fb = _req(fw, torch.float32, "filter_w", 2), _req(fb, torch.float32, "filter_b", 1)
        else:
            wpre = _req(wpre, torch.float32, "filter_pre", 2)
            fw = fb = None
        n, Fd = q.shape
        s = edges.s
        q_out, mu_out = torch.empty_like(q), torch.empty_like(mu)
        _timed("painn_message_fwd", lambda: _lib.load().geossl_painn_message_fwd(
            _p(q), _p(mu), _p(x), _p(fw), _p(fb), _p(edges.offsets), _p(edges.widths), edges.offsets.numel(), Fd, _p(edges.dist),
            _p(edges.dir), _p(edges.fcut), _p(s.t_rowptr), _p(s.t_eid), _p(s.t_tgt), n, _p(q_out), _p(mu_out), _p(wpre), _stream()))
        ctx.edges, ctx.has_pre = edges, wpre is not None
        if wpre is None:
            ctx.save_for_backward(mu, x, fw, fb)
        else:
            ctx.save_for_backward(mu, x, wpre)
        return q_out, mu_out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gq_out, gmu_out):
        edges, s = ctx.edges, ctx.edges.s
        lib = _lib.load()
        n, Fd = gq_out.shape
        e = s.rei.size(1)
        R = edges.offsets.numel()
        gq_out, gmu_out = gq_out.contiguous(), gmu_out.contiguous()
        if ctx.has_pre:
            mu, x, wpre = ctx.saved_tensors
            gx, gmu = torch.empty_like(x), torch.empty_like(mu)
            # (rows past the live edge count of a capacity-padded list are zeroed by the kernel itself: the filter GEMM's
            #  weight-gradient kernel contracts over ALL rows)
            gpre = torch.empty((max(e, 1), 3 * Fd), dtype=torch.float32, device=x.device)
            _timed("painn_message_bwd", lambda: lib.geossl_painn_message_bwd(
                _p(gq_out), _p(gmu_out), _p(mu), _p(x), None, None, _p(edges.offsets), _p(edges.widths), R, Fd, _p(edges.dist),
                _p(edges.dir), _p(edges.fcut), _p(s.rowptr), _p(s.src), n, e, _p(gx), _p(gmu), _p(gpre), None, None, None,
                _p(wpre), _stream()))
            return gq_out, gmu, gx, None, None, None, gpre[:e] if e > 0 else torch.zeros_like(wpre)
        mu, x, fw, fb = ctx.saved_tensors
        gx, gmu = torch.empty_like(x), torch.empty_like(mu)
        gw, gb = torch.empty_like(fw), torch.empty_like(fb)
        scratch = torch.empty((max(e, 1), 3 * Fd), dtype=torch.float32, device=x.device)
        ws = torch.empty(lib.geossl_painn_workspace(R, Fd), dtype=torch.float32, device=x.device)
        _timed("painn_message_bwd", lambda: lib.geossl_painn_message_bwd(
            _p(gq_out), _p(gmu_out), _p(mu), _p(x), _p(fw), _p(fb), _p(edges.offsets), _p(edges.widths), R, Fd, _p(edges.dist),
            _p(edges.dir), _p(edges.fcut), _p(s.rowptr), _p(s.src), n, e, _p(gx), _p(gmu), _p(scratch), _p(ws), _p(gw), _p(gb),
            None, _stream()))
        return gq_out, gmu, gx, gw, gb, None, None


# =====================================================================================================
# PaiNN update (mixing) block, fused elementwise parts
# =====================================================================================================
class PaiNNMixPre(torch.autograd.Function):
    """(q (N,F), mu_mix (N,3,2F)) -> ctx (N,2F) = [q, |mu_V|_eps], dot (N,F) = <mu_V, mu_W>   (painn.py:100-105,111)."""

    @staticmethod
    def forward(ctx_, q, mu_mix, epsilon):
        q, mu_mix = _req(q, torch.float32, "q", 2), _req(mu_mix, torch.float32, "mu_mix", 3)
        n, Fd = q.shape
        out = torch.empty((n, 2 * Fd), dtype=torch.float32, device=q.device)
        dot = torch.empty((n, Fd), dtype=torch.float32, device=q.device)
        check(_lib.load().geossl_painn_mix_pre(_p(q), _p(mu_mix), n, Fd, float(epsilon), _p(out), _p(dot), _stream()), "painn_mix_pre")
        ctx_.save_for_backward(mu_mix, out)
        return out, dot

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx_, g_ctx, g_dot):
        mu_mix, out = ctx_.saved_tensors
        n, Fd = out.size(0), out.size(1) // 2
        g_q = torch.empty((n, Fd), dtype=torch.float32, device=out.device)
        g_mm = torch.empty_like(mu_mix)
        check(_lib.load().geossl_painn_mix_pre_bwd(_p(mu_mix), _p(out), _p(g_ctx.contiguous()), _p(g_dot.contiguous()), n, Fd, _p(g_q),
                                                   _p(g_mm), _stream()), "painn_mix_pre_bwd")
        return g_q, g_mm, None


class PaiNNMixPost(torch.autograd.Function):
    """(q, mu, y = [a|b|c], mu_mix, dot) -> q + a + c*dot, mu + b*mu_W   (painn.py:108-113)."""

    @staticmethod
    def forward(ctx_, q, mu, y, mu_mix, dot):
        q, mu, y = _req(q, torch.float32, "q", 2), _req(mu, torch.float32, "mu", 3), _req(y, torch.float32, "y", 2)
        mu_mix, dot = _req(mu_mix, torch.float32, "mu_mix", 3), _req(dot, torch.float32, "dot", 2)
        n, Fd = q.shape
        q_out, mu_out = torch.empty_like(q), torch.empty_like(mu)
        check(_lib.load().geossl_painn_mix_post(_p(q), _p(mu), _p(y), _p(mu_mix), _p(dot), n, Fd, _p(q_out), _p(mu_out), _stream()),
              "painn_mix_post")
        ctx_.save_for_backward(y, mu_mix, dot)
        return q_out, mu_out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx_, gq, gmu):
        y, mu_mix, dot = ctx_.saved_tensors
        n, Fd = dot.shape
        gq, gmu = gq.contiguous(), gmu.contiguous()
        g_y, g_mm, g_dot = torch.empty_like(y), torch.empty_like(mu_mix), torch.empty_like(dot)
        check(_lib.load().geossl_painn_mix_post_bwd(_p(gq), _p(gmu), _p(y), _p(mu_mix), _p(dot), n, Fd, _p(g_y), _p(g_mm), _p(g_dot),
                                                    _stream()), "painn_mix_post_bwd")
        return gq, gmu, g_y, g_mm, g_dot
