"""Inner steps of the fine-tune callers of the encoder.

``md17_losses`` is the body of /root/reference/examples/finetune_md17.py::train (:30-48): energy from the readout,
force = -dE/dpos by autograd with ``create_graph=True``, and the weighted L1 objective whose backward differentiates
through the force (double backward; served by the closed-under-differentiation cfconv primitives in ops.py).
``lba_loss`` is the body of finetune_lba.py::train (:33-47).
"""
import torch
from torch.autograd import grad

from . import ops


def _readout(args, model, batch_data, positions):
    x = batch_data.x[:, 0] if batch_data.x.dim() == 2 else batch_data.x
    if args.model_3d == "schnet":
        out = model(x, positions, batch_data.batch, num_graphs=getattr(batch_data, "n_graphs", None),
                    graph=getattr(batch_data, "extras", {}).get("graph"))
    elif args.model_3d == "painn":
        out = model(x, positions, batch_data.radius_edge_index, batch_data.batch,
                    num_graphs=getattr(batch_data, "n_graphs", None))
    else:
        raise Exception("3D model {} not included.".format(args.model_3d))
    live = getattr(batch_data, "extras", {}).get("n_graphs_live")       # capacity-padded batch: drop the padding graph
    return out if live is None else out[:live]


def md17_losses(args, batch_data, model, graph_pred_linear, criterion, energy_coeff=0.05, force_coeff=0.95):
    """Returns ``(loss, pred_energy, pred_force)``; ``batch_data`` carries ``y`` (B,) and ``force`` (N,3) in extras or
    as attributes.  Coefficients default to submit_finetune_md17_schnet.sh's 0.05 / 0.95."""
    positions = batch_data.positions
    positions.requires_grad_()
    molecule_3D_repr = _readout(args, model, batch_data, positions)
    pred_energy = (graph_pred_linear(molecule_3D_repr) if graph_pred_linear is not None else molecule_3D_repr).squeeze(1)
    with ops.param_grads_disabled():        # only d/dpos is asked for: no (discarded) weight-gradient launches in the force pass
        pred_force = -grad(outputs=pred_energy, inputs=positions, grad_outputs=torch.ones_like(pred_energy),
                           create_graph=True, retain_graph=True)[0]
    y = getattr(batch_data, "y", None)
    y = batch_data.extras["y"] if y is None else y
    f = getattr(batch_data, "force", None)
    f = batch_data.extras["force"] if f is None else f
    loss = energy_coeff * criterion(pred_energy, y) + force_coeff * criterion(pred_force, f)
    return loss, pred_energy, pred_force


def md17_train_step(args, batch_data, model, graph_pred_linear, criterion, optimizer, grad_sync=None, zero_grad=True, **coeffs):
    """One iteration of finetune_md17.py::train (:30-56); ``grad_sync`` = the data-parallel gradient exchange."""
    loss, _, _ = md17_losses(args, batch_data, model, graph_pred_linear, criterion, **coeffs)
    if zero_grad:
        optimizer.zero_grad()
    loss.backward()
    if grad_sync is not None:
        grad_sync()
    optimizer.step()
    return loss.detach()


def lba_loss(args, batch, model, graph_pred_linear, criterion):
    molecule_3D_repr = _readout(args, model, batch, batch.positions)
    pred = (graph_pred_linear(molecule_3D_repr) if graph_pred_linear is not None else molecule_3D_repr).squeeze()
    y = getattr(batch, "y", None)
    return criterion(pred, batch.extras["y"] if y is None else y)


def lba_train_step(args, batch, model, graph_pred_linear, criterion, optimizer, grad_sync=None, zero_grad=True):
    """One iteration of finetune_lba.py::train (:33-51)."""
    loss = lba_loss(args, batch, model, graph_pred_linear, criterion)
    if zero_grad:
        optimizer.zero_grad(set_to_none=True)
    with ops.side_stream_wgrads():          # first-order path: the small weight-gradient kernels overlap the backward chain
        loss.backward()
    if grad_sync is not None:
        grad_sync()
    optimizer.step()
    return loss.detach()


class GraphedFinetuneStep:
    """A fine-tune iteration (``step_fn(batch) -> loss``: forward, loss, backward, optimizer) captured in one CUDA graph.

    Pockets differ in size, so every batch is padded to one atom capacity (``data.pad_batch``: the padding atoms form an
    extra, edge-less graph whose readout row ``_readout`` drops); targets ride along in ``batch.extras``.  Works for any
    step without host synchronisation -- the first-order SchNet path has none (edge counts stay on the device).  The MD17
    force step (second order) trims its edge list on the host (``RadiusCSR.exact``) and is refused by the capture."""

    def __init__(self, step_fn, example_batches, optimizer, warmup=2, margin=1.02):
        from .data import pad_batch
        self._pad = pad_batch
        n_max = max(b.positions.size(0) for b in example_batches)
        same = len({b.positions.size(0) for b in example_batches}) == 1
        self.n_cap = n_max if same else -(-int(margin * n_max) // 128) * 128
        self.padded = not same
        b = self.pad(example_batches[0])
        extras = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.extras.items()}
        self.static = type(b)(b.x.clone(), b.positions.clone(), b.batch.clone(), None, None, b.num_graphs,
                              None if b.graph_ptr is None else b.graph_ptr.clone(), extras)
        dev = b.positions.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                optimizer.zero_grad(set_to_none=True)
                step_fn(self.static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = step_fn(self.static)

    @staticmethod
    def _tensors(b):
        """x, positions, batch + the tensor extras (targets); fine-tune batches carry no atom pairs."""
        return [b.x, b.positions, b.batch] + [b.extras[k] for k in sorted(b.extras) if torch.is_tensor(b.extras[k])]

    def pad(self, batch):
        if not self.padded or batch.extras.get("n_graphs_live") is not None:
            return batch
        return self._pad(type(batch)(batch.x, batch.positions, batch.batch, None, None, batch.n_graphs, batch.graph_ptr,
                                     dict(batch.extras)), self.n_cap, 0)

    def __call__(self, batch):
        for d, t in zip(self._tensors(self.static), self._tensors(self.pad(batch))):
            d.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.loss


class GraphedMD17Step:
    """The MD17 force-training iteration (finetune_md17.py::train :30-56: energy, force = -dE/dpos with create_graph=True,
    weighted L1, loss.backward() through the force -- double backward -- and the optimizer) replayed from a CUDA graph.

    The second-order path trims its edge list to the exact edge count, which is one host read (``RadiusCSR.exact``).  That
    read stays OUTSIDE the graph: per batch the neighbour search runs eagerly, its CSR (+ transpose) is copied into static
    buffers, and the captured forward / force / double-backward / optimizer sequence is replayed on them.  One graph is
    captured per distinct edge count (conformers of one molecule inside the cutoff share one complete graph, so MD17
    needs exactly one); ``max_graphs`` bounds the cache, beyond it the step runs eagerly."""

    def __init__(self, args, example_batch, model, graph_pred_linear, criterion, optimizer, grad_sync=None, max_graphs=4, **coeffs):
        self.args, self.model, self.lin, self.crit, self.opt, self.sync, self.coeffs = args, model, graph_pred_linear, criterion, optimizer, grad_sync, coeffs
        self.max_graphs = max_graphs
        self.graphs = {}
        self.capture_error = None
        self.template = example_batch

    def _structure(self, batch):
        g = ops.radius_csr(batch.positions.detach(), batch.batch, self.model.cutoff, num_graphs=batch.num_graphs)
        if ops.COMPOSED_PAIRS and ops.SHARE_PAIR_FILTERS and ops.FILTER_MODE != "simt":
            g.ensure_pairs()                                       # pair index on the device first: both counts come back ...
        return g.exact()                                           # ... in ONE host read (outside any capture)

    def _capture(self, batch, ge):
        e = ge.num_edges
        sg = ops.RadiusCSR(ge.n_atoms, e, ge.rowptr.clone(), ge.src.clone(), ge.tgt.clone(), None, batch.batch.clone(), ge.graph_ptr.clone())
        sg.t_rowptr, sg.t_eid, sg.t_tgt = ge.t_rowptr.clone(), ge.t_eid.clone(), ge.t_tgt.clone()
        sg._n_edges, sg._exact = e, sg
        if ge.pair_rowptr is not None:
            sg.pair_rowptr, sg.pair_of_edge, sg.pair_atoms = ge.pair_rowptr.clone(), ge.pair_of_edge.clone(), ge.pair_atoms.clone()
            sg._n_pairs = ge.num_pairs
        extras = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.extras.items() if k != "graph"}
        extras["graph"] = sg
        static = type(batch)(batch.x.clone(), batch.positions.detach().clone(), sg.batch, None, None, batch.num_graphs, sg.graph_ptr, extras)

        def run():
            # a FRESH leaf over the static coordinates every time: its AccumulateGrad node then belongs to the stream of this
            # run (a leaf reused from the warm-up would carry the warm-up stream's node into the capture and invalidate it)
            view = type(batch)(static.x, static.positions.detach().requires_grad_(), static.batch, None, None, static.n_graphs,
                               static.graph_ptr, static.extras)
            return md17_train_step(self.args, view, self.model, self.lin, self.crit, self.opt, grad_sync=self.sync, zero_grad=False,
                                   **self.coeffs)
        dev = static.positions.device
        self.opt.zero_grad(set_to_none=True)
        run()                                               # once on the caller's stream (library handles, allocator pools)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self.opt.zero_grad(set_to_none=True)
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # create_graph=True leaves reference cycles behind (the force graph references itself through its saved tensors): until
        # they are collected, the AccumulateGrad nodes of earlier iterations -- tied to THEIR streams -- stay alive and the
        # autograd engine would synchronise the capture stream with them, which invalidates the capture
        import gc
        gc.collect()
        graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            loss = run()
        gc.collect()
        return graph, static, sg, loss

    def __call__(self, batch):
        ge = self._structure(batch)
        e = (ge.num_edges, ge.num_pairs if ge.pair_rowptr is not None else -1)
        entry = self.graphs.get(e)
        if entry is None:
            if len(self.graphs) >= self.max_graphs or self.capture_error is not None:
                return md17_train_step(self.args, batch, self.model, self.lin, self.crit, self.opt, grad_sync=self.sync, **self.coeffs)
            try:
                entry = self.graphs[e] = self._capture(batch, ge)
            except RuntimeError as exc:                     # a refused capture is reported, and the step stays correct (eager)
                self.capture_error = exc
                torch.cuda.synchronize()
                return md17_train_step(self.args, batch, self.model, self.lin, self.crit, self.opt, grad_sync=self.sync, **self.coeffs)
        graph, static, sg, loss = entry
        with torch.no_grad():
            static.x.copy_(batch.x, non_blocking=True)
            static.positions.copy_(batch.positions.detach(), non_blocking=True)
            sg.batch.copy_(batch.batch, non_blocking=True)
            for k in ("y", "force"):
                static.extras[k].copy_(batch.extras[k], non_blocking=True)
            names = ("rowptr", "src", "tgt", "t_rowptr", "t_eid", "t_tgt", "graph_ptr")
            if sg.pair_rowptr is not None:
                names += ("pair_rowptr", "pair_of_edge", "pair_atoms")
            for name in names:
                getattr(sg, name).copy_(getattr(ge, name), non_blocking=True)
        graph.replay()
        return loss
