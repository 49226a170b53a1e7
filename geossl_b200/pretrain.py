"""The DDM pretraining step -- drop-in for ``perturb`` / ``do_DDM`` and the inner loop of ``train`` in
/root/reference/examples/pretrain_GeoSSL.py (:68-74, :179-212, :234-260), plus the data-parallel wrapper
the reference lacks (one process per GPU, one NCCL all-reduce of a flat fp32 gradient buffer per step).
"""
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from . import ops

# The reference's do_DDM reads these two module globals (pretrain_GeoSSL.py:207-208).
NCSN_model_01 = None
NCSN_model_02 = None


def set_ddm_heads(head_01, head_02):
    global NCSN_model_01, NCSN_model_02
    NCSN_model_01, NCSN_model_02 = head_01, head_02


def perturb(x, positions, mu, sigma, device_noise=False):
    """pretrain_GeoSSL.py:68-74.  Default: the reference's CPU ``torch.normal`` draw then H2D (RNG contract);
    ``device_noise=True`` draws on the GPU generator instead (different stream of random numbers)."""
    if device_noise:
        noise = torch.randn(positions.size(), device=positions.device, dtype=positions.dtype) * sigma + mu
    else:
        noise = torch.normal(mu, sigma, size=positions.size()).to(positions.device, non_blocking=True)
    return x, positions + noise


def _encode(args, model, x, positions, batch):
    n_graphs = getattr(batch, "n_graphs", None)
    if args.model_3d == "schnet":
        _, rep = model(x, positions, batch.batch, return_latent=True, num_graphs=n_graphs)
    elif args.model_3d == "painn":
        _, rep = model(x, positions, batch.radius_edge_index, batch.batch, return_latent=True, num_graphs=n_graphs,
                       assume_sorted=bool(getattr(batch, "extras", {}).get("rei_sorted", False)))
    else:
        raise Exception("3D model {} not included.".format(args.model_3d))
    return rep


def _encode_stacked(args, model, x_01, positions_01, x_02, positions_02, batch):
    """Both views through the encoder as ONE batch of 2B independent graphs (same weights, graphs never interact:
    radius graphs are per graph, schnet.py:91) -- halves the launches and doubles every kernel's row count."""
    n, b = positions_01.size(0), batch.num_graphs
    x = torch.cat([x_01, x_02])
    pos = torch.cat([positions_01, positions_02])
    bvec = torch.cat([batch.batch, batch.batch + b])
    if args.model_3d == "schnet":
        _, rep = model(x, pos, bvec, return_latent=True, num_graphs=2 * b)
    else:
        rei = batch.radius_edge_index
        _, rep = model(x, pos, torch.cat([rei, rei + n], dim=1), bvec, return_latent=True, num_graphs=2 * b,
                       assume_sorted=bool(getattr(batch, "extras", {}).get("rei_sorted", False)))
    return rep[:n], rep[n:]


def do_DDM(args, batch, model, criterion=None, mu=0.0, sigma=0.3, num_neg=1, heads=None, draws=None,
           positions_02=None, device_noise=False, stack_views=True):
    """Same call shape and return value ``(loss, 0)`` as the reference.  Extensions (all optional):
    ``heads=(head_01, head_02)`` instead of the module globals; ``draws=((level_1, eps_1), (level_2, eps_2))``
    and ``positions_02`` inject the random draws for parity tests; ``stack_views=False`` runs the encoder twice
    like the reference instead of once on the stacked 2B-graph batch."""
    head_01, head_02 = heads if heads is not None else (NCSN_model_01, NCSN_model_02)
    positions = batch.positions
    x_01 = batch.x[:, 0]
    positions_01 = positions
    if positions_02 is None:
        x_02, positions_02 = perturb(x_01, positions, mu, sigma, device_noise=device_noise)
    else:
        x_02 = x_01

    if stack_views and getattr(batch, "n_graphs", None) is not None:
        repr_01, repr_02 = _encode_stacked(args, model, x_01, positions_01, x_02, positions_02, batch)
    else:
        repr_01 = _encode(args, model, x_01, positions_01, batch)
        repr_02 = _encode(args, model, x_02, positions_02, batch)
    if getattr(args, "normalize", False):
        repr_01 = F.normalize(repr_01, dim=-1)
        repr_02 = F.normalize(repr_02, dim=-1)

    sei = batch.super_edge_index
    distance_01 = ops.pair_distance(positions_01, sei)
    distance_02 = ops.pair_distance(positions_02, sei)
    d1 = draws[0] if draws is not None else (None, None)
    d2 = draws[1] if draws is not None else (None, None)
    loss_01 = head_01(batch, repr_01, distance_02, noise_level=d1[0], distance_noise=d1[1])
    loss_02 = head_02(batch, repr_02, distance_01, noise_level=d2[0], distance_noise=d2[1])
    loss = (loss_01 + loss_02) / 2
    return loss, 0


def default_args(model_3d="schnet", normalize=False):
    return SimpleNamespace(model_3d=model_3d, normalize=normalize)


class FlatGradAllReduce:
    """Data-parallel gradient exchange: one all-reduce (sum) of a single flat fp32 buffer per step, then 1/R.
    519,556 floats (2.08 MB) for SchNet-DDM -- latency bound over NVLink, so one message, not one per tensor."""

    def __init__(self, params, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def __call__(self):
        if self.world == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, grads)
        self.dist.all_reduce(self.flat, op=self.dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / self.world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)


def broadcast_parameters(modules, src=0):
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src)


def train_step(args, batch, model, heads, optimizer, mu=0.0, sigma=0.3, grad_sync=None, device_noise=False,
               draws=None, positions_02=None):
    """One iteration of train() (pretrain_GeoSSL.py:234-260) without its per-step ``.item()`` host sync:
    returns the loss tensor (still on the device)."""
    loss, _ = do_DDM(args, batch, model, None, mu, sigma, heads=heads, device_noise=device_noise, draws=draws,
                     positions_02=positions_02)
    optimizer.zero_grad(set_to_none=True)
    with ops.side_stream_wgrads():          # small weight-gradient kernels overlap the backward chain; joined on exit
        loss.backward()
    if grad_sync is not None:
        grad_sync()
    optimizer.step()
    return loss.detach()


class GraphedTrainStep:
    """The whole training iteration (perturb, two encoder passes, two DDM heads, backward, gradient all-reduce,
    Adam) captured once in a CUDA graph and replayed per batch.

    Possible because nothing on the path synchronises with the host: the data-dependent edge count lives in device
    memory (``rowptr[N]``), every buffer is sized at a host-known capacity, ``num_graphs`` is carried by the batch,
    and the random draws use the graph-safe device generator.  Batches must have the captured shapes (same atom
    and pair counts, e.g. fixed-size molecules); ``matches(batch)`` tells, and callers fall back to ``train_step``.
    The optimizer must be capturable (``torch.optim.Adam(..., capturable=True)``).
    """

    def __init__(self, args, example_batch, model, heads, optimizer, mu=0.0, sigma=0.3, grad_sync=None, warmup=3):
        self.args, self.model, self.heads, self.optimizer = args, model, heads, optimizer
        b = example_batch
        dev = b.positions.device
        self.static = type(b)(b.x.clone(), b.positions.clone(), b.batch.clone(), b.super_edge_index.clone(),
                              None if b.radius_edge_index is None else b.radius_edge_index.clone(), b.num_graphs,
                              None if b.graph_ptr is None else b.graph_ptr.clone(), dict(getattr(b, "extras", {})))

        def run():
            loss, _ = do_DDM(args, self.static, model, None, mu, sigma, heads=heads, device_noise=True)
            with ops.side_stream_wgrads():
                loss.backward()
            if grad_sync is not None:
                grad_sync()
            optimizer.step()
            return loss.detach()

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                optimizer.zero_grad(set_to_none=True)
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = run()

    def matches(self, batch):
        s = self.static
        return (batch.positions.shape == s.positions.shape and batch.super_edge_index.shape == s.super_edge_index.shape
                and batch.num_graphs == s.num_graphs
                and (batch.radius_edge_index is None) == (s.radius_edge_index is None)
                and (s.radius_edge_index is None or batch.radius_edge_index.shape == s.radius_edge_index.shape))

    def load(self, batch, non_blocking=True):
        """Copy a batch (device or pinned host) into the graph's static input buffers."""
        s = self.static
        s.x.copy_(batch.x, non_blocking=non_blocking)
        s.positions.copy_(batch.positions, non_blocking=non_blocking)
        s.batch.copy_(batch.batch, non_blocking=non_blocking)
        s.super_edge_index.copy_(batch.super_edge_index, non_blocking=non_blocking)
        if s.radius_edge_index is not None:
            s.radius_edge_index.copy_(batch.radius_edge_index, non_blocking=non_blocking)

    def __call__(self, batch=None):
        if batch is not None:
            self.load(batch)
        self.graph.replay()
        return self.loss
