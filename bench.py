#!/usr/bin/env python
"""bench.py -- GeoSSL-DDM SchNet pretraining throughput (molecules/s) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA kernels via the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

A "step" is one full pretraining iteration of BASELINE.json configs[1] on one synthetic batch per GPU:
perturb -> SchNet encoder x2 (radius graph, 6 interactions) -> two DDM heads -> backward -> gradient
all-reduce (N>1) -> Adam.  Prints ONE JSON line (contract in the task statement / DESIGN.md section 6).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import torch  # noqa: E402

METRIC = "GeoSSL-DDM SchNet train molecules/s at 1/2/4/8 B200; cfconv % HBM roofline"
CFG = dict(batch_per_gpu=256, atoms=30, cutoff=10.0, num_gaussians=50, hidden=128, filters=128, interactions=6,
           sigma_levels=50, anneal_power=2.0, pos_sigma=0.3, lr=5e-4)
NCU_TRAFFIC_CFCONV_FWD = 123.69e6 + 6.86e6     # bytes per launch of the bench workload, row-gather kernel (profiles/r01_v32_ncu_full.txt)
NCU_TRAFFIC_CFCONV_PAIRS = 121.96e6 + 4.25e6   # pair-centric kernel: dram read + mean write per launch (profiles/r02_v22_ncu_full.txt)
PAINN_METRIC = "GeoSSL-DDM PaiNN train molecules/s (BASELINE configs[2], secondary)"
PAINN_WORKLOAD = ("configs[2]: PaiNN GeoSSL-DDM pretraining step, F=128, 3 interactions, 20 RBF, cutoff 5 A, synthetic Molecule3D-shaped "
                  "conformers, batch 256 per GPU x 30 atoms, data-parallel")
WORKLOAD = ("configs[1]: SchNet GeoSSL-DDM pretraining step, synthetic Molecule3D-shaped conformers, "
            "batch 256 per GPU x 30 atoms, cutoff 10 A, 50 RBF, hidden 128, 6 interactions, data-parallel")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="geossl_b200", choices=["geossl_b200", "reference"])
    ap.add_argument("--pool", type=int, default=8, help="distinct synthetic batches cycled through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timers", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    ap.add_argument("--atoms-max", type=int, default=0,
                    help="secondary variant of configs[1]: n ~ U{atoms..atoms_max} atoms per molecule (rows beyond 32 neighbours "
                         "are truncated, shapes differ per batch => eager launches)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: total molecules per step, split evenly over the GPUs (default 0 = weak scaling, 256 per GPU)")
    ap.add_argument("--workload", default="ddm", choices=["ddm", "md17", "lba"],
                    help="ddm = the DDM pretraining step (headline); md17 = configs[3]: SchNet energy + autograd-force fine-tune step "
                         "(double backward) on aspirin-size conformers; lba = configs[4]: SchNet fine-tune step on ~600-atom pockets, "
                         "cutoff 6 A, batch 32")
    ap.add_argument("--model", default="schnet", choices=["schnet", "painn"],
                    help="schnet = the headline workload (configs[1]); painn = configs[2] (F=128, 3 interactions, 20 RBF, cutoff 5 A)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference(steps, warmup, sample_graphs=None, threads=None, budget_s=None, model_3d="schnet"):
    """The reference's own CPU implementation of the step on this box's host cores.

    kind "reference": the UNMODIFIED reference modules (Geom3D/models/schnet.py | painn.py, examples/NCSN.py; placed in
    the git-ignored oracle/_ref by oracle/make_ref.py, imported under oracle/shims) driven with do_DDM's call sequence
    (oracle/reference_loader.ddm_step) + loss.backward() + torch.optim.Adam, as train() does (pretrain_GeoSSL.py:249-260).
    kind "port": the functional restatement oracle/models.py -- only when oracle/_ref is absent.
    Each step is one full batch of the bench workload (``sample_graphs`` molecules, default the whole 256-molecule batch);
    ``budget_s`` stops early (after >= 2 timed steps) if the host is too slow, and the line says how many steps ran."""
    from oracle import reference_loader
    from geossl_b200.data import synthetic_batch
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sample_graphs = sample_graphs or CFG["batch_per_gpu"]
    torch.manual_seed(42)
    painn = model_3d == "painn"
    kind = "reference" if reference_loader.available() else "port"
    if kind == "reference":
        SchNet, PaiNN, NCSN = reference_loader.load()
    else:
        from geossl_b200.Geom3D.models import PaiNN, SchNet
        from geossl_b200.NCSN import NCSN_version_03 as NCSN
        from oracle import models as O
    if painn:
        model = PaiNN(n_atom_basis=CFG["hidden"], n_interactions=3, n_rbf=20, cutoff=5.0, max_z=9, n_out=1, readout="add")
    else:
        model = SchNet(hidden_channels=CFG["hidden"], num_filters=CFG["filters"], num_interactions=CFG["interactions"],
                       num_gaussians=CFG["num_gaussians"], cutoff=CFG["cutoff"], node_class=9)
    heads = [NCSN(CFG["hidden"], 10, 0.01, CFG["sigma_levels"], "symmetry", CFG["anneal_power"]) for _ in range(2)]
    if kind == "reference":
        leaves = [p for m in [model] + heads for p in m.parameters() if p.requires_grad]
    else:
        sd = {k: v.clone().requires_grad_(v.is_floating_point() and v.dtype == torch.float32) for k, v in model.state_dict().items()}
        sdh = [{k: v.clone().requires_grad_(k != "sigmas") for k, v in h.state_dict().items()} for h in heads]
        leaves = [v for v in sd.values() if v.requires_grad] + [v for d in sdh for v in d.values() if v.requires_grad]
    opt = torch.optim.Adam(leaves, lr=CFG["lr"])
    from oracle.radius import radius_graph
    times = []
    t_start = time.perf_counter()
    for it in range(warmup + steps):
        b = synthetic_batch(sample_graphs, CFG["atoms"], seed=1000 + it)
        if painn:                      # dataset-time radius graph (datasets_3D_Radius.py:120): outside the timed step
            b.radius_edge_index = radius_graph(b.positions, 5.0, b.batch)
        t0 = time.perf_counter()
        if kind == "reference":
            loss = reference_loader.ddm_step(model_3d, model, heads, b, 0.0, CFG["pos_sigma"])
        else:
            _, pos2 = O.perturb(None, b.positions, 0.0, CFG["pos_sigma"])
            if painn:
                enc = lambda z, p: O.painn_forward(sd, z, p, b.radius_edge_index, b.batch, readout="add")[1]
            else:
                enc = lambda z, p: O.schnet_forward(sd, z, p, b.batch, cutoff=CFG["cutoff"])[1]
            n_pairs = b.super_edge_index.shape[1]
            draws = [(torch.randint(0, CFG["sigma_levels"], (sample_graphs,)), torch.randn(n_pairs, 1)) for _ in range(2)]
            loss, _ = O.ddm_loss(enc, sdh[0], sdh[1], b.x[:, 0], b.positions, pos2, b.batch, b.super_edge_index, draws[0], draws[1],
                                 CFG["anneal_power"])
        opt.zero_grad()
        loss.backward()
        opt.step()
        float(loss.detach())
        if it >= warmup:
            times.append(time.perf_counter() - t0)
            if budget_s and len(times) >= 2 and time.perf_counter() - t_start > budget_s:
                break
    total = sum(times)
    what = ("unmodified reference modules (oracle/_ref) under oracle/shims" if kind == "reference"
            else "oracle/models.py port (oracle/_ref absent)")
    return {"value": sample_graphs * len(times) / total, "unit": "molecules/s", "cores": threads, "kind": kind,
            "sample": f"{len(times)} steps x {sample_graphs} molecules x {CFG['atoms']} atoms = the bench batch "
                      f"({what}, torch CPU fp32, {warmup} warm-up, fwd+bwd+Adam)",
            "ms_per_step": 1e3 * total / len(times), "steps": len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    painn = args.model == "painn"
    cb = cpu_reference(args.steps, args.warmup, budget_s=240.0, model_3d=args.model)
    line = {"metric": PAINN_METRIC if painn else METRIC, "value": cb["value"], "unit": "molecules/s", "n_gpus": args.gpus,
            "steps": cb["steps"], "warmup": args.warmup,
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": PAINN_WORKLOAD if painn else WORKLOAD, **CFG, "device": "host CPU",
                       "reference_step": "one 256-molecule batch per step on rank 0's host cores"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(json.dumps(line))


# ------------------------------------------------------------------------------------------- product arm
def run_product(args):
    import torch.distributed as dist
    from geossl_b200 import _lib, ops
    from geossl_b200.Geom3D.models import SchNet
    from geossl_b200.NCSN import NCSN_version_03
    from geossl_b200.data import synthetic_batch
    from geossl_b200.pretrain import FlatGradAllReduce, GraphedTrainStep, broadcast_parameters, default_args, train_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- geossl_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    torch.manual_seed(42)
    if args.model == "painn":
        from geossl_b200.Geom3D.models import PaiNN
        model = PaiNN(n_atom_basis=CFG["hidden"], n_interactions=3, n_rbf=20, cutoff=5.0, max_z=9, n_out=1, readout="add").to(dev)
    else:
        model = SchNet(hidden_channels=CFG["hidden"], num_filters=CFG["filters"], num_interactions=CFG["interactions"],
                       num_gaussians=CFG["num_gaussians"], cutoff=CFG["cutoff"], node_class=9).to(dev)
    heads = [NCSN_version_03(CFG["hidden"], 10, 0.01, CFG["sigma_levels"], "symmetry", CFG["anneal_power"]).to(dev) for _ in range(2)]
    broadcast_parameters([model] + heads)
    groups = [{"params": model.parameters()}] + [{"params": [p for p in h.parameters() if p.requires_grad]} for h in heads]
    opt = torch.optim.Adam(groups, lr=CFG["lr"], fused=True, capturable=not args.no_graph)
    sync = FlatGradAllReduce([p for g in groups for p in g["params"]]) if world > 1 else None
    if os.environ.get("GEOSSL_CFCONV_PAIRS"):             # A/B only: "0" = row-gather cfconv kernels, else the tuning code
        from geossl_b200 import ops as _o
        _o.CFCONV_PAIRS = os.environ["GEOSSL_CFCONV_PAIRS"] != "0"
        _o.CFCONV_PAIRS_TUNING = int(os.environ["GEOSSL_CFCONV_PAIRS"])
    if os.environ.get("GEOSSL_WGRAD_MAIN_STREAM"):        # A/B only: weight-gradient kernels on the main stream (same work, no overlap)
        import contextlib
        import geossl_b200.pretrain as _pt
        _pt.ops.side_stream_wgrads = contextlib.nullcontext
    if os.environ.get("GEOSSL_PAIR_OWNER_SMALL"):         # tuning only
        from geossl_b200 import ops as _o
        _o.PAIR_OWNER_SMALL = os.environ["GEOSSL_PAIR_OWNER_SMALL"] != "0"
    targs = default_args(args.model)
    torch.manual_seed(1234 + rank)

    B = CFG["batch_per_gpu"]
    if args.global_batch:
        assert args.global_batch % world == 0, "--global-batch must be divisible by the number of GPUs"
        B = args.global_batch // world
    host_pool = [synthetic_batch(B, CFG["atoms"] if not args.atoms_max else 10, args.atoms_max or None, seed=10_000 * rank + i)
                 for i in range(args.pool)]
    if args.atoms_max:
        # variable-size molecules: every batch is padded to ONE capacity (data.pad_batch) so that a single captured graph
        # serves the stream; capacity = the largest batch of the pool + 2 %, rounded up
        from geossl_b200.data import pad_batch
        n_cap = -(-int(1.02 * max(hb.positions.size(0) for hb in host_pool)) // 128) * 128
        p_cap = -(-int(1.02 * max(hb.super_edge_index.size(1) for hb in host_pool)) // 1024) * 1024
        live_atoms = [hb.positions.size(0) for hb in host_pool]
        host_pool = [pad_batch(hb, n_cap, p_cap, max_graph_atoms_cap=args.atoms_max) for hb in host_pool]
    if args.model == "painn":
        # dataset-time radius graph (datasets_3D_Radius.py:120) on the clean coordinates, reused for both views
        from geossl_b200 import ops as _ops
        from geossl_b200.data import pad_batch
        for hb in host_pool:
            g0 = _ops.radius_csr(hb.positions.to(dev), hb.batch.to(dev), 5.0, num_graphs=B, transpose=False)
            hb.radius_edge_index = g0.edge_index.cpu()
            hb.extras["rei_sorted"] = True
        # edge lists differ in length per batch: pad every batch to one edge capacity so that ONE captured graph serves all
        live_edges = [hb.radius_edge_index.size(1) for hb in host_pool]
        e_cap = -(-int(1.03 * max(live_edges)) // 1024) * 1024
        host_pool = [pad_batch(hb, hb.positions.size(0), hb.super_edge_index.size(1), e_cap) for hb in host_pool]
    host_pool = [hb.pin_memory() for hb in host_pool]
    dev_pool = [b.to(dev) for b in host_pool]
    h2d_bytes = sum(t.numel() * t.element_size() for t in (host_pool[0].x, host_pool[0].positions, host_pool[0].batch,
                                                           host_pool[0].super_edge_index, host_pool[0].radius_edge_index)
                    if t is not None)

    def step(batch):
        return train_step(targs, batch, model, heads, opt, 0.0, CFG["pos_sigma"], grad_sync=sync, device_noise=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(n):
            fn(i)
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up (allocator, cuBLAS handles, kernel attribute setup), then capture the step in a CUDA graph
    for i in range(max(args.warmup, 3)):
        step(dev_pool[i % args.pool])
    barrier()
    eager_step = step
    if not args.no_graph:
        graphed = GraphedTrainStep(targs, dev_pool[0], model, heads, opt, 0.0, CFG["pos_sigma"], grad_sync=sync)
        step = graphed
        for i in range(max(10, args.warmup)):                # replays of the captured graph are part of the warm-up too
            step(dev_pool[i % args.pool])
        barrier()

    # ---- (1) device-resident throughput + clocks + per-kernel timers
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda i: step(dev_pool[i % args.pool]), args.steps)
    value = world * B * args.steps / (ms / 1e3)
    # per-kernel durations, measured INSIDE a replayed graph: the step is captured a second time with external event-record
    # nodes around the named kernels (ops.KERNEL_TIMERS in-graph mode) and replayed n_k times; the launch counter runs
    # during that capture (a replay launches exactly the captured kernels).  Eager runs (--no-graph, variable shapes) fall
    # back to eager event brackets.
    ktimes = {}
    n_k = min(args.steps, 10)
    timer_names = ("cfconv_fwd", "filter_fwd", "filter_bwd", "cfconv_bwd_x", "ddm_head_fwd", "ddm_head_bwd", "ddm_head_fused",
                   "linear_fwd", "linear_dgrad", "linear_wgrad", "painn_message_fwd", "painn_message_bwd", "dense_fwd", "dense_dgrad",
                   "dense_wgrad", "dense_chain_fwd", "dense_chain_bwd")
    if step is not eager_step:
        _lib.launch_count(reset=True)
        probe = GraphedTrainStep(targs, dev_pool[0], model, heads, opt, 0.0, CFG["pos_sigma"], grad_sync=sync, warmup=0,
                                 kernel_timers=None if args.no_kernel_timers else timer_names)
        launches = _lib.launch_count() * args.steps
        acc = None
        for i in range(n_k):
            probe(dev_pool[i % args.pool])
            if not args.no_kernel_timers:
                acc = ops.KERNEL_TIMERS.collect_replay(acc)
        barrier()
        if acc:
            ktimes = ops.KERNEL_TIMERS.stats(acc)
        kernel_timing = f"event-record nodes inside a replayed CUDA graph of the step, {n_k} replays"
    else:
        _lib.launch_count(reset=True)
        if not args.no_kernel_timers:
            ops.KERNEL_TIMERS.enable(timer_names)
        ms_eager = timed(lambda i: eager_step(dev_pool[i % args.pool]), n_k)
        launches = _lib.launch_count() * args.steps // n_k
        if not args.no_kernel_timers:
            ktimes = ops.KERNEL_TIMERS.collect()
            ops.KERNEL_TIMERS.disable()
        kernel_timing = f"CUDA-event brackets over {n_k} eager steps ({ms_eager / n_k:.2f} ms/step eager)"
    clocks = sampler.stop() if rank == 0 else None

    # ---- (2) end to end: host (pinned) batches -> H2D every step -> step -> D2H loss every step
    # The loss of EVERY step is read back to pinned host memory, one step behind the launches (the host waits for step
    # i-1's loss after it has enqueued step i), so the launch latency of a step hides behind the previous step's compute.
    sinks = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    landed = [torch.cuda.Event(), torch.cuda.Event()]
    losses_seen = []

    piped = not args.no_graph
    staged = [-1, -1]                                        # which step's batch sits in each staging slot

    def e2e_step(i):
        if not piped:
            loss = step(host_pool[i % args.pool].to(dev, non_blocking=True))
        else:
            # pinned host -> staging slot on a copy stream (overlaps the previous step) -> static inputs -> replay;
            # every step's batch crosses PCIe inside the timed region, the next one while this step computes
            if staged[i & 1] != i:
                step.prefetch(host_pool[i % args.pool], i & 1)
                staged[i & 1] = i
            loss = step.run_prefetched(i & 1)
            step.prefetch(host_pool[(i + 1) % args.pool], (i + 1) & 1)
            staged[(i + 1) & 1] = i + 1
        sinks[i & 1].copy_(loss.view(1), non_blocking=True)
        landed[i & 1].record()
        if i > 0:
            landed[(i - 1) & 1].synchronize()
            losses_seen.append(float(sinks[(i - 1) & 1][0]))

    def e2e_run(i):
        e2e_step(i)
        if i == args.steps - 1:                              # last step: its loss is read inside the timed region too
            landed[i & 1].synchronize()
            losses_seen.append(float(sinks[i & 1][0]))

    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    losses_seen.clear()
    staged[0] = staged[1] = -1                               # the timed region stages its own first batch
    ms_e2e = timed(e2e_run, args.steps)
    assert len(losses_seen) == args.steps and all(v == v for v in losses_seen), "every step's loss must reach the host"
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    if rank != 0:
        _finish(world)
        return

    # ---- roofline of the cfconv forward kernel (BASELINE metric) + the other hot kernels
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
    pk = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = {**json.load(open(pk)), "src": "measured"}
    if args.model == "painn":
        # roofline of the message kernel (north_star (3)): algorithmic HBM bytes = every per-atom row once (ctx 3F, mu in/out
        # 3F + 3F, q in/out) + the per-edge scalars and indices; the kernel is in fact fp32-FMA bound (the 3F filter values of
        # an edge are rebuilt from 20 rbf values: 2*20*3F + 10F FLOP per edge), so the FMA fraction is reported beside it.
        F_, R = CFG["hidden"], 20
        n_atoms2, n_edges2 = 2 * dev_pool[0].positions.size(0), 2 * live_edges[0]
        msg_bytes = 4 * F_ * n_atoms2 * (3 + 3 + 3 + 1 + 1) + n_edges2 * (4 * 5 + 8)
        msg_flops = n_edges2 * (2 * R * 3 * F_ + 2 * 5 * F_)
        roof, others = None, {}
        if "painn_message_fwd" in ktimes:
            t = ktimes["painn_message_fwd"]["mean_ms"] / 1e3
            roof = {"kernel": "painn_message_fwd_kernel<128> (PaiNN scalar/vector message, filter rebuilt per edge)", "bound": "hbm",
                    "achieved": msg_bytes / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": msg_bytes / t / 1e9 / peaks["hbm_gbs"],
                    "traffic": None, "peak_source": peaks["src"], "algorithmic_bytes_per_launch": msg_bytes, "mean_ms": 1e3 * t,
                    "fp32_fma": {"achieved_tflops": msg_flops / t / 1e12, "flops_per_launch": msg_flops,
                                 "note": "the kernel is bound by the fp32 FMA pipe / shared-memory reads of the filter slice, not by HBM"},
                    "share_of_step": ktimes["painn_message_fwd"]["total_ms"] / n_k / (ms / args.steps)}
        for k in ("painn_message_bwd", "ddm_head_fused", "ddm_head_fwd", "ddm_head_bwd", "dense_fwd", "dense_dgrad", "dense_wgrad"):
            if k in ktimes:
                others[k] = {"mean_ms": ktimes[k]["mean_ms"], "launches_per_step": ktimes[k]["n"] // n_k,
                             "share_of_step": ktimes[k]["total_ms"] / n_k / (ms / args.steps)}
        cpu = None
        if not args.no_cpu_baseline:
            cb = cpu_reference(3, 1, budget_s=40.0, model_3d="painn")
            cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {"metric": PAINN_METRIC, "value": value, "unit": "molecules/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": PAINN_WORKLOAD, **CFG, "global_batch": world * B, "parallelism": f"dp{world}",
                           "launch": "eager" if args.no_graph else "whole step captured in one CUDA graph (edge lists padded to "
                                     f"{e_cap} columns, {min(live_edges)}..{max(live_edges)} live)",
                           "kernel_timing": kernel_timing, "edges_per_view": live_edges[0],
                           "l2": f"{args.pool} distinct batches cycled"},
                "e2e": {"value": e2e_value, "unit": "molecules/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_other_kernels": others, "cpu_baseline": cpu}
        _emit(json.dumps(line))
        _finish(world)
        return

    # one encoder launch processes BOTH views stacked as 2B graphs (pretrain._encode_stacked): size the work from that graph
    b0 = dev_pool[0]
    pos2 = torch.cat([b0.positions, b0.positions + CFG["pos_sigma"] * torch.randn_like(b0.positions)])
    g = ops.radius_csr(pos2, torch.cat([b0.batch, b0.batch + B]), CFG["cutoff"], num_graphs=2 * B,
                       max_graph_atoms=b0.extras.get("max_graph_atoms"))
    n_atoms, n_edges, F_, G = pos2.shape[0], g.num_edges, CFG["filters"], CFG["num_gaussians"]
    n_pairs = int(b0.extras["n_pairs_live"]) if args.atoms_max else b0.super_edge_index.shape[1]
    # filter rows: one per undirected atom pair when the two directions share it (ops.SHARE_PAIR_FILTERS), else one per edge
    shared = ops.SHARE_PAIR_FILTERS and ops.FILTER_MODE != "simt"
    n_rows_w = int(g.ensure_pairs().n_pairs_dev.item()) if shared else n_edges
    pair_kernel = shared and ops._pairs_kernel_applies(g)       # geossl_cfconv_pairs: one CTA per molecule, rows read once
    # algorithmic bytes of cfconv forward: every filter row once + x in + m out + the index arrays the kernel reads
    # (row gather: src ids (+ the pair map) + rowptr; pair-centric: the (s, t) pair records + pair_rowptr + graph_ptr)
    idx_bytes = (8 * n_rows_w + 4 * (n_atoms + 1) + 4 * (2 * B + 1)) if pair_kernel else (4 * n_edges * (2 if shared else 1) + 4 * (n_atoms + 1))
    cf_bytes = 4 * F_ * n_rows_w + 2 * 4 * F_ * n_atoms + idx_bytes
    cf_bytes_per_edge_form = 4 * F_ * n_edges + 2 * 4 * F_ * n_atoms + 4 * n_edges + 4 * (n_atoms + 1)   # 551 B/edge
    roof = None
    others = {}
    if "cfconv_fwd" in ktimes:
        t = ktimes["cfconv_fwd"]["mean_ms"] / 1e3
        ach = cf_bytes / t / 1e9
        full_size = shared and n_edges > 400_000 and not args.atoms_max
        roof = {"kernel": ("cfconv_pairs_kernel<2, 8, false> (cfconv forward, F = 128, one CTA per molecule)" if pair_kernel
                           else "cfconv_gather_async_kernel<false> (cfconv forward, F = 128)"), "bound": "hbm", "achieved": ach,
                "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"],
                "traffic": (NCU_TRAFFIC_CFCONV_PAIRS if pair_kernel else NCU_TRAFFIC_CFCONV_FWD) if full_size else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, "
                                  + ("profiles/r02_v22_ncu_full.txt" if pair_kernel else "profiles/r01_v32_ncu_full.txt"),
                "peak_source": peaks["src"],
                "algorithmic_bytes_per_launch": cf_bytes, "filter_rows_per_launch": n_rows_w, "mean_ms": 1e3 * t,
                "launches_timed": ktimes["cfconv_fwd"]["n"],
                "per_edge_form": {"bytes_per_launch": cf_bytes_per_edge_form, "achieved": cf_bytes_per_edge_form / t / 1e9,
                                  "note": "SURVEY 8d figure (551 B/edge, one filter row per DIRECTED edge) over the same duration: "
                                          "the rate a per-edge kernel would need to match this one; not a DRAM rate"}}
        flops = {"filter_fwd": n_rows_w * (2 * G * F_ + 2 * F_ * F_),
                 "filter_bwd": n_rows_w * (2 * (2 * F_ * F_) + 2 * G * F_),      # 78,336 FLOP/row (no recompute counted)
                 "ddm_head_fwd": n_pairs * 50_048, "ddm_head_bwd": n_pairs * 3 * 50_048,
                 "ddm_head_fused": n_pairs * 3 * 50_048}       # one pass = forward + backward (2x forward FLOPs) of one head
        for k, fl in flops.items():
            if k in ktimes:
                tt = ktimes[k]["mean_ms"] / 1e3
                others[k] = {"bound": "tensor", "achieved": fl / tt / 1e12, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                             "frac": fl / tt / 1e12 / peaks["bf16_tflops_sustained"], "mean_ms": 1e3 * tt,
                             "share_of_step": ktimes[k]["total_ms"] / n_k / (ms / args.steps),
                             "note": ("3 split-precision MMAs per product (fp32-grade); FLOPs of the rows actually processed "
                                      f"({n_rows_w} filter rows for {n_edges} edges) against the bf16 tensor peak")
                             if k.startswith("filter") else "3 split-precision MMAs per product; useful FLOPs against the bf16 tensor peak"}
        if "cfconv_bwd_x" in ktimes:
            tt = ktimes["cfconv_bwd_x"]["mean_ms"] / 1e3
            bb = cf_bytes if pair_kernel else 4 * F_ * n_rows_w + 2 * 4 * F_ * n_atoms + 4 * n_edges * (3 if shared else 2) + 4 * (n_atoms + 1)
            others["cfconv_bwd_x"] = {"bound": "hbm", "achieved": bb / tt / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                      "frac": bb / tt / 1e9 / peaks["hbm_gbs"], "mean_ms": 1e3 * tt,
                                      "share_of_step": ktimes["cfconv_bwd_x"]["total_ms"] / n_k / (ms / args.steps)}
        roof["share_of_step"] = ktimes["cfconv_fwd"]["total_ms"] / n_k / (ms / args.steps)
        for k in ("linear_fwd", "linear_dgrad", "linear_wgrad", "dense_chain_fwd", "dense_chain_bwd"):
            if k in ktimes:
                tt = ktimes[k]["mean_ms"] / 1e3
                others[k] = {"bound": "latency", "mean_ms": 1e3 * tt, "launches_per_step": ktimes[k]["n"] // n_k,
                             "share_of_step": ktimes[k]["total_ms"] / n_k / (ms / args.steps),
                             "note": "128->128 atom-wise layers on tcgen05 (0.5 GFLOP per layer; dense_chain_* = 2-3 layers per launch); "
                                     "linear_wgrad runs on the side stream"}

    # whole-step floor: every kernel at its own roofline, back to back (the kernels are serial on the critical path).
    # Tensor work is counted with the 3 split-precision MMAs per product the fp32-grade path issues.
    L = CFG["interactions"]
    fl_filter = n_rows_w * (2 * G * F_ + 2 * F_ * F_ + 2 * (2 * F_ * F_) + 2 * G * F_) * L            # fwd + bwd, per step
    fl_dense = n_atoms * 2 * F_ * F_ * 3 * (3 * L + 2)                                                # fwd + dgrad + wgrad
    fl_head = n_pairs * 50_048 * 3 * 2                                                                # two heads, fwd + bwd
    bytes_hbm = L * (cf_bytes + (cf_bytes if pair_kernel else 4 * F_ * n_rows_w + 2 * 4 * F_ * n_atoms + 4 * n_edges * (3 if shared else 2))   # cfconv fwd + dx
                     + 4 * F_ * n_rows_w)                                                             # filter rows written once
    t_tensor = 3 * (fl_filter + fl_dense + fl_head) / (peaks["bf16_tflops_sustained"] * 1e12)
    t_hbm = bytes_hbm / (peaks["hbm_gbs"] * 1e9)
    floor = {"floor_ms": 1e3 * (t_tensor + t_hbm), "tensor_ms": 1e3 * t_tensor, "hbm_ms": 1e3 * t_hbm,
             "frac_of_step": (t_tensor + t_hbm) / (ms / args.steps / 1e3),
             "flops_per_step": {"filter_mlp": fl_filter, "atomwise_dense": fl_dense, "ddm_heads": fl_head, "mma_per_product": 3},
             "hbm_bytes_per_step": bytes_hbm,
             "note": "sum of per-kernel roofline times: split-precision tensor work at the sustained bf16 peak + the materialised-"
                     "filter traffic (written once, read by cfconv forward and by dx) at the measured HBM peak"}

    cpu = None
    if not args.no_cpu_baseline:
        cb = cpu_reference(3, 1, budget_s=40.0)
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": METRIC, "value": value, "unit": "molecules/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if not args.atoms_max else
                       WORKLOAD + f" -- VARIABLE-SIZE VARIANT: 10..{args.atoms_max} atoms per molecule, batches padded to the capacity "
                       f"({n_cap} atoms, {p_cap} pairs; {min(live_atoms)}..{max(live_atoms)} live atoms) of one captured graph", **{**CFG, "batch_per_gpu": B}, "global_batch": world * B, "atoms_per_launch": n_atoms, "edges_per_launch": n_edges, "views_stacked": 2,
                       "pairs": n_pairs, "parallelism": f"dp{world}", "optimizer": "torch.optim.Adam(fused)", "launch": "eager" if args.no_graph else "whole step captured in one CUDA graph",
                       "kernel_timing": kernel_timing,
                       "filter_rows_per_launch": n_rows_w,
                       "l2": f"{args.pool} distinct batches cycled; per-step working set (6 x {4 * F_ * n_rows_w / 1e6:.0f} MB filter "
                             "tensors) exceeds the 126 MB L2", "position_noise": "device generator"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "molecules/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "roofline": roof, "roofline_other_kernels": others, "step_floor": floor, "cpu_baseline": cpu}
    _emit(json.dumps(line))
    _finish(world)


# ------------------------------------------------------------------------------------------- fine-tune workloads (configs[3], [4])
# (md17: 21 atoms in a 5.6 A cube -- like aspirin, every pair lies inside the 10 A cutoff, so all conformers share one complete graph)
FT = {"md17": dict(batch=256, atoms=21, density=0.12, cutoff=10.0, unit="conformers/s",
                   metric="SchNet MD17-shaped force fine-tune conformers/s (BASELINE configs[3], secondary)",
                   workload="configs[3]: SchNet MD17-shaped force fine-tune step (energy + autograd forces, double backward through cfconv, "
                            "L1 losses 0.05/0.95, Adam), synthetic aspirin-size conformers (21 atoms), batch 256 per GPU"),
      "lba": dict(batch=32, atoms=600, density=0.05, cutoff=6.0, unit="pockets/s",
                  metric="SchNet LBA-shaped pocket fine-tune pockets/s (BASELINE configs[4], secondary)",
                  workload="configs[4]: SchNet LBA-shaped fine-tune step (readout -> Linear -> MSE, Adam) on synthetic protein-ligand "
                           "pockets (~600 atoms, cutoff 6 A: every neighbour row truncates at 32/33), batch 32 per GPU")}


def _ft_batch(kind, seed):
    from geossl_b200.data import synthetic_batch
    f = FT[kind]
    lo, hi = (f["atoms"], None) if kind == "md17" else (560, 640)
    b = synthetic_batch(f["batch"], lo, hi, seed=seed, density=f["density"], with_pairs=False)
    g = torch.Generator().manual_seed(seed + 7)
    b.extras["y"] = torch.randn(f["batch"], generator=g)
    if kind == "md17":
        b.extras["force"] = torch.randn(b.positions.shape, generator=g)
    return b


def ft_cpu_reference(kind, steps, warmup, budget_s=None):
    """The reference's CPU path of the fine-tune step: unmodified reference SchNet (oracle/_ref) + nn.Linear(128,1), loss
    as finetune_md17.py:30-51 / finetune_lba.py:33-47 (oracle/reference_loader.{md17,lba}_step), backward, Adam."""
    from oracle import reference_loader
    f = FT[kind]
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(42)
    if reference_loader.available():
        SchNet, _, _ = reference_loader.load()
        how = "reference"
    else:
        raise SystemExit("bench.py: the fine-tune reference arm needs oracle/_ref (run oracle/make_ref.py in the build container)")
    model = SchNet(hidden_channels=128, num_filters=128, num_interactions=6, num_gaussians=50, cutoff=f["cutoff"], node_class=9)
    lin = torch.nn.Linear(128, 1)
    opt = torch.optim.Adam(list(model.parameters()) + list(lin.parameters()), lr=5e-4)
    times, t_start = [], time.perf_counter()
    for it in range(warmup + steps):
        b = _ft_batch(kind, 2000 + it)
        t0 = time.perf_counter()
        loss = (reference_loader.md17_step(model, lin, b, b.extras["y"], b.extras["force"]) if kind == "md17"
                else reference_loader.lba_step(model, lin, b, b.extras["y"]))
        opt.zero_grad()
        loss.backward()
        opt.step()
        float(loss.detach())
        if it >= warmup:
            times.append(time.perf_counter() - t0)
            if budget_s and len(times) >= 2 and time.perf_counter() - t_start > budget_s:
                break
    total = sum(times)
    return {"value": f["batch"] * len(times) / total, "unit": f["unit"], "cores": os.cpu_count(), "kind": how,
            "sample": f"{len(times)} steps x the bench batch ({f['batch']} x ~{f['atoms']} atoms; unmodified reference SchNet under "
                      f"oracle/shims, torch CPU fp32, {warmup} warm-up, fwd+bwd+Adam)",
            "ms_per_step": 1e3 * total / len(times), "steps": len(times)}


def run_finetune(args):
    """configs[3] / configs[4]: same JSON contract as the headline line, on the fine-tune callers of the encoder."""
    kind, f = args.workload, FT[args.workload]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    base = {"metric": f["metric"], "unit": f["unit"], "n_gpus": args.gpus, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    cfg = {"workload": f["workload"], "batch_per_gpu": f["batch"], "atoms": f["atoms"], "cutoff": f["cutoff"], "num_gaussians": 50,
           "hidden": 128, "interactions": 6, "lr": 5e-4}
    if args.impl == "reference":
        if rank != 0:
            return
        cb = ft_cpu_reference(kind, args.steps, args.warmup, budget_s=240.0)
        _emit(json.dumps({**base, "value": cb["value"], "steps": cb["steps"], "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
                          "impl": "reference", "config": {**cfg, "device": "host CPU"},
                          "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                          "e2e": {"value": cb["value"], "unit": f["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return
    import torch.distributed as dist
    from geossl_b200 import _lib, ops
    from geossl_b200.Geom3D.models import SchNet
    from geossl_b200.finetune import GraphedFinetuneStep, GraphedMD17Step, lba_train_step, md17_train_step
    from geossl_b200.pretrain import FlatGradAllReduce, broadcast_parameters, default_args
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- geossl_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    torch.manual_seed(42)
    model = SchNet(hidden_channels=128, num_filters=128, num_interactions=6, num_gaussians=50, cutoff=f["cutoff"], node_class=9).to(dev)
    lin = torch.nn.Linear(128, 1).to(dev)
    broadcast_parameters([model, lin])
    params = list(model.parameters()) + list(lin.parameters())
    opt = torch.optim.Adam(params, lr=5e-4, fused=True, capturable=not args.no_graph)
    sync = FlatGradAllReduce(params) if world > 1 else None
    targs = default_args("schnet")
    crit = torch.nn.L1Loss() if kind == "md17" else torch.nn.MSELoss()
    host_pool = [_ft_batch(kind, 10_000 * rank + i).pin_memory() for i in range(args.pool)]
    dev_pool = [b.to(dev) for b in host_pool]
    step_fn = md17_train_step if kind == "md17" else lba_train_step

    def eager(batch):
        return step_fn(targs, batch, model, lin, crit, opt, grad_sync=sync)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(n):
            fn(i)
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(max(args.warmup, 3)):
        eager(dev_pool[i % args.pool])
    barrier()
    step, launch = eager, "eager launches"
    same_shape = len({tuple(b.positions.shape) for b in dev_pool}) == 1
    if not args.no_graph:
        try:
            if kind == "md17":
                step = GraphedMD17Step(targs, dev_pool[0], model, lin, crit, opt, grad_sync=sync)
                step(dev_pool[0])
                launch = ("neighbour search + edge count eager per batch, then forward / force / double backward / Adam replayed "
                          "from one CUDA graph per edge count")
            else:
                step = GraphedFinetuneStep(lambda b: step_fn(targs, b, model, lin, crit, opt, grad_sync=sync, zero_grad=False), dev_pool, opt)
                launch = "whole step captured in one CUDA graph (batches padded to one atom capacity)"
        except Exception as exc:                              # noqa: BLE001 -- report, never hide
            launch = f"eager launches (graph capture refused: {type(exc).__name__}: {str(exc)[:120]})"
            step = eager
    for i in range(2):
        step(dev_pool[i % args.pool])
    if getattr(step, "capture_error", None) is not None:      # GraphedMD17Step keeps running eagerly after a refused capture: say so
        launch = f"eager launches (graph capture refused: {type(step.capture_error).__name__}: {str(step.capture_error)[:160]})"
    elif kind == "md17" and not args.no_graph and hasattr(step, "graphs"):
        launch += f" ({len(step.graphs)} graph(s) captured)"
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda i: step(dev_pool[i % args.pool]), args.steps)
    value = world * f["batch"] * args.steps / (ms / 1e3)
    # per-kernel durations (eager event brackets) + launch count
    names = ("cfconv_fwd", "cfconv_bwd_x", "cfconv_bwd_w", "filter_fwd", "filter_bwd", "linear_fwd", "linear_dgrad", "linear_wgrad", "radius_csr")
    _lib.launch_count(reset=True)
    ops.KERNEL_TIMERS.enable(names)
    n_k = min(args.steps, 5)
    timed(lambda i: eager(dev_pool[i % args.pool]), n_k)
    launches = _lib.launch_count() * args.steps // n_k
    ktimes = ops.KERNEL_TIMERS.collect()
    ops.KERNEL_TIMERS.disable()
    clocks = sampler.stop() if rank == 0 else None
    # end to end: pinned host batch -> H2D -> step -> loss D2H, every step
    sink = torch.empty(1, dtype=torch.float32).pin_memory()

    def e2e(i):
        loss = step(host_pool[i % args.pool].to(dev, non_blocking=True))
        sink.copy_(loss.view(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        assert float(sink[0]) == float(sink[0])
    e2e(0)
    ms_e2e = timed(e2e, args.steps)
    if rank != 0:
        _finish(world)
        return
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
    pk = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = {**json.load(open(pk)), "src": "measured"}
    b0 = dev_pool[0]
    g = ops.radius_csr(b0.positions, b0.batch, f["cutoff"], num_graphs=f["batch"])
    n_atoms, n_edges = b0.positions.size(0), g.num_edges
    shared = kind == "lba" and ops.SHARE_PAIR_FILTERS
    n_rows_w = int(g.ensure_pairs().n_pairs_dev.item()) if shared else n_edges
    cf_bytes = 4 * 128 * n_rows_w + 2 * 4 * 128 * n_atoms + 4 * n_edges * (2 if shared else 1) + 4 * (n_atoms + 1)
    roof = None
    if "cfconv_fwd" in ktimes:
        t = ktimes["cfconv_fwd"]["mean_ms"] / 1e3
        roof = {"kernel": "cfconv_gather_async_kernel<false> (cfconv forward, F = 128)", "bound": "hbm", "achieved": cf_bytes / t / 1e9,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": cf_bytes / t / 1e9 / peaks["hbm_gbs"], "traffic": None,
                "peak_source": peaks["src"], "algorithmic_bytes_per_launch": cf_bytes, "mean_ms": 1e3 * t,
                "note": "eager event brackets (include the host enqueue gap); the working set of this workload "
                        + ("fits the 126 MB L2, so this is not a DRAM-bound launch" if kind == "md17" else "exceeds the L2")}
    others = {k: {"mean_ms": v["mean_ms"], "launches_per_step": v["n"] // n_k} for k, v in ktimes.items() if k != "cfconv_fwd"}
    cpu = None
    if not args.no_cpu_baseline:
        cb = ft_cpu_reference(kind, 2, 1, budget_s=45.0)
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    h2d = sum(t.numel() * t.element_size() for t in (host_pool[0].x, host_pool[0].positions, host_pool[0].batch)) + \
        sum(v.numel() * v.element_size() for v in host_pool[0].extras.values() if torch.is_tensor(v))
    _emit(json.dumps({**base, "value": value, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                      "config": {**cfg, "global_batch": world * f["batch"], "atoms_per_launch": n_atoms, "edges_per_launch": n_edges,
                                 "filter_rows_per_launch": n_rows_w, "launch": launch, "parallelism": f"dp{world}",
                                 "l2": f"{args.pool} distinct batches cycled"},
                      "clocks": clocks, "e2e": {"value": world * f["batch"] * args.steps / (ms_e2e / 1e3), "unit": f["unit"],
                                                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
                      "gpu_launches": launches, "roofline": roof, "roofline_other_kernels": others, "cpu_baseline": cpu}))
    _finish(world)


def _finish(world):
    """Multi-rank exit: the captured CUDA graph keeps references into the NCCL communicator, and tearing the process
    group down under it can block; everything is flushed, so leave without the teardown."""
    if world > 1:
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def _emit(line):
    """The ONE JSON line goes to the real stdout; everything else this process (or NCCL's C code) prints was sent to stderr."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    a = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                       # library chatter on fd 1 (e.g. "NCCL version ...") must not pollute the JSON line
    if a.workload != "ddm":
        run_finetune(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_product(a)
