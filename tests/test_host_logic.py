"""CPU: host-side logic -- drop-in module surface (state_dict keys, seeded init bit-identical to the
reference's), the AtomTuple batch contract, and the data-parallel gradient exchange on gloo (world size 2)."""
import itertools
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from _build import full_sd
from _golden import Golden
from geossl_b200.data import pair_count, super_edges_host, synthetic_batch


def _same(sd_a, sd_b):
    assert sorted(sd_a) == sorted(sd_b)
    for k in sd_a:
        assert sd_a[k].dtype == sd_b[k].dtype and sd_a[k].shape == sd_b[k].shape, k
        if k != "atomic_mass":
            assert torch.equal(sd_a[k], sd_b[k]), k


@pytest.mark.parametrize("name", ["schnet_small", "schnet_trunc"])
def test_schnet_seeded_init_matches_reference(name):
    from geossl_b200.Geom3D.models import SchNet
    g = Golden(name)
    c = g.cfg
    torch.manual_seed(c["seed"])
    m = SchNet(hidden_channels=c["hidden"], num_filters=c["filters"], num_interactions=c["layers"],
               num_gaussians=c["gaussians"], cutoff=c["cutoff"], node_class=9, readout=c["readout"])
    _same(m.state_dict(), full_sd(g.sd()))
    assert m.state_dict()["atomic_mass"].dtype == torch.float64 and m.state_dict()["atomic_mass"].numel() == 119
    # quirk kept: mlp[2].bias is NOT zeroed (schnet.py:155-158)
    assert m.interactions[0].mlp[2].bias.abs().max() > 0 and m.interactions[0].mlp[0].bias.abs().max() == 0
    assert m.interactions[0].conv.nn[0].weight is m.interactions[0].mlp[0].weight


@pytest.mark.parametrize("name", ["painn_small", "painn_full"])
def test_painn_seeded_init_matches_reference(name):
    from geossl_b200.Geom3D.models import PaiNN
    g = Golden(name)
    c = g.cfg
    torch.manual_seed(c["seed"])
    m = PaiNN(n_atom_basis=c["feat"], n_interactions=c["layers"], n_rbf=c["rbf"], cutoff=c["cutoff"], max_z=9, n_out=1,
              readout=c["readout"])
    _same(m.state_dict(), g.sd())
    assert m.embedding.weight[0].abs().max() == 0       # padding_idx=0 (painn.py:174)


def test_ncsn_seeded_init_and_sigma_schedule():
    from geossl_b200.NCSN import NCSN_version_03
    g = Golden("ncsn_h128")
    c = g.cfg
    torch.manual_seed(c["seed"])
    m = NCSN_version_03(c["emb"], sigma_begin=10, sigma_end=0.01, num_noise_level=c["levels"], noise_type="symmetry",
                        anneal_power=c["anneal_power"])
    _same(m.state_dict(), g.sd())
    assert not m.sigmas.requires_grad and "sigmas" in dict(m.named_parameters())   # frozen nn.Parameter (NCSN.py:179)


def test_super_edges_match_itertools():
    counts = [1, 2, 5, 0, 7, 3]
    for option, gen in (("combination", itertools.combinations), ("permutation", itertools.permutations)):
        sei = super_edges_host(counts, option)
        ref, off = [], 0
        for n in counts:
            if n >= 2:
                ref.append(np.array(list(gen(np.arange(n), 2))).T + off)
            off += n
        assert torch.equal(sei, torch.from_numpy(np.concatenate(ref, axis=1)))
        assert sei.shape[1] == int(pair_count(counts, option).sum())


def test_pair_subsampling_consumes_numpy_stream_like_the_extractor():
    """--distance_sample_ratio < 1 (dataloaders_AtomTuple.py:15-37): literal replay of AtomTupleExtractor.__call__ per
    molecule + the collate offset (:60-66) with the same np.random seed must give the same columns in the same order."""
    from geossl_b200.data import sample_pairs_host, sampled_pair_count
    counts, ratio = [1, 2, 5, 0, 7, 3, 12], 0.4
    for option, gen in (("combination", itertools.combinations), ("permutation", itertools.permutations)):
        np.random.seed(7)
        ref, off = [], 0
        for n in counts:
            if n >= 2:
                sei = np.array(list(gen(np.arange(n), 2))).T
                M = sei.shape[1]
                sampled = np.random.choice(M, int(M * ratio), replace=False)
                ref.append(sei[:, sampled] + off)
            off += n
        ref = torch.from_numpy(np.concatenate(ref, axis=1))
        np.random.seed(7)
        got = super_edges_host(counts, option, ratio=ratio)
        assert torch.equal(got, ref)
        assert got.shape[1] == int(sampled_pair_count(counts, option, ratio).sum())
        np.random.seed(7)
        sel = sample_pairs_host(counts, option, ratio)
        assert torch.equal(super_edges_host(counts, option, ratio=ratio, selection=sel), ref)


def test_synthetic_batch_contract():
    b = synthetic_batch(8, 10, 20, seed=3)
    n = b.positions.shape[0]
    assert b.x.shape == (n, 2) and b.x.dtype == torch.int64 and b.positions.dtype == torch.float32
    assert b.batch.dtype == torch.int64 and bool((b.batch[1:] >= b.batch[:-1]).all())
    assert b.num_graphs == 8 and int(b.x[:, 0].max()) <= 8
    assert bool((b.batch[b.super_edge_index[0]] == b.batch[b.super_edge_index[1]]).all())
    assert int(b.graph_ptr[-1]) == n


def test_batches_carry_the_largest_molecule_and_padding_keeps_or_replaces_it():
    """``extras['max_graph_atoms']`` is the host-side bound that selects the pair-centric cfconv kernel without a device sync
    (ops.CFCONV_PAIRS): set at collate time, kept by ``to`` / ``pin_memory`` / ``pad_batch``, replaced by the capacity of a
    padded stream when one is given, and read back by ``pretrain._max_graph_atoms``."""
    from geossl_b200.data import pad_batch
    from geossl_b200.pretrain import _max_graph_atoms
    b = synthetic_batch(8, 10, 20, seed=3)
    largest = int(torch.bincount(b.batch).max())
    assert b.extras["max_graph_atoms"] == largest == _max_graph_atoms(b)
    assert b.to("cpu").extras["max_graph_atoms"] == largest
    n, p = b.positions.shape[0], b.super_edge_index.shape[1]
    padded = pad_batch(b, n + 40, p + 100)
    assert padded.extras["max_graph_atoms"] == largest and padded.num_graphs == 9          # + the edge-less padding graph
    assert int(padded.graph_ptr[-1]) == n + 40 and int(padded.extras["n_pairs_live"]) == p
    assert pad_batch(b, n + 40, p + 100, max_graph_atoms_cap=32).extras["max_graph_atoms"] == 32
    ref_style = type(b)(b.x, b.positions, b.batch, b.super_edge_index)                     # a reference batch has no bound
    assert _max_graph_atoms(ref_style) is None


def _dp_worker(rank, world, port, tmp):
    import torch.distributed as dist
    from geossl_b200.pretrain import FlatGradAllReduce, broadcast_parameters
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # different init per rank ...
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    broadcast_parameters([model])                       # ... made identical by the broadcast
    sync = FlatGradAllReduce(model.parameters())
    torch.manual_seed(7)
    x = torch.randn(8, 6)
    shard = x[rank * 4:(rank + 1) * 4]                  # equal shards => mean of means == global mean
    model(shard).pow(2).mean().backward()
    sync()
    torch.save({"grads": [p.grad.clone() for p in model.parameters()],
                "params": [p.detach().clone() for p in model.parameters()]}, os.path.join(tmp, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_flat_grad_allreduce_gloo_world2(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    for a, b in zip(r0["params"], r1["params"]):
        assert torch.equal(a, b)
    for a, b in zip(r0["grads"], r1["grads"]):
        assert torch.equal(a, b)
    # single-process reference on the full batch
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    with torch.no_grad():
        for p, v in zip(model.parameters(), r0["params"]):
            p.copy_(v)
    torch.manual_seed(7)
    x = torch.randn(8, 6)
    model(x).pow(2).mean().backward()
    for p, gsync in zip(model.parameters(), r0["grads"]):
        assert torch.allclose(p.grad, gsync, rtol=1e-5, atol=1e-7)


def test_embedding_lookup_backward_matches_torch_embedding():
    """ops.EmbeddingLookup (one-hot GEMM backward) vs nn.Embedding's own backward, with and without padding_idx
    (pure torch ops, so the check runs on the CPU; the CUDA path uses the same function)."""
    from geossl_b200 import ops
    g = torch.Generator().manual_seed(0)
    for rows, pad in ((9, None), (100, 0)):
        emb = torch.nn.Embedding(rows, 16, padding_idx=pad)
        z = torch.randint(0, rows, (257,), generator=g)
        w_out = torch.randn(257, 16, generator=g)
        ref = emb(z)
        (ref * w_out).sum().backward()
        gref = emb.weight.grad.clone()
        emb.weight.grad = None
        out = ops.EmbeddingLookup.apply(emb.weight, z, pad)
        assert torch.equal(out, ref)
        (out * w_out).sum().backward()
        assert torch.allclose(emb.weight.grad, gref, rtol=1e-5, atol=1e-6)
        if pad is not None:
            assert emb.weight.grad[pad].abs().max() == 0
    # the lookup is a torch op either way (not one of the CUDA kernels); off the GPU the dispatcher leaves nn.Embedding alone
    # -- the model itself still refuses CPU tensors at the first kernel (tests/test_abi.py::test_no_cpu_fallback)
    assert torch.equal(ops.embedding(emb, z), emb(z))
