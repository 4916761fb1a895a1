"""CPU, world_size 2 over gloo: query sharding, flat gradient all-reduce(SUM) and the rank-statistic
reduction reproduce the single-process results (the N>1 host logic of redgnn_b200/dist.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from redgnn_b200 import dist as rgd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class Toy(torch.nn.Module):
    """Stand-in with the model's calling convention (subs, rels) -> scores (n, n_ent) and one
    parameter that gets no gradient (like Ws_attn of layer 0 before the explicit-zero fix)."""

    def __init__(self, n_ent=11):
        super().__init__()
        torch.manual_seed(0)
        self.emb = torch.nn.Embedding(n_ent, 6)
        self.rel = torch.nn.Embedding(5, 6)
        self.out = torch.nn.Linear(6, n_ent)
        self.unused = torch.nn.Parameter(torch.ones(3))

    def forward(self, subs, rels):
        s = torch.as_tensor(np.asarray(subs), dtype=torch.long)
        r = torch.as_tensor(np.asarray(rels), dtype=torch.long)
        return self.out(torch.tanh(self.emb(s) + self.rel(r)))


def _worker(rank, world, port, triples, ranks_all, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = Toy()
        if rank == 1:                                   # de-synchronise, then broadcast from rank 0
            with torch.no_grad():
                model.out.weight.add_(1.0)
        rgd.broadcast_parameters(model)
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        loss = rgd.sharded_train_step(model, opt, triples)
        lo, hi = rgd.shard_bounds(len(ranks_all), rank, world)
        stats = rgd.reduce_rank_stats(ranks_all[lo:hi])
        q.put((rank, float(loss), {k: v.detach().numpy().copy() for k, v in model.state_dict().items()},
               model.unused.grad.numpy().copy(), stats))       # numpy: pickled by value
    finally:
        dist.destroy_process_group()


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            cuts = [rgd.shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    a, b = rgd.shard_batch((np.arange(10), np.arange(10) * 2), 1, 4)
    assert list(a) == [2, 3, 4] and list(b) == [4, 6, 8]


def test_two_rank_train_step_matches_single_process():
    rng = np.random.default_rng(0)
    triples = np.stack([rng.integers(0, 11, 9), rng.integers(0, 5, 9), rng.integers(0, 11, 9)], 1)
    ranks_all = rng.integers(1, 40, 25).astype(np.float64)
    # single process reference: sum-loss over the whole batch, one SGD step
    ref = Toy()
    opt = torch.optim.SGD(ref.parameters(), lr=0.1)
    ref_loss = rgd.sharded_train_step(ref, opt, triples, rank=0, world=1)
    ref_stats = rgd.reduce_rank_stats(ranks_all)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, triples, ranks_all, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert abs(res[0][1] + res[1][1] - float(ref_loss)) < 1e-4 * abs(float(ref_loss))
    for k, v in ref.state_dict().items():
        assert np.allclose(res[0][2][k], v.numpy(), rtol=1e-5, atol=1e-6), k
        assert np.array_equal(res[0][2][k], res[1][2][k]), "ranks diverged on " + k
    assert float(np.abs(res[0][3]).max()) == 0.0          # grad-less parameter: explicit zeros on every rank
    for r in res:
        assert np.allclose(r[4][:3], ref_stats[:3]) and r[4][3] == 25


class FlatToy(Toy):
    """Toy whose gradients live in ONE flat buffer (what RedGNN.flat_grad() provides with
    `grads_in_place`): every p.grad is a view of it, so the exchange is a single in-place all-reduce."""

    def __init__(self):
        super().__init__()
        n = sum(p.numel() for p in self.parameters())
        self._flat = torch.zeros(n)
        off = 0
        for p in self.parameters():
            p.grad = self._flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def flat_grad(self):
        return self._flat


def _flat_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = FlatToy()
        with torch.no_grad():
            for p in model.parameters():
                p.grad.fill_(float(rank + 1))            # "local" gradients, written in place
        n = rgd.allreduce_model_gradients(model)
        q.put((rank, n, [float(p.grad.flatten()[0]) for p in model.parameters()],
               all(p.grad.data_ptr() >= model._flat.data_ptr() for p in model.parameters())))
    finally:
        dist.destroy_process_group()


def test_flat_in_place_gradient_allreduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_param = sum(p.numel() for p in Toy().parameters())
    for rank, n, firsts, still_views in res:
        assert n == n_param and still_views
        assert all(abs(v - 3.0) < 1e-6 for v in firsts)       # 1 + 2 summed, identical on both ranks
