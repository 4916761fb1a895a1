"""GPU: behaviours around the captured-graph paths that a long training / serving run depends on:
recovery from a diverged step (the reference loop survives those, base_model.py:65-69), the in-place
device re-split of shuffle_train (transductive/load_data.py:152-164) under live CUDA graphs, the
lifetime of the expansion scratch under many batch sizes, and input validation."""
import numpy as np
import pytest
import torch

from oracle import redgnn_oracle as O
from helpers import assert_close, assert_expansion_equal
from redgnn_b200.synth import Options

pytestmark = pytest.mark.gpu


def loss_backward(model, tri):
    model.zero_grad(set_to_none=True)
    out = model(tri[:, 0], tri[:, 1])
    pos = out[torch.arange(len(out)).cuda(), torch.as_tensor(tri[:, 2]).cuda()]
    mx = out.max(1, keepdim=True)[0]
    loss = torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(out - mx), 1)))
    loss.backward()
    return loss.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def test_graph_training_recovers_after_a_diverged_step(tiny_dir):
    """A step with NaN parameters poisons the runner's persistent per-node buffers; the next step has
    FEWER nodes (stale rows past its count), and must still give finite, correct gradients."""
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    L = TransductiveLoader(tiny_dir)
    model = RED_GNN_trans(Options(hidden_dim=48, attn_dim=5, n_layer=2, dropout=0.0, act="relu", n_rel=L.n_rel), L).cuda()
    model.train()
    train = np.asarray(L.train_data)
    deg = np.bincount(L.KG[:, 0], minlength=L.n_ent)
    order = np.argsort(deg[train[:, 0]], kind="stable")
    small, big = train[order[:10]], train[order[-10:]]
    model.graph_train = False
    want_loss, want = loss_backward(model, small)            # eager autograd path: no persistent state
    model.graph_train = True
    runner_counts = []
    loss_backward(model, big)
    runner = next(iter(model._train_graph_cache.values()))
    runner_counts.append(runner.node_counts())
    saved = {k: p.detach().clone() for k, p in model.named_parameters()}
    with torch.no_grad():
        model.gate.weight_ih_l0.fill_(float("nan"))          # the diverged step: NaN gates -> NaN hidden / saved rows
    bad_loss, bad = loss_backward(model, big)
    assert not torch.isfinite(bad_loss)
    with torch.no_grad():                                    # base_model.py:65-69 re-randomises NaN parameters
        for k, p in model.named_parameters():
            p.copy_(saved[k])
    got_loss, got = loss_backward(model, small)
    runner_counts.append(runner.node_counts())
    assert any(a > b for a, b in zip(*runner_counts)), "the second step must leave stale rows behind"
    assert len(model._train_graph_cache) == 1
    assert torch.isfinite(got_loss) and all(torch.isfinite(g).all() for g in got.values())
    assert_close(got_loss, want_loss, 1e-5, "loss after the diverged step")
    floor = 1e-7 * max(float(g.abs().max()) for g in want.values())
    for k in want:
        err, scale = float((got[k] - want[k]).abs().max()), float(want[k].abs().max())
        assert err <= 2e-4 * scale or err <= floor, "grad %s after the diverged step: err %.3e scale %.3e" % (k, err, scale)


def test_shuffle_train_resplits_on_the_device_in_place_and_graphs_survive(tiny_dir):
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    L = TransductiveLoader(tiny_dir)
    model = RED_GNN_trans(Options(hidden_dim=48, attn_dim=5, n_layer=3, dropout=0.0, act="relu", n_rel=L.n_rel), L).cuda()
    model.train()
    g0 = L.graph_for("train")
    addr = [t.data_ptr() for t in (g0.head, g0.rel, g0.tail, g0.in_ptr, g0.in_adj, g0.out_ptr, g0.out_adj)]
    loss_backward(model, L.get_batch(np.arange(12)))         # captures the training graphs on the train KG
    pool = np.concatenate([np.array(L.fact_triple), np.array(L.train_triple)], 0)
    for seed in (7, 8):
        np.random.seed(seed)
        L.shuffle_train()
        np.random.seed(seed)
        perm = np.random.permutation(len(pool))               # transductive/load_data.py:157
        n_keep = len(pool) * 3 // 4
        facts = pool[perm][:n_keep]
        want_graph = O.Graph(O.add_inverse_block(facts.tolist(), L.n_rel), L.n_ent, L.n_rel)
        g1 = L.graph_for("train")
        assert g1 is g0 and addr == [t.data_ptr() for t in (g1.head, g1.rel, g1.tail, g1.in_ptr, g1.in_adj,
                                                             g1.out_ptr, g1.out_adj)]
        assert np.array_equal(g1.kg_numpy(), want_graph.KG.astype(np.int64))
        assert np.array_equal(L.KG, want_graph.KG.astype(np.int64))                  # lazy host copy
        assert np.array_equal(np.asarray(L.fact_data), np.array(O.add_inverse_block(facts.tolist(), L.n_rel)))
        assert np.array_equal(L.train_data, np.array(O.add_inverse_block(pool[perm][n_keep:].tolist(), L.n_rel)))
        assert L.n_train == len(L.train_data) and L.n_fact == len(want_graph.KG)
        # expansion on the rebuilt CSR views: bit-exact against the oracle on the re-split KG
        tri = L.get_batch(np.arange(9))
        nodes = np.stack([np.arange(9), tri[:, 0]], 1)
        for _ in range(2):
            want = O.get_neighbors(want_graph, nodes)
            assert_expansion_equal(L.get_neighbors(nodes, "train"), want, "after shuffle_train")
            nodes = want[0].numpy()
        # the captured training step keeps working on the rebuilt KG (same addresses): equals the eager path
        res = {}
        for graph_train in (True, False):
            model.graph_train = graph_train
            res[graph_train] = loss_backward(model, tri)
        assert len(model._train_graph_cache) <= 2 and g1.epoch <= 1
        assert_close(res[True][0], res[False][0], 1e-5, "loss after in-place re-split")
        floor = 1e-7 * max(float(g.abs().max()) for g in res[False][1].values())
        for k in res[True][1]:
            err = float((res[True][1][k] - res[False][1][k]).abs().max())
            assert err <= 2e-4 * float(res[False][1][k].abs().max()) or err <= floor, k
    want_sc = O.model_forward({k: v.detach().cpu() for k, v in model.state_dict().items()}, want_graph,
                              tri[:, 0], tri[:, 1], 3, "relu")
    model.eval()
    with torch.no_grad():
        assert_close(model(tri[:, 0], tri[:, 1], mode="train"), want_sc, 1e-4, "scores on the re-split KG")


def test_expansion_scratch_outlives_many_batch_sizes(tiny_dir):
    """More distinct batch sizes than any cache holds: every captured inference graph must keep
    replaying correctly (its expansion scratch is referenced by the cache entry, never freed)."""
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    L = TransductiveLoader(tiny_dir)
    model = RED_GNN_trans(Options(n_rel=L.n_rel), L).cuda().eval()
    model.MAX_CACHED_GRAPHS = 16
    q = np.array(L.test_q)
    first = {}
    with torch.no_grad():
        for n in list(range(1, 13)) + [40, 3, 1, 12, 7]:
            got = model(q[:n, 0], q[:n, 1], mode="test")
            if n in first:
                assert torch.equal(got, first[n]), "replay of batch size %d changed after other sizes ran" % n
            else:
                first[n] = got.clone()
        torch.cuda.empty_cache()
        junk = [torch.full((1 << 20,), float("nan"), device="cuda") for _ in range(8)]   # reuse freed blocks, if any
        for n in (1, 5, 12, 40):
            assert torch.equal(model(q[:n, 0], q[:n, 1], mode="test"), first[n])
        del junk


def test_invalid_ids_are_refused(tiny_dir):
    from redgnn_b200 import DeviceGraph, TransductiveLoader, RED_GNN_trans, _lib
    with pytest.raises(_lib.RgError):
        DeviceGraph(np.array([[0, 0, 5]]), 5, 2, "cuda")           # entity id == n_ent
    with pytest.raises(_lib.RgError):
        DeviceGraph(np.array([[0, 5, 1]]), 5, 2, "cuda")           # relation id > 2 * n_rel
    L = TransductiveLoader(tiny_dir)
    model = RED_GNN_trans(Options(n_rel=L.n_rel), L).cuda().eval()
    bad = torch.tensor([1, L.n_ent], device="cuda")
    rels = torch.tensor([0, 1], device="cuda")
    with pytest.raises(_lib.RgError):
        model(bad, rels, mode="test")                              # tensor inputs: device-side range check
    model.check_tensor_inputs = False                              # sync-free serving: surfaces lazily
    with torch.no_grad():
        model(bad, rels, mode="test")
    with pytest.raises(_lib.RgError):
        model.last_stats
    model.train()                                                  # graph-captured training step: no host sync at all
    model(bad, rels).sum().backward()
    with pytest.raises(_lib.RgError):
        model.last_stats
