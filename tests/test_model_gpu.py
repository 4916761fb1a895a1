"""GPU parity of the whole path behind RED_GNN_trans / RED_GNN_induc.forward: per-entity scores
within 1e-4 relative (fp32), unvisited entities exactly 0, parameter gradients of the reference
training loss within 1e-4, filtered ranks / MRR / Hits identical to 3 decimals."""
import numpy as np
import pytest
import torch

from oracle import redgnn_oracle as O
from helpers import golden, golden_state_dict, assert_close, assert_grad_close, grad_floor, to64
from redgnn_b200.synth import Options

pytestmark = pytest.mark.gpu


def oracle_loss_grads(sd, graph, tri, n_layer, act):
    """Gradients of the reference training loss (base_model.py:58-60) in fp32 and fp64."""
    out = []
    for conv in ((lambda t: t.clone()), (lambda t: t.double())):
        sd_g = {k: conv(v).requires_grad_(True) for k, v in sd.items()}
        O.train_loss(O.model_forward(sd_g, graph, tri[:, 0], tri[:, 1], n_layer, act), tri[:, 2]).backward()
        out.append({k: v.grad for k, v in sd_g.items()})
    return out


def cuda_loss_backward(model, tri):
    model.inference_in_eval = False            # differentiate through an eval-mode (dropout-free) forward
    out = model(tri[:, 0], tri[:, 1])
    pos = out[torch.arange(len(out)).cuda(), torch.as_tensor(tri[:, 2]).cuda()]
    mx = out.max(1, keepdim=True)[0]
    model.zero_grad()
    loss = torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(out - mx), 1)))
    loss.backward()
    return loss


def write_family_dir(tmp_path, fx):
    """Rebuild a dataset directory in the reference's text layout from the family fixture."""
    d = tmp_path / "family"
    d.mkdir()
    n_ent, n_rel = int(fx["n_ent"]), int(fx["n_rel"])
    (d / "entities.txt").write_text("".join("e%d\n" % i for i in range(n_ent)))
    (d / "relations.txt").write_text("".join("r%d\n" % i for i in range(n_rel)))
    dump = lambda tri: "".join("e%d r%d e%d\n" % tuple(t) for t in tri.tolist())
    (d / "facts.txt").write_text(dump(fx["fact_triple"]))
    (d / "train.txt").write_text(dump(fx["train_triple"]))
    (d / "valid.txt").write_text("e0 r0 e1\n")
    (d / "test.txt").write_text("e0 r0 e1\n")
    return str(d)


def test_golden_family_scores_ranks_grads(tmp_path):
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    fx = golden("family")
    L = TransductiveLoader(write_family_dir(tmp_path, fx))
    opts = Options(hidden_dim=48, attn_dim=5, n_layer=3, dropout=0.29, act="relu", n_rel=L.n_rel)
    model = RED_GNN_trans(opts, L).cuda()
    model.load_state_dict(golden_state_dict(fx))          # reference parameter names load unchanged
    model.eval()
    scores = model(fx["eval_subs"], fx["eval_rels"], mode="test")
    want = torch.from_numpy(fx["eval_scores"])
    assert_close(scores, want, 1e-4, "family scores")
    assert torch.equal(scores.cpu() == 0, want == 0), "visited / unvisited entity sets differ"
    ranks = O.cal_ranks(scores.detach().cpu().numpy(), fx["eval_objs"].astype(np.float64),
                        fx["eval_filters"].astype(np.float64))
    assert np.array_equal(np.array(ranks), fx["eval_ranks"])
    for a, b in zip(O.cal_performance(ranks), O.cal_performance(fx["eval_ranks"])):
        assert round(a, 3) == round(b, 3)
    tri = fx["train_triples"]
    model.zero_grad()
    model.inference_in_eval = False
    out = model(tri[:, 0], tri[:, 1])                     # mode='train' graph, eval() => no dropout
    pos = out[torch.arange(len(out)).cuda(), torch.as_tensor(tri[:, 2]).cuda()]
    mx = out.max(1, keepdim=True)[0]
    loss = torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(out - mx), 1)))
    loss.backward()
    assert abs(loss.item() - float(fx["train_loss"])) <= 1e-4 * abs(float(fx["train_loss"]))
    # yardstick for cancellation-prone sums (w_alpha.bias ...): the oracle evaluated in fp64
    from helpers import family_graphs
    _, g64 = oracle_loss_grads(golden_state_dict(fx), family_graphs(fx)[0], tri, 3, "relu")
    for k, p in model.named_parameters():
        assert_grad_close(p.grad, torch.from_numpy(fx["train_grad." + k]), g64[k], 1e-4, "family grad " + k,
                          floor=grad_floor(g64))


@pytest.mark.parametrize("act,d,a,n_layer", [("relu", 48, 5, 3), ("tanh", 32, 3, 4), ("idd", 64, 5, 2)])
def test_transductive_model_vs_oracle(tiny_dir, act, d, a, n_layer):
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    L, D = TransductiveLoader(tiny_dir), O.TransductiveData(tiny_dir)
    sd = O.init_state_dict(n_layer, d, a, D.n_rel, seed=7)
    opts = Options(hidden_dim=d, attn_dim=a, n_layer=n_layer, dropout=0.1, act=act, n_rel=L.n_rel)
    model = RED_GNN_trans(opts, L).cuda()
    model.load_state_dict(sd)
    model.eval()
    subs, rels, objs = L.get_batch(np.arange(40), data="test")
    for mode in ("train", "test"):
        got = model(subs, rels, mode=mode)
        want = O.model_forward(sd, D.graph_for(mode), subs, rels, n_layer, act)
        assert_close(got, want, 1e-4, "scores " + mode)
        assert torch.equal(got.cpu() == 0, want == 0)
    # gradients of the training loss (base_model.py:58-60)
    tri = L.get_batch(np.arange(10))
    g32, g64 = oracle_loss_grads(sd, D.graph, tri, n_layer, act)
    cuda_loss_backward(model, tri)
    for k, p in model.named_parameters():
        assert_grad_close(p.grad, g32[k], g64[k], 1e-4, "grad " + k, floor=grad_floor(g64))
    # a second identical forward is bit-identical (no atomics in the forward path)
    assert torch.equal(model(subs, rels, mode="test"), model(subs, rels, mode="test"))


def test_hub_model_vs_oracle(hub_dir):
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    L, D = TransductiveLoader(hub_dir), O.TransductiveData(hub_dir)
    sd = O.init_state_dict(3, 48, 5, D.n_rel, seed=9)
    model = RED_GNN_trans(Options(n_rel=L.n_rel), L).cuda()
    model.load_state_dict(sd)
    model.eval()
    subs, rels, _ = L.get_batch(np.arange(8), data="valid")
    got = model(subs, rels, mode="valid")
    want = O.model_forward(sd, D.test_graph, subs, rels, 3, "relu")
    assert_close(got, want, 1e-4, "hub scores")
    tri = L.get_batch(np.arange(4))
    g32, g64 = oracle_loss_grads(sd, D.graph, tri, 3, "relu")
    cuda_loss_backward(model, tri)
    for k, p in model.named_parameters():
        assert_grad_close(p.grad, g32[k], g64[k], 1e-4, "hub grad " + k, floor=grad_floor(g64))


def test_many_relations_short_segment_path_vs_oracle(manyrel_dir):
    """601 relation-table rows and 32 x 3000 >= 75,776 (upper-bound) segments: forward / backward run the
    non-persistent kernels with 8 segments per warp (rg_edge.cu: fwd_block8 / bwd_group)."""
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    L, D = TransductiveLoader(manyrel_dir), O.TransductiveData(manyrel_dir)
    sd = O.init_state_dict(3, 48, 5, D.n_rel, seed=13)
    model = RED_GNN_trans(Options(n_rel=L.n_rel), L).cuda()
    model.load_state_dict(sd)
    model.eval()
    subs, rels, _ = L.get_batch(np.arange(32), data="valid")
    got = model(subs, rels, mode="valid")
    want = O.model_forward(sd, D.test_graph, subs, rels, 3, "relu")
    assert_close(got, want, 1e-4, "many-relation scores")
    assert torch.equal(got, model(subs, rels, mode="valid"))
    model.train()
    tri = L.get_batch(np.arange(32))
    g32, g64 = oracle_loss_grads(sd, D.graph, tri, 3, "relu")
    for graph_train in (True, False):           # captured training step, then the eager autograd path
        model.graph_train = graph_train
        model.zero_grad(set_to_none=True)
        cuda_loss_backward(model, tri)
        for k, p in model.named_parameters():
            assert_grad_close(p.grad, g32[k], g64[k], 1e-4, "many-relation grad %s (graph=%s)" % (k, graph_train),
                              floor=grad_floor(g64))


def test_inductive_model_vs_oracle_and_golden(induc_dir):
    from redgnn_b200 import InductiveLoader, RED_GNN_induc
    L, D = InductiveLoader(induc_dir), O.InductiveData(induc_dir)
    sd = O.init_state_dict(3, 48, 5, D.n_rel, seed=5)
    model = RED_GNN_induc(Options(n_rel=L.n_rel), L).cuda()
    model.load_state_dict(sd)
    model.eval()
    subs, rels, _ = L.get_batch(np.arange(15), data="valid")
    got = model(subs, rels)
    assert got.shape == (15, L.n_ent)
    assert_close(got, O.model_forward(sd, D.tra_graph, subs, rels, 3, "relu"), 1e-4, "transductive mode")
    subs, rels, _ = L.get_batch(np.arange(15), data="test")
    got = model(subs, rels, "inductive")
    assert got.shape == (15, L.n_ent_ind)
    want = O.model_forward(sd, D.ind_graph, subs, rels, 3, "relu", n_ent_out=D.n_ent_ind)
    assert_close(got, want, 1e-4, "inductive mode")


def test_golden_fb237_v2_scores(tmp_path):
    """Scores the live reference produced for fb237_v2 (inductive mode) vs the CUDA path, with the
    graph rebuilt from the fixture through DeviceGraph directly."""
    from redgnn_b200 import RED_GNN_induc, DeviceGraph

    fx = golden("fb237_v2")

    class FixtureLoader(object):
        n_ent, n_ent_ind, n_rel = int(fx["n_ent"]), int(fx["n_ent_ind"]), int(fx["n_rel"])

        def __init__(self):
            self.g = {"transductive": DeviceGraph(fx["tra_triples"].astype(np.int64), self.n_ent, self.n_rel, "cuda"),
                      "inductive": DeviceGraph(fx["ind_triples"].astype(np.int64), self.n_ent_ind, self.n_rel, "cuda")}

        def graph_for(self, mode, device=None):
            return self.g["transductive" if mode == "transductive" else "inductive"]

        def n_ent_for(self, mode):
            return self.n_ent if mode == "transductive" else self.n_ent_ind

    L = FixtureLoader()
    model = RED_GNN_induc(Options(n_rel=L.n_rel, dropout=0.3), L).cuda()
    model.load_state_dict(golden_state_dict(fx))
    model.eval()
    scores = model(fx["eval_subs"], fx["eval_rels"], "inductive")
    want = torch.from_numpy(fx["eval_scores"])
    assert_close(scores, want, 1e-4, "fb237_v2 scores")
    assert torch.equal(scores.cpu() == 0, want == 0)
    ranks = O.cal_ranks(scores.detach().cpu().numpy(), fx["eval_objs"].astype(np.float64),
                        fx["eval_filters"].astype(np.float64))
    assert np.array_equal(np.array(ranks), fx["eval_ranks"])
    tri = fx["train_triples"]
    model.inference_in_eval = False
    out = model(tri[:, 0], tri[:, 1])
    pos = out[torch.arange(len(out)).cuda(), torch.as_tensor(tri[:, 2]).cuda()]
    mx = out.max(1, keepdim=True)[0]
    model.zero_grad()
    torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(out - mx), 1))).backward()
    from test_oracle_golden import fb237_graphs
    _, g64 = oracle_loss_grads(golden_state_dict(fx), fb237_graphs(fx)[0], tri, 3, "relu")
    for k, p in model.named_parameters():
        assert_grad_close(p.grad, torch.from_numpy(fx["train_grad." + k]), g64[k], 1e-4, "fb237_v2 grad " + k,
                          floor=grad_floor(g64))


@pytest.mark.parametrize("act,d,a,n_layer", [("relu", 48, 5, 3), ("tanh", 32, 3, 4), ("idd", 64, 5, 2),
                                             ("relu", 16, 8, 3)])
def test_fused_inference_path_vs_oracle(tiny_dir, act, d, a, n_layer):
    """torch.no_grad(): the per-node work runs in the fused rg_node_update kernel."""
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans, _lib
    L, D = TransductiveLoader(tiny_dir), O.TransductiveData(tiny_dir)
    sd = O.init_state_dict(n_layer, d, a, D.n_rel, seed=11)
    model = RED_GNN_trans(Options(hidden_dim=d, attn_dim=a, n_layer=n_layer, dropout=0.2, act=act, n_rel=L.n_rel),
                          L).cuda()
    model.load_state_dict(sd)
    model.eval()
    subs, rels, _ = L.get_batch(np.arange(70), data="test")
    _lib.Stats.timing = []
    with torch.no_grad():
        got = model(subs, rels, mode="test")
    names = [t[0] for t in _lib.Stats.timing]
    _lib.Stats.timing = None
    assert names.count("node_update") == n_layer and names.count("edge_fwd") == n_layer
    want = O.model_forward(sd, D.test_graph, subs, rels, n_layer, act)
    assert_close(got, want, 1e-4, "fused scores")
    assert torch.equal(got.cpu() == 0, want == 0)
    model.inference_in_eval = False
    composed = model(subs, rels, mode="test")
    model.inference_in_eval = True
    assert composed.grad_fn is not None and got.grad_fn is None
    assert_close(got, composed, 1e-5, "fused vs torch-composed path")
    assert model(subs, rels, mode="test").grad_fn is None      # eval(): inference kernels by default
    with torch.no_grad():
        assert torch.equal(got, model(subs, rels, mode="test")), "inference must be bit-reproducible"


def test_fused_inference_golden_family_and_hub(tmp_path, hub_dir):
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    fx = golden("family")
    L = TransductiveLoader(write_family_dir(tmp_path, fx))
    model = RED_GNN_trans(Options(n_rel=L.n_rel, dropout=0.29), L).cuda()
    model.load_state_dict(golden_state_dict(fx))
    model.eval()
    with torch.no_grad():
        scores = model(fx["eval_subs"], fx["eval_rels"], mode="test")
    want = torch.from_numpy(fx["eval_scores"])
    assert_close(scores, want, 1e-4, "family fused scores")
    assert torch.equal(scores.cpu() == 0, want == 0)
    ranks = O.cal_ranks(scores.cpu().numpy(), fx["eval_objs"].astype(np.float64), fx["eval_filters"].astype(np.float64))
    assert np.array_equal(np.array(ranks), fx["eval_ranks"])
    L2, D2 = TransductiveLoader(hub_dir), O.TransductiveData(hub_dir)
    sd = O.init_state_dict(3, 48, 5, D2.n_rel, seed=9)
    m2 = RED_GNN_trans(Options(n_rel=L2.n_rel), L2).cuda()
    m2.load_state_dict(sd)
    m2.eval()
    subs, rels, _ = L2.get_batch(np.arange(8), data="valid")
    with torch.no_grad():
        got = m2(subs, rels, mode="valid")
    assert_close(got, O.model_forward(sd, D2.test_graph, subs, rels, 3, "relu"), 1e-4, "hub fused scores")


def test_cuda_graph_replay_matches_eager_and_tracks_parameters(tiny_dir):
    """The captured inference forward: new inputs, in-place parameter updates and a new KG
    (shuffle_train) all give the same scores as the eager sync-free path."""
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans, _lib
    L = TransductiveLoader(tiny_dir)
    model = RED_GNN_trans(Options(n_rel=L.n_rel), L).cuda()
    model.eval()
    batches = [L.get_batch(np.arange(k * 24, (k + 1) * 24), data="test")[:2] for k in range(3)]

    def eager(subs, rels, mode="test"):
        model.use_cuda_graph = False
        try:
            with torch.no_grad():
                return model(subs, rels, mode=mode)
        finally:
            model.use_cuda_graph = True

    with torch.no_grad():
        for subs, rels in batches:                       # first call captures, the others replay
            before = _lib.Stats.launches
            got = model(subs, rels, mode="test")
            assert _lib.Stats.launches > before
            assert torch.equal(got, eager(subs, rels))
        assert len(model._graph_cache) == 1
        stats = model.last_stats
        assert len(stats["edges"]) == 3 and stats["edges"][0] > 0
        for p in model.parameters():                     # optimiser-style in-place update
            p.add_(0.01 * torch.randn_like(p))
        subs, rels = batches[0]
        assert torch.equal(model(subs, rels, mode="test"), eager(subs, rels))
        assert len(model._graph_cache) == 1
        np.random.seed(3)
        L.shuffle_train()                                # new train graph object -> new capture
        got = model(subs, rels, mode="train")
        assert len(model._graph_cache) == 2
        assert torch.equal(got, eager(subs, rels, "train"))


def test_fb15k237_scale_scores_vs_oracle_and_batch_invariance():
    """BASELINE configs[2] at full size (14,541 entities, 272,115 triples, n_layer=4): two queries
    against the oracle, and batch-composition invariance (a query's scores do not depend on which
    other queries share its batch) on a 48-query batch -- a size-independent property of the path."""
    from redgnn_b200 import RED_GNN_trans
    from redgnn_b200.synth import ArrayLoader
    L = ArrayLoader("fb15k237", seed=0)
    n_layer = 4
    sd = O.init_state_dict(n_layer, 48, 5, L.n_rel, seed=1234)
    model = RED_GNN_trans(Options(n_layer=n_layer, n_rel=L.n_rel), L).cuda()
    model.load_state_dict(sd)
    model.eval()
    q = np.array(L.test_q)
    og = O.Graph(L._test_graph.triples, L.n_ent, L.n_rel)
    want = O.model_forward(sd, og, q[:2, 0], q[:2, 1], n_layer, "relu")
    with torch.no_grad():
        got2 = model(q[:2, 0], q[:2, 1], mode="test")
        got48 = model(q[:48, 0], q[:48, 1], mode="test")
        rev = model(q[:48, 0][::-1].copy(), q[:48, 1][::-1].copy(), mode="test")
    assert_close(got2, want, 1e-4, "fb15k237-scale scores")
    assert torch.equal(got2.cpu() == 0, want == 0)
    # not bit-equal: the tiny per-query projection Wqr(rela[q_rel]) is a cuBLAS call whose reduction
    # order depends on the batch size; everything per edge / per node is batch independent
    assert_close(got48[:2], got2, 1e-5, "batch invariance")
    assert_close(rev.flip(0), got48, 1e-5, "batch order invariance")
    st = model.last_stats
    assert len(st["edges"]) == n_layer and st["edges"][-1] > 20_000_000


@pytest.mark.parametrize("act,d", [("relu", 48), ("tanh", 32), ("idd", 16)])
def test_fused_train_node_update_matches_composed_path(tiny_dir, act, d):
    """Training path: node update in the tcgen05 kernel + explicit backward (NodeUpdateTrain) against
    the torch-composed path (autograd through W_h / act / index_copy / GRU) -- values and all grads."""
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    L = TransductiveLoader(tiny_dir)
    model = RED_GNN_trans(Options(hidden_dim=d, attn_dim=5, n_layer=3, dropout=0.0, act=act, n_rel=L.n_rel), L).cuda()
    model.train()
    model.graph_train = False                      # eager autograd path (the graph path has its own test)
    tri = L.get_batch(np.arange(12))
    res = {}
    for fused in (True, False):
        model.fused_train_node_update = fused
        loss = cuda_loss_backward(model, tri)
        res[fused] = (loss.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()})
    assert_close(res[True][0], res[False][0], 1e-5, "loss")
    for k in res[True][1]:
        assert_close(res[True][1][k], res[False][1][k], 2e-4, "grad " + k)
    # dropout active: runs, finite, and differs from the dropout-free result
    model.fused_train_node_update = True
    model.dropout.p = 0.3
    loss_d = cuda_loss_backward(model, tri)
    assert torch.isfinite(loss_d) and all(torch.isfinite(p.grad).all() for p in model.parameters())
    assert abs(float(loss_d.detach()) - float(res[True][0])) > 0


@pytest.mark.parametrize("act,d,fixture", [("relu", 48, "tiny_dir"), ("tanh", 32, "tiny_dir"), ("relu", 48, "hub_dir")])
def test_graph_captured_training_step_matches_eager(request, act, d, fixture):
    """train_graph.py: forward + hand-written backward replayed as CUDA graphs vs the eager autograd
    path -- loss and every parameter gradient, over several batches and across an optimiser step."""
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    L = TransductiveLoader(request.getfixturevalue(fixture))
    model = RED_GNN_trans(Options(hidden_dim=d, attn_dim=5, n_layer=3, dropout=0.0, act=act, n_rel=L.n_rel), L).cuda()
    model.train()
    opt = torch.optim.SGD(model.parameters(), lr=1e-2)
    nb = 6 if fixture == "hub_dir" else 12
    for it in range(3):
        tri = L.get_batch(np.arange(it * nb, (it + 1) * nb))
        res = {}
        for graph in (True, False):
            model.graph_train = graph
            loss = cuda_loss_backward(model, tri)
            res[graph] = (loss.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()})
        assert_close(res[True][0], res[False][0], 1e-5, "loss it%d" % it)
        floor = 1e-7 * max(float(g.abs().max()) for g in res[False][1].values())
        for k in res[True][1]:
            a, b = res[True][1][k], res[False][1][k]
            err, scale = (a - b).abs().max().item(), b.abs().max().item()
            assert err <= 2e-4 * scale or err <= floor, "it%d grad %s: err %.3e scale %.3e" % (it, k, err, scale)
        model.graph_train = True
        opt.step()                                           # parameters change in place; graphs follow
    assert len(model._train_graph_cache) == 1
    st = model.last_stats
    assert len(st["edges"]) == 3 and st["edges"][-1] > 0
    # dropout inside the captured graph: fresh mask per replay, finite gradients
    model.dropout.p = 0.25
    l1 = float(cuda_loss_backward(model, tri).detach())
    l2 = float(cuda_loss_backward(model, tri).detach())
    assert l1 != l2 and all(torch.isfinite(p.grad).all() for p in model.parameters())


def test_training_loop_graph_path_tracks_eager_path(tiny_dir):
    """25 Adam steps from the same initialisation (dropout 0): the graph-captured step and the eager
    autograd step follow the same loss curve, and the loss goes down."""
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    L = TransductiveLoader(tiny_dir)
    curves = {}
    for graph in (True, False):
        torch.manual_seed(5)
        model = RED_GNN_trans(Options(hidden_dim=48, attn_dim=5, n_layer=3, dropout=0.0, act="relu", n_rel=L.n_rel),
                              L).cuda()
        model.train()
        model.graph_train = graph
        opt = torch.optim.Adam(model.parameters(), lr=5e-3, weight_decay=1e-5)
        losses = []
        for it in range(25):
            tri = L.get_batch(np.arange((it % 5) * 20, (it % 5 + 1) * 20))
            opt.zero_grad()
            out = model(tri[:, 0], tri[:, 1])
            pos = out[torch.arange(len(out)).cuda(), torch.as_tensor(tri[:, 2]).cuda()]
            mx = out.max(1, keepdim=True)[0]
            loss = torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(out - mx), 1)))
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        curves[graph] = np.array(losses)
    assert curves[True][-5:].mean() < 0.9 * curves[True][:5].mean(), curves[True]
    rel = np.abs(curves[True] - curves[False]) / np.abs(curves[False])
    assert rel[:5].max() < 1e-4 and rel.max() < 2e-2, rel       # fp32 chaos grows slowly with Adam steps


def test_edge_case_batches(tiny_dir):
    """Empty batch, a single query, a batch with repeated queries, and an out-of-range subject."""
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans, _lib
    L, D = TransductiveLoader(tiny_dir), O.TransductiveData(tiny_dir)
    sd = O.init_state_dict(3, 48, 5, D.n_rel, seed=2)
    model = RED_GNN_trans(Options(n_rel=L.n_rel), L).cuda()
    model.load_state_dict(sd)
    model.eval()
    assert model(np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), mode="test").shape == (0, L.n_ent)
    subs, rels, _ = L.get_batch(np.arange(5), data="test")
    one = model(subs[:1], rels[:1], mode="test")
    assert_close(one, O.model_forward(sd, D.test_graph, subs[:1], rels[:1], 3, "relu"), 1e-4, "single query")
    rep_s, rep_r = np.repeat(subs[:2], 3), np.repeat(rels[:2], 3)
    rep = model(rep_s, rep_r, mode="test")
    assert torch.equal(rep[0], rep[1]) and torch.equal(rep[3], rep[5])
    assert_close(rep, O.model_forward(sd, D.test_graph, rep_s, rep_r, 3, "relu"), 1e-4, "repeated queries")
    with pytest.raises(_lib.RgError):
        model(np.array([L.n_ent]), np.array([0]), mode="test")
    model.train()                                   # same edge cases through the graph-captured training path
    out = model(subs[:1], rels[:1])
    out.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_optimizer_steps_track_the_oracle(tiny_dir):
    """Five Adam steps of the reference training recipe (sum loss, weight decay; base_model.py:27,49-62)
    run through the oracle on the CPU and through the CUDA path from the same initialisation:
    the loss sequences agree step by step (gradients are right, parameters stay in sync)."""
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    L, D = TransductiveLoader(tiny_dir), O.TransductiveData(tiny_dir)
    sd0 = O.init_state_dict(3, 48, 5, D.n_rel, seed=21)
    batches = [L.get_batch(np.arange(k * 10, (k + 1) * 10)) for k in range(5)]
    # oracle loop
    sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    opt = torch.optim.Adam(list(sd.values()), lr=3e-3, weight_decay=1e-5)
    want = []
    for tri in batches:
        opt.zero_grad()
        loss = O.train_loss(O.model_forward(sd, D.graph, tri[:, 0], tri[:, 1], 3, "relu"), tri[:, 2])
        loss.backward()
        opt.step()
        want.append(float(loss.detach()))
    # CUDA loop (graph-captured training step)
    model = RED_GNN_trans(Options(hidden_dim=48, attn_dim=5, n_layer=3, dropout=0.0, act="relu", n_rel=L.n_rel), L).cuda()
    model.load_state_dict(sd0)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=3e-3, weight_decay=1e-5)
    got = []
    for tri in batches:
        opt.zero_grad()
        got.append(float(cuda_loss_backward_train(model, tri).detach()))
        opt.step()
    rel = np.abs(np.array(got) - np.array(want)) / np.abs(np.array(want))
    assert rel.max() < 5e-4, (got, want)
    for k, p in model.named_parameters():
        assert_close(p, sd[k], 2e-3, "parameter after 5 steps " + k)


def cuda_loss_backward_train(model, tri):
    out = model(tri[:, 0], tri[:, 1])
    pos = out[torch.arange(len(out)).cuda(), torch.as_tensor(tri[:, 2]).cuda()]
    mx = out.max(1, keepdim=True)[0]
    loss = torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(out - mx), 1)))
    loss.backward()
    return loss


@pytest.mark.parametrize("fixture,n_layer", [("tiny_dir", 3), ("hub_dir", 2)])
def test_fused_loss_matches_dense_loss_and_gradients(request, fixture, n_layer):
    """model.loss(subs, rels, objs) (rg_node_loss on the per-node scores, no dense (n, n_ent) matrix) against
    the reference's dense formula (base_model.py:58-60) on model.forward's scores: value and every gradient;
    includes queries whose target entity is NOT visited (its score is the implicit 0)."""
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    L = TransductiveLoader(request.getfixturevalue(fixture))
    model = RED_GNN_trans(Options(hidden_dim=48, attn_dim=5, n_layer=n_layer, dropout=0.0, act="relu", n_rel=L.n_rel), L).cuda()
    model.train()
    tri = L.get_batch(np.arange(14)).copy()
    with torch.no_grad():
        sc = model(tri[:, 0], tri[:, 1])
    unvisited = (sc[0] == 0).nonzero().flatten()
    if len(unvisited):
        tri[0, 2] = int(unvisited[0])                       # a target outside the query's subgraph
    res = {}
    for fused in (False, True):
        model.zero_grad(set_to_none=True)
        if fused:
            loss = model.loss(tri[:, 0], tri[:, 1], tri[:, 2])
        else:
            out = model(tri[:, 0], tri[:, 1])
            pos = out[torch.arange(len(out)).cuda(), torch.as_tensor(tri[:, 2]).cuda()]
            mx = out.max(1, keepdim=True)[0]
            loss = torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(out - mx), 1)))
        loss.backward()
        res[fused] = (loss.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()})
    assert_close(res[True][0], res[False][0], 1e-5, "fused loss value")
    floor = 1e-7 * max(float(g.abs().max()) for g in res[False][1].values())
    for k in res[False][1]:
        a, b = res[True][1][k], res[False][1][k]
        err = float((a - b).abs().max())
        assert err <= 2e-4 * float(b.abs().max()) or err <= floor, "fused-loss grad %s: err %.3e" % (k, err)
    # scaled upstream gradient
    model.zero_grad(set_to_none=True)
    (0.5 * model.loss(tri[:, 0], tri[:, 1], tri[:, 2])).backward()
    for k, p in model.named_parameters():
        assert_close(p.grad, 0.5 * res[True][1][k], 1e-5, "scaled " + k)
