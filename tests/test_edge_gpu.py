"""GPU parity: fused edge kernels (explicit edge lists through GNNLayer.forward, and the implicit
pull path) vs the oracle / torch fp32 reference -- tolerance 1e-4 relative (north_star)."""
import numpy as np
import pytest
import torch

from oracle import redgnn_oracle as O
from helpers import device_graph, assert_close, assert_grad_close, to64

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _check_heavy_queues():
    from redgnn_b200 import ops
    ops.DEBUG_CHECK_HEAVY = True
    yield
    ops.DEBUG_CHECK_HEAVY = False


ACT = {"relu": torch.relu, "tanh": torch.tanh, "idd": lambda x: x}


def make_layer(d, a, n_rel, act, seed):
    from redgnn_b200 import GNNLayer
    torch.manual_seed(seed)
    layer = GNNLayer(d, d, a, n_rel, act=ACT[act])
    sd = {"gnn_layers.0." + k: v.detach().clone() for k, v in layer.state_dict().items()}
    return layer.cuda(), sd


def oracle_layer(sd, q_rel, hidden, edges, n_node, act, w=None):
    """fp32 oracle (and, given the output cotangent w, its gradients in fp32 and fp64)."""
    res = []
    for conv in ((lambda t: t.clone()), (lambda t: t.double())):
        sdc = {k: conv(v).requires_grad_(True) for k, v in sd.items()}
        hid = conv(hidden).requires_grad_(True)
        out = O.gnn_layer_forward(sdc, 0, q_rel, hid, edges, n_node, act)
        if w is not None:
            (out * conv(w)).sum().backward()
        res.append((out, sdc, hid))
    return res


def check_layer_grads(layer, hid_c, res, tag):
    (_, sd32, hid32), (_, sd64, hid64) = res
    assert_grad_close(hid_c.grad, hid32.grad, hid64.grad, 1e-4, tag + " grad hidden")
    for k, p in layer.named_parameters():
        key = "gnn_layers.0." + k
        assert_grad_close(p.grad, sd32[key].grad, sd64[key].grad, 1e-4, tag + " grad " + k)


@pytest.mark.parametrize("d,a,act", [(48, 5, "relu"), (32, 3, "tanh"), (64, 5, "idd"), (16, 8, "relu")])
def test_explicit_layer_forward_backward(tiny_dir, d, a, act):
    D = O.TransductiveData(tiny_dir)
    g = D.test_graph
    n = 6
    rng = np.random.default_rng(d)
    subs, rels = rng.integers(0, D.n_ent, n), rng.integers(0, 2 * D.n_rel, n)
    nodes = np.stack([np.arange(n), subs], 1)
    nodes1, _, _ = O.get_neighbors(g, nodes)
    nodes2, edges, remap = O.get_neighbors(g, nodes1.numpy())
    torch.manual_seed(0)
    hidden = torch.randn(nodes1.shape[0], d)
    q_rel = torch.as_tensor(rels)
    layer, sd = make_layer(d, a, D.n_rel, act, seed=1)
    torch.manual_seed(3)
    w = torch.randn(nodes2.shape[0], d)
    res = oracle_layer(sd, q_rel, hidden, edges, nodes2.shape[0], act, w)
    hid_c = hidden.cuda().requires_grad_(True)
    got = layer(torch.as_tensor(subs).cuda(), q_rel.cuda(), hid_c, edges.cuda(), nodes2.shape[0], remap.cuda())
    assert_close(got, res[0][0], 1e-4, "forward")
    (got * w.cuda()).sum().backward()
    check_layer_grads(layer, hid_c, res, "explicit")


def test_explicit_layer_empty_segments_and_zero_hidden(tiny_dir):
    D = O.TransductiveData(tiny_dir)
    nodes = np.stack([np.arange(3), np.array([1, 2, 3])], 1)
    nodes1, edges, remap = O.get_neighbors(D.graph, nodes)
    layer, sd = make_layer(48, 5, D.n_rel, "relu", seed=2)
    hidden = torch.zeros(3, 48)
    q_rel = torch.tensor([0, 3, 5])
    n_node = nodes1.shape[0] + 5                      # trailing segments without edges -> zeros
    want = O.gnn_layer_forward(sd, 0, q_rel, hidden, edges, n_node, "relu")
    with torch.no_grad():
        got = layer(None, q_rel.cuda(), hidden.cuda(), edges.cuda(), n_node, remap.cuda())
    assert_close(got, want, 1e-4, "zero hidden")
    assert float(got[-5:].abs().max()) == 0.0


def test_heavy_segments_explicit_and_deterministic(hub_dir):
    """Segments far longer than RG_HEAVY_CHUNK (hub entities): chunk queue + ordered fix-up."""
    D = O.TransductiveData(hub_dir)
    g = D.graph
    n = 6
    subs = np.arange(n)
    nodes0, _, _ = O.get_neighbors(g, np.stack([np.arange(n), subs], 1))
    nodes1, _, _ = O.get_neighbors(g, nodes0.numpy())
    nodes2, edges, remap = O.get_neighbors(g, nodes1.numpy())
    seg_len = torch.bincount(edges[:, 5])
    assert int(seg_len.max()) > 1024, "fixture must contain heavy segments (max %d)" % int(seg_len.max())
    layer, sd = make_layer(48, 5, D.n_rel, "relu", seed=3)
    torch.manual_seed(1)
    hidden = torch.randn(nodes1.shape[0], 48)
    q_rel = torch.arange(n) % (2 * D.n_rel)
    w = torch.randn(nodes2.shape[0], 48)
    res = oracle_layer(sd, q_rel, hidden, edges, nodes2.shape[0], "relu", w)
    hid_c = hidden.cuda().requires_grad_(True)
    got = layer(None, q_rel.cuda(), hid_c, edges.cuda(), nodes2.shape[0], None)
    assert_close(got, res[0][0], 1e-4, "heavy forward")
    got2 = layer(None, q_rel.cuda(), hid_c, edges.cuda(), nodes2.shape[0], None)
    assert torch.equal(got, got2), "forward must be bit-reproducible"
    (got * w.cuda()).sum().backward()
    check_layer_grads(layer, hid_c, res, "heavy")


@pytest.mark.parametrize("fixture,extra_hops", [("tiny_dir", 0), ("hub_dir", 1)])
def test_implicit_matches_explicit(request, fixture, extra_hops):
    """The pull path (no edge list in HBM) against the explicit path on the same hop."""
    from redgnn_b200.ops import Segments
    D = O.TransductiveData(request.getfixturevalue(fixture))
    g = D.graph
    dg = device_graph(g)
    n = 9
    subs = torch.arange(n) * 3 % D.n_ent
    fr0 = dg.frontier_from_nodes(torch.stack([torch.arange(n), subs], 1).cuda(), n)
    for _ in range(extra_hops):
        fr0 = dg.step(fr0)
        fr0.read_counts()
    fr1 = dg.step(fr0)
    _, _, n1, _ = fr1.read_counts()
    fr2 = dg.step(fr1)
    _, e2, n2, _ = fr2.read_counts()
    b1, e1 = fr1.nodes32(n1)
    b2, ent2 = fr2.nodes32(n2)
    nodes1 = torch.stack([b1, e1], 1).long()
    _, edges, _ = dg.get_neighbors(nodes1, n_query=n)
    assert edges.shape[0] == e2
    layer, _ = make_layer(48, 5, D.n_rel, "relu", seed=4)
    torch.manual_seed(2)
    q_rel = (torch.arange(n) % (2 * D.n_rel)).cuda()
    h_exp = torch.randn(n1, 48, device="cuda").requires_grad_(True)
    h_imp = h_exp.detach().clone().requires_grad_(True)
    out_exp = layer(None, q_rel, h_exp, edges, n2, None)
    fwd = Segments.implicit(b2, ent2, dg.in_ptr, dg.in_adj, fr1, dg.heavy_in)
    bwd = Segments.implicit(b1, e1, dg.out_ptr, dg.out_adj, fr2, dg.heavy_out)
    out_imp = layer.propagate(q_rel, h_imp, fwd, bwd)
    assert_close(out_imp, out_exp, 1e-5, "implicit forward")
    w = torch.randn_like(out_exp)
    grads_exp = torch.autograd.grad((out_exp * w).sum(), [h_exp] + list(layer.parameters()))
    grads_imp = torch.autograd.grad((out_imp * w).sum(), [h_imp] + list(layer.parameters()))
    for ge, gi, name in zip(grads_exp, grads_imp, ["hidden"] + [k for k, _ in layer.named_parameters()]):
        assert_close(gi, ge, 1e-4, "implicit grad " + name)


def test_layer_with_no_edges():
    layer, sd = make_layer(48, 5, 4, "tanh", seed=6)
    hidden = torch.randn(3, 48, device="cuda")
    out = layer(None, torch.tensor([0, 1, 2]).cuda(), hidden, torch.zeros((0, 6), dtype=torch.long, device="cuda"), 7,
                None)
    assert out.shape == (7, 48) and float(out.abs().max()) == 0.0
