"""Shared test helpers (the oracle is imported here and only here / in tests)."""
import hashlib
import os

import numpy as np
import torch

from oracle import redgnn_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def sha(t):
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().contiguous().numpy()
    return hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()


def golden_state_dict(fx):
    return {k[3:]: torch.from_numpy(fx[k].copy()) for k in fx.files if k.startswith("sd.")}


def family_graphs(fx):
    """(train graph, test graph) oracle Graphs rebuilt from the family fixture
    (transductive/load_data.py:43-44)."""
    n_ent, n_rel = int(fx["n_ent"]), int(fx["n_rel"])
    fact = fx["fact_triple"].astype(np.int64).tolist()
    train = fx["train_triple"].astype(np.int64).tolist()
    g_train = O.Graph(O.add_inverse_block(fact, n_rel), n_ent, n_rel)
    g_test = O.Graph(O.add_inverse_block(fact, n_rel) + O.add_inverse_block(train, n_rel), n_ent, n_rel)
    return g_train, g_test


def graph_triples(g):
    """triples (without the self-loop block) of an oracle Graph as int64 [T,3]."""
    return g.KG[:-g.n_ent].astype(np.int64)


def device_graph(g, device="cuda"):
    from redgnn_b200 import DeviceGraph
    return DeviceGraph(graph_triples(g), g.n_ent, g.n_rel, device)


def assert_expansion_equal(got, want, tag=""):
    names = ("tail_nodes", "edges", "old_nodes_new_idx")
    for name, a, b in zip(names, got, want):
        a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
        b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
        assert a.dtype == np.int64, "%s %s dtype %s" % (tag, name, a.dtype)
        assert a.shape == b.shape, "%s %s shape %s vs %s" % (tag, name, a.shape, b.shape)
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            same_set = name == "edges" and np.array_equal(np.sort(a.view([("", a.dtype)] * a.shape[1]), axis=0),
                                                          np.sort(b.view([("", b.dtype)] * b.shape[1]), axis=0))
            raise AssertionError("%s %s differs at %d entries; first %s got %s want %s; same multiset of rows: %s"
                                 % (tag, name, len(bad), bad[0], a[tuple(bad[0])], b[tuple(bad[0])], same_set))


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def assert_close(a, b, rtol=1e-4, tag=""):
    """max |a-b| <= rtol * max|b|  (north_star: scores within 1e-4 relative in fp32)."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, "%s shape %s vs %s" % (tag, tuple(a.shape), tuple(b.shape))
    scale = b.abs().max().clamp_min(1e-30)
    err = (a - b).abs().max()
    assert err <= rtol * scale, "%s: max abs err %.3e vs scale %.3e (rel %.3e > %.1e)" % (
        tag, err.item(), scale.item(), (err / scale).item(), rtol)


def assert_close_yardstick(a, want64, ref_err32, rtol=1e-4, slack=2.0, tag=""):
    """Scores at the fp32 noise floor (hub entities: sums over 10^3..10^5 in-edges, activations ~10^3): the
    bar is `rtol` against the fp64 evaluation of the reference formula, or -- where the reference's OWN fp32
    evaluation is already that far from fp64 (`ref_err32`, relative to max |score|) -- `slack` times that."""
    a, b = a.detach().cpu().double(), want64.detach().cpu().double()
    scale = b.abs().max().clamp_min(1e-30)
    err = float((a - b).abs().max() / scale)
    bound = max(rtol, slack * float(ref_err32))
    assert err <= bound, "%s: rel err %.3e vs fp64 > %.3e (reference fp32 vs fp64: %.3e)" % (tag, err, bound, ref_err32)
    return err


def grad_floor(grads64):
    """Absolute floor for near-zero gradient tensors: 1e-7 x the largest gradient entry of the whole
    model ("rtol 1e-4, looser atol for near-zero entries")."""
    return 1e-7 * max(float(g.abs().max()) for g in grads64.values())


GRAD_BRANCHES = {"rtol": 0, "fp32-yardstick": 0, "floor": 0}      # how often each acceptance branch fired (conftest prints it)


def assert_grad_close(mine, ref32, ref64, rtol=1e-4, tag="", floor=0.0):
    """Gradient parity.  Pass if max|mine - ref64| <= rtol * max|ref64|, or -- for sums with heavy
    cancellation, where fp32 itself cannot hold rtol -- if the error is within 4x the error the
    reference's own fp32 arithmetic (ref32) makes against the fp64 evaluation of the same formula,
    or below `floor` (see grad_floor)."""
    m, r32, r64 = (t.detach().cpu().double() for t in (mine, ref32, ref64))
    assert m.shape == r64.shape, "%s shape %s vs %s" % (tag, tuple(m.shape), tuple(r64.shape))
    scale = r64.abs().max().clamp_min(1e-30)
    err, err_ref = (m - r64).abs().max(), (r32 - r64).abs().max()
    branch = "rtol" if err <= rtol * scale else ("fp32-yardstick" if err <= 4 * err_ref else ("floor" if err <= floor else None))
    if branch:
        GRAD_BRANCHES[branch] += 1
    assert err <= rtol * scale or err <= 4 * err_ref or err <= floor, \
        "%s: max abs err %.3e (fp32 reference itself: %.3e) vs scale %.3e (rel %.3e > %.1e)" % (
            tag, err.item(), err_ref.item(), scale.item(), (err / scale).item(), rtol)


def to64(sd):
    return {k: v.double() for k, v in sd.items()}
