"""CPU, build container only: pin oracle/redgnn_oracle.py against the LIVE unmodified reference
(skipped where /root/reference is not mounted, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ref_import as R
from oracle import redgnn_oracle as O

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def family():
    return (R.make_loader("transductive", R.data_dir("transductive", "family")),
            O.TransductiveData(R.data_dir("transductive", "family")))


@pytest.fixture(scope="module")
def fb237():
    return (R.make_loader("inductive", R.data_dir("inductive", "fb237_v2")),
            O.InductiveData(R.data_dir("inductive", "fb237_v2")))


def test_transductive_data_matches(family):
    L, D = family
    assert np.array_equal(L.KG, D.graph.KG) and np.array_equal(L.tKG, D.test_graph.KG)
    assert np.array_equal(L.train_data, D.train_data)
    assert L.valid_q == D.valid_q and L.test_q == D.test_q
    assert all(np.array_equal(a, b) for a, b in zip(L.test_a, D.test_a))
    assert {k: set(v) for k, v in L.filters.items()} == {k: set(v) for k, v in D.filters.items()}


def test_inductive_data_matches(fb237):
    L, D = fb237
    assert np.array_equal(L.tra_KG, D.tra_graph.KG) and np.array_equal(L.ind_KG, D.ind_graph.KG)
    assert np.array_equal(L.tra_train, D.train_data)
    assert L.valid_q == D.valid_q and L.test_q == D.test_q
    assert {k: set(v) for k, v in L.val_filters.items()} == {k: set(v) for k, v in D.val_filters.items()}
    assert {k: set(v) for k, v in L.tst_filters.items()} == {k: set(v) for k, v in D.tst_filters.items()}


@pytest.mark.parametrize("mode,n", [("train", 20), ("test", 7)])
def test_get_neighbors_port_and_definition(family, mode, n):
    L, D = family
    g = D.graph_for(mode)
    h, r, t = g.int_arrays()
    nodes = np.stack([np.arange(n), L.train_data[:n, 0]], 1)
    for _ in range(3):
        want = L.get_neighbors(nodes, mode)
        got = O.get_neighbors(g, nodes)
        dfn = O.expand_definition(h, r, t, g.n_ent, nodes)
        for a, b, c in zip(want, got, dfn):
            assert torch.equal(a, b)
            assert np.array_equal(a.numpy(), c)
        nodes = want[0].numpy()


def test_get_neighbors_unsorted_duplicate_input(family):
    L, D = family
    rng = np.random.default_rng(0)
    nodes = np.stack([rng.integers(0, 5, 40), rng.integers(0, L.n_ent, 40)], 1)
    nodes = np.concatenate([nodes, nodes[:7]], 0)
    want = L.get_neighbors(nodes, "train")
    h, r, t = D.graph.int_arrays()
    dfn = O.expand_definition(h, r, t, D.n_ent, nodes)
    for a, c in zip(want, dfn):
        assert np.array_equal(a.numpy(), c)


def test_model_forward_and_loss_grads(family):
    L, D = family
    _, M, U = R.load_reference("transductive")
    opts = R.family_options(L)
    torch.manual_seed(1234)
    model = M.RED_GNN_trans(opts, L)
    model.eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    assert all(torch.equal(sd[k], v) for k, v in O.init_state_dict(3, 48, 5, L.n_rel, 1234).items())
    subs, rels, objs = L.get_batch(np.arange(12), data="test")
    ref = model(subs, rels, mode="test").detach()
    mine = O.model_forward(sd, D.test_graph, subs, rels, 3, "relu")
    assert (ref - mine).abs().max().item() < 2e-6
    assert torch.equal(ref == 0, mine == 0)
    # grads of the training loss
    tri = L.train_data[:6]
    scores = model(tri[:, 0], tri[:, 1])
    model.zero_grad()
    pos = scores[torch.arange(6), torch.LongTensor(tri[:, 2])]
    mx = torch.max(scores, 1, keepdim=True)[0]
    torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(scores - mx), 1))).backward()
    sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    O.train_loss(O.model_forward(sd_g, D.graph, tri[:, 0], tri[:, 1], 3, "relu"), tri[:, 2]).backward()
    for k, p in model.named_parameters():
        scale = p.grad.abs().max().item() + 1e-12
        assert (p.grad - sd_g[k].grad).abs().max().item() <= 1e-4 * scale, k
    # metrics
    filt = np.zeros((12, L.n_ent))
    for i in range(12):
        filt[i][np.array(L.filters[(subs[i], rels[i])])] = 1
    assert U.cal_ranks(ref.numpy(), objs, filt) == O.cal_ranks(ref.numpy(), objs, filt)


def test_inductive_forward(fb237):
    L, D = fb237
    _, M, _ = R.load_reference("inductive")
    torch.manual_seed(1234)
    model = M.RED_GNN_induc(R.fb237_v2_options(L), L)
    model.eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    subs, rels, _ = L.get_batch(np.arange(6), data="test")
    ref = model(subs, rels, "inductive").detach()
    mine = O.model_forward(sd, D.ind_graph, subs, rels, 3, "relu", n_ent_out=D.n_ent_ind)
    assert ref.shape == mine.shape and (ref - mine).abs().max().item() < 2e-6


def test_staged_reference_copy_is_byte_identical():
    """oracle/_ref (what the GPU box runs as `kind: "reference"`) == the mounted reference files."""
    import filecmp
    import os
    from oracle import build_ref
    if not R.is_live():
        pytest.skip("reference tree not mounted (running from the staged copy)")
    assert build_ref.stage() > 0
    for setting in ("transductive", "inductive"):
        for f in build_ref.FILES:
            assert filecmp.cmp(os.path.join(build_ref.SRC, "Static", setting, f),
                               os.path.join(build_ref.DEST, "Static", setting, f), shallow=False), f
