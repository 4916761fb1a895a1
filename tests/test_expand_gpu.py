"""GPU parity: CUDA frontier expansion (through the C ABI) vs the oracle -- bit-exact, in the
reference's order (north_star: "expanded node/edge lists and index remaps bit-exact")."""
import numpy as np
import pytest
import torch

from oracle import redgnn_oracle as O
from helpers import golden, family_graphs, device_graph, assert_expansion_equal, sha
from test_oracle_golden import fb237_graphs, SURVEY_HASHES

pytestmark = pytest.mark.gpu


def run_layers(g, dg, nodes, n_layer, tag, n_query=None):
    for l in range(n_layer):
        want = O.get_neighbors(g, nodes)
        got = dg.get_neighbors(nodes, n_query=n_query)
        assert all(t.is_cuda for t in got)
        assert_expansion_equal(got, want, "%s layer %d" % (tag, l))
        nodes = want[0].numpy()


def test_golden_family_and_survey_hashes():
    fx = golden("family")
    g_train, g_test = family_graphs(fx)
    for tag, g in (("train20", g_train), ("test50", g_test)):
        dg = device_graph(g)
        subs = fx[tag + "_subs"]
        nodes = np.stack([np.arange(len(subs)), subs], 1)
        for l in range(3):
            tn, ed, rm = dg.get_neighbors(nodes)
            assert (sha(tn), sha(ed), sha(rm)) == tuple(fx["%s_L%d_sha" % (tag, l)])
            assert (sha(tn)[:16], sha(ed)[:16], sha(rm)[:16]) == SURVEY_HASHES[("family", tag)][l]
            nodes = tn.cpu().numpy()


def test_golden_fb237_v2_inductive():
    fx = golden("fb237_v2")
    g_tra, g_ind = fb237_graphs(fx)
    for tag, g in (("tra10", g_tra), ("ind10", g_ind)):
        dg = device_graph(g)
        subs = fx[tag + "_subs"]
        nodes = np.stack([np.arange(len(subs)), subs], 1)
        for l in range(3):
            tn, ed, rm = dg.get_neighbors(nodes)
            assert (sha(tn), sha(ed), sha(rm)) == tuple(fx["%s_L%d_sha" % (tag, l)])
            nodes = tn.cpu().numpy()


@pytest.mark.parametrize("n_query", [1, 5, 32, 33, 70, 200, 1000])
def test_synthetic_vs_oracle(tiny_dir, n_query):
    D = O.TransductiveData(tiny_dir)
    g = D.test_graph
    dg = device_graph(g)
    rng = np.random.default_rng(n_query)
    subs = rng.integers(0, D.n_ent, n_query)
    run_layers(g, dg, np.stack([np.arange(n_query), subs], 1), 3, "tiny n=%d" % n_query)


def test_hub_graph_vs_oracle(hub_dir):
    D = O.TransductiveData(hub_dir)
    dg = device_graph(D.graph)
    subs = np.arange(40) % D.n_ent
    run_layers(D.graph, dg, np.stack([np.arange(40), subs], 1), 3, "hub")


def test_unsorted_duplicate_and_sparse_batch_input(tiny_dir):
    D = O.TransductiveData(tiny_dir)
    g = D.graph
    dg = device_graph(g)
    rng = np.random.default_rng(1)
    nodes = np.stack([rng.integers(0, 9, 60), rng.integers(0, D.n_ent, 60)], 1)
    nodes = np.concatenate([nodes, nodes[:11]], 0)          # duplicates, unsorted, batch ids with gaps
    want = O.get_neighbors(g, nodes)
    got = dg.get_neighbors(nodes)
    assert_expansion_equal(got, want, "unsorted")
    # also as a CUDA tensor input
    got = dg.get_neighbors(torch.as_tensor(nodes).cuda())
    assert_expansion_equal(got, want, "unsorted-cuda")


def test_out_of_range_node_is_reported(tiny_dir):
    from redgnn_b200 import _lib
    D = O.TransductiveData(tiny_dir)
    dg = device_graph(D.graph)
    with pytest.raises(_lib.RgError):
        dg.get_neighbors(np.array([[0, D.n_ent]]))


def test_loader_get_neighbors_modes(tiny_dir, induc_dir):
    from redgnn_b200 import TransductiveLoader, InductiveLoader
    L, D = TransductiveLoader(tiny_dir), O.TransductiveData(tiny_dir)
    nodes = np.stack([np.arange(6), L.train_data[:6, 0]], 1)
    for mode in ("train", "valid", "test"):
        assert_expansion_equal(L.get_neighbors(nodes, mode), O.get_neighbors(D.graph_for(mode), nodes), mode)
    L2, D2 = InductiveLoader(induc_dir), O.InductiveData(induc_dir)
    for mode, subs in (("transductive", L2.tra_train[:5, 0]), ("inductive", np.array(L2.test_q[:5])[:, 0])):
        nodes = np.stack([np.arange(5), subs], 1)
        assert_expansion_equal(L2.get_neighbors(nodes, mode), O.get_neighbors(D2.graph_for(mode), nodes), mode)


def test_size_independent_properties_at_scale():
    """FB15k-237-shaped synthetic KG at full size: properties that need no oracle run."""
    from redgnn_b200 import synth, DeviceGraph
    n_ent, n_rel, n_tri = 14541, 237, 272115
    tri = synth.zipf_triples(n_ent, n_rel, n_tri, 0.8, 0.8, seed=0)
    inv = np.stack([tri[:, 2], tri[:, 1] + n_rel, tri[:, 0]], 1)
    dg = DeviceGraph(np.concatenate([tri, inv], 0), n_ent, n_rel, "cuda")
    n = 16
    nodes = torch.stack([torch.arange(n), torch.arange(n) * 7], 1).cuda()
    deg = torch.bincount(dg.head.long(), minlength=n_ent)
    for l in range(3):
        tn, ed, rm = dg.get_neighbors(nodes, n_query=n)
        key_in = nodes[:, 0] * n_ent + nodes[:, 1]
        key_out = tn[:, 0] * n_ent + tn[:, 1]
        assert bool((key_out[1:] > key_out[:-1]).all())                       # sorted unique
        assert ed.shape[0] == int(deg[nodes[:, 1]].sum())                     # E = sum of degrees
        assert torch.equal(tn[rm], nodes)                                     # remap round trip
        assert torch.equal(nodes[ed[:, 4]], ed[:, [0, 1]])                    # head_index consistent
        assert torch.equal(tn[ed[:, 5]], ed[:, [0, 3]])                       # tail_index consistent
        fact_key = ed[:, 1] * (2 * n_rel + 1) * n_ent + ed[:, 2] * n_ent + ed[:, 3]
        assert torch.isin(key_out, ed[:, 0] * n_ent + ed[:, 3]).all()
        loops = ed[ed[:, 2] == 2 * n_rel]
        assert loops.shape[0] == nodes.shape[0]
        again = dg.get_neighbors(nodes, n_query=n)
        assert all(torch.equal(a, b) for a, b in zip((tn, ed, rm), again))    # deterministic
        del fact_key, key_in
        nodes = tn


@pytest.mark.parametrize("fixture", ["tiny_dir", "hub_dir"])
def test_graph_build_matches_host_construction(request, fixture):
    """rg_graph_build (device radix sort) against the host construction of the same CSR views."""
    from redgnn_b200 import DeviceGraph
    from helpers import graph_triples
    D = O.TransductiveData(request.getfixturevalue(fixture))
    tri = graph_triples(D.test_graph)
    g_dev = DeviceGraph(tri, D.n_ent, D.n_rel, "cuda")
    g_cpu = DeviceGraph(tri, D.n_ent, D.n_rel, "cpu")
    for name in ("head", "rel", "tail", "in_ptr", "in_adj", "out_ptr", "out_adj"):
        assert torch.equal(getattr(g_dev, name).cpu(), getattr(g_cpu, name)), name
    assert g_dev.heavy_in == g_cpu.heavy_in and g_dev.heavy_out == g_cpu.heavy_out
