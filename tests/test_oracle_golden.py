"""CPU, runs everywhere: the oracle against the committed fixtures the live reference produced
(oracle/make_golden.py) and against the SURVEY section 4 known-answer hashes."""
import numpy as np
import pytest
import torch

from oracle import redgnn_oracle as O
from helpers import golden, golden_state_dict, family_graphs, sha

# SHA-256 prefixes of the raw int64 outputs of the unmodified reference
SURVEY_HASHES = {
    ("family", "train20"): [("ab688a95807e767f", "659cf6d1c31237f3", "107a4f20d2e55ad9"),
                            ("3aa1d652a32fc87b", "2681debf866e99d5", "dae329fe52736442"),
                            ("a207741253bac756", "1096bc6f63296737", "b93cd18a59d95940")],
    ("family", "test50"): [("a3d7629a533e3e3f", "80ef4f7392726575", "7bc2d5a4ea1d9bf5"),
                           ("f9858396d088bde7", "543bed9e216f2b52", "42775cc12df7aa3b"),
                           ("2386b41279a14cbf", "12eb9de6bd41c012", "75a562bc51d3b9ef")],
    ("fb237_v2", "tra10"): [("89054d4990f16f9a", "7630af3ad0c89c38", "32f115a1f24a0dae"),
                            ("c74466ef8e7bc660", "88832b142131533b", "e51b6961e2993963"),
                            ("5acd31e7c1759ebb", "5210705b4bd7b77b", "5332c6340b55b6a8")],
    ("fb237_v2", "ind10"): [("68f379fc0565aeb3", "e554e2022a4d74e8", "417fd17381434c0d"),
                            ("9ddd105ae6825ad6", "75c4267c3e42cb19", "7b73392b2dbb51d6"),
                            ("d322b4385af73029", "ca75f3db41046cc7", "a026087076ca712f")],
}


def fb237_graphs(fx):
    n_rel = int(fx["n_rel"])
    return (O.Graph(fx["tra_triples"].astype(np.int64), int(fx["n_ent"]), n_rel),
            O.Graph(fx["ind_triples"].astype(np.int64), int(fx["n_ent_ind"]), n_rel))


def cases():
    fam, fb = golden("family"), golden("fb237_v2")
    g_train, g_test = family_graphs(fam)
    g_tra, g_ind = fb237_graphs(fb)
    return [("family", "train20", fam, g_train), ("family", "test50", fam, g_test),
            ("fb237_v2", "tra10", fb, g_tra), ("fb237_v2", "ind10", fb, g_ind)]


@pytest.mark.parametrize("case", cases(), ids=lambda c: c[0] + "-" + c[1])
def test_expansion_hashes(case):
    ds, tag, fx, g = case
    subs = fx[tag + "_subs"]
    nodes = np.stack([np.arange(len(subs)), subs], 1)
    for l in range(3):
        tn, ed, rm = O.get_neighbors(g, nodes)
        assert list(fx["%s_L%d_sizes" % (tag, l)]) == [len(nodes), len(ed), len(tn)]
        got = (sha(tn), sha(ed), sha(rm))
        assert got == tuple(fx["%s_L%d_sha" % (tag, l)])
        assert tuple(h[:16] for h in got) == SURVEY_HASHES[(ds, tag)][l]
        nodes = tn.numpy()


def test_expansion_full_arrays_and_definition():
    fx = golden("family")
    _, g = family_graphs(fx)
    h, r, t = g.int_arrays()
    nodes = np.stack([np.arange(4), fx["test50_subs"][:4]], 1)
    for l in range(3):
        tn, ed, rm = O.get_neighbors(g, nodes)
        tn2, ed2, rm2 = O.expand_definition(h, r, t, g.n_ent, nodes)
        for a, b, key in ((tn, tn2, "nodes"), (ed, ed2, "edges"), (rm, rm2, "remap")):
            want = fx["test4_L%d_%s" % (l, key)].astype(np.int64)
            assert np.array_equal(a.numpy(), want) and np.array_equal(b, want)
        nodes = tn.numpy()


@pytest.mark.parametrize("name", ["family", "fb237_v2"])
def test_scores_grads_and_ranks(name):
    fx = golden(name)
    sd = golden_state_dict(fx)
    if name == "family":
        g_train, g_eval = family_graphs(fx)
        n_out = None
    else:
        g_train, g_eval = fb237_graphs(fx)
        n_out = int(fx["n_ent_ind"])
    scores = O.model_forward(sd, g_eval, fx["eval_subs"], fx["eval_rels"], 3, "relu", n_ent_out=n_out)
    want = torch.from_numpy(fx["eval_scores"])
    assert (scores - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    assert torch.equal(scores == 0, want == 0)
    ranks = O.cal_ranks(fx["eval_scores"].astype(np.float32), fx["eval_objs"].astype(np.float64),
                        fx["eval_filters"].astype(np.float64))
    assert np.array_equal(np.array(ranks), fx["eval_ranks"])
    mrr, h1, h10 = O.cal_performance(ranks)
    assert 0 < mrr <= 1 and 0 <= h1 <= h10 <= 1
    tri = fx["train_triples"]
    sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss = O.train_loss(O.model_forward(sd_g, g_train, tri[:, 0], tri[:, 1], 3, "relu"), tri[:, 2])
    loss.backward()
    assert abs(loss.item() - float(fx["train_loss"])) <= 1e-5 * abs(float(fx["train_loss"]))
    for k in sd:
        want_g = torch.from_numpy(fx["train_grad." + k])
        scale = want_g.abs().max().item() + 1e-12
        assert (sd_g[k].grad - want_g).abs().max().item() <= 1e-4 * scale, k
