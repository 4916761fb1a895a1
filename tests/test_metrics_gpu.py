"""GPU parity of the device-side filtered ranking (rg_filtered_ranks) with utils.cal_ranks
(oracle restatement pinned against the reference): identical ranks, including heavy ties
(unvisited entities all score exactly 0) and answers outside their own filter set."""
import numpy as np
import pytest
import torch

from oracle import redgnn_oracle as O
from redgnn_b200.synth import Options

pytestmark = pytest.mark.gpu


def dense(lists, n_ent):
    m = np.zeros((len(lists), n_ent))
    for i, l in enumerate(lists):
        m[i][np.asarray(l, dtype=np.int64)] = 1
    return m


@pytest.mark.parametrize("n,n_ent,seed", [(7, 300, 0), (33, 1000, 1), (4, 14541, 2)])
def test_ranks_equal_cal_ranks_with_ties(n, n_ent, seed):
    from redgnn_b200.metrics import filtered_ranks, rank_metrics
    rng = np.random.default_rng(seed)
    scores = rng.normal(size=(n, n_ent)).astype(np.float32)
    scores[rng.random((n, n_ent)) < 0.6] = 0.0                      # unvisited entities: exact ties
    scores[:, : n_ent // 10] = np.round(scores[:, : n_ent // 10], 1)   # more duplicate values
    answers = [np.unique(rng.integers(0, n_ent, rng.integers(1, 6))) for _ in range(n)]
    filters = [np.unique(np.concatenate([a, rng.integers(0, n_ent, rng.integers(0, 30))])) for a in answers]
    filters[0] = filters[0][~np.isin(filters[0], answers[0][:1])]     # one answer outside its filter set
    want = np.array(O.cal_ranks(scores, dense(answers, n_ent), dense(filters, n_ent)))
    got = filtered_ranks(torch.as_tensor(scores).cuda(), answers, filters)
    assert got.dtype == torch.float64 and np.array_equal(got.cpu().numpy(), want)
    for a, b in zip(rank_metrics(got), O.cal_performance(want)):
        assert abs(a - b) < 1e-12


def test_model_eval_metrics_on_device(tiny_dir):
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans
    from redgnn_b200.metrics import filtered_ranks
    L = TransductiveLoader(tiny_dir)
    model = RED_GNN_trans(Options(n_rel=L.n_rel), L).cuda().eval()
    idx = np.arange(60)
    subs, rels, objs = L.get_batch(idx, data="test")
    with torch.no_grad():
        scores = model(subs, rels, mode="test")
    filt = [L.filters[(s, r)] for s, r in zip(subs, rels)]
    want = np.array(O.cal_ranks(scores.cpu().numpy(), objs, dense(filt, L.n_ent)))
    got = filtered_ranks(scores, [L.test_a[i] for i in idx], filt)
    assert np.array_equal(got.cpu().numpy(), want)
