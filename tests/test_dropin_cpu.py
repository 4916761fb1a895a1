"""CPU: host-side mirror of the reference surface -- loaders against the oracle's data readers,
state_dict interchange, and (build container only) the reference's own base_model.py resolving
`from models import ...` / `from load_data import ...` to this package by sys.path order."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from oracle import redgnn_oracle as O
from oracle import ref_import as R
from redgnn_b200.synth import Options

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_transductive_loader_matches_oracle_reader(tiny_dir):
    from redgnn_b200 import TransductiveLoader
    L, D = TransductiveLoader(tiny_dir, device="cpu"), O.TransductiveData(tiny_dir)
    assert (L.n_ent, L.n_rel) == (D.n_ent, D.n_rel)
    assert np.array_equal(L.KG, D.graph.KG.astype(np.int64)) and np.array_equal(L.tKG, D.test_graph.KG.astype(np.int64))
    assert np.array_equal(L.train_data, D.train_data)
    assert L.valid_q == D.valid_q and L.test_q == D.test_q
    assert all(np.array_equal(a, b) for a, b in zip(L.test_a, D.test_a))
    assert {k: set(v) for k, v in L.filters.items()} == {k: set(v) for k, v in D.filters.items()}
    subs, rels, objs = L.get_batch(np.arange(5), data="test")
    assert objs.shape == (5, L.n_ent) and objs.sum() == sum(len(D.test_a[i]) for i in range(5))
    assert L.get_batch(np.arange(3)).shape == (3, 3)
    g = L.graph_for("train", "cpu")
    assert g.n_fact == L.n_fact == len(D.graph.KG)
    # CSR views: every fact appears once, grouped by tail / head, stable in fact order
    kg = L.KG
    order = np.argsort(kg[:, 2], kind="stable")
    assert np.array_equal(g.in_adj.numpy(), kg[order][:, [0, 1]])
    order = np.argsort(kg[:, 0], kind="stable")
    assert np.array_equal(g.out_adj.numpy(), kg[order][:, [2, 1]])
    assert int(g.in_ptr[-1]) == g.n_fact and int(g.out_ptr[-1]) == g.n_fact


def test_shuffle_train_follows_reference_rng_stream(tiny_dir):
    from redgnn_b200 import TransductiveLoader
    L = TransductiveLoader(tiny_dir, device="cpu")
    np.random.seed(7)
    L.shuffle_train()
    np.random.seed(7)
    allt = np.concatenate([np.array(L.fact_triple), np.array(L.train_triple)], 0)
    perm = np.random.permutation(len(allt))                       # transductive/load_data.py:157
    n_fact = len(allt) * 3 // 4
    assert np.array_equal(np.array(L.fact_data)[:n_fact], allt[perm][:n_fact])
    assert L.n_train == 2 * (len(allt) - n_fact) and L.n_fact == 2 * n_fact + L.n_ent


def test_inductive_loader_matches_oracle_reader(induc_dir):
    from redgnn_b200 import InductiveLoader
    L, D = InductiveLoader(induc_dir, device="cpu"), O.InductiveData(induc_dir)
    assert (L.n_ent, L.n_ent_ind, L.n_rel) == (D.n_ent, D.n_ent_ind, D.n_rel)
    assert np.array_equal(L.tra_KG, D.tra_graph.KG.astype(np.int64))
    assert np.array_equal(L.ind_KG, D.ind_graph.KG.astype(np.int64))
    assert np.array_equal(L.tra_train, D.train_data)
    assert L.valid_q == D.valid_q and L.test_q == D.test_q
    assert {k: set(v) for k, v in L.tst_filters.items()} == {k: set(v) for k, v in D.tst_filters.items()}
    assert L.get_batch(np.arange(4), data="test")[2].shape == (4, L.n_ent_ind)


def test_state_dict_names_match_reference_and_no_cpu_path(tiny_dir):
    from redgnn_b200 import TransductiveLoader, RED_GNN_trans, _lib
    L = TransductiveLoader(tiny_dir, device="cpu")
    sd = O.init_state_dict(3, 48, 5, L.n_rel, seed=1)
    model = RED_GNN_trans(Options(n_rel=L.n_rel), L)
    assert set(model.state_dict()) == set(sd)
    model.load_state_dict(sd)
    with pytest.raises(_lib.RgError):
        model(np.array([0, 1]), np.array([0, 1]))                 # CPU model: refuses, no fallback
    with pytest.raises(ValueError):
        RED_GNN_trans(Options(n_rel=L.n_rel, hidden_dim=40), L)


@pytest.mark.skipif(not R.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("setting,cls", [("transductive", "RED_GNN_trans"), ("inductive", "RED_GNN_induc")])
def test_reference_base_model_resolves_to_this_package(setting, cls, tiny_dir, induc_dir):
    """Static/<setting>/base_model.py imported UNMODIFIED with drop_in/<setting> first on sys.path."""
    R._install_shims()                                            # identity .cuda() on the CPU-only box
    shim = os.path.join(ROOT, "redgnn_b200", "drop_in", setting)
    saved_path, saved_mods = list(sys.path), {k: sys.modules.pop(k, None) for k in ("models", "load_data", "utils")}
    sys.path[:0] = [shim, os.path.join(R.REF_ROOT, "Static", setting)]
    try:
        spec = importlib.util.spec_from_file_location(
            "ref_base_model_" + setting, os.path.join(R.REF_ROOT, "Static", setting, "base_model.py"))
        bm = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bm)
        import load_data
        import redgnn_b200
        assert getattr(bm, cls) is getattr(redgnn_b200, cls)
        loader = load_data.DataLoader(tiny_dir if setting == "transductive" else induc_dir)
        assert isinstance(loader, (redgnn_b200.TransductiveLoader, redgnn_b200.InductiveLoader))
        opts = Options(n_rel=loader.n_rel, n_ent=loader.n_ent)
        trainer = bm.BaseModel(opts, loader)                      # Adam over our parameters, unchanged code
        assert isinstance(trainer.model, getattr(redgnn_b200, cls))
        assert len(list(trainer.model.parameters())) == len(trainer.optimizer.param_groups[0]["params"])
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def test_vectorised_query_grouping_equals_the_reference_row_loop():
    """data._group_queries (lexsort + split) against the reference's sort + dict loop
    (transductive/load_data.py:91-104) on random, duplicate-heavy and empty inputs."""
    from collections import defaultdict
    from redgnn_b200.data import _group_queries

    def naive(triples):
        triples = [list(map(int, t)) for t in triples]
        triples.sort(key=lambda x: (x[0], x[1]))
        table = defaultdict(list)
        for h, r, t in triples:
            table[(h, r)].append(t)
        return list(table.keys()), [np.array(v) for v in table.values()]

    rng = np.random.default_rng(0)
    for n, hi in ((0, 5), (1, 5), (50, 3), (500, 7), (2000, 40)):
        tri = rng.integers(0, hi, size=(n, 3))
        keys, ans = _group_queries(tri.tolist())
        keys0, ans0 = naive(tri.tolist())
        assert keys == keys0
        assert len(ans) == len(ans0) and all(np.array_equal(a, b) for a, b in zip(ans, ans0))
