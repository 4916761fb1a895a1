"""Pure-numpy generator (no dependency on the CUDA library: the reference / CPU arm of bench.py and
the oracle-side fixture scripts import it without loading libredgnn_b200.so).

Synthetic knowledge graphs of the BASELINE shapes (SURVEY.md section 8d recipe) written in the
reference's own text layout, so both the drop-in loaders and the oracle read the same files.

Heads are drawn from a Zipf-like law p(rank) ~ rank^-alpha_h over a random entity permutation,
tails likewise (independent permutation, alpha_t), relations uniform; h == t rows are dropped and
triples de-duplicated, topped up to the exact count, then shuffled.
"""
import os

import numpy as np

SHAPES = {
    # name: n_ent, n_rel, n_triples, n_valid, n_test, alpha_h, alpha_t, n_layer
    "tiny": (300, 7, 3000, 200, 200, 0.8, 0.8, 3),
    "family": (3007, 12, 23483, 2038, 2835, 0.6, 0.6, 3),
    "fb15k237": (14541, 237, 272115, 17535, 20466, 0.8, 0.8, 4),
    "yago310": (123182, 37, 1079040, 5000, 5000, 0.3, 1.0, 5),
    "powerlaw": (1000000, 500, 10000000, 5000, 5000, 1.0, 1.0, 6),
}


def zipf_triples(n_ent, n_rel, n_triples, alpha_h, alpha_t, seed=0):
    """Returns int64 [n_triples, 3] unique (h, r, t) rows with h != t."""
    rng = np.random.default_rng(seed)
    ranks = np.arange(1, n_ent + 1, dtype=np.float64)
    ph = ranks ** (-alpha_h)
    ph /= ph.sum()
    pt = ranks ** (-alpha_t)
    pt /= pt.sum()
    perm_h, perm_t = rng.permutation(n_ent), rng.permutation(n_ent)
    ch, ct = np.cumsum(ph), np.cumsum(pt)
    keys = np.empty(0, dtype=np.int64)
    while len(keys) < n_triples:
        m = int((n_triples - len(keys)) * 1.3) + 1024
        h = perm_h[np.minimum(np.searchsorted(ch, rng.random(m)), n_ent - 1)]
        t = perm_t[np.minimum(np.searchsorted(ct, rng.random(m)), n_ent - 1)]
        r = rng.integers(0, n_rel, size=m)
        ok = h != t
        k = (h[ok].astype(np.int64) * n_rel + r[ok]) * n_ent + t[ok]
        keys = np.unique(np.concatenate([keys, k]))
    keys = rng.permutation(keys)[:n_triples]
    h, rem = keys // (n_rel * n_ent), keys % (n_rel * n_ent)
    return np.stack([h, rem // n_ent, rem % n_ent], axis=1)


def _write_triples(path, tri):
    with open(path, "w") as f:
        f.write("".join("e%d\tr%d\te%d\n" % (h, r, t) for h, r, t in tri.tolist()))


def write_transductive(task_dir, shape="tiny", seed=0, override=None):
    """entities.txt / relations.txt (one name per line) + facts/train/valid/test.txt;
    facts : train = 3 : 1 (reference README)."""
    n_ent, n_rel, n_tri, n_valid, n_test, ah, at, _ = override or SHAPES[shape]
    tri = zipf_triples(n_ent, n_rel, n_tri + n_valid + n_test, ah, at, seed)
    os.makedirs(task_dir, exist_ok=True)
    with open(os.path.join(task_dir, "entities.txt"), "w") as f:
        f.write("".join("e%d\n" % i for i in range(n_ent)))
    with open(os.path.join(task_dir, "relations.txt"), "w") as f:
        f.write("".join("r%d\n" % i for i in range(n_rel)))
    n_fact = n_tri * 3 // 4
    _write_triples(os.path.join(task_dir, "facts.txt"), tri[:n_fact])
    _write_triples(os.path.join(task_dir, "train.txt"), tri[n_fact:n_tri])
    _write_triples(os.path.join(task_dir, "valid.txt"), tri[n_tri:n_tri + n_valid])
    _write_triples(os.path.join(task_dir, "test.txt"), tri[n_tri + n_valid:])
    return task_dir


def write_inductive(task_dir, n_ent=400, n_ent_ind=260, n_rel=9, n_train=3500, n_ind_train=1500, n_eval=250,
                    seed=0):
    """<dir>/ and <dir>_ind/ with `name<TAB>id` entity / relation tables (inductive layout)."""
    for d, ne, nt, sd in ((task_dir, n_ent, n_train, seed), (task_dir + "_ind", n_ent_ind, n_ind_train, seed + 7)):
        os.makedirs(d, exist_ok=True)
        tri = zipf_triples(ne, n_rel, nt + 2 * n_eval, 0.7, 0.7, sd)
        with open(os.path.join(d, "entities.txt"), "w") as f:
            f.write("".join("e%d\t%d\n" % (i, i) for i in range(ne)))
        with open(os.path.join(d, "relations.txt"), "w") as f:
            f.write("".join("r%d\t%d\n" % (i, i) for i in range(n_rel)))
        _write_triples(os.path.join(d, "train.txt"), tri[:nt])
        _write_triples(os.path.join(d, "valid.txt"), tri[nt:nt + n_eval])
        _write_triples(os.path.join(d, "test.txt"), tri[nt + n_eval:])
    return task_dir


class Options(object):
    """The bare options object the reference's train.py builds (transductive/train.py:46-111)."""

    def __init__(self, **kw):
        self.lr, self.decay_rate, self.lamb = 0.003, 0.99, 1e-5
        self.hidden_dim, self.attn_dim, self.n_layer = 48, 5, 3
        self.dropout, self.act, self.n_batch, self.n_tbatch = 0.0, 'relu', 20, 50
        self.__dict__.update(kw)


class ArraySplits(object):
    """Splits of a LARGE synthetic KG built straight from arrays (no text files, no Python loops).
    Row order follows the transductive reference: train graph = [facts | inverse facts], test graph =
    [facts | inv | train | inv train] (self-loops are appended by the graph builders)."""

    def __init__(self, shape="powerlaw", seed=0, override=None):
        n_ent, n_rel, n_tri, n_valid, n_test, ah, at, self.n_layer = override or SHAPES[shape]
        tri = zipf_triples(n_ent, n_rel, n_tri + n_valid + n_test, ah, at, seed)
        self.n_ent, self.n_rel = n_ent, n_rel
        n_fact = n_tri * 3 // 4
        inv = lambda t: np.stack([t[:, 2], t[:, 1] + n_rel, t[:, 0]], axis=1)
        dbl = lambda t: np.concatenate([t, inv(t)], axis=0)
        fact, train, test = tri[:n_fact], tri[n_fact:n_tri], tri[n_tri + n_valid:]
        self.train_data = dbl(train)
        self.train_graph_triples = dbl(fact)
        self.test_graph_triples = np.concatenate([dbl(fact), dbl(train)], axis=0)
        q = np.unique(dbl(test)[:, :2], axis=0)
        self.test_q = [tuple(x) for x in q.tolist()]
