"""CPU oracle for RED-GNN's query-conditioned propagation path.

TEST INFRASTRUCTURE ONLY.  This file is a CPU restatement (scipy / numpy / torch-CPU) of the
reference algorithm.  It is imported only by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`; the product package `redgnn_b200`
never imports it and has no CPU fallback.

Parity status: PINNED.  The reference ships no tests or golden vectors, so the
oracle is pinned against the reference itself, imported unmodified in the build container
(`oracle/ref_import.py`): `tests/test_oracle_vs_reference.py` compares every function below with
the live reference where `/root/reference` is mounted, and `tests/test_oracle_golden.py` compares
it with the committed fixtures `tests/golden/*.npz` that `oracle/make_golden.py` generated from
the live reference (those run everywhere).

All `file:line` citations are relative to /root/reference/Static/.

Third-party arithmetic restated here:
  * torch_scatter 2.0.9 `scatter(..., reduce='sum')` (transductive/models.py:3,39) == `index_add_`.
  * scipy.sparse `csr_matrix.dot` (load_data.py:115-116): the edge ORDER depends on the SMMP
    output order of `csr_matmat` -- pinned empirically (fact row ascending, batch index
    descending) by `expand_definition()` below and the SURVEY section 4 hashes.
"""
from collections import defaultdict
import os

import numpy as np
import torch
from scipy.sparse import csr_matrix
from scipy.stats import rankdata


# --------------------------------------------------------------------------------------
# a1. graph construction  (transductive/load_data.py:69-89, inductive/load_data.py:76-98)
# --------------------------------------------------------------------------------------
def add_inverse_block(triples, n_rel):
    """transductive `double_triple` (load_data.py:69-74): inverses appended as ONE block."""
    triples = [list(t) for t in triples]
    return triples + [[t, r + n_rel, h] for h, r, t in triples]


def add_inverse_interleaved(triples, n_rel):
    """inductive `read_triples` (inductive/load_data.py:84-85): (h,r,t),(t,r+R,h) per line."""
    out = []
    for h, r, t in triples:
        out.append([h, r, t])
        out.append([t, r + n_rel, h])
    return out


class Graph(object):
    """KG rows [triples ; self-loops (e, 2R, e)] + one-hot CSR of the head column.

    The reference keeps KG as float64 because the identity block is built with np.ones
    (load_data.py:77-79); values are exact small integers, so the oracle keeps that dtype to
    reproduce the reference's fp64->int64 conversion cost in the CPU baseline.
    """

    def __init__(self, triples, n_ent, n_rel):
        ids = np.arange(n_ent, dtype=np.float64)[:, None]
        loops = np.concatenate([ids, np.full((n_ent, 1), 2.0 * n_rel), ids], axis=1)
        tri = np.asarray(triples, dtype=np.float64).reshape(-1, 3)
        self.KG = np.concatenate([tri, loops], axis=0)
        self.n_fact = self.KG.shape[0]
        self.n_ent = n_ent
        self.n_rel = n_rel
        rows = np.arange(self.n_fact)
        self.M_sub = csr_matrix((np.ones(self.n_fact), (rows, self.KG[:, 0])),
                                shape=(self.n_fact, n_ent))

    def int_arrays(self):
        kg = self.KG.astype(np.int64)
        return kg[:, 0].copy(), kg[:, 1].copy(), kg[:, 2].copy()


# --------------------------------------------------------------------------------------
# a2. one hop of frontier expansion  (transductive/load_data.py:106-131,
#                                     inductive/load_data.py:115-143)
# --------------------------------------------------------------------------------------
def get_neighbors(graph, nodes):
    """Port of the reference algorithm with the same library calls (the CPU baseline).

    nodes: np.ndarray[N,2] int (batch_idx, entity).  Returns torch int64 CPU tensors
    (tail_nodes[N',2], sampled_edges[E,6], old_nodes_new_idx[N]).
    """
    nodes = np.asarray(nodes)
    sel = csr_matrix((np.ones(len(nodes)), (nodes[:, 1], nodes[:, 0])),
                     shape=(graph.n_ent, nodes.shape[0]))               # load_data.py:115
    hit = graph.M_sub.dot(sel)                                          # :116  (n_fact x N)
    fact_id, batch_id = np.nonzero(hit)                                 # :117
    rows = np.concatenate([batch_id[:, None], graph.KG[fact_id]], axis=1)   # :118
    edges = torch.LongTensor(rows)                                      # :119

    head_nodes, head_index = torch.unique(edges[:, [0, 1]], dim=0, sorted=True, return_inverse=True)
    tail_nodes, tail_index = torch.unique(edges[:, [0, 3]], dim=0, sorted=True, return_inverse=True)
    edges = torch.cat([edges, head_index[:, None], tail_index[:, None]], dim=1)   # :125

    loops = edges[:, 2] == 2 * graph.n_rel                              # :127
    order = head_index[loops].sort()[1]                                 # :128
    old_nodes_new_idx = tail_index[loops][order]                        # :129
    return tail_nodes, edges, old_nodes_new_idx


def expand_definition(head, rel, tail, n_ent, nodes):
    """Library-independent DEFINITION of the same hop (numpy only), used to pin the semantics
    the CUDA kernels implement:

      * frontier = unique (b, e) rows of `nodes`, sorted lexicographically;
      * edge set = {(f, b) : (b, head[f]) in frontier}, ordered by f ASCENDING then b DESCENDING
        (scipy SMMP emits each output row's columns in reverse insertion order);
      * tail_nodes = sorted unique (b, tail[f]); head_index / tail_index = row ranks;
      * old_nodes_new_idx[i] = rank of frontier row i inside tail_nodes.

    head/rel/tail: int arrays [F] in reference row order (self-loops last).
    Returns numpy int64 arrays (tail_nodes, edges, old_nodes_new_idx).
    """
    head = np.asarray(head, dtype=np.int64)
    rel = np.asarray(rel, dtype=np.int64)
    tail = np.asarray(tail, dtype=np.int64)
    nodes = np.asarray(nodes, dtype=np.int64).reshape(-1, 2)
    key_in = np.unique(nodes[:, 0] * n_ent + nodes[:, 1])               # sorted unique frontier
    fb, fe = key_in // n_ent, key_in % n_ent
    # facts grouped by head (stable) so that each frontier node lists its facts
    order = np.argsort(head, kind="stable")
    ptr = np.zeros(n_ent + 1, dtype=np.int64)
    np.add.at(ptr, head + 1, 1)
    ptr = np.cumsum(ptr)
    deg = ptr[fe + 1] - ptr[fe]
    e_b = np.repeat(fb, deg)
    starts = np.repeat(ptr[fe], deg)
    within = np.arange(deg.sum()) - np.repeat(np.cumsum(deg) - deg, deg)
    e_f = order[starts + within]
    # reference order: fact ascending, batch descending
    perm = np.lexsort((-e_b, e_f))
    e_b, e_f = e_b[perm], e_f[perm]
    e_h, e_r, e_t = head[e_f], rel[e_f], tail[e_f]
    key_out = np.unique(e_b * n_ent + e_t)
    head_index = np.searchsorted(key_in, e_b * n_ent + e_h)
    tail_index = np.searchsorted(key_out, e_b * n_ent + e_t)
    edges = np.stack([e_b, e_h, e_r, e_t, head_index, tail_index], axis=1)
    tail_nodes = np.stack([key_out // n_ent, key_out % n_ent], axis=1)
    old_new = np.searchsorted(key_out, key_in)
    return tail_nodes, edges, old_new


# --------------------------------------------------------------------------------------
# a3/a4. GNNLayer  (transductive/models.py:5-43 == inductive/models.py:5-43)
# --------------------------------------------------------------------------------------
ACTS = {"relu": torch.relu, "tanh": torch.tanh, "idd": lambda x: x}


def layer_param_names(i):
    p = "gnn_layers.%d." % i
    return dict(rela=p + "rela_embed.weight", Ws=p + "Ws_attn.weight", Wr=p + "Wr_attn.weight",
                Wqr=p + "Wqr_attn.weight", bqr=p + "Wqr_attn.bias", wa=p + "w_alpha.weight",
                ba=p + "w_alpha.bias", Wh=p + "W_h.weight")


def gnn_layer_forward(sd, i, q_rel, hidden, edges, n_node, act):
    """models.py:23-43 with parameters taken from a state_dict `sd` (layer i)."""
    nm = layer_param_names(i)
    rela = sd[nm["rela"]]
    sub, rel, obj, r_idx = edges[:, 4], edges[:, 2], edges[:, 5], edges[:, 0]
    hs = hidden[sub]                                                     # :29
    hr = rela[rel]                                                       # :30
    h_qr = rela[q_rel][r_idx]                                            # :33
    pre = hs @ sd[nm["Ws"]].t() + hr @ sd[nm["Wr"]].t() + h_qr @ sd[nm["Wqr"]].t() + sd[nm["bqr"]]
    alpha = torch.sigmoid(torch.relu(pre) @ sd[nm["wa"]].t() + sd[nm["ba"]])     # :36
    message = alpha * (hs + hr)                                          # :35,37
    agg = torch.zeros(n_node, hidden.shape[1], dtype=hidden.dtype).index_add_(0, obj, message)  # :39
    return ACTS[act](agg @ sd[nm["Wh"]].t())                             # :41


def gru_step(sd, x, h):
    """Single-step nn.GRU(d, d) (models.py:63,83): gate order (r, z, n)."""
    gi = x @ sd["gate.weight_ih_l0"].t() + sd["gate.bias_ih_l0"]
    gh = h @ sd["gate.weight_hh_l0"].t() + sd["gate.bias_hh_l0"]
    i_r, i_z, i_n = gi.chunk(3, dim=1)
    h_r, h_z, h_n = gh.chunk(3, dim=1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1.0 - z) * n + z * h


def model_forward(sd, graph, subs, rels, n_layer, act, n_ent_out=None, dropout_masks=None,
                  return_trace=False):
    """RED_GNN_trans / RED_GNN_induc.forward (models.py:65-89), dropout off unless explicit
    per-layer keep-masks (already scaled by 1/(1-p)) are supplied.

    sd: state_dict (fp32 CPU tensors) with the reference parameter names.
    Returns scores_all (n, n_ent_out) and optionally the per-layer (nodes, edges, remap) trace.
    """
    n = len(subs)
    d = sd["W_final.weight"].shape[1]
    dt = sd["W_final.weight"].dtype          # fp32 like the reference; fp64 for error yardsticks in tests
    q_sub = torch.as_tensor(np.asarray(subs), dtype=torch.long)
    q_rel = torch.as_tensor(np.asarray(rels), dtype=torch.long)
    h0 = torch.zeros(n, d, dtype=dt)
    nodes = torch.stack([torch.arange(n), q_sub], dim=1)                 # :73
    hidden = torch.zeros(n, d, dtype=dt)
    trace = []
    for i in range(n_layer):
        nodes, edges, remap = get_neighbors(graph, nodes.numpy())        # :78
        hidden = gnn_layer_forward(sd, i, q_rel, hidden, edges, nodes.shape[0], act)   # :80
        h0 = torch.zeros(nodes.shape[0], d, dtype=dt).index_copy_(0, remap, h0)    # :81
        if dropout_masks is not None:
            hidden = hidden * dropout_masks[i]                           # :82
        hidden = gru_step(sd, hidden, h0)                                # :83
        h0 = hidden
        if return_trace:
            trace.append((nodes, edges, remap))
    scores = (hidden @ sd["W_final.weight"].t()).squeeze(-1)             # :86
    n_ent_out = graph.n_ent if n_ent_out is None else n_ent_out
    scores_all = torch.zeros(n, n_ent_out, dtype=dt)                     # :87
    scores_all[nodes[:, 0], nodes[:, 1]] = scores                        # :88
    if return_trace:
        return scores_all, trace
    return scores_all


def train_loss(scores, objs):
    """transductive/base_model.py:58-60 (sum over the batch, not mean)."""
    pos = scores[torch.arange(len(scores)), torch.as_tensor(objs, dtype=torch.long)]
    mx = scores.max(dim=1, keepdim=True)[0]
    return torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(scores - mx), dim=1)))


def init_state_dict(n_layer, hidden_dim, attn_dim, n_rel, seed=1234):
    """Random-init parameters with the reference names/shapes (models.py:6-21,46-63), built
    from torch.nn modules so that default initialisers match the reference's."""
    g = torch.Generator().manual_seed(seed)
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        import torch.nn as nn
        sd = {}
        for i in range(n_layer):
            nm = layer_param_names(i)
            sd[nm["rela"]] = nn.Embedding(2 * n_rel + 1, hidden_dim).weight.detach().clone()
            sd[nm["Ws"]] = nn.Linear(hidden_dim, attn_dim, bias=False).weight.detach().clone()
            sd[nm["Wr"]] = nn.Linear(hidden_dim, attn_dim, bias=False).weight.detach().clone()
            lin = nn.Linear(hidden_dim, attn_dim)
            sd[nm["Wqr"]], sd[nm["bqr"]] = lin.weight.detach().clone(), lin.bias.detach().clone()
            lin = nn.Linear(attn_dim, 1)
            sd[nm["wa"]], sd[nm["ba"]] = lin.weight.detach().clone(), lin.bias.detach().clone()
            sd[nm["Wh"]] = nn.Linear(hidden_dim, hidden_dim, bias=False).weight.detach().clone()
        sd["W_final.weight"] = nn.Linear(hidden_dim, 1, bias=False).weight.detach().clone()
        gru = nn.GRU(hidden_dim, hidden_dim)
        for k, v in gru.state_dict().items():
            sd["gate." + k] = v.detach().clone()
    finally:
        torch.random.set_rng_state(state)
    del g
    return sd


# --------------------------------------------------------------------------------------
# filtered ranking metrics  (transductive/utils.py:7-21)
# --------------------------------------------------------------------------------------
def cal_ranks(scores, labels, filters):
    scores = scores - np.min(scores, axis=1, keepdims=True) + 1e-8
    full = rankdata(-scores, method="average", axis=1)
    filt = rankdata(-(scores * filters), method="min", axis=1)
    ranks = (full - filt + 1) * labels
    return list(ranks[np.nonzero(ranks)])


def cal_performance(ranks):
    ranks = np.asarray(ranks, dtype=np.float64)
    return (1.0 / ranks).mean(), float((ranks <= 1).mean()), float((ranks <= 10).mean())


# --------------------------------------------------------------------------------------
# dataset readers  (transductive/load_data.py:8-67,91-104; inductive/load_data.py:8-113)
# --------------------------------------------------------------------------------------
def _read_names(path, with_id):
    table = {}
    with open(path) as f:
        for k, line in enumerate(f):
            if with_id:
                name, idx = line.strip().split()
                table[name] = int(idx)
            else:
                table[line.strip()] = k
    return table


def _read_triples(path, ent, rel):
    out = []
    with open(path) as f:
        for line in f:
            h, r, t = line.strip().split()
            out.append([ent[h], rel[r], ent[t]])
    return out


def group_queries(triples):
    """load_query (load_data.py:91-104): sort by (h, r); group answers per (h, r)."""
    table = defaultdict(list)
    for h, r, t in sorted(triples, key=lambda x: (x[0], x[1])):
        table[(h, r)].append(t)
    return list(table.keys()), [np.array(v) for v in table.values()]


class TransductiveData(object):
    """Oracle-side view of transductive/load_data.py DataLoader (initial split, no shuffle)."""

    def __init__(self, task_dir):
        ent = _read_names(os.path.join(task_dir, "entities.txt"), False)
        rel = _read_names(os.path.join(task_dir, "relations.txt"), False)
        self.n_ent, self.n_rel = len(ent), len(rel)
        R = self.n_rel
        facts = _read_triples(os.path.join(task_dir, "facts.txt"), ent, rel)
        train = _read_triples(os.path.join(task_dir, "train.txt"), ent, rel)
        valid = _read_triples(os.path.join(task_dir, "valid.txt"), ent, rel)
        test = _read_triples(os.path.join(task_dir, "test.txt"), ent, rel)
        self.filters = defaultdict(set)
        for h, r, t in facts + train + valid + test:                    # load_data.py:65-66
            self.filters[(h, r)].add(t)
            self.filters[(t, r + R)].add(h)
        self.train_data = np.array(add_inverse_block(train, R))
        self.graph = Graph(add_inverse_block(facts, R), self.n_ent, R)                 # :43
        self.test_graph = Graph(add_inverse_block(facts, R) + add_inverse_block(train, R),
                                self.n_ent, R)                                          # :44
        self.valid_q, self.valid_a = group_queries(add_inverse_block(valid, R))
        self.test_q, self.test_a = group_queries(add_inverse_block(test, R))

    def graph_for(self, mode):
        return self.graph if mode == "train" else self.test_graph       # :107-112


class InductiveData(object):
    """Oracle-side view of inductive/load_data.py DataLoader."""

    def __init__(self, task_dir):
        ind_dir = task_dir + "_ind"
        ent = _read_names(os.path.join(task_dir, "entities.txt"), True)
        rel = _read_names(os.path.join(task_dir, "relations.txt"), True)
        ent_ind = _read_names(os.path.join(ind_dir, "entities.txt"), True)
        self.n_ent, self.n_rel, self.n_ent_ind = len(ent), len(rel), len(ent_ind)
        R = self.n_rel
        rd = lambda d, f, e: add_inverse_interleaved(_read_triples(os.path.join(d, f), e, rel), R)
        tra_train, tra_valid, tra_test = (rd(task_dir, f, ent) for f in ("train.txt", "valid.txt", "test.txt"))
        ind_train, ind_valid, ind_test = (rd(ind_dir, f, ent_ind) for f in ("train.txt", "valid.txt", "test.txt"))
        self.val_filters, self.tst_filters = defaultdict(set), defaultdict(set)
        for h, r, t in tra_train + tra_valid + tra_test:                # inductive/load_data.py:177-186
            self.val_filters[(h, r)].add(t)
        for h, r, t in ind_train + ind_valid + ind_test:                # :187-197
            self.tst_filters[(h, r)].add(t)
        self.tra_graph = Graph(tra_train, self.n_ent, R)                 # :56
        self.ind_graph = Graph(ind_train, self.n_ent_ind, R)             # :57
        self.train_data = np.array(tra_valid)                            # :60
        self.valid_q, self.valid_a = group_queries(tra_test)             # :61,65
        q1, a1 = group_queries(ind_valid)
        q2, a2 = group_queries(ind_test)
        self.test_q, self.test_a = q1 + q2, a1 + a2                      # :66

    def graph_for(self, mode):
        return self.tra_graph if mode == "transductive" else self.ind_graph   # :118-125
