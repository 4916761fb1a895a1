"""Generate tests/golden/*.npz from the LIVE, UNMODIFIED reference (run in the build container,
where /root/reference is mounted):   python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  Each fixture holds the inputs (graph triples as data, queries, a
seeded state_dict) and what the reference itself produced for them: per-layer get_neighbors
outputs (full arrays for a small batch, SHA-256 for the SURVEY section 4 batches), RED_GNN_*.forward
scores, parameter gradients of the training loss, and filtered ranks.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def expansion_trace(loader, nodes, mode, n_layer, keep_arrays):
    out = {}
    for l in range(n_layer):
        tn, ed, rm = loader.get_neighbors(nodes, mode)
        out["L%d_sizes" % l] = np.array([len(nodes), len(ed), len(tn)], dtype=np.int64)
        out["L%d_sha" % l] = np.array([sha(tn), sha(ed), sha(rm)])
        if keep_arrays:
            out["L%d_nodes" % l] = tn.numpy().astype(np.int32)
            out["L%d_edges" % l] = ed.numpy().astype(np.int32)
            out["L%d_remap" % l] = rm.numpy().astype(np.int32)
        nodes = tn.numpy()
    return out


def pack(prefix, d):
    return {prefix + k: v for k, v in d.items()}


def model_block(model, loader, subs, rels, objs_idx, fwd_mode, filters, labels):
    """scores (eval mode), grads of the reference training loss (dropout inactive in eval mode),
    filtered ranks via the reference's cal_ranks."""
    out = {}
    model.eval()
    scores = model(subs, rels, fwd_mode) if fwd_mode is not None else model(subs, rels)
    out["scores"] = scores.detach().numpy().astype(np.float32)
    # base_model.py:58-60
    pos = scores[torch.arange(len(scores)), torch.LongTensor(objs_idx)]
    mx = torch.max(scores, 1, keepdim=True)[0]
    loss = torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(scores - mx), 1)))
    model.zero_grad()
    loss.backward()
    out["loss"] = np.array(loss.item(), dtype=np.float64)
    for k, p in model.named_parameters():
        out["grad." + k] = p.grad.detach().numpy().astype(np.float32)
    return out


def family():
    ld, M, U = R.load_reference("transductive")
    L = R.make_loader("transductive", R.data_dir("transductive", "family"))
    fx = {"n_ent": np.int64(L.n_ent), "n_rel": np.int64(L.n_rel),
          "fact_triple": np.array(L.fact_triple, dtype=np.int16),
          "train_triple": np.array(L.train_triple, dtype=np.int16)}
    # SURVEY section 4 batches (hash only)
    nodes = np.stack([np.arange(20), L.train_data[:20, 0]], 1)
    fx.update(pack("train20_", expansion_trace(L, nodes, "train", 3, False)))
    fx["train20_subs"] = L.train_data[:20, 0].astype(np.int64)
    tq = np.array(L.test_q[:50])
    fx.update(pack("test50_", expansion_trace(L, np.stack([np.arange(50), tq[:, 0]], 1), "test", 3, False)))
    fx["test50_subs"] = tq[:, 0].astype(np.int64)
    # small batch, full arrays
    fx.update(pack("test4_", expansion_trace(L, np.stack([np.arange(4), tq[:4, 0]], 1), "test", 3, True)))
    # model: eval scores for 16 test queries + training-loss grads for 8 train triples
    opts = R.family_options(L)
    torch.manual_seed(1234)
    model = M.RED_GNN_trans(opts, L)
    for k, v in model.state_dict().items():
        fx["sd." + k] = v.numpy().astype(np.float32)
    q = np.array(L.test_q[:16])
    subs, rels, objs = L.get_batch(np.arange(16), data="test")
    model.eval()
    scores = model(subs, rels, mode="test").detach().numpy()
    fx["eval_subs"], fx["eval_rels"] = subs.astype(np.int64), rels.astype(np.int64)
    fx["eval_scores"] = scores.astype(np.float32)
    fx["eval_objs"] = objs.astype(np.uint8)
    filt = np.zeros((16, L.n_ent))
    for i in range(16):
        filt[i][np.array(L.filters[(subs[i], rels[i])])] = 1
    fx["eval_filters"] = filt.astype(np.uint8)
    fx["eval_ranks"] = np.array(U.cal_ranks(scores, objs, filt), dtype=np.float64)
    tri = L.train_data[:8]
    blk = model_block(model, L, tri[:, 0], tri[:, 1], tri[:, 2], None, None, None)
    fx["train_triples"] = tri.astype(np.int64)
    fx.update(pack("train_", blk))
    np.savez_compressed(os.path.join(OUT, "family.npz"), **fx)
    print("family.npz", os.path.getsize(os.path.join(OUT, "family.npz")))


def fb237_v2():
    ld, M, U = R.load_reference("inductive")
    L = R.make_loader("inductive", R.data_dir("inductive", "fb237_v2"))
    fx = {"n_ent": np.int64(L.n_ent), "n_ent_ind": np.int64(L.n_ent_ind), "n_rel": np.int64(L.n_rel),
          "tra_triples": L.tra_KG[:-L.n_ent].astype(np.int16),       # interleaved inverses, no self-loops
          "ind_triples": L.ind_KG[:-L.n_ent_ind].astype(np.int16)}
    tr = L.tra_train[:10]
    fx.update(pack("tra10_", expansion_trace(L, np.stack([np.arange(10), tr[:, 0]], 1), "transductive", 3, False)))
    fx["tra10_subs"] = tr[:, 0].astype(np.int64)
    tq = np.array(L.test_q[:10])
    fx.update(pack("ind10_", expansion_trace(L, np.stack([np.arange(10), tq[:, 0]], 1), "inductive", 3, False)))
    fx["ind10_subs"] = tq[:, 0].astype(np.int64)
    fx.update(pack("ind3_", expansion_trace(L, np.stack([np.arange(3), tq[:3, 0]], 1), "inductive", 3, True)))
    opts = R.fb237_v2_options(L)
    torch.manual_seed(1234)
    model = M.RED_GNN_induc(opts, L)
    for k, v in model.state_dict().items():
        fx["sd." + k] = v.numpy().astype(np.float32)
    subs, rels, objs = L.get_batch(np.arange(10), data="test")
    model.eval()
    scores = model(subs, rels, "inductive").detach().numpy()
    fx["eval_subs"], fx["eval_rels"] = subs.astype(np.int64), rels.astype(np.int64)
    fx["eval_scores"] = scores.astype(np.float32)
    fx["eval_objs"] = objs.astype(np.uint8)
    filt = np.zeros((10, L.n_ent_ind))
    for i in range(10):
        filt[i][np.array(L.tst_filters[(subs[i], rels[i])])] = 1
    fx["eval_filters"] = filt.astype(np.uint8)
    fx["eval_ranks"] = np.array(U.cal_ranks(scores, objs, filt), dtype=np.float64)
    tri = L.tra_train[:6]
    blk = model_block(model, L, tri[:, 0], tri[:, 1], tri[:, 2], None, None, None)
    fx["train_triples"] = tri.astype(np.int64)
    fx.update(pack("train_", blk))
    np.savez_compressed(os.path.join(OUT, "fb237_v2.npz"), **fx)
    print("fb237_v2.npz", os.path.getsize(os.path.join(OUT, "fb237_v2.npz")))


if __name__ == "__main__":
    if not R.available():
        raise SystemExit("reference tree not mounted; fixtures can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    family()
    fb237_v2()
