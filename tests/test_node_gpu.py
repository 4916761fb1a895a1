"""GPU parity of the fused node update (rg_node_update): tensor-core (tcgen05/TMEM, 3xTF32) and
CUDA-core variants against a torch fp32/fp64 statement of models.py:41,81-86."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import assert_close

pytestmark = pytest.mark.gpu
ACTS = {0: lambda x: x, 1: torch.relu, 2: torch.tanh}


def reference(agg, h_prev, src, W_h, gru, act, Ws8, W_final):
    """double-precision statement of the node update."""
    dd = lambda t: t.double()
    x = ACTS[act](dd(agg) @ dd(W_h).t())
    h0 = torch.zeros_like(x)
    if h_prev is not None:
        ok = src >= 0
        h0[ok] = dd(h_prev)[src[ok].long()]
    gi = x @ dd(gru.weight_ih_l0).t() + dd(gru.bias_ih_l0)
    gh = h0 @ dd(gru.weight_hh_l0).t() + dd(gru.bias_hh_l0)
    i_r, i_z, i_n = gi.chunk(3, 1)
    h_r, h_z, h_n = gh.chunk(3, 1)
    r, z = torch.sigmoid(i_r + h_r), torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    hid = (1 - z) * n + z * h0
    return hid, hid @ dd(Ws8).t(), hid @ dd(W_final).reshape(-1)


@pytest.mark.parametrize("d,n,act,has_h0", [(48, 1000, 1, True), (48, 128, 2, True), (48, 77, 1, False),
                                            (32, 513, 0, True), (16, 300, 1, True), (64, 400, 2, True),
                                            (48, 20000, 1, True)])
def test_node_update_variants(d, n, act, has_h0):
    from redgnn_b200.ops import node_update
    torch.manual_seed(d + n)
    dev = "cuda"
    agg = torch.randn(n, d, device=dev) * 2
    n_prev = max(1, n // 2)
    h_prev = torch.randn(n_prev, d, device=dev) if has_h0 else None
    src = None
    if has_h0:
        src = torch.full((n,), -1, dtype=torch.int32, device=dev)
        pick = torch.randperm(n, device=dev)[:n_prev].sort()[0]
        src[pick] = torch.arange(n_prev, dtype=torch.int32, device=dev)
    W_h = torch.randn(d, d, device=dev) / d ** 0.5
    gru = torch.nn.GRU(d, d).to(dev)
    Ws8 = F.pad(torch.randn(5, d, device=dev) / d ** 0.5, (0, 0, 0, 3)).contiguous()
    W_final = torch.randn(1, d, device=dev)
    want = reference(agg, h_prev, src, W_h, gru, act, Ws8, W_final)
    outs = {}
    for variant in ("tc", "simt"):
        os.environ["REDGNN_NODE_SIMT"] = "1" if variant == "simt" else "0"
        try:
            with torch.no_grad():
                got = node_update(agg, h_prev, src, W_h, gru, act, Ws8, W_final)
        finally:
            os.environ.pop("REDGNN_NODE_SIMT", None)
        torch.cuda.synchronize()
        outs[variant] = got
        for g, w, name in zip(got, want, ("hidden", "as8", "score")):
            assert_close(g, w.float(), 2e-5, "%s %s d=%d" % (variant, name, d))
    # device-side count: only the first n_true rows are produced
    n_true = torch.tensor([n // 3], dtype=torch.int64, device=dev)
    with torch.no_grad():
        part = node_update(agg, h_prev, src, W_h, gru, act, Ws8, W_final, n_dev=n_true)
    assert torch.equal(part[0][:n // 3], outs["tc" if d <= 48 else "simt"][0][:n // 3])
