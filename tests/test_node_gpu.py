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


@pytest.mark.parametrize("d,n,act,has_h0,drop,fold", [(48, 1000, 1, True, False, True), (48, 130, 2, True, True, False),
                                                      (48, 77, 1, False, False, True), (32, 513, 0, True, True, True),
                                                      (16, 300, 1, True, False, False), (48, 20000, 1, True, False, True)])
def test_node_backward_native_kernels_vs_fp64_autograd(d, n, act, has_h0, drop, fold):
    """rg_node_update_train -> rg_node_bwd (tcgen05, A operand in TMEM) + rg_node_wgrad (CUDA cores) against
    torch.autograd of the fp64 statement: g_agg, g_h0, every weight / bias gradient; `fold` adds the
    g_small . w_small and g_h0_next[remap] upstream parts and the dW_small output."""
    import ctypes as C
    from redgnn_b200 import _lib
    from redgnn_b200._lib import lib, check, ptr, stream_ptr
    torch.manual_seed(7 * d + n)
    dev = "cuda"
    f = lambda *s: torch.randn(*s, device=dev)
    agg = f(n, d) * 2
    n_prev = max(1, n // 2)
    h_prev = f(n_prev, d) if has_h0 else None
    src = None
    if has_h0:
        src = torch.full((n,), -1, dtype=torch.int32, device=dev)
        pick = torch.randperm(n, device=dev)[:n_prev].sort()[0]
        src[pick] = torch.arange(n_prev, dtype=torch.int32, device=dev)
    W_h = f(d, d) / d ** 0.5
    gru = torch.nn.GRU(d, d).to(dev)
    mask = ((torch.rand(n, d, device=dev) < 0.7).float() / 0.7) if drop else None
    g_hidden = f(n, d)
    g_small = (f(n, 8) if fold else None)
    w_small = (f(8, d) if fold else None)
    n_next = n + 50
    g_h0_next = f(n_next, d) if fold else None
    remap = torch.randperm(n_next, device=dev)[:n].sort()[0].to(torch.int32) if fold else None
    # forward through the training kernel (saves the gates)
    hidden = torch.empty(n, d, device=dev)
    pl = _lib.il_plane_floats(n, d)               # saved / G4 / g_pre: lane-interleaved planes (csrc/rg_tc.cuh)
    saved = torch.empty(6, pl, device=dev)
    check(lib.rg_node_update_train(d, n, None, ptr(agg), ptr(h_prev), ptr(src), ptr(W_h), ptr(gru.weight_ih_l0),
                                   ptr(gru.weight_hh_l0), ptr(gru.bias_ih_l0), ptr(gru.bias_hh_l0), act, ptr(mask),
                                   ptr(hidden), ptr(saved), None, 0, None, None, None, stream_ptr()))
    e = lambda *s: torch.empty(*s, device=dev)
    G4, g_pre, g_agg, g_h0 = e(4, pl), e(pl), e(n, d), (e(n, d) if has_h0 else None)
    check(lib.rg_node_bwd(d, n, None, ptr(g_hidden), ptr(g_small), 8, ptr(w_small), 8, ptr(g_h0_next), ptr(remap), ptr(saved), n,
                          ptr(mask), ptr(W_h), ptr(gru.weight_ih_l0), ptr(gru.weight_hh_l0), act, int(has_h0), ptr(G4),
                          ptr(g_pre), ptr(g_agg), ptr(g_h0), stream_ptr()))
    out_floats = int(lib.rg_node_wgrad_out_floats(d))
    partial, wg = e(int(lib.rg_node_wgrad_ctas()) * out_floats), e(out_floats)
    check(lib.rg_node_wgrad(d, n, None, ptr(saved), n, ptr(mask), ptr(agg), ptr(hidden), ptr(G4), ptr(g_pre), ptr(g_small), 8,
                            int(has_h0), ptr(partial), ptr(wg), None, None, None, None, None, None, 0, stream_ptr()))
    torch.cuda.synchronize()
    # fp64 autograd of the same op
    dd = lambda t: t.detach().double().requires_grad_(True)
    a64, Wh64, wih, whh, bih, bhh = dd(agg), dd(W_h), dd(gru.weight_ih_l0), dd(gru.weight_hh_l0), dd(gru.bias_ih_l0), dd(gru.bias_hh_l0)
    h064 = torch.zeros(n, d, dtype=torch.float64, device=dev)
    if has_h0:
        ok = src >= 0
        h064[ok] = h_prev.double()[src[ok].long()]
    h064.requires_grad_(True)
    x = ACTS[act](a64 @ Wh64.t())
    xin = x * mask.double() if drop else x
    gi, gh = xin @ wih.t() + bih, h064 @ whh.t() + bhh
    i_r, i_z, i_n = gi.chunk(3, 1)
    h_r, h_z, h_n = gh.chunk(3, 1)
    r, z = torch.sigmoid(i_r + h_r), torch.sigmoid(i_z + h_z)
    hid = (1 - z) * torch.tanh(i_n + r * h_n) + z * h064
    assert_close(hidden, hid.detach().float(), 2e-5, "forward hidden")
    up = g_hidden.double()
    if fold:
        up = up + g_small.double() @ w_small.double() + g_h0_next.double()[remap.long()]
    hid.backward(up)
    o_whh, o_wh, o_ws, o_b = 3 * d * d, 6 * d * d, 7 * d * d, 7 * d * d + 8 * d
    tol = 5e-5
    assert_close(g_agg, a64.grad.float(), tol, "g_agg")
    assert_close(wg[:o_whh].view(3 * d, d), wih.grad.float(), tol, "dW_ih")
    assert_close(wg[o_wh:o_ws].view(d, d), Wh64.grad.float(), tol, "dW_h")
    b4 = wg[o_b:].view(4, d)
    assert_close(b4[:3].reshape(-1), bih.grad.float(), tol, "db_ih")
    assert_close(torch.cat([b4[0], b4[1], b4[3]]), bhh.grad.float(), tol, "db_hh")
    if has_h0:
        assert_close(g_h0, h064.grad.float(), tol, "g_h0")
        assert_close(wg[o_whh:o_wh].view(3 * d, d), whh.grad.float(), tol, "dW_hh")
    else:
        assert float(wg[o_whh:o_wh].abs().max()) == 0.0
    if fold:
        assert_close(wg[o_ws:o_b].view(8, d), (g_small.double().t() @ hid.detach()).float(), tol, "dW_small")
    # device-side count: rows past n_true are neither read nor written, weight gradients cover n_true rows only
    n_true = torch.tensor([n // 3], dtype=torch.int64, device=dev)
    g_agg2 = torch.full_like(g_agg, 7.0)
    check(lib.rg_node_bwd(d, n, ptr(n_true), ptr(g_hidden), ptr(g_small), 8, ptr(w_small), 8, ptr(g_h0_next), ptr(remap), ptr(saved),
                          n, ptr(mask), ptr(W_h), ptr(gru.weight_ih_l0), ptr(gru.weight_hh_l0), act, int(has_h0), ptr(G4),
                          ptr(g_pre), ptr(g_agg2), ptr(g_h0), stream_ptr()))
    assert torch.equal(g_agg2[:n // 3], g_agg[:n // 3]) and float((g_agg2[n // 3:] - 7.0).abs().max()) == 0.0
