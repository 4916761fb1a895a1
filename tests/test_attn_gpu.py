"""GPU parity of the per-relation / per-query kernels (rg_attn_tables, rg_attn_param_grads) with a torch fp64
statement of the attention-side of GNNLayer (reference Static/transductive/models.py:29-36) and its autograd."""
import numpy as np
import pytest
import torch

from helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d,a,rows,n,copies", [(48, 5, 475, 16, 8), (32, 3, 25, 7, 1), (48, 8, 1001, 64, 8), (16, 1, 15, 1, 2)])
def test_attn_tables_and_param_grads(d, a, rows, n, copies):
    from redgnn_b200._lib import lib, check, ptr, stream_ptr
    torch.manual_seed(d + rows)
    dev = "cuda"
    f = lambda *s: torch.randn(*s, device=dev)
    rela, Wr, Wqr, bqr, w_alpha = f(rows, d), f(a, d), f(a, d), f(a), f(1, a)
    q_rel = torch.randint(0, rows, (n,), device=dev)
    q_rel[n // 2:] = int(q_rel[0])                             # repeated query relations
    ar8, aq8, w8 = (torch.full(s, 9.0, device=dev) for s in ((rows, 8), (n, 8), (8,)))
    check(lib.rg_attn_tables(d, a, rows, n, ptr(rela), ptr(Wr), ptr(Wqr), ptr(bqr), ptr(w_alpha), ptr(q_rel), ptr(ar8),
                             ptr(aq8), ptr(w8), stream_ptr()))
    dd = lambda t: t.detach().double().requires_grad_(True)
    rela64, Wr64, Wqr64, bqr64 = dd(rela), dd(Wr), dd(Wqr), dd(bqr)
    want_ar = rela64 @ Wr64.t()
    want_aq = rela64[q_rel] @ Wqr64.t() + bqr64
    assert_close(ar8[:, :a], want_ar.float(), 1e-5, "ar8")
    assert_close(aq8[:, :a], want_aq.float(), 1e-5, "aq8")
    assert float(ar8[:, a:].abs().max() if a < 8 else 0) == 0 and float(aq8[:, a:].abs().max() if a < 8 else 0) == 0
    assert torch.equal(w8[:a], w_alpha.reshape(-1)) and float(w8[a:].abs().sum()) == 0
    # upstream gradients as the edge backward / query sums deliver them
    g_rela_c, g_ar8_c = f(copies, rows, d), torch.zeros(copies, rows, 8, device=dev)
    g_ar8_c[:, :, :a] = f(copies, rows, a)
    q_slices = 32
    q_part = f(n, q_slices, 24)
    q_part[:, :, a:8] = 0
    outs = {k: torch.full(s, 7.0, device=dev) for k, s in (("rela", (rows, d)), ("Wr", (a, d)), ("Wqr", (a, d)), ("bqr", (a,)),
                                                           ("w", (a,)), ("b", (1,)))}
    check(lib.rg_attn_param_grads(d, a, rows, n, copies, ptr(rela), ptr(Wr), ptr(Wqr), ptr(q_rel), ptr(g_rela_c), ptr(g_ar8_c),
                                  ptr(q_part), q_slices, ptr(outs["rela"]), ptr(outs["Wr"]), ptr(outs["Wqr"]), ptr(outs["bqr"]),
                                  ptr(outs["w"]), ptr(outs["b"]), stream_ptr()))
    g_ar = g_ar8_c.double().sum(0)[:, :a]
    g_aq = q_part.double().sum(1)[:, :a]
    (want_ar * g_ar).sum().backward(retain_graph=True)
    (want_aq * g_aq).sum().backward()
    assert_close(outs["rela"], (rela64.grad + g_rela_c.double().sum(0)).float(), 2e-5, "rela_embed grad")
    assert_close(outs["Wr"], Wr64.grad.float(), 2e-5, "Wr grad")
    assert_close(outs["Wqr"], Wqr64.grad.float(), 2e-5, "Wqr grad")
    assert_close(outs["bqr"], bqr64.grad.float(), 2e-5, "bqr grad")
    assert_close(outs["w"], q_part.double().sum((0, 1))[8:8 + a].float(), 2e-5, "w_alpha grad")
    assert_close(outs["b"], q_part.double().sum((0, 1))[16:17].float(), 2e-5, "b_alpha grad")
    again = torch.empty_like(outs["rela"])
    check(lib.rg_attn_param_grads(d, a, rows, n, copies, ptr(rela), ptr(Wr), ptr(Wqr), ptr(q_rel), ptr(g_rela_c), ptr(g_ar8_c),
                                  ptr(q_part), q_slices, ptr(again), ptr(outs["Wr"]), ptr(outs["Wqr"]), ptr(outs["bqr"]),
                                  ptr(outs["w"]), ptr(outs["b"]), stream_ptr()))
    assert torch.equal(again, outs["rela"]), "fixed summation order: bit-reproducible"
