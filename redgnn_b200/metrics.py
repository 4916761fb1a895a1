"""Device-side evaluation metrics: the filtered ranking of
`utils.cal_ranks` / `cal_performance` (reference Static/transductive/utils.py:7-21) computed on the
GPU from the (n, n_ent) score matrix the model returns, so that an evaluation loop does not have to
copy the scores to the host and call scipy.stats.rankdata per batch."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr


def _csr(lists, device):
    lens = np.fromiter((len(x) for x in lists), dtype=np.int64, count=len(lists))
    p = np.zeros(len(lists) + 1, dtype=np.int32)
    np.cumsum(lens, out=p[1:])
    flat = np.concatenate([np.asarray(x, dtype=np.int32).reshape(-1) for x in lists]) if len(lists) and p[-1] else \
        np.zeros(0, dtype=np.int32)
    return torch.as_tensor(p).to(device), torch.as_tensor(flat).to(device), int(p[-1])


def filtered_ranks(scores, answers, filters):
    """scores: cuda float32 (n, n_ent); answers[i] / filters[i]: entity ids of query i's true answers /
    of all known tails for its (subject, relation) -- e.g. loader.test_a[idx], loader.filters[(s, r)].
    Returns a cuda float64 tensor of ranks in the order of cal_ranks (query-major, entity ascending)."""
    _lib.require_cuda(scores)
    scores = scores.detach().to(torch.float32).contiguous()
    n, n_ent = scores.shape
    assert len(answers) == n and len(filters) == n
    ans_ptr, ans_idx, total = _csr([np.sort(np.asarray(a)) for a in answers], scores.device)
    flt_ptr, flt_idx, _ = _csr(filters, scores.device)
    ranks = torch.empty(total, dtype=torch.float64, device=scores.device)
    if flt_idx.numel() == 0:
        flt_idx = torch.zeros(1, dtype=torch.int32, device=scores.device)
    if total:
        check(lib.rg_filtered_ranks(n, n_ent, ptr(scores), ptr(ans_ptr), ptr(ans_idx), ptr(flt_ptr), ptr(flt_idx),
                                    ptr(ranks), stream_ptr()))
        _lib.Stats.launches += 1
    return ranks


def rank_metrics(ranks):
    """cal_performance (utils.py:17-21): (MRR, Hits@1, Hits@10) of a rank tensor, still on the device."""
    r = ranks.to(torch.float64)
    n = max(r.numel(), 1)
    return float((1.0 / r).sum() / n), float((r <= 1).sum()) / n, float((r <= 10).sum()) / n
