"""Multi-GPU plumbing for the path: one process per GPU, queries sharded, model + KG replicated.

The reference is single-process.  Every query expands and propagates in its
own (batch_idx, .) node space (reference Static/transductive/load_data.py:115-123,
models.py:33), so the path shards over queries with NO data-path collective.  Training adds one
flat-buffer all-reduce(SUM) of the gradients per step: the reference loss is a SUM over the
batch (base_model.py:58-60), so summing shard gradients reproduces the single-GPU update for
the concatenated batch.  Evaluation all-reduces four scalars of rank statistics.
"""
import numpy as np
import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, world):
    """Contiguous slice [lo, hi) of n items for `rank`; sizes differ by at most one."""
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_batch(arrays, rank=None, world=None):
    """Slice every array of a batch (subs, rels, objs, ...) along dim 0 for this rank."""
    if rank is None:
        rank, world = world_info()
    lo, hi = shard_bounds(len(arrays[0]), rank, world)
    return tuple(a[lo:hi] for a in arrays)


def allreduce_gradients(parameters, group=None):
    """ONE flat all-reduce(SUM) over all gradients (message <= ~1.2 MB: latency-bound, so a single
    call; parameters without a gradient contribute zeros so that every rank sends the same layout)."""
    params = [p for p in parameters if p.requires_grad]
    if not params:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        g = flat[off:off + p.numel()].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += p.numel()
    return flat.numel()


def allreduce_flat(flat, group=None):
    """All-reduce(SUM) of an already-flat gradient buffer in place (RedGNN.flat_grad() with
    `grads_in_place`: the parameters' .grad are views of it, so there is nothing to pack or unpack)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat.numel()


def allreduce_model_gradients(model, group=None):
    """The gradient exchange of one training step: the flat in-place buffer when the model provides
    one, else the pack / all-reduce / unpack path over the parameters."""
    flat = model.flat_grad() if hasattr(model, "flat_grad") else None
    if flat is not None:
        return allreduce_flat(flat, group)
    return allreduce_gradients(model.parameters(), group)


def broadcast_parameters(module, src=0, group=None):
    """Make every rank start from rank `src`'s weights (after that identical SUM-reduced gradients
    and identical optimiser steps keep them in sync without further broadcasts)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def sharded_train_step(model, optimizer, triples, rank=None, world=None, group=None):
    """One training step of base_model.py:49-70 on this rank's shard of `triples` [B,3]=(h,r,t):
    forward, sum-loss, backward, gradient all-reduce(SUM), optimiser step.  Returns the local loss."""
    if rank is None:
        rank, world = world_info()
    (tri,) = shard_batch((np.asarray(triples),), rank, world)
    optimizer.zero_grad(set_to_none=True)
    if len(tri):
        scores = model(tri[:, 0], tri[:, 1])
        dev = scores.device
        pos = scores[torch.arange(len(scores), device=dev), torch.as_tensor(tri[:, 2], device=dev)]
        mx = scores.max(1, keepdim=True)[0]
        loss = torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(scores - mx), 1)))
        loss.backward()
        allreduce_model_gradients(model, group)
    else:
        # empty shard: contribute zeros with the same message layout (the flat in-place buffer would
        # still hold the previous step's gradients)
        loss = torch.zeros((), device=next(model.parameters()).device)
        allreduce_gradients(model.parameters(), group)
    optimizer.step()
    return loss.detach()


def reduce_rank_stats(ranks, device=None, group=None):
    """Filtered-ranking metrics over all ranks' queries (utils.py:17-21 cal_performance):
    all-reduce(SUM) of (sum 1/rank, #rank<=1, #rank<=10, count) -> (MRR, H@1, H@10, count)."""
    r = np.asarray(ranks, dtype=np.float64)
    stats = torch.tensor([(1.0 / r).sum() if len(r) else 0.0, float((r <= 1).sum()), float((r <= 10).sum()),
                          float(len(r))], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    s = stats.cpu().numpy()
    n = max(s[3], 1.0)
    return s[0] / n, s[1] / n, s[2] / n, int(s[3])
