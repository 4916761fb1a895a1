#!/usr/bin/env python
"""bench.py -- throughput of RED-GNN's query-conditioned propagation path on B200.

A "step" is one pass of the hot path (per-layer frontier expansion + fused attention message
passing + node update + score scatter, i.e. RED_GNN_trans.forward) over one batch of queries of a
synthetic KG of the BASELINE shape.  Default workload = BASELINE.json configs[2]: FB15k-237-shaped
synthetic KG (14,541 entities, 237 relations + inverses, 272,115 triples), n_layer=4, hidden 48,
attn 5, filtered-eval forward, per-GPU query batch fixed (weak scaling over 1/2/4/8 GPUs).

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path)
  python bench.py --impl reference --steps K --warmup W    # reference algorithm on host cores (oracle port)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the byte model behind `roofline`.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {  # name -> (synth shape, n_layer, default per-GPU batch)
    "fb15k237": ("fb15k237", 4, 64),
    "family": ("family", 3, 256),
    "yago310": ("yago310", 5, 8),
    "tiny": ("tiny", 3, 32),
    "powerlaw": ("powerlaw", 6, 4),      # built from arrays (ArrayLoader): 10 M triples, no text round trip
}
ARRAY_WORKLOADS = ("powerlaw",)
# BASELINE configs[1]: Static/inductive on an fb237_v2-shaped pair of KGs (train graph 2,608 entities /
# 9,739 triples, unseen-entity graph 1,660 entities / 4,145 triples, 200 relations), n_layer 3
INDUCTIVE_WORKLOADS = {"fb237v2": dict(n_ent=2608, n_ent_ind=1660, n_rel=200, n_train=9739, n_ind_train=4145,
                                       n_eval=1170)}
WORKLOADS["fb237v2"] = ("fb237v2", 3, 128)
HIDDEN, ATTN = 48, 5


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="fb15k237", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="queries per GPU per step (0 = workload default)")
    ap.add_argument("--train", action="store_true", help="time forward+backward+Adam instead of eval forward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-queries", type=int, default=2, help="queries per reference/CPU step (bounded sample)")
    return ap.parse_args()


def make_dataset(workload):
    from redgnn_b200 import synth
    shape, n_layer, batch = WORKLOADS[workload]
    tmp = tempfile.mkdtemp(prefix="rg_bench_")
    task = synth.write_transductive(os.path.join(tmp, shape), shape, seed=0)
    return task, n_layer, batch


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region: the poller is started early
    (nvidia-smi needs ~0.2 s before its first row) and only rows stamped inside [begin, end] count."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.t0, self.t1 = [], None, index, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        self.thread.join(timeout=2)
        t0, t1 = self.t0 or 0.0, self.t1 or float("inf")
        rows = [r for t, r in self.rows if len(r) >= 7 and t0 <= t <= t1 + 0.03]
        if not rows:                                   # region shorter than the polling period: nearest rows
            rows = [r for t, r in self.rows if len(r) >= 7 and t >= t0 - 0.05][:3]
        num = lambda x: x.replace(".", "").isdigit()
        sm = [float(r[1]) for r in rows if num(r[1])]
        mx = [float(r[2]) for r in rows if num(r[2])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_step(oracle, data, sd, n_layer, subs, rels):
    """One step of the reference algorithm on host cores (oracle port: scipy SpGEMM + torch.unique +
    torch CPU ops, same library calls as the reference)."""
    with torch.no_grad():
        scores, trace = oracle.model_forward(sd, data.test_graph, subs, rels, n_layer, "relu", return_trace=True)
    return scores, sum(int(t[1].shape[0]) for t in trace)


def run_reference(args):
    """--impl reference: rank 0 only; the other ranks exit 0 without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import redgnn_oracle as O
    if args.workload in ARRAY_WORKLOADS:
        from redgnn_b200 import synth

        class _D(object):
            pass
        ld = synth.ArrayLoader(WORKLOADS[args.workload][0], seed=0, device="cpu")
        n_layer = WORKLOADS[args.workload][1]
        data = _D()
        data.test_graph = O.Graph(ld._test_graph.triples, ld.n_ent, ld.n_rel)
        data.test_q, data.n_rel = ld.test_q, ld.n_rel
    elif args.workload in INDUCTIVE_WORKLOADS:
        from redgnn_b200 import synth

        class _D(object):
            pass
        n_layer = WORKLOADS[args.workload][1]
        task = synth.write_inductive(os.path.join(tempfile.mkdtemp(prefix="rg_bench_"), args.workload), seed=0,
                                     **INDUCTIVE_WORKLOADS[args.workload])
        ind = O.InductiveData(task)
        data = _D()
        data.test_graph, data.test_q, data.n_rel = ind.ind_graph, ind.test_q, ind.n_rel
    else:
        task, n_layer, _ = make_dataset(args.workload)
        data = O.TransductiveData(task)
    sd = O.init_state_dict(n_layer, HIDDEN, ATTN, data.n_rel, seed=1234)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    q = np.array(data.test_q)
    nq = max(1, args.cpu_queries)
    # bounded sample: if the first step projects the whole --steps/--warmup run beyond ~4 minutes, a step
    # shrinks to one query (the metric is per query, so the value is unaffected)
    t0 = time.perf_counter()
    cpu_reference_step(O, data, sd, n_layer, q[:nq, 0], q[:nq, 1])
    if nq > 1 and (time.perf_counter() - t0) * (args.warmup + args.steps) > 240.0:
        nq = 1
    batches = [q[i * nq:(i + 1) * nq] for i in range(args.warmup + args.steps)]
    for b in batches[:max(0, args.warmup - 1)]:
        cpu_reference_step(O, data, sd, n_layer, b[:, 0], b[:, 1])
    t0 = time.perf_counter()
    edges = 0
    for b in batches[args.warmup:]:
        edges += cpu_reference_step(O, data, sd, n_layer, b[:, 0], b[:, 1])[1]
    dt = time.perf_counter() - t0
    qps = nq * args.steps / dt
    sample = "%d queries/step x %d steps of the %s workload, oracle port of the reference CPU path" % (
        nq, args.steps, args.workload)
    emit({
        "impl": "reference", "metric": "queries/s", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "edges_per_s": edges / dt,
        "config": {"workload": workload_name(args.workload, n_layer), "queries_per_step": nq,
                   "hidden_dim": HIDDEN, "attn_dim": ATTN, "n_layer": n_layer},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def workload_name(workload, n_layer):
    from redgnn_b200 import synth
    if workload in INDUCTIVE_WORKLOADS:
        w = INDUCTIVE_WORKLOADS[workload]
        return ("%s-shaped synthetic inductive pair (train KG %d entities / %d triples, unseen-entity KG %d entities "
                "/ %d triples, %d relations + inverses), n_layer=%d, eval forward on the unseen-entity graph" % (
                    workload, w["n_ent"], w["n_train"], w["n_ent_ind"], w["n_ind_train"], w["n_rel"], n_layer))
    ne, nr, nt = synth.SHAPES[WORKLOADS[workload][0]][:3]
    return "%s-shaped synthetic KG (%d entities, %d relations + inverses, %d triples), n_layer=%d, eval forward" % (
        workload, ne, nr, nt, n_layer)


_REAL_STDOUT = None


def guard_stdout():
    """Libraries (NCCL, torch.distributed) print banners to stdout; the contract is ONE JSON line there.
    Route fd 1 to stderr for the whole run and keep a private handle for the result line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse_args()
    guard_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import redgnn_b200
    from redgnn_b200 import synth, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    inductive = args.workload in INDUCTIVE_WORKLOADS
    if args.workload in ARRAY_WORKLOADS:
        shape, n_layer, batch = WORKLOADS[args.workload]
        loader, task = synth.ArrayLoader(shape, seed=0, device=dev), None
    elif inductive:
        _, n_layer, batch = WORKLOADS[args.workload]
        task = synth.write_inductive(os.path.join(tempfile.mkdtemp(prefix="rg_bench_"), args.workload), seed=0,
                                     **INDUCTIVE_WORKLOADS[args.workload])
        with contextlib.redirect_stdout(io.StringIO()):
            loader = redgnn_b200.InductiveLoader(task, device=dev)
    else:
        task, n_layer, batch = make_dataset(args.workload)
        with contextlib.redirect_stdout(io.StringIO()):
            loader = redgnn_b200.TransductiveLoader(task, device=dev)
    batch = args.batch or batch
    opts = synth.Options(hidden_dim=HIDDEN, attn_dim=ATTN, n_layer=n_layer, n_rel=loader.n_rel, dropout=0.0)
    torch.manual_seed(1234)
    model = (redgnn_b200.RED_GNN_induc if inductive else redgnn_b200.RED_GNN_trans)(opts, loader).to(dev)
    optim = torch.optim.Adam(model.parameters(), lr=1e-3) if args.train else None
    model.train() if args.train else model.eval()
    if inductive:       # train on the training graph, evaluate on the unseen-entity graph
        mode = "transductive" if args.train else "inductive"
    else:
        mode = "train" if args.train else "test"

    # query stream: rank r takes batches r, r+world, ... of the test queries (train triples for --train)
    if args.train:
        pool = (loader.tra_train if inductive else loader.train_data)[:, :3]
    else:
        tq = np.array(loader.test_q)
        pool = np.concatenate([tq, np.zeros((len(tq), 1), dtype=tq.dtype)], 1)
    n_steps_total = args.warmup + args.steps
    need = batch * world * n_steps_total * 2
    reps = -(-need // len(pool))
    pool = np.concatenate([pool] * reps, 0)

    def step_batch(phase, i):
        k = ((phase * n_steps_total + i) * world + rank) * batch
        return pool[k:k + batch]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def run_step(subs, rels, objs):
        if args.train:
            optim.zero_grad(set_to_none=True)
            scores = model(subs, rels, mode)
            pos = scores[torch.arange(len(scores), device=dev), objs]
            mx = scores.max(1, keepdim=True)[0]
            loss = torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(scores - mx), 1)))
            loss.backward()
            if world > 1:
                flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
                dist.all_reduce(flat)            # SUM: the loss is a sum over queries
                off = 0
                for p in model.parameters():
                    p.grad.copy_(flat[off:off + p.numel()].view_as(p))
                    off += p.numel()
            optim.step()
            return loss
        with torch.no_grad():
            return model(subs, rels, mode=mode)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value) ----------------
    dev_batches = []
    for i in range(n_steps_total):
        b = step_batch(0, i)
        dev_batches.append(tuple(torch.as_tensor(b[:, c].astype(np.int64)).to(dev) for c in range(3)))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        run_step(*dev_batches[i])
    barrier()
    _lib.Stats.launches = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    edges_total = 0
    barrier()
    sampler.begin()
    for i in range(args.steps):
        flush.zero_()                                   # L2 flush between timed iterations (untimed)
        ev[i][0].record()
        run_step(*dev_batches[args.warmup + i])
        ev[i][1].record()
    barrier()
    sampler.end()
    launches = _lib.Stats.launches
    clocks = sampler.stop() if rank == 0 else None
    t_dev = sum(a.elapsed_time(b) for a, b in ev) * 1e-3

    # dominant kernel: fused edge forward, timed with CUDA events around every rg_edge_agg_fwd launch in
    # an instrumented pass over the SAME batches (the timed region above replays CUDA graphs, which
    # cannot hold event records).  Algorithmic bytes per launch:
    #   (16 + 4d) * E + 4d * N'   with the hidden-row gather (layers >= 1),  16 * E + 4d * N' at layer 0
    _lib.Stats.timing = []                            # routes every step through the eager, instrumented path
    for i in range(2):                                # untimed warm-up of exactly that path (allocator pools)
        run_step(*dev_batches[i])
    barrier()
    _lib.Stats.timing = []
    for i in range(args.steps):
        flush.zero_()
        run_step(*dev_batches[args.warmup + i])
        edges_total += sum(model.last_stats["edges"])
    barrier()
    timing, _lib.Stats.timing = _lib.Stats.timing, None
    # every step issues the same launches in the same order: per launch slot take the MEDIAN over the
    # steps (one-off stalls of hundreds of ms were seen right after the nvidia-smi poller exits), then add
    # the slots up -> milliseconds and algorithmic bytes of one step
    edge_ms, edge_bytes, bwd_ms, bwd_bytes = 0.0, 0.0, 0.0, 0.0
    kernel_ms = {}
    per_step = len(timing) // max(1, args.steps)
    med = lambda xs: float(np.median(xs))
    for j in range(per_step):
        slot = [timing[st * per_step + j] for st in range(args.steps)]
        name = slot[0][0]
        ms = med([a.elapsed_time(b) for _, _, a, b in slot])
        kernel_ms[name] = kernel_ms.get(name, 0.0) + ms
        if os.environ.get("RG_BENCH_DEBUG"):
            print("timing", name, [round(a.elapsed_time(b), 3) for _, _, a, b in slot], file=sys.stderr)
        if name == "edge_bwd":
            # push backward over the CSR-by-head: 16 B of structure per edge, the g_agg row of the tail node
            # per edge (4d), the hidden row of the head segment once per segment plus the g_hidden row written
            # (8d per head node, layers >= 1); alpha is recomputed, not stored
            nbytes = []
            for _, (seg, d, has_hidden), _, _ in slot:
                nbytes.append((16 + 4 * d) * (seg.n_edges or 0) + (8 * d if has_hidden else 0) * seg.n_seg)
            bwd_ms += ms
            bwd_bytes += med(nbytes)
        if name == "edge_fwd":
            nbytes = []
            for _, (seg, d, has_hidden), _, _ in slot:
                fr = getattr(seg, "frontier", None)      # sync-free path: counts resolved by model.last_stats
                n_seg, e_l = (fr.n_nodes, fr.n_edges) if fr is not None else (seg.n_seg, seg.n_edges)
                nbytes.append(((16 + 4 * d) if has_hidden else 16) * e_l + 4 * d * n_seg)
            edge_ms += ms
            edge_bytes += med(nbytes)
    t_instr = edge_ms * 1e-3 * args.steps          # same units as t_dev (all timed steps)

    # subsystem (1): the drop-in get_neighbors chain (explicit sampled_edges / tail_nodes / remap emission,
    # reference load_data.py:106-131) over the first timed batch, device time per hop between CUDA events
    # (the 16-byte count read-back between them is excluded).  Bytes: SURVEY 8(d) B_exp.
    kg = loader.graph_for(mode, dev)
    exp_ms, exp_emit_ms, exp_bytes, exp_edges = 0.0, 0.0, 0.0, 0
    for rep in range(2):                                  # rep 0 = warm-up (allocator), rep 1 = measured
        subs0 = dev_batches[args.warmup][0]
        nodes = torch.stack([torch.arange(batch, device=dev), subs0], 1)
        spans = []
        for l in range(n_layer):
            e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            e0.record()
            fr_in = kg.frontier_from_nodes(nodes, batch)
            fr_out = kg.step(fr_in)
            e1.record()
            n_in, n_e, n_out, _ = fr_out.read_counts(also=fr_in)
            if (n_e + n_out) * 48 > (40 << 30):           # keep the explicit edge list within HBM
                break
            e2.record()
            tail_nodes = fr_out.nodes64(n_out)
            remap = fr_in.remap_to(fr_out, n_in)
            edges = kg.emit_edges(fr_in, fr_out, n_e)
            e3.record()
            spans.append((e0, e1, e2, e3, 56 * n_e + 4 * kg.n_fact + 24 * n_in + 16 * n_out, n_e))
            nodes = tail_nodes
            del edges, remap
        torch.cuda.synchronize()
        if rep == 1:
            for e0, e1, e2, e3, nbytes, n_e in spans:
                exp_ms += e0.elapsed_time(e1) + e2.elapsed_time(e3)
                exp_emit_ms += e2.elapsed_time(e3)
                exp_bytes += nbytes
                exp_edges += n_e
    del nodes, tail_nodes

    # ---------------- end-to-end timing through the public API with host buffers (e2e) ----------------
    host_batches = [step_batch(1, i) for i in range(n_steps_total)]
    n_ent_out = loader.n_ent_for(mode)
    pinned_out = torch.empty((batch, n_ent_out), dtype=torch.float32).pin_memory()
    for i in range(args.warmup):
        b = host_batches[i]
        out = run_step(b[:, 0], b[:, 1], torch.as_tensor(b[:, 2]).to(dev))
        if not args.train:
            pinned_out.copy_(out)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        b = host_batches[args.warmup + i]
        if args.train:
            loss = run_step(b[:, 0], b[:, 1], torch.as_tensor(b[:, 2]).to(dev))
            loss.item()                                 # D2H read of the step's result
        else:
            out = run_step(b[:, 0], b[:, 1], None)      # numpy subs/rels: H2D inside model.forward
            pinned_out.copy_(out, non_blocking=True)    # D2H of the (n, n_ent) score matrix
            torch.cuda.current_stream().synchronize()
    barrier()
    t_e2e = time.perf_counter() - t0

    times = torch.tensor([t_dev, t_e2e, float(edges_total), edge_ms, edge_bytes], dtype=torch.float64, device=dev)
    if world > 1:
        mx = times.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = times.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_dev, t_e2e = float(mx[0]), float(mx[1])
        edges_all = float(sm[2])
    else:
        edges_all = float(edges_total)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    traffic = None                                   # DRAM bytes per launch from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "edge_fwd_traffic.json")
    if os.path.isfile(tpath) and not args.train:
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("workload") == args.workload:
            traffic = tj["traffic_bytes_per_launch"]
    qps = world * batch * args.steps / t_dev
    gbps = lambda nbytes, ms: (nbytes / 1e9) / (ms * 1e-3) if ms > 0 else 0.0
    achieved = gbps(edge_bytes, edge_ms)
    line = {
        "metric": "queries/s", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "edges_per_s": edges_all / t_dev,
        "config": {"workload": workload_name(args.workload, n_layer) + (" + backward + Adam" if args.train else ""),
                   "queries_per_gpu_per_step": batch, "hidden_dim": HIDDEN, "attn_dim": ATTN, "n_layer": n_layer,
                   "edges_per_step_per_gpu": edges_total / args.steps,
                   "l2": "flushed between timed steps (256 MiB write, untimed)", "parallelism": "dp%d" % world},
        "e2e": {"value": world * batch * args.steps / t_e2e, "unit": "queries/s",
                "h2d_bytes_per_step": int(batch * 16 + (batch * 8 if args.train else 0)),
                "d2h_bytes_per_step": int(4 if args.train else batch * n_ent_out * 4)},
        "gpu_launches": int(launches),
        "kernel_ms_per_step": {k: round(v, 4) for k, v in kernel_ms.items()},
        "roofline": {"bound": "hbm", "kernel": "k_edge_fwd (fused gather+attention+segmented reduce)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "share_of_step": t_instr / t_dev,
                     "timing": "CUDA events around every rg_edge_agg_fwd launch, instrumented pass over the same batches",
                     "bytes_model": "(16+4d)*E + 4d*N' per launch (16*E + 4d*N' at layer 0)"},
        "subsystems": {
            "expand": {"kernels": "rg_frontier_from_nodes + rg_frontier_step + rg_frontier_nodes + rg_frontier_remap "
                                  "+ rg_edges_emit (explicit get_neighbors outputs, %d hops, one batch)" % n_layer,
                       "ms": exp_ms, "ms_emit_part": exp_emit_ms, "edges": exp_edges, "achieved": gbps(exp_bytes, exp_ms), "peak": peak,
                       "unit": "GB/s", "frac": gbps(exp_bytes, exp_ms) / peak,
                       "bytes_model": "56*E + 4*n_fact + 24*N + 16*N' per hop"},
            "edge_fwd": {"ms_per_step": edge_ms, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak},
            "edge_bwd": None if not args.train else {
                "ms_per_step": bwd_ms, "achieved": gbps(bwd_bytes, bwd_ms), "peak": peak,
                "unit": "GB/s", "frac": gbps(bwd_bytes, bwd_ms) / peak,
                "bytes_model": "(16+4d)*E + 8d*N per launch (16+4d)*E at layer 0"},
        },
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        from oracle import redgnn_oracle as O
        if inductive:
            class _D(object):
                pass
            ind = O.InductiveData(task)
            data = _D()
            data.test_graph, data.test_q = ind.ind_graph, ind.test_q
        elif task is None:
            class _D(object):
                pass
            data = _D()
            data.test_graph = O.Graph(loader._test_graph.triples, loader.n_ent, loader.n_rel)
            data.test_q = loader.test_q
        else:
            data = O.TransductiveData(task)
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        nq = max(1, args.cpu_queries)
        q = np.array(data.test_q)
        cpu_reference_step(O, data, sd, n_layer, q[:nq, 0], q[:nq, 1])           # warm-up
        t0 = time.perf_counter()
        reps, edges_cpu = 0, 0
        while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 50):
            b = q[(reps + 1) * nq:(reps + 2) * nq]
            edges_cpu += cpu_reference_step(O, data, sd, n_layer, b[:, 0], b[:, 1])[1]
            reps += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": nq * reps / dt, "unit": "queries/s", "cores": cores, "kind": "port",
                                "edges_per_s": edges_cpu / dt,
                                "sample": "%d batches of %d queries of the same workload through the oracle port "
                                          "(scipy SpGEMM + torch.unique + torch CPU), %.1f s" % (reps, nq, dt)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
