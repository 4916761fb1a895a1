"""get_neighbors chain (explicit emission) on the FB15k-237-shaped KG, n queries, 4 hops; prints per-hop device ms."""
import sys, os, tempfile, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import redgnn_b200
from redgnn_b200 import synth
dev = torch.device("cuda", 0)
task = synth.write_transductive(os.path.join(tempfile.mkdtemp(), "fb"), "fb15k237", seed=0)
with contextlib.redirect_stdout(io.StringIO()):
    L = redgnn_b200.TransductiveLoader(task, device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kg = L.graph_for("test", dev)
q = np.array(L.test_q)[:n]
for rep in range(3):
    nodes = torch.stack([torch.arange(n, device=dev), torch.as_tensor(q[:, 0], device=dev)], 1)
    out = []
    for l in range(4):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        fr_in = kg.frontier_from_nodes(nodes, n)
        fr_out = kg.step(fr_in)
        e[1].record()
        n_in, n_e, n_out, _ = fr_out.read_counts(also=fr_in)
        tail_nodes = fr_out.nodes64(n_out)
        remap = fr_in.remap_to(fr_out, n_in)
        e[2].record()
        edges = kg.emit_edges(fr_in, fr_out, n_e)
        e[3].record()
        torch.cuda.synchronize()
        out.append((n_e, round(e[0].elapsed_time(e[1]), 3), round(e[2].elapsed_time(e[3]), 3),
                    round(48e-6 * n_e / max(e[2].elapsed_time(e[3]), 1e-9), 1)))
        nodes = tail_nodes
        del edges
    print("rep", rep, "(E, step ms, emit ms, emit GB/s):", out)
