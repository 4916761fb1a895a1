"""Kernel-level breakdown (torch profiler / CUPTI) of the eval forward on a bench workload.
usage: profile_eval.py <workload: fb15k237|yago310|powerlaw> <batch>"""
import sys, os, tempfile, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import redgnn_b200
from redgnn_b200 import synth
import bench
dev = torch.device("cuda", 0)
wl = sys.argv[1]
shape, n_layer, batch = bench.WORKLOADS[wl]
n = int(sys.argv[2]) if len(sys.argv) > 2 else batch
if wl in bench.ARRAY_WORKLOADS:
    L = synth.ArrayLoader(shape, seed=0, device=dev)
else:
    task = synth.write_transductive(os.path.join(tempfile.mkdtemp(), shape), shape, seed=0)
    with contextlib.redirect_stdout(io.StringIO()):
        L = redgnn_b200.TransductiveLoader(task, device=dev)
torch.manual_seed(1234)
model = redgnn_b200.RED_GNN_trans(synth.Options(n_layer=n_layer, n_rel=L.n_rel, dropout=0.0), L).to(dev).eval()
q = np.array(L.test_q)
def step(i):
    b = q[i * n:(i + 1) * n]
    with torch.no_grad():
        return model(b[:, 0], b[:, 1], mode="test")
for i in range(3): step(i)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3, 6): step(i)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print("workload", wl, "batch", n, "edges", model.last_stats["edges"], "GPU ms/step %.3f" % (tot / 3e3))
for e in rows[:14]:
    print("%-70s %8.3f ms/step  x%d" % (e.key[:70], e.device_time_total / 3e3, e.count // 3))
