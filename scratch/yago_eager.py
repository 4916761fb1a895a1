import sys, os, tempfile, io, contextlib, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import redgnn_b200
from redgnn_b200 import synth, _lib
import bench
dev = torch.device("cuda", 0)
shape, n_layer, batch = bench.WORKLOADS["yago310"]
task = synth.write_transductive(os.path.join(tempfile.mkdtemp(), shape), shape, seed=0)
with contextlib.redirect_stdout(io.StringIO()):
    L = redgnn_b200.TransductiveLoader(task, device=dev)
torch.manual_seed(1234)
model = redgnn_b200.RED_GNN_trans(synth.Options(n_layer=n_layer, n_rel=L.n_rel, dropout=0.0), L).to(dev).eval()
q = np.array(L.test_q)
n = 8
def step(i):
    b = q[i * n:(i + 1) * n]
    with torch.no_grad():
        return model(b[:, 0], b[:, 1], mode="test")
for graph in (True, False):
    model.use_cuda_graph = graph
    for i in range(3): step(i)
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(3, 8): step(i)
    torch.cuda.synchronize(); print("graph", graph, "ms/step", (time.perf_counter() - t) / 5 * 1e3)
_lib.Stats.timing = []
for i in range(8, 10): step(i)
torch.cuda.synchronize()
_lib.Stats.timing = []
for i in range(10, 13): step(i)
torch.cuda.synchronize()
for name, meta, a, b in _lib.Stats.timing:
    print(name, round(a.elapsed_time(b), 3))
