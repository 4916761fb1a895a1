import sys, time, os, tempfile, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import redgnn_b200
from redgnn_b200 import synth, _lib
dev = torch.device("cuda", 0)
task = synth.write_transductive(os.path.join(tempfile.mkdtemp(), "fb"), "fb15k237", seed=0)
with contextlib.redirect_stdout(io.StringIO()):
    L = redgnn_b200.TransductiveLoader(task, device=dev)
torch.manual_seed(1234)
model = redgnn_b200.RED_GNN_trans(synth.Options(n_layer=4, n_rel=L.n_rel, dropout=0.0), L).to(dev)
model.train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
def step(i):
    tri = L.train_data[i * n:(i + 1) * n]
    opt.zero_grad(set_to_none=True)
    scores = model(tri[:, 0], tri[:, 1])
    pos = scores[torch.arange(n, device=dev), torch.as_tensor(tri[:, 2], device=dev)]
    mx = scores.max(1, keepdim=True)[0]
    loss = torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(scores - mx), 1)))
    loss.backward()
    opt.step()
    return loss
for i in range(3): step(i)
torch.cuda.synchronize()
_lib.Stats.timing = []
t = time.perf_counter()
for i in range(3, 8): step(i)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 5
tm = {}
for name, meta, a, b in _lib.Stats.timing:
    tm[name] = tm.get(name, 0) + a.elapsed_time(b) / 5
_lib.Stats.timing = None
print("step ms", dt * 1e3, "kernels ms/step", tm, "edges", sum(model.last_stats["edges"]))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(8, 10): step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=34, max_name_column_width=60))
