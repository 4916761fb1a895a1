import sys, time, os, tempfile, io, contextlib, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import redgnn_b200
from redgnn_b200 import synth, _lib
dev = torch.device("cuda", 0)
task = synth.write_inductive(os.path.join(tempfile.mkdtemp(), "fb237v2"), seed=0, n_ent=2608, n_ent_ind=1660, n_rel=200, n_train=9739, n_ind_train=4145, n_eval=1170)
with contextlib.redirect_stdout(io.StringIO()):
    L = redgnn_b200.InductiveLoader(task, device=dev)
torch.manual_seed(1234)
model = redgnn_b200.RED_GNN_induc(synth.Options(n_layer=3, n_rel=L.n_rel, dropout=0.1), L).to(dev)
model.train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
n = 10
def step(i):
    tri = L.tra_train[i * n:(i + 1) * n]
    opt.zero_grad(set_to_none=True)
    scores = model(tri[:, 0], tri[:, 1])
    pos = scores[torch.arange(n, device=dev), torch.as_tensor(tri[:, 2], device=dev)]
    mx = scores.max(1, keepdim=True)[0]
    loss = torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(scores - mx), 1)))
    loss.backward()
    opt.step()
    return loss
for i in range(5): step(i)
torch.cuda.synchronize()
t = time.perf_counter()
for i in range(5, 55): step(i)
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t) / 50 * 1e3)
pr = cProfile.Profile(); pr.enable()
for i in range(55, 105): step(i)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
ms0 = torch.cuda.memory_stats()
for i in range(105, 155): step(i)
torch.cuda.synchronize()
ms1 = torch.cuda.memory_stats()
for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "allocation.all.allocated", "segment.all.allocated"):
    print(k, ms1.get(k, 0) - ms0.get(k, 0))
print("reserved MB", torch.cuda.memory_reserved() / 2**20, "allocated MB", torch.cuda.memory_allocated() / 2**20)
import torch.utils.benchmark as tb
x = torch.randn(20000, 24, device=dev)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(200): y = x[:, 8:16].sum(0)
print("sum call us (async)", (time.perf_counter() - t) / 200 * 1e6)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(200): y = torch.empty(1000, 48, device=dev)
print("empty call us", (time.perf_counter() - t) / 200 * 1e6)
