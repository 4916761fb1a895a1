"""A few graph-captured training steps on the FB15k-237-shaped KG (profiling driver for ncu)."""
import sys, os, tempfile, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import redgnn_b200
from redgnn_b200 import synth
dev = torch.device("cuda", 0)
task = synth.write_transductive(os.path.join(tempfile.mkdtemp(), "fb"), "fb15k237", seed=0)
with contextlib.redirect_stdout(io.StringIO()):
    L = redgnn_b200.TransductiveLoader(task, device=dev)
torch.manual_seed(1234)
model = redgnn_b200.RED_GNN_trans(synth.Options(n_layer=4, n_rel=L.n_rel, dropout=0.0), L).to(dev)
model.train()
model.grads_in_place = True
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for i in range(steps):
    tri = L.train_data[i * n:(i + 1) * n]
    opt.zero_grad(set_to_none=True)
    scores = model(tri[:, 0], tri[:, 1])
    pos = scores[torch.arange(n, device=dev), torch.as_tensor(tri[:, 2], device=dev)]
    mx = scores.max(1, keepdim=True)[0]
    loss = torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(scores - mx), 1)))
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("ok", float(loss))
