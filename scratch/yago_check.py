import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import redgnn_b200
from redgnn_b200 import synth, _lib
from oracle import redgnn_oracle as O
dev = torch.device("cuda", 0)
L = synth.ArrayLoader("yago310", seed=0, device=dev)
n_layer = 5
torch.manual_seed(1234)
model = redgnn_b200.RED_GNN_trans(synth.Options(n_layer=n_layer, n_rel=L.n_rel), L).to(dev).eval()
q = np.array(L.test_q)
subs, rels = q[:8, 0], q[:8, 1]
g = L.graph_for("test")
print("heavy_in", g.heavy_in, "heavy_out", g.heavy_out, "n_fact", g.n_fact)
with torch.no_grad():
    model.use_cuda_graph = False
    for rep in range(3):
        _lib.Stats.timing = []
        torch.cuda.synchronize(); t = time.perf_counter()
        out_e = model(subs, rels, mode="test")
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        tm = {}
        for name, meta, a, b in _lib.Stats.timing:
            tm.setdefault(name, []).append(round(a.elapsed_time(b), 3))
        _lib.Stats.timing = None
        print("eager", rep, round(dt * 1e3, 2), "ms", tm, model.last_stats)
    model.use_cuda_graph = True
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        out_g = model(subs, rels, mode="test")
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print("graph", rep, round(dt * 1e3, 2), "ms")
    print("graph == eager", torch.equal(out_e, out_g))
# oracle on 2 queries
t = time.perf_counter()
og = O.Graph(L._test_graph.triples, L.n_ent, L.n_rel)
sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
want = O.model_forward(sd, og, subs[:2], rels[:2], n_layer, "relu")
print("oracle s", time.perf_counter() - t)
with torch.no_grad():
    got = model(subs[:2], rels[:2], mode="test").cpu()
err = (got - want).abs().max() / want.abs().max()
print("rel err", err.item(), "zeros equal", torch.equal(got == 0, want == 0))
