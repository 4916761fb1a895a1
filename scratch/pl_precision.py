"""Where does the score error at the power-law shape come from?  CUDA path (tensor-core vs CUDA-core node
update) against the fp64 / fp32 oracle fixture."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import redgnn_oracle as O
from redgnn_b200 import RED_GNN_trans
from redgnn_b200.synth import Options, ArrayLoader
fx = np.load("tests/golden/plscaled.npz")
L = ArrayLoader(override=(100000, 500, 1000000, 500, 500, 1.0, 1.0, 6), seed=0)
sd = O.init_state_dict(6, 48, 5, L.n_rel, seed=1234)
model = RED_GNN_trans(Options(n_layer=6, n_rel=L.n_rel), L).cuda()
model.load_state_dict(sd); model.eval(); model.use_cuda_graph = False
q = fx["queries"]
w64 = torch.from_numpy(fx["scores64"]).double(); w32 = torch.from_numpy(fx["scores"]).double()
scale = float(w64.abs().max())
deg = np.bincount(L._test_graph.triples[:, 2], minlength=L.n_ent)
for simt in ("0", "1"):
    os.environ["REDGNN_NODE_SIMT"] = simt
    with torch.no_grad():
        got = model(q[:, 0], q[:, 1], mode="test").cpu().double()
    e64 = (got - w64).abs(); e32 = (got - w32).abs()
    i = int(e64.argmax()) % L.n_ent
    print("SIMT=%s  vs fp64 %.3e  vs fp32 %.3e  | worst entity %d in-degree %d | #entities with err>1e-4*scale: %d" % (
        simt, float(e64.max()) / scale, float(e32.max()) / scale, i, deg[i], int((e64 > 1e-4 * scale).sum())))
print("fp32 oracle vs fp64 %.3e" % (float((w32 - w64).abs().max()) / scale))
