"""Full-split fixtures from the LIVE, UNMODIFIED reference trainer (run in the build container):
    python oracle/make_golden_full.py

TEST INFRASTRUCTURE ONLY.  For the bundled `family` (transductive, BASELINE configs[0]) and
`fb237_v2` (inductive, configs[1]) datasets the reference's own `BaseModel` (base_model.py, imported
unmodified, CPU under the `.cuda()` identity shim) is driven exactly as train.py drives it:

  * seed 1234 (train.py:20-21), the dataset's hyper-parameters (train.py:46-56 / inductive
    train.py:87-96) with dropout set to 0 so the loss sequence is a deterministic function of the
    weights (the CUDA path draws its dropout masks from a different RNG stream);
  * `train_batch()` over the first N_STEPS batches of the epoch (`n_train` capped, nothing else
    touched): per-step training loss (recomputed from the scores a forward hook captures, with
    base_model.py:58-60's formula), the parameters before and after;
  * `evaluate()` with the trained parameters over the COMPLETE valid and test splits: the
    out_str metrics and every filtered rank (`cal_ranks` outputs captured in call order).

tests/ compare the CUDA path with these: through `O.cal_ranks`, through `metrics.filtered_ranks`, and
by running the same unmodified `base_model.py` on top of the drop-in (oracle/_ref copy on the GPU box).
"""
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
N_STEPS = 30


def parse_out_str(s):
    v = [float(x) for x in re.findall(r"(?:MRR|H@1|H@10):([0-9.]+)", s)]
    return np.array(v[:6], dtype=np.float64)          # v_mrr v_h1 v_h10 t_mrr t_h1 t_h10 (4 decimals)


def drive(setting, dataset, options_fn):
    np.random.seed(1234)
    torch.manual_seed(1234)
    with R.reference_trainer(setting) as bm:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            loader = bm.rg_load_data.DataLoader(R.data_dir(setting, dataset))
        R.fix_ragged_answers(loader)
        opts = options_fn(loader)
        opts.dropout = 0.0
        opts.mem_file = None
        trainer = bm.BaseModel(opts, loader)
        fx = {"n_steps": np.int64(N_STEPS), "n_batch": np.int64(opts.n_batch)}
        for k, v in trainer.model.state_dict().items():
            fx["sd0." + k] = v.detach().numpy().astype(np.float32).copy()
        # capture scores of every training forward; cap the epoch at N_STEPS batches
        captured = []
        hook = trainer.model.register_forward_hook(lambda m, i, o: captured.append(o.detach().clone()))
        full_n_train = loader.n_train
        train_rows = np.array(loader.get_batch(np.arange(N_STEPS * opts.n_batch))).copy()
        loader.n_train = N_STEPS * opts.n_batch
        trainer.n_train = loader.n_train                      # the inductive trainer copies it in __init__
        ranks = []
        real_cal_ranks = bm.cal_ranks

        def spy(scores, labels, filters):
            r = real_cal_ranks(scores, labels, filters)
            ranks.append(np.array(r, dtype=np.float64))
            return r

        bm.cal_ranks = spy
        _, out_str = trainer.train_batch()
        hook.remove()
        loader.n_train = trainer.n_train = full_n_train
        losses = []
        for i in range(N_STEPS):
            tri = train_rows[i * opts.n_batch:(i + 1) * opts.n_batch]
            sc = captured[i].double()
            pos = sc[torch.arange(len(sc)), torch.as_tensor(tri[:, 2])]
            mx = sc.max(1, keepdim=True)[0]
            losses.append(float(torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(sc - mx), 1)))))
        fx["train_rows"] = train_rows.astype(np.int64)
        fx["losses"] = np.array(losses)
        for k, v in trainer.model.state_dict().items():
            fx["sd1." + k] = v.detach().numpy().astype(np.float32).copy()
        fx["metrics"] = parse_out_str(out_str)
        n_valid_batches = -(-loader.n_valid // (opts.n_tbatch if setting == "transductive" else opts.n_batch))
        fx["valid_ranks"] = np.concatenate(ranks[:n_valid_batches]).astype(np.float32)
        fx["test_ranks"] = np.concatenate(ranks[n_valid_batches:]).astype(np.float32)
        fx["n_valid"], fx["n_test"] = np.int64(loader.n_valid), np.int64(loader.n_test)
        fx["eval_batch"] = np.int64(opts.n_tbatch if setting == "transductive" else opts.n_batch)
        # the ranks reproduce the printed metrics
        U = sys.modules["utils"]
        chk = list(U.cal_performance(fx["valid_ranks"].astype(np.float64))) + \
            list(U.cal_performance(fx["test_ranks"].astype(np.float64)))
        assert np.allclose(np.round(chk, 4), fx["metrics"], atol=1.01e-4), (chk, fx["metrics"])
        fx["metrics_exact"] = np.array(chk, dtype=np.float64)
    path = os.path.join(OUT, "%s_full.npz" % dataset)
    np.savez_compressed(path, **fx)
    print(dataset, "losses", np.round(fx["losses"][:5], 4), "... metrics", fx["metrics"], os.path.getsize(path))


if __name__ == "__main__":
    if not R.is_live():
        raise SystemExit("reference tree not mounted; fixtures can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["family", "fb237_v2"]
    if "family" in which:
        drive("transductive", "family", R.family_options)
    if "fb237_v2" in which:
        drive("inductive", "fb237_v2", R.fb237_v2_options)
