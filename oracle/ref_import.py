"""Import the UNMODIFIED reference modules (LARS-research/RED-GNN, Static/*) under shims.

TEST INFRASTRUCTURE ONLY.  Used in the build container (where /root/reference is mounted)
to (1) pin `oracle/redgnn_oracle.py` against the live reference and (2) generate the golden
fixtures under tests/golden/ (see oracle/make_golden*.py).  `/root/reference` does not exist on
the GPU box: there the byte-identical copies staged by `oracle/build_ref.py` under `oracle/_ref/`
(git-ignored, shipped by gpurun) are used -- by `bench.py --impl reference` / `cpu_baseline`
(`kind: "reference"`) and by the tests that run the reference's own `base_model.py` trainer on
top of the drop-in.  `available()` gates every use.

Shims:
  1. `torch_scatter` is not installed -> stub module, scatter(sum) == zeros().index_add_().
     (torch_scatter 2.0.9 `scatter_sum` is `scatter_add_`; reference call site
     Static/transductive/models.py:3,39.)
  2. hard-coded `.cuda()` (models.py:69-74,81,87; load_data.py:119) -> identity on a CPU-only box.
  3. numpy>=2 rejects the ragged `np.array(self.valid_a)` (transductive/load_data.py:137,139;
     inductive/load_data.py:149,152) -> pre-convert the answer lists to 1-D object arrays.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick_root():
    live = os.environ.get("REDGNN_REFERENCE", "/root/reference")
    if os.path.isfile(os.path.join(live, "Static", "transductive", "models.py")):
        return live
    return _STAGED


REF_ROOT = _pick_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "Static", "transductive", "models.py"))


def is_live() -> bool:
    """True when REF_ROOT is the mounted reference tree (build container), False for the staged copy."""
    return REF_ROOT != _STAGED


def _install_shims(force_cpu=False):
    """force_cpu: make `.cuda()` an identity even where a GPU exists -- the CPU arm of bench.py times
    the reference's CPU path on the host cores of the GPU box."""
    if "torch_scatter" not in sys.modules:
        mod = types.ModuleType("torch_scatter")

        def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
            assert dim == 0 and reduce == "sum"
            out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
            return out.index_add_(0, index, src)

        mod.scatter = scatter
        sys.modules["torch_scatter"] = mod
    if (force_cpu or not torch.cuda.is_available()) and not getattr(torch.Tensor.cuda, "_rg_identity", False):
        _saved.update(tensor=torch.Tensor.cuda, module=torch.nn.Module.cuda)

        def _tensor_cuda(self, *a, **k):
            return self
        _tensor_cuda._rg_identity = True
        torch.Tensor.cuda = _tensor_cuda

        def _module_cuda(self, *a, **k):
            return self
        torch.nn.Module.cuda = _module_cuda


_saved = {}


def remove_cpu_shim():
    """Undo the identity `.cuda()` patch (tests that mix the CPU reference with the CUDA path)."""
    if _saved:
        torch.Tensor.cuda, torch.nn.Module.cuda = _saved.pop("tensor"), _saved.pop("module")


def _load(setting: str, name: str):
    """Load Static/<setting>/<name>.py under the module name ref_<setting>_<name>."""
    path = os.path.join(REF_ROOT, "Static", setting, name + ".py")
    modname = "ref_%s_%s" % (setting, name)
    if modname in sys.modules:
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference(setting: str, force_cpu=False):
    """Returns (load_data, models, utils) modules of Static/<setting>, unmodified."""
    assert setting in ("transductive", "inductive")
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_shims(force_cpu)
    return _load(setting, "load_data"), _load(setting, "models"), _load(setting, "utils")


def _object_array(lst):
    arr = np.empty(len(lst), dtype=object)
    for i, a in enumerate(lst):
        arr[i] = a
    return arr


def make_loader(setting: str, task_dir: str):
    """Instantiate the reference DataLoader and apply shim 3."""
    load_data, _, _ = load_reference(setting)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        loader = load_data.DataLoader(task_dir)
    loader.valid_a = _object_array(loader.valid_a)
    loader.test_a = _object_array(loader.test_a)
    return loader


def data_dir(setting: str, name: str) -> str:
    return os.path.join(REF_ROOT, "Static", setting, "data", name)


class Options(object):
    pass


def family_options(loader):
    """Static/transductive/train.py:46-56."""
    o = Options()
    o.lr, o.decay_rate, o.lamb = 0.0036, 0.999, 0.000017
    o.hidden_dim, o.attn_dim, o.n_layer = 48, 5, 3
    o.dropout, o.act, o.n_batch, o.n_tbatch = 0.29, "relu", 20, 50
    o.n_ent, o.n_rel = loader.n_ent, loader.n_rel
    return o


def fb237_v2_options(loader):
    """Static/inductive/train.py:87-96."""
    o = Options()
    o.lr, o.decay_rate, o.lamb = 0.0077, 0.993, 0.0002
    o.hidden_dim, o.attn_dim, o.n_layer = 48, 5, 3
    o.dropout, o.act, o.n_batch, o.n_tbatch = 0.3, "relu", 10, 10
    o.n_ent, o.n_rel, o.n_ent_ind = loader.n_ent, loader.n_rel, loader.n_ent_ind
    return o


import contextlib  # noqa: E402


@contextlib.contextmanager
def reference_trainer(setting: str, drop_in_dir=None, force_cpu=False):
    """Import Static/<setting>/base_model.py UNMODIFIED and yield the module.

    Its bare imports (`from models import ...`, `from utils import ...`; train.py does
    `from load_data import DataLoader`) resolve by sys.path order: with `drop_in_dir` first they land
    on the B200 drop-in (redgnn_b200/drop_in/<setting>), otherwise on the reference's own files.
    `models`, `load_data`, `utils`, `base_model` are removed from sys.modules on exit so that both
    resolutions can be used in one process.  The module `load_data` is reachable as
    `<yielded>.rg_load_data`."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_shims(force_cpu)
    names = ("models", "load_data", "utils", "base_model")
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.pop(k, None) for k in names}
    ref_dir = os.path.join(REF_ROOT, "Static", setting)
    sys.path[:0] = ([drop_in_dir] if drop_in_dir else []) + [ref_dir]
    try:
        spec = importlib.util.spec_from_file_location("base_model", os.path.join(ref_dir, "base_model.py"))
        bm = importlib.util.module_from_spec(spec)
        sys.modules["base_model"] = bm
        spec.loader.exec_module(bm)
        import load_data
        bm.rg_load_data = load_data
        yield bm
    finally:
        sys.path[:] = saved_path
        for k in names:
            sys.modules.pop(k, None)
            if saved_mods[k] is not None:
                sys.modules[k] = saved_mods[k]


def fix_ragged_answers(loader):
    """Shim 3 for a loader built by the reference's own DataLoader."""
    loader.valid_a = _object_array(loader.valid_a)
    loader.test_a = _object_array(loader.test_a)
    return loader
