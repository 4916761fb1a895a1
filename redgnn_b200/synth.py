"""Synthetic KGs for the package's tests and bench: the generator itself lives in the top-level,
numpy-only module `kg_synth` (importable without the CUDA library); this module re-exports it and
adds the device-backed `ArrayLoader`."""
import numpy as np

from kg_synth import SHAPES, zipf_triples, write_transductive, write_inductive, Options, ArraySplits  # noqa: F401


class ArrayLoader(object):
    """Loader for LARGE synthetic KGs built straight from arrays (kg_synth.ArraySplits):
    the subset of the DataLoader surface the model and bench.py need (`graph_for`, `n_ent_for`,
    `n_ent`, `n_rel`, `train_data`, `test_q`)."""

    def __init__(self, shape="powerlaw", seed=0, device=None, override=None):
        import torch
        from .data import _GraphSlot
        sp = ArraySplits(shape, seed, override)
        self.n_ent, self.n_rel, self.n_layer = sp.n_ent, sp.n_rel, sp.n_layer
        self.device = device if device is not None else (
            torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else None)
        self.train_data = sp.train_data
        self._train_graph = _GraphSlot(sp.train_graph_triples, sp.n_ent, sp.n_rel)
        self._test_graph = _GraphSlot(sp.test_graph_triples, sp.n_ent, sp.n_rel)
        self.n_fact, self.tn_fact = self._train_graph.n_fact, self._test_graph.n_fact
        self.test_q = sp.test_q
        self.n_train, self.n_test = len(self.train_data), len(self.test_q)

    def graph_for(self, mode, device=None):
        slot = self._train_graph if mode == 'train' else self._test_graph
        return slot.on(device if device is not None else self.device)

    def n_ent_for(self, mode):
        return self.n_ent

    def get_neighbors(self, nodes, mode='train'):
        return self.graph_for(mode).get_neighbors(nodes)
