"""Drop-in DataLoaders: same files, attributes and methods as the reference's
Static/transductive/load_data.py and Static/inductive/load_data.py `DataLoader`, with the graph
held on the GPU (`DeviceGraph`) and `get_neighbors` served by the CUDA expansion kernels.

Only the graph store and `get_neighbors` are on the accelerated path; the text parsing, query
grouping, filters and batching below are host-side harness kept behaviour-compatible so that the
reference's base_model.py / train.py run unchanged.
"""
import os

import numpy as np
import torch

from .graph import DeviceGraph
from .text import NameTable, filter_table, parse_triples, read_id_table as _read_id_table


def _default_device():
    if not torch.cuda.is_available():
        return None
    return torch.device('cuda', torch.cuda.current_device())


def _group_queries(triples):
    """load_query (transductive/load_data.py:91-104): queries keyed by (h, r) in sorted order, the
    answers of a query in file order (the reference's list.sort is stable; so is lexsort)."""
    a = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    if len(a) == 0:
        return [], []
    a = a[np.lexsort((a[:, 1], a[:, 0]))]
    starts = np.flatnonzero(np.r_[True, (a[1:, 0] != a[:-1, 0]) | (a[1:, 1] != a[:-1, 1])])
    keys = [(int(h), int(r)) for h, r in a[starts, :2]]
    return keys, np.split(a[:, 2].copy(), starts[1:])


def _ragged(lst):
    arr = np.empty(len(lst), dtype=object)
    for i, a in enumerate(lst):
        arr[i] = a
    return arr


def _multi_hot(answers, batch_idx, n_ent):
    objs = np.zeros((len(batch_idx), n_ent))
    for i, k in enumerate(batch_idx):
        objs[i][answers[k]] = 1
    return objs


def _cached_query_array(obj, name):
    """np.array(self.valid_q) per call (load_data.py:139-146) converts the whole list every batch;
    the list never changes after __init__, so the array is built once."""
    cache = obj.__dict__.setdefault('_q_arrays', {})
    arr = cache.get(name)
    if arr is None:
        arr = cache[name] = np.array(getattr(obj, name))
    return arr


class _GraphSlot(object):
    """Host triples of one KG (or a thunk producing them) + its lazily built device copy."""

    def __init__(self, triples, n_ent, n_rel, n_triples=None):
        self._triples = triples if callable(triples) else np.asarray(triples, dtype=np.int64).reshape(-1, 3)
        self.n_ent, self.n_rel = n_ent, n_rel
        self.n_fact = (n_triples if n_triples is not None else len(self._triples)) + n_ent
        self._dev = {}

    @property
    def triples(self):
        if callable(self._triples):
            self._triples = np.asarray(self._triples(), dtype=np.int64).reshape(-1, 3)
        return self._triples

    def on(self, device):
        key = str(device)
        g = self._dev.get(key)
        if g is None:
            g = DeviceGraph(self.triples, self.n_ent, self.n_rel, device)
            self._dev = {key: g}
        return g

    def device_graph(self):
        """The device copy if one has been built (any device), else None."""
        return next(iter(self._dev.values()), None)

    def adopt(self, other):
        """Take over `other`'s device copy after it was rebuilt IN PLACE with this slot's triples:
        the DeviceGraph object (and every device address captured CUDA graphs hold) stays the same."""
        self._dev = other._dev
        other._dev = {}

    def kg(self):
        """The reference's KG array ([triples ; self-loops]) as int64."""
        ids = np.arange(self.n_ent, dtype=np.int64)
        loops = np.stack([ids, np.full_like(ids, 2 * self.n_rel), ids], axis=1)
        return np.concatenate([self.triples, loops], axis=0)


class TransductiveLoader(object):
    """Static/transductive/load_data.py:8-164."""

    def __init__(self, task_dir, device=None):
        self.task_dir = task_dir
        self.device = device if device is not None else _default_device()
        self.entity2id = _read_id_table(os.path.join(task_dir, 'entities.txt'), False)
        self.relation2id = _read_id_table(os.path.join(task_dir, 'relations.txt'), False)
        self.n_ent = len(self.entity2id)
        self.n_rel = len(self.relation2id)

        self._names = (NameTable(self.entity2id), NameTable(self.relation2id))
        self.fact_triple = self.read_triples('facts.txt')
        self.train_triple = self.read_triples('train.txt')
        self.valid_triple = self.read_triples('valid.txt')
        self.test_triple = self.read_triples('test.txt')

        self.fact_data = self.double_triple(self.fact_triple)
        self.train_data = self.double_triple(self.train_triple)
        self.valid_data = self.double_triple(self.valid_triple)
        self.test_data = self.double_triple(self.test_triple)

        self.load_graph(self.fact_data)
        self.load_test_graph(np.concatenate([self.fact_data, self.train_data], axis=0))

        self.valid_q, self.valid_a = _group_queries(self.valid_data)
        self.test_q, self.test_a = _group_queries(self.test_data)
        self.valid_a, self.test_a = _ragged(self.valid_a), _ragged(self.test_a)

        self.n_train = len(self.train_data)
        self.n_valid = len(self.valid_q)
        self.n_test = len(self.test_q)

        # load_data.py:64-65 + :51-52: (h, r) -> tails and (t, r + n_rel) -> heads over all four files
        self.filters = filter_table([self.fact_data, self.train_data, self.valid_data, self.test_data], self.n_ent)

        print('n_train:', self.n_train, 'n_valid:', self.n_valid, 'n_test:', self.n_test)

    def read_triples(self, filename):
        """load_data.py:58-67 on the host threads (rg_text_parse_triples); an (n, 3) int64 array where the
        reference builds a list of lists.  The filter entries of the same loop are built once, in
        __init__, from the doubled arrays."""
        return parse_triples(os.path.join(self.task_dir, filename), *self._names)

    def double_triple(self, triples):
        """Inverse triples appended as one block after the originals (load_data.py:69-74).  Returns an
        int64 array (the reference builds a list of lists row by row; at 10 M facts that loop alone
        takes minutes per epoch in `shuffle_train`)."""
        a = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
        inv = np.stack([a[:, 2], a[:, 1] + self.n_rel, a[:, 0]], axis=1)
        return np.concatenate([a, inv], axis=0)

    def load_graph(self, triples):
        """load_data.py:76-81.  A re-load of the same size (shuffle_train, every epoch) rebuilds the device
        copy in place, so CUDA graphs captured on the train KG survive the epoch boundary."""
        old = getattr(self, '_train_graph', None)
        self._train_graph = _GraphSlot(triples, self.n_ent, self.n_rel)
        self.n_fact = self._train_graph.n_fact
        g = old.device_graph() if old is not None else None
        if g is not None and g.device.type == 'cuda' and g.n_fact == self.n_fact:
            g.rebuild(self._train_graph.triples)
            self._train_graph.adopt(old)

    def load_test_graph(self, triples):
        self._test_graph = _GraphSlot(triples, self.n_ent, self.n_rel)
        self.tn_fact = self._test_graph.n_fact

    @property
    def KG(self):
        return self._train_graph.kg()

    @property
    def tKG(self):
        return self._test_graph.kg()

    def graph_for(self, mode, device=None):
        slot = self._train_graph if mode == 'train' else self._test_graph     # load_data.py:107-112
        return slot.on(device if device is not None else self.device)

    def n_ent_for(self, mode):
        return self.n_ent

    def get_neighbors(self, nodes, mode='train'):
        return self.graph_for(mode).get_neighbors(nodes)

    def get_batch(self, batch_idx, steps=2, data='train'):
        if data == 'train':
            return np.asarray(self.train_data)[batch_idx]      # fancy indexing copies the batch only
        if data == 'valid':
            query, answer = self._query_array('valid_q'), self.valid_a
        if data == 'test':
            query, answer = self._query_array('test_q'), self.test_a
        subs = query[batch_idx, 0]
        rels = query[batch_idx, 1]
        return subs, rels, _multi_hot(answer, batch_idx, self.n_ent)

    def _query_array(self, name):
        return _cached_query_array(self, name)

    def shuffle_train(self):
        """load_data.py:152-164: re-split facts+train 3:1 with np.random (same stream, same permutation
        as the reference) and rebuild the graph.  With the train KG on a GPU the re-split itself runs
        there (rg_graph_resplit on the resident fact+train pool, then rg_graph_build, both in place):
        only the 4-byte-per-triple permutation crosses PCIe, and the host keeps just the quarter that
        becomes `train_data`; `fact_data` / `KG` are materialised lazily if someone asks."""
        pool = self.__dict__.get('_pool')
        if pool is None:
            pool = self._pool = np.concatenate([np.array(self.fact_triple, dtype=np.int64).reshape(-1, 3),
                                                np.array(self.train_triple, dtype=np.int64).reshape(-1, 3)], axis=0)
        n_all = len(pool)
        rand_idx = np.random.permutation(n_all)
        n_keep = n_all * 3 // 4
        self.train_data = self.double_triple(pool[rand_idx[n_keep:]])
        self.n_train = len(self.train_data)
        g = self._train_graph.device_graph()
        if g is None or g.device.type != 'cuda' or 2 * n_keep + self.n_ent != g.n_fact:
            self.fact_data = self.double_triple(pool[rand_idx[:n_keep]])
            self.load_graph(self.fact_data)
            return
        pool_dev = self.__dict__.get('_pool_dev')
        if pool_dev is None or pool_dev.device != g.device:
            pool_dev = self._pool_dev = torch.as_tensor(pool.astype(np.int32)).to(g.device)
        perm_dev = torch.as_tensor(rand_idx[:n_keep].astype(np.int32)).to(g.device, non_blocking=True)
        g.resplit(pool_dev, perm_dev, n_keep)
        keep = rand_idx[:n_keep]
        slot = _GraphSlot(lambda: self.double_triple(pool[keep]), self.n_ent, self.n_rel, n_triples=2 * n_keep)
        slot.adopt(self._train_graph)
        self._train_graph = slot
        self.n_fact = slot.n_fact
        self.__dict__.pop('fact_data', None)          # stale; the property below rebuilds it on demand

    def __getattr__(self, name):
        if name == 'fact_data':                       # only reached when the attribute is not set
            tg = self.__dict__.get('_train_graph')
            if tg is not None:
                return tg.triples
        raise AttributeError(name)


class InductiveLoader(object):
    """Static/inductive/load_data.py:8-197."""

    def __init__(self, task_dir, device=None):
        self.trans_dir = task_dir
        self.ind_dir = task_dir + '_ind'
        self.device = device if device is not None else _default_device()
        self.entity2id = _read_id_table(os.path.join(task_dir, 'entities.txt'), True)
        self.relation2id = _read_id_table(os.path.join(task_dir, 'relations.txt'), True)
        self.entity2id_ind = _read_id_table(os.path.join(self.ind_dir, 'entities.txt'), True)
        id2relation = list(self.relation2id.keys())
        self.id2relation = id2relation + [r + '_inv' for r in id2relation] + ['idd']

        self.n_ent = len(self.entity2id)
        self.n_rel = len(self.relation2id)
        self.n_ent_ind = len(self.entity2id_ind)
        rel_names = NameTable(self.relation2id)
        self._names = {'transductive': (NameTable(self.entity2id), rel_names),
                       'inductive': (NameTable(self.entity2id_ind), rel_names)}

        self.tra_train = self.read_triples(self.trans_dir, 'train.txt')
        self.tra_valid = self.read_triples(self.trans_dir, 'valid.txt')
        self.tra_test = self.read_triples(self.trans_dir, 'test.txt')
        self.ind_train = self.read_triples(self.ind_dir, 'train.txt', 'inductive')
        self.ind_valid = self.read_triples(self.ind_dir, 'valid.txt', 'inductive')
        self.ind_test = self.read_triples(self.ind_dir, 'test.txt', 'inductive')

        self.val_filters = self.get_filter('valid')
        self.tst_filters = self.get_filter('test')

        self._tra_graph = self.load_graph(self.tra_train)
        self._ind_graph = self.load_graph(self.ind_train, 'inductive')

        self.tra_train = np.array(self.tra_valid)
        self.tra_val_qry, self.tra_val_ans = _group_queries(self.tra_test)
        self.ind_val_qry, self.ind_val_ans = _group_queries(self.ind_valid)
        self.ind_tst_qry, self.ind_tst_ans = _group_queries(self.ind_test)
        self.valid_q, self.valid_a = self.tra_val_qry, _ragged(self.tra_val_ans)
        self.test_q = self.ind_val_qry + self.ind_tst_qry
        self.test_a = _ragged(self.ind_val_ans + self.ind_tst_ans)

        self.n_train = len(self.tra_train)
        self.n_valid = len(self.valid_q)
        self.n_test = len(self.test_q)

        print('n_train:', self.n_train, 'n_valid:', self.n_valid, 'n_test:', self.n_test)

    def read_triples(self, directory, filename, mode='transductive'):
        """(h,r,t) and its inverse interleaved per line (inductive/load_data.py:76-86)."""
        a = parse_triples(os.path.join(directory, filename), *self._names[mode])
        out = np.empty((2 * len(a), 3), dtype=np.int64)
        out[0::2] = a
        out[1::2, 0], out[1::2, 1], out[1::2, 2] = a[:, 2], a[:, 1] + self.n_rel, a[:, 0]
        return out

    def load_graph(self, triples, mode='transductive'):
        n_ent = self.n_ent if mode == 'transductive' else self.n_ent_ind
        return _GraphSlot(triples, n_ent, self.n_rel)

    @property
    def tra_KG(self):
        return self._tra_graph.kg()

    @property
    def ind_KG(self):
        return self._ind_graph.kg()

    def graph_for(self, mode, device=None):
        slot = self._tra_graph if mode == 'transductive' else self._ind_graph    # :118-125
        return slot.on(device if device is not None else self.device)

    def n_ent_for(self, mode):
        return self.n_ent if mode == 'transductive' else self.n_ent_ind

    def get_neighbors(self, nodes, mode='transductive'):
        return self.graph_for(mode).get_neighbors(nodes)

    def get_batch(self, batch_idx, steps=2, data='train'):
        if data == 'train':
            return self.tra_train[batch_idx]
        if data == 'valid':
            query, answer, n_ent = self._query_array('valid_q'), self.valid_a, self.n_ent
        if data == 'test':
            query, answer, n_ent = self._query_array('test_q'), self.test_a, self.n_ent_ind
        subs = query[batch_idx, 0]
        rels = query[batch_idx, 1]
        return subs, rels, _multi_hot(answer, batch_idx, n_ent)

    def _query_array(self, name):
        return _cached_query_array(self, name)

    def shuffle_train(self):
        rand_idx = np.random.permutation(self.n_train)
        self.tra_train = self.tra_train[rand_idx]

    def get_filter(self, data='valid'):
        """inductive/load_data.py:170-197; the values are lists already (:43-46)."""
        if data == 'valid':
            return filter_table([self.tra_train, self.tra_valid, self.tra_test], self.n_ent)
        return filter_table([self.ind_train, self.ind_valid, self.ind_test], self.n_ent_ind)
