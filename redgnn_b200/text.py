"""Text side of the graph store: entity / relation dictionaries and triples files -> id arrays.

`parse_triples` is the host-threaded native reader (csrc/rg_text.cpp, `rg_text_parse_triples`) of
the "head relation tail" files that DataLoader.read_triples walks line by line in Python
(Static/transductive/load_data.py:58-67, Static/inductive/load_data.py:76-86); `filter_table`
builds the (h, r) -> known-tails table of the same loops (`self.filters[(h,r)].add(t)`,
transductive :64-65; inductive get_filter :170-197) by one sort instead of one set insertion per
triple.  There is no Python fallback: without the library the import of this package fails.
"""
from collections.abc import Mapping
import ctypes as C
import os

import numpy as np

from . import _lib


def read_id_table(path, with_id):
    """entity2id / relation2id.  with_id=False: the id is the line number and the name the stripped
    line (transductive/load_data.py:11-25); True: "name id" lines (inductive/load_data.py:14-40)."""
    table = {}
    with open(path) as f:
        for k, line in enumerate(f):
            if with_id:
                name, idx = line.strip().split()
                table[name] = int(idx)
            else:
                table[line.strip()] = k
    return table


class NameTable(object):
    """A name -> id dictionary flattened for the C ABI (`rg_name_table`): the UTF-8 names back to back,
    their offsets and their ids.  Keeps the arrays alive for as long as the struct is used."""

    def __init__(self, table):
        names = [k.encode('utf-8') for k in table]
        ids = np.fromiter(table.values(), dtype=np.int64, count=len(names))
        if len(ids) and (ids.min() < -2 ** 31 or ids.max() >= 2 ** 31):
            raise ValueError("an id of the dictionary does not fit 32 bits")
        self.blob = b''.join(names)
        self.off = np.zeros(len(names) + 1, dtype=np.int64)
        if names:
            np.cumsum(np.fromiter(map(len, names), dtype=np.int64, count=len(names)), out=self.off[1:])
        self.ids = ids.astype(np.int32)
        self.struct = _lib.RgNameTable(C.cast(C.c_char_p(self.blob), C.c_void_p), self.off.ctypes.data,
                                       self.ids.ctypes.data, len(names))


def _as_table(t):
    return t if isinstance(t, NameTable) else NameTable(t)


def count_lines(path):
    n = C.c_int64(0)
    rc = _lib.lib.rg_text_count_lines(os.fsencode(path), C.byref(n))
    if rc == _lib.RG_ERR_IO:
        raise FileNotFoundError(2, "cannot read the triples file", path)
    _lib.check(rc)
    return n.value


def parse_triples(path, entity2id, relation2id, n_threads=0):
    """(n, 3) int64 [h, r, t] of the file's lines in file order.  `entity2id` / `relation2id` are
    dicts or prebuilt `NameTable`s.  Raises what the reference's loop raises at its first bad line:
    FileNotFoundError, ValueError (not exactly three names), KeyError (unknown name).  One difference: bytes
    that are not valid UTF-8 are compared as they are (an unknown name, KeyError) where Python's text
    decoding of the file would raise UnicodeDecodeError."""
    ent, rel = _as_table(entity2id), _as_table(relation2id)
    try:
        cap = (os.path.getsize(path) + 1) // 6 + 1        # "a b c\n": no line that parses is shorter than 6 bytes
    except OSError:
        raise FileNotFoundError(2, "cannot read the triples file", path)
    out = np.empty((cap, 3), dtype=np.int32)              # untouched pages cost nothing
    rows, bad = C.c_int64(0), C.c_int64(-1)
    rc = _lib.lib.rg_text_parse_triples(os.fsencode(path), C.byref(ent.struct), C.byref(rel.struct),
                                        out.ctypes.data, cap, C.byref(rows), C.byref(bad), int(n_threads))
    if rc == _lib.RG_ERR_BAD_ARG and rows.value > cap:    # more lines than 6-byte ones fit: some line is too short
        raise ValueError("%s: a line holds fewer than 3 names" % path)
    if rc == _lib.RG_ERR_IO:
        raise FileNotFoundError(2, "cannot read the triples file", path)
    if rc == _lib.RG_ERR_PARSE:
        raise ValueError("%s line %d: expected 3 names (head relation tail)" % (path, bad.value + 1))
    if rc == _lib.RG_ERR_UNKNOWN_NAME:
        raise KeyError("%s line %d: a name is missing from entity2id / relation2id" % (path, bad.value + 1))
    _lib.check(rc)
    return out[:rows.value].astype(np.int64)


class FilterTable(Mapping):
    """{(h, r): [t, ...]} held as sorted arrays: the keys in (h, r) order, one offset per key, the tails
    ascending and listed once.  Indexing returns the list the reference's dictionary holds after its
    `list(set)` pass (in ascending order; the reference's order is the set's); a missing key yields an
    empty set, as `defaultdict(lambda: set())` does (without inserting it).  Iteration, `items()`,
    `len()` and `in` behave like the dictionary's.  `rows(subs, rels)` serves a whole batch at once."""

    def __init__(self, h, r, ptr, tails):
        self.h, self.r, self.ptr, self.tails = h, r, ptr, tails

    def _find(self, key):
        try:
            kh, kr = key
            kh, kr = int(kh), int(kr)
        except (TypeError, ValueError):
            return -1
        lo, hi = np.searchsorted(self.h, kh, 'left'), np.searchsorted(self.h, kh, 'right')
        i = lo + np.searchsorted(self.r[lo:hi], kr, 'left')
        return int(i) if i < hi and self.r[i] == kr else -1

    def __getitem__(self, key):
        i = self._find(key)
        return self.tails[self.ptr[i]:self.ptr[i + 1]].tolist() if i >= 0 else set()

    def __contains__(self, key):
        return self._find(key) >= 0

    def __iter__(self):
        return iter(zip(self.h.tolist(), self.r.tolist()))

    def __len__(self):
        return len(self.h)

    def rows(self, subs, rels):
        """CSR (ptr int64 [n+1], tails int64) of the filter lists of the queries (subs[i], rels[i])."""
        subs, rels = np.asarray(subs, dtype=np.int64).reshape(-1), np.asarray(rels, dtype=np.int64).reshape(-1)
        idx = np.fromiter((self._find(k) for k in zip(subs.tolist(), rels.tolist())), dtype=np.int64, count=len(subs))
        lo = np.where(idx >= 0, self.ptr[np.maximum(idx, 0)], 0)
        n = np.where(idx >= 0, self.ptr[np.maximum(idx, 0) + 1] - lo, 0)
        ptr = np.zeros(len(subs) + 1, dtype=np.int64)
        np.cumsum(n, out=ptr[1:])
        take = np.repeat(lo - ptr[:-1], n) + np.arange(ptr[-1])
        return ptr, self.tails[take]


def filter_table(triples, n_ent):
    """The (h, r) -> known-tails table over the rows of `triples` ((n, 3) ints; several arrays are
    concatenated): same keys and the same sets of values as the reference's per-triple
    `filters[(h, r)].add(t)` followed by `list(...)`, built by one sort (see `FilterTable`)."""
    if isinstance(triples, (list, tuple)):
        triples = np.concatenate([np.zeros((0, 3), dtype=np.int64)]
                                 + [np.asarray(t, dtype=np.int64).reshape(-1, 3) for t in triples], axis=0)
    a = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    if len(a) == 0:
        z = np.zeros(0, dtype=np.int64)
        return FilterTable(z, z, np.zeros(1, dtype=np.int64), z)
    n_ent = max(int(n_ent), int(a[:, 2].max()) + 1)
    n_r = int(a[:, 1].max()) + 1
    if int(a.min()) >= 0 and (int(a[:, 0].max()) + 1) * n_r * n_ent < 2 ** 62:
        key = np.sort((a[:, 0] * n_r + a[:, 1]) * n_ent + a[:, 2])     # (h, r, t) order
        key = key[np.r_[True, key[1:] != key[:-1]]]                    # every tail once (np.unique is ~100x slower here)
        hr, t = key // n_ent, key % n_ent
        h, r = hr // n_r, hr % n_r
    else:
        a = a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
        a = a[np.r_[True, np.any(a[1:] != a[:-1], axis=1)]]
        h, r, t = a[:, 0], a[:, 1], a[:, 2]
    starts = np.flatnonzero(np.r_[True, (h[1:] != h[:-1]) | (r[1:] != r[:-1])])
    return FilterTable(h[starts].copy(), r[starts].copy(), np.r_[starts, len(t)].astype(np.int64), np.ascontiguousarray(t))
