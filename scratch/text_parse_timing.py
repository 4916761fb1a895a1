"""Host timing of the text side (f3): native reader + sorted filter table against the reference's
Python loops (restated in the oracle), on a power-law-sized triples file.  CPU only.
    python scratch/text_parse_timing.py [n_triples] [n_ent] [n_rel]"""
import os
import sys
import tempfile
import time
from collections import defaultdict

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import redgnn_oracle as O   # noqa: E402
from redgnn_b200 import text            # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
n_ent = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
n_rel = int(sys.argv[3]) if len(sys.argv) > 3 else 500
rng = np.random.default_rng(0)
h, r, t = rng.integers(0, n_ent, n), rng.integers(0, n_rel, n), rng.integers(0, n_ent, n)
with tempfile.TemporaryDirectory() as tmp:
    with open(os.path.join(tmp, "entities.txt"), "w") as f:
        f.write("".join("e%d\n" % i for i in range(n_ent)))
    with open(os.path.join(tmp, "relations.txt"), "w") as f:
        f.write("".join("r%d\n" % i for i in range(n_rel)))
    path = os.path.join(tmp, "facts.txt")
    with open(path, "w") as f:
        f.write("".join("e%d\tr%d\te%d\n" % x for x in zip(h.tolist(), r.tolist(), t.tolist())))
    mb = os.path.getsize(path) / 1e6
    t0 = time.perf_counter()
    ent = text.read_id_table(os.path.join(tmp, "entities.txt"), False)
    rel = text.read_id_table(os.path.join(tmp, "relations.txt"), False)
    t1 = time.perf_counter()
    tabs = (text.NameTable(ent), text.NameTable(rel))
    t2 = time.perf_counter()
    got = text.parse_triples(path, *tabs)
    t3 = time.perf_counter()
    got1 = text.parse_triples(path, *tabs, n_threads=1)
    t4 = time.perf_counter()
    inv = np.stack([got[:, 2], got[:, 1] + n_rel, got[:, 0]], 1)
    filt = text.filter_table([got, inv], n_ent)
    t5 = time.perf_counter()
    want = O._read_triples(path, ent, rel)
    t6 = time.perf_counter()
    ref = defaultdict(set)
    for a, b, c in want:
        ref[(a, b)].add(c)
        ref[(c, b + n_rel)].add(a)
    for k in ref:
        ref[k] = list(ref[k])
    t7 = time.perf_counter()
    assert np.array_equal(got, np.array(want)) and np.array_equal(got, got1)
    assert len(filt) == len(ref)
print("%d triples, %.0f MB, %d host threads" % (n, mb, os.cpu_count()))
print("dictionaries (python)      %.2f s   flatten for the C ABI %.2f s" % (t1 - t0, t2 - t1))
print("rg_text_parse_triples      %.3f s (%.0f MB/s)   1 thread %.3f s" % (t3 - t2, mb / (t3 - t2), t4 - t3))
print("reference loop (oracle)    %.2f s   -> x%.0f" % (t6 - t5, (t6 - t5) / (t3 - t2)))
print("filter_table (sorted)      %.2f s   reference set insertions %.2f s" % (t5 - t4, t7 - t6))
