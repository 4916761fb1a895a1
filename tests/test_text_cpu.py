"""CPU: the native text reader (csrc/rg_text.cpp through redgnn_b200.text) against the oracle's
line-by-line restatement of DataLoader.read_triples (Static/transductive/load_data.py:58-67,
Static/inductive/load_data.py:76-86) -- same ids in file order, same first error -- and the sorted
filter table against the reference's per-triple set insertions (:64-65, inductive :170-197)."""
from collections import defaultdict
import ctypes
import os

import numpy as np
import pytest

from oracle import redgnn_oracle as O
from redgnn_b200 import _lib, text

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_read(path, ent, rel):
    return np.array(O._read_triples(path, ent, rel), dtype=np.int64).reshape(-1, 3)


def write(path, data):
    with open(path, "wb") as f:
        f.write(data if isinstance(data, bytes) else data.encode("utf-8"))
    return str(path)


ENT = {"a": 0, "b": 1, "c": 2, "Zürich": 3, "東京": 4, "x" * 300: 5, "": 6, "two words": 7}
REL = {"likes": 0, "r/2": 1, "é": 2}


def test_plain_file_and_every_line_ending(tmp_path):
    body = ["a likes b", "b\tr/2\tc", "  c   likes   a  ", "Zürich é 東京", "x" * 300 + " likes a"]
    for name, sep, tail in (("lf", "\n", "\n"), ("crlf", "\r\n", "\r\n"), ("cr", "\r", "\r"), ("nolast", "\n", ""),
                            ("mixed", None, "")):
        if sep is None:
            data = body[0] + "\n" + body[1] + "\r\n" + body[2] + "\r" + body[3] + "\n" + body[4]
        else:
            data = sep.join(body) + tail
        p = write(tmp_path / name, data)
        want = oracle_read(p, ENT, REL)
        got = text.parse_triples(p, ENT, REL)
        assert got.dtype == np.int64 and np.array_equal(got, want), name
        assert text.count_lines(p) == len(body)
        assert np.array_equal(text.parse_triples(p, ENT, REL, n_threads=3), want)


def test_separators_are_those_of_str_split(tmp_path):
    seps = [" ", "\t", "\x0b", "\x0c", "\x1c", "\x1d", "\x1e", "\x1f", "\x85", "\xa0", "\u1680", "\u2000", "\u2005",
            "\u200a", "\u2028", "\u2029", "\u202f", "\u205f", "\u3000"]
    lines = ["%sa%slikes%s%sb%s" % (s, s, s, seps[(i + 1) % len(seps)], s) for i, s in enumerate(seps)]
    p = write(tmp_path / "seps", "\n".join(lines) + "\n")
    want = oracle_read(p, ENT, REL)
    assert len(want) == len(seps)
    assert np.array_equal(text.parse_triples(p, ENT, REL), want)
    # characters that are NOT separators stay inside the name: U+200B (zero width space), U+00AD, a lone 0xe2-lead
    for inside in ("\u200b", "\u00ad", "\u2060", "\u20ac", "\u180e", "\u0080"):
        ent = dict(ENT)
        ent["a" + inside + "b"] = 9
        p = write(tmp_path / "inside", "a%sb likes c\n" % inside)
        assert np.array_equal(text.parse_triples(p, ent, REL), oracle_read(p, ent, REL))
        with pytest.raises(KeyError):
            text.parse_triples(p, ENT, REL)


def test_empty_file_and_missing_file(tmp_path):
    p = write(tmp_path / "empty", "")
    got = text.parse_triples(p, ENT, REL)
    assert got.shape == (0, 3) and text.count_lines(p) == 0
    with pytest.raises(FileNotFoundError):
        text.parse_triples(str(tmp_path / "nope.txt"), ENT, REL)
    with pytest.raises(FileNotFoundError):
        O._read_triples(str(tmp_path / "nope.txt"), ENT, REL)
    with pytest.raises(FileNotFoundError):
        text.parse_triples(str(tmp_path), ENT, REL)          # a directory


@pytest.mark.parametrize("body,exc", [
    ("a likes b\n\nb likes c\n", ValueError),                # blank line: nothing to unpack
    ("a likes b\n   \n", ValueError),
    ("a likes\n", ValueError),
    ("a likes b c\n", ValueError),
    ("a likes b\nq likes b\n", KeyError),                    # unknown entity
    ("a hates b\n", KeyError),                               # unknown relation
    ("a likes likes\n", KeyError),                           # a relation name where an entity belongs
    ("likes a b\n", KeyError),
    ("two words likes a\n", ValueError),                     # a dictionary key with a blank can never match
    ("a likes b\nq likes b\na likes\n", KeyError),           # the FIRST bad line decides
    ("a likes b\na likes\nq likes b\n", ValueError),
    ("q q\n", ValueError),                                   # the split fails before any lookup
    ("a likes b\n\n", ValueError),                           # trailing blank line
])
def test_first_error_matches_the_reference_loop(tmp_path, body, exc):
    p = write(tmp_path / "bad", body)
    with pytest.raises(exc):
        O._read_triples(p, ENT, REL)
    with pytest.raises(exc):
        text.parse_triples(p, ENT, REL)
    with pytest.raises(exc):
        text.parse_triples(p, ENT, REL, n_threads=4)


def test_error_line_is_reported_through_the_c_abi(tmp_path):
    p = write(tmp_path / "bad", "a likes b\nb likes c\nq likes a\na likes\n")
    ent, rel = text.NameTable(ENT), text.NameTable(REL)
    out = np.full((4, 3), -7, dtype=np.int32)
    rows, bad = ctypes.c_int64(0), ctypes.c_int64(-1)
    rc = _lib.lib.rg_text_parse_triples(p.encode(), ctypes.byref(ent.struct), ctypes.byref(rel.struct),
                                        out.ctypes.data, 4, ctypes.byref(rows), ctypes.byref(bad), 1)
    assert rc == _lib.RG_ERR_UNKNOWN_NAME and bad.value == 2 and rows.value == 4
    assert out[:2].tolist() == [[0, 0, 1], [1, 0, 2]]
    assert b"missing" in _lib.lib.rg_strerror(rc) and b"three" in _lib.lib.rg_strerror(_lib.RG_ERR_PARSE)
    # a buffer smaller than the file is refused before anything is written, with the needed row count
    out[:] = -7
    rc = _lib.lib.rg_text_parse_triples(p.encode(), ctypes.byref(ent.struct), ctypes.byref(rel.struct),
                                        out.ctypes.data, 3, ctypes.byref(rows), ctypes.byref(bad), 1)
    assert rc == -1 and rows.value == 4 and (out == -7).all()
    assert _lib.lib.rg_text_parse_triples(None, None, None, None, 0, None, None, 0) == -1
    assert _lib.lib.rg_text_count_lines(None, None) == -1
    assert _lib.lib.rg_text_count_lines(b"/nonexistent/file", ctypes.byref(rows)) == _lib.RG_ERR_IO


def test_duplicate_names_later_entry_wins(tmp_path):
    # a dict cannot hold a name twice, the flat C table can: it behaves like repeated dict assignment
    blob = b"aba"
    off = np.array([0, 1, 2, 3], dtype=np.int64)
    ids = np.array([10, 11, 12], dtype=np.int32)
    ent = _lib.RgNameTable(ctypes.cast(ctypes.c_char_p(blob), ctypes.c_void_p), off.ctypes.data, ids.ctypes.data, 3)
    rel = text.NameTable({"r": 5})
    p = write(tmp_path / "dup", "a r b\n")
    out = np.zeros((1, 3), dtype=np.int32)
    rows, bad = ctypes.c_int64(0), ctypes.c_int64(-1)
    assert _lib.lib.rg_text_parse_triples(p.encode(), ctypes.byref(ent), ctypes.byref(rel.struct), out.ctypes.data, 1,
                                          ctypes.byref(rows), ctypes.byref(bad), 1) == 0
    assert out.tolist() == [[12, 5, 11]]


def test_large_file_threads_agree_and_match_the_oracle(tmp_path):
    rng = np.random.default_rng(5)
    n_ent, n_rel, n = 20000, 57, 300000
    ent = {"/m/%06x_%s" % (i * 7919 % 1000003, "é" * (i % 3)): i for i in range(n_ent)}
    rel = {"/rel/%d/%s" % (i, "x" * (i % 11)): i for i in range(n_rel)}
    en, rn = list(ent), list(rel)
    h, r, t = rng.integers(0, n_ent, n), rng.integers(0, n_rel, n), rng.integers(0, n_ent, n)
    seps = ["\t", " ", "  ", " \t"]
    lines = [en[a] + seps[i & 3] + rn[b] + seps[(i >> 2) & 3] + en[c] for i, (a, b, c) in enumerate(zip(h, r, t))]
    p = write(tmp_path / "big", "\n".join(lines) + "\n")
    assert os.path.getsize(p) > 4 << 20                       # several MB: the file is split over the threads
    want = np.stack([h, r, t], 1)
    assert np.array_equal(oracle_read(p, ent, rel)[:5000], want[:5000])
    tabs = (text.NameTable(ent), text.NameTable(rel))
    for threads in (1, 2, 5, 0):
        assert np.array_equal(text.parse_triples(p, *tabs, n_threads=threads), want), threads
    assert text.count_lines(p) == n
    # an error deep inside a late chunk is still reported as that line, not as an earlier chunk's
    lines[250001] = "nobody " + rn[0] + " " + en[0]
    lines[270000] = "broken line"
    p = write(tmp_path / "big_bad", "\n".join(lines))
    with pytest.raises(KeyError, match="line 250002"):
        text.parse_triples(p, *tabs, n_threads=8)


def test_bundled_and_synthetic_datasets_match_the_oracle_reader(tiny_dir, induc_dir):
    cases = [(tiny_dir, False), (induc_dir, True), (induc_dir + "_ind", True)]
    staged = os.path.join(ROOT, "oracle", "_ref", "Static")
    for sub, with_id in (("transductive/data/family", False), ("inductive/data/fb237_v2", True),
                         ("inductive/data/fb237_v2_ind", True)):
        if os.path.isdir(os.path.join(staged, sub)):
            cases.append((os.path.join(staged, sub), with_id))
    for d, with_id in cases:
        rel_dir = d[:-4] if d.endswith("_ind") else d         # the _ind split shares the relation dictionary
        ent = text.read_id_table(os.path.join(d, "entities.txt"), with_id)
        rel = text.read_id_table(os.path.join(rel_dir, "relations.txt"), with_id)
        assert ent == O._read_names(os.path.join(d, "entities.txt"), with_id)
        for f in ("facts.txt", "train.txt", "valid.txt", "test.txt"):
            p = os.path.join(d, f)
            if os.path.isfile(p):
                assert np.array_equal(text.parse_triples(p, ent, rel), oracle_read(p, ent, rel)), p


def test_filter_table_equals_the_per_triple_set_insertions():
    rng = np.random.default_rng(11)
    n_ent, n_rel = 50, 4
    a = np.stack([rng.integers(0, n_ent, 3000), rng.integers(0, 2 * n_rel, 3000), rng.integers(0, n_ent, 3000)], 1)
    want = defaultdict(set)
    for h, r, t in a.tolist():                                # load_data.py:64 / inductive :170-197
        want[(h, r)].add(t)
    got = text.filter_table([a[:1000], a[1000:]], n_ent)
    assert set(got) == set(want)
    assert all(isinstance(v, list) and len(v) == len(set(v)) and set(v) == want[k] for k, v in got.items())
    assert all(type(x) is int for k in got for x in k) and type(next(iter(got.values()))[0]) is int
    assert got[(np.int64(a[0, 0]), np.int64(a[0, 1]))] == got[(int(a[0, 0]), int(a[0, 1]))]   # numpy keys hit
    assert got[(10 ** 6, 0)] == set()                         # missing key: the reference's defaultdict(set)
    assert text.filter_table(np.zeros((0, 3), dtype=np.int64), n_ent) == {} and (3, 1) not in text.filter_table([], 5)
    assert len(got) == len(want) and list(got) == sorted(want) and (10 ** 6, 0) not in got and "x" not in got
    # a whole batch at once, unknown queries included
    subs, rels = np.r_[a[:7, 0], 10 ** 6], np.r_[a[:7, 1], 0]
    ptr, tails = got.rows(subs, rels)
    assert [tails[ptr[i]:ptr[i + 1]].tolist() for i in range(8)] == [sorted(want.get((int(s), int(r)), ()))
                                                                      for s, r in zip(subs, rels)]
    # ids too large for the packed 62-bit key take the lexsort route; same table
    big = a.copy()
    big[:, 0] += 2 ** 40
    big[:, 2] += 2 ** 30
    g2 = text.filter_table(big, n_ent)
    assert {(h - 2 ** 40, r): {t - 2 ** 30 for t in v} for (h, r), v in g2.items()} == dict(want)


def test_name_table_refuses_ids_beyond_32_bits():
    with pytest.raises(ValueError):
        text.NameTable({"a": 2 ** 31})


def test_fuzz_against_the_reference_loop(tmp_path):
    """Random files over an alphabet of names, separators, line endings and look-alikes: the native
    reader returns the oracle loop's array, or raises the exception type the loop raises."""
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies
    ent = {"a": 0, "b": 1, "ab": 2, "é": 3, "a\u200bb": 4}
    rel = {"r": 0, "a": 1}
    tabs = (text.NameTable(ent), text.NameTable(rel))
    alphabet = ["a", "b", "r", "é", "q", " ", " ", "\t", "\n", "\n", "\r", "\r\n", "\xa0", "\u2003", "\x1c", "\u200b",
                "\u2028", "\x0c"]
    path = str(tmp_path / "fuzz")

    def outcome(fn):
        try:
            return np.asarray(fn(), dtype=np.int64).reshape(-1, 3).tolist()
        except (ValueError, KeyError) as e:
            return type(e).__name__

    @hyp.settings(max_examples=400, deadline=None, database=None, derandomize=True,
                  suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(st.lists(st.sampled_from(alphabet), max_size=40), st.integers(1, 3))
    def run(parts, threads):
        write(path, "".join(parts))
        want = outcome(lambda: O._read_triples(path, ent, rel))
        assert outcome(lambda: text.parse_triples(path, *tabs, n_threads=threads)) == want

    run()
