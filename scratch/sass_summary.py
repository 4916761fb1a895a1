"""Per-kernel SASS evidence of the built library: tensor-core / TMEM / TMA / atomic mnemonic counts from
`cuobjdump -sass` and registers / spills / shared memory from the ptxas logs the Makefile keeps.
    python scratch/sass_summary.py > profiles/r02/sass_summary.txt        (CPU only)"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "redgnn_b200", "libredgnn_b200.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "REDG", "RED.", "ATOMG", "ATOMS", "LDG.E.128",
        "STG.E.128", "SHFL", "VOTE", "MUFU", "STL", "LDL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for s in out:
        s = re.sub(r"\(anonymous namespace\)::", "", s)
        s = re.sub(r"^void ", "", s)
        short.append(re.sub(r"\(.*$", "", s))
    return dict(zip(names, short))


sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, order, cur = collections.defaultdict(collections.Counter), [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        order.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_n"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1

res = {}
for log in glob.glob(os.path.join(ROOT, "redgnn_b200", "csrc", "*.ptxas.log")):
    name = None
    for line in open(log):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
        m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?", line)
        if m and name:
            res.setdefault(name, {})["regs"] = int(m.group(1))
        m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and name:
            res.setdefault(name, {})["spill"] = (int(m.group(1)), int(m.group(2)))
        m = re.search(r"(\d+) bytes smem", line)
        if m and name:
            res.setdefault(name, {})["smem"] = int(m.group(1))

names = demangle(order)
print("libredgnn_b200.so (sm_100a): per kernel -- SASS instructions, registers, static smem, spills, then the counts of")
print("the mnemonics that matter (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,")
print("UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops, REDG = global reduction atomics, STL / LDL = local memory)")
print()
for fn in sorted(order, key=lambda f: names[f]):
    c, r = counts[fn], res.get(fn, {})
    keys = " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])
    print("%-58s insts=%-5d regs=%-3s smem=%-6s spill=%s  %s" % (names[fn][:58], c["_n"], r.get("regs", "?"), r.get("smem", 0),
                                                               "%d/%d" % r["spill"] if "spill" in r else "?", keys))
