"""CPU: bench.py's reference arm prints exactly one JSON line with the contract's keys (the CUDA arm
needs a GPU; its line is checked by tests/test_bench_gpu.py)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_on_tiny_workload():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "2", "--warmup", "1", "--cpu-queries", "2"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "queries/s" and d["unit"] == "queries/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    from oracle import ref_import as R
    # the reference's own files when mounted / staged (oracle/_ref), else the oracle port
    assert cb["kind"] == ("reference" if R.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["config"]["queries_per_step"] == 32 and cb["queries_per_cpu_step"] == 2
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--gpus", "2"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_does_not_load_the_cuda_library():
    """The CPU arm must not even import the product package (driver check: native_so_loaded empty)."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'tiny', '--steps', '1', "
            "'--warmup', '1', '--cpu-queries', '1']; runpy.run_path(%r, run_name='__main__'); "
            "assert not any(m.startswith('redgnn_b200') for m in sys.modules), 'product package imported'; "
            "maps = open('/proc/self/maps').read(); assert 'libredgnn_b200' not in maps, 'CUDA library mapped'"
            % os.path.join(ROOT, "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
