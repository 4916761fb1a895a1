"""ncu logs of scratch/ncu_traffic.sh -> profiles/edge_fwd_traffic.json (+ copies of the logs under profiles/r02/)."""
import collections, csv, hashlib, json, os, shutil, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(root, "gpurun_out", sys.argv[1])
bid = hashlib.sha256(open(root + "/redgnn_b200/csrc/rg_edge.cu", "rb").read()).hexdigest()[:12]
out = {}
for w, n_layer in (("fb15k237", 4), ("yago310", 5), ("powerlaw", 6)):
    rows = list(csv.reader(open(os.path.join(src, "edge_fwd_%s.csv" % w))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hi], rows[hi + 1:]
    ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    launches = collections.OrderedDict()
    for r in data:
        d = launches.setdefault(r[0], {"kernel": r[ki].split("(")[0]})
        v = float(r[vi].replace(",", ""))
        if r[mi].startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ui]]
        if r[mi] == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1e-3)
        d[r[mi]] = v
    main = [d for d in launches.values() if "chunks" not in d["kernel"]]
    last = main[-n_layer:]
    tr = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in last) / len(last)
    out[w] = {"kernel_build_id": bid, "kernels": sorted(set(d["kernel"] for d in last)), "traffic_bytes_per_launch": tr,
              "launches": [{"layer": i, "duration_ms": round(d["gpu__time_duration.sum"], 4), "dram_read": d["dram__bytes_read.sum"],
                            "dram_write": d["dram__bytes_write.sum"], "l2_hit_pct": round(d["lts__t_sector_hit_rate.pct"], 1),
                            "l1_wavefront_pct": round(d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"], 1),
                            "ipc": round(d["sm__inst_executed.avg.per_cycle_active"], 2)} for i, d in enumerate(last)],
              "source": "profiles/r02/ncu_edge_fwd_%s.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... --clock-control none "
                        "-k regex:k_edge_fwd, python bench.py --workload %s --steps 1 --warmup 1 --no-cpu-baseline --no-train-subsystem); "
                        "mean over the %d launches (layers) of the last forward of the run" % (w, w, n_layer)}
    shutil.copy(os.path.join(src, "edge_fwd_%s.csv" % w), os.path.join(root, "profiles", "r02", "ncu_edge_fwd_%s.csv" % w))
    print(w, "traffic/launch %.1f MB" % (tr / 1e6), [(x["duration_ms"], x["l2_hit_pct"], x["l1_wavefront_pct"], x["ipc"]) for x in out[w]["launches"]])
json.dump(out, open(root + "/profiles/edge_fwd_traffic.json", "w"), indent=1)
