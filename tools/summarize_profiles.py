"""Condense ncu outputs from gpurun_out/ into small, committed summaries under profiles/.

  launch lists (ncu --metrics gpu__time_duration.sum): one training step, aggregated per kernel -> profiles/<name>_step.csv
  ncu --set full raw pages (ncu -i X.ncu-rep --page raw --csv): key metrics per kernel        -> profiles/<name>_metrics.csv
"""
import csv, collections, re, sys, os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def short(nm):
    m = re.search(r"(\w+_kernel)(<[^>]*>)?", nm)
    return (m.group(1) + (m.group(2) or "")) if m else re.sub(r"\(.*", "", nm)[:70]


def launches(src, dst):
    with open(src) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
    names = [r["Kernel Name"] for r in rows]
    vals = [float(r["Metric Value"].replace(",", "")) for r in rows]
    idx = [i for i, n in enumerate(names) if "stem_im2col" in n]
    a, b = idx[1], idx[2]  # the timed step
    agg = collections.OrderedDict()
    for i in range(a, b):
        k = short(names[i])
        e = agg.setdefault(k, [0, 0.0])
        e[0] += 1
        e[1] += vals[i]
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("kernel,launches_per_step,total_us,share_of_step\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{v[0]},{v[1]/1e3:.1f},{v[1]/tot:.4f}\n")
        f.write(f"\"TOTAL (serialised, cold cache)\",{b-a},{tot/1e3:.1f},1.0\n")


def launches_dram(src, dst, algo_gb=None):
    """ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum launch list -> per-kernel time AND DRAM traffic of one
    training step, plus the step-level DRAM / algorithmic-bytes ratio (bench.py's roofline.traffic)."""
    with open(src) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
    recs = collections.OrderedDict()
    for r in rows:
        d = recs.setdefault(r["ID"], {"name": r["Kernel Name"], "us": 0.0, "rd": 0.0, "wr": 0.0})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u, m = r["Metric Unit"], r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if u.startswith("n") else (v if u.startswith("u") else v * 1e3)
        elif m.startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            d["rd" if "read" in m else "wr"] = v * mult
    L = list(recs.values())
    idx = [i for i, d in enumerate(L) if "stem_im2col" in d["name"]]
    a, b = idx[0], idx[1]
    agg = collections.OrderedDict()
    for d in L[a:b]:
        e = agg.setdefault(short(d["name"]), [0, 0.0, 0.0, 0.0])
        e[0] += 1
        e[1] += d["us"]
        e[2] += d["rd"]
        e[3] += d["wr"]
    tot = sum(v[1] for v in agg.values())
    dram = sum(v[2] + v[3] for v in agg.values())
    with open(dst, "w") as f:
        f.write("kernel,launches_per_step,total_us,share_of_step,dram_read_MB,dram_write_MB,dram_GBps\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{v[0]},{v[1]:.1f},{v[1]/tot:.4f},{v[2]/1e6:.1f},{v[3]/1e6:.1f},{(v[2]+v[3])/max(v[1],1e-9)/1e3:.0f}\n")
        f.write(f"\"TOTAL (serialised, cold cache)\",{b-a},{tot:.1f},1.0,{sum(v[2] for v in agg.values())/1e6:.1f},{sum(v[3] for v in agg.values())/1e6:.1f},{dram/tot/1e3:.0f}\n")
        if algo_gb:
            f.write(f"\"STEP dram bytes / algorithmic bytes ({algo_gb} GB = 190.8 MB/img x 128)\",,,,,,{dram/1e9/algo_gb:.3f}\n")
    return dram


def metrics(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("kernel," + ",".join(f"{k} [{units[idx[k]]}]" for k in KEYS if k in idx) + "\n")
        for r in rows[2:]:
            f.write("\"" + short(r[idx["Kernel Name"]]) + "\"," + ",".join(r[idx[k]].replace(",", "") for k in KEYS if k in idx) + "\n")


if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    if kind == "launches_dram":
        launches_dram(src, os.path.join(ROOT, "profiles", dst), float(sys.argv[4]) if len(sys.argv) > 4 else None)
    else:
        (launches if kind == "launches" else metrics)(src, os.path.join(ROOT, "profiles", dst))
