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
    (launches if kind == "launches" else metrics)(src, os.path.join(ROOT, "profiles", dst))
