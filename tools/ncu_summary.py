#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small committed files under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches_r01.csv profiles/r01_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof_r01.ncu-rep profiles/r01_top_kernels.csv
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
           "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
           "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# source: {src}; total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
        f.write(f"{'kernel':48s} {'launches':>8s} {'total_us':>12s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:48s} {v[0]:8d} {v[1]:12.1f} {100 * v[1] / tot:6.1f}%\n")
    print(open(dst).read())


def _raw(src):
    """raw page of a .ncu-rep, or a CSV already exported with `ncu -i x.ncu-rep --page raw --csv` (large reports stay on the GPU box)"""
    if src.endswith(".csv"):
        return open(src).read()
    return subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout


def full(src, dst):
    raw = _raw(src)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([m for m, _ in idx])
        w.writerow([units[i] for _, i in idx])
        for r in rows[2:]:
            w.writerow([re.sub(r"\(.*", "", r[i]) if m == "Kernel Name" else r[i] for m, i in idx])
    print(open(dst).read()[:3000])


def traffic(src, dst, note=""):
    """per-kernel DRAM traffic (read + write) averaged over the captured launches -> JSON read by bench.py"""
    import json
    raw = _raw(src)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ik, ir, iw, it = (hdr.index(m) for m in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ik])
        a = agg[name]
        a[0] += 1
        a[1] += float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
        a[2] += float(r[it].replace(",", "")) * tscale[units[it]]
    out = {k: {"launches_captured": v[0], "dram_bytes_per_launch": v[1] / v[0], "ms_per_launch_under_ncu": v[2] / v[0]} for k, v in agg.items()}
    out["_note"] = note
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
