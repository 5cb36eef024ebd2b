#!/usr/bin/env python
"""Turn the scratch ncu outputs of a GPU session into the committed summaries under profiles/.

    python tools/ncu_summary.py <tag>

reads   gpurun_out/prof_<tag>.ncu-rep      (ncu --set full of the decode and post kernels)
        gpurun_out/launches_<tag>.csv      (ncu --metrics gpu__time_duration.sum launch list)
writes  profiles/<tag>_full.csv            one row per captured launch, the counters the roofline
                                           discussion in DESIGN.md cites
        profiles/<tag>_launches.csv        per kernel: launches, mean us, share of the library time
        profiles/traffic.json              dram bytes per launch of the dominant kernel; bench.py
                                           copies it into roofline.traffic
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = ("track_step_kernel", "zone_step_kernel", "step_kernel", "letterbox_decimate3_wide_kernel", "letterbox_tile_kernel", "letterbox_decimate3_kernel", "decode_tma_kernel", "decode_ldg_kernel", "post_kernel", "nms_kernel",
        "letterbox_kernel", "pred_candidates_kernel", "decode_head_kernel", "kalman", "lapjv")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warp_latency_issue_stalled_barrier.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def short(name):
    for k in OURS:
        if k in name:
            return k
    return name.split("(")[0][-60:]


def full(tag):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    old = json.load(open(tpath)) if os.path.exists(tpath) else {}
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    out = os.path.join(ROOT, "profiles", f"{tag}_full.csv")
    traffic = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{m} [{units[i]}]" for m, i in cols])
        for r in data:
            w.writerow([short(r[ki])] + [r[i] for _, i in cols])
            k = short(r[ki])
            rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            traffic.setdefault(k, []).append((rd, wr))
    summary = {k: {"dram_bytes_read_per_launch": sum(a for a, _ in v) / len(v),
                   "dram_bytes_write_per_launch": sum(b for _, b in v) / len(v),
                   "launches_captured": len(v)} for k, v in traffic.items()}
    for v in summary.values():
        v["source"] = f"ncu --set full --clock-control none, gpurun_out/prof_{tag}.ncu-rep -> profiles/{tag}_full.csv"
    old.pop("source", None)
    old.update(summary)                                  # kernels captured in other sessions keep their entries
    with open(tpath, "w") as f:
        json.dump(old, f, indent=1)
    print("wrote", out)


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)  # -> us
        agg.setdefault(short(r[ki]), []).append(v)
    ours = {k: v for k, v in agg.items() if any(o in k for o in OURS)}
    lib_total = sum(sum(v) for v in ours.values())
    all_total = sum(sum(v) for v in agg.values())
    out = os.path.join(ROOT, "profiles", f"{tag}_launches.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "avg_us", "share_of_library_time", "share_of_all_gpu_time", "ours"])
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            w.writerow([k, len(v), f"{sum(v) / len(v):.2f}", f"{sum(v) / lib_total:.3f}" if k in ours else "",
                        f"{sum(v) / all_total:.3f}", int(k in ours)])
    print("wrote", out)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    full(tag)
    launches(tag)
