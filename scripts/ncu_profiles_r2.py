#!/usr/bin/env python
"""Turns the round-2 ncu captures (gpurun_out/prof_r2_*.ncu-rep, launches_r2_*.csv) into the committed
summaries under profiles/: ncu_r2_key_metrics.json (per captured launch) and traffic_r2.json (DRAM bytes per
launch of the dominant kernels, tagged with the hash of the kernel sources they were captured with --
bench.py only reports `traffic` when that tag matches the sources it runs)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (kernel_source_tag)

KEYS = {
    "gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "registers", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_instruction",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "sm__inst_executed.sum": "warp_instructions", "launch__grid_size": "grid",
}
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "msecond": 1e3, "usecond": 1.0, "second": 1e6, "nsecond": 1e-3,
        "ms": 1e3, "us": 1.0, "s": 1e6, "ns": 1e-3}


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    res = []
    for r in rows[2:]:
        d = {"kernel": r[ki].split("(")[0].replace("void ", "").replace("rtb::", "")}
        for i, n in enumerate(h):
            if n in KEYS and r[i] != "":
                v = float(r[i].replace(",", ""))
                d[KEYS[n]] = v * UNIT.get(units[i], 1.0) if KEYS[n] in ("dram_read", "dram_write", "duration_us") else v
        res.append(d)
    return res


def main():
    gp = os.path.join(ROOT, "gpurun_out")
    out = {"kernel_source_tag": bench.kernel_source_tag(),
           "how": "ncu --set full --clock-control none (scripts/ncu_round2.sh, scripts/gpu_round2_f.sh) on one B200; each capture directly after the same command ran clean without ncu",
           "captures": {}}
    for name in ("prof_r2_mixed100k", "prof_r2_soup1m", "prof_r2_mixed100k_stage21"):
        p = os.path.join(gp, name + ".ncu-rep")
        if os.path.exists(p):
            out["captures"][name] = report(p)
    with open(os.path.join(ROOT, "profiles", "ncu_r2_key_metrics.json"), "w") as f:
        json.dump(out, f, indent=1)
    traffic = {"kernel_source_tag": out["kernel_source_tag"], "source": "profiles/ncu_r2_key_metrics.json (dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full)"}
    for wl, cap in (("mixed100k", "prof_r2_mixed100k"), ("soup1m", "prof_r2_soup1m")):
        ks = [k for k in out["captures"].get(cap, []) if "trace" in k["kernel"] or "shadow" in k["kernel"]]
        if ks:
            dom = max(ks, key=lambda k: k.get("duration_us", 0))
            traffic[wl] = {"kernel": dom["kernel"], "dram_bytes_per_launch": int(dom.get("dram_read", 0) + dom.get("dram_write", 0)),
                           "duration_us": dom.get("duration_us")}
    with open(os.path.join(ROOT, "profiles", "traffic_r2.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    for f in ("launches_r2_mixed100k.csv",):
        if os.path.exists(os.path.join(gp, f)):
            shutil.copy(os.path.join(gp, f), os.path.join(ROOT, "profiles", f))
    for cap, ks in out["captures"].items():
        for k in ks:
            print(f"{cap:28s} {k['kernel']:28s} {k.get('duration_us', 0):9.1f} us  lanes {k.get('lanes_per_instruction', 0):5.2f}  issue {k.get('issue_slot_pct', 0):5.1f}%  "
                  f"L1 {k.get('l1_hit_pct', 0):5.1f}%  L2 {k.get('l2_hit_pct', 0):5.1f}%  long_sb {k.get('stall_long_scoreboard', 0):5.2f}  dram {int(k.get('dram_read', 0) + k.get('dram_write', 0)) / 1e6:8.1f} MB  regs {int(k.get('registers', 0))}")


if __name__ == "__main__":
    main()
