#!/usr/bin/env python
"""ncu report of tests/tools/profile_workload.py <workload>  ->  profiles/traffic.json entry (the counters bench.py's
roofline reads) + a key-metrics csv under profiles/.
    python tests/tools/ncu_traffic.py <workload> gpurun_out/prof_<workload>.ncu-rep [tag]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
from ncu_summary import KEYS  # noqa: E402


def main():
    workload, rep = sys.argv[1], sys.argv[2]
    tag = sys.argv[3] if len(sys.argv) > 3 else "r02"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")

    def val(r, key):
        return float(r[hdr.index(key)].replace(",", "")) if key in hdr and r[hdr.index(key)] not in ("", "n/a") else None

    def unit_scale(key):  # ncu reports bytes in the unit of its choice
        u = units[hdr.index(key)].lower()
        return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)

    path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(path)) if os.path.exists(path) else {}
    entry = traffic.setdefault(workload, {})
    for r in data:
        kname = r[ki]
        # trace_kernel<(int)MODE, (int)VARIANT>: MODE 1 = primary, 2 = shadow
        mode = kname.split("trace_kernel<")[1].split(">")[0] if "trace_kernel<" in kname else ""
        which = "primary" if mode.replace("(int)", "").replace(" ", "").startswith("1,") else "shadow"
        dram = sum((val(r, k) or 0.0) * unit_scale(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        entry[f"trace_kernel<{which}>"] = {
            "dram_bytes": dram, "warp_inst": val(r, "smsp__inst_executed.sum"),
            "lanes_per_inst": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "l1_hit_pct": val(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers": val(r, "launch__registers_per_thread"), "ncu_time_ms": (val(r, "gpu__time_duration.sum") or 0.0) *
            {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")].lower(), 1.0),
            "kernel": kname.split("(")[0].strip()[-60:], "source": f"profiles/{tag}_ncu_{workload}_key_metrics.csv"}
    json.dump(traffic, open(path, "w"), indent=1)
    out = [["metric", "unit"] + [r[ki].split("(dodrt")[0][-44:] for r in data]]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out.append([k, units[i]] + [r[i] for r in data])
    csv.writer(open(os.path.join(ROOT, "profiles", f"{tag}_ncu_{workload}_key_metrics.csv"), "w")).writerows(out)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
