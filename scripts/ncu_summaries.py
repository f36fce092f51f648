#!/usr/bin/env python
"""Turns the two ncu captures of a round into the summaries committed under profiles/.

  python scripts/ncu_summaries.py <tag>      # reads gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_full.ncu-rep

<tag>_launch_shares.txt : per-kernel launch count, total / average duration and share of the summed kernel time (launch list of
                          `ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)
<tag>_ncu_full_summary.csv : per-launch duration, DRAM bytes, DRAM / SM / tensor / FP64 utilisation, registers, occupancy of the
                          `ncu --set full` capture (read with `ncu -i ... --page raw --csv`)."""
import collections
import csv
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = os.path.join(ROOT, "gpurun_out")
prof = os.path.join(ROOT, "profiles")

launches = os.path.join(out, tag + "_launches.csv")
if os.path.exists(launches):
    rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
    hdr = rows[0]
    iK, iV, iM = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        if r[iM] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[iK]).strip()
        us = float(r[iV].replace(",", "")) / 1e3
        t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += us
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(prof, tag + "_launch_shares.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 300 python bench.py --steps 2 --warmup 1 --no-cpu   (C3 100k states, 1xB200)\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES\n")
        for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("%-45s launches=%4d total_us=%10.1f avg_us=%8.1f share=%5.1f%%\n" % (name, n, us, us / n, 100 * us / total))
    shutil.copy(launches, os.path.join(prof, tag + "_launches.csv"))

rep = os.path.join(out, tag + "_full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers", "sm__cycles_elapsed.max"]
    idx = [hdr.index(k) for k in keep if k in hdr]
    with open(os.path.join(prof, tag + "_ncu_full_summary.csv"), "w", newline="") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:k_lin_gp|k_panel4|k_spine|k_assemble|k_lin_extra|k_bwd|k_level_ws|k_small -s 51 -c 17 "
                "python bench.py --steps 1 --warmup 3 --no-cpu   (one GN iteration of C3, 100k SE(3) states, 1xB200; per launch)\n")
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
print("wrote", [p for p in os.listdir(prof) if p.startswith(tag)])
