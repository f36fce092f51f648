#!/usr/bin/env python
"""Per-kernel stall-reason and instruction-mix summary from the source page of an `ncu --set full --import-source on` capture.

  python scripts/ncu_stalls.py <tag> [kernel-regex ...]     # reads gpurun_out/<tag>_full.ncu-rep, writes profiles/<tag>_stall_breakdown.txt
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
kernels = sys.argv[2:] or ["k_lin_gp", "k_assemble_mma", "k_spine", "k_panel4", "k_level_ws", "k_bwd", "k_small_solve", "k_lin_extra"]
rep = os.path.join(ROOT, "gpurun_out", tag + "_full.ncu-rep")
out = open(os.path.join(ROOT, "profiles", tag + "_stall_breakdown.txt"), "w")
out.write("# ncu -i gpurun_out/%s_full.ncu-rep --page source --csv --kernel-name regex:<kernel>   (first launch of each kernel in the capture;\n" % tag)
out.write("# warp-stall samples by reason, then the opcodes that executed most with their share of the samples)\n")
for kn in kernels:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kn], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    name = rows[0][1] if len(rows[0]) > 1 else kn
    hdr = rows[1]
    iS, iA, iE = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    seen, tot, ex, st, n = set(), collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= max(iA, iE) or len(r) < len(hdr) or r[0] in seen:   # the page lists every instruction twice (SASS and source-correlated view)
            continue
        seen.add(r[0])
        toks = r[iS].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
        try:
            a, e = float(r[iA] or 0), float(r[iE] or 0)
        except ValueError:
            continue
        ex[op] += e; st[op] += a; n[op] += 1
        for h in reasons:
            try:
                tot[h] += float(r[hdr.index(h)] or 0)
            except ValueError:
                pass
    S, E = sum(st.values()), sum(ex.values())
    if S == 0:
        continue
    out.write("\n%s\n  static instructions %d, executed warp-instructions %.0f, samples %.0f\n" % (name[:110], sum(n.values()), E, S))
    out.write("  stalls: " + ", ".join("%s %.1f%%" % (h.replace("stall_", ""), 100 * v / max(1.0, sum(tot.values()))) for h, v in tot.most_common(8)) + "\n")
    out.write("  opcodes: " + ", ".join("%s %.1f%% exec / %.1f%% samples" % (k, 100 * v / E, 100 * st[k] / S) for k, v in ex.most_common(9)) + "\n")
out.close()
print(open(os.path.join(ROOT, "profiles", tag + "_stall_breakdown.txt")).read())
