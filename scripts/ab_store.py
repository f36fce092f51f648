#!/usr/bin/env python
"""Floor of k_lin_gp's store pattern (gpb_debug_store_peak): microseconds to write 99 999 SE(3) [A|b] records (240 MB)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpslam_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 99999
names = {0: "cudaMemsetAsync", 1: "row pairs NFp*16 B apart, 128-thread CTAs", 2: "same, st.global.cs", 3: "same, 256-thread CTAs", 4: "tiled layout (ab_off), 128-thread CTAs"}
for mode in (0, 1, 2, 3, 4):
    us = capi.store_peak(mode, n)
    print(json.dumps({"mode": names[mode], "us": round(us, 2), "GB_per_s": round(2400.0 * n / us / 1e3, 1)}), flush=True)
