"""gpslam_b200 — B200-native sparse-GP factor-graph linearise-and-solve engine (gpslam-compatible factor surface).

Python host mirror over the C ABI (include/gpb.h, gpslam_b200/libgpb.so).  The library is CUDA-only: there is no CPU
fallback, and importing `Graph` / calling `lib()` fails loudly when libgpb.so is missing.
"""
from .capi import (GPB_LINEAR, GPB_POSE2, GPB_POSE3, GPB_POSE3VW, GPB_ROT3, Graph, Params, Stats, build_library, default_params, device_count, lib,
                   library_path)

__all__ = ["Graph", "Params", "Stats", "default_params", "device_count", "lib", "build_library", "library_path", "GPB_POSE3", "GPB_POSE2",
           "GPB_ROT3", "GPB_LINEAR", "GPB_POSE3VW"]
