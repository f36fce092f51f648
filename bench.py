#!/usr/bin/env python
"""bench.py — GN iterations/sec on BASELINE.json's 100k-state SE(3) GP trajectory (config C3) + linearise HBM GB/s.

  python bench.py --gpus N --steps K --warmup W            # CUDA engine (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (restated oracle) on the host cores

A "step" is one Gauss-Newton iteration of the hot path over the whole graph: batched linearise of every factor ->
normal-equation assembly -> bordered block-tridiagonal Cholesky solve -> retract -> error.  Prints ONE JSON line.
"""
import argparse
import datetime
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C3: SE(3) GP-prior + interpolated range factors, 100k states, 50k ranges, 16 landmarks (BASELINE.json configs[2]; interpolated range only - the reference has no interpolated bearing factor)"
METRIC = "GN iterations/sec on 100k-state SE(3) GP trajectory"
# CPU legs: the sample is the largest cut of C3 (same factor densities) whose (warm-up + timed) iterations fit CPU_BUDGET_S at
# ~30 us per state and iteration on one core - the whole 100k-state workload for the default 1 + 3 iterations (about 12 s),
# never less than 10k states
CPU_BUDGET_S = 120.0
CPU_US_PER_STATE = 30.0


def cpu_sample_states(iterations, full=100000):
    return int(max(10000, min(full, CPU_BUDGET_S / (max(1, iterations) * CPU_US_PER_STATE * 1e-6))))
# dram__bytes_read.sum + dram__bytes_write.sum of one launch on C3, from the ncu --set full capture of THIS round's build summarised in
# profiles/rd2w_ncu_full_summary.csv (ncu cannot run inside the timed bench, so these are constants tied to that capture):
# k_lin_gp 15.7 MB read + 182.0 MB written (the tail of the 240 MB of [A|b] is still in L2 at kernel end and is written back during
# the next kernel); k_panel0<12,4> (level 0) 261.4 MB read + 19.8 MB written (k_panel4 in round 1: 827.3 MB)
TRAFFIC_LIN_GP = 197.7e6
TRAFFIC_PANEL = 281.2e6
PANEL_EXECUTED_FRACTION = 0.58  # 11.38 M DMMA executed by k_panel0 on C3 (ncu source page, r2f) of the dense panel's 19.6 M
# algorithmic FLOPs of the level-0 panel per state (SE(3), w = 61 columns): Y = L^-1 P (12*13/2*61 MAC), P' = Le Y (12*12*61),
# S += Y^T Y (61*62/2*12)
PANEL_FLOP_PER_STATE = 2.0 * (78 * 61 + 144 * 61 + 61 * 62 // 2 * 12)


def po_threads():
    from oracle import pyoracle as po
    return po.hardware_threads()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms; started before the graph is built (nvidia-smi needs a second or
    more to come up on an 8-GPU box) and reduced to the samples whose timestamps fall inside the loaded window
    [mark_begin, mark_end] = warm-up + timed iterations + end-to-end leg + stage timings"""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = datetime.datetime.now()

    def mark_end(self):
        self.t1 = datetime.datetime.now()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(c[1]), float(c[2]), c[4:8]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        if not inside and rows and self.t0 is not None:  # window shorter than the sampling period: the nearest sample to it
            inside = [min(rows, key=lambda r: abs((r[0] - self.t0).total_seconds()))]
        for ts, s_, m_, flags in inside:
            sm.append(s_); mx.append(m_)
            for n, v in zip(names, flags):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        os.unlink(self.f.name)
        return out


def cpu_reference_run(steps, warmup, threads, n_sample=None, keep_first=False):
    """the reference's CPU path (oracle/: restated gpslam factors + GTSAM-style GN over a bordered block-tridiagonal Cholesky)
    on a bounded sample of the workload: C3 cut to n_sample states (same factor densities).  Per-iteration cost is linear in
    the number of states, so iterations/sec on the 100k-state graph = sample rate * n_sample / 100000."""
    from gpslam_b200 import synth
    from oracle import pyoracle as po
    cfg = synth.config("C3"); full = cfg.n_states
    if n_sample is None:
        n_sample = cpu_sample_states(steps + warmup, full)
    cfg.n_states = n_sample
    o, _ = synth.build(cfg, lambda grp, n, l: po.Graph(grp, n, l))
    o.set_threads(threads)
    first = None
    if warmup:
        if keep_first:  # the values after the FIRST Gauss-Newton iteration: what the engine's parity gate is compared with
            o.optimize(n_iter=1, use_lm=False)
            first = o.get_values()
            if warmup > 1:
                o.optimize(n_iter=warmup - 1, use_lm=False)
        else:
            o.optimize(n_iter=warmup, use_lm=False)
    t0 = time.perf_counter()
    st = o.optimize(n_iter=steps, use_lm=False)
    dt = time.perf_counter() - t0
    rate_sample = steps / dt
    return {"value": rate_sample * n_sample / full, "n_sample": n_sample, "first_iteration_values": first, "seconds_per_iteration_sample": dt / steps, "lin_seconds": st.lin_seconds, "solve_seconds": st.solve_seconds,
            "sample": ("the whole workload (C3, %d states), %d GN iterations after %d warm-up" % (full, steps, warmup)) if n_sample == full else
                      ("C3 cut to %d of %d states (same factor densities), %d GN iterations after %d warm-up; rate scaled by %d/%d (cost is linear in states)"
                       % (n_sample, full, steps, warmup, n_sample, full))}


def parity_after_one_gn(make_graph, n_states, oracle_values):
    """SURVEY.md §8(d) parity gate next to the throughput number: a fresh engine graph of the same workload, ONE Gauss-Newton
    iteration from the same initial values, compared with the oracle's values after its first iteration (largest absolute
    difference of the wire entries; the tests hold this to 1e-6, north_star's tolerance)"""
    from gpslam_b200 import synth
    cfg = synth.config("C3"); cfg.n_states = n_states
    g, _ = synth.build(cfg, make_graph)
    g.optimize(n_iter=1, use_lm=False)
    P, V, L = g.get_values()
    Po, Vo, Lo = oracle_values
    return {"after_gn_iterations": 1, "states": n_states, "max_pose_diff": float(np.abs(P - Po).max()), "max_velocity_diff": float(np.abs(V - Vo).max()),
            "max_landmark_diff": float(np.abs(L - Lo).max()) if np.size(Lo) else 0.0, "tolerance": 1e-6}


def parity_converged(make_graph, threads):
    """north_star's parity criterion beside the throughput number: the engine's and the oracle's LevenbergMarquardtOptimizer runs
    to convergence (GTSAM's stop rule, matlab/PlazaPose2.m:208-230) from the same initial values on the whole workload;
    per-state |Log(T_cpu^-1 T_gpu)|_inf, velocity and landmark differences of the two converged solutions"""
    from gpslam_b200 import synth
    from oracle import pyoracle as po
    from tests.test_gpu_fullsize import pose_error
    rec, _ = synth.record(synth.config("C3"))
    g = rec.replay(make_graph); o = rec.replay(lambda grp, n, l: po.Graph(grp, n, l))
    o.set_threads(threads)
    t0 = time.perf_counter(); sg = g.optimize(use_lm=True); tg = time.perf_counter() - t0
    t0 = time.perf_counter(); so = o.optimize(use_lm=True); to = time.perf_counter() - t0
    P, V, L = g.get_values(); Po, Vo, Lo = o.get_values()
    return {"after": "converged", "optimizer": "LevenbergMarquardt", "states": g.N, "iterations_engine": sg.iterations, "iterations_oracle": so.iterations,
            "error_engine": sg.error_final, "error_oracle": so.error_final, "max_pose_log_err": float(pose_error(0, Po, P).max()),
            "max_velocity_diff": float(np.abs(V - Vo).max()), "max_landmark_diff": float(np.abs(L - Lo).max()), "tolerance": 1e-6,
            "engine_seconds": tg, "oracle_seconds": to, "oracle_threads": threads}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import pyoracle as po
    threads = po.hardware_threads()
    r = cpu_reference_run(args.steps, args.warmup, threads)
    line = {"metric": METRIC, "value": r["value"], "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / r["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "impl": "reference", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": r["value"], "unit": "iterations/s", "cores": threads, "kind": "port", "sample": r["sample"],
                             "linearise_s_per_iteration_sample": r["lin_seconds"] / args.steps, "solve_s_per_iteration_sample": r["solve_seconds"] / args.steps,
                             "note": "linearise is threaded over factors; the bordered block-tridiagonal Cholesky of the restated reference is sequential scalar C++ (no Eigen / BLAS)"},
            "e2e": {"value": r["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def run_engine(args, rank, world, local_rank):
    import gpslam_b200 as gb
    from gpslam_b200 import shard, synth
    # GPB_BENCH_ONE_DEVICE=1: debugging aid - every rank on cuda:0 with the gloo backend, to walk the N-rank code path on a
    # one-GPU box (NCCL refuses two ranks on one device).  Its line is marked "debug_one_device" and is not a bench value.
    one_device = world > 1 and os.environ.get("GPB_BENCH_ONE_DEVICE") == "1"
    if one_device:
        local_rank = 0
    if gb.device_count() <= local_rank:
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        if one_device:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank)
    cfg = synth.config(args.config)
    if args.states:
        cfg.n_states = args.states
        if cfg.n_closures:
            cfg.closure_min_gap = min(cfg.closure_min_gap, max(2, cfg.n_states // 10))
    se3_wide = cfg.group == 0 and 11 <= cfg.n_landmarks <= 17   # SE(3) with a 64-column panel: the spine / panel stage timers exist
    t0 = time.perf_counter()
    if world > 1:
        # strong scaling: the SAME 100k-state graph, cut into contiguous segments, one per GPU; one NCCL all-reduce of the
        # boundary Schur system per GN iteration and no other collective on the data path
        g, _ = synth.build(cfg, lambda grp, n, l: shard.ShardBuilder(lambda g_, n_, l_: gb.Graph(g_, n_, l_), grp, n, l, rank, world), finalize=False)
        gen_s = time.perf_counter() - t0
        g.finalize(local_rank)
        g = g.g
        if one_device:
            g.set_allreduce(shard.torch_allreduce(local_rank))
        else:
            # the engine's own NCCL communicator: the all-reduce is enqueued by the engine inside the captured CUDA graph of an
            # iteration (no Python, no torch dispatch on the hot loop); torch.distributed only carries the 128-byte unique id
            shard.init_engine_nccl(g, rank, world)
    else:
        g, _ = synth.build(cfg, lambda grp, n, l: gb.Graph(grp, n, l), finalize=False)
        gen_s = time.perf_counter() - t0
        g.finalize(local_rank)
    build_s = time.perf_counter() - t0
    sz = g.sizes()
    err0 = g.linearize()

    def sync_all():
        if dist is not None:
            import torch
            dist.barrier(); torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up, then K timed GN iterations: CUDA events on the engine's stream, synchronised on both sides (gpb_optimize),
    #      barrier before, max over ranks after
    sampler.mark_begin()
    g.optimize(n_iter=max(args.warmup, 3), use_lm=False)
    sync_all()
    st = g.optimize(n_iter=args.steps, use_lm=False)
    sync_all()
    launches = g.launches()
    n_allreduce = g.allreduces()
    ms_per_step = max_over_ranks(st.total_ms) / args.steps
    value = 1e3 / ms_per_step
    # ---- end to end through the C ABI with host buffers: H2D of the values, one iteration, D2H of the result, every step
    #      (values live in page-locked host arrays, as a caller that cares about the transfer would hold them)
    #      through gpb_optimize_batch: step k+1's H2D and step k-1's D2H run on copy streams while step k computes.  Every step
    #      uploads its own input values (two page-locked input sets, alternating) and downloads its result and error.
    ins = [g.get_values(out=g.alloc_values()) for _ in range(2)]
    outs = [g.alloc_values() for _ in range(2)]
    P, V, Lm = ins[0]
    g.optimize_batch([ins[k & 1] for k in range(3)], [outs[k & 1] for k in range(3)])
    sync_all()
    t0 = time.perf_counter()
    st_b, errs_b = g.optimize_batch([ins[k & 1] for k in range(args.steps)], [outs[k & 1] for k in range(args.steps)])
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    io_bytes = int(P.nbytes + V.nbytes + Lm.nbytes)
    # the same steps one call at a time (Values in, iterate, Values out; nothing overlapped) for comparison
    for _ in range(2):
        g.set_values(P, V, Lm); g.optimize(n_iter=1, use_lm=False); g.get_values(out=outs[0])
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.set_values(P, V, Lm); g.optimize(n_iter=1, use_lm=False); g.get_values(out=outs[0])
    e2e_serial_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    # ---- per-stage device times and the linearise roofline (local shard)
    stages = {n: g.time_stage(k, 20) for k, n in ((0, "linearise_gp"), (1, "linearise_other"), (2, "assemble"), (3, "solve"), (4, "retract"), (5, "solve_fwd_level0"),
                                                   (6, "solve_spine_level0"), (7, "solve_panel_level0"), (8, "solve_backward"), (9, "linearise_all")) if se3_wide or k not in (6, 7)}
    from gpslam_b200 import capi
    dmma_peak = capi.dmma_peak(local_rank)
    # what the [A|b] store pattern costs with no arithmetic in front of it, and a plain memset of the same bytes (context for
    # the linearise roofline: the kernel is bound by its stores)
    store_floor_us = capi.store_peak(4, sz.n_gp, local_rank); memset_us = capi.store_peak(0, sz.n_gp, local_rank)
    sampler.mark_end()
    clocks = sampler.stop()
    peak, peak_src = peaks()
    D_ = 6 if cfg.group == 0 else 3
    SR_ = {0: 18, 1: 6, 2: 12, 3: 6}[cfg.group]
    gp_bytes = g.N * 8.0 * SR_ + sz.n_gp * (8.0 + 8.0 * 2 * D_ * (4 * D_ + 1))  # states once + per factor (param + [A|b])
    achieved = gp_bytes / (stages["linearise_gp"] * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD if args.config == "C3" and not args.states else "%s shape, %d states%s (bench.py --config / --states: not the headline workload)" % (args.config, cfg.n_states, ", %d loop closures" % cfg.n_closures if cfg.n_closures else ""),
                   "states": cfg.n_states, "states_per_gpu": g.N, "gp_factors_rank0": sz.n_gp, "other_factors_rank0": sz.n_extra,
                   "landmark_dims": sz.border_dim, "solver_levels": sz.levels, "optimizer": "Gauss-Newton",
                   "parallelism": "trajectory segments x%d, one NCCL all-reduce of the boundary Schur system per iteration" % world if world > 1 else "single GPU",
                   "allreduces_per_step": (n_allreduce - 1) / args.steps if world > 1 else 0,
                   "l2": "inputs larger than L2: [A|b] buffers 2 x %.0f MB, resident %.0f MB per GPU (L2 126 MB)" % (sz.n_gp * 2400 / 1e6, sz.hbm_bytes / 1e6),
                   "hbm_resident_mb": sz.hbm_bytes / 1e6, "graph_build_s": build_s, "generator_s": gen_s, "finalize_s": build_s - gen_s},
        "e2e": {"value": 1.0 / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": io_bytes * world, "d2h_bytes_per_step": io_bytes * world + 8 * world,
                "api": "gpb_optimize_batch (host values in -> one GN iteration -> host values + error out, every step; copies double-buffered against compute)",
                "unpipelined_value": 1.0 / e2e_serial_s, "unpipelined_api": "gpb_set_values + gpb_optimize(1) + gpb_get_values per step"},
        "gpu_launches": launches,
        "roofline": {"kernel": "k_lin_gp<%s> (batched GP-prior linearise%s)" % ({0: "POSE3", 1: "POSE2", 2: "ROT3", 3: "LINEAR"}[cfg.group], "; diagonal-Qc instantiation" if cfg.group == 0 else ""), "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "algorithmic_bytes": gp_bytes, "ms": stages["linearise_gp"], "traffic": TRAFFIC_LIN_GP if world == 1 and args.config == "C3" and not args.states else None,
                     "store_pattern_floor_ms": store_floor_us * 1e-3, "memset_same_bytes_ms": memset_us * 1e-3,
                     # every factor of the graph (SURVEY §8d LINEARISE_BYTES incl. the measurement rows), timed as the iteration runs it:
                     # k_lin_gp on the main stream, the other factors beside it on the side stream, joined before the error reduction
                     "whole_linearise": {"algorithmic_bytes": sz.linearise_bytes, "ms": stages["linearise_all"],
                                         "achieved": sz.linearise_bytes / (stages["linearise_all"] * 1e-3) / 1e9, "frac": sz.linearise_bytes / (stages["linearise_all"] * 1e-3) / 1e9 / peak,
                                         "streams": "overlapped (two streams, fork / join inside the stage)"}},
        # the kernel that dominates the iteration by time: the level-0 panel, bound by the FP64 tensor pipe
        "stages_ms": stages, "clocks": clocks, "error": {"initial": err0, "final": st.error_final},
    }
    if se3_wide:
        # the kernel that dominates the iteration by time: the level-0 panel.  `achieved` counts the DENSE panel's flops (every one of
        # the 61 columns at every state: what the elimination order costs on paper and what k_panel4 executed in round 1); k_panel0
        # skips the column tiles that are still exactly zero, so it executes fewer (executed_fraction, from the ncu DMMA count)
        pf = PANEL_FLOP_PER_STATE * g.N
        line["roofline_solver"] = {"kernel": "k_panel0<12,4> (level-0 panel on the active columns: Y = L^-1 P, P' = -Le Y, S += Y^T Y on mma.sync.m8n8k4.f64)", "bound": "tensor",
                                   "achieved": pf / (stages["solve_panel_level0"] * 1e-3) / 1e12, "peak": dmma_peak, "unit": "TFLOP/s",
                                   "frac": pf / (stages["solve_panel_level0"] * 1e-3) / 1e12 / dmma_peak,
                                   "peak_source": "measured in this run: FP64 mma.sync m8n8k4 issue loop on all SMs (gpb_debug_dmma_peak); MEASURED_PEAKS.json has no FP64 figure",
                                   "algorithmic_flops": pf, "executed_fraction": PANEL_EXECUTED_FRACTION if args.config == "C3" and not args.states and world == 1 else None,
                                   "ms": stages["solve_panel_level0"], "traffic": TRAFFIC_PANEL if world == 1 and args.config == "C3" and not args.states else None}
    if rank == 0 and world == 1 and not args.no_cpu and args.config == "C3" and not args.states:
        r = cpu_reference_run(3, 1, 1, keep_first=True)
        line["cpu_baseline"] = {"value": r["value"], "unit": "iterations/s", "cores": 1, "kind": "port", "sample": r["sample"],
                                "seconds_per_iteration_sample": r["seconds_per_iteration_sample"],
                                "linearise_s_per_iteration_sample": r["lin_seconds"] / 3, "solve_s_per_iteration_sample": r["solve_seconds"] / 3}
        try:  # reported beside the numbers, never allowed to cost them
            line["parity"] = parity_after_one_gn(lambda grp, n, l: gb.Graph(grp, n, l), r["n_sample"], r["first_iteration_values"])
            if not args.no_converged:
                line["parity"].update(parity_converged(lambda grp, n, l: gb.Graph(grp, n, l), po_threads()))
        except Exception as e:  # noqa: BLE001
            line["parity"] = dict(line.get("parity") or {}, error="%s: %s" % (type(e).__name__, e))
    if one_device:
        line["debug_one_device"] = True
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE JSON line: everything else a library prints (NCCL's version banner, torchrun notices) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--states", type=int, default=0, help="override the number of states (parity/debug runs; not a bench value)")
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4", "C5"], help="BASELINE.json config shape (default C3, the headline workload; others are side runs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-converged", action="store_true", help="skip the converged-LM parity run against the oracle (about 1.5 minutes of host time)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_engine(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
