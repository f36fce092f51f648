#!/usr/bin/env python
"""Value-level check of the engine-owned NCCL path (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/nccl_check.py

Every rank builds its shard of the same graph (C3 / C5 shapes, loop closures across shards included), the engine creates its own
NCCL communicator (gpb_graph_init_nccl) and runs (i) fixed-count Gauss-Newton - ONE CUDA graph per iteration with the all-reduce
captured inside - and (ii) LM to convergence.  Rank 0 also solves the unsharded graph on its GPU and runs the CPU oracle; the
gathered sharded solution must match both (1e-9 / 1e-6), every halo copy must be bit-identical to its owner's state, and the
number of all-reduces must be exactly one per GN iteration (+ the two 4-double error reports)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gpslam_b200 as gb  # noqa: E402
from gpslam_b200 import shard, synth  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [("C3", 4000, dict(n_landmarks=16, prior_every=50), 3), ("C5", 6000, dict(n_landmarks=16, prior_every=50, n_closures=6, closure_min_gap=1500), 3),
             ("C4", 5000, dict(), 3)]
    for name, n, kw, iters in cases:
        cfg = synth.config(name); cfg.n_states = n
        for k, v in kw.items():
            setattr(cfg, k, v)
        for use_lm in (False, True):
            sb, _ = synth.build(cfg, lambda grp, N, L: shard.ShardBuilder(lambda g_, n_, l_: gb.Graph(g_, n_, l_), grp, N, L, rank, world), finalize=False)
            sb.finalize(local)
            g = sb.g
            shard.init_engine_nccl(g, rank, world)
            st = g.optimize(n_iter=0 if use_lm else iters, use_lm=use_lm)
            nar = g.allreduces()
            p, v, l = g.get_values()
            a, b = shard.owned_range(n, rank, world)
            own = [p[a - sb.lo:b - sb.lo], v[a - sb.lo:b - sb.lo]]
            halo = [p[0].copy(), v[0].copy()] if rank > 0 else None
            gathered = [None] * world
            dist.all_gather_object(gathered, (own, halo, float(st.error_final), int(st.iterations), int(nar)))
            if rank == 0:
                P = np.concatenate([x[0][0] for x in gathered]); V = np.concatenate([x[0][1] for x in gathered])
                ref, _ = synth.build(cfg, lambda grp, N, L: gb.Graph(grp, N, L))
                sr = ref.optimize(n_iter=0 if use_lm else iters, use_lm=use_lm)
                P0, V0, L0 = ref.get_values()
                from oracle import pyoracle as po
                o, _ = synth.build(cfg, lambda grp, N, L: po.Graph(grp, N, L))
                o.set_threads(po.hardware_threads())
                so = o.optimize(n_iter=0 if use_lm else iters, use_lm=use_lm)
                Po, Vo, Lo = o.get_values()
                d_single = max(np.abs(P - P0).max(), np.abs(V - V0).max(), np.abs(l - L0).max() if L0.size else 0.0)
                d_oracle = max(np.abs(P - Po).max(), np.abs(V - Vo).max(), np.abs(l - Lo).max() if Lo.size else 0.0)
                halos_ok = all(np.array_equal(gathered[r][1][0], P[shard.owned_range(n, r, world)[0] - 1]) and np.array_equal(gathered[r][1][1], V[shard.owned_range(n, r, world)[0] - 1])
                               for r in range(1, world))
                its = [x[3] for x in gathered]; nars = [x[4] for x in gathered]
                good = d_single < (1e-7 if use_lm else 1e-8) and d_oracle < 1e-6 and halos_ok and all(k == sr.iterations for k in its) and sr.iterations == so.iterations
                if not use_lm:
                    good = good and all(k == iters + 2 for k in nars)
                ok = ok and good
                print("%s nccl_check %s n=%d world=%d %s: |sharded - single| = %.2e, |sharded - oracle| = %.2e, halos bit-identical: %s, iterations %s (single %d, oracle %d), all-reduces %s"
                      % ("PASS" if good else "FAIL", name, n, world, "LM" if use_lm else "GN x%d" % iters, d_single, d_oracle, halos_ok, its, sr.iterations, so.iterations, nars), flush=True)
            del g, sb
            dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
