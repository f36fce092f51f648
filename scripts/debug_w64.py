"""GPU debugging aid: the 64-column (16-landmark) SE(3) forward sweep in its three implementations
(default = k_spine + k_panel4, GPB_OLD_PANEL = k_spine + k_panel, GPB_GENERIC_FWD = k_fwd<12,64>) against a dense solve of
the engine's own normal equations (small n) and against each other (full size)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpslam_b200 as gb
from gpslam_b200 import synth

MODES = {"default": {}, "old_panel": {"GPB_OLD_PANEL": "1"}, "generic": {"GPB_GENERIC_FWD": "1"}}


def build(n, mode, seglen=None, n_land=16):
    for k in ("GPB_OLD_PANEL", "GPB_GENERIC_FWD"):
        os.environ.pop(k, None)
    os.environ.update(MODES[mode])
    cfg = synth.config("C3"); cfg.n_states = n; cfg.n_landmarks = n_land
    if n < 2000:
        cfg.prior_every = 40
    def mk(grp, n_, l):
        g = gb.Graph(grp, n_, l)
        if seglen:
            g.set_segment_length(*seglen)
        return g
    g, _ = synth.build(cfg, mk)
    return g


def small():
    for n in (13, 47, 48, 49, 95, 200, 333, 600):
        for seglen in (None, (12, 4), (46, 8), (20, 5)):
            ref = None
            for mode in MODES:
                g = build(n, mode, seglen)
                g.linearize()
                H, rhs = g.normal_equations_dense()
                for lam in (0.0, 1e-2):
                    ds, dl = g.solve_delta(lam)
                    x = np.concatenate([ds.ravel(), dl])
                    ref = np.linalg.solve(H + lam * np.eye(len(rhs)), rhs)
                    err = np.abs(x - ref).max() / max(1.0, np.abs(ref).max())
                    print("n=%4d seglen=%-9s %-10s lam=%g  rel err %.2e %s" % (n, seglen, mode, lam, err, "" if err < 1e-8 else "  <<<<<< BAD"), flush=True)


def full():
    sols = {}
    for mode in ("generic", "default"):
        g = build(100000, mode)
        g.linearize()
        ds, dl = g.solve_delta(0.0)
        sols[mode] = (ds.copy(), dl.copy())
        print(mode, "delta max", np.abs(ds).max(), np.abs(dl).max(), flush=True)
    a, b = sols["generic"], sols["default"]
    d = np.abs(a[0] - b[0]).max(axis=1)
    print("generic vs default: max |ds| diff %.3e at state %d ; landmarks %.3e" % (d.max(), int(d.argmax()), np.abs(a[1] - b[1]).max()))
    bad = np.nonzero(d > 1e-8 * max(1.0, np.abs(a[0]).max()))[0]
    print("states differing:", len(bad), bad[:10], bad[-10:] if len(bad) else "")
    # convergence trace of the tail
    for mode in ("generic", "default"):
        g = build(100000, mode)
        for it in range(8):
            g.optimize(n_iter=1, use_lm=True)
        st = g.optimize(use_lm=True)
        print(mode, "LM done: iterations", st.iterations, "error", st.error_final, "lambda", st.lambda_)
        for it in range(6):
            ds, dl = g.solve_delta(0.0)
            m = np.abs(ds).max(axis=1)
            print("  %s GN %d: max|ds| %.3e at %d ; tail %.3e ; dl %.3e" % (mode, it, m.max(), int(m.argmax()), m[-50:].max(), np.abs(dl).max()), flush=True)
            st = g.optimize(n_iter=1, use_lm=False)
            print("     error", st.error_final)


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] == "small":
        small()
    if len(sys.argv) < 2 or sys.argv[1] == "full":
        full()
