// Solver kernels: multi-level segment-parallel bordered block-tridiagonal Cholesky.
//
// The chain of n states (block size BS = 2D) with a dense border of nb landmark dimensions is cut, at every level, into
// segments of M states: interior states [p+1, q-1] between two separators p and q = p + M.  One CTA eliminates the
// interior of a segment sequentially (block Cholesky), carrying a panel of w = BS + nb + 1 columns
//   [ coupling to the left separator p ("spike", BS) | landmark border (nb) | right-hand side (1) ],
// one panel column per thread, and accumulates the Schur complement onto {p, q, landmarks}.  The separators form the
// chain of the next level (n' = (n-1)/M states) with the same structure; the last level is a single segment whose
// elimination leaves only the landmark system, solved by k_landmark_solve.  Back-substitution walks the levels in
// reverse.  This is the same elimination GTSAM's multifrontal Cholesky performs on a GP chain with landmarks ordered
// last, re-ordered by nested dissection so that segments run in parallel.
//
// Level records (FP64, HBM):
//   level 0 input : HREC[i] = [D_i | E_i | g_i]  + sparse border rows (XR, CSR by interval)
//   level >= 1    : REC[j]  = [D1 | D2 | E | g1 | g2],  BREC[j] = [B1 | B2]   (two parts: written by the segment to the
//                   left (q role) and to the right (p role) of separator j — no atomics, run-to-run deterministic)
//   factors       : FREC[i] = [L_ii | Le_i = E_i L_ii^-T | Y_i = L_ii^-1 panel (BS x w)]
#pragma once
#include "kernels_lin.cuh"

struct FwdArgs {
  int n, M, S, nseg, first_level;
  const double* rec;
  const double* brec;
  const double* XR;
  const int* rowoff;
  const int* rowland;
  int NXRp, nint, nb, DL;
  double lambda;
  double* rec_out;
  double* brec_out;
  double* frec;
  int fstride;
  double* cseg;
  int* flag;
};

struct BwdArgs {
  int n, M, S, nseg, nb;
  const double* frec;
  int fstride;
  const double* xup;  // solution of the next level [S][BS]
  const double* xl;   // landmark solution [nb]
  double* xsol;       // [n][BS]
};

template <int BS, int W>
__global__ void __launch_bounds__((W < 32 ? 32 : W)) k_fwd(const FwdArgs a) {
  constexpr int NT = (W < 32 ? 32 : W);
  constexpr int REC0 = 2 * BS * BS + BS, REC1 = 3 * BS * BS + 2 * BS, YS = BS + 1, NA = W / 2 + 1;
  __shared__ double Dm[BS * BS], Dn[BS * BS], Em[BS * BS], Ysm[W * YS], invd[BS];
  const int c = threadIdx.x;
  const int nb = a.nb, w = BS + nb + 1, M = a.M;
  const bool first = a.first_level != 0;
  const int RECS = first ? REC0 : REC1;
  const bool is_spike = c < BS, is_border = (c >= BS) && (c < BS + nb), is_rhs = (c == BS + nb), active = c < w;
  const int lb = c - BS;

  double acc[NA];
#pragma unroll
  for (int j = 0; j < NA; j++) acc[j] = 0.0;

  // D of state i (both parts summed, + lambda on the diagonal at level 0), element k
  auto D_at = [&](int i, int k) -> double {
    const double* r = a.rec + (size_t)i * RECS;
    if (first) return r[k] + ((k % (BS + 1)) == 0 ? a.lambda : 0.0);
    return r[k] + r[BS * BS + k];
  };
  auto E_ptr = [&](int i) -> const double* { return a.rec + (size_t)i * RECS + (first ? BS * BS : 2 * BS * BS); };
  // own (not yet eliminated) entries of this thread's panel column for state i, added into P
  auto add_own = [&](int i, bool with_spike, int p, double* P) {
    if (is_spike) {
      if (with_spike) {
        const double* E = E_ptr(p);  // rows: state p+1, cols: state p
#pragma unroll
        for (int r = 0; r < BS; r++) P[r] += E[r + c * BS];
      }
    } else if (is_border) {
      if (first) {
        const int l = lb / a.DL, d = lb - l * a.DL;
#pragma unroll 1
        for (int side = 0; side < 2; side++) {
          const int t = side == 0 ? i : i - 1;  // interval whose a-part (side 0) / b-part (side 1) is state i
          if (t < 0 || t >= a.nint) continue;
          for (int row = a.rowoff[t]; row < a.rowoff[t + 1]; row++) {
            if (a.rowland[row] != l) continue;
            const double coef = a.XR[(size_t)(2 * BS + d) * a.NXRp + row];
#pragma unroll
            for (int r = 0; r < BS; r++) P[r] += a.XR[(size_t)(side * BS + r) * a.NXRp + row] * coef;
          }
        }
      } else {
        const double* B = a.brec + (size_t)i * (2 * BS * nb);
#pragma unroll
        for (int r = 0; r < BS; r++) P[r] += B[r + lb * BS] + B[BS * nb + r + lb * BS];
      }
    } else if (is_rhs) {
      const double* r0 = a.rec + (size_t)i * RECS;
      if (first) {
#pragma unroll
        for (int r = 0; r < BS; r++) P[r] += r0[2 * BS * BS + r];
      } else {
#pragma unroll
        for (int r = 0; r < BS; r++) P[r] += r0[3 * BS * BS + r] + r0[3 * BS * BS + BS + r];
      }
    }
  };

  for (int seg = blockIdx.x; seg < a.nseg; seg += gridDim.x) {
    const int p = seg > 0 ? seg * M - 1 : -1;
    const int q = seg < a.S ? (seg + 1) * M - 1 : -1;
    const int i0 = p + 1, i1 = (q >= 0) ? q - 1 : a.n - 1;
    double P[BS];
#pragma unroll
    for (int r = 0; r < BS; r++) P[r] = 0.0;
    for (int k = c; k < BS * BS; k += NT) Dn[k] = 0.0;
    __syncthreads();

    for (int i = i0; i <= i1; i++) {
      const bool has_next = (i < i1) || (q >= 0);
      for (int k = c; k < BS * BS; k += NT) Dm[k] = D_at(i, k) + Dn[k];
      if (has_next) {
        const double* E = E_ptr(i);
        for (int k = c; k < BS * BS; k += NT) Em[k] = E[k];
      }
      add_own(i, (i == i0) && (p >= 0), p, P);
      __syncthreads();
      // ---- in-place Cholesky of Dm (lower, column-major) by the first warp
      if (c < 32) {
#pragma unroll 1
        for (int j = 0; j < BS; j++) {
          const double djj = Dm[j + j * BS];
          if (!(djj > 0.0) && c == 0) *a.flag = 1;
          const double sq = sqrt(djj > 0.0 ? djj : 1.0);
          const double inv = 1.0 / sq;
          __syncwarp();
          if (c == j) { Dm[j + j * BS] = sq; invd[j] = inv; }
          if (c > j && c < BS) Dm[c + j * BS] *= inv;
          __syncwarp();
          if (c > j && c < BS) {
            const double lrj = Dm[c + j * BS];
            for (int cc = j + 1; cc <= c; cc++) Dm[c + cc * BS] -= lrj * Dm[cc + j * BS];
          }
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- Y = L^-1 P  (one column per thread)
      double Y[BS];
#pragma unroll
      for (int r = 0; r < BS; r++) {
        double s = P[r];
#pragma unroll
        for (int t = 0; t < r; t++) s -= Dm[r + t * BS] * Y[t];
        Y[r] = s * invd[r];
      }
      if (c < W) {
#pragma unroll
        for (int r = 0; r < BS; r++) Ysm[c * YS + r] = active ? Y[r] : 0.0;
      }
      double* F = a.frec + (size_t)i * a.fstride;
      if (active) {
#pragma unroll
        for (int r = 0; r < BS; r++) F[2 * BS * BS + c * BS + r] = Y[r];
      }
      // ---- Le = E L^-T, one row per thread (in place in Em)
      if (has_next && c < BS) {
#pragma unroll
        for (int cc = 0; cc < BS; cc++) {
          double s = Em[c + cc * BS];
#pragma unroll
          for (int t = 0; t < cc; t++) s -= Em[c + t * BS] * Dm[cc + t * BS];
          Em[c + cc * BS] = s * invd[cc];
        }
      }
      __syncthreads();
      for (int k = c; k < BS * BS; k += NT) F[k] = Dm[k];
      if (has_next) {
        for (int k = c; k < BS * BS; k += NT) F[BS * BS + k] = Em[k];
        // next panel column: P = -Le Y ; Schur update of the next diagonal block: Dn = -Le Le^T
#pragma unroll
        for (int r = 0; r < BS; r++) {
          double s = 0.0;
#pragma unroll
          for (int t = 0; t < BS; t++) s += Em[r + t * BS] * Y[t];
          P[r] = -s;
        }
        for (int k = c; k < BS * BS; k += NT) {
          const int r = k % BS, cc = k / BS;
          double s = 0.0;
#pragma unroll
          for (int t = 0; t < BS; t++) s += Em[r + t * BS] * Em[cc + t * BS];
          Dn[k] = -s;
        }
      } else {
#pragma unroll
        for (int r = 0; r < BS; r++) P[r] = 0.0;
      }
      // ---- Schur accumulation: acc[j] += Y_c . Y_{(c+j) mod W}
      if (c < W) {
#pragma unroll
        for (int j = 0; j < NA; j++) {
          const double* y2 = Ysm + ((c + j) % W) * YS;
          double s = 0.0;
#pragma unroll
          for (int r = 0; r < BS; r++) s += Y[r] * y2[r];
          acc[j] += s;
        }
      }
      __syncthreads();
    }

    // ---- segment end: hand the Schur complement to the next level
    if (q >= 0) {
      double* R = a.rec_out + (size_t)seg * REC1;
      for (int k = c; k < BS * BS; k += NT) R[k] = D_at(q, k) + Dn[k];  // D1
      add_own(q, false, -1, P);                                          // own border / rhs of q on top of the updates
      if (is_border) {
        double* B = a.brec_out + (size_t)seg * (2 * BS * nb);
#pragma unroll
        for (int r = 0; r < BS; r++) B[r + lb * BS] = P[r];
      } else if (is_rhs) {
#pragma unroll
        for (int r = 0; r < BS; r++) R[3 * BS * BS + r] = P[r];
      } else if (is_spike && p >= 0) {
        double* Ep = a.rec_out + (size_t)(seg - 1) * REC1 + 2 * BS * BS;  // E_{seg-1}: rows separator seg, cols separator seg-1
#pragma unroll
        for (int r = 0; r < BS; r++) Ep[r + c * BS] = P[r];
      }
    }
    if (c < W) {
      double* Rp = (p >= 0) ? a.rec_out + (size_t)(seg - 1) * REC1 : nullptr;
      double* Bp = (p >= 0) ? a.brec_out + (size_t)(seg - 1) * (2 * BS * nb) + BS * nb : nullptr;
#pragma unroll
      for (int j = 0; j < NA; j++) {
        const int c2 = (c + j) % W;
        const int lo = c < c2 ? c : c2, hi = c < c2 ? c2 : c;
        if (lo < BS) {  // spike involved: belongs to separator p, flushed per segment
          if (p >= 0 && hi < w) {
            const double v = -acc[j];
            if (hi < BS) { Rp[BS * BS + lo + hi * BS] = v; Rp[BS * BS + hi + lo * BS] = v; }  // D2
            else if (hi < BS + nb) Bp[lo + (hi - BS) * BS] = v;                                // B2
            else Rp[3 * BS * BS + BS + lo] = v;                                                // g2
          }
          acc[j] = 0.0;
        }
      }
    }
    __syncthreads();
  }
  // ---- landmark x landmark and landmark x rhs parts stay in registers across this CTA's segments
  if (nb > 0 && c < W) {
    double* Cs = a.cseg + (size_t)blockIdx.x * (nb * nb + nb);
#pragma unroll
    for (int j = 0; j < NA; j++) {
      const int c2 = (c + j) % W;
      const int lo = c < c2 ? c : c2, hi = c < c2 ? c2 : c;
      if (lo >= BS && hi < w) {
        const double v = -acc[j];
        if (hi < BS + nb) { Cs[(lo - BS) + (hi - BS) * nb] = v; Cs[(hi - BS) + (lo - BS) * nb] = v; }
        else if (lo < BS + nb) Cs[nb * nb + (lo - BS)] = v;
      }
    }
  }
}

// Back-substitution of one level: x_i = L_ii^-T ( y_i - Yspike_i x_p - Yborder_i x_l - Le_i^T x_{i+1} ), right to left.
template <int BS, int W>
__global__ void __launch_bounds__((W < 32 ? 32 : W)) k_bwd(const BwdArgs a) {
  constexpr int NT = (W < 32 ? 32 : W), NP = NT / BS;
  __shared__ double coef[W], part[NP][BS], Lm[BS * BS], Le[BS * BS], xn[BS], cv[BS];
  const int c = threadIdx.x;
  const int nb = a.nb, w = BS + nb + 1, M = a.M;
  for (int seg = blockIdx.x; seg < a.nseg; seg += gridDim.x) {
    const int p = seg > 0 ? seg * M - 1 : -1;
    const int q = seg < a.S ? (seg + 1) * M - 1 : -1;
    const int i0 = p + 1, i1 = (q >= 0) ? q - 1 : a.n - 1;
    if (c < W) {
      double v = 0.0;
      if (c < BS) v = (p >= 0) ? -a.xup[(size_t)(seg - 1) * BS + c] : 0.0;
      else if (c < BS + nb) v = -a.xl[c - BS];
      else if (c == BS + nb) v = 1.0;
      coef[c] = v;
    }
    if (c < BS) {
      xn[c] = (q >= 0) ? a.xup[(size_t)seg * BS + c] : 0.0;
      if (q >= 0) a.xsol[(size_t)q * BS + c] = xn[c];
    }
    __syncthreads();
    bool has_next = (q >= 0);
    for (int i = i1; i >= i0; i--) {
      const double* F = a.frec + (size_t)i * a.fstride;
      for (int k = c; k < BS * BS; k += NT) { Lm[k] = F[k]; if (has_next) Le[k] = F[BS * BS + k]; }
      // partial sums of Y * coef over a slice of columns
      if (c < NP * BS) {
        const int r = c % BS, pt = c / BS;
        double s = 0.0;
        for (int col = pt; col < w; col += NP) s += F[2 * BS * BS + col * BS + r] * coef[col];
        part[pt][r] = s;
      }
      __syncthreads();
      if (c < BS) {
        double s = 0.0;
#pragma unroll
        for (int pt = 0; pt < NP; pt++) s += part[pt][c];
        if (has_next) {
#pragma unroll
          for (int t = 0; t < BS; t++) s -= Le[t + c * BS] * xn[t];
        }
        cv[c] = s;
      }
      __syncwarp();
      // L^T x = cv, by lane 0..BS-1 of the first warp (column sweep from the bottom)
      if (c < 32) {
        double rhs = (c < BS) ? cv[c] : 0.0;
        double xr = 0.0;
#pragma unroll 1
        for (int r = BS - 1; r >= 0; r--) {
          const double xv = __shfl_sync(0xffffffffu, rhs, r) / Lm[r + r * BS];
          if (c == r) xr = xv;
          if (c < r) rhs -= Lm[r + c * BS] * xv;
        }
        if (c < BS) { xn[c] = xr; a.xsol[(size_t)i * BS + c] = xr; }
      }
      has_next = true;
      __syncthreads();
    }
  }
}

// landmark x landmark base: C0 = sum_rows l^T l (block diagonal), gl0 = sum_rows l^T rhs.  One CTA per landmark.
template <int NT>
__global__ void __launch_bounds__(NT) k_landmark_base(const double* __restrict__ XR, const int* __restrict__ lmoff, const int* __restrict__ lmrows,
                                                      int NXRp, int colL, int DL, int nb, double* __restrict__ Cbase) {
  __shared__ double sred[NT / 32];
  const int l = blockIdx.x;
  double a[12];
#pragma unroll
  for (int k = 0; k < 12; k++) a[k] = 0.0;
  for (int t = lmoff[l] + threadIdx.x; t < lmoff[l + 1]; t += NT) {
    const int row = lmrows[t];
    double lv[3] = {0, 0, 0};
    for (int d = 0; d < DL; d++) lv[d] = XR[(size_t)(colL + d) * NXRp + row];
    const double rh = XR[(size_t)(colL + DL) * NXRp + row];
    for (int d1 = 0; d1 < DL; d1++) {
      for (int d2 = 0; d2 < DL; d2++) a[d1 * 3 + d2] += lv[d1] * lv[d2];
      a[9 + d1] += lv[d1] * rh;
    }
  }
  for (int k = 0; k < 12; k++) {
    const double t = block_sum<NT>(a[k], sred);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (k < 9) { const int d1 = k / 3, d2 = k % 3; if (d1 < DL && d2 < DL) Cbase[(l * DL + d1) + (size_t)(l * DL + d2) * nb] = t; }
      else { const int d1 = k - 9; if (d1 < DL) Cbase[(size_t)nb * nb + l * DL + d1] = t; }
    }
  }
}

// stage 1 of the deterministic reduction of the per-CTA landmark Schur blocks: out[slice][e] = sum_{b = slice mod R} cseg[b][e]
__global__ void k_cseg_reduce(const double* __restrict__ cseg, int nblocks, int entries, int R, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int slice = blockIdx.y;
  if (e >= entries) return;
  double s = 0.0;
  for (int b = slice; b < nblocks; b += R) s += cseg[(size_t)b * entries + e];
  out[(size_t)slice * entries + e] = s;
}

// landmark system: C = Cbase + sum parts + lambda I, Cholesky, solve.  Single CTA.  parts: [nparts][nb*nb+nb]
template <int NT>
__global__ void __launch_bounds__(NT) k_landmark_solve(const double* __restrict__ Cbase, const double* __restrict__ parts, int nparts, int nb,
                                                       double lambda, double* __restrict__ xl, int* __restrict__ flag) {
  extern __shared__ double sm[];
  double* C = sm;            // nb x nb column-major
  double* g = sm + nb * nb;  // nb
  const int entries = nb * nb + nb;
  for (int e = threadIdx.x; e < entries; e += NT) {
    double s = Cbase[e];
    for (int k = 0; k < nparts; k++) s += parts[(size_t)k * entries + e];
    if (e < nb * nb && (e % (nb + 1)) == 0) s += lambda;
    sm[e] = s;
  }
  __syncthreads();
  for (int j = 0; j < nb; j++) {
    const double djj = C[j + j * nb];
    if (!(djj > 0.0) && threadIdx.x == 0) *flag = 2;
    const double sq = sqrt(djj > 0.0 ? djj : 1.0);
    __syncthreads();
    for (int r = j + threadIdx.x; r < nb; r += NT) C[r + j * nb] = (r == j) ? sq : C[r + j * nb] / sq;
    __syncthreads();
    for (int k = threadIdx.x; k < (nb - j - 1) * (nb - j - 1); k += NT) {
      const int r = j + 1 + k % (nb - j - 1), cc = j + 1 + k / (nb - j - 1);
      if (r >= cc) C[r + cc * nb] -= C[r + j * nb] * C[cc + j * nb];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (int r = 0; r < nb; r++) { double s = g[r]; for (int t = 0; t < r; t++) s -= C[r + t * nb] * g[t]; g[r] = s / C[r + r * nb]; }
    for (int r = nb - 1; r >= 0; r--) { double s = g[r]; for (int t = r + 1; t < nb; t++) s -= C[t + r * nb] * g[t]; g[r] = s / C[r + r * nb]; }
    for (int r = 0; r < nb; r++) xl[r] = g[r];
  }
}

// x <- x (+) delta for every state (Pose3 / Rot3: Expmap; Pose2: GTSAM's default chart; vectors: add), plus the two dot
// products LM needs: g.delta and |delta|^2 (block partials).
template <int G, int NT>
__global__ void __launch_bounds__(NT) k_retract(const double* __restrict__ X, const double* __restrict__ xsol, const double* __restrict__ HREC,
                                                double* __restrict__ Xt, double* __restrict__ part_gd, double* __restrict__ part_dd, int N) {
  constexpr int D = GroupTraits<G>::D, PS = GroupTraits<G>::PS, SR = PS + D, bs = 2 * D, REC = 2 * bs * bs + bs;
  __shared__ double sred[NT / 32];
  const int i = blockIdx.x * NT + threadIdx.x;
  double gd = 0.0, dd = 0.0;
  if (i < N) {
    double d[bs];
#pragma unroll
    for (int k = 0; k < bs; k++) d[k] = xsol[(size_t)i * bs + k];
    const double* g = HREC + (size_t)i * REC + 2 * bs * bs;
#pragma unroll
    for (int k = 0; k < bs; k++) { gd += g[k] * d[k]; dd += d[k] * d[k]; }
    const double* x = X + (size_t)i * SR;
    double* y = Xt + (size_t)i * SR;
    if constexpr (G == G_POSE3) {
      const P3 T = p3_compose(p3_from_wire(x), se3_expmap(x6_from(d)));
      p3_to_wire(T, y);
    } else if constexpr (G == G_ROT3) {
      m3_to_wire(m3_from_wire(x) * so3_expmap(v3(d[0], d[1], d[2])), y);
    } else if constexpr (G == G_POSE2) {
      const P2 T = p2_compose(p2(x[0], x[1], x[2]), p2(d[0], d[1], d[2]));
      y[0] = T.x; y[1] = T.y; y[2] = T.th;
    } else {
#pragma unroll
      for (int k = 0; k < D; k++) y[k] = x[k] + d[k];
    }
#pragma unroll
    for (int k = 0; k < D; k++) y[PS + k] = x[PS + k] + d[D + k];
  }
  const double t1 = block_sum<NT>(gd, sred);
  __syncthreads();
  const double t2 = block_sum<NT>(dd, sred);
  if (threadIdx.x == 0) { part_gd[blockIdx.x] = t1; part_dd[blockIdx.x] = t2; }
}

__global__ void k_retract_land(const double* __restrict__ land, const double* __restrict__ xl, const double* __restrict__ gl, double* __restrict__ landt,
                               int n, double* __restrict__ scal) {
  // single block: landmarks are few; also folds their share of g.delta / |delta|^2 into scal[1], scal[2]
  __shared__ double sred[8];
  double gd = 0, dd = 0;
  for (int k = threadIdx.x; k < n; k += 256) { landt[k] = land[k] + xl[k]; gd += gl[k] * xl[k]; dd += xl[k] * xl[k]; }
  const double t1 = block_sum<256>(gd, sred);
  __syncthreads();
  const double t2 = block_sum<256>(dd, sred);
  if (threadIdx.x == 0) { scal[1] += t1; scal[2] += t2; }
}
