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
  int n, M, S, nseg, first_level, extL, extR;   // S = number of interior separators; nseg = S + 1
  int nreal;           // level 0: chain entries >= nreal are ghost replicas (damped by their owner rank, not here)
  int lamL;            // level 0: the pinned first state is owned by this graph -> its pass-through block gets the LM damping here
  const int* sep;      // [S] positions of the interior separators in this level's chain (ascending)
  const double* rec;
  const double* brec;
  const double* XR;
  const int* bsoff;    // level 0: per-state CSR of landmark-bearing rows
  const int* bsrow;    //          row index
  const int* bsside;   //          0: state is the row's a-part, 1: b-part
  const int* rowland;
  const double* bent;  // level 0: packed border entries (k_border_pack), 16 doubles each, CSR order of bsoff
  int NXRp, nb, DL;
  const double* lambda_ptr;  // LM damping lives in device memory so a captured CUDA graph can be replayed with a new value
  double* rec_out;
  double* brec_out;
  double* frec;
  int fstride;
  double* cseg;
  int* flag;
  int store_y;         // 1: Y = L^-1 [spike | border | rhs] is written behind (L^-1 | Le) for the Y-reading back-substitution k_bwd (A/B switch GPB_OLD_BWD)
};

struct BwdArgs {
  int n, M, S, nseg, nb, extL, extR;
  const int* sep;
  const double* frec;
  int fstride;
  const double* xup;  // solution of the next level [extL + S + extR][BS]
  const double* xl;   // landmark solution [nb]
  double* xsol;       // [n][BS]
  // k_bwd2 only: the level's own records (right-hand side, border, coupling to the left separator) are re-read instead of a stored Y
  int first_level, DL;
  const double* rec;    // level 0: HREC; above: [n][3 BS^2 + 2 BS]
  const double* brec;   // level >= 1: [n][2 BS nb]
  const int* bsoff;     // level 0: per-state CSR of the packed border entries
  const double* bent;   // level 0: packed border entries (k_border_pack)
};

// Segment geometry shared by the two sweeps.  Chain states 0..n-1; a pinned first / last state (extL / extR: a shard's external
// separators, or loop-closure endpoints sitting on the chain ends) is never eliminated; interior separators are listed in
// sep[] - every M-th ordinary state plus every pinned state (loop-closure endpoint), which stays a separator at every level
// and so reaches the reduced top system.
struct SegGeom { int p, q, po, qo, i0, i1; };
__device__ __forceinline__ SegGeom seg_geom(int seg, int n, const int* __restrict__ sep, int S, int extL, int extR) {
  SegGeom g;
  g.p = seg > 0 ? sep[seg - 1] : (extL ? 0 : -1);
  g.q = seg < S ? sep[seg] : (extR ? n - 1 : -1);
  g.po = seg > 0 ? extL + seg - 1 : 0;
  g.qo = seg < S ? extL + seg : extL + S + extR - 1;
  g.i0 = g.p + 1;
  g.i1 = g.q >= 0 ? g.q - 1 : n - 1;
  return g;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// lower-triangular 8x8 tile grid of the 64x64 Schur block, interleaved over the two warps: tile t = 2u + warp
__constant__ unsigned char c_tileI[36] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 7, 7};
__constant__ unsigned char c_tileJ[36] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 6, 0, 1, 2, 3, 4, 5, 6, 7};

// Fused Cholesky + triangular inverse of an SPD block (BS <= 12) by ONE warp, right-looking, with the lower-triangular
// entries spread over all 32 lanes (entry e = lane + 32k).  Per pivot step j: the pivot travels by one shuffle, the owners of
// column j of L and of row j of X = L^-1 publish them in shared memory, then every lane applies one FMA per owned entry to the
// Cholesky trailing block (A[r][c] -= L[r][j] L[c][j]) and to the inverse accumulators (Z[r][c] += L[r][j] X[j][c]).
// Dsrc: the SPD block (column-major, lower part read).  Outputs: Li = L^-1 (lower, column-major; strictly-upper untouched),
// Lout (optional) = L, invd[j] = 1 / L[j][j].  colL / rowX: 2 x BS scratch each (double-buffered by step parity).  Returns false if not PD.
template <int BS> struct CholMap { static constexpr int NE = BS * (BS + 1) / 2, K = (NE + 31) / 32; int er[K], ec[K]; };
template <int BS> __device__ __forceinline__ CholMap<BS> chol_map(int lane) {  // (row, col) of the entries a lane owns; computed once per kernel
  CholMap<BS> m;
#pragma unroll
  for (int k = 0; k < CholMap<BS>::K; k++) {
    int e = lane + 32 * k, cidx = 0;
    if (e >= CholMap<BS>::NE) e = CholMap<BS>::NE - 1;  // surplus lanes shadow the last entry (their stores are suppressed)
    while (e >= BS - cidx) { e -= BS - cidx; cidx++; }
    m.ec[k] = cidx; m.er[k] = cidx + e;
  }
  return m;
}
template <int BS>
__device__ __forceinline__ bool warp_chol_inverse(const CholMap<BS>& cm, const double* Dsrc, double* Li, double* Lout, double* invd, double* colL, double* rowX, int lane) {
  constexpr int NE = CholMap<BS>::NE, K = CholMap<BS>::K;
  const int* er = cm.er; const int* ec = cm.ec;
  double A[K], Z[K];
#pragma unroll
  for (int k = 0; k < K; k++) { A[k] = Dsrc[er[k] + ec[k] * BS]; Z[k] = 0.0; }
  bool ok = true;
#pragma unroll 1
  for (int j = 0; j < BS; j++) {
    const int didx = j * BS - (j * (j - 1)) / 2;  // position of (j, j) in the column-major lower enumeration
    const int dslot = didx >> 5;
    double mine = A[0];
#pragma unroll
    for (int k = 1; k < K; k++) if (dslot == k) mine = A[k];
    const double piv = __shfl_sync(0xffffffffu, mine, didx & 31);
    ok &= (piv > 0.0);
    const double inv = rsqrt(piv > 0.0 ? piv : 1.0);
    double* cl = colL + (j & 1) * BS;
    double* rx = rowX + (j & 1) * BS;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const bool live = (lane + 32 * k) < NE;
      if (live && ec[k] == j) {                    // column j of L
        const double l = (er[k] == j) ? piv * inv : A[k] * inv;
        cl[er[k]] = l;
        if (Lout) Lout[er[k] + j * BS] = l;
      }
      if (live && er[k] == j) {                    // row j of X = L^-1
        const double x = (ec[k] == j) ? inv : -Z[k] * inv;
        rx[ec[k]] = x;
        Li[j + ec[k] * BS] = x;
      }
    }
    if (lane == 0) invd[j] = inv;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (ec[k] > j) A[k] -= cl[er[k]] * cl[ec[k]];
      else if (er[k] > j) Z[k] += cl[er[k]] * rx[ec[k]];
    }
  }
  return ok;
}

template <int BS, int W>
__global__ void __launch_bounds__((W < 32 ? 32 : W), (W == 128 ? 3 : W == 64 ? 6 : 8)) k_fwd(const FwdArgs a) {
  constexpr bool MMA = (BS == 12 && W == 64);  // Schur SYRK + panel update on the FP64 tensor pipe
  constexpr int NT = (W < 32 ? 32 : W);
  constexpr int REC0 = 2 * BS * BS + BS, REC1 = 3 * BS * BS + 2 * BS, YS = (MMA || W == 128) ? BS : BS + 1, NA = MMA ? 36 : W / 2 + 1, NBP = W == 128 ? W - BS : W;  // NBP: padded border width (W = 128: trimmed so the static shared memory stays under 48 KB)
  __shared__ __align__(16) double Rb[2][REC1];
  __shared__ double Dm[BS * BS], Dn[BS * BS], Em[BS * BS], Li[BS * BS], Ysm[W * YS], invd[BS], Bn[BS * NBP];
  __shared__ double colL[2 * BS], rowX[2 * BS];
  __shared__ double Psm[W * BS];  // panel columns P (one per thread), kept in shared memory between phases
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gi = lane >> 2, ti = lane & 3;
  const CholMap<BS> cmap = chol_map<BS>(lane);
  const int c = threadIdx.x;
  const int nb = a.nb, w = BS + nb + 1, M = a.M;
  const bool first = a.first_level != 0;
  const int RECS = first ? REC0 : REC1;
  const int oE = first ? BS * BS : 2 * BS * BS, oG = first ? 2 * BS * BS : 3 * BS * BS;
  const bool is_spike = c < BS, is_border = (c >= BS) && (c < BS + nb), is_rhs = (c == BS + nb), active = c < w;
  const int lb = c - BS;

  double acc[NA];
#pragma unroll
  for (int j = 0; j < NA; j++) acc[j] = 0.0;

  auto prefetch = [&](int i, int buf) {  // whole record of state i -> Rb[buf] (16-byte cp.async chunks)
    const double* src = a.rec + (size_t)i * RECS;
    for (int k = c; k < RECS / 2; k += NT) cp_async16(&Rb[buf][2 * k], src + 2 * k);
  };
  // level 0: gather the landmark border of state i into Bn (12 x nb) from the sparse measurement rows; `lane` in [0, 32)
  auto gather_border = [&](int i, int lane) {
    for (int e = a.bsoff[i]; e < a.bsoff[i + 1]; e++) {
      const int row = a.bsrow[e], side = a.bsside[e];
      const int l = a.rowland[row];
      if (lane < BS) {
        const double av = a.XR[(size_t)(side * BS + lane) * a.NXRp + row];
        for (int d = 0; d < a.DL; d++) Bn[lane + (l * a.DL + d) * BS] += av * a.XR[(size_t)(2 * BS + d) * a.NXRp + row];
      }
    }
  };
  // own (not yet eliminated) entries of this thread's panel column for state i (whose record sits in Rb[buf]), added into P
  auto add_own = [&](int i, int buf) {
    double* P = Psm + c * BS;
    if (is_border) {
      if (first) {
#pragma unroll
        for (int r = 0; r < BS; r++) { P[r] += Bn[r + lb * BS]; Bn[r + lb * BS] = 0.0; }
      } else {
        const double* B = a.brec + (size_t)i * (2 * BS * nb);
#pragma unroll
        for (int r = 0; r < BS; r++) P[r] += B[r + lb * BS] + B[BS * nb + r + lb * BS];
      }
    } else if (is_rhs) {
#pragma unroll
      for (int r = 0; r < BS; r++) P[r] += Rb[buf][oG + r] + (first ? 0.0 : Rb[buf][oG + BS + r]);
    }
  };
  auto D_of = [&](int buf, int k) -> double {  // D (parts summed; + lambda on the diagonal at level 0)
    if (first) return Rb[buf][k] + ((k % (BS + 1)) == 0 ? (*a.lambda_ptr) : 0.0);
    return Rb[buf][k] + Rb[buf][BS * BS + k];
  };
  for (int k = c; k < BS * BS; k += NT) Li[k] = 0.0;  // strictly-upper part of L^-1 stays zero

  if (first) { for (int k = c; k < BS * NBP; k += NT) Bn[k] = 0.0; }
  __syncthreads();

  for (int seg = blockIdx.x; seg < a.nseg; seg += gridDim.x) {
    const SegGeom sg = seg_geom(seg, a.n, a.sep, a.S, a.extL, a.extR);
    const int p = sg.p, q = sg.q, i0 = sg.i0, i1 = sg.i1;
    const int ilast = (q >= 0) ? q : i1;  // last record that must be fetched
    for (int k = c; k < BS * BS; k += NT) Dn[k] = 0.0;
    if (i0 <= ilast) prefetch(i0, 0);
    cp_async_commit();
    if (first && nb > 0 && i0 <= ilast && c < 32) gather_border(i0, c);
    // panel init; spike columns start as the coupling of the first interior state to the left separator, E_p
    if (c < W) {
      const bool sp = is_spike && p >= 0 && i0 <= i1;
      const double* E = a.rec + (size_t)(sp ? p : 0) * RECS + oE;
#pragma unroll
      for (int r = 0; r < BS; r++) Psm[c * BS + r] = sp ? E[r + c * BS] : 0.0;
    }
    __syncthreads();

    int buf = 0;
    for (int i = i0; i <= i1; i++, buf ^= 1) {
      const bool has_next = (i < i1) || (q >= 0);
      if (i + 1 <= ilast) prefetch(i + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      for (int k = c; k < BS * BS; k += NT) Dm[k] = D_of(buf, k) + Dn[k];
      if (!MMA && has_next) { for (int k = c; k < BS * BS; k += NT) Em[k] = Rb[buf][oE + k]; }
      add_own(i, buf);
      __syncthreads();
      // ---- factor phase (the serial critical path of the chain): fused Cholesky + L^-1 by the first warp, entries spread over
      //      all 32 lanes (warp_chol_inverse).  Meanwhile the second warp gathers the next state's landmark border.
      if (c < 32) {
        if (NT == 32 && first && nb > 0 && i + 1 <= ilast) gather_border(i + 1, c);
        const bool ok = warp_chol_inverse<BS>(cmap, Dm, Li, MMA ? nullptr : Dm, invd, colL, rowX, c);
        if (!ok && c == 0) *a.flag = 1;
      } else if (c < 64) {
        if (first && nb > 0 && i + 1 <= ilast) gather_border(i + 1, c - 32);
      }
      __syncthreads();
      // ---- Y = L^-1 P  (one column per thread)
      double* F = a.frec + (size_t)i * a.fstride;
      if constexpr (MMA) {
        // ---- Y = L^-1 P on the tensor pipe: warp w owns column tiles 4w..4w+3; row tile 0 skips the zero k-slice 2
        {
          double aLi[2][3];
#pragma unroll
          for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int sK = 0; sK < 3; sK++) aLi[mt][sK] = (8 * mt + gi < BS) ? Li[(8 * mt + gi) + (4 * sK + ti) * BS] : 0.0;
#pragma unroll
          for (int jt = 0; jt < 4; jt++) {
            const int J = 4 * warp + jt;
            double d[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
            for (int sK = 0; sK < 3; sK++) {
              const double bP = Psm[(8 * J + gi) * BS + 4 * sK + ti];
              if (sK < 2) dmma884(d[0][0], d[0][1], aLi[0][sK], bP);
              dmma884(d[1][0], d[1][1], aLi[1][sK], bP);
            }
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
              if (8 * mt + gi < BS) { Ysm[(8 * J + 2 * ti) * YS + 8 * mt + gi] = d[mt][0]; Ysm[(8 * J + 2 * ti + 1) * YS + 8 * mt + gi] = d[mt][1]; }
          }
        }
        // ---- Le = E L^-T : warp w owns row tile w; E is still in the record buffer
        if (has_next) {
          const double* E0 = &Rb[buf][oE];
#pragma unroll
          for (int nt = 0; nt < 2; nt++) {
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int sK = 0; sK < 3; sK++) {
              const double aE = (8 * warp + gi < BS) ? E0[(8 * warp + gi) + (4 * sK + ti) * BS] : 0.0;
              const double bL = (8 * nt + gi < BS) ? Li[(8 * nt + gi) + (4 * sK + ti) * BS] : 0.0;  // (L^-T)[k][j] = L^-1[j][k]
              dmma884(d0, d1, aE, bL);
            }
            if (8 * warp + gi < BS && 8 * nt + 2 * ti < BS) { Em[(8 * warp + gi) + (8 * nt + 2 * ti) * BS] = d0; Em[(8 * warp + gi) + (8 * nt + 2 * ti + 1) * BS] = d1; }
          }
        }
      } else {
        if (c < W) {
          double* yc = Ysm + c * YS;
          const double* pc = Psm + c * BS;
#pragma unroll 1
          for (int r = 0; r < BS; r++) {  // forward substitution through shared memory (own column: no synchronisation needed)
            double sv = active ? pc[r] : 0.0;
            for (int t = 0; t < r; t++) sv -= Dm[r + t * BS] * yc[t];
            yc[r] = sv * invd[r];
          }
        }
        // ---- Le = E L^-T, one row per thread (in place in Em)
        if (has_next && c < BS) {
#pragma unroll 1
          for (int cc = 0; cc < BS; cc++) {
            double sv = Em[c + cc * BS];
            for (int t = 0; t < cc; t++) sv -= Em[c + t * BS] * Dm[cc + t * BS];
            Em[c + cc * BS] = sv * invd[cc];
          }
        }
      }
      __syncthreads();
      if (active && a.store_y) {
        const double* yc = Ysm + c * YS;
#pragma unroll
        for (int r = 0; r < BS; r += 2) st128(F + 2 * BS * BS + c * BS + r, yc[r], yc[r + 1]);
      }
      for (int k = c; k < BS * BS; k += NT) F[k] = Li[k];  // the back-substitution uses L^-1
      if (has_next) {
        for (int k = c; k < BS * BS; k += NT) F[BS * BS + k] = Em[k];
        // Schur update of the next diagonal block: Dn = -Le Le^T
        for (int k = c; k < BS * BS; k += NT) {
          const int r = k % BS, cc = k / BS;
          double s = 0.0;
#pragma unroll
          for (int t = 0; t < BS; t++) s += Em[r + t * BS] * Em[cc + t * BS];
          Dn[k] = -s;
        }
      }
      if constexpr (MMA) {
        // ---- tensor-pipe part.  Y fragments: element (k = 4s + ti) of panel column 8T + gi.
        // (1) next panel  P' = -Le Y : warp w owns column tiles 4w..4w+3, two row tiles (rows 8..11 of the second are padding)
        if (has_next) {
          double aLe[2][3];
#pragma unroll
          for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int sK = 0; sK < 3; sK++) aLe[mt][sK] = (8 * mt + gi < BS) ? Em[(8 * mt + gi) + (4 * sK + ti) * BS] : 0.0;
#pragma unroll
          for (int jt = 0; jt < 4; jt++) {
            const int J = 4 * warp + jt;
            double d[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
            for (int sK = 0; sK < 3; sK++) {
              const double bY = Ysm[(8 * J + gi) * YS + 4 * sK + ti];
              dmma884(d[0][0], d[0][1], aLe[0][sK], bY);
              dmma884(d[1][0], d[1][1], aLe[1][sK], bY);
            }
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
              if (8 * mt + gi < BS) { Psm[(8 * J + 2 * ti) * BS + 8 * mt + gi] = -d[mt][0]; Psm[(8 * J + 2 * ti + 1) * BS + 8 * mt + gi] = -d[mt][1]; }
          }
        }
        else {
#pragma unroll
          for (int r = 0; r < BS; r++) Psm[c * BS + r] = 0.0;
        }
        // (2) Schur accumulation  S += Y^T Y  on the lower-triangular 8x8 tile grid: this warp's tiles t = 2u + warp
#pragma unroll
        for (int u = 0; u < 18; u++) {
          const int t = 2 * u + warp;
          const int I = c_tileI[t], J = c_tileJ[t];
#pragma unroll
          for (int sK = 0; sK < 3; sK++) {
            const double aY = Ysm[(8 * I + gi) * YS + 4 * sK + ti];
            const double bY = Ysm[(8 * J + gi) * YS + 4 * sK + ti];
            dmma884(acc[2 * u], acc[2 * u + 1], aY, bY);
          }
          if ((u & 3) == 3) asm volatile("" ::: "memory");  // keep the scheduler from hoisting every tile's fragment loads (register pressure)
        }
      } else {
        double Y[BS];
#pragma unroll
        for (int r = 0; r < BS; r++) Y[r] = (c < W) ? Ysm[c * YS + r] : 0.0;
        if (c < W) {  // next panel column: P = -Le Y
#pragma unroll
          for (int r = 0; r < BS; r++) {
            double s = 0.0;
            if (has_next) {
#pragma unroll
              for (int t = 0; t < BS; t++) s += Em[r + t * BS] * Y[t];
            }
            Psm[c * BS + r] = -s;
          }
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
      }
      __syncthreads();
    }

    // ---- segment end: hand the Schur complement to the next level (record of q is in Rb[buf])
    cp_async_wait<0>();
    __syncthreads();
    if (q >= 0) {
      double* R = a.rec_out + (size_t)sg.qo * REC1;
      for (int k = c; k < BS * BS; k += NT) R[k] = D_of(buf, k) + Dn[k] - ((first && q >= a.nreal && (k % (BS + 1)) == 0) ? (*a.lambda_ptr) : 0.0);  // D1 (a ghost is damped by its owner)
      add_own(q, buf);                                                      // own border / rhs of q on top of the updates
      const double* P = Psm + (c < W ? c : 0) * BS;
      if (is_border) {
        double* B = a.brec_out + (size_t)sg.qo * (2 * BS * nb);
#pragma unroll
        for (int r = 0; r < BS; r++) B[r + lb * BS] = P[r];
      } else if (is_rhs) {
#pragma unroll
        for (int r = 0; r < BS; r++) R[3 * BS * BS + r] = P[r];
      } else if (is_spike && p >= 0) {
        double* Ep = a.rec_out + (size_t)sg.po * REC1 + 2 * BS * BS;  // rows: separator q, cols: separator p
        if (i0 <= i1) {
#pragma unroll
          for (int r = 0; r < BS; r++) Ep[r + c * BS] = P[r];
        } else {  // no interior state: p and q are directly coupled
          const double* E = a.rec + (size_t)p * RECS + oE;
#pragma unroll
          for (int r = 0; r < BS; r++) Ep[r + c * BS] = E[r + c * BS];
        }
      }
      if (a.extR && seg == a.S) {  // external right separator: no segment to its right -> its part 2 is zero
        for (int k = c; k < BS * BS; k += NT) R[BS * BS + k] = 0.0;
        if (c < BS) R[3 * BS * BS + BS + c] = 0.0;
        if (is_border) { double* B = a.brec_out + (size_t)sg.qo * (2 * BS * nb) + BS * nb;
#pragma unroll
          for (int r = 0; r < BS; r++) B[r + lb * BS] = 0.0; }
      }
    }
    if (a.extL && seg == 0) {
      // external left separator: nobody is to its left here -> part 1 carries this shard's own (undamped) share of it
      double* R = a.rec_out;  // po == 0
      const double* src = a.rec + (size_t)p * RECS;
      for (int k = c; k < BS * BS; k += NT) R[k] = first ? src[k] + ((a.lamL && (k % (BS + 1)) == 0) ? (*a.lambda_ptr) : 0.0) : src[k] + src[BS * BS + k];
      if (c < BS) R[3 * BS * BS + c] = first ? src[oG + c] : src[oG + c] + src[oG + BS + c];
      __syncthreads();  // Bn is free (q consumed it); gather p's border when it exists
      if (first && nb > 0) { if (c < 32) gather_border(p, c); }
      __syncthreads();
      if (is_border) {
        double* B = a.brec_out;
        if (first) {
#pragma unroll
          for (int r = 0; r < BS; r++) { B[r + lb * BS] = Bn[r + lb * BS]; Bn[r + lb * BS] = 0.0; }
        } else {
          const double* Bs = a.brec + (size_t)p * (2 * BS * nb);
#pragma unroll
          for (int r = 0; r < BS; r++) B[r + lb * BS] = Bs[r + lb * BS] + Bs[BS * nb + r + lb * BS];
        }
      }
    }
    {
      double* Rp = (p >= 0) ? a.rec_out + (size_t)sg.po * REC1 : nullptr;
      double* Bp = (p >= 0) ? a.brec_out + (size_t)sg.po * (2 * BS * nb) + BS * nb : nullptr;
      // entry (x, y) of the accumulated Y^T Y with a spike column involved belongs to separator p: flush and reset
      auto flush_spike = [&](int x, int y, double& av) {
        const int lo = x < y ? x : y, hi = x < y ? y : x;
        if (lo < BS) {
          if (p >= 0 && hi < w) {
            const double v = -av;
            if (hi < BS) { Rp[BS * BS + lo + hi * BS] = v; Rp[BS * BS + hi + lo * BS] = v; }  // D2
            else if (hi < BS + nb) Bp[lo + (hi - BS) * BS] = v;                                // B2
            else Rp[3 * BS * BS + BS + lo] = v;                                                // g2
          }
          av = 0.0;
        }
      };
      if constexpr (MMA) {
#pragma unroll
        for (int u = 0; u < 18; u++) {
          const int t = 2 * u + warp;
          const int x = 8 * c_tileI[t] + gi, y = 8 * c_tileJ[t] + 2 * ti;
          if (x >= y) flush_spike(x, y, acc[2 * u]); else if (x < BS || y < BS) acc[2 * u] = 0.0;
          if (x >= y + 1) flush_spike(x, y + 1, acc[2 * u + 1]); else if (x < BS || y + 1 < BS) acc[2 * u + 1] = 0.0;
        }
      } else if (c < W) {
#pragma unroll
        for (int j = 0; j < NA; j++) flush_spike(c, (c + j) % W, acc[j]);
      }
    }
    __syncthreads();
  }
  // ---- landmark x landmark and landmark x rhs parts stay in registers across this CTA's segments
  if (nb > 0) {
    double* Cs = a.cseg + (size_t)blockIdx.x * (nb * nb + nb);
    auto flush_land = [&](int x, int y, double av) {
      const int lo = x < y ? x : y, hi = x < y ? y : x;
      if (lo >= BS && hi < w) {
        const double v = -av;
        if (hi < BS + nb) { Cs[(lo - BS) + (hi - BS) * nb] = v; Cs[(hi - BS) + (lo - BS) * nb] = v; }
        else if (lo < BS + nb) Cs[nb * nb + (lo - BS)] = v;
      }
    };
    if constexpr (MMA) {
#pragma unroll
      for (int u = 0; u < 18; u++) {
        const int t = 2 * u + warp;
        const int x = 8 * c_tileI[t] + gi, y = 8 * c_tileJ[t] + 2 * ti;
        if (x >= y) flush_land(x, y, acc[2 * u]);
        if (x >= y + 1) flush_land(x, y + 1, acc[2 * u + 1]);
      }
    } else if (c < W) {
#pragma unroll
      for (int j = 0; j < NA; j++) flush_land(c, (c + j) % W, acc[j]);
    }
  }
}

// 1/sqrt(x) for a positive, normal x: the hardware seed (MUFU.RSQ64H, ~20 bits) and one cubically convergent step; no special-case branches
__device__ __forceinline__ double rsqrt_pos(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  // one third-order step  y <- y (1 + e/2 + 3 e^2 / 8),  e = 1 - x y^2 : seed error 2^-21 -> e^3 ~ 2^-63, four dependent operations
  const double e = fma(-(x * y), y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
}
// ---------------------------------------------------------------------------------------------------------------------
// Two-kernel forward sweep (BS = 12, panel width 64).  The chain "spine"  D'_i = D_i - Le_{i-1} Le_{i-1}^T -> L_i^-1 ->
// Le_i = E_i L_i^-T  never depends on the panel, so it is factored first by k_spine: ONE WARP PER SEGMENT, 16 independent warps
// per SM - the recurrence is latency-bound (12 dependent pivots per state), and only thread-level parallelism hides that.
// k_panel4 then streams the stored (L^-1, Le) and does nothing but tensor-pipe products: Y = L^-1 P, P' = own - Le Y, S += Y^T Y.
//
// Spine, second version: the block column [D_i ; E_i] (24 x 12) is factored as ONE panel.  Lanes 0..11 own the rows of D_i,
// lanes 12..23 the rows of E_i = H_{i+1,i}; the right-looking pivot loop (one shuffle per column entry) then leaves the rows of
// L_i in lanes 0..11 and the rows of Le_i = E_i L_i^-T in lanes 12..23 - Le costs no instruction of its own, and it no longer
// waits for the inverse.  Lanes 0..11 also build column `lane` of L_i^-1 with one extra FMA per shuffle (needed by the panel
// kernel and the back-substitution) and stream it straight to HBM from registers.  Only the Schur update Dn = -Le Le^T goes
// through shared memory (Le out, 9 DMMA on the lower tiles, Dn back in the row-per-lane layout, read as 128-bit column loads
// thanks to symmetry).  The next pivot is broadcast from an early copy (arow[j+1] - l^2 on its owner lane) so the pivot chain
// is  shfl -> rsqrt -> mul  per column instead of waiting for the column broadcast.
// `ready` (fused upper-level kernel only): shared-memory counter of the states this warp has finished, published after a
// fence so that the panel warps of the same CTA may fetch their (L^-1, Le) from HBM / L2.
template <int BS, bool FIRST, bool FUSED>
__device__ __forceinline__ void spine_body(const FwdArgs& a, const int lane, volatile int* ready) {
  static_assert(BS == 12, "spine kernel is specialised for 12 x 12 state blocks");
  constexpr int REC0 = 2 * BS * BS + BS, REC1 = 3 * BS * BS + 2 * BS;
  __shared__ __align__(16) double Les[BS * BS], Dn[BS * BS];
  const int gi = lane >> 2, ti = lane & 3;
  int done = 0;
  const bool dl = lane < BS, el = lane >= BS && lane < 2 * BS;
  const int rr = dl ? lane : (el ? lane - BS : 0);
  constexpr bool first = FIRST;
  constexpr int RECS = first ? REC0 : REC1, oE = first ? BS * BS : 2 * BS * BS;
  const double lambda = first ? *a.lambda_ptr : 0.0;
  // next state's row, fetched one state ahead: D on lanes 0..11, E on lanes 12..23 (every other lane re-reads lane 0's row and is
  // masked when the row is consumed - nothing may depend on the loaded values here, or the warp would wait for HBM on the spot);
  // above level 0 the second part of D arrives in nrow2
  double nrow[BS], nrow2[first ? 1 : BS];
  bool nact = false;
  auto fetch = [&](int i, bool want_e) {
    const double* r = a.rec + (size_t)i * RECS;
    nact = dl || (el && want_e);
    const double* p0 = r + (el ? oE + rr : rr);
#pragma unroll
    for (int cc = 0; cc < BS; cc++) {
      nrow[cc] = p0[cc * BS];
      if constexpr (!first) nrow2[cc] = r[BS * BS + rr + cc * BS];
    }
  };
  for (int seg = blockIdx.x; seg < a.nseg; seg += gridDim.x) {
    const SegGeom sg = seg_geom(seg, a.n, a.sep, a.S, a.extL, a.extR);
    const int q = sg.q, i0 = sg.i0, i1 = sg.i1;
    const int ilast = (q >= 0) ? q : i1;
    for (int k = lane; k < BS * BS; k += 32) Dn[k] = 0.0;
    if (i0 <= ilast) fetch(i0, (i0 < i1) || (q >= 0 && i0 <= i1));
    __syncwarp();
    bool ok = true;
#pragma unroll 1
    for (int i = i0; i <= i1; i++) {
      const bool has_next = (i < i1) || (q >= 0);
      double arow[BS], sv[BS];
      {
        const double* dn = Dn + rr * BS;  // column rr == row rr (symmetric)
#pragma unroll
        for (int cc = 0; cc < BS; cc += 2) {
          const double2 t = *reinterpret_cast<const double2*>(dn + cc);
          double v0 = nrow[cc], v1 = nrow[cc + 1];
          if constexpr (!first) { v0 += dl ? nrow2[cc] : 0.0; v1 += dl ? nrow2[cc + 1] : 0.0; }
          arow[cc] = (nact ? v0 : 0.0) + (dl ? t.x : 0.0) + ((dl && cc == rr) ? lambda : 0.0);
          arow[cc + 1] = (nact ? v1 : 0.0) + (dl ? t.y : 0.0) + ((dl && cc + 1 == rr) ? lambda : 0.0);
        }
      }
      if (i + 1 <= ilast) fetch(i + 1, (i + 1 < i1) || (q >= 0 && i + 1 <= i1));  // next state's loads fly during this factorisation
      double* F = a.frec + (size_t)i * a.fstride;
#pragma unroll
      for (int cc = 0; cc < BS; cc++) sv[cc] = 0.0;
      double piv = __shfl_sync(0xffffffffu, arow[0], 0), xprev = 0.0;
#pragma unroll
      for (int j = 0; j < BS; j++) {
        ok &= (piv > 0.0);
        const double inv = rsqrt_pos(piv > 0.0 ? piv : 1.0);
        const double lrj = (lane == j) ? piv * inv : arow[j] * inv;   // lanes 0..11: L[lane][j] (lane >= j); lanes 12..23: Le[rr][j]
        if (j + 1 < BS) piv = __shfl_sync(0xffffffffu, fma(-lrj, lrj, arow[j + 1 < BS ? j + 1 : j]), j + 1);  // early copy of the next pivot
        const double xj = (lane == j) ? inv : ((lane > j) ? 0.0 : -sv[j] * inv);   // X[j][lane], X = L^-1 (lanes 0..11)
        if (j & 1) { if (dl) st128(F + (j - 1) + lane * BS, xprev, xj); } else xprev = xj;
        if (el && has_next) { Les[rr + j * BS] = lrj; F[BS * BS + rr + j * BS] = lrj; }
#pragma unroll
        for (int cc = j + 1; cc < BS; cc++) {
          const double lcj = __shfl_sync(0xffffffffu, lrj, cc);
          arow[cc] = fma(-lrj, lcj, arow[cc]);
          sv[cc] = fma(lcj, xj, sv[cc]);
        }
      }
      __syncwarp();
      if (has_next) {
        // Dn = -Le Le^T on the lower 8x8 tiles; written back symmetric so that lane r reads its row as a contiguous column
        double f[2][3];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
          for (int sK = 0; sK < 3; sK++) f[mt][sK] = (8 * mt + gi < BS) ? Les[(8 * mt + gi) + (4 * sK + ti) * BS] : 0.0;
        double c00a = 0.0, c00b = 0.0, c10a = 0.0, c10b = 0.0, c11a = 0.0, c11b = 0.0;
#pragma unroll
        for (int sK = 0; sK < 3; sK++) {
          dmma884(c00a, c00b, f[0][sK], f[0][sK]);
          dmma884(c10a, c10b, f[1][sK], f[0][sK]);
          dmma884(c11a, c11b, f[1][sK], f[1][sK]);
        }
        __syncwarp();
        // tile (0,0): element (gi, 2ti + {0,1}) -> stored at its transposed position (a contiguous pair)
        st128(Dn + 2 * ti + gi * BS, -c00a, -c00b);
        if (gi < 4) {
          // tile (1,0): element (8 + gi, 2ti + {0,1}): upper copy as a pair, lower copy as two scalars
          st128(Dn + 2 * ti + (8 + gi) * BS, -c10a, -c10b);
          Dn[(8 + gi) + (2 * ti) * BS] = -c10a; Dn[(8 + gi) + (2 * ti + 1) * BS] = -c10b;
          // tile (1,1): element (8 + gi, 8 + 2ti + {0,1}), 2ti + 1 < 4
          if (ti < 2) st128(Dn + 8 + 2 * ti + (8 + gi) * BS, -c11a, -c11b);
        }
      }
      __syncwarp();
      if constexpr (FUSED) {  // L^-1 and Le of state i are in HBM / L2: let the panel warps go
        __threadfence();
        __syncwarp();
        done++;
        if (lane == 0) atomicExch(const_cast<int*>(ready), done);   // published with an atomic: the counter is a flag, not barrier-protected data
      }
    }
    if (!ok && lane == 0) *a.flag = 1;
    if (q >= 0 && dl) {  // D1 of the right separator: its own block (fetched last) + the last Schur update (+ damping at level 0)
      double* R = a.rec_out + (size_t)sg.qo * REC1;
#pragma unroll
      for (int cc = 0; cc < BS; cc++) { double v = nrow[cc]; if constexpr (!first) v += nrow2[cc]; R[rr + cc * BS] = v + Dn[cc + rr * BS] + ((cc == rr && q < a.nreal) ? lambda : 0.0); }
    }
    __syncwarp();
  }
}
template <int BS, bool FIRST>
__global__ void __launch_bounds__(32, 16) k_spine(const FwdArgs a) { spine_body<BS, FIRST, false>(a, threadIdx.x, nullptr); }


// ---------------------------------------------------------------------------------------------------------------------
// Panel kernel, four warps.  Warp pw owns column tiles 2pw, 2pw+1 of the 12 x 64 panel [spike | border | rhs] for the
// thread-per-column steps and for Y = L^-1 P, P' = -Le Y; the symmetric update S += Y^T Y (36 lower 8x8 tiles) is split nine
// tiles per warp with compile-time tile lists, so its fragments and accumulators never leave registers.  ONE block barrier
// per state (Y of all tiles visible before the rank-12 update; Y is double-buffered so the next state's Y can be written
// while slower warps still read this one).  (L^-1 | Le), g and the packed border entries of state i+2 stream in with
// cp.async (three stages) while state i is processed.
__host__ __device__ constexpr int syrk_tile_I(int pw, int u) {
  constexpr int t[4][9] = {{0, 1, 1, 2, 2, 2, 3, 3, 3}, {4, 4, 4, 4, 5, 5, 5, 5, 3}, {6, 6, 6, 6, 7, 7, 7, 7, 7}, {4, 5, 5, 6, 6, 6, 7, 7, 7}};
  return t[pw][u];
}
__host__ __device__ constexpr int syrk_tile_J(int pw, int u) {
  constexpr int t[4][9] = {{0, 0, 1, 0, 1, 2, 1, 2, 3}, {0, 1, 2, 3, 0, 1, 2, 3, 0}, {0, 1, 2, 3, 0, 1, 2, 3, 4}, {4, 4, 5, 4, 5, 6, 5, 6, 7}};
  return t[pw][u];
}
__host__ __device__ constexpr bool syrk_needs(int pw, int t) {
  for (int u = 0; u < 9; u++) if (syrk_tile_I(pw, u) == t || syrk_tile_J(pw, u) == t) return true;
  return false;
}
template <int PW> __device__ __forceinline__ void syrk_accum(const double* __restrict__ Y, double (&acc)[18], int gi, int ti) {
  constexpr int BS = 12;
#pragma unroll
  for (int sK = 0; sK < 3; sK++) {
    double yf[8];
    static_for<0, 8>([&](auto T) {
      constexpr int t = decltype(T)::value;
      if constexpr (syrk_needs(PW, t)) yf[t] = Y[(8 * t + gi) * BS + 4 * sK + ti];
    });
    static_for<0, 9>([&](auto U) {
      constexpr int u = decltype(U)::value;
      dmma884(acc[2 * u], acc[2 * u + 1], yf[syrk_tile_I(PW, u)], yf[syrk_tile_J(PW, u)]);
    });
  }
}

// level-0 border entries, one 128-byte record per landmark-bearing row in per-state CSR order:
// [0..11] the row's coefficients on the state, [12..12+DL) its coefficients on the landmark, [15] the landmark index
__global__ void k_border_pack(const double* __restrict__ XR, const int* __restrict__ bsrow, const int* __restrict__ bsside, const int* __restrict__ rowland,
                              int nent, int BS, int DL, int NXRp, double* __restrict__ bent) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, e = t >> 4, k = t & 15;
  if (e >= nent) return;
  const int row = bsrow[e], side = bsside[e];
  double v = 0.0;
  if (k < BS) v = XR[(size_t)(side * BS + k) * NXRp + row];
  else if (k < BS + DL) v = XR[(size_t)(2 * BS + (k - BS)) * NXRp + row];
  else if (k == 15) v = (double)rowland[row];
  bent[(size_t)e * 16 + k] = v;
}

// FUSED (upper-level kernel): the 128 panel threads share the CTA with a spine warp; block barriers become a named barrier of
// the panel threads, and a state's (L^-1, Le) is fetched only after the spine warp has published it (`ready`).
template <bool FUSED> __device__ __forceinline__ void panel_sync() {
  if constexpr (FUSED) asm volatile("bar.sync 1, 128;" ::: "memory"); else __syncthreads();
}
template <int BS, bool FUSED>
__device__ __forceinline__ void panel_body(const FwdArgs& a, volatile const int* ready) {
  static_assert(BS == 12, "panel kernel is specialised for 12 x 12 state blocks");
  constexpr int W = 64, NT = 128, REC0 = 2 * BS * BS + BS, REC1 = 3 * BS * BS + 2 * BS, NST = 3, MAXE = 6, HB = BS / 2;
  __shared__ __align__(16) double Fb[NST][2 * BS * BS];  // (L^-1 | Le)
  __shared__ __align__(16) double Gb[NST][2 * BS];       // rhs block(s) g
  __shared__ __align__(16) double Eb[NST][MAXE * 16];    // packed border entries (level 0)
  __shared__ __align__(16) double Psm[W * BS], Ysm[2][W * BS];
  const int c = threadIdx.x, pw = c >> 5, lane = c & 31, gi = lane >> 2, ti = lane & 3;
  const int col = 16 * pw + (lane & 15), r0 = HB * (lane >> 4);  // thread-per-half-column steps: column, first row
  const int nb = a.nb, w = BS + nb + 1, M = a.M;
  const bool first = a.first_level != 0;
  const bool ent = first && nb > 0;
  const int RECS = first ? REC0 : REC1;
  const int oE = first ? BS * BS : 2 * BS * BS, oG = first ? 2 * BS * BS : 3 * BS * BS;
  const bool is_border = (col >= BS) && (col < BS + nb), is_rhs = (col == BS + nb), is_spike = col < BS, active = col < w;
  const int lb = col - BS;
  const int myl = is_border ? lb / a.DL : -1, myd = is_border ? lb % a.DL : 0;
  double acc[18];
#pragma unroll
  for (int j = 0; j < 18; j++) acc[j] = 0.0;
  double bn1[HB], bn2[HB];   // upper levels: the next state's dense border half-column, in flight
#pragma unroll
  for (int r = 0; r < HB; r++) { bn1[r] = 0.0; bn2[r] = 0.0; }
  auto with_pw = [&](auto fn) {
    if (pw == 0) fn(std::integral_constant<int, 0>{});
    else if (pw == 1) fn(std::integral_constant<int, 1>{});
    else if (pw == 2) fn(std::integral_constant<int, 2>{});
    else fn(std::integral_constant<int, 3>{});
  };
  auto bso = [&](int i) { return ent ? a.bsoff[i < a.n ? i : a.n] : 0; };

  int seg_base = 0;  // states the spine warp finished before this segment (it walks the same segments in the same order)
  for (int seg = blockIdx.x; seg < a.nseg; seg += gridDim.x) {
    const SegGeom sg = seg_geom(seg, a.n, a.sep, a.S, a.extL, a.extR);
    const int p = sg.p, q = sg.q, i0 = sg.i0, i1 = sg.i1;
    const int ilast = (q >= 0) ? q : i1;
    int b0 = bso(i0), b1 = bso(i0 + 1), b2 = bso(i0 + 2), b3 = bso(i0 + 3);  // rolling window of CSR offsets: states i .. i+3
    auto prefetch = [&](int i, int st, int e0, int e1) {
      if (i <= i1) {
        if constexpr (FUSED) {  // the spine warp has published state i (one poller per warp, atomic read of the flag)
          if (lane == 0) { while (atomicAdd(const_cast<int*>(ready), 0) <= seg_base + (i - i0)) { } }
          __syncwarp();
        }
        const double* src = a.frec + (size_t)i * a.fstride;
        const int n2 = (((i < i1) || (q >= 0)) ? 2 * BS * BS : BS * BS) / 2;
        for (int k = c; k < n2; k += NT) cp_async16(&Fb[st][2 * k], src + 2 * k);
      }
      if (i <= ilast) {
        if (c < (first ? BS : 2 * BS) / 2) cp_async16(&Gb[st][2 * c], a.rec + (size_t)i * RECS + oG + 2 * c);
        if (ent) { const int ne = min(e1 - e0, MAXE); if (c < 8 * ne) cp_async16(&Eb[st][2 * c], a.bent + (size_t)e0 * 16 + 2 * c); }
      }
    };
    // own half-column of state i: border entries / dense border blocks, rhs
    auto add_own = [&](int i, int st, int e0, int e1) {
      double* P = Psm + col * BS + r0;
      if (is_border) {
        if (first) {
          const int ne = e1 - e0;
          for (int k = 0; k < ne; k++) {
            if (k < MAXE) {
              const double* en = &Eb[st][16 * k];
              if ((int)en[15] == myl) { const double h = en[BS + myd];
#pragma unroll
                for (int r = 0; r < HB; r++) P[r] += en[r0 + r] * h; }
            } else {
              const double* en = a.bent + (size_t)(e0 + k) * 16;
              if ((int)en[15] == myl) { const double h = en[BS + myd];
#pragma unroll
                for (int r = 0; r < HB; r++) P[r] += en[r0 + r] * h; }
            }
          }
        } else {
          // dense border blocks of the upper levels: this state's half-column was loaded one state ahead (a global load issued here
          // would sit on the per-state critical path: it was the largest stall of k_level_ws); fetch the next state's now
#pragma unroll
          for (int r = 0; r < HB; r++) P[r] += bn1[r] + bn2[r];
          if (i + 1 <= ilast) {
            const double* B = a.brec + (size_t)(i + 1) * (2 * BS * nb) + r0 + lb * BS;
#pragma unroll
            for (int r = 0; r < HB; r++) { bn1[r] = B[r]; bn2[r] = B[BS * nb + r]; }
          }
        }
      } else if (is_rhs) {
#pragma unroll
        for (int r = 0; r < HB; r++) P[r] += Gb[st][r0 + r] + (first ? 0.0 : Gb[st][BS + r0 + r]);
      }
    };
    if (!first && is_border && i0 <= ilast) {
      const double* B = a.brec + (size_t)i0 * (2 * BS * nb) + r0 + lb * BS;
#pragma unroll
      for (int r = 0; r < HB; r++) { bn1[r] = B[r]; bn2[r] = B[BS * nb + r]; }
    }
    prefetch(i0, 0, b0, b1); cp_async_commit();
    prefetch(i0 + 1, 1, b1, b2); cp_async_commit();
    {
      const bool sp = is_spike && p >= 0 && i0 <= i1;
      const double* E = a.rec + (size_t)(sp ? p : 0) * RECS + oE + r0 + col * BS;
#pragma unroll
      for (int r = 0; r < HB; r++) Psm[col * BS + r0 + r] = sp ? E[r] : 0.0;
    }
    cp_async_wait<1>();
    panel_sync<FUSED>();
    int st = 0, ys = 0;
    for (int i = i0; i <= i1; i++) {
      const bool has_next = (i < i1) || (q >= 0);
      const int st2 = st == 0 ? 2 : st - 1;  // stage of state i+2 == stage of state i-1, free since the previous barrier
      prefetch(i + 2, st2, b2, b3);
      cp_async_commit();
      const int b4 = bso(i + 4);
      add_own(i, st, b0, b1);
      __syncwarp();
      const double* Li = Fb[st];
      const double* Le = Fb[st] + BS * BS;
      double* Y = Ysm[ys];
      // ---- Y = L^-1 P (own column tiles)
      {
        double aLi[2][3];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
          for (int sK = 0; sK < 3; sK++) aLi[mt][sK] = (8 * mt + gi < BS) ? Li[(8 * mt + gi) + (4 * sK + ti) * BS] : 0.0;
        double d[2][2][2];
#pragma unroll
        for (int jt = 0; jt < 2; jt++) { d[jt][0][0] = d[jt][0][1] = d[jt][1][0] = d[jt][1][1] = 0.0; }
#pragma unroll
        for (int sK = 0; sK < 3; sK++)
#pragma unroll
          for (int jt = 0; jt < 2; jt++) {
            const double bP = Psm[(8 * (2 * pw + jt) + gi) * BS + 4 * sK + ti];
            if (sK < 2) dmma884(d[jt][0][0], d[jt][0][1], aLi[0][sK], bP);
            dmma884(d[jt][1][0], d[jt][1][1], aLi[1][sK], bP);
          }
#pragma unroll
        for (int jt = 0; jt < 2; jt++) {
          const int J = 2 * pw + jt;
#pragma unroll
          for (int mt = 0; mt < 2; mt++)
            if (8 * mt + gi < BS) { Y[(8 * J + 2 * ti) * BS + 8 * mt + gi] = d[jt][mt][0]; Y[(8 * J + 2 * ti + 1) * BS + 8 * mt + gi] = d[jt][mt][1]; }
        }
      }
      __syncwarp();
      // ---- P' = -Le Y (own column tiles)
      if (has_next) {
        double aLe[2][3];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
          for (int sK = 0; sK < 3; sK++) aLe[mt][sK] = (8 * mt + gi < BS) ? Le[(8 * mt + gi) + (4 * sK + ti) * BS] : 0.0;
        double d[2][2][2];
#pragma unroll
        for (int jt = 0; jt < 2; jt++) { d[jt][0][0] = d[jt][0][1] = d[jt][1][0] = d[jt][1][1] = 0.0; }
#pragma unroll
        for (int sK = 0; sK < 3; sK++)
#pragma unroll
          for (int jt = 0; jt < 2; jt++) {
            const double bY = Y[(8 * (2 * pw + jt) + gi) * BS + 4 * sK + ti];
            dmma884(d[jt][0][0], d[jt][0][1], aLe[0][sK], bY);
            dmma884(d[jt][1][0], d[jt][1][1], aLe[1][sK], bY);
          }
#pragma unroll
        for (int jt = 0; jt < 2; jt++) {
          const int J = 2 * pw + jt;
#pragma unroll
          for (int mt = 0; mt < 2; mt++)
            if (8 * mt + gi < BS) { Psm[(8 * J + 2 * ti) * BS + 8 * mt + gi] = -d[jt][mt][0]; Psm[(8 * J + 2 * ti + 1) * BS + 8 * mt + gi] = -d[jt][mt][1]; }
        }
      } else {
#pragma unroll
        for (int r = 0; r < HB; r++) Psm[col * BS + r0 + r] = 0.0;
      }
      // ---- Y half-column to HBM (only for the Y-reading back-substitution)
      if (active && a.store_y) {
        double* F = a.frec + (size_t)i * a.fstride + 2 * BS * BS + col * BS + r0;
        const double* yc = Y + col * BS + r0;
#pragma unroll
        for (int r = 0; r < HB; r += 2) st128(F + r, yc[r], yc[r + 1]);
      }
      cp_async_wait<1>();
      panel_sync<FUSED>();
      // ---- S += Y^T Y (this warp's nine tiles)
      with_pw([&](auto PW) { syrk_accum<decltype(PW)::value>(Y, acc, gi, ti); });
      b0 = b1; b1 = b2; b2 = b3; b3 = b4;
      st = st == 2 ? 0 : st + 1; ys ^= 1;
    }
    cp_async_wait<0>();
    panel_sync<FUSED>();
    // ---- segment end (D1 of q was written by k_spine): the closing separator's own border / rhs, then hand the panel off
    if (q >= 0) {
      double* R = a.rec_out + (size_t)sg.qo * REC1;
      add_own(q, st, b0, b1);
      const double* P = Psm + col * BS + r0;
      if (is_border) {
        double* B = a.brec_out + (size_t)sg.qo * (2 * BS * nb) + r0 + lb * BS;
#pragma unroll
        for (int r = 0; r < HB; r++) B[r] = P[r];
      } else if (is_rhs) {
#pragma unroll
        for (int r = 0; r < HB; r++) R[3 * BS * BS + r0 + r] = P[r];
      } else if (is_spike && p >= 0) {
        double* Ep = a.rec_out + (size_t)sg.po * REC1 + 2 * BS * BS + r0 + col * BS;
        if (i0 <= i1) {
#pragma unroll
          for (int r = 0; r < HB; r++) Ep[r] = P[r];
        } else {
          const double* E = a.rec + (size_t)p * RECS + oE + r0 + col * BS;
#pragma unroll
          for (int r = 0; r < HB; r++) Ep[r] = E[r];
        }
      }
      if (a.extR && seg == a.S) {
        for (int k = c; k < BS * BS; k += NT) R[BS * BS + k] = 0.0;
        if (c < BS) R[3 * BS * BS + BS + c] = 0.0;
        if (is_border) { double* B = a.brec_out + (size_t)sg.qo * (2 * BS * nb) + BS * nb + r0 + lb * BS;
#pragma unroll
          for (int r = 0; r < HB; r++) B[r] = 0.0; }
      }
    }
    if (a.extL && seg == 0) {  // the left external separator (halo state p): its own blocks pass through to the next level
      double* R = a.rec_out;
      const double* src = a.rec + (size_t)p * RECS;
      for (int k = c; k < BS * BS; k += NT) R[k] = first ? src[k] + ((a.lamL && (k % (BS + 1)) == 0) ? (*a.lambda_ptr) : 0.0) : src[k] + src[BS * BS + k];
      if (c < BS) R[3 * BS * BS + c] = first ? src[oG + c] : src[oG + c] + src[oG + BS + c];
      if (is_border) {
        double* B = a.brec_out + r0 + lb * BS;
        double v[HB];
#pragma unroll
        for (int r = 0; r < HB; r++) v[r] = 0.0;
        if (first) {
          for (int e = a.bsoff[p]; e < a.bsoff[p + 1]; e++) {
            const double* en = a.bent + (size_t)e * 16;
            if ((int)en[15] == myl) { const double h = en[BS + myd];
#pragma unroll
              for (int r = 0; r < HB; r++) v[r] += en[r0 + r] * h; }
          }
        } else {
          const double* Bs = a.brec + (size_t)p * (2 * BS * nb) + r0 + lb * BS;
#pragma unroll
          for (int r = 0; r < HB; r++) v[r] = Bs[r] + Bs[BS * nb + r];
        }
#pragma unroll
        for (int r = 0; r < HB; r++) B[r] = v[r];
      }
    }
    {
      double* Rp = (p >= 0) ? a.rec_out + (size_t)sg.po * REC1 : nullptr;
      double* Bp = (p >= 0) ? a.brec_out + (size_t)sg.po * (2 * BS * nb) + BS * nb : nullptr;
      auto flush_spike = [&](int x, int y, double& av) {
        const int lo = x < y ? x : y, hi = x < y ? y : x;
        if (lo < BS) {
          if (p >= 0 && hi < w) {
            const double v = -av;
            if (hi < BS) { Rp[BS * BS + lo + hi * BS] = v; Rp[BS * BS + hi + lo * BS] = v; }
            else if (hi < BS + nb) Bp[lo + (hi - BS) * BS] = v;
            else Rp[3 * BS * BS + BS + lo] = v;
          }
          av = 0.0;
        }
      };
      with_pw([&](auto PW) {
        constexpr int pwc = decltype(PW)::value;
        static_for<0, 9>([&](auto U) {
          constexpr int u = decltype(U)::value;
          if constexpr (syrk_tile_J(pwc, u) < 2) {  // only tiles whose columns reach into the spike block (columns 0..11)
            const int x = 8 * syrk_tile_I(pwc, u) + gi, y = 8 * syrk_tile_J(pwc, u) + 2 * ti;
            if (x >= y) flush_spike(x, y, acc[2 * u]); else if (x < BS || y < BS) acc[2 * u] = 0.0;
            if (x >= y + 1) flush_spike(x, y + 1, acc[2 * u + 1]); else if (x < BS || y + 1 < BS) acc[2 * u + 1] = 0.0;
          }
        });
      });
    }
    panel_sync<FUSED>();
    if (i1 >= i0) seg_base += i1 - i0 + 1;
  }
  if (nb > 0) {
    double* Cs = a.cseg + (size_t)blockIdx.x * (nb * nb + nb);
    auto flush_land = [&](int x, int y, double av) {
      const int lo = x < y ? x : y, hi = x < y ? y : x;
      if (lo >= BS && hi < w) {
        const double v = -av;
        if (hi < BS + nb) { Cs[(lo - BS) + (hi - BS) * nb] = v; Cs[(hi - BS) + (lo - BS) * nb] = v; }
        else if (lo < BS + nb) Cs[nb * nb + (lo - BS)] = v;
      }
    };
    with_pw([&](auto PW) {
      constexpr int pwc = decltype(PW)::value;
      static_for<0, 9>([&](auto U) {
        constexpr int u = decltype(U)::value;
        const int x = 8 * syrk_tile_I(pwc, u) + gi, y = 8 * syrk_tile_J(pwc, u) + 2 * ti;
        if (x >= y) flush_land(x, y, acc[2 * u]);
        if (x >= y + 1) flush_land(x, y + 1, acc[2 * u + 1]);
      });
    });
  }
}
template <int BS>
__global__ void __launch_bounds__(128, 5) k_panel4(const FwdArgs a) { panel_body<BS, false>(a, nullptr); }

// ---------------------------------------------------------------------------------------------------------------------
// Level-0 panel kernel with ACTIVE-COLUMN tracking (default for SE(3) graphs with a 64-column panel).  A segment's panel starts
// as [coupling to the left separator | 0 | 0 ...]: the columns of a landmark stay exactly zero in P, Y and S until the first
// state of the segment that observes it.  The columns are therefore ordered per segment as
//   [ spike (12) | rhs (1) | landmarks in order of first appearance in this segment (3 each) | landmarks it never meets ]
// (lorder[seg][k] = landmark of rank k, a permutation computed once at finalize: the factor structure is static), and state i
// touches only the first ntile[i] = ceil((13 + 3 seen_i) / 8) column tiles: Y = L^-1 P and P' = -Le Y on those tiles, S += Y^T Y on
// the lower-triangular tile pairs among them.  Skipped work multiplies exact zeros, so the results equal the dense kernel's
// (k_panel4) bit for bit up to the order of the landmark sums.  On C3 (16 landmarks, one range factor per two states, 46-state
// segments) this executes 52 % of the dense kernel's DMMAs.  Column tiles are dealt to the four warps round-robin (tile T ->
// warp T mod 4) and the Schur tiles by their index in the row-major enumeration of the lower triangle (t -> warp t mod 4), so the
// warps stay balanced for every ntile.  The landmark x landmark part of S is un-permuted into a shared-memory accumulator at
// the end of each segment (each entry owned by one thread: deterministic), which persists across the CTA's segments.
__host__ __device__ constexpr int tri_I(int t) { int i = 0; while ((i + 1) * (i + 2) / 2 <= t) i++; return i; }
__host__ __device__ constexpr int tri_J(int t) { return t - tri_I(t) * (tri_I(t) + 1) / 2; }
template <int BS, int MINB>
__global__ void __launch_bounds__(128, MINB) k_panel0(const FwdArgs a, const unsigned char* __restrict__ lorder, const unsigned char* __restrict__ ntile) {
  static_assert(BS == 12, "panel kernel is specialised for 12 x 12 state blocks");
  constexpr int W = 64, NT = 128, REC0 = 2 * BS * BS + BS, REC1 = 3 * BS * BS + 2 * BS, NST = 3, MAXE = 6, HB = BS / 2, C0 = BS + 1, DL = 3, LMAX = 17;
  constexpr int CSN = LMAX * DL * (LMAX * DL + 1) / 2 + LMAX * DL;   // packed lower triangle of the landmark block + landmark x rhs
  __shared__ __align__(16) double Fb[NST][2 * BS * BS];  // (L^-1 | Le)
  __shared__ __align__(16) double Gb[NST][BS];           // rhs block g
  __shared__ __align__(16) double Eb[NST][MAXE * 16];    // packed border entries
  __shared__ __align__(16) double Psm[W * BS], Ysm[2][W * BS];
  __shared__ double Csm[CSN];
  __shared__ int gdim[W];                                // physical column -> global landmark dimension (or -1), per segment
  __shared__ unsigned char nts[64];                      // active column tiles of the segment's first 64 states (a global load per state would sit on the critical path)
  const int c = threadIdx.x, pw = c >> 5, lane = c & 31, gi = lane >> 2, ti = lane & 3;
  const int ctile = pw + 4 * ((lane & 15) >> 3);         // thread-per-half-column steps: this thread's column tile, column, first row
  const int col = 8 * ctile + (lane & 7), r0 = HB * (lane >> 4);
  const int nb = a.nb, w = BS + nb + 1, nl = nb / DL;
  const bool is_rhs = (col == BS), is_spike = col < BS, is_border = (col >= C0) && (col < C0 + nb), active = col < w;
  const int myk = is_border ? (col - C0) / DL : 0, myd = is_border ? (col - C0) % DL : 0;
  double acc[18];
#pragma unroll
  for (int j = 0; j < 18; j++) acc[j] = 0.0;
  for (int k = c; k < CSN; k += NT) Csm[k] = 0.0;
  auto bso = [&](int i) { return nb > 0 ? a.bsoff[i < a.n ? i : a.n] : 0; };

  for (int seg = blockIdx.x; seg < a.nseg; seg += gridDim.x) {
    const SegGeom sg = seg_geom(seg, a.n, a.sep, a.S, a.extL, a.extR);
    const int p = sg.p, q = sg.q, i0 = sg.i0, i1 = sg.i1;
    const int ilast = (q >= 0) ? q : i1;
    const int myl = is_border ? (int)lorder[(size_t)seg * LMAX + myk] : -1;   // landmark of this thread's column in this segment
    const int mygd = is_border ? myl * DL + myd : -1;                         // its global border dimension
    if (c < W) gdim[c] = (c >= C0 && c < C0 + nb) ? (int)lorder[(size_t)seg * LMAX + (c - C0) / DL] * DL + (c - C0) % DL : -1;
    if (c >= 64 && i0 + (c - 64) <= i1) nts[c - 64] = ntile[i0 + (c - 64)];
    int b0 = bso(i0), b1 = bso(i0 + 1), b2 = bso(i0 + 2), b3 = bso(i0 + 3);
    auto prefetch = [&](int i, int st, int e0, int e1) {
      if (i <= i1) {
        const double* src = a.frec + (size_t)i * a.fstride;
        const int n2 = (((i < i1) || (q >= 0)) ? 2 * BS * BS : BS * BS) / 2;
        for (int k = c; k < n2; k += NT) cp_async16(&Fb[st][2 * k], src + 2 * k);
      }
      if (i <= ilast) {
        if (c < BS / 2) cp_async16(&Gb[st][2 * c], a.rec + (size_t)i * REC0 + 2 * BS * BS + 2 * c);
        if (nb > 0) { const int ne = min(e1 - e0, MAXE); if (c < 8 * ne) cp_async16(&Eb[st][2 * c], a.bent + (size_t)e0 * 16 + 2 * c); }
      }
    };
    // own half-column of state i: border entries of this column's landmark, rhs
    auto add_own = [&](int st, int e0, int e1) {
      double* P = Psm + col * BS + r0;
      if (is_border) {
        const int ne = e1 - e0;
        for (int k = 0; k < ne; k++) {
          const double* en = (k < MAXE) ? &Eb[st][16 * k] : a.bent + (size_t)(e0 + k) * 16;
          if ((int)en[15] == myl) { const double h = en[BS + myd];
#pragma unroll
            for (int r = 0; r < HB; r++) P[r] += en[r0 + r] * h; }
        }
      } else if (is_rhs) {
#pragma unroll
        for (int r = 0; r < HB; r++) P[r] += Gb[st][r0 + r];
      }
    };
    prefetch(i0, 0, b0, b1); cp_async_commit();
    prefetch(i0 + 1, 1, b1, b2); cp_async_commit();
    {
      const bool sp = is_spike && p >= 0 && i0 <= i1;
      const double* E = a.rec + (size_t)(sp ? p : 0) * REC0 + BS * BS + r0 + col * BS;
#pragma unroll
      for (int r = 0; r < HB; r++) Psm[col * BS + r0 + r] = sp ? E[r] : 0.0;
    }
    cp_async_wait<1>();
    __syncthreads();
    int st = 0, ys = 0, ntmax = 2;
    for (int i = i0; i <= i1; i++) {
      const bool has_next = (i < i1) || (q >= 0);
      const int nt = (i - i0 < 64) ? (int)nts[i - i0] : (int)ntile[i];   // active column tiles at this state (uniform over the CTA)
      ntmax = nt;
      const int st2 = st == 0 ? 2 : st - 1;
      prefetch(i + 2, st2, b2, b3);
      cp_async_commit();
      const int b4 = bso(i + 4);
      add_own(st, b0, b1);
      __syncwarp();
      const double* Li = Fb[st];
      const double* Le = Fb[st] + BS * BS;
      double* Y = Ysm[ys];
      // ---- Y = L^-1 P and P' = -Le Y on this warp's active column tiles (pw and pw + 4)
#pragma unroll
      for (int jt = 0; jt < 2; jt++) {
        const int J = pw + 4 * jt;
        if (J < nt) {
          double d[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
          for (int sK = 0; sK < 3; sK++) {
            const double bP = Psm[(8 * J + gi) * BS + 4 * sK + ti];
            const double a0 = Li[gi + (4 * sK + ti) * BS], a1 = (8 + gi < BS) ? Li[(8 + gi) + (4 * sK + ti) * BS] : 0.0;
            if (sK < 2) dmma884(d[0][0], d[0][1], a0, bP);   // rows 0..7 of L^-1 have no entries in columns 8..11
            dmma884(d[1][0], d[1][1], a1, bP);
          }
          Y[(8 * J + 2 * ti) * BS + gi] = d[0][0]; Y[(8 * J + 2 * ti + 1) * BS + gi] = d[0][1];
          if (8 + gi < BS) { Y[(8 * J + 2 * ti) * BS + 8 + gi] = d[1][0]; Y[(8 * J + 2 * ti + 1) * BS + 8 + gi] = d[1][1]; }
        }
      }
      __syncwarp();
      if (has_next) {
#pragma unroll
        for (int jt = 0; jt < 2; jt++) {
          const int J = pw + 4 * jt;
          if (J < nt) {
            double d[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
            for (int sK = 0; sK < 3; sK++) {
              const double bY = Y[(8 * J + gi) * BS + 4 * sK + ti];
              const double a0 = Le[gi + (4 * sK + ti) * BS], a1 = (8 + gi < BS) ? Le[(8 + gi) + (4 * sK + ti) * BS] : 0.0;
              dmma884(d[0][0], d[0][1], a0, bY);
              dmma884(d[1][0], d[1][1], a1, bY);
            }
            Psm[(8 * J + 2 * ti) * BS + gi] = -d[0][0]; Psm[(8 * J + 2 * ti + 1) * BS + gi] = -d[0][1];
            if (8 + gi < BS) { Psm[(8 * J + 2 * ti) * BS + 8 + gi] = -d[1][0]; Psm[(8 * J + 2 * ti + 1) * BS + 8 + gi] = -d[1][1]; }
          }
        }
      } else if (ctile < nt) {
#pragma unroll
        for (int r = 0; r < HB; r++) Psm[col * BS + r0 + r] = 0.0;
      }
      cp_async_wait<1>();
      __syncthreads();
      // ---- S += Y^T Y on the active lower-triangular tile pairs; this warp's tiles: t = 4 u + pw in row-major order of the triangle
      {
        double yf[8][3];
#pragma unroll
        for (int T = 0; T < 8; T++)
          if (T < nt) {
#pragma unroll
            for (int sK = 0; sK < 3; sK++) yf[T][sK] = Y[(8 * T + gi) * BS + 4 * sK + ti];
          }
        auto tiles = [&](auto PW) {
          constexpr int pwc = decltype(PW)::value;
          static_for<0, 9>([&](auto U) {
            constexpr int u = decltype(U)::value, t = 4 * u + pwc, I = tri_I(t), J = tri_J(t);
            if (I < nt) {
#pragma unroll
              for (int sK = 0; sK < 3; sK++) dmma884(acc[2 * u], acc[2 * u + 1], yf[I][sK], yf[J][sK]);
            }
          });
        };
        if (pw == 0) tiles(std::integral_constant<int, 0>{});
        else if (pw == 1) tiles(std::integral_constant<int, 1>{});
        else if (pw == 2) tiles(std::integral_constant<int, 2>{});
        else tiles(std::integral_constant<int, 3>{});
      }
      b0 = b1; b1 = b2; b2 = b3; b3 = b4;
      st = st == 2 ? 0 : st + 1; ys ^= 1;
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- segment end (D1 of q was written by k_spine): the closing separator's own border / rhs, then hand the panel off
    if (q >= 0) {
      double* R = a.rec_out + (size_t)sg.qo * REC1;
      add_own(st, b0, b1);
      const double* P = Psm + col * BS + r0;
      if (is_border) {
        double* B = a.brec_out + (size_t)sg.qo * (2 * BS * nb) + r0 + mygd * BS;
#pragma unroll
        for (int r = 0; r < HB; r++) B[r] = P[r];
      } else if (is_rhs) {
#pragma unroll
        for (int r = 0; r < HB; r++) R[3 * BS * BS + r0 + r] = P[r];
      } else if (is_spike && p >= 0) {
        double* Ep = a.rec_out + (size_t)sg.po * REC1 + 2 * BS * BS + r0 + col * BS;
        if (i0 <= i1) {
#pragma unroll
          for (int r = 0; r < HB; r++) Ep[r] = P[r];
        } else {
          const double* E = a.rec + (size_t)p * REC0 + BS * BS + r0 + col * BS;
#pragma unroll
          for (int r = 0; r < HB; r++) Ep[r] = E[r];
        }
      }
      if (a.extR && seg == a.S) {
        for (int k = c; k < BS * BS; k += NT) R[BS * BS + k] = 0.0;
        if (c < BS) R[3 * BS * BS + BS + c] = 0.0;
        if (is_border) { double* B = a.brec_out + (size_t)sg.qo * (2 * BS * nb) + BS * nb + r0 + mygd * BS;
#pragma unroll
          for (int r = 0; r < HB; r++) B[r] = 0.0; }
      }
    }
    if (a.extL && seg == 0) {  // the left external separator (halo state p): its own blocks pass through to the next level
      double* R = a.rec_out;
      const double* src = a.rec + (size_t)p * REC0;
      for (int k = c; k < BS * BS; k += NT) R[k] = src[k] + ((a.lamL && (k % (BS + 1)) == 0) ? (*a.lambda_ptr) : 0.0);
      if (c < BS) R[3 * BS * BS + c] = src[2 * BS * BS + c];
      if (is_border) {
        double* B = a.brec_out + r0 + mygd * BS;
        double v[HB];
#pragma unroll
        for (int r = 0; r < HB; r++) v[r] = 0.0;
        for (int e = a.bsoff[p]; e < a.bsoff[p + 1]; e++) {
          const double* en = a.bent + (size_t)e * 16;
          if ((int)en[15] == myl) { const double h = en[BS + myd];
#pragma unroll
            for (int r = 0; r < HB; r++) v[r] += en[r0 + r] * h; }
        }
#pragma unroll
        for (int r = 0; r < HB; r++) B[r] = v[r];
      }
    }
    {
      // flush the accumulated Y^T Y of this segment: entries touching a spike column belong to separator p (D2 | B2 | g2); the
      // rest (landmark x landmark, landmark x rhs) is un-permuted into the CTA's accumulator.  Every accumulator is reset.
      double* Rp = (p >= 0) ? a.rec_out + (size_t)sg.po * REC1 : nullptr;
      double* Bp = (p >= 0) ? a.brec_out + (size_t)sg.po * (2 * BS * nb) + BS * nb : nullptr;
      auto flush = [&](int x, int y, double& av) {   // x >= y: physical columns
        const double v = -av;
        av = 0.0;
        if (y < BS) {            // a spike column
          if (p < 0) return;
          if (x < BS) { Rp[BS * BS + y + x * BS] = v; Rp[BS * BS + x + y * BS] = v; }            // D2
          else if (x == BS) Rp[3 * BS * BS + BS + y] = v;                                        // g2
          else if (x < C0 + nb) Bp[y + gdim[x] * BS] = v;                                        // B2
        } else if (x < C0 + nb && x > BS) {
          const int gx = gdim[x];
          if (y == BS) Csm[nl * DL * (nl * DL + 1) / 2 + gx] += v;                               // landmark x rhs
          else { const int gy = gdim[y]; const int hi = gx > gy ? gx : gy, lo = gx > gy ? gy : gx; Csm[hi * (hi + 1) / 2 + lo] += v; }
        }
      };
      auto tiles = [&](auto PW) {
        constexpr int pwc = decltype(PW)::value;
        static_for<0, 9>([&](auto U) {
          constexpr int u = decltype(U)::value, t = 4 * u + pwc, I = tri_I(t), J = tri_J(t);
          // tiles that were never active hold zeros: spike-related ones are still written (B2 of unseen landmarks must read zero)
          if (I < ntmax || J < 2) {
            const int x = 8 * I + gi, y = 8 * J + 2 * ti;
            if (x >= y) flush(x, y, acc[2 * u]); else acc[2 * u] = 0.0;
            if (x >= y + 1) flush(x, y + 1, acc[2 * u + 1]); else acc[2 * u + 1] = 0.0;
          }
        });
      };
      if (pw == 0) tiles(std::integral_constant<int, 0>{});
      else if (pw == 1) tiles(std::integral_constant<int, 1>{});
      else if (pw == 2) tiles(std::integral_constant<int, 2>{});
      else tiles(std::integral_constant<int, 3>{});
    }
    __syncthreads();
  }
  if (nb > 0) {  // expand the packed accumulator into this CTA's full symmetric slot
    double* Cs = a.cseg + (size_t)blockIdx.x * (nb * nb + nb);
    for (int e = c; e < nb * nb; e += NT) { const int r = e % nb, cc = e / nb, hi = r > cc ? r : cc, lo = r > cc ? cc : r; Cs[e] = Csm[hi * (hi + 1) / 2 + lo]; }
    for (int e = c; e < nb; e += NT) Cs[nb * nb + e] = Csm[nb * (nb + 1) / 2 + e];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Level-0 panel, register-resident (A/B switch GPB_PANELW).  Two warps per segment (one 64-thread CTA); the panel P and
// Y = L^-1 P never touch shared memory as matrices: each warp keeps its four column tiles of P^T as mma A-fragments
// (pa[j][s] = P[k(s, ti)][8 T + gi], T = 2 j + warp), computes Y^T = P^T L^-T and P'^T = -Y^T Le^T with the (L^-1, Le) fragments
// loaded straight from L2 one state ahead, and turns each C-fragment back into an A/B-fragment with ONE shuffle per tile thanks
// to the k-slot order k(0, ti) = 2 ti, k(1, ti) = 2 ti + 1, k(2, ti) = {8, 10, 9, 11}[ti] (a permutation of the summation index
// applied to both operands of every product).  The only exchange is Y's fragments (12 doubles per lane and state, double
// buffered, one 64-thread barrier per state) so that both warps can form their 18 tiles of S += Y^T Y.  Column order, active
// tiles and the end-of-segment hand-off are k_panel0's.
constexpr int PW_MAXE = 6, PW_OWN = 12 + 16 * PW_MAXE;   // per-warp staging of a state's right-hand side and border entries
template <int W>
__device__ __forceinline__ void panelw_body(const FwdArgs& a, const unsigned char* __restrict__ lorder, const unsigned char* __restrict__ ntile,
                                            double* Csm, int* gdim, int* lrank, unsigned char* nts, double (*Yx)[8][3][32], double (*Own)[PW_OWN]) {
  constexpr int BS = 12, REC0 = 2 * BS * BS + BS, REC1 = 3 * BS * BS + 2 * BS, C0 = BS + 1, DL = 3, LMAX = 17, MAXE = PW_MAXE;
  const int tid = threadIdx.x, lane = tid & 31, gi = lane >> 2, ti = lane & 3;
  const int nb = a.nb, nl = nb / DL;
  const int ks[3] = {2 * ti, 2 * ti + 1, ti < 2 ? 8 + 2 * ti : 9 + 2 * (ti - 2)};
  const int shsrc = (lane & ~3) | (ti & 1);   // lane holding rows 9 / 11 of this column (second register of its rows-8..11 fragment)
  double acc[18][2];
#pragma unroll
  for (int u = 0; u < 18; u++) { acc[u][0] = 0.0; acc[u][1] = 0.0; }
  for (int seg = blockIdx.x; seg < a.nseg; seg += gridDim.x) {
    const SegGeom sg = seg_geom(seg, a.n, a.sep, a.S, a.extL, a.extR);
    const int p = sg.p, q = sg.q, i0 = sg.i0, i1 = sg.i1;
    if (tid < 64) gdim[tid] = (tid >= C0 && tid < C0 + nb) ? (int)lorder[(size_t)seg * LMAX + (tid - C0) / DL] * DL + (tid - C0) % DL : -1;
    if (tid < nl) lrank[(int)lorder[(size_t)seg * LMAX + tid]] = tid;
    if (i0 + tid <= i1) nts[tid] = ntile[i0 + tid];
    __syncthreads();
    double pa[4][3];
    {  // spike columns start as the coupling to the left separator
      const double* E = a.rec + (size_t)(p >= 0 ? p : 0) * REC0 + BS * BS;
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int s = 0; s < 3; s++) { const int col = 8 * (2 * j + W) + gi; pa[j][s] = (p >= 0 && col < BS) ? E[ks[s] + col * BS] : 0.0; }
    }
    // own part of state i: right-hand side (column 12) and the border entries of the state.  Both were staged into this warp's
    // shared-memory slot one state ahead with cp.async (a global load here would sit on the per-state critical path: in the first
    // version of this kernel a third of the stall samples were exactly that)
    auto stage_own = [&](int i, int e0, int e1, int st) {
      double* dst = Own[st];
      if (lane < BS / 2) cp_async16(dst + 2 * lane, a.rec + (size_t)i * REC0 + 2 * BS * BS + 2 * lane);
      const int ne = min(e1 - e0, MAXE);
      for (int k = lane; k < 8 * ne; k += 32) cp_async16(dst + BS + 2 * k, a.bent + (size_t)e0 * 16 + 2 * k);
    };
    auto add_own = [&](int st, int e0, int e1) {
      const double* own = Own[st];
      if (W == 1 && gi == 4) {
#pragma unroll
        for (int s = 0; s < 3; s++) pa[0][s] += own[ks[s]];
      }
      for (int e = e0; e < e1; e++) {
        const double* en = (e - e0 < MAXE) ? own + BS + 16 * (e - e0) : a.bent + (size_t)e * 16;
        const int k = lrank[(int)en[15]];
        const double a0 = en[ks[0]], a1 = en[ks[1]], a2 = en[ks[2]];
#pragma unroll
        for (int d = 0; d < DL; d++) {
          const int col = C0 + DL * k + d, T = col >> 3;
          if ((T & 1) == W && (col & 7) == gi) {
            const double h = en[BS + d];
#pragma unroll
            for (int j = 0; j < 4; j++) if (j == (T >> 1)) { pa[j][0] = fma(a0, h, pa[j][0]); pa[j][1] = fma(a1, h, pa[j][1]); pa[j][2] = fma(a2, h, pa[j][2]); }
          }
        }
      }
    };
    // fragments of (L^-1 | Le) of one state, loaded straight from L2: B[k][n] = L^-1[n][k] (rows 0..7: slices 0, 1 only - the block is
    // lower triangular; rows 8..11) and Le[n][k]
    double fL0[2], fL1[3], fE0[3], fE1[3];
    auto load_frags = [&](int i, double (&l0)[2], double (&l1)[3], double (&e0)[3], double (&e1)[3], bool want_e) {
      const double* F = a.frec + (size_t)i * a.fstride;
#pragma unroll
      for (int s = 0; s < 3; s++) {
        if (s < 2) l0[s] = F[gi + ks[s] * BS];
        l1[s] = (gi < 4) ? F[(8 + gi) + ks[s] * BS] : 0.0;
        e0[s] = want_e ? F[BS * BS + gi + ks[s] * BS] : 0.0;
        e1[s] = (want_e && gi < 4) ? F[BS * BS + (8 + gi) + ks[s] * BS] : 0.0;
      }
    };
    if (i0 <= i1) load_frags(i0, fL0, fL1, fE0, fE1, (i0 < i1) || (q >= 0));
    int b0 = nb ? a.bsoff[i0 <= a.n ? i0 : a.n] : 0, b1 = nb ? a.bsoff[i0 + 1 <= a.n ? i0 + 1 : a.n] : 0, b2 = nb ? a.bsoff[i0 + 2 <= a.n ? i0 + 2 : a.n] : 0;
    const int ilast = (q >= 0) ? q : i1;
    int par = 0, ntmax = 2, ost = 0;
    __syncwarp();
    if (i0 <= ilast) stage_own(i0, b0, b1, 0);
    cp_async_commit();
    for (int i = i0; i <= i1; i++) {
      const bool has_next = (i < i1) || (q >= 0);
      const int nt = (i - i0 < 64) ? (int)nts[i - i0] : (int)ntile[i];
      ntmax = nt;
      const int b3 = nb ? a.bsoff[i + 3 <= a.n ? i + 3 : a.n] : 0;
      double nL0[2], nL1[3], nE0[3], nE1[3];
      if (i + 1 <= i1) load_frags(i + 1, nL0, nL1, nE0, nE1, (i + 1 < i1) || (q >= 0));   // next state's fragments fly during this state
      if (i + 1 <= ilast) stage_own(i + 1, b1, b2, ost ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
      __syncwarp();
      add_own(ost, b0, b1);
      __syncwarp();
      ost ^= 1;
      double yall[8][3];
      // own tiles two at a time, slice-outer: the mma chains of the two tiles (and of the two row halves) are independent and
      // issue back to back instead of each waiting for its own accumulator
#pragma unroll
      for (int jp = 0; jp < 4; jp += 2) {
        const int TA = 2 * jp + W, TB = 2 * (jp + 1) + W;
        const bool actA = TA < nt, actB = TB < nt;
        if (actA) {
          double dA0[2] = {0.0, 0.0}, dA1[2] = {0.0, 0.0}, dB0[2] = {0.0, 0.0}, dB1[2] = {0.0, 0.0};
#pragma unroll
          for (int s = 0; s < 3; s++) {
            if (s < 2) dmma884(dA0[0], dA0[1], pa[jp][s], fL0[s]);
            dmma884(dA1[0], dA1[1], pa[jp][s], fL1[s]);
            if (actB) { if (s < 2) dmma884(dB0[0], dB0[1], pa[jp + 1][s], fL0[s]); dmma884(dB1[0], dB1[1], pa[jp + 1][s], fL1[s]); }
          }
          const double vA = __shfl_sync(0xffffffffu, dA1[1], shsrc), vB = __shfl_sync(0xffffffffu, dB1[1], shsrc);
          yall[TA][0] = dA0[0]; yall[TA][1] = dA0[1]; yall[TA][2] = ti < 2 ? dA1[0] : vA;
          yall[TB][0] = dB0[0]; yall[TB][1] = dB0[1]; yall[TB][2] = ti < 2 ? dB1[0] : vB;
#pragma unroll
          for (int s = 0; s < 3; s++) { Yx[par][TA][s][lane] = yall[TA][s]; if (actB) Yx[par][TB][s][lane] = yall[TB][s]; }
          if (has_next) {
            double eA0[2] = {0.0, 0.0}, eA1[2] = {0.0, 0.0}, eB0[2] = {0.0, 0.0}, eB1[2] = {0.0, 0.0};
#pragma unroll
            for (int s = 0; s < 3; s++) {
              dmma884(eA0[0], eA0[1], yall[TA][s], fE0[s]); dmma884(eA1[0], eA1[1], yall[TA][s], fE1[s]);
              if (actB) { dmma884(eB0[0], eB0[1], yall[TB][s], fE0[s]); dmma884(eB1[0], eB1[1], yall[TB][s], fE1[s]); }
            }
            const double wA = __shfl_sync(0xffffffffu, eA1[1], shsrc), wB = __shfl_sync(0xffffffffu, eB1[1], shsrc);
            pa[jp][0] = -eA0[0]; pa[jp][1] = -eA0[1]; pa[jp][2] = -(ti < 2 ? eA1[0] : wA);
            if (actB) { pa[jp + 1][0] = -eB0[0]; pa[jp + 1][1] = -eB0[1]; pa[jp + 1][2] = -(ti < 2 ? eB1[0] : wB); }
          } else {
#pragma unroll
            for (int s = 0; s < 3; s++) { pa[jp][s] = 0.0; pa[jp + 1][s] = 0.0; }
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int T = 2 * j + (1 - W);
        if (T < nt) {
#pragma unroll
          for (int s = 0; s < 3; s++) yall[T][s] = Yx[par][T][s][lane];
        }
      }
      // S += Y^T Y, slice-outer: this warp's active tiles (row tile I < nt; their number is monotone in u) form independent chains
      int umax = 0;
      static_for<0, 18>([&](auto U) { constexpr int u = decltype(U)::value, t = 2 * u + W; if (tri_I(t) < nt) umax = u + 1; });
#pragma unroll
      for (int s = 0; s < 3; s++)
        static_for<0, 18>([&](auto U) {
          constexpr int u = decltype(U)::value, t = 2 * u + W, I = tri_I(t), J = tri_J(t);
          if (u < umax) dmma884(acc[u][0], acc[u][1], yall[I][s], yall[J][s]);
        });
      par ^= 1;
      b0 = b1; b1 = b2; b2 = b3;
      if (i + 1 <= i1) {
#pragma unroll
        for (int s = 0; s < 3; s++) { if (s < 2) fL0[s] = nL0[s]; fL1[s] = nL1[s]; fE0[s] = nE0[s]; fE1[s] = nE1[s]; }
      }
    }
    // ---- segment end: the closing separator's own border / rhs, then hand the panel off (D1 of q was written by k_spine)
    cp_async_wait<0>();
    __syncwarp();
    if (q >= 0) {
      double* R = a.rec_out + (size_t)sg.qo * REC1;
      add_own(ost, b0, b1);
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int s = 0; s < 3; s++) {
          const int col = 8 * (2 * j + W) + gi, r = ks[s];
          if (col < BS) { if (p >= 0) a.rec_out[(size_t)sg.po * REC1 + 2 * BS * BS + r + col * BS] = pa[j][s]; }
          else if (col == BS) R[3 * BS * BS + r] = pa[j][s];
          else if (col < C0 + nb) a.brec_out[(size_t)sg.qo * (2 * BS * nb) + r + gdim[col] * BS] = pa[j][s];
        }
      if (a.extR && seg == a.S) {
        for (int k = tid; k < BS * BS; k += 64) R[BS * BS + k] = 0.0;
        if (tid < BS) R[3 * BS * BS + BS + tid] = 0.0;
        for (int k = tid; k < BS * nb; k += 64) a.brec_out[(size_t)sg.qo * (2 * BS * nb) + BS * nb + k] = 0.0;
      }
    }
    if (a.extL && seg == 0) {  // the left external separator (halo state p): its own blocks pass through to the next level
      double* R = a.rec_out;
      const double* src = a.rec + (size_t)p * REC0;
      for (int k = tid; k < BS * BS; k += 64) R[k] = src[k] + ((a.lamL && (k % (BS + 1)) == 0) ? (*a.lambda_ptr) : 0.0);
      if (tid < BS) R[3 * BS * BS + tid] = src[2 * BS * BS + tid];
      for (int k = tid; k < BS * nb; k += 64) {
        const int r = k % BS, gd = k / BS;
        double v = 0.0;
        for (int e = a.bsoff[p]; e < a.bsoff[p + 1]; e++) { const double* en = a.bent + (size_t)e * 16; if ((int)en[15] == gd / DL) v += en[r] * en[BS + gd % DL]; }
        a.brec_out[k] = v;
      }
    }
    {
      double* Rp = (p >= 0) ? a.rec_out + (size_t)sg.po * REC1 : nullptr;
      double* Bp = (p >= 0) ? a.brec_out + (size_t)sg.po * (2 * BS * nb) + BS * nb : nullptr;
      auto flush = [&](int x, int y, double& av) {   // x >= y: physical columns
        const double v = -av;
        av = 0.0;
        if (y < BS) {
          if (p < 0) return;
          if (x < BS) { Rp[BS * BS + y + x * BS] = v; Rp[BS * BS + x + y * BS] = v; }
          else if (x == BS) Rp[3 * BS * BS + BS + y] = v;
          else if (x < C0 + nb) Bp[y + gdim[x] * BS] = v;
        } else if (x < C0 + nb && x > BS) {
          const int gx = gdim[x];
          if (y == BS) Csm[nb * (nb + 1) / 2 + gx] += v;
          else { const int gy = gdim[y]; const int hi = gx > gy ? gx : gy, lo = gx > gy ? gy : gx; Csm[hi * (hi + 1) / 2 + lo] += v; }
        }
      };
      static_for<0, 18>([&](auto U) {
        constexpr int u = decltype(U)::value, t = 2 * u + W, I = tri_I(t), J = tri_J(t);
        if (I < ntmax || J < 2) {
          const int x = 8 * I + gi, y = 8 * J + 2 * ti;
          if (x >= y) flush(x, y, acc[u][0]); else acc[u][0] = 0.0;
          if (x >= y + 1) flush(x, y + 1, acc[u][1]); else acc[u][1] = 0.0;
        }
      });
    }
    __syncthreads();
  }
}
template <int BS>
__global__ void __launch_bounds__(64, 4) k_panel_w(const FwdArgs a, const unsigned char* __restrict__ lorder, const unsigned char* __restrict__ ntile) {
  static_assert(BS == 12, "panel kernel is specialised for 12 x 12 state blocks");
  constexpr int LMAX = 17, DL = 3, CSN = LMAX * DL * (LMAX * DL + 1) / 2 + LMAX * DL;
  __shared__ double Csm[CSN];
  __shared__ double Yx[2][8][3][32];
  __shared__ int gdim[64], lrank[LMAX];
  __shared__ unsigned char nts[64];
  __shared__ __align__(16) double Own[2][2][PW_OWN];   // [warp][stage]
  for (int k = threadIdx.x; k < CSN; k += 64) Csm[k] = 0.0;
  __syncthreads();
  if (threadIdx.x < 32) panelw_body<0>(a, lorder, ntile, Csm, gdim, lrank, nts, Yx, Own[0]);
  else panelw_body<1>(a, lorder, ntile, Csm, gdim, lrank, nts, Yx, Own[1]);
  __syncthreads();
  const int nb = a.nb;
  if (nb > 0) {
    double* Cs = a.cseg + (size_t)blockIdx.x * (nb * nb + nb);
    for (int e = threadIdx.x; e < nb * nb; e += 64) { const int r = e % nb, cc = e / nb, hi = r > cc ? r : cc, lo = r > cc ? cc : r; Cs[e] = Csm[hi * (hi + 1) / 2 + lo]; }
    for (int e = threadIdx.x; e < nb; e += 64) Cs[nb * nb + e] = Csm[nb * (nb + 1) / 2 + e];
  }
}

// Upper elimination levels (a few hundred segments at most: latency, not throughput): ONE kernel per level.  Warp 4 of each
// CTA walks the spine of the CTA's segments and never waits; warps 0-3 run the panel a couple of states behind it, so a level
// costs about one spine pass instead of a spine launch followed by a panel launch.
template <int BS>
__global__ void __launch_bounds__(160) k_level_ws(const FwdArgs a) {
  __shared__ int ready;
  if (threadIdx.x == 0) ready = 0;
  __syncthreads();
  if (threadIdx.x >= 128) spine_body<BS, false, true>(a, threadIdx.x - 128, &ready);
  else panel_body<BS, true>(a, &ready);
}

// The same warp-specialised kernel at level 0 (A/B switch GPB_FUSE_L0): persistent CTAs, one resident wave; the spine warp's
// (L^-1, Le) reach the panel warps through L2 (they are written to HBM anyway: the back-substitution needs them).
template <int BS>
__global__ void __launch_bounds__(160, 4) k_level0_ws(const FwdArgs a) {
  __shared__ int ready;
  if (threadIdx.x == 0) ready = 0;
  __syncthreads();
  if (threadIdx.x >= 128) spine_body<BS, true, true>(a, threadIdx.x - 128, &ready);
  else panel_body<BS, true>(a, &ready);
}

// Back-substitution of one level: x_i = L_ii^-T ( y_i - Yspike_i x_p - Yborder_i x_l - Le_i^T x_{i+1} ), right to left.
// The factor record of the next state to visit streams into shared memory (cp.async double buffer) while the current one is used.
template <int BS, int W>
__global__ void __launch_bounds__((W < 32 ? 32 : W)) k_bwd(const BwdArgs a) {
  constexpr int NT = (W < 32 ? 32 : W), NP = NT / BS, FSM = 2 * BS * BS + BS * W;
  __shared__ __align__(16) double Fb[2][FSM];
  __shared__ double coef[W], part[NP][BS], xn[BS], cv[BS];
  const int c = threadIdx.x;
  const int nb = a.nb, w = BS + nb + 1, M = a.M, fs = a.fstride;
  auto prefetch = [&](int i, int buf) {
    const double* src = a.frec + (size_t)i * fs;
    for (int k = c; k < fs / 2; k += NT) cp_async16(&Fb[buf][2 * k], src + 2 * k);
  };
  for (int seg = blockIdx.x; seg < a.nseg; seg += gridDim.x) {
    const SegGeom sg = seg_geom(seg, a.n, a.sep, a.S, a.extL, a.extR);
    const int p = sg.p, q = sg.q, i0 = sg.i0, i1 = sg.i1;
    if (i1 >= i0) prefetch(i1, 0);
    cp_async_commit();
    if (c < W) {
      double v = 0.0;
      if (c < BS) v = (p >= 0) ? -a.xup[(size_t)sg.po * BS + c] : 0.0;
      else if (c < BS + nb) v = -a.xl[c - BS];
      else if (c == BS + nb) v = 1.0;
      coef[c] = v;
    }
    if (c < BS) {
      xn[c] = (q >= 0) ? a.xup[(size_t)sg.qo * BS + c] : 0.0;
      if (q >= 0) a.xsol[(size_t)q * BS + c] = xn[c];
      if (a.extL && seg == 0) a.xsol[c] = a.xup[c];  // external left separator: copy its solution down
    }
    bool has_next = (q >= 0);
    int buf = 0;
    for (int i = i1; i >= i0; i--, buf ^= 1) {
      if (i - 1 >= i0) prefetch(i - 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      const double* F = Fb[buf];
      // partial sums of Y * coef over a slice of columns
      if (c < NP * BS) {
        const int r = c % BS, pt = c / BS;
        double s = 0.0;
        for (int col = pt; col < w; col += NP) s += F[2 * BS * BS + col * BS + r] * coef[col];
        part[pt][r] = s;
      }
      __syncthreads();
      // first warp: reduce, subtract Le^T x_next, then L^T x = cv by a column sweep from the bottom
      if (c < 32) {
        double rhs = 0.0;
        if (c < BS) {
#pragma unroll
          for (int pt = 0; pt < NP; pt++) rhs += part[pt][c];
          if (has_next) {
#pragma unroll
            for (int t = 0; t < BS; t++) rhs -= F[BS * BS + t + c * BS] * xn[t];
          }
        }
        // x = L^-T rhs: the factor record stores L^-1 (lower triangular, column-major), so this is a mat-vec
        if (c < BS) cv[c] = rhs;
        __syncwarp();
        double xr = 0.0;
        if (c < BS) {
#pragma unroll
          for (int t = 0; t < BS; t++) xr += F[t + c * BS] * cv[t];
        }
        __syncwarp();
        if (c < BS) { xn[c] = xr; a.xsol[(size_t)i * BS + c] = xr; }
      }
      has_next = true;
      __syncthreads();
    }
    cp_async_wait<0>();
    __syncthreads();
  }
}

// Back-substitution without a stored Y (default).  With the separator and landmark solutions known, the interior of a segment
// solves  A_II x_I = b_I - A_Ip x_p - A_Iq x_q - A_Il x_l  through its block-bidiagonal factor: a forward sweep
//   z_i = [g_i - B_i x_l - (i == i0) E_p x_p] - Le_{i-1} y_{i-1},   y_i = L_i^-1 z_i
// and a backward sweep  x_i = L_i^-T (y_i - Le_i^T x_{i+1}),  x_{i1+1} = x_q.  Only (L^-1 | Le) (2 BS^2 doubles per state, twice), the
// right-hand side and the SPARSE border entries of a state are read: 0.5 GB on C3 instead of the 0.8 GB of [L^-1 | Le | Y] - and the
// forward elimination no longer writes Y (0.6 GB).  One WARP per segment (the sweeps are 12-pivot-free mat-vec chains: latency,
// hidden by 16 independent warps per SM); lane r owns row r; (L^-1 | Le) stream through a per-warp cp.async ring; y_i waits in
// xsol[i] between the sweeps.  The same arithmetic as eliminating the right-hand side column with the known separator values
// substituted, i.e. the result equals the Y-based form up to rounding order.
// CTASEG (upper levels: few segments, dense border blocks): one CTA per segment - the four warps share the right-hand-side pass
// (a state each at a time), warp 0 then runs the two sweeps.  Otherwise (level 0) one warp per segment, four segments per CTA.
template <int BS, bool CTASEG>
__global__ void __launch_bounds__(128, 4) k_bwd2(const BwdArgs a) {
  constexpr int NW = 4, NST = 3, F2 = 2 * BS * BS, REC0 = 2 * BS * BS + BS, REC1 = 3 * BS * BS + 2 * BS, YC = 48, HC = BS / 2;
  constexpr int DLC = BS == 12 ? 3 : 2;   // landmark dimension of the groups with this block size (SO(3) chains carry no border)
  constexpr int NPART = 32 / BS, KMAX = (64 + NPART - 1) / NPART, KH = (KMAX + 1) / 2;
  __shared__ __align__(16) double Fb[NW][NST][F2];
  __shared__ __align__(16) double ysm[NW][YC * BS];   // right-hand sides, then y, of the segment's first YC states (later ones wait in xsol)
  __shared__ double vec[NW][BS];
  __shared__ double xls[128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = a.nb;
  const bool first = a.first_level != 0, rl = lane < BS;
  // mat-vec lanes: (row or column rr, half hh of the summation index); the two halves meet through one shuffle
  const int rr = lane % BS, hh = (lane / BS) & 1;
  const bool mv = lane < 2 * BS;
  const int RECS = first ? REC0 : REC1, oE = first ? BS * BS : 2 * BS * BS, oG = first ? 2 * BS * BS : 3 * BS * BS;
  for (int k = threadIdx.x; k < 128; k += 128) xls[k] = k < nb ? a.xl[k] : 0.0;
  __syncthreads();
  double* const v = vec[warp];
  const int yw = CTASEG ? 0 : warp;           // whose y slots / staging ring the segment uses
  const bool sweeper = !CTASEG || warp == 0;  // this warp runs the sweeps of the segment
  // y = M v (row form) or M^T v (column form) for a BS x BS column-major block in shared memory; valid on lanes < BS
  auto matvec = [&](const double* M, bool transposed) -> double {
    double s0 = 0.0, s1 = 0.0;
    if (mv) {
#pragma unroll
      for (int k = 0; k < HC; k++) {
        const int c = hh * HC + k;
        const double m = transposed ? M[c + rr * BS] : M[rr + c * BS];
        if (k & 1) s1 = fma(m, v[c], s1); else s0 = fma(m, v[c], s0);
      }
    }
    const double s = s0 + s1;
    return s + __shfl_down_sync(0xffffffffu, s, BS);
  };
  for (int seg = CTASEG ? blockIdx.x : blockIdx.x * NW + warp; seg < a.nseg; seg += CTASEG ? gridDim.x : gridDim.x * NW) {
    const SegGeom sg = seg_geom(seg, a.n, a.sep, a.S, a.extL, a.extR);
    const int p = sg.p, q = sg.q, i0 = sg.i0, i1 = sg.i1;
    double xq = 0.0, xp = 0.0;   // lane c: entry c of the right / left separator solution
    if (rl && sweeper) {
      if (q >= 0) { xq = a.xup[(size_t)sg.qo * BS + lane]; a.xsol[(size_t)q * BS + lane] = xq; }
      if (p >= 0) xp = a.xup[(size_t)sg.po * BS + lane];
      if (a.extL && seg == 0) a.xsol[lane] = a.xup[lane];  // external left separator: copy its solution down
    }
    if (i1 < i0) continue;
    auto fetch = [&](int i, int st) {
      const double* src = a.frec + (size_t)i * a.fstride;
      const int n2 = (((i < i1) || (q >= 0)) ? F2 : BS * BS) / 2;   // Le of the last interior state exists only in front of a right separator
      for (int k = lane; k < n2; k += 32) cp_async16(&Fb[yw][st][2 * k], src + 2 * k);
    };
    auto slot = [&](int i) -> double* { return (i - i0 < YC) ? &ysm[yw][(i - i0) * BS] : a.xsol + (size_t)i * BS; };
    // the (L^-1 | Le) of the first states start streaming in while the right-hand sides are formed
    // CTA per segment and the segment fits the staging buffers (NW * NST states: every upper level with M <= 12): all of its
    // (L^-1 | Le) are loaded once, by the whole CTA, and both sweeps run out of shared memory with no memory latency in the chain
    const bool resident = CTASEG && (i1 - i0 + 1) <= NW * NST;
    double* const Fall = &Fb[0][0][0];
    if (resident) {
      for (int sidx = 0; sidx <= i1 - i0; sidx++) {
        const double* src = a.frec + (size_t)(i0 + sidx) * a.fstride;
        for (int k = threadIdx.x; k < F2 / 2; k += 128) cp_async16(Fall + (size_t)sidx * F2 + 2 * k, src + 2 * k);
      }
      cp_async_commit();
    } else if (sweeper) {
      fetch(i0, 0); cp_async_commit();
      if (i0 + 1 <= i1) fetch(i0 + 1, 1);
      cp_async_commit();
    }
    // ---- pass A: o_i = g_i - B_i x_l for every interior state (no recurrence: all loads of the segment are in flight together)
    if (first) {
      // level 0: one LANE per state; the border is sparse (packed 128-byte entries, CSR by state)
      for (int i = i0 + lane + (sweeper ? 0 : a.n); i <= i1; i += 32) {
        const double* r = a.rec + (size_t)i * REC0 + oG;
        double o[BS];
#pragma unroll
        for (int k = 0; k < BS; k += 2) { const double2 t = *reinterpret_cast<const double2*>(r + k); o[k] = t.x; o[k + 1] = t.y; }
        if (nb) {
          for (int e = a.bsoff[i]; e < a.bsoff[i + 1]; e++) {
            const double* en = a.bent + (size_t)e * 16;
            double ev[16];
#pragma unroll
            for (int k = 0; k < 16; k += 2) { const double2 t = *reinterpret_cast<const double2*>(en + k); ev[k] = t.x; ev[k + 1] = t.y; }
            const int l = (int)ev[15];
            double sc = 0.0;
#pragma unroll
            for (int d = 0; d < DLC; d++) sc += ev[BS + d] * xls[l * DLC + d];
#pragma unroll
            for (int k = 0; k < BS; k++) o[k] -= ev[k] * sc;
          }
        }
        double* dst = slot(i);
#pragma unroll
        for (int k = 0; k < BS; k += 2) st128(dst + k, o[k], o[k + 1]);
      }
    } else {
      // upper levels: dense border blocks B1 + B2 (BS x nb); lane = (row r, column group part), every load of a batch issued before its FMAs
      const int r = lane % BS, part = lane / BS;
      for (int i = i0 + (CTASEG ? warp : 0); i <= i1; i += CTASEG ? NW : 1) {
        const double* B = a.brec + (size_t)i * (2 * BS * nb);
        double s0 = 0.0, s1 = 0.0;
        if (nb > 64) {   // wide borders (the 128-column panel): plain loop over this lane group's columns
          if (part < NPART) for (int l = part; l < nb; l += NPART) s0 = fma(B[r + l * BS] + B[BS * nb + r + l * BS], xls[l], s0);
        } else
#pragma unroll
        for (int h = 0; h < 2; h++) {
          double b1[KH], b2[KH];
#pragma unroll
          for (int k = 0; k < KH; k++) {
            const int l = part + NPART * (h * KH + k);
            const bool ok = part < NPART && l < nb;
            b1[k] = ok ? B[r + l * BS] : 0.0; b2[k] = ok ? B[BS * nb + r + l * BS] : 0.0;
          }
#pragma unroll
          for (int k = 0; k < KH; k++) {
            const double x = xls[(part + NPART * (h * KH + k)) & 63];   // xls is zero beyond nb
            if (k & 1) s1 = fma(b1[k] + b2[k], x, s1); else s0 = fma(b1[k] + b2[k], x, s0);
          }
        }
        double tot = s0 + s1;
        if (part >= NPART) tot = 0.0;
        double sum = 0.0;
#pragma unroll
        for (int pp = 0; pp < NPART; pp++) sum += __shfl_sync(0xffffffffu, tot, (r + pp * BS) & 31);
        if (rl) { const double* g = a.rec + (size_t)i * REC1 + oG; slot(i)[lane] = g[lane] + g[BS + lane] - sum; }
      }
    }
    if (resident) cp_async_wait<0>();
    if constexpr (CTASEG) __syncthreads(); else __syncwarp();
    if (sweeper) {
    // ---- forward sweep:  z_i = o_i - [i == i0] E_p x_p - Le_{i-1} y_{i-1},  y_i = L_i^-1 z_i
    double t = 0.0;
    if (p >= 0) {  // coupling of the first interior state to the left separator
      if (rl) v[lane] = xp;
      __syncwarp();
      t = matvec(a.rec + (size_t)p * RECS + oE, false);
      __syncwarp();
    }
    int st = 0;
    for (int i = i0; i <= i1; i++) {
      if (!resident && i + 2 <= i1) fetch(i + 2, st == 0 ? 2 : st - 1);
      cp_async_commit();
      double* ys = slot(i);
      const double own = rl ? ys[lane] : 0.0;
      cp_async_wait<2>();
      __syncwarp();
      const double* Li = resident ? Fall + (size_t)(i - i0) * F2 : Fb[yw][st];
      if (rl) v[lane] = own - t;
      __syncwarp();
      const double y = matvec(Li, false);   // the strictly-upper part of L^-1 is stored as zeros
      __syncwarp();
      if (rl) { v[lane] = y; ys[lane] = y; }
      __syncwarp();
      t = (i < i1) ? matvec(Li + BS * BS, false) : 0.0;
      st = st == NST - 1 ? 0 : st + 1;
      __syncwarp();   // every lane is done with this state's stage before a later prefetch may overwrite it
    }
    cp_async_wait<0>();
    __syncwarp();
    // ---- backward sweep:  x_i = L_i^-T (y_i - Le_i^T x_{i+1}),  x_{i1+1} = x_q
    if (!resident) {
      fetch(i1, 0); cp_async_commit();
      if (i1 - 1 >= i0) fetch(i1 - 1, 1);
      cp_async_commit();
    }
    bool hn = q >= 0;
    double xn = xq;   // lane r: entry r of x_{i+1}
    st = 0;
    for (int i = i1; i >= i0; i--) {
      if (!resident && i - 2 >= i0) fetch(i - 2, st == 0 ? 2 : st - 1);
      cp_async_commit();
      const double ycur = rl ? slot(i)[lane] : 0.0;
      cp_async_wait<2>();
      __syncwarp();
      const double* Li = resident ? Fall + (size_t)(i - i0) * F2 : Fb[yw][st];
      double w = ycur;
      if (hn) {
        if (rl) v[lane] = xn;
        __syncwarp();
        w -= matvec(Li + BS * BS, true);   // (Le^T x_{i+1})[lane]
        __syncwarp();
      }
      if (rl) v[lane] = w;
      __syncwarp();
      xn = matvec(Li, true);               // (L^-T w)[lane]
      if (rl) a.xsol[(size_t)i * BS + lane] = xn;
      hn = true;
      __syncwarp();
      st = st == NST - 1 ? 0 : st + 1;
    }
    cp_async_wait<0>();
    __syncwarp();
    }  // sweeper
    if constexpr (CTASEG) __syncthreads();   // the y slots and the staging ring are free for the CTA's next segment
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Chains of 6 x 6 blocks without a landmark border (SO(3) AHRS graphs - BASELINE config C4 -, Pose2 / Linear chains with no
// landmarks): ONE THREAD PER SEGMENT, everything in registers.  The panel is 7 columns (6 spike + the right-hand side), a state costs
// about 900 FMAs and no synchronisation at all; a million states in 26-state segments are 38 000 independent threads - one
// resident wave of the chip.  Same elimination, record formats and end-of-segment hand-off as k_fwd<6, 16> (which needed a CTA and
// ~5 us per state for the same work).  tri(r, c): packed lower triangle.
__host__ __device__ constexpr int tri(int r, int c) { return r * (r + 1) / 2 + c; }
template <bool FIRST>
__global__ void __launch_bounds__(64) k_fwd6t(const FwdArgs a) {
  constexpr int BS = 6, REC0 = 2 * BS * BS + BS, REC1 = 3 * BS * BS + 2 * BS, RECS = FIRST ? REC0 : REC1;
  constexpr int oE = FIRST ? BS * BS : 2 * BS * BS, oG = FIRST ? 2 * BS * BS : 3 * BS * BS;
  const int seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= a.nseg) return;
  const SegGeom sg = seg_geom(seg, a.n, a.sep, a.S, a.extL, a.extR);
  const int p = sg.p, q = sg.q, i0 = sg.i0, i1 = sg.i1;
  const double lambda = FIRST ? *a.lambda_ptr : 0.0;
  double Dn[21], P[7][BS], acc[27];
#pragma unroll
  for (int k = 0; k < 21; k++) Dn[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 27; k++) acc[k] = 0.0;
  {
    const bool sp = p >= 0 && i0 <= i1;
    const double* E = a.rec + (size_t)(sp ? p : 0) * RECS + oE;
#pragma unroll
    for (int c = 0; c < BS; c++)
#pragma unroll
      for (int r = 0; r < BS; r++) P[c][r] = sp ? E[r + c * BS] : 0.0;
#pragma unroll
    for (int r = 0; r < BS; r++) P[6][r] = 0.0;
  }
  bool ok = true;
  for (int i = i0; i <= i1; i++) {
    const double* rec = a.rec + (size_t)i * RECS;
    const bool has_next = (i < i1) || (q >= 0);
    double L[21], X[21], dinv[BS];
    // the record is read with 128-bit loads (a thread's records are not coalesced with its neighbours': every load instruction
    // touches 32 sectors, so their number is what the load / store unit feels)
#pragma unroll
    for (int c = 0; c < BS; c++)
#pragma unroll
      for (int r = 0; r < BS; r += 2) {
        double2 d = *reinterpret_cast<const double2*>(rec + r + c * BS);
        if constexpr (!FIRST) { const double2 d2 = *reinterpret_cast<const double2*>(rec + BS * BS + r + c * BS); d.x += d2.x; d.y += d2.y; }
        if (r >= c) L[tri(r, c)] = d.x + Dn[tri(r, c)] + (r == c ? lambda : 0.0);
        if (r + 1 >= c) L[tri(r + 1, c)] = d.y + Dn[tri(r + 1, c)] + (r + 1 == c ? lambda : 0.0);
      }
    // Cholesky, right-looking, in place
#pragma unroll
    for (int j = 0; j < BS; j++) {
      const double piv = L[tri(j, j)];
      ok &= piv > 0.0;
      const double inv = rsqrt_pos(piv > 0.0 ? piv : 1.0);
      dinv[j] = inv;
      L[tri(j, j)] = piv * inv;
#pragma unroll
      for (int r = j + 1; r < BS; r++) L[tri(r, j)] *= inv;
#pragma unroll
      for (int c = j + 1; c < BS; c++)
#pragma unroll
        for (int r = c; r < BS; r++) L[tri(r, c)] = fma(-L[tri(r, j)], L[tri(c, j)], L[tri(r, c)]);
    }
    // X = L^-1 (lower): column by column
#pragma unroll
    for (int c = 0; c < BS; c++) {
      X[tri(c, c)] = dinv[c];
#pragma unroll
      for (int r = c + 1; r < BS; r++) {
        double sacc = 0.0;
#pragma unroll
        for (int k = c; k < r; k++) sacc = fma(L[tri(r, k)], X[tri(k, c)], sacc);
        X[tri(r, c)] = -sacc * dinv[r];
      }
    }
    double* F = a.frec + (size_t)i * a.fstride;
#pragma unroll
    for (int c = 0; c < BS; c++)
#pragma unroll
      for (int r = 0; r < BS; r += 2) st128(F + c * BS + r, r >= c ? X[tri(r, c)] : 0.0, r + 1 >= c ? X[tri(r + 1, c)] : 0.0);
    double Le[BS][BS];   // Le[c][r]: column c
    if (has_next) {
      // Le L^T = E: column c of Le from columns < c
#pragma unroll
      for (int c = 0; c < BS; c++)
#pragma unroll
        for (int r = 0; r < BS; r += 2) {
          const double2 e = *reinterpret_cast<const double2*>(rec + oE + r + c * BS);
          double v0 = e.x, v1 = e.y;
#pragma unroll
          for (int k = 0; k < c; k++) { v0 = fma(-Le[k][r], L[tri(c, k)], v0); v1 = fma(-Le[k][r + 1], L[tri(c, k)], v1); }
          Le[c][r] = v0 * dinv[c]; Le[c][r + 1] = v1 * dinv[c];
        }
#pragma unroll
      for (int c = 0; c < BS; c++)
#pragma unroll
        for (int r = 0; r < BS; r += 2) st128(F + BS * BS + c * BS + r, Le[c][r], Le[c][r + 1]);
#pragma unroll
      for (int c = 0; c < BS; c++)
#pragma unroll
        for (int r = c; r < BS; r++) {
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < BS; k++) sacc = fma(Le[k][r], Le[k][c], sacc);
          Dn[tri(r, c)] = -sacc;
        }
    }
    // own right-hand side, then Y = L^-1 P in place (rows from the bottom: y_r needs p_k, k <= r)
#pragma unroll
    for (int r = 0; r < BS; r += 2) {
      double2 gq = *reinterpret_cast<const double2*>(rec + oG + r);
      if constexpr (!FIRST) { const double2 g2 = *reinterpret_cast<const double2*>(rec + oG + BS + r); gq.x += g2.x; gq.y += g2.y; }
      P[6][r] += gq.x; P[6][r + 1] += gq.y;
    }
#pragma unroll
    for (int c = 0; c < 7; c++)
#pragma unroll
      for (int r = BS - 1; r >= 0; r--) {
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k <= r; k++) sacc = fma(X[tri(r, k)], P[c][k], sacc);
        P[c][r] = sacc;
      }
    // S += Y^T Y: spike x spike (lower) and spike x rhs
#pragma unroll
    for (int x = 0; x < BS; x++) {
#pragma unroll
      for (int y = 0; y <= x; y++) {
        double sacc = acc[tri(x, y)];
#pragma unroll
        for (int k = 0; k < BS; k++) sacc = fma(P[x][k], P[y][k], sacc);
        acc[tri(x, y)] = sacc;
      }
      double sr = acc[21 + x];
#pragma unroll
      for (int k = 0; k < BS; k++) sr = fma(P[x][k], P[6][k], sr);
      acc[21 + x] = sr;
    }
    // P' = -Le Y
#pragma unroll
    for (int c = 0; c < 7; c++) {
      double t[BS];
#pragma unroll
      for (int r = 0; r < BS; r++) {
        double sacc = 0.0;
        if (has_next) {
#pragma unroll
          for (int k = 0; k < BS; k++) sacc = fma(Le[k][r], P[c][k], sacc);
        }
        t[r] = -sacc;
      }
#pragma unroll
      for (int r = 0; r < BS; r++) P[c][r] = t[r];
    }
  }
  if (!ok) *a.flag = 1;
  // ---- segment end: the same hand-off as k_fwd (nb = 0)
  if (q >= 0) {
    const double* rq = a.rec + (size_t)q * RECS;
    double* R = a.rec_out + (size_t)sg.qo * REC1;
    const double lamq = (FIRST && q < a.nreal) ? lambda : 0.0;   // a ghost is damped by its owner
#pragma unroll
    for (int c = 0; c < BS; c++)
#pragma unroll
      for (int r = 0; r < BS; r++)
        R[r + c * BS] = rq[r + c * BS] + (FIRST ? 0.0 : rq[BS * BS + r + c * BS]) + Dn[r >= c ? tri(r, c) : tri(c, r)] + (r == c ? lamq : 0.0);   // D1
#pragma unroll
    for (int r = 0; r < BS; r++) R[3 * BS * BS + r] = P[6][r] + rq[oG + r] + (FIRST ? 0.0 : rq[oG + BS + r]);                                   // g1
    if (p >= 0) {
      double* Ep = a.rec_out + (size_t)sg.po * REC1 + 2 * BS * BS;   // rows: separator q, cols: separator p
      if (i0 <= i1) {
#pragma unroll
        for (int c = 0; c < BS; c++)
#pragma unroll
          for (int r = 0; r < BS; r++) Ep[r + c * BS] = P[c][r];
      } else {  // no interior state: p and q are directly coupled
        const double* E = a.rec + (size_t)p * RECS + oE;
#pragma unroll
        for (int k = 0; k < BS * BS; k++) Ep[k] = E[k];
      }
    }
    if (a.extR && seg == a.S) {  // external right separator: no segment to its right -> its part 2 is zero
#pragma unroll
      for (int k = 0; k < BS * BS; k++) R[BS * BS + k] = 0.0;
#pragma unroll
      for (int r = 0; r < BS; r++) R[3 * BS * BS + BS + r] = 0.0;
    }
  }
  if (a.extL && seg == 0) {  // external left separator: part 1 carries this shard's own (undamped unless owned) share of it
    double* R = a.rec_out;
    const double* src = a.rec + (size_t)p * RECS;
#pragma unroll
    for (int k = 0; k < BS * BS; k++) R[k] = FIRST ? src[k] + ((a.lamL && (k % (BS + 1)) == 0) ? lambda : 0.0) : src[k] + src[BS * BS + k];
#pragma unroll
    for (int r = 0; r < BS; r++) R[3 * BS * BS + r] = FIRST ? src[oG + r] : src[oG + r] + src[oG + BS + r];
  }
  if (p >= 0) {  // Schur complement onto the left separator: D2, g2
    double* Rp = a.rec_out + (size_t)sg.po * REC1;
#pragma unroll
    for (int x = 0; x < BS; x++) {
#pragma unroll
      for (int y = 0; y <= x; y++) { Rp[BS * BS + y + x * BS] = -acc[tri(x, y)]; Rp[BS * BS + x + y * BS] = -acc[tri(x, y)]; }
      Rp[3 * BS * BS + BS + x] = -acc[21 + x];
    }
  }
}

// Back-substitution of the same chains, one thread per segment (the recurrences of k_bwd2 on 6-vectors in registers; y_i waits in xsol[i])
template <bool FIRST>
__global__ void __launch_bounds__(128) k_bwd6t(const BwdArgs a) {
  constexpr int BS = 6, REC0 = 2 * BS * BS + BS, REC1 = 3 * BS * BS + 2 * BS, RECS = FIRST ? REC0 : REC1;
  constexpr int oE = FIRST ? BS * BS : 2 * BS * BS, oG = FIRST ? 2 * BS * BS : 3 * BS * BS;
  const int seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= a.nseg) return;
  const SegGeom sg = seg_geom(seg, a.n, a.sep, a.S, a.extL, a.extR);
  const int p = sg.p, q = sg.q, i0 = sg.i0, i1 = sg.i1;
  double xq[BS], t[BS];
#pragma unroll
  for (int k = 0; k < BS; k++) { xq[k] = q >= 0 ? a.xup[(size_t)sg.qo * BS + k] : 0.0; t[k] = 0.0; }
  if (q >= 0) {
#pragma unroll
    for (int k = 0; k < BS; k++) a.xsol[(size_t)q * BS + k] = xq[k];
  }
  if (a.extL && seg == 0) {
#pragma unroll
    for (int k = 0; k < BS; k++) a.xsol[k] = a.xup[k];
  }
  if (i1 < i0) return;
  if (p >= 0) {  // E_p x_p enters the first interior state
    const double* E = a.rec + (size_t)p * RECS + oE;
#pragma unroll
    for (int c = 0; c < BS; c++) {
      const double xc = a.xup[(size_t)sg.po * BS + c];
#pragma unroll
      for (int r = 0; r < BS; r++) t[r] = fma(E[r + c * BS], xc, t[r]);
    }
  }
  for (int i = i0; i <= i1; i++) {   // forward: z = g_i - t, y = L^-1 z, t = Le y
    const double* rec = a.rec + (size_t)i * RECS;
    const double* F = a.frec + (size_t)i * a.fstride;
    double z[BS], y[BS], M[BS * BS];
#pragma unroll
    for (int k = 0; k < BS * BS; k += 2) { const double2 m = *reinterpret_cast<const double2*>(F + k); M[k] = m.x; M[k + 1] = m.y; }   // L^-1
#pragma unroll
    for (int r = 0; r < BS; r += 2) {
      double2 gq = *reinterpret_cast<const double2*>(rec + oG + r);
      if constexpr (!FIRST) { const double2 g2 = *reinterpret_cast<const double2*>(rec + oG + BS + r); gq.x += g2.x; gq.y += g2.y; }
      z[r] = gq.x - t[r]; z[r + 1] = gq.y - t[r + 1];
    }
#pragma unroll
    for (int r = 0; r < BS; r++) {
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k <= r; k++) sacc = fma(M[r + k * BS], z[k], sacc);
      y[r] = sacc;
    }
#pragma unroll
    for (int r = 0; r < BS; r += 2) st128(a.xsol + (size_t)i * BS + r, y[r], y[r + 1]);
    if (i < i1) {
#pragma unroll
      for (int k = 0; k < BS * BS; k += 2) { const double2 m = *reinterpret_cast<const double2*>(F + BS * BS + k); M[k] = m.x; M[k + 1] = m.y; }   // Le
#pragma unroll
      for (int r = 0; r < BS; r++) {
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < BS; k++) sacc = fma(M[r + k * BS], y[k], sacc);
        t[r] = sacc;
      }
    }
  }
  bool hn = q >= 0;
  for (int i = i1; i >= i0; i--) {   // backward: x = L^-T (y - Le^T x_next)
    const double* F = a.frec + (size_t)i * a.fstride;
    double w[BS], x[BS], M[BS * BS];
#pragma unroll
    for (int c = 0; c < BS; c += 2) { const double2 yv = *reinterpret_cast<const double2*>(a.xsol + (size_t)i * BS + c); w[c] = yv.x; w[c + 1] = yv.y; }
    if (hn) {
#pragma unroll
      for (int k = 0; k < BS * BS; k += 2) { const double2 m = *reinterpret_cast<const double2*>(F + BS * BS + k); M[k] = m.x; M[k + 1] = m.y; }   // Le
#pragma unroll
      for (int c = 0; c < BS; c++) {
        double sacc = w[c];
#pragma unroll
        for (int r = 0; r < BS; r++) sacc = fma(-M[r + c * BS], xq[r], sacc);
        w[c] = sacc;
      }
    }
#pragma unroll
    for (int k = 0; k < BS * BS; k += 2) { const double2 m = *reinterpret_cast<const double2*>(F + k); M[k] = m.x; M[k + 1] = m.y; }   // L^-1
#pragma unroll
    for (int c = 0; c < BS; c++) {
      double sacc = 0.0;
#pragma unroll
      for (int r = c; r < BS; r++) sacc = fma(M[r + c * BS], w[r], sacc);
      x[c] = sacc;
    }
#pragma unroll
    for (int c = 0; c < BS; c += 2) { st128(a.xsol + (size_t)i * BS + c, x[c], x[c + 1]); xq[c] = x[c]; xq[c + 1] = x[c + 1]; }
    hn = true;
  }
}

// landmark x landmark base: C0 = sum_rows l^T l (block diagonal), gl0 = sum_rows l^T rhs.  One CTA per landmark; each thread
// takes a few of the landmark's rows, the 12 sums are reduced by warp shuffles and one shared-memory pass (fixed order).
template <int NT>
__global__ void __launch_bounds__(NT) k_landmark_base(const double* __restrict__ XR, const int* __restrict__ lmoff, const int* __restrict__ lmrows,
                                                      int NXRp, int colL, int DL, int nb, double* __restrict__ Cbase) {
  __shared__ double sred[NT / 32][12];
  const int l = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double a[12];
#pragma unroll
  for (int k = 0; k < 12; k++) a[k] = 0.0;
  for (int t = lmoff[l] + threadIdx.x; t < lmoff[l + 1]; t += NT) {
    const int row = lmrows[t];
    double lv[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < 3; d++) if (d < DL) lv[d] = XR[(size_t)(colL + d) * NXRp + row];
    const double rh = XR[(size_t)(colL + DL) * NXRp + row];
#pragma unroll
    for (int d1 = 0; d1 < 3; d1++) {
#pragma unroll
      for (int d2 = 0; d2 < 3; d2++) a[d1 * 3 + d2] += lv[d1] * lv[d2];
      a[9 + d1] += lv[d1] * rh;
    }
  }
#pragma unroll
  for (int k = 0; k < 12; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_down_sync(0xffffffffu, a[k], o);
    if (lane == 0) sred[wid][k] = a[k];
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    const int k = threadIdx.x;
    double t = 0.0;
    for (int w = 0; w < NT / 32; w++) t += sred[w][k];
    if (k < 9) { const int d1 = k / 3, d2 = k % 3; if (d1 < DL && d2 < DL) Cbase[(l * DL + d1) + (size_t)(l * DL + d2) * nb] = t; }
    else { const int d1 = k - 9; if (d1 < DL) Cbase[(size_t)nb * nb + l * DL + d1] = t; }
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

// the same over every level at once (one launch per solve instead of one per level): slice s sums block b of level v when
// (b + first block of v) mod R == s, walking the levels in order - fixed order, deterministic
struct CsegLevels { const double* ptr[24]; int ncta[24]; int n; };
__global__ void k_cseg_reduce_all(const CsegLevels lv, int entries, int R, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int slice = blockIdx.y;
  if (e >= entries) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;   // four independent chains: the loads of a batch are in flight together
  int base = 0;
  for (int v = 0; v < lv.n; v++) {
    const double* c = lv.ptr[v] + e;
    const int n = lv.ncta[v];
    int b = slice - base % R; if (b < 0) b += R;
    for (; b + 3 * R < n; b += 4 * R) {
      s0 += c[(size_t)b * entries]; s1 += c[(size_t)(b + R) * entries]; s2 += c[(size_t)(b + 2 * R) * entries]; s3 += c[(size_t)(b + 3 * R) * entries];
    }
    for (; b < n; b += R) s0 += c[(size_t)b * entries];
    base += n;
  }
  out[(size_t)slice * entries + e] = (s0 + s1) + (s2 + s3);
}

// stage 2: C = Cbase + sum over all level/slice partials  (many CTAs; fixed order -> deterministic)
__global__ void k_cseg_final(const double* __restrict__ Cbase, const double* __restrict__ parts, int nparts, int entries, double* __restrict__ Csum) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= entries) return;
  double s0 = Cbase[e], s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int k = 0;
  for (; k + 3 < nparts; k += 4) { s0 += parts[(size_t)k * entries + e]; s1 += parts[(size_t)(k + 1) * entries + e]; s2 += parts[(size_t)(k + 2) * entries + e]; s3 += parts[(size_t)(k + 3) * entries + e]; }
  for (; k < nparts; k++) s0 += parts[(size_t)k * entries + e];
  Csum[e] = (s0 + s1) + (s2 + s3);
}

// Blocked Cholesky of a matrix held in shared memory (lower triangle, column-major, leading dimension ld; nrows >= R rows: row R,
// when present, is a right-hand side riding along, so the factorisation also forward-substitutes it).  8-column block steps:
//   (1) warp 0 factors the 8 x 8 diagonal block in registers (lane r = row r; pivots and column multipliers travel by shuffles),
//   (2) every row below is forward-substituted against it (one thread per row, L broadcast from shared memory),
//   (3) the trailing matrix gets C[I][J] -= X_I X_J^T as 8 x 8 tiles on the FP64 tensor pipe (two m8n8k4 per tile), tiles dealt to warps.
// Three block barriers per 8 columns instead of one per column with a whole-CTA rank-1 update in between: the sequential part
// of a 48-unknown landmark system drops from 48 to 6 steps.  dinv[j] = 1 / L[j][j].  Called by all NT threads of the CTA.
template <int NT>
__device__ __forceinline__ void chol_blocked_smem(double* sm, int ld, int R, int nrows, double* dinv, int* flag, int flagval) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gi = lane >> 2, ti = lane & 3;
  constexpr int NWARP = NT / 32;
  for (int c0 = 0; c0 < R; c0 += 8) {
    const int nc = min(8, R - c0);
    if (warp == 0) {
      // ---- (1) diagonal block; rows / columns beyond R are padded with the identity
      const int r = lane & 7;
      double a[8];
#pragma unroll
      for (int c = 0; c < 8; c++) a[c] = (r < nc && c <= r && c < nc) ? sm[(c0 + r) + (c0 + c) * ld] : ((r == c) ? 1.0 : 0.0);
      bool ok = true;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const double piv = __shfl_sync(0xffffffffu, a[j], j);
        ok &= piv > 0.0;
        const double inv = rsqrt_pos(piv > 0.0 ? piv : 1.0);
        const double lrj = (r == j) ? piv * inv : a[j] * inv;   // L[r][j], r >= j
        a[j] = lrj;
        if (lane == j && j < nc) dinv[c0 + j] = inv;
#pragma unroll
        for (int c = j + 1; c < 8; c++) {
          const double lcj = __shfl_sync(0xffffffffu, lrj, c);
          a[c] = fma(-lrj, lcj, a[c]);
        }
      }
      if (!ok && lane == 0) *flag = flagval;
      if (lane < nc) {
#pragma unroll
        for (int c = 0; c < 8; c++) if (c <= r && c < nc) sm[(c0 + r) + (c0 + c) * ld] = a[c];
      }
    }
    __syncthreads();
    // ---- (2) rows below: x L^T = a, in place
    for (int row = c0 + nc + tid; row < nrows; row += NT) {
      double x[8];
#pragma unroll
      for (int c = 0; c < 8; c++) {
        if (c < nc) {
          double v = sm[row + (c0 + c) * ld];
#pragma unroll
          for (int k = 0; k < c; k++) v = fma(-x[k], sm[(c0 + c) + (c0 + k) * ld], v);
          x[c] = v * dinv[c0 + c];
        } else x[c] = 0.0;
      }
#pragma unroll
      for (int c = 0; c < 8; c++) if (c < nc) sm[row + (c0 + c) * ld] = x[c];
    }
    __syncthreads();
    // ---- (3) trailing update on 8 x 8 tiles (I >= J) of the rows / columns behind this block column
    const int b0 = c0 + nc;                       // first trailing row / column
    const int ntr = (nrows - b0 + 7) / 8;         // row tiles (the rhs row included)
    const int ntc = (R - b0 + 7) / 8;             // column tiles
    if (ntc > 0) {
      const int ntiles = ntc * (ntc + 1) / 2 + (ntr - ntc) * ntc;   // lower triangle of the square part + the rows below it
      for (int t = warp; t < ntiles; t += NWARP) {
        int I, J;
        const int tri_n = ntc * (ntc + 1) / 2;
        if (t < tri_n) { I = 0; while ((I + 1) * (I + 2) / 2 <= t) I++; J = t - I * (I + 1) / 2; }
        else { const int u = t - tri_n; I = ntc + u / ntc; J = u % ntc; }
        const int ra = b0 + 8 * I + gi, rb = b0 + 8 * J + gi;      // fragment rows of X_I (A operand) and X_J (B operand: B[k][n] = X_J[n][k])
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int sK = 0; sK < 2; sK++) {
          const int k = 4 * sK + ti;
          const double fa = (ra < nrows && k < nc) ? sm[ra + (c0 + k) * ld] : 0.0;
          const double fb = (rb < R && k < nc) ? sm[rb + (c0 + k) * ld] : 0.0;
          dmma884(d0, d1, fa, fb);
        }
        const int cr = b0 + 8 * I + gi, cc = b0 + 8 * J + 2 * ti;   // C fragment: (cr, cc), (cr, cc + 1)
        if (cr < nrows) {
          if (cc < R && cc <= cr) sm[cr + cc * ld] -= d0;
          if (cc + 1 < R && cc + 1 <= cr) sm[cr + (cc + 1) * ld] -= d1;
        }
      }
    }
    __syncthreads();
  }
}

// Small dense SPD solve in shared memory, single CTA:  (A + lambda * diag[loff..R)) x = rhs,  R <= SMALL_SOLVE_MAX.
// Used for the landmark system (no pinned states; loff = 0) and for reduced systems of a few separators + landmarks (sharded
// graphs, a handful of loop closures).  The right-hand side rides along as row R of the lower triangle, so the Cholesky
// factorisation performs the forward substitution; one block barrier per column (the trailing update reads the unscaled
// pivot column and scales on the fly, the column itself is scaled afterwards by other threads).  Back-substitution is done by
// one warp, row-oriented, the solution entries living in registers (entry r on lane r mod 32).
// KB > 0 (R + 1 <= 16 KB; instantiated for KB = 2, 4, 6, 9): the trailing update of a column is register-blocked - a thread's (up to KB x KB) targets, the 2 KB - 1
// pivot-column entries and KB multipliers they need are all loaded before the first FMA, so the column costs one shared-memory
// latency instead of KB^2 dependent load-FMA-store round trips (R = 48, the landmark system of C3: 44 us -> see DESIGN.md).
// KB = 0: plain loops, any R <= SMALL_SOLVE_MAX.  Same operations on the same operands in every instantiation.
constexpr int SMALL_SOLVE_MAX = 160;
template <int NT, int KB = 0>
__global__ void __launch_bounds__(NT) k_small_solve(const double* __restrict__ A, int lda, const double* rhs, int rstride, int R, int loff,
                                                    const double* __restrict__ lambda_ptr, double* x, int* __restrict__ flag, int flagval) {
  extern __shared__ double sm[];
  const int ld = R + 1 + ((R & 1) ? 1 : 0);  // odd leading dimension: row walks are bank-conflict free
  double* dinv = sm + (size_t)ld * R;        // 1 / L[j][j]
  const double lambda = *lambda_ptr;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, TY = NT / 16;
  for (int c = ty; c < R; c += TY) {
    for (int r = c + tx; r < R; r += 16) sm[r + c * ld] = A[r + (size_t)c * lda] + ((r == c && r >= loff) ? lambda : 0.0);
    if (tx == 0) sm[R + c * ld] = rhs[(size_t)c * rstride];
  }
  __syncthreads();
  if constexpr (KB < 0) {   // blocked factorisation: 8-column steps, tensor-pipe trailing updates
    chol_blocked_smem<NT>(sm, ld, R, R + 1, dinv, flag, flagval);
  } else
  for (int j = 0; j < R; j++) {
    const double djj = sm[j + j * ld];
    if (!(djj > 0.0) && tid == 0) *flag = flagval;
    const double inv = rsqrt_pos(djj > 0.0 ? djj : 1.0);
    if constexpr (KB > 0) {
      static_assert(NT == 256, "16 x 16 thread tile");
      // targets (r, cc) = (j + 1 + ty + tx + 16 (a + b), j + 1 + ty + 16 a): pivot-column rows depend on a + b only
      const int base = j + 1 + ty;
      double pv[2 * KB - 1], lc[KB], tv[KB][KB];
#pragma unroll
      for (int m = 0; m < 2 * KB - 1; m++) { const int r = base + tx + 16 * m; pv[m] = r <= R ? sm[r + j * ld] : 0.0; }
#pragma unroll
      for (int a = 0; a < KB; a++) { const int cc = base + 16 * a; lc[a] = cc < R ? sm[cc + j * ld] : 0.0; }
#pragma unroll
      for (int a = 0; a < KB; a++)
#pragma unroll
        for (int b = 0; b + a < KB; b++) { const int cc = base + 16 * a, r = cc + tx + 16 * b; tv[a][b] = (cc < R && r <= R) ? sm[r + cc * ld] : 0.0; }
#pragma unroll
      for (int a = 0; a < KB; a++)
#pragma unroll
        for (int b = 0; b + a < KB; b++) {
          const int cc = base + 16 * a, r = cc + tx + 16 * b;
          if (cc < R && r <= R) sm[r + cc * ld] = fma(-(pv[a + b] * inv), lc[a] * inv, tv[a][b]);
        }
    } else {
      for (int cc = j + 1 + ty; cc < R; cc += TY) {
        const double lc = sm[cc + j * ld] * inv;
        for (int r = cc + tx; r <= R; r += 16) sm[r + cc * ld] = fma(-(sm[r + j * ld] * inv), lc, sm[r + cc * ld]);
      }
    }
    __syncthreads();
    for (int r = j + tid; r <= R; r += NT) sm[r + j * ld] *= inv;
    if (tid == 0) dinv[j] = inv;
  }
  __syncthreads();
  if (tid < 32) {
    constexpr int KMAX = SMALL_SOLVE_MAX / 32;
    double yv[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; k++) { const int r = tid + 32 * k; yv[k] = r < R ? sm[R + r * ld] : 0.0; }
    for (int j = R - 1; j >= 0; j--) {
      double mine = 0.0;
#pragma unroll
      for (int k = 0; k < KMAX; k++) if ((j >> 5) == k) mine = yv[k];
      const double xj = __shfl_sync(0xffffffffu, mine, j & 31) * dinv[j];
#pragma unroll
      for (int k = 0; k < KMAX; k++) {
        const int r = tid + 32 * k;
        if (r < j) yv[k] = fma(-sm[j + r * ld], xj, yv[k]);
        else if (r == j) yv[k] = xj;
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; k++) { const int r = tid + 32 * k; if (r < R) x[r] = yv[k]; }
  }
}
static size_t small_solve_smem(int R) { return ((size_t)(R + 2) * R + R) * sizeof(double); }
// the instantiation that fits R (allow_blocked = false: the plain-loop form, for A/B and tests)
// mode 0 (default): blocked factorisation (8-column steps, tensor-pipe trailing updates); 1: the per-column kernel with a register-
// blocked trailing update (round 1; A/B switch GPB_OLD_TINY); 2: the per-column kernel with plain loops (tests)
static void launch_small_solve(cudaStream_t stream, const double* A, int lda, const double* rhs, int rstride, int R, int loff, const double* lambda_ptr, double* x,
                               int* flag, int flagval, int mode = 0) {
  if (mode == 0) k_small_solve<256, -1><<<1, 256, small_solve_smem(R), stream>>>(A, lda, rhs, rstride, R, loff, lambda_ptr, x, flag, flagval);
  else if (mode == 1 && R + 1 <= 16 * 2) k_small_solve<256, 2><<<1, 256, small_solve_smem(R), stream>>>(A, lda, rhs, rstride, R, loff, lambda_ptr, x, flag, flagval);
  else if (mode == 1 && R + 1 <= 16 * 4) k_small_solve<256, 4><<<1, 256, small_solve_smem(R), stream>>>(A, lda, rhs, rstride, R, loff, lambda_ptr, x, flag, flagval);
  else if (mode == 1 && R + 1 <= 16 * 6) k_small_solve<256, 6><<<1, 256, small_solve_smem(R), stream>>>(A, lda, rhs, rstride, R, loff, lambda_ptr, x, flag, flagval);
  else if (mode == 1 && R + 1 <= 16 * 9) k_small_solve<256, 9><<<1, 256, small_solve_smem(R), stream>>>(A, lda, rhs, rstride, R, loff, lambda_ptr, x, flag, flagval);  // 8 shards: R = 7 * 12 + 48 = 132
  else k_small_solve<256, 0><<<1, 256, small_solve_smem(R), stream>>>(A, lda, rhs, rstride, R, loff, lambda_ptr, x, flag, flagval);
}


// x <- x (+) delta for every state (Pose3 / Rot3: Expmap; Pose2: GTSAM's default chart; vectors: add), plus the two dot
// products LM needs: g.delta and |delta|^2 (block partials).
template <int G, int NT>
__global__ void __launch_bounds__(NT) k_retract(const double* __restrict__ X, const double* __restrict__ xsol, const double* __restrict__ HREC,
                                                double* __restrict__ Xt, double* __restrict__ part_gd, double* __restrict__ part_dd, int N, int dd_from, int dd_to) {
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
    for (int k = 0; k < bs; k++) { gd += g[k] * d[k]; if (i >= dd_from && i < dd_to) dd += d[k] * d[k]; }  // a halo / ghost copy's |delta|^2 is counted by its owner
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

// values cross the C ABI as separate pose / velocity arrays; HBM holds one AoS record per state.  dir 0: arrays -> records, 1: back.
__global__ void k_pack_values(double* __restrict__ poses, double* __restrict__ vels, double* __restrict__ X, int N, int PS, int D, int dir) {
  const int SR = PS + D;
  const size_t total = (size_t)N * SR;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t i = e / SR; const int k = (int)(e % SR);
    double* a = k < PS ? poses + i * PS + k : vels + i * D + (k - PS);
    if (dir == 0) X[e] = *a; else *a = X[e];
  }
}
__global__ void k_retract_land(const double* __restrict__ land, const double* __restrict__ xl, const double* __restrict__ gl, double* __restrict__ landt,
                               int n, double* __restrict__ scal, int count_dd) {
  // single block: landmarks are few; also folds their share of g.delta / |delta|^2 into scal[1], scal[2]
  __shared__ double sred[8];
  double gd = 0, dd = 0;
  for (int k = threadIdx.x; k < n; k += 256) { landt[k] = land[k] + xl[k]; gd += gl[k] * xl[k]; if (count_dd) dd += xl[k] * xl[k]; }
  const double t1 = block_sum<256>(gd, sred);
  __syncthreads();
  const double t2 = block_sum<256>(dd, sred);
  if (threadIdx.x == 0) { scal[1] += t1; scal[2] += t2; }
}

// ===================================================================== the reduced ("top") system
// Unknowns: [top states (bs each) | landmarks (nb)], R = ntop * bs + nb.  Top states are the pinned states of the chain: the
// external separators of a shard (its halo and its last state) and the endpoints of loop closures - states the level recursion
// never eliminates.  Buffer: T ((R+1) x R column-major, leading dimension R+1: rows 0..R-1 the symmetric matrix, row R the
// right-hand side) | scalars[4] = {local error sum, not-PD flag sum, unused, unused}.  A graph adds the Schur complement of its
// chain (top-level records of its pinned states at their global top indices gtop[], its landmark block, its loop-closure cross
// blocks); on sharded graphs ONE all-reduce (sum) over NVLink completes the system and every rank factors it redundantly.
struct PackArgs {
  int bs, nb, R, ntop, P;   // ntop: top states of the whole (global) system; P: this graph's top-level chain length
  const int* gtop;          // [P] global top index of local top state k
  const double* rec;        // top-level records [P][3 bs^2 + 2 bs]
  const double* brec;       // [P][2 bs nb]
  const double* Csum;       // local landmark block nb*nb + nb
  double err_local;
  const double* err_ptr;    // when non-null the local error is read from device memory (asynchronous Gauss-Newton) instead of err_local
  const int* flag;
  double* buf;              // zero-filled before the launch
};
// one CTA per local top state (+ one for the landmark block and the scalars)
__global__ void k_pack_top(const PackArgs a) {
  const int R = a.R, ld = R + 1, bs = a.bs, nb = a.nb, REC1 = 3 * bs * bs + 2 * bs, loff = a.ntop * bs;
  const int k = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
  double* T = a.buf;
  if (k < a.P) {
    const double* rec = a.rec + (size_t)k * REC1;
    const int o = a.gtop[k] * bs;
    for (int e = tid; e < bs * bs; e += NT) { const int r = e % bs, c = e / bs; T[(o + r) + (size_t)(o + c) * ld] = rec[e] + rec[bs * bs + e]; }
    if (k + 1 < a.P) {  // E: rows top state k+1, cols top state k (the segment between them wrote it)
      const int o1 = a.gtop[k + 1] * bs;
      for (int e = tid; e < bs * bs; e += NT) { const int r = e % bs, c = e / bs; const double v = rec[2 * bs * bs + e]; T[(o1 + r) + (size_t)(o + c) * ld] = v; T[(o + c) + (size_t)(o1 + r) * ld] = v; }
    }
    const double* B = a.brec + (size_t)k * (2 * bs * nb);
    for (int e = tid; e < bs * nb; e += NT) { const int r = e % bs, l = e / bs; const double v = B[e] + B[bs * nb + e]; T[(o + r) + (size_t)(loff + l) * ld] = v; T[(loff + l) + (size_t)(o + r) * ld] = v; }
    for (int r = tid; r < bs; r += NT) T[R + (size_t)(o + r) * ld] = rec[3 * bs * bs + r] + rec[3 * bs * bs + bs + r];
  } else {
    for (int e = tid; e < nb * nb; e += NT) { const int r = e % nb, c = e / nb; T[(loff + r) + (size_t)(loff + c) * ld] = a.Csum[e]; }
    for (int r = tid; r < nb; r += NT) T[R + (size_t)(loff + r) * ld] = a.Csum[(size_t)nb * nb + r];
    if (tid == 0) { double* sc = T + (size_t)ld * R; sc[0] = a.err_ptr ? *a.err_ptr : a.err_local; sc[1] = (double)(*a.flag); sc[2] = 0.0; sc[3] = 0.0; }
  }
}

// Loop closures (BetweenFactor between non-adjacent states i < j; their whitened rows sit at the tail of the extra-row table,
// a-part = state i, b-part = state j, pose columns only).  Diagonal share: one CTA per endpoint state adds sum A_s^T A_s and
// sum A_s^T b into the state's assembled record (after k_assemble; fixed order -> deterministic).
__global__ void k_assemble_closures(const double* __restrict__ XR, int NXRp, int bs, int m, int xrhs, const int* __restrict__ epstate, const int* __restrict__ epoff,
                                    const int* __restrict__ eprow, const int* __restrict__ epside, double* __restrict__ HREC) {
  const int s = epstate[blockIdx.x], REC = 2 * bs * bs + bs, tid = threadIdx.x;
  double* rec = HREC + (size_t)s * REC;
  for (int e = tid; e < bs * bs + bs; e += blockDim.x) {
    const bool isg = e >= bs * bs;
    const int r = isg ? e - bs * bs : e % bs, c = isg ? 0 : e / bs;
    double acc = 0.0;
    for (int t = epoff[blockIdx.x]; t < epoff[blockIdx.x + 1]; t++) {
      const int row0 = eprow[t], o = epside[t] * bs;
      for (int q = 0; q < m; q++) {
        const double ar = XR[(size_t)(o + r) * NXRp + row0 + q];
        acc += ar * (isg ? XR[(size_t)xrhs * NXRp + row0 + q] : XR[(size_t)(o + c) * NXRp + row0 + q]);
      }
    }
    rec[isg ? 2 * bs * bs + r : e] += acc;
  }
}
// Cross share: one CTA per unique endpoint pair adds sum A_j^T A_i at (rows top(j), cols top(i)) of the packed top system and
// its transpose (+= : a pair adjacent in the top chain already holds its chain coupling there).
__global__ void k_pack_closures(const double* __restrict__ XR, int NXRp, int bs, int m, int ld, const int* __restrict__ pair_a, const int* __restrict__ pair_b,
                                const int* __restrict__ pairoff, const int* __restrict__ pairrow, double* __restrict__ T) {
  const int oa = pair_a[blockIdx.x] * bs, ob = pair_b[blockIdx.x] * bs;  // global top offsets of the a-side (state i) and b-side (state j)
  for (int e = threadIdx.x; e < bs * bs; e += blockDim.x) {
    const int r = e % bs, c = e / bs;  // r: column of the b-part, c: column of the a-part
    double acc = 0.0;
    for (int t = pairoff[blockIdx.x]; t < pairoff[blockIdx.x + 1]; t++) {
      const int row0 = pairrow[t];
      for (int q = 0; q < m; q++) acc += XR[(size_t)(bs + r) * NXRp + row0 + q] * XR[(size_t)c * NXRp + row0 + q];
    }
    T[(ob + r) + (size_t)(oa + c) * ld] += acc;
    T[(oa + c) + (size_t)(ob + r) * ld] += acc;
  }
}

// scatter the reduced solution: this graph's top states -> top-level xsol, landmarks -> xl
__global__ void k_top_scatter(const double* __restrict__ x, int bs, int nb, int ntop, int P, const int* __restrict__ gtop, double* __restrict__ xsol_top,
                              double* __restrict__ xl) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < P * bs) xsol_top[t] = x[gtop[t / bs] * bs + t % bs];
  if (t < nb) xl[t] = x[ntop * bs + t];
}

// ---- blocked dense Cholesky solve for reduced systems beyond the shared-memory solver (many loop closures).  Right-looking,
// block size 64, in place on the lower triangle of T (leading dimension ld = R + 1; row R = right-hand side, so the
// factorisation also performs the forward substitution).  Per block column: k_dense_diag (one CTA: Cholesky of the 64 x 64
// diagonal tile) -> k_dense_trsm (one thread per row below: row <- row * L_jj^-T) -> k_dense_syrk (64 x 64 tiles of the trailing
// matrix, 4 x 4 register micro-tiles).  Back-substitution walks the block columns in reverse (k_dense_bwd).
constexpr int DNB = 64;
__global__ void __launch_bounds__(256) k_dense_diag(double* __restrict__ T, int ld, int R, int j0, int loff, const double* __restrict__ lambda_ptr, int* __restrict__ flag) {
  __shared__ double sm[DNB * (DNB + 1)];
  __shared__ double dinv_s[DNB];
  const int n = min(DNB, R - j0), tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, ldt = DNB + 1;
  const double lambda = *lambda_ptr;
  // the LM damping of the landmark diagonal is applied when a diagonal tile is first touched (entries were only updated, never read, before)
  for (int c = ty; c < n; c += 16) for (int r = c + tx; r < n; r += 16) sm[r + c * ldt] = T[(j0 + r) + (size_t)(j0 + c) * ld] + ((r == c && j0 + r >= loff) ? lambda : 0.0);
  __syncthreads();
  chol_blocked_smem<256>(sm, ldt, n, n, dinv_s, flag, 3);   // 8-column block steps (was: one block barrier and a whole-CTA rank-1 update per column)
  for (int c = ty; c < n; c += 16) for (int r = c + tx; r < n; r += 16) T[(j0 + r) + (size_t)(j0 + c) * ld] = sm[r + c * ldt];
}
// rows j0+n .. R (inclusive: the rhs row) of block column j0: X L^T = A, one row per thread, L_jj broadcast from shared memory
__global__ void __launch_bounds__(128) k_dense_trsm(double* __restrict__ T, int ld, int R, int j0) {
  __shared__ double L[DNB * DNB], dinv[DNB];
  const int n = min(DNB, R - j0), tid = threadIdx.x;
  for (int e = tid; e < n * n; e += 128) { const int r = e % n, c = e / n; L[r + c * DNB] = (r >= c) ? T[(j0 + r) + (size_t)(j0 + c) * ld] : 0.0; }
  __syncthreads();
  if (tid < n) dinv[tid] = 1.0 / L[tid + tid * DNB];
  __syncthreads();
  const int row = j0 + n + blockIdx.x * 128 + tid;
  if (row > R) return;
  double x[DNB];
  if (n == DNB) {
#pragma unroll
    for (int c = 0; c < DNB; c++) {
      double v = T[row + (size_t)(j0 + c) * ld];
#pragma unroll
      for (int k = 0; k < c; k++) v = fma(-x[k], L[c + k * DNB], v);
      x[c] = v * dinv[c];
    }
#pragma unroll
    for (int c = 0; c < DNB; c++) T[row + (size_t)(j0 + c) * ld] = x[c];
  } else {  // ragged last block column: in place through global memory
    for (int c = 0; c < n; c++) {
      double v = T[row + (size_t)(j0 + c) * ld];
      for (int k = 0; k < c; k++) v = fma(-T[row + (size_t)(j0 + k) * ld], L[c + k * DNB], v);
      T[row + (size_t)(j0 + c) * ld] = v * dinv[c];
    }
  }
}
// trailing update C[I][J] -= A_I A_J^T over the rows / columns beyond block column j0 (rows up to R inclusive, columns < R), I >= J
__global__ void __launch_bounds__(256) k_dense_syrk(double* __restrict__ T, int ld, int R, int j0) {
  if (blockIdx.y > blockIdx.x) return;
  __shared__ double As[16][DNB + 1], Bs[16][DNB + 1];
  const int n = min(DNB, R - j0), base = j0 + n;
  const int r0 = base + blockIdx.x * DNB, c0 = base + blockIdx.y * DNB;  // tile origin (rows, cols)
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < n; k0 += 16) {
    for (int e = tid; e < 16 * DNB; e += 256) {
      const int rr = e % DNB, kk = e / DNB;
      const bool kv = k0 + kk < n;
      As[kk][rr] = (kv && r0 + rr <= R) ? T[(r0 + rr) + (size_t)(j0 + k0 + kk) * ld] : 0.0;
      Bs[kk][rr] = (kv && c0 + rr <= R) ? T[(c0 + rr) + (size_t)(j0 + k0 + kk) * ld] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
      double av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { av[i] = As[kk][tx + 16 * i]; bv[i] = Bs[kk][ty + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = r0 + tx + 16 * i, c = c0 + ty + 16 * j;
      if (r <= R && c < R && r >= c) T[r + (size_t)c * ld] -= acc[i][j];
    }
}
// The same trailing update on the FP64 tensor pipe (default; north_star: "tensor cores ... where that panel is a genuine dense
// contraction" - this SYRK over 64-column panels is the one place in the path that is).  CTA = 4 warps, one 64 x 64 tile of the
// trailing matrix, K = 64 staged through shared memory in 16-deep chunks; warp w owns row tiles 2w, 2w+1 x the 8 column tiles
// (16 accumulator fragments); per k-slice 2 + 8 fragment loads feed 16 mma.sync.m8n8k4.f64.
__global__ void __launch_bounds__(128) k_dense_syrk_mma(double* __restrict__ T, int ld, int R, int j0) {
  if (blockIdx.y > blockIdx.x) return;
  constexpr int KC = 16, LDS_ = DNB + 8;   // row stride 72 doubles: the four k rows of a fragment load fall into two bank halves
  __shared__ double As[KC][LDS_], Bs[KC][LDS_];
  const int n = min(DNB, R - j0), base = j0 + n;
  const int r0 = base + blockIdx.x * DNB, c0 = base + blockIdx.y * DNB;  // tile origin (rows, cols)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gi = lane >> 2, ti = lane & 3;
  double acc[2][8][2];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  for (int k0 = 0; k0 < n; k0 += KC) {
    for (int e = tid; e < KC * DNB; e += 128) {
      const int rr = e % DNB, kk = e / DNB;
      const bool kv = k0 + kk < n;
      As[kk][rr] = (kv && r0 + rr <= R) ? T[(r0 + rr) + (size_t)(j0 + k0 + kk) * ld] : 0.0;
      Bs[kk][rr] = (kv && c0 + rr < R) ? T[(c0 + rr) + (size_t)(j0 + k0 + kk) * ld] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int sK = 0; sK < KC / 4; sK++) {
      const double a0 = As[4 * sK + ti][8 * (2 * warp) + gi], a1 = As[4 * sK + ti][8 * (2 * warp + 1) + gi];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const double b = Bs[4 * sK + ti][8 * j + gi];   // B[k][n] = A_J[n][k]
        dmma884(acc[0][j][0], acc[0][j][1], a0, b);
        dmma884(acc[1][j][0], acc[1][j][1], a1, b);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int r = r0 + 8 * (2 * warp + i) + gi, c = c0 + 8 * j + 2 * ti + h;
        if (r <= R && c < R && r >= c) T[r + (size_t)c * ld] -= acc[i][j][h];
      }
}
// back-substitution step for block column j0: x_j = L_jj^-T y_j (every CTA, redundantly, by its first warp), then
// y_c -= L[j-rows][c]^T x_j for the columns c < j0 (one per thread).  y lives in row R of T; x goes to xout.
__global__ void __launch_bounds__(256) k_dense_bwd(double* __restrict__ T, int ld, int R, int j0, double* __restrict__ xout) {
  __shared__ double L[DNB * (DNB + 1)], xs[DNB];
  const int n = min(DNB, R - j0), tid = threadIdx.x, ldt = DNB + 1;
  for (int e = tid; e < n * n; e += 256) { const int r = e % n, c = e / n; L[r + c * ldt] = (r >= c) ? T[(j0 + r) + (size_t)(j0 + c) * ld] : 0.0; }
  __syncthreads();
  if (tid < 32) {
    double y0 = tid < n ? T[R + (size_t)(j0 + tid) * ld] : 0.0, y1 = tid + 32 < n ? T[R + (size_t)(j0 + tid + 32) * ld] : 0.0;
    for (int r = n - 1; r >= 0; r--) {
      const double xr = __shfl_sync(0xffffffffu, r < 32 ? y0 : y1, r & 31) / L[r + r * ldt];
      // row r of L: L[r][k], k < r
      if (tid < r) y0 = fma(-L[r + tid * ldt], xr, y0); else if (tid == r) y0 = xr;
      if (tid + 32 < r) y1 = fma(-L[r + (tid + 32) * ldt], xr, y1); else if (tid + 32 == r) y1 = xr;
    }
    if (tid < n) xs[tid] = y0;
    if (tid + 32 < n) xs[tid + 32] = y1;
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid < n) xout[j0 + tid] = xs[tid];
  const int c = blockIdx.x * 256 + tid;
  if (c < j0) {
    const double* col = T + (size_t)c * ld + j0;
    double acc = 0.0;
    for (int r = 0; r < n; r++) acc = fma(col[r], xs[r], acc);
    T[R + (size_t)c * ld] -= acc;
  }
}

__global__ void k_set_scalar(double* p, double v) { *p = v; }
__global__ void k_clear_flag(int* f) { *f = 0; }
