// Normal-equation assembly kernel: H = A^T A restricted to the chain (see engine.cu for layouts).
#pragma once
#include "kernels_lin.cuh"

// D_i (bs x bs), E_i = H_{i+1,i} (rows: state i+1, cols: state i), g_i = (A^T b)_i, all column-major inside the state's
// HREC record [D | E | g].  One thread per (state, tile): bs = 12 -> four 12x6 register tiles per state (D left/right
// + g, E left/right); bs = 6 -> one thread per state.  Operands stream from the SoA [A|b] (coalesced 128-bit loads:
// consecutive threads of a tile class read consecutive factors) and from the extra-row table; results leave as 128-bit
// stores.  Landmark columns of the extra rows are not touched here (they form the border, see the solver).
template <int G, int NT>
__global__ void __launch_bounds__(NT) k_assemble(const double* __restrict__ AB, const double* __restrict__ dt, const double* __restrict__ XR,
                                                 const int* __restrict__ rowoff, double* __restrict__ HREC, int N, int NFp, int NXRp) {
  constexpr int D = GroupTraits<G>::D, DL = GroupTraits<G>::DL, bs = 2 * D, REC = 2 * bs * bs + bs;
  constexpr int TILES = (bs == 12) ? 4 : 1;
  constexpr int XRHS = 2 * bs + DL;
  // tile-major thread order inside the grid: threads of the same tile class are contiguous, so a warp reads consecutive factors
  const int gid = blockIdx.x * NT + threadIdx.x;
  const int Npad = (N + 31) & ~31;  // tile classes start on a warp boundary
  const int tile = gid / Npad, i = gid % Npad;
  if (tile >= TILES || i >= N) return;
  const int nint = N - 1;
  const bool doD = (TILES == 1) || tile < 2;
  const bool doE = (TILES == 1) || tile >= 2;
  const bool doG = (TILES == 1) || tile == 0;
  const int c0 = (TILES == 1) ? 0 : (tile & 1) * 6;
  double acc[bs][6];   // D tile (or E tile when this thread only does E)
  double acc2[TILES == 1 ? bs : 1][6];  // E tile for the single-thread-per-state case
  double g[bs];
#pragma unroll
  for (int r = 0; r < bs; r++) {
    g[r] = 0;
#pragma unroll
    for (int c = 0; c < 6; c++) acc[r][c] = 0;
  }
  if constexpr (TILES == 1) {
#pragma unroll
    for (int r = 0; r < bs; r++)
#pragma unroll
      for (int c = 0; c < 6; c++) acc2[r][c] = 0;
  }
  auto ld2 = [&](int col, int rp, int f) { return *reinterpret_cast<const double2*>(AB + ((size_t)(col * D + rp) * NFp + f) * 2); };
  // ---- GP prior of interval i: columns 0..bs-1 belong to state i, bs..2bs-1 to state i+1
  if (i < nint && dt[i] > 0.0) {
#pragma unroll 1
    for (int rp = 0; rp < D; rp++) {
      double2 a[bs];
#pragma unroll
      for (int c = 0; c < bs; c++) a[c] = ld2(c, rp, i);
      if (doD) {
#pragma unroll
        for (int r = 0; r < bs; r++)
#pragma unroll
          for (int c = 0; c < 6; c++) acc[r][c] += a[r].x * a[c0 + c].x + a[r].y * a[c0 + c].y;
      }
      if (doE) {
        double2 b2[bs];
#pragma unroll
        for (int c = 0; c < bs; c++) b2[c] = ld2(bs + c, rp, i);
        if constexpr (TILES == 1) {
#pragma unroll
          for (int r = 0; r < bs; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) acc2[r][c] += b2[r].x * a[c].x + b2[r].y * a[c].y;
        } else {
#pragma unroll
          for (int r = 0; r < bs; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) acc[r][c] += b2[r].x * a[c0 + c].x + b2[r].y * a[c0 + c].y;
        }
      }
      if (doG) {
        const double2 rh = ld2(2 * bs, rp, i);
#pragma unroll
        for (int r = 0; r < bs; r++) g[r] += a[r].x * rh.x + a[r].y * rh.y;
      }
    }
  }
  // ---- GP prior of interval i-1: state i is its second state
  if (doD && i >= 1 && dt[i - 1] > 0.0) {
#pragma unroll 1
    for (int rp = 0; rp < D; rp++) {
      double2 a[bs];
#pragma unroll
      for (int c = 0; c < bs; c++) a[c] = ld2(bs + c, rp, i - 1);
#pragma unroll
      for (int r = 0; r < bs; r++)
#pragma unroll
        for (int c = 0; c < 6; c++) acc[r][c] += a[r].x * a[c0 + c].x + a[r].y * a[c0 + c].y;
      if (doG) {
        const double2 rh = ld2(2 * bs, rp, i - 1);
#pragma unroll
        for (int r = 0; r < bs; r++) g[r] += a[r].x * rh.x + a[r].y * rh.y;
      }
    }
  }
  // ---- extra rows of interval i (a-part = state i, b-part = state i+1) and of interval i-1 (b-part = state i)
  if (XR != nullptr) {
    if (i < nint) {
      for (int row = rowoff[i]; row < rowoff[i + 1]; row++) {
        double a[bs];
#pragma unroll
        for (int c = 0; c < bs; c++) a[c] = XR[(size_t)c * NXRp + row];
        if (doD) {
#pragma unroll
          for (int r = 0; r < bs; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) acc[r][c] += a[r] * a[c0 + c];
        }
        if (doE) {
          double b2[bs];
#pragma unroll
          for (int c = 0; c < bs; c++) b2[c] = XR[(size_t)(bs + c) * NXRp + row];
          if constexpr (TILES == 1) {
#pragma unroll
            for (int r = 0; r < bs; r++)
#pragma unroll
              for (int c = 0; c < 6; c++) acc2[r][c] += b2[r] * a[c];
          } else {
#pragma unroll
            for (int r = 0; r < bs; r++)
#pragma unroll
              for (int c = 0; c < 6; c++) acc[r][c] += b2[r] * a[c0 + c];
          }
        }
        if (doG) {
          const double rh = XR[(size_t)XRHS * NXRp + row];
#pragma unroll
          for (int r = 0; r < bs; r++) g[r] += a[r] * rh;
        }
      }
    }
    if (doD && i >= 1) {
      for (int row = rowoff[i - 1]; row < rowoff[i]; row++) {
        double b2[bs];
#pragma unroll
        for (int c = 0; c < bs; c++) b2[c] = XR[(size_t)(bs + c) * NXRp + row];
#pragma unroll
        for (int r = 0; r < bs; r++)
#pragma unroll
          for (int c = 0; c < 6; c++) acc[r][c] += b2[r] * b2[c0 + c];
        if (doG) {
          const double rh = XR[(size_t)XRHS * NXRp + row];
#pragma unroll
          for (int r = 0; r < bs; r++) g[r] += b2[r] * rh;
        }
      }
    }
  }
  // ---- store: tile = 6 full columns = 6*bs contiguous doubles of the column-major block
  double* rec = HREC + (size_t)i * REC;
  double* dst = rec + (doD ? 0 : bs * bs) + c0 * bs;
#pragma unroll
  for (int c = 0; c < 6; c++)
#pragma unroll
    for (int r = 0; r < bs; r += 2) st128(dst + c * bs + r, acc[r][c], acc[r + 1][c]);
  if constexpr (TILES == 1) {
    double* dstE = rec + bs * bs;
#pragma unroll
    for (int c = 0; c < 6; c++)
#pragma unroll
      for (int r = 0; r < bs; r += 2) st128(dstE + c * bs + r, acc2[r][c], acc2[r + 1][c]);
  }
  if (doG) {
#pragma unroll
    for (int r = 0; r < bs; r += 2) st128(rec + 2 * bs * bs + r, g[r], g[r + 1]);
  }
}
