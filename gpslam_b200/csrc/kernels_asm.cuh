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
  // a warp is one tile class over 32 consecutive states (coalesced 128-bit loads, no divergence); the TILES warps that
  // work on the same 32 states sit in the same CTA so the [A|b] columns they share are fetched from HBM once (L1 hits)
  static_assert(NT % (32 * TILES) == 0, "CTA = groups of TILES warps");
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = wid % TILES;
  const int i = (blockIdx.x * (NT / (32 * TILES)) + wid / TILES) * 32 + lane;
  if (i >= N) return;
  const int nint = N - 1;
  const int c0 = (TILES == 1) ? 0 : (tile & 1) * 6;
  auto ld2 = [&](int col, int rp, int f) { return *reinterpret_cast<const double2*>(AB + ((size_t)(col * D + rp) * NFp + f) * 2); };
  double* rec = HREC + (size_t)i * REC;

  // One tile = 6 columns of D_i (DOD) and/or of E_i (DOE); C0, DOD, DOE, DOG are compile-time so every operand array stays in
  // registers and each tile class only carries the operands it needs.
  auto run = [&](auto c0tag, auto dtag, auto etag, auto gtag) {
    constexpr int C0 = decltype(c0tag)::value;
    constexpr bool DOD = decltype(dtag)::value, DOE = decltype(etag)::value, DOG = decltype(gtag)::value;
    double accD[DOD ? bs : 1][6], accE[DOE ? bs : 1][6], g[DOG ? bs : 1];
#pragma unroll
    for (int r = 0; r < (DOD ? bs : 1); r++)
#pragma unroll
      for (int c = 0; c < 6; c++) accD[r][c] = 0;
#pragma unroll
    for (int r = 0; r < (DOE ? bs : 1); r++)
#pragma unroll
      for (int c = 0; c < 6; c++) accE[r][c] = 0;
#pragma unroll
    for (int r = 0; r < (DOG ? bs : 1); r++) g[r] = 0;

    // ---- GP prior of interval i: columns 0..bs-1 belong to state i, bs..2bs-1 to state i+1
    if (i < nint && dt[i] > 0.0) {
#pragma unroll 1
      for (int rp = 0; rp < D; rp++) {
        if constexpr (DOD) {
          double2 a[bs];
#pragma unroll
          for (int c = 0; c < bs; c++) a[c] = ld2(c, rp, i);
#pragma unroll
          for (int r = 0; r < bs; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) accD[r][c] += a[r].x * a[C0 + c].x + a[r].y * a[C0 + c].y;
          if constexpr (DOG) {
            const double2 rh = ld2(2 * bs, rp, i);
#pragma unroll
            for (int r = 0; r < bs; r++) g[r] += a[r].x * rh.x + a[r].y * rh.y;
          }
        }
        if constexpr (DOE) {
          double2 a6[6], b2[bs];
#pragma unroll
          for (int c = 0; c < 6; c++) a6[c] = ld2(C0 + c, rp, i);
#pragma unroll
          for (int c = 0; c < bs; c++) b2[c] = ld2(bs + c, rp, i);
#pragma unroll
          for (int r = 0; r < bs; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) accE[r][c] += b2[r].x * a6[c].x + b2[r].y * a6[c].y;
        }
      }
    }
    // ---- GP prior of interval i-1: state i is its second state
    if constexpr (DOD) {
      if (i >= 1 && dt[i - 1] > 0.0) {
#pragma unroll 1
        for (int rp = 0; rp < D; rp++) {
          double2 a[bs];
#pragma unroll
          for (int c = 0; c < bs; c++) a[c] = ld2(bs + c, rp, i - 1);
#pragma unroll
          for (int r = 0; r < bs; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) accD[r][c] += a[r].x * a[C0 + c].x + a[r].y * a[C0 + c].y;
          if constexpr (DOG) {
            const double2 rh = ld2(2 * bs, rp, i - 1);
#pragma unroll
            for (int r = 0; r < bs; r++) g[r] += a[r].x * rh.x + a[r].y * rh.y;
          }
        }
      }
    }
    // ---- extra rows of interval i (a-part = state i, b-part = state i+1) and of interval i-1 (b-part = state i)
    if (XR != nullptr) {
      if (i < nint) {
        for (int row = rowoff[i]; row < rowoff[i + 1]; row++) {
          if constexpr (DOD) {
            double a[bs];
#pragma unroll
            for (int c = 0; c < bs; c++) a[c] = XR[(size_t)c * NXRp + row];
#pragma unroll
            for (int r = 0; r < bs; r++)
#pragma unroll
              for (int c = 0; c < 6; c++) accD[r][c] += a[r] * a[C0 + c];
            if constexpr (DOG) {
              const double rh = XR[(size_t)XRHS * NXRp + row];
#pragma unroll
              for (int r = 0; r < bs; r++) g[r] += a[r] * rh;
            }
          }
          if constexpr (DOE) {
            double a6[6], b2[bs];
#pragma unroll
            for (int c = 0; c < 6; c++) a6[c] = XR[(size_t)(C0 + c) * NXRp + row];
#pragma unroll
            for (int c = 0; c < bs; c++) b2[c] = XR[(size_t)(bs + c) * NXRp + row];
#pragma unroll
            for (int r = 0; r < bs; r++)
#pragma unroll
              for (int c = 0; c < 6; c++) accE[r][c] += b2[r] * a6[c];
          }
        }
      }
      if constexpr (DOD) {
        if (i >= 1) {
          for (int row = rowoff[i - 1]; row < rowoff[i]; row++) {
            double b2[bs];
#pragma unroll
            for (int c = 0; c < bs; c++) b2[c] = XR[(size_t)(bs + c) * NXRp + row];
#pragma unroll
            for (int r = 0; r < bs; r++)
#pragma unroll
              for (int c = 0; c < 6; c++) accD[r][c] += b2[r] * b2[C0 + c];
            if constexpr (DOG) {
              const double rh = XR[(size_t)XRHS * NXRp + row];
#pragma unroll
              for (int r = 0; r < bs; r++) g[r] += b2[r] * rh;
            }
          }
        }
      }
    }
    // ---- store: a tile = 6 full columns = 6*bs contiguous doubles of the column-major block
    if constexpr (DOD) {
      double* dst = rec + C0 * bs;
#pragma unroll
      for (int c = 0; c < 6; c++)
#pragma unroll
        for (int r = 0; r < bs; r += 2) st128(dst + c * bs + r, accD[r][c], accD[r + 1][c]);
    }
    if constexpr (DOE) {
      double* dst = rec + bs * bs + C0 * bs;
#pragma unroll
      for (int c = 0; c < 6; c++)
#pragma unroll
        for (int r = 0; r < bs; r += 2) st128(dst + c * bs + r, accE[r][c], accE[r + 1][c]);
    }
    if constexpr (DOG) {
#pragma unroll
      for (int r = 0; r < bs; r += 2) st128(rec + 2 * bs * bs + r, g[r], g[r + 1]);
    }
  };
  using T = std::true_type; using F = std::false_type;
  if constexpr (TILES == 1) {
    run(std::integral_constant<int, 0>{}, T{}, T{}, T{});
  } else {
    if (tile == 0) run(std::integral_constant<int, 0>{}, T{}, F{}, T{});
    else if (tile == 1) run(std::integral_constant<int, 6>{}, T{}, F{}, F{});
    else if (tile == 2) run(std::integral_constant<int, 0>{}, F{}, T{}, F{});
    else run(std::integral_constant<int, 6>{}, F{}, T{}, F{});
  }
  (void)c0;
}
