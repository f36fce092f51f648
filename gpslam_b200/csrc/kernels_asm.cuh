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
  auto ld2 = [&](int col, int rp, int f) { return *reinterpret_cast<const double2*>(AB + ab_off(col * D + rp, f, (4 * D + 1) * D)); };
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

// ---------------------------------------------------------------------------------------------------------------------
// SE(3) assembly on the FP64 tensor pipe.  One CTA (4 warps) builds the records of TS = 8 consecutive states.
//   load    : the whitened [A|b] of the 9 GP priors touching those states is staged from the SoA linearisation buffer into
//             shared memory, factor-major, with cp.async (LDGSTS: ~11 independent 16-byte copies per thread in flight, no
//             register staging).  Intervals without a GP prior hold zeros in the buffer (cleared once at finalize).
//   compute : per state (two per warp)  D_i | g_i = Fa_a^T [Fa_a | b] + Fp_b^T [Fp_b | b]  and  E_i = Fa_b^T Fa_a  as
//             mma.sync.m8n8k4.f64 tiles (36 DMMA), Fa / Fp = priors of the intervals i / i-1, _a / _b their columns on the first /
//             second state; the rhs rides along as column 12 of the second column tile.  The few measurement / prior rows of
//             the two intervals are further k-steps of the same products, their fragments read straight from the row table.
//   store   : each state's record straight from its accumulator fragments (whole 32-byte sectors per store instruction)
// This replaces the thread-per-tile kernel for SE(3) (255 registers, 8 warps per SM: latency-bound at a quarter of the HBM rate).
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(valid ? 16 : 0) : "memory");
}
struct AsmGeom { static constexpr int D = 6, bs = 12, REC = 2 * bs * bs + bs, TS = 8, NF = TS + 1, NCOL = 4 * D + 1, FS = NCOL * bs + 2; };  // FS: doubles per staged factor (padded)
struct AsmRows {   // per warp: row ranges of its two states and the prefetched first k-step of their extra rows
  int rr0[2], rr1[2], rr2[2];
  double xc0[2], xc1[2], xd0[2], xd1[2], xq0[2], xq1[2];
};
// issue the cp.async copies of the 9 GP priors of the tile starting at state i0 into one staging buffer and commit them as one group
__device__ __forceinline__ void asm_issue_load(double* Fsm, const double* __restrict__ AB, int i0, int NFp, int tid) {
  constexpr int D = AsmGeom::D, NF = AsmGeom::NF, NCOL = AsmGeom::NCOL, FS = AsmGeom::FS;
  // ---- load: item = (row pair pr = column * D + rp, factor ff), 16 bytes each; consecutive threads take consecutive factors.
  // Thread (ff, g) = (tid % NF, tid / NF) of the first 14 * NF = 126 threads copies row pairs g, g + 14, ...: its source
  // (ab_off: + one row of the tile per row pair) and destination (Fsm[ff][2 pr]) advance by constants - no index arithmetic
  // per copy (the straightforward item -> (ff, c, rp) decode was a quarter of this kernel's instructions)
  {
    constexpr int NPR = NCOL * D, NG = 128 / NF;
    const int ff = tid % NF, g = tid / NF;
    if (g < NG) {
      const int f = i0 - 1 + ff;
      const bool ok = f >= 0 && f < NFp;
      const double* src = AB + ab_off(g, ok ? f : 0, NPR);
      double* dst = &Fsm[ff * FS + 2 * g];
#pragma unroll
      for (int j = 0; j < (NPR + NG - 1) / NG; j++)
        if (g + NG * j < NPR) cp_async16_zfill(dst + 2 * NG * j, src + (size_t)(AB_TF * 2) * NG * j, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
}
__device__ __forceinline__ void asm_prefetch_rows(AsmRows& R, const double* __restrict__ XR, const int* __restrict__ rowoff, int i0, int N, int NXRp, int xrhs,
                                                  int warp, int gi, int ti) {
  const int nint = N - 1;
  int (&rr0)[2] = R.rr0, (&rr1)[2] = R.rr1, (&rr2)[2] = R.rr2;
  double (&xc0)[2] = R.xc0, (&xc1)[2] = R.xc1, (&xd0)[2] = R.xd0, (&xd1)[2] = R.xd1, (&xq0)[2] = R.xq0, (&xq1)[2] = R.xq1;
  // row ranges of this warp's two states (interval i-1: [r0, r1), interval i: [r1, r2)) while the copies fly
#pragma unroll
  for (int sidx = 0; sidx < 2; sidx++) {
    const int i = i0 + 2 * warp + sidx;
    rr0[sidx] = rr1[sidx] = rr2[sidx] = 0;
    if (XR != nullptr && i < N) {
      rr1[sidx] = rowoff[i < nint ? i : nint];
      rr0[sidx] = i >= 1 ? rowoff[i - 1] : rr1[sidx];
      rr2[sidx] = i < nint ? rowoff[i + 1] : rr1[sidx];
    }
  }
  // first k-step of the extra rows of both states, fetched while the staged factors are still in flight: the measurement rows
  // sit in a [column][row] table, each fragment is a scattered 8-byte load, and almost every interval has at most four rows -
  // taken after the barrier these loads were the kernel's longest stall (long scoreboard 26 % of the samples).
  // xc1 / xq1: column 8 + gi (resp. 20 + gi) for gi < 4, the rhs column for gi == 4
#pragma unroll
  for (int sidx = 0; sidx < 2; sidx++) {
    {
      const int base = rr1[sidx], row = base + ti;
      const bool v = row < rr2[sidx];
      const double* x = XR + (v ? row : base);
      xc0[sidx] = v ? x[(size_t)gi * NXRp] : 0.0;
      xc1[sidx] = (v && gi <= 4) ? x[(size_t)(gi < 4 ? 8 + gi : xrhs) * NXRp] : 0.0;
      xd0[sidx] = v ? x[(size_t)(12 + gi) * NXRp] : 0.0;
      xd1[sidx] = (v && gi < 4) ? x[(size_t)(20 + gi) * NXRp] : 0.0;
    }
    {
      const int base = rr0[sidx], row = base + ti;
      const bool v = row < rr1[sidx];
      const double* x = XR + (v ? row : base);
      xq0[sidx] = v ? x[(size_t)(12 + gi) * NXRp] : 0.0;
      xq1[sidx] = (v && gi <= 4) ? x[(size_t)(gi < 4 ? 20 + gi : xrhs) * NXRp] : 0.0;
    }
  }
}
__device__ __forceinline__ void asm_compute_store(const double* Fsm, const AsmRows& R, const double* __restrict__ XR, double* __restrict__ HREC, int i0, int N,
                                                  int NXRp, int xrhs, int warp, int gi, int ti) {
  constexpr int bs = AsmGeom::bs, REC = AsmGeom::REC, FS = AsmGeom::FS;
  const int (&rr0)[2] = R.rr0, (&rr1)[2] = R.rr1, (&rr2)[2] = R.rr2;
  const double (&xc0)[2] = R.xc0, (&xc1)[2] = R.xc1, (&xd0)[2] = R.xd0, (&xd1)[2] = R.xd1, (&xq0)[2] = R.xq0, (&xq1)[2] = R.xq1;
  double accD[2][2][2][2], accE[2][2][2][2];  // [state][mt][nt][2]
#pragma unroll
  for (int sidx = 0; sidx < 2; sidx++) {
    const int sl = 2 * warp + sidx;             // state slot in the tile
    const double* Fp = Fsm + sl * FS;           // prior of interval i-1
    const double* Fa = Fsm + (sl + 1) * FS;     // prior of interval i
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int nt = 0; nt < 2; nt++) { accD[sidx][mt][nt][0] = accD[sidx][mt][nt][1] = 0.0; accE[sidx][mt][nt][0] = accE[sidx][mt][nt][1] = 0.0; }
    // one k-step (4 rows) of the products; a*: columns of the first state, b*: of the second state, *1b: second column tile
    // carrying the rhs as column 12.  cur: rows of interval i (all three products); else rows of interval i-1 (their second-
    // state columns feed D and g only)
    auto kstep_cur = [&](double a0, double a1, double a1b, double b0, double b1) {
      dmma884(accD[sidx][0][0][0], accD[sidx][0][0][1], a0, a0);
      dmma884(accD[sidx][0][1][0], accD[sidx][0][1][1], a0, a1b);
      dmma884(accD[sidx][1][0][0], accD[sidx][1][0][1], a1, a0);
      dmma884(accD[sidx][1][1][0], accD[sidx][1][1][1], a1, a1b);
      dmma884(accE[sidx][0][0][0], accE[sidx][0][0][1], b0, a0);
      dmma884(accE[sidx][0][1][0], accE[sidx][0][1][1], b0, a1);
      dmma884(accE[sidx][1][0][0], accE[sidx][1][0][1], b1, a0);
      dmma884(accE[sidx][1][1][0], accE[sidx][1][1][1], b1, a1);
    };
    auto kstep_prev = [&](double p0, double p1, double p1b) {
      dmma884(accD[sidx][0][0][0], accD[sidx][0][0][1], p0, p0);
      dmma884(accD[sidx][0][1][0], accD[sidx][0][1][1], p0, p1b);
      dmma884(accD[sidx][1][0][0], accD[sidx][1][0][1], p1, p0);
      dmma884(accD[sidx][1][1][0], accD[sidx][1][1][1], p1, p1b);
    };
#pragma unroll
    for (int ks = 0; ks < 3; ks++) {
      const int k = 4 * ks + ti;
      const double a0 = Fa[gi * bs + k], a1 = (gi < 4) ? Fa[(8 + gi) * bs + k] : 0.0, a1b = (gi < 4) ? a1 : (gi == 4 ? Fa[24 * bs + k] : 0.0);
      const double b0 = Fa[(12 + gi) * bs + k], b1 = (gi < 4) ? Fa[(20 + gi) * bs + k] : 0.0;
      kstep_cur(a0, a1, a1b, b0, b1);
      const double p0 = Fp[(12 + gi) * bs + k], p1 = (gi < 4) ? Fp[(20 + gi) * bs + k] : 0.0, p1b = (gi < 4) ? p1 : (gi == 4 ? Fp[24 * bs + k] : 0.0);
      kstep_prev(p0, p1, p1b);
    }
    // extra rows, four per k-step, fragments straight from the row table XR[column][row]; the first k-step was prefetched
    if (rr1[sidx] < rr2[sidx]) kstep_cur(xc0[sidx], gi < 4 ? xc1[sidx] : 0.0, xc1[sidx], xd0[sidx], xd1[sidx]);
    for (int base = rr1[sidx] + 4; base < rr2[sidx]; base += 4) {
      const int row = base + ti;
      const bool v = row < rr2[sidx];
      const double* x = XR + (v ? row : base);
      const double a0 = v ? x[(size_t)gi * NXRp] : 0.0, a1 = (v && gi < 4) ? x[(size_t)(8 + gi) * NXRp] : 0.0;
      const double a1b = (gi < 4) ? a1 : ((v && gi == 4) ? x[(size_t)xrhs * NXRp] : 0.0);
      const double b0 = v ? x[(size_t)(12 + gi) * NXRp] : 0.0, b1 = (v && gi < 4) ? x[(size_t)(20 + gi) * NXRp] : 0.0;
      kstep_cur(a0, a1, a1b, b0, b1);
    }
    if (rr0[sidx] < rr1[sidx]) kstep_prev(xq0[sidx], gi < 4 ? xq1[sidx] : 0.0, xq1[sidx]);
    for (int base = rr0[sidx] + 4; base < rr1[sidx]; base += 4) {
      const int row = base + ti;
      const bool v = row < rr1[sidx];
      const double* x = XR + (v ? row : base);
      const double p0 = v ? x[(size_t)(12 + gi) * NXRp] : 0.0, p1 = (v && gi < 4) ? x[(size_t)(20 + gi) * NXRp] : 0.0;
      const double p1b = (gi < 4) ? p1 : ((v && gi == 4) ? x[(size_t)xrhs * NXRp] : 0.0);
      kstep_prev(p0, p1, p1b);
    }
    // ---- store: straight from the accumulator fragments.  For one (mt, nt, h) the eight lanes of equal ti hold eight
    // consecutive rows of one column - a 64-byte run (32 bytes for rows 8..11) that starts on a 32-byte boundary (records are
    // 2400 bytes, columns 96) - so every store instruction writes whole sectors and nothing needs to pass through shared memory
    if (i0 + sl < N) {
      double* O = HREC + (size_t)(i0 + sl) * REC;
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int row = 8 * mt + gi, col = 8 * nt + 2 * ti + h;
            if (row < bs) {
              if (col < bs) { O[row + col * bs] = accD[sidx][mt][nt][h]; O[bs * bs + row + col * bs] = accE[sidx][mt][nt][h]; }
              else if (col == bs) O[2 * bs * bs + row] = accD[sidx][mt][nt][h];
            }
          }
    }
  }
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_assemble_mma(const double* __restrict__ AB, const double* __restrict__ XR,
                                                      const int* __restrict__ rowoff, double* __restrict__ HREC, int N, int NFp, int NXRp, int xrhs, int blk0) {
  __shared__ __align__(16) double Fsm[AsmGeom::NF * AsmGeom::FS];   // staged factors [factor][column][row]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gi = lane >> 2, ti = lane & 3;
  const int i0 = (blockIdx.x + blk0) * AsmGeom::TS;   // blk0: first tile of this launch (the chunked linearise / assemble pipeline launches ranges of tiles)
  asm_issue_load(Fsm, AB, i0, NFp, tid);
  AsmRows R;
  asm_prefetch_rows(R, XR, rowoff, i0, N, NXRp, xrhs, warp, gi, ti);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  asm_compute_store(Fsm, R, XR, HREC, i0, N, NXRp, xrhs, warp, gi, ti);
}
// The same tile pipeline as a persistent kernel: a resident wave of CTAs walks the tiles with stride gridDim.x, the copies of
// the next tile in flight (second staging buffer) while the current one is multiplied and stored - the one-tile CTAs above spend
// most of their ~10 us life between launch, the first copy landing and the barrier, with nothing on the SM saturated.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_assemble_mma_p(const double* __restrict__ AB, const double* __restrict__ XR,
                                                        const int* __restrict__ rowoff, double* __restrict__ HREC, int N, int NFp, int NXRp, int xrhs, int ntiles) {
  __shared__ __align__(16) double Fsm[2][AsmGeom::NF * AsmGeom::FS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gi = lane >> 2, ti = lane & 3;
  int tile = blockIdx.x, st = 0;
  if (tile < ntiles) asm_issue_load(Fsm[0], AB, tile * AsmGeom::TS, NFp, tid);
  for (; tile < ntiles; tile += gridDim.x, st ^= 1) {
    const int i0 = tile * AsmGeom::TS, next = tile + gridDim.x;
    if (next < ntiles) asm_issue_load(Fsm[st ^ 1], AB, next * AsmGeom::TS, NFp, tid);   // that buffer was released by the barrier closing the previous round
    else asm volatile("cp.async.commit_group;" ::: "memory");                           // keep one group per round
    AsmRows R;
    asm_prefetch_rows(R, XR, rowoff, i0, N, NXRp, xrhs, warp, gi, ti);
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // everything but the newest group: this tile has landed
    __syncthreads();
    asm_compute_store(Fsm[st], R, XR, HREC, i0, N, NXRp, xrhs, warp, gi, ti);
    __syncthreads();                                       // every warp is done with this buffer before the next round refills it
  }
}
