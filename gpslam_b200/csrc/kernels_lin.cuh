// Shared device helpers + batched linearise kernels (see engine.cu for the data-layout overview).
#pragma once
#include <cuda_runtime.h>
#include <type_traits>
#include "factors.cuh"

using namespace gpb;


// ===================================================================== device helpers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_tile(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  const unsigned b = smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// FP64 tensor-core MMA, D(8x8) = A(8x4, row) * B(4x8, col) + C.  Fragments: a = A[lane>>2][lane&3], b = B[lane&3][lane>>2],
// c/d = C[lane>>2][2*(lane&3) + {0,1}]   (PTX ISA, mma.m8n8k4 .f64; SASS DMMA)
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void st128(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// block-wide sum, result valid in thread 0 (deterministic order)
template <int NT> __device__ __forceinline__ double block_sum(double v, double* sred) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sred[wid] = v;
  __syncthreads();
  double t = 0;
  if (threadIdx.x == 0) for (int k = 0; k < NT / 32; k++) t += sred[k];
  return t;
}

// ===================================================================== kernel: batched GP-prior linearise
// One thread per GP prior factor (interval i -> i+1).  The tile's NT+1 state records arrive in shared memory through one
// TMA bulk copy; every whitened column of [A|b] is produced in registers and stored as 128-bit row pairs into the SoA
// layout, so a warp's store instruction covers 512 contiguous bytes.
// DIAG (SE(3) only): every Qc model of the graph is diagonal - whitening is an element-wise scale (gp_prior_pose3_emit).
template <int G, int NT, bool DIAG = false, int MINB = 1>
__global__ void __launch_bounds__(NT, MINB) k_lin_gp(const double* __restrict__ X, const double* __restrict__ dt, const int* __restrict__ qc,
                                               const double* __restrict__ RqTab, double* __restrict__ AB, double* __restrict__ errpart,
                                               int nint, int NFp, int wantJ) {
  constexpr int D = GroupTraits<G>::D, SR = GroupTraits<G>::PS + D;
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ double sred[NT / 32];
  const int tile0 = blockIdx.x * NT;
  const int cnt = min(NT, nint - tile0);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) tma_load_tile(sm, X + (size_t)tile0 * SR, (unsigned)((cnt + 1) * SR * sizeof(double)), &bar);
  const int f = tile0 + threadIdx.x;
  // the factor's own parameters are fetched while the tile is in flight (nothing below the wait depends on a fresh global load)
  const double h = threadIdx.x < cnt ? dt[f] : 0.0;
  const int qi = threadIdx.x < cnt ? qc[f] : 0;
  mbar_wait(&bar, 0);
  double err = 0.0;
  if (threadIdx.x < cnt) {
    if (h > 0.0) {
      const double* s1 = sm + threadIdx.x * SR;
      const double* s2 = s1 + SR;
      const double* Rq = RqTab + qi * D * D;
      const GpWhiten w = gp_whiten(h);
      double col[2 * D];
      // this CTA's factors are one tile of the [A|b] layout (ab_off): entry (column c, row pair rp) = base + a compile-time offset
      static_assert(NT == AB_TF, "one linearise CTA per [A|b] tile");
      double* const abase = AB + ab_off(0, f, (4 * D + 1) * D);
      auto store_col = [&](int c, const double* v) {
#pragma unroll
        for (int rp = 0; rp < D; rp++) st128(abase + (size_t)(c * D + rp) * (AB_TF * 2), v[2 * rp], v[2 * rp + 1]);
      };
      auto store = [&](int c) { store_col(c, col); };
      if constexpr (G == G_POSE3) {
        err = gp_prior_pose3_emit<DIAG>(s1, s2, h, wantJ != 0, w, Rq, store_col);
      } else if constexpr (G == G_POSE3VW) {
        GpPose3VW o;
        gp_prior_pose3vw_eval(s1, s2, h, wantJ != 0, o);
        gp_prior_pose3vw_col<4, 0>(o, w, Rq, h, col);
#pragma unroll
        for (int k = 0; k < 12; k++) err += col[k] * col[k];
        if (wantJ) {
          store(24);
          static_for<0, 6>([&](auto c) { gp_prior_pose3vw_col<0, decltype(c)::value>(o, w, Rq, h, col); store(decltype(c)::value); });
          static_for<0, 6>([&](auto c) { gp_prior_pose3vw_col<1, decltype(c)::value>(o, w, Rq, h, col); store(6 + decltype(c)::value); });
          static_for<0, 6>([&](auto c) { gp_prior_pose3vw_col<2, decltype(c)::value>(o, w, Rq, h, col); store(12 + decltype(c)::value); });
          static_for<0, 6>([&](auto c) { gp_prior_pose3vw_col<3, decltype(c)::value>(o, w, Rq, h, col); store(18 + decltype(c)::value); });
        }
      } else {
        GpD3 o;
        gp_prior_d3_eval<G>(s1, s2, h, wantJ != 0, o);
        gp_prior_d3_col<4, 0>(o, w, Rq, h, col);
#pragma unroll
        for (int k = 0; k < 6; k++) err += col[k] * col[k];
        if (wantJ) {
          store(12);
          static_for<0, 3>([&](auto c) { gp_prior_d3_col<0, decltype(c)::value>(o, w, Rq, h, col); store(decltype(c)::value); });
          static_for<0, 3>([&](auto c) { gp_prior_d3_col<1, decltype(c)::value>(o, w, Rq, h, col); store(3 + decltype(c)::value); });
          static_for<0, 3>([&](auto c) { gp_prior_d3_col<2, decltype(c)::value>(o, w, Rq, h, col); store(6 + decltype(c)::value); });
          static_for<0, 3>([&](auto c) { gp_prior_d3_col<3, decltype(c)::value>(o, w, Rq, h, col); store(9 + decltype(c)::value); });
        }
      }
    }
  }
  const double tot = block_sum<NT>(0.5 * err, sred);
  if (threadIdx.x == 0) errpart[blockIdx.x] = tot;
}

// ===================================================================== kernel: measurement / prior / between rows
// One thread per "extra" factor; writes its m whitened rows over [state a (2D) | state b (2D) | landmark (DL) | rhs].
// CLS 0: interpolated measurement factors (range / attitude) — the volume; CLS 1: priors, between, plain 2-D factors;
// CLS 2: the multi-row SE(3) interpolated factors (GPS, projection), kept out of CLS 0 so the range path keeps its registers.
// Two kernels so the lean interpolated path does not inherit the generic path's local arrays and divergence.
template <int G, int CLS>
__device__ __forceinline__ void extra_rows(int kind, const double* __restrict__ X, const double* __restrict__ land, int sa, int sb, int l,
                                           const double* __restrict__ prm, bool wantJ, double* __restrict__ XR, int NXRp, int row0,
                                           double& err) {
  constexpr int D = GroupTraits<G>::D, PS = GroupTraits<G>::PS, SR = PS + D, DL = GroupTraits<G>::DL, bs = 2 * D;
  constexpr int NC = 2 * bs + DL + 1;
  const double* Rm = prm + 20;
  // helper: write one full row (coefficients c[NC-1] then rhs)
  auto put = [&](int row, int col, double v) { XR[(size_t)col * NXRp + row] = v; };
  if constexpr (CLS == 0) {
  if (kind == X_INTERP_RANGE) {
    const double isg = Rm[0];
    if constexpr (G == G_POSE3) {
      Range3Out o;
      interp_range_pose3(X + (size_t)sa * SR, X + (size_t)sb * SR, land + (size_t)l * 3, prm, wantJ, o);
      err += 0.5 * isg * isg * o.e * o.e;
      if (wantJ) {
#pragma unroll
        for (int k = 0; k < 6; k++) { put(row0, k, isg * elem(o.H1, k)); put(row0, 6 + k, isg * elem(o.H2, k)); put(row0, 12 + k, isg * elem(o.H3, k)); put(row0, 18 + k, isg * elem(o.H4, k)); }
        put(row0, 24, isg * o.H5.x); put(row0, 25, isg * o.H5.y); put(row0, 26, isg * o.H5.z);
        put(row0, 27, -isg * o.e);
      }
    } else if constexpr (G == G_POSE2 || G == G_LINEAR) {
      Range2Out o;
      interp_range_2d<G>(X + (size_t)sa * SR, X + (size_t)sb * SR, land + (size_t)l * 2, prm, wantJ, o);
      err += 0.5 * isg * isg * o.e * o.e;
      if (wantJ) {
#pragma unroll
        for (int k = 0; k < 3; k++) { put(row0, k, isg * elem(o.H1, k)); put(row0, 3 + k, isg * elem(o.H2, k)); put(row0, 6 + k, isg * elem(o.H3, k)); put(row0, 9 + k, isg * elem(o.H4, k)); }
        put(row0, 12, isg * o.H5[0]); put(row0, 13, isg * o.H5[1]);
        put(row0, 14, -isg * o.e);
      }
    }
    return;
  }
  if (kind == X_INTERP_ATTITUDE) {
    if constexpr (G == G_ROT3) {
      AttOut o;
      interp_attitude_rot3(X + (size_t)sa * SR, X + (size_t)sb * SR, prm, wantJ, o);
      // R is 2x2 upper triangular (isotropic in practice)
      const double r00 = Rm[0], r01 = Rm[2], r11 = Rm[3];
      const double w0 = r00 * o.e[0] + r01 * o.e[1], w1 = r11 * o.e[1];
      err += 0.5 * (w0 * w0 + w1 * w1);
      if (wantJ) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double a0[4] = {elem(o.H1[0], k), elem(o.H2[0], k), elem(o.H3[0], k), elem(o.H4[0], k)};
          const double a1[4] = {elem(o.H1[1], k), elem(o.H2[1], k), elem(o.H3[1], k), elem(o.H4[1], k)};
#pragma unroll
          for (int v = 0; v < 4; v++) { put(row0, 3 * v + k, r00 * a0[v] + r01 * a1[v]); put(row0 + 1, 3 * v + k, r11 * a1[v]); }
        }
        put(row0, 12, -w0); put(row0 + 1, 12, -w1);
      }
    }
    return;
  }
  }
  if constexpr (CLS == 2) {
  if (kind == X_INTERP_GPS || kind == X_INTERP_GPS_VW) {
    if constexpr (G == G_POSE3) {
      Gps3Out o;
      if (kind == X_INTERP_GPS) interp_gps_pose3(X + (size_t)sa * SR, X + (size_t)sb * SR, prm, wantJ, o);
      else interp_gps_pose3vw(X + (size_t)sa * SR, X + (size_t)sb * SR, prm, wantJ, o);
      // whiten with the dense upper-triangular 3x3 sqrt information: row r = sum_{k >= r} R[r,k] (.)
      const double ev[3] = {o.e.x, o.e.y, o.e.z};
#pragma unroll
      for (int r = 0; r < 3; r++) {
        double be = 0.0;
#pragma unroll
        for (int k = r; k < 3; k++) be += Rm[r + 3 * k] * ev[k];
        err += 0.5 * be * be;
        if (wantJ) {
#pragma unroll
          for (int v = 0; v < 4; v++)
#pragma unroll
            for (int c = 0; c < 6; c++) {
              double a = 0.0;
#pragma unroll
              for (int k = r; k < 3; k++) a += Rm[r + 3 * k] * elem(o.H[k][v], c);
              put(row0 + r, 6 * v + c, a);
            }
          put(row0 + r, 24, 0.0); put(row0 + r, 25, 0.0); put(row0 + r, 26, 0.0);
          put(row0 + r, 27, -be);
        }
      }
    }
    return;
  }
  if (kind == X_INTERP_PROJECTION) {
    if constexpr (G == G_POSE3) {
      Proj3Out o;
      interp_projection_pose3(X + (size_t)sa * SR, X + (size_t)sb * SR, land + (size_t)l * 3, prm, wantJ, o);
#pragma unroll
      for (int r = 0; r < 2; r++) {
        double be = 0.0;
#pragma unroll
        for (int k = r; k < 2; k++) be += Rm[r + 2 * k] * o.e[k];
        err += 0.5 * be * be;
        if (wantJ) {
#pragma unroll
          for (int v = 0; v < 4; v++)
#pragma unroll
            for (int c = 0; c < 6; c++) {
              double a = 0.0;
#pragma unroll
              for (int k = r; k < 2; k++) a += Rm[r + 2 * k] * elem(o.H[k][v], c);
              put(row0 + r, 6 * v + c, a);
            }
#pragma unroll
          for (int c = 0; c < 3; c++) {
            double a = 0.0;
#pragma unroll
            for (int k = r; k < 2; k++) a += Rm[r + 2 * k] * elem(o.H5[k], c);
            put(row0 + r, 24 + c, a);
          }
          put(row0 + r, 27, -be);
        }
      }
    }
    return;
  }
  }
}

// CLS 1 (priors, between incl. loop closures, plain 2-D factors): ONE WARP PER FACTOR.  Every lane evaluates the factor's
// residual e (m <= 6) and its small core Jacobian in registers (uniform across the warp - same factor, no divergence), lane c
// then forms column c of the unwhitened Jacobian over the row layout [state a | state b | landmark] (lane NC-1: the rhs),
// whitens it with the dense m x m sqrt information R and stores its m entries.  No local arrays, ~30x more parallelism than
// a thread per factor - these factors are few (priors every 100th state) but sit on the iteration's critical path.
template <int G>
__device__ __forceinline__ void extra_rows_warp(int kind, const double* __restrict__ X, const double* __restrict__ land, int sa, int sb, int l,
                                                const double* __restrict__ prm, bool wantJ, double* __restrict__ XR, int NXRp, int row0,
                                                int lane, double& err) {
  constexpr int D = GroupTraits<G>::D, PS = GroupTraits<G>::PS, SR = PS + D, DL = GroupTraits<G>::DL, bs = 2 * D;
  constexpr int NC = 2 * bs + DL + 1;
  static_assert(NC <= 32, "one lane per column of the row layout");
  const double* Rm = prm + 20;
  const int c = lane;
  double e[6], h[6];
#pragma unroll
  for (int k = 0; k < 6; k++) { e[k] = 0.0; h[k] = 0.0; }
  int m = 0;
  const int side = (sa >= 0) ? 0 : 1;          // single-state factors: a-part when sa valid, else b-part
  const int s = (sa >= 0) ? sa : sb;
  const int off = side * bs;
  if (kind == X_PRIOR_POSE) {
    m = D;
    const double* x = X + (size_t)s * SR;
    if constexpr (G == G_POSE3) { const X6 d = se3_logmap(p3_between(p3_from_wire(prm + 4), p3_from_wire(x)));
#pragma unroll
      for (int k = 0; k < 6; k++) e[k] = elem(d, k); }
    else if constexpr (G == G_ROT3) { const V3 d = so3_logmap(transpose(m3_from_wire(prm + 4)) * m3_from_wire(x)); e[0] = d.x; e[1] = d.y; e[2] = d.z; }
    else if constexpr (G == G_POSE2) { const P2 d = p2_between(p2(x[0], x[1], x[2]), p2(prm[4], prm[5], prm[6])); e[0] = -d.x; e[1] = -d.y; e[2] = -p2_theta(d); }
    else {
#pragma unroll
      for (int k = 0; k < 3; k++) e[k] = x[k] - prm[4 + k]; }
#pragma unroll
    for (int k = 0; k < D; k++) h[k] = (c == off + k) ? 1.0 : 0.0;
  } else if (kind == X_PRIOR_VEL) {
    m = D;
    const double* x = X + (size_t)s * SR + PS;
#pragma unroll
    for (int k = 0; k < D; k++) { e[k] = x[k] - prm[4 + k]; h[k] = (c == off + D + k) ? 1.0 : 0.0; }
  } else if (kind == X_PRIOR_LANDMARK) {
    if constexpr (DL > 0) {
      m = DL;
#pragma unroll
      for (int k = 0; k < DL; k++) { e[k] = land[(size_t)l * DL + k] - prm[4 + k]; h[k] = (c == 2 * bs + k) ? 1.0 : 0.0; }
    }
  } else if (kind == X_BETWEEN) {
    m = D;
    const double* x1 = X + (size_t)sa * SR;
    const double* x2 = X + (size_t)sb * SR;
    const bool swapped = prm[17] != 0.0;  // measured is (b -> a) when the factor was added as (i, i-1)
    const double* p = swapped ? x2 : x1;
    const double* q = swapped ? x1 : x2;
    const int o1 = swapped ? bs : 0, o2 = swapped ? 0 : bs;
    if constexpr (G == G_POSE3) {
      const P3 hx = p3_between(p3_from_wire(p), p3_from_wire(q));
      const X6 d = se3_logmap(p3_between(p3_from_wire(prm + 4), hx));
#pragma unroll
      for (int k = 0; k < 6; k++) e[k] = elem(d, k);
      const L6 A = l6_adjoint(p3_inverse(hx));
#pragma unroll
      for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int cc = 0; cc < 6; cc++) if (c == o1 + cc) h[r] = -elem(A, r, cc);
        if (c == o2 + r) h[r] = 1.0;
      }
    } else if constexpr (G == G_ROT3) {
      const M3 hx = transpose(m3_from_wire(p)) * m3_from_wire(q);
      const V3 d = so3_logmap(transpose(m3_from_wire(prm + 4)) * hx);
      e[0] = d.x; e[1] = d.y; e[2] = d.z;
#pragma unroll
      for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int cc = 0; cc < 3; cc++) if (c == o1 + cc) h[r] = -hx.m[3 * cc + r];
        if (c == o2 + r) h[r] = 1.0;
      }
    } else if constexpr (G == G_POSE2) {
      const P2 hx = p2_between(p2(p[0], p[1], p[2]), p2(q[0], q[1], q[2]));
      const P2 d = p2_between(p2(prm[4], prm[5], prm[6]), hx);
      e[0] = d.x; e[1] = d.y; e[2] = p2_theta(d);
      const M3 A = p2_adjoint(p2_inverse(hx));
#pragma unroll
      for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int cc = 0; cc < 3; cc++) if (c == o1 + cc) h[r] = -A.m[3 * r + cc];
        if (c == o2 + r) h[r] = 1.0;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; k++) { e[k] = (q[k] - p[k]) - prm[4 + k]; h[k] = (c == o1 + k) ? -1.0 : ((c == o2 + k) ? 1.0 : 0.0); }
    }
  } else if (kind == X_RANGE_2D) {
    if constexpr (G == G_POSE2 || G == G_LINEAR) {
      m = 1;
      const double* x = X + (size_t)s * SR;
      const double dx = land[(size_t)l * 2] - x[0], dy = land[(size_t)l * 2 + 1] - x[1];
      const double r = sqrt(dx * dx + dy * dy);
      e[0] = r - prm[2];
      double hx, hy, h0, h1;
      if (G == G_LINEAR && !(fabs(r) > 1e-10)) { hx = 1; hy = 1; } else { hx = dx / r; hy = dy / r; }
      if constexpr (G == G_POSE2) { const double cs = cos(x[2]), sn = sin(x[2]); h0 = -(hx * cs + hy * sn); h1 = hx * sn - hy * cs; }
      else { h0 = -hx; h1 = -hy; }
      h[0] = (c == off) ? h0 : (c == off + 1) ? h1 : (c == 2 * bs) ? hx : (c == 2 * bs + 1) ? hy : 0.0;
    }
  } else if (kind == X_RANGE_BEARING_2D) {
    if constexpr (G == G_LINEAR) {
      m = 2;
      const double* x = X + (size_t)s * SR;
      const double cs = cos(x[2]), sn = sin(x[2]);
      const double dx = land[(size_t)l * 2] - x[0], dy = land[(size_t)l * 2 + 1] - x[1];
      const double rx = cs * dx + sn * dy, ry = -sn * dx + cs * dy;
      const double n = sqrt(rx * rx + ry * ry);
      const double ec = rx / n, es = ry / n, bc = cos(prm[3]), bsn = sin(prm[3]);
      const double d = sqrt(dx * dx + dy * dy);
      e[0] = atan2(bc * es - bsn * ec, bc * ec + bsn * es);
      e[1] = d - prm[2];
      double hx, hy;
      if (fabs(d) > 1e-10) { hx = dx / d; hy = dy / d; } else { hx = 1; hy = 1; }
      double t0 = 0, t1 = 0;
      if (d > 1e-5) { t0 = -ry / (d * d); t1 = rx / (d * d); }
      // H11 = tmp * [ -R^T , (ry, -rx)^T ],  H12 = tmp * R^T ; R^T = [[c, s],[-s, c]]
      const double g0 = t0 * cs - t1 * sn, g1 = t0 * sn + t1 * cs;
      h[0] = (c == off) ? -g0 : (c == off + 1) ? -g1 : (c == off + 2) ? (t0 * ry - t1 * rx) : (c == 2 * bs) ? g0 : (c == 2 * bs + 1) ? g1 : 0.0;
      h[1] = (c == off) ? -hx : (c == off + 1) ? -hy : (c == 2 * bs) ? hx : (c == 2 * bs + 1) ? hy : 0.0;
    }
  } else if (kind == X_ODOMETRY_2D) {
    if constexpr (G == G_LINEAR) {
      m = 3;
      const double* x1 = X + (size_t)sa * SR;
      const double* x2 = X + (size_t)sb * SR;
      const double cs = cos(x1[2]), sn = sin(x1[2]);
      const double vx = x2[0] - x1[0], vy = x2[1] - x1[1];
      const double qx = cs * vx + sn * vy, qy = -sn * vx + cs * vy;
      e[0] = qx - prm[4]; e[1] = qy - prm[5]; e[2] = (x2[2] - x1[2]) - prm[6];
      h[0] = (c == 0) ? -cs : (c == 1) ? -sn : (c == 2) ? qy : (c == bs) ? cs : (c == bs + 1) ? sn : 0.0;
      h[1] = (c == 0) ? sn : (c == 1) ? -cs : (c == 2) ? -qx : (c == bs) ? -sn : (c == bs + 1) ? cs : 0.0;
      h[2] = (c == 2) ? -1.0 : (c == bs + 2) ? 1.0 : 0.0;
    }
  }
  // whiten: row r = sum_{k >= r} R[r,k] (.)   (R upper triangular, m x m column-major)
#pragma unroll
  for (int r = 0; r < 6; r++) {
    if (r < m) {
      double be = 0.0, a = 0.0;
#pragma unroll
      for (int k = r; k < 6; k++) if (k < m) { const double rk = Rm[r + k * m]; be += rk * e[k]; a += rk * h[k]; }
      if (lane == 0) err += 0.5 * be * be;
      if (wantJ) {
        if (c < NC - 1) XR[(size_t)c * NXRp + row0 + r] = a;
        else if (c == NC - 1) XR[(size_t)c * NXRp + row0 + r] = -be;
      }
    }
  }
}

template <int G, int CLS, int NT>
__global__ void __launch_bounds__(NT) k_lin_extra(const int* __restrict__ list, int nlist, const double* __restrict__ X, const double* __restrict__ land, const int* __restrict__ xkind,
                                                  const int* __restrict__ xsa, const int* __restrict__ xsb, const int* __restrict__ xl,
                                                  const int* __restrict__ xrow, const double* __restrict__ xprm, double* __restrict__ XR,
                                                  double* __restrict__ errpart, int nx, int NXRp, int wantJ) {
  __shared__ double sred[NT / 32];
  const int t = blockIdx.x * NT + threadIdx.x;
  double err = 0.0;
  if constexpr (CLS != 1) {
    if (t < nlist) {
      const int f = list[t];
      extra_rows<G, CLS>(xkind[f], X, land, xsa[f], xsb[f], xl[f], xprm + (size_t)f * XP_STRIDE, wantJ != 0, XR, NXRp, xrow[f], err);
    }
  } else {
    if ((t >> 5) < nlist) {  // one warp per factor
      const int f = list[t >> 5];
      extra_rows_warp<G>(xkind[f], X, land, xsa[f], xsb[f], xl[f], xprm + (size_t)f * XP_STRIDE, wantJ != 0, XR, NXRp, xrow[f], threadIdx.x & 31, err);
    }
  }
  (void)nx;
  const double tot = block_sum<NT>(err, sred);
  if (threadIdx.x == 0) errpart[blockIdx.x] = tot;
}

// interpolatePose queries (gpb_interpolate_poses / gpb_graph_interpolate): one thread per query, support records ia[k], ib[k] of X
template <int G>
__global__ void __launch_bounds__(128) k_interp_query(const double* __restrict__ X, const int* __restrict__ ia, const int* __restrict__ ib, const double* __restrict__ dt,
                                                      const double* __restrict__ tau, int n, double* __restrict__ poses, double* __restrict__ H) {
  constexpr int D = GroupTraits<G>::D, PS = GroupTraits<G>::PS, SR = PS + D;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double pose[PS], Hl[4 * D * D];
  interp_pose<G>(X + (size_t)ia[k] * SR, X + (size_t)ib[k] * SR, dt[k], tau[k], H != nullptr, pose, Hl);
#pragma unroll
  for (int t = 0; t < PS; t++) poses[(size_t)k * PS + t] = pose[t];
  if (H != nullptr) {
#pragma unroll
    for (int t = 0; t < 4 * D * D; t++) H[(size_t)k * 4 * D * D + t] = Hl[t];
  }
}

// GaussianProcessInterpolatorLinear::interpolateVelocity (gp/GaussianProcessInterpolatorLinear.h:106-126): the lower D rows of
// Lambda x1 + Psi x2.  Like the upper rows (interp_coef) every D x D block is a scalar times I and does not depend on Qc: with
// s = tau / dt,  Psi21 = 6 (s - s^2) / dt,  Psi22 = 3 s^2 - 2 s,  Lambda21 = -Psi21,  Lambda22 = 1 - 4 s + 3 s^2
// (Q(tau) Phi(dt - tau)^T Q(dt)^-1 and Phi(tau) - Psi Phi(dt) of gp/GPutils.h:54-71 multiplied out).
// One thread per query; H (or null): the four scalars (H1 = Lambda21 I, H2 = Lambda22 I, H3 = Psi21 I, H4 = Psi22 I).
__global__ void __launch_bounds__(128) k_interp_velocity_linear(const double* __restrict__ x1, const double* __restrict__ v1, const double* __restrict__ x2, const double* __restrict__ v2,
                                                                const double* __restrict__ dt, const double* __restrict__ tau, int n, int D, double* __restrict__ vel, double* __restrict__ H) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double s = tau[k] / dt[k];
  const double psi21 = 6.0 * (s - s * s) / dt[k], psi22 = 3.0 * s * s - 2.0 * s, lam21 = -psi21, lam22 = 1.0 - 4.0 * s + 3.0 * s * s;
  for (int d = 0; d < D; d++) {
    const size_t o = (size_t)k * D + d;
    vel[o] = lam21 * x1[o] + lam22 * v1[o] + psi21 * x2[o] + psi22 * v2[o];
  }
  if (H != nullptr) { H[4 * k] = lam21; H[4 * k + 1] = lam22; H[4 * k + 2] = psi21; H[4 * k + 3] = psi22; }
}

// deterministic final sum of block partials (single block)
__global__ void k_sum_partials(const double* __restrict__ part, int n, double* __restrict__ out, int slot) {
  __shared__ double sred[8];
  double v = 0;
  for (int k = threadIdx.x; k < n; k += 256) v += part[k];
  const double t = block_sum<256>(v, sred);
  if (threadIdx.x == 0) out[slot] = t;
}

