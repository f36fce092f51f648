// B200-native (sm_100a) sparse-GP factor-graph engine: batched linearise, normal-equation assembly,
// multi-level segment-parallel bordered block-tridiagonal Cholesky, GN/LM loop.  C ABI in include/gpb.h.
//
// Device data layout (all FP64, resident in HBM for the life of the graph):
//   X      [N][SR]            state records  [pose (PS) | velocity (D)]            (AoS, 16 B-aligned records,
//                              staged per tile into shared memory by one TMA bulk copy, cp.async.bulk)
//   AB     [NFp/128][(4D+1) D][128][2]   whitened GP-prior JacobianFactors [A|b]: tiles of 128 factors, inside a tile SoA by
//          (column, row pair) - factors.cuh:ab_off  (128-bit stores,
//                              consecutive lanes -> consecutive 16 B)
//   XR     [2bs+DL+1][NXRp]    whitened rows of every other factor (range, attitude, priors, between, 2-D factors)
//   HREC   [N][2bs^2+bs]       assembled normal equations per state: D_i | E_i (= H_{i+1,i}) | g_i
//   level records / factor records of the solver: see "solver" below.
//
// Reference path replaced: NonlinearFactorGraph::linearize -> GaussianFactorGraph ->
// eliminateMultifrontal (Cholesky) -> back-substitution -> Values::retract -> graph.error, driven by
// GaussNewtonOptimizer / LevenbergMarquardtOptimizer::iterate.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/gpb.h"
#include "kernels_lin.cuh"
#include "kernels_asm.cuh"
#include "kernels_solve.cuh"

// ===================================================================== errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(x)                                                                                           \
  do {                                                                                                        \
    cudaError_t e_ = (x);                                                                                     \
    if (e_ != cudaSuccess) return fail(GPB_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_));        \
  } while (0)


// ===================================================================== NCCL (optional, bound at run time)
// The boundary all-reduce of a sharded graph is issued by the engine itself on its own stream, so that a whole
// Gauss-Newton iteration - collective included - is ONE captured CUDA graph.  libnccl is bound with dlopen when a communicator is
// first requested (gpb_graph_init_nccl): a process that already holds NCCL (torch.distributed) shares that copy, single-GPU
// users never load it.  Only the five entry points below are used; types follow nccl.h (ncclUniqueId = 128 opaque bytes,
// ncclDouble = 8, ncclSum = 0).
namespace nccl_rt {
struct UniqueId { char internal[128]; };
typedef void* Comm;
typedef int (*GetUniqueId_t)(UniqueId*);
typedef int (*CommInitRank_t)(Comm*, int, UniqueId, int);
typedef int (*AllReduce_t)(const void*, void*, size_t, int, int, Comm, cudaStream_t);
typedef int (*CommDestroy_t)(Comm);
typedef const char* (*GetErrorString_t)(int);
static void* handle = nullptr;
static GetUniqueId_t GetUniqueId = nullptr;
static CommInitRank_t CommInitRank = nullptr;
static AllReduce_t AllReduce = nullptr;
static CommDestroy_t CommDestroy = nullptr;
static GetErrorString_t GetErrorString = nullptr;
static const char* load() {  // returns nullptr on success, else what went wrong
  if (handle) return nullptr;
  void* h = nullptr;
  if (const char* env = getenv("GPB_NCCL_LIB")) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy this process already holds (torch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return "libnccl.so.2 not found (set GPB_NCCL_LIB)";
  GetUniqueId = (GetUniqueId_t)dlsym(h, "ncclGetUniqueId"); CommInitRank = (CommInitRank_t)dlsym(h, "ncclCommInitRank");
  AllReduce = (AllReduce_t)dlsym(h, "ncclAllReduce"); CommDestroy = (CommDestroy_t)dlsym(h, "ncclCommDestroy");
  GetErrorString = (GetErrorString_t)dlsym(h, "ncclGetErrorString");
  if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy || !GetErrorString) return "libnccl lacks an expected symbol";
  handle = h;
  return nullptr;
}
}  // namespace nccl_rt

// ===================================================================== host-side graph
struct Extra {
  int kind, sa, sb, l, m, order, interval, closure;  // closure: BetweenFactor between non-adjacent states (rows at the tail of the row table)
  double prm[XP_STRIDE];
};

struct Level {
  int n = 0, M = 0, S = 0, nseg = 0, ncta = 0, ncta_bwd = 0;
  bool top = false;        // storage-only level holding the Schur complement on the pinned states (external separators, loop-closure endpoints)
  int* d_sep = nullptr;    // [S] positions of the interior separators in this level's chain
  double* rec = nullptr;   // level >= 1: [n][3 bs^2 + 2 bs]  (D1 | D2 | E | g1 | g2)
  double* brec = nullptr;  // level >= 1: [n][2 bs nb]
  double* frec = nullptr;  // [n][2 bs^2 + bs w]   (Lii | Le | Y)
  double* xsol = nullptr;  // [n][bs]
  double* cseg = nullptr;  // [ncta][nb nb + nb]
};

struct gpb_graph {
  int group = 0, D = 0, PS = 0, DL = 0, N = 0, L = 0, bs = 0, SR = 0, nb = 0, w = 0, W = 0;
  int qc_diag = 0;   // every Qc model is diagonal: SE(3) priors use the element-wise whitening kernel class
  int lin_variant = 0;  // k_lin_gp<G_POSE3> instantiation (see launch_lin_gp_pose3)
  int vw = 0;  // GPB_POSE3VW: a GPB_POSE3 graph whose velocities are [v_world | w_world] (selects the VW linearise kernels only)
  std::vector<std::vector<double>> Rq;  // chol_upper(Qc^-1), D x D column-major
  std::vector<double> dt;               // per interval (0 = no GP prior)
  std::vector<int> gp_qc;
  std::vector<Extra> extras;
  std::vector<double> h_X, h_land;   // host values before finalize()
  bool pinned = false;  // h_X / h_land are page-locked (cudaHostRegister) after finalize()
  bool finalized = false, linearized = false, assembled = false;
  int device = -1;
  int M0 = 0, Mup = 0;
  int nint = 0, NFp = 0, NX = 0, NXR = 0, NXRp = 0, ncolsX = 0, ngp = 0;
  double *d_X = nullptr, *d_Xt = nullptr, *d_land = nullptr, *d_landt = nullptr;
  double *d_dt = nullptr, *d_Rq = nullptr;
  int* d_qc = nullptr;
  double *d_AB[2] = {nullptr, nullptr}, *d_XR[2] = {nullptr, nullptr};
  int cur = 0;
  int *d_xkind = nullptr, *d_xsa = nullptr, *d_xsb = nullptr, *d_xl = nullptr, *d_xrow = nullptr;
  double* d_xprm = nullptr;
  int *d_rowoff = nullptr, *d_rowland = nullptr, *d_lmoff = nullptr, *d_lmrows = nullptr;
  int *d_bsoff = nullptr, *d_bsrow = nullptr, *d_bsside = nullptr;  // per-state CSR of landmark-bearing rows (level-0 border gather)
  double* d_bent = nullptr; int nbent = 0;                          // the same rows packed as 128-byte entries (k_border_pack)
  int rank = 0, world = 1, R = 0, sms = 148;
  int ntop = 0, P = 0;            // top states of the global reduced system / of this graph's top-level chain
  int pinL = 0, pinR = 0;         // first / last state of the chain is pinned (external separator or loop-closure endpoint)
  int* d_gtop = nullptr;          // [P] global top index of local top state k
  double* d_topx = nullptr;       // [R] solution of the reduced system
  int n_real = 0;                 // chain entries [n_real, N) are ghost replicas of remote loop-closure endpoints (sharded graphs)
  int map_ntop = -1; std::vector<int> map_local, map_gtop;  // explicit top map of a shard (gpb_graph_set_top_map)
  int nclos = 0, nep = 0, npair = 0;  // loop closures: factors, endpoint states, unique endpoint pairs
  int *d_epstate = nullptr, *d_epoff = nullptr, *d_eprow = nullptr, *d_epside = nullptr;
  int *d_pair_a = nullptr, *d_pair_b = nullptr, *d_pairoff = nullptr, *d_pairrow = nullptr;
  bool generic_fwd = false, force_blocked = false, old_assemble = false, split_levels = false, no_tiny = false, fuse_l0 = false, old_bwd = false;
  int tiny_mode = 0;
  bool asm_persist = false;   // A/B switch GPB_ASM_PERSIST: k_assemble_mma_p (persistent CTAs, next tile's copies in flight during the products)
  int asm_occ = 6;            // A/B switch GPB_ASM_OCC: resident CTAs per SM of k_assemble_mma (5 / 6 / 7)
  bool fma_syrk = false;      // A/B switch GPB_FMA_SYRK: trailing update of the multi-CTA dense solve on FP64 FMAs (round 1) instead of the tensor pipe
  bool thread_chain = false;  // 6 x 6 chains without a landmark border: thread-per-segment kernels (k_fwd6t / k_bwd6t); GPB_NO_THREAD_CHAIN = generic kernels
  int panel0_occ = 4;         // CTAs per SM the level-0 active-column panel kernel is compiled for (GPB_PANEL0_OCC = 4: 128 registers, no spills)
  bool panelw = false;        // A/B switch GPB_PANELW: register-resident two-warp level-0 panel kernel (k_panel_w)
  bool dense_panel = false;   // A/B switch GPB_DENSE_PANEL: k_panel4 (all 64 columns at every state) instead of k_panel0 (active columns only)
  unsigned char *d_lorder = nullptr, *d_ntile = nullptr;  // k_panel0: per-segment landmark order [nseg][17], active column tiles per state [N]
  int fstride = 0;  // doubles per state of a level's factor record: (L^-1 | Le), + Y for the Y-reading back-substitution
  gpb_allreduce_fn allreduce = nullptr; void* allreduce_ctx = nullptr;
  nccl_rt::Comm nccl = nullptr;   // engine-owned communicator (gpb_graph_init_nccl): the all-reduce is captured inside the iteration graph
  cudaGraphExec_t gn_graph[4] = {nullptr, nullptr, nullptr, nullptr}; int gn_graph_launches[4] = {0, 0, 0, 0};  // [parity + 2 * variant]  // whole asynchronous GN iteration per buffer parity
  bool failed = false;            // gpb_graph_finalize failed half-way: the graph can only be destroyed
  // pipelined batch interface (gpb_optimize_batch): staging buffers, copy streams, events
  double* d_stage_in[2] = {nullptr, nullptr}; double* d_stage_out[2] = {nullptr, nullptr}; double* h_batch_err = nullptr;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_in_ready[2] = {nullptr, nullptr}, ev_in_free[2] = {nullptr, nullptr}, ev_out_ready[2] = {nullptr, nullptr}, ev_out_free[2] = {nullptr, nullptr};
  double* d_topbuf = nullptr; double cur_error_local = 0; int n_allreduce = 0;
  double* d_lambda = nullptr;
  cudaGraphExec_t iter_graph[2] = {nullptr, nullptr}; int iter_graph_launches[2] = {0, 0};  // captured GN/LM trial per buffer parity
  cudaGraphExec_t dist_graph[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}; int dist_graph_launches[2] = {0, 0};  // sharded GN: [parity][before / after the all-reduce]
  int extL = 0, extR = 0;  // shard: first / last state is an external separator (owned by the global reduced system)
  int *d_listA = nullptr, *d_listB = nullptr, *d_listC = nullptr;  // extra factors by kind class: interpolated range / attitude, generic, GPS / projection
  int nA = 0, nB = 0, nC = 0;
  double* d_Csum = nullptr;
  std::vector<int> sorted_of_order, h_xrow;
  std::vector<Extra> sorted;
  double* d_HREC = nullptr;
  double *d_errpart = nullptr, *d_scal = nullptr;
  int nerrpart = 0;
  int* d_flag = nullptr;
  double *d_Cbase = nullptr, *d_xlm = nullptr, *d_Cpart = nullptr;
  std::vector<Level> levels;
  cudaStream_t stream = nullptr, stream2 = nullptr;  // stream2: forked side branch (joined back before anything consumes its output)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
  bool chunk_lin = false;  // A/B switch GPB_CHUNK: linearise and assemble as a chunked pipeline (measured slower: DESIGN.md §7) instead of two whole-graph stages
  double cur_error = 0;
  int launches = 0;
  size_t hbm_bytes = 0;
  std::vector<void*> allocs;
};

constexpr int BS_PANEL_C0 = 13;  // k_panel0: first landmark column (12 spike columns + the right-hand side)
static int pose_storage(int group, int D) { return group == GPB_POSE3 ? 12 : group == GPB_ROT3 ? 9 : group == GPB_POSE2 ? 3 : D; }
static int land_dim(int group) { return group == GPB_POSE3 ? 3 : group == GPB_ROT3 ? 0 : 2; }
static int extra_rows_of(const gpb_graph* g, int kind) {
  switch (kind) {
    case X_INTERP_RANGE: case X_RANGE_2D: return 1;
    case X_INTERP_ATTITUDE: case X_RANGE_BEARING_2D: case X_INTERP_PROJECTION: return 2;
    case X_INTERP_GPS: case X_INTERP_GPS_VW: return 3;
    case X_PRIOR_POSE: case X_PRIOR_VEL: case X_BETWEEN: return g->D;
    case X_PRIOR_LANDMARK: return g->DL;
    case X_ODOMETRY_2D: return 3;
  }
  return 0;
}

template <class T> static int dev_alloc(gpb_graph* g, T** p, size_t count) {
  void* q = nullptr;
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  CUDA_TRY(cudaMalloc(&q, bytes));
  g->allocs.push_back(q);
  g->hbm_bytes += bytes;
  *p = (T*)q;
  return GPB_OK;
}
template <class T> static int dev_upload(gpb_graph* g, T** p, const std::vector<T>& v) {
  int rc = dev_alloc(g, p, v.size());
  if (rc) return rc;
  if (!v.empty()) CUDA_TRY(cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return GPB_OK;
}

// upper Cholesky of the inverse of a small SPD matrix (column-major), for Rq = chol(Qc^-1)
static bool chol_upper_of_inverse(const double* Q, int n, double* R) {
  std::vector<double> A(Q, Q + n * n), I(n * n, 0.0);
  for (int k = 0; k < n; k++) I[k + k * n] = 1.0;
  for (int c = 0; c < n; c++) {  // Gauss-Jordan with partial pivoting
    int p = c;
    for (int r = c + 1; r < n; r++) if (std::fabs(A[r + c * n]) > std::fabs(A[p + c * n])) p = r;
    if (A[p + c * n] == 0.0) return false;
    if (p != c) for (int k = 0; k < n; k++) { std::swap(A[c + k * n], A[p + k * n]); std::swap(I[c + k * n], I[p + k * n]); }
    const double d = 1.0 / A[c + c * n];
    for (int k = 0; k < n; k++) { A[c + k * n] *= d; I[c + k * n] *= d; }
    for (int r = 0; r < n; r++) if (r != c) { const double f = A[r + c * n]; if (f != 0) for (int k = 0; k < n; k++) { A[r + k * n] -= f * A[c + k * n]; I[r + k * n] -= f * I[c + k * n]; } }
  }
  for (int k = 0; k < n * n; k++) R[k] = 0.0;
  for (int j = 0; j < n; j++) {
    double d = I[j + j * n];
    for (int k = 0; k < j; k++) d -= R[k + j * n] * R[k + j * n];
    if (!(d > 0)) return false;
    R[j + j * n] = std::sqrt(d);
    for (int c = j + 1; c < n; c++) { double t = I[j + c * n]; for (int k = 0; k < j; k++) t -= R[k + j * n] * R[k + c * n]; R[j + c * n] = t / R[j + j * n]; }
  }
  return true;
}

// ===================================================================== C ABI: construction
extern "C" {

const char* gpb_last_error(void) { return g_err.c_str(); }
int gpb_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
void gpb_default_params(gpb_params* p, int use_lm) {
  p->max_iterations = 100; p->rel_tol = 1e-5; p->abs_tol = 1e-5; p->err_tol = 0.0; p->lambda_initial = 1e-5; p->lambda_factor = 10.0;
  p->lambda_upper = 1e5; p->lambda_lower = 0.0; p->min_model_fidelity = 1e-3; p->use_lm = use_lm;
}

gpb_graph* gpb_graph_create(int group, int dim, int n_states, int n_landmarks) {
  if (group < 0 || group > GPB_POSE3VW || n_states < 2 || n_landmarks < 0) { fail(GPB_ERR_ARG, "gpb_graph_create: bad arguments (need group 0..4, n_states >= 2)"); return nullptr; }
  if (group == GPB_LINEAR && dim != 3) { fail(GPB_ERR_UNSUPPORTED, "gpb_graph_create: GPB_LINEAR supports dim 3 (the reference's 2DLinear states)"); return nullptr; }
  gpb_graph* g = new gpb_graph();
  if (group == GPB_POSE3VW) { g->vw = 1; group = GPB_POSE3; }
  g->group = group; g->D = group == GPB_POSE3 ? 6 : 3; g->PS = pose_storage(group, g->D); g->DL = land_dim(group);
  g->N = n_states; g->L = g->DL ? n_landmarks : 0; g->bs = 2 * g->D; g->SR = g->PS + g->D; g->nint = n_states - 1;
  g->dt.assign(g->nint, 0.0); g->gp_qc.assign(g->nint, 0);
  g->h_X.assign((size_t)n_states * g->SR, 0.0); g->h_land.assign((size_t)g->L * std::max(g->DL, 1), 0.0);
  return g;
}

int gpb_graph_group(const gpb_graph* g) { return !g ? GPB_ERR_ARG : (g->vw ? GPB_POSE3VW : g->group); }

void gpb_graph_destroy(gpb_graph* g) {
  if (!g) return;
  if (g->device >= 0) cudaSetDevice(g->device);
  if (g->pinned) { cudaHostUnregister(g->h_X.data()); if (g->L) cudaHostUnregister(g->h_land.data()); }
  for (int k = 0; k < 2; k++) if (g->iter_graph[k]) cudaGraphExecDestroy(g->iter_graph[k]);
  for (int k = 0; k < 2; k++) for (int h = 0; h < 2; h++) if (g->dist_graph[k][h]) cudaGraphExecDestroy(g->dist_graph[k][h]);
  for (int k = 0; k < 4; k++) if (g->gn_graph[k]) cudaGraphExecDestroy(g->gn_graph[k]);
  if (g->nccl && nccl_rt::CommDestroy) { if (g->stream) cudaStreamSynchronize(g->stream); nccl_rt::CommDestroy(g->nccl); }
  for (int k = 0; k < 2; k++) {
    if (g->ev_in_ready[k]) cudaEventDestroy(g->ev_in_ready[k]);
    if (g->ev_in_free[k]) cudaEventDestroy(g->ev_in_free[k]);
    if (g->ev_out_ready[k]) cudaEventDestroy(g->ev_out_ready[k]);
    if (g->ev_out_free[k]) cudaEventDestroy(g->ev_out_free[k]);
  }
  if (g->s_h2d) cudaStreamDestroy(g->s_h2d);
  if (g->s_d2h) cudaStreamDestroy(g->s_d2h);
  if (g->h_batch_err) cudaFreeHost(g->h_batch_err);
  for (void* p : g->allocs) cudaFree(p);
  if (g->ev_fork) cudaEventDestroy(g->ev_fork);
  if (g->ev_join) cudaEventDestroy(g->ev_join);
  if (g->ev_join2) cudaEventDestroy(g->ev_join2);
  if (g->stream2) cudaStreamDestroy(g->stream2);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
}

#define CHECK_OPEN(g) do { if (!(g)) return fail(GPB_ERR_ARG, "null graph"); if ((g)->finalized) return fail(GPB_ERR_STATE, "graph already finalized"); } while (0)

int gpb_add_qc_model(gpb_graph* g, const double* Qc) {
  CHECK_OPEN(g);
  // getQc (gp/GPutils.cpp:16-20) dereferences a failed dynamic_cast when the model is not Gaussian; here a non-SPD Qc is a checked error
  std::vector<double> R(g->D * g->D);
  if (!chol_upper_of_inverse(Qc, g->D, R.data())) return fail(GPB_ERR_ARG, "gpb_add_qc_model: Qc is not symmetric positive definite");
  g->Rq.push_back(R);
  return (int)g->Rq.size() - 1;
}

int gpb_add_gp_prior(gpb_graph* g, int n, const int* i, const double* delta_t, int qc) {
  CHECK_OPEN(g);
  if (qc < 0 || qc >= (int)g->Rq.size()) return fail(GPB_ERR_ARG, "gpb_add_gp_prior: unknown Qc model id");
  for (int k = 0; k < n; k++) {
    if (i[k] < 0 || i[k] >= g->nint) return fail(GPB_ERR_ARG, "gpb_add_gp_prior: interval index out of range");
    if (!(delta_t[k] > 0.0)) return fail(GPB_ERR_ARG, "gpb_add_gp_prior: delta_t must be positive");
    if (g->dt[i[k]] > 0.0) return fail(GPB_ERR_UNSUPPORTED, "gpb_add_gp_prior: at most one GP prior per interval");
    g->dt[i[k]] = delta_t[k]; g->gp_qc[i[k]] = qc;
  }
  return GPB_OK;
}

static Extra make_extra(int kind) { Extra e; std::memset(&e, 0, sizeof(e)); e.kind = kind; e.sa = e.sb = e.l = -1; return e; }
static void set_R(Extra& e, int m, const double* R) { for (int k = 0; k < m * m; k++) e.prm[20 + k] = R[k]; }
// place a single-state factor: a-part of interval s, or b-part of interval N-2 for the last state
static void place_single(const gpb_graph* g, Extra& e, int s) {
  if (s <= g->nint - 1) { e.sa = s; e.sb = -1; e.interval = s; } else { e.sa = -1; e.sb = s; e.interval = s - 1; }
}

int gpb_add_interp_range(gpb_graph* g, int n, const int* i, const int* l, const double* z, const double* sigma, const double* delta_t,
                         const double* tau, int qc, const double* body_P_sensor) {
  CHECK_OPEN(g);
  if (g->group == GPB_ROT3) return fail(GPB_ERR_UNSUPPORTED, "gpb_add_interp_range: no range factor on Rot3 trajectories");
  if (g->vw) return fail(GPB_ERR_UNSUPPORTED, "gpb_add_interp_range: the reference has no range factor for Pose3 VW states");
  if (body_P_sensor && g->group == GPB_LINEAR) return fail(GPB_ERR_UNSUPPORTED, "GPInterpolatedRangeFactor2DLinear has no body_P_sensor");
  (void)qc;  // Lambda/Psi do not depend on Qc
  for (int k = 0; k < n; k++) {
    if (i[k] < 0 || i[k] >= g->nint || l[k] < 0 || l[k] >= g->L) return fail(GPB_ERR_ARG, "gpb_add_interp_range: index out of range");
    if (!(sigma[k] > 0.0) || !(delta_t[k] > 0.0)) return fail(GPB_ERR_ARG, "gpb_add_interp_range: sigma and delta_t must be positive");
    Extra e = make_extra(X_INTERP_RANGE);
    e.sa = i[k]; e.sb = i[k] + 1; e.l = l[k]; e.interval = i[k]; e.m = 1;
    e.prm[0] = delta_t[k]; e.prm[1] = tau[k]; e.prm[2] = z[k]; e.prm[20] = 1.0 / sigma[k];
    if (body_P_sensor) { for (int t = 0; t < g->PS; t++) e.prm[4 + t] = body_P_sensor[t]; e.prm[16] = 1.0; }
    e.order = (int)g->extras.size(); g->extras.push_back(e);
  }
  return GPB_OK;
}

int gpb_add_interp_gps(gpb_graph* g, int n, const int* i, const double* measured, const double* sqrt_info, const double* delta_t, const double* tau, int qc,
                       const double* body_P_sensor) {
  CHECK_OPEN(g);
  if (g->group != GPB_POSE3) return fail(GPB_ERR_UNSUPPORTED, "gpb_add_interp_gps: Pose3 trajectories only (GPInterpolatedGPSFactorPose3)");
  (void)qc;  // Lambda/Psi do not depend on Qc
  for (int k = 0; k < n; k++) {
    if (i[k] < 0 || i[k] >= g->nint) return fail(GPB_ERR_ARG, "gpb_add_interp_gps: index out of range");
    if (!(delta_t[k] > 0.0)) return fail(GPB_ERR_ARG, "gpb_add_interp_gps: delta_t must be positive");
    Extra e = make_extra(g->vw ? X_INTERP_GPS_VW : X_INTERP_GPS);  // GPInterpolatedGPSFactorPose3VW on a GPB_POSE3VW graph
    e.sa = i[k]; e.sb = i[k] + 1; e.interval = i[k]; e.m = 3;
    e.prm[0] = delta_t[k]; e.prm[1] = tau[k];
    for (int t = 0; t < 3; t++) e.prm[40 + t] = measured[3 * k + t];
    set_R(e, 3, sqrt_info);
    if (body_P_sensor) { for (int t = 0; t < 12; t++) e.prm[4 + t] = body_P_sensor[t]; e.prm[16] = 1.0; }
    e.order = (int)g->extras.size(); g->extras.push_back(e);
  }
  return GPB_OK;
}

int gpb_add_interp_projection(gpb_graph* g, int n, const int* i, const int* l, const double* measured, const double* sqrt_info, const double* delta_t,
                              const double* tau, int qc, const double* K, const double* body_P_sensor) {
  CHECK_OPEN(g);
  if (g->group != GPB_POSE3 || g->vw) return fail(GPB_ERR_UNSUPPORTED, "gpb_add_interp_projection: Pose3 trajectories only (GPInterpolatedProjectionFactorPose3; no VW variant in the reference)");
  if (!K) return fail(GPB_ERR_ARG, "gpb_add_interp_projection: null calibration");
  (void)qc;
  for (int k = 0; k < n; k++) {
    if (i[k] < 0 || i[k] >= g->nint || l[k] < 0 || l[k] >= g->L) return fail(GPB_ERR_ARG, "gpb_add_interp_projection: index out of range");
    if (!(delta_t[k] > 0.0)) return fail(GPB_ERR_ARG, "gpb_add_interp_projection: delta_t must be positive");
    Extra e = make_extra(X_INTERP_PROJECTION);
    e.sa = i[k]; e.sb = i[k] + 1; e.l = l[k]; e.interval = i[k]; e.m = 2;
    e.prm[0] = delta_t[k]; e.prm[1] = tau[k];
    for (int t = 0; t < 2; t++) e.prm[40 + t] = measured[2 * k + t];
    for (int t = 0; t < 5; t++) e.prm[43 + t] = K[t];
    set_R(e, 2, sqrt_info);
    if (body_P_sensor) { for (int t = 0; t < 12; t++) e.prm[4 + t] = body_P_sensor[t]; e.prm[16] = 1.0; }
    e.order = (int)g->extras.size(); g->extras.push_back(e);
  }
  return GPB_OK;
}

int gpb_add_interp_attitude(gpb_graph* g, int n, const int* i, const double* delta_t, const double* tau, int qc, const double* nZ,
                            const double* bRef, const double* sigma) {
  CHECK_OPEN(g);
  if (g->group != GPB_ROT3) return fail(GPB_ERR_UNSUPPORTED, "gpb_add_interp_attitude: Rot3 trajectories only");
  (void)qc;
  for (int k = 0; k < n; k++) {
    if (i[k] < 0 || i[k] >= g->nint) return fail(GPB_ERR_ARG, "gpb_add_interp_attitude: index out of range");
    Extra e = make_extra(X_INTERP_ATTITUDE);
    e.sa = i[k]; e.sb = i[k] + 1; e.interval = i[k]; e.m = 2;
    e.prm[0] = delta_t[k]; e.prm[1] = tau[k];
    for (int t = 0; t < 3; t++) { e.prm[4 + t] = nZ[3 * k + t]; e.prm[7 + t] = bRef[3 * k + t]; }
    e.prm[20] = 1.0 / sigma[k]; e.prm[23] = 1.0 / sigma[k];
    e.order = (int)g->extras.size(); g->extras.push_back(e);
  }
  return GPB_OK;
}

int gpb_add_prior_pose(gpb_graph* g, int i, const double* value, const double* sqrt_info) {
  CHECK_OPEN(g);
  if (i < 0 || i >= g->N) return fail(GPB_ERR_ARG, "gpb_add_prior_pose: index out of range");
  Extra e = make_extra(X_PRIOR_POSE); place_single(g, e, i); e.m = g->D;
  for (int t = 0; t < g->PS; t++) e.prm[4 + t] = value[t];
  set_R(e, g->D, sqrt_info); e.order = (int)g->extras.size(); g->extras.push_back(e);
  return GPB_OK;
}
int gpb_add_prior_vel(gpb_graph* g, int i, const double* value, const double* sqrt_info) {
  CHECK_OPEN(g);
  if (i < 0 || i >= g->N) return fail(GPB_ERR_ARG, "gpb_add_prior_vel: index out of range");
  Extra e = make_extra(X_PRIOR_VEL); place_single(g, e, i); e.m = g->D;
  for (int t = 0; t < g->D; t++) e.prm[4 + t] = value[t];
  set_R(e, g->D, sqrt_info); e.order = (int)g->extras.size(); g->extras.push_back(e);
  return GPB_OK;
}
int gpb_add_prior_landmark(gpb_graph* g, int l, const double* value, const double* sqrt_info) {
  CHECK_OPEN(g);
  if (l < 0 || l >= g->L) return fail(GPB_ERR_ARG, "gpb_add_prior_landmark: index out of range");
  Extra e = make_extra(X_PRIOR_LANDMARK); e.l = l; e.interval = 0; e.sa = 0; e.m = g->DL;
  for (int t = 0; t < g->DL; t++) e.prm[4 + t] = value[t];
  set_R(e, g->DL, sqrt_info); e.order = (int)g->extras.size(); g->extras.push_back(e);
  return GPB_OK;
}
int gpb_add_between(gpb_graph* g, int i, int j, const double* measured, const double* sqrt_info) {
  CHECK_OPEN(g);
  if (i < 0 || i >= g->N || j < 0 || j >= g->N || i == j) return fail(GPB_ERR_ARG, "gpb_add_between: index out of range");
  if (std::abs(i - j) != 1 && g->group == GPB_LINEAR) return fail(GPB_ERR_UNSUPPORTED, "gpb_add_between: loop closures need a pose group");
  Extra e = make_extra(X_BETWEEN); e.m = g->D;
  // odometry (|i-j| == 1) lives on its interval; a loop closure couples two distant states: both become pinned separators of the
  // elimination and its rows go to the tail of the row table (interval == nint)
  e.sa = std::min(i, j); e.sb = std::max(i, j); e.closure = std::abs(i - j) != 1; e.interval = e.closure ? g->nint : e.sa; e.prm[17] = (j < i) ? 1.0 : 0.0;
  for (int t = 0; t < g->PS; t++) e.prm[4 + t] = measured[t];
  set_R(e, g->D, sqrt_info); e.order = (int)g->extras.size(); g->extras.push_back(e);
  return GPB_OK;
}
int gpb_add_range_2d(gpb_graph* g, int i, int l, double z, double sigma) {
  CHECK_OPEN(g);
  if (g->group != GPB_POSE2 && g->group != GPB_LINEAR) return fail(GPB_ERR_UNSUPPORTED, "gpb_add_range_2d: Pose2 / Linear<3> trajectories only");
  if (i < 0 || i >= g->N || l < 0 || l >= g->L) return fail(GPB_ERR_ARG, "gpb_add_range_2d: index out of range");
  Extra e = make_extra(X_RANGE_2D); place_single(g, e, i); e.l = l; e.m = 1; e.prm[2] = z; e.prm[20] = 1.0 / sigma;
  e.order = (int)g->extras.size(); g->extras.push_back(e);
  return GPB_OK;
}
int gpb_add_range_bearing_2d(gpb_graph* g, int i, int l, double range, double bearing, const double* sqrt_info) {
  CHECK_OPEN(g);
  if (g->group != GPB_LINEAR) return fail(GPB_ERR_UNSUPPORTED, "gpb_add_range_bearing_2d: Linear<3> trajectories only");
  if (i < 0 || i >= g->N || l < 0 || l >= g->L) return fail(GPB_ERR_ARG, "gpb_add_range_bearing_2d: index out of range");
  Extra e = make_extra(X_RANGE_BEARING_2D); place_single(g, e, i); e.l = l; e.m = 2; e.prm[2] = range; e.prm[3] = bearing; set_R(e, 2, sqrt_info);
  e.order = (int)g->extras.size(); g->extras.push_back(e);
  return GPB_OK;
}
int gpb_add_odometry_2d(gpb_graph* g, int i, int j, const double* measured, const double* sqrt_info) {
  CHECK_OPEN(g);
  if (g->group != GPB_LINEAR) return fail(GPB_ERR_UNSUPPORTED, "gpb_add_odometry_2d: Linear<3> trajectories only");
  if (i < 0 || j != i + 1 || j >= g->N) return fail(GPB_ERR_UNSUPPORTED, "gpb_add_odometry_2d: needs consecutive states (j == i+1)");
  Extra e = make_extra(X_ODOMETRY_2D); e.sa = i; e.sb = j; e.interval = i; e.m = 3;
  for (int t = 0; t < 3; t++) e.prm[4 + t] = measured[t];
  set_R(e, 3, sqrt_info); e.order = (int)g->extras.size(); g->extras.push_back(e);
  return GPB_OK;
}

static int upload_values(gpb_graph* g) {
  CUDA_TRY(cudaMemcpyAsync(g->d_X, g->h_X.data(), g->h_X.size() * sizeof(double), cudaMemcpyHostToDevice, g->stream));
  if (g->L) CUDA_TRY(cudaMemcpyAsync(g->d_land, g->h_land.data(), (size_t)g->L * g->DL * sizeof(double), cudaMemcpyHostToDevice, g->stream));
  g->linearized = false; g->assembled = false;
  return GPB_OK;
}

// stage the caller's arrays in the (idle) trial buffer and interleave them into state records on the device: with page-locked
// caller memory (gpb_alloc_host) the copies run at PCIe rate and no host loop touches the 14 MB
static int values_via_device(gpb_graph* g, double* poses, double* vels, double* landmarks, int dir) {
  CUDA_TRY(cudaSetDevice(g->device));
  const size_t np = (size_t)g->N * g->PS, nv = (size_t)g->N * g->D;
  double *sp = g->d_Xt, *sv = g->d_Xt + np;
  const int nblk = (int)std::min<size_t>((np + nv + 255) / 256, (size_t)g->sms * 8);
  if (dir == 0) {
    CUDA_TRY(cudaMemcpyAsync(sp, poses, np * sizeof(double), cudaMemcpyHostToDevice, g->stream));
    CUDA_TRY(cudaMemcpyAsync(sv, vels, nv * sizeof(double), cudaMemcpyHostToDevice, g->stream));
    k_pack_values<<<nblk, 256, 0, g->stream>>>(sp, sv, g->d_X, g->N, g->PS, g->D, 0);
    if (landmarks && g->L) CUDA_TRY(cudaMemcpyAsync(g->d_land, landmarks, (size_t)g->L * g->DL * sizeof(double), cudaMemcpyHostToDevice, g->stream));
    g->linearized = false; g->assembled = false;
  } else {
    k_pack_values<<<nblk, 256, 0, g->stream>>>(sp, sv, g->d_X, g->N, g->PS, g->D, 1);
    CUDA_TRY(cudaMemcpyAsync(poses, sp, np * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaMemcpyAsync(vels, sv, nv * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
    if (landmarks && g->L) CUDA_TRY(cudaMemcpyAsync(landmarks, g->d_land, (size_t)g->L * g->DL * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return GPB_OK;
}
int gpb_alloc_host(void** ptr, long long bytes) {
  if (!ptr || bytes <= 0) return fail(GPB_ERR_ARG, "gpb_alloc_host: bad arguments");
  CUDA_TRY(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
  return GPB_OK;
}
int gpb_free_host(void* ptr) {
  if (ptr) CUDA_TRY(cudaFreeHost(ptr));
  return GPB_OK;
}

int gpb_set_values(gpb_graph* g, const double* poses, const double* vels, const double* landmarks) {
  if (!g) return fail(GPB_ERR_ARG, "null graph");
  if (g->finalized && poses && vels) return values_via_device(g, const_cast<double*>(poses), const_cast<double*>(vels), const_cast<double*>(landmarks), 0);
  if (g->finalized) {  // partial update: refresh the host mirror first
    CUDA_TRY(cudaSetDevice(g->device));
    CUDA_TRY(cudaMemcpyAsync(g->h_X.data(), g->d_X, g->h_X.size() * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
    if (g->L) CUDA_TRY(cudaMemcpyAsync(g->h_land.data(), g->d_land, (size_t)g->L * g->DL * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
  }
  for (int i = 0; i < g->N; i++) {
    if (poses) for (int k = 0; k < g->PS; k++) g->h_X[(size_t)i * g->SR + k] = poses[(size_t)i * g->PS + k];
    if (vels) for (int k = 0; k < g->D; k++) g->h_X[(size_t)i * g->SR + g->PS + k] = vels[(size_t)i * g->D + k];
  }
  if (landmarks && g->L) std::copy(landmarks, landmarks + (size_t)g->L * g->DL, g->h_land.begin());
  if (g->finalized) { CUDA_TRY(cudaSetDevice(g->device)); int rc = upload_values(g); if (rc) return rc; CUDA_TRY(cudaStreamSynchronize(g->stream)); }
  return GPB_OK;
}

int gpb_get_values(gpb_graph* g, double* poses, double* vels, double* landmarks) {
  if (!g) return fail(GPB_ERR_ARG, "null graph");
  if (g->finalized && poses && vels) return values_via_device(g, poses, vels, landmarks, 1);
  if (g->finalized) {
    CUDA_TRY(cudaSetDevice(g->device));
    CUDA_TRY(cudaMemcpyAsync(g->h_X.data(), g->d_X, g->h_X.size() * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
    if (g->L) CUDA_TRY(cudaMemcpyAsync(g->h_land.data(), g->d_land, (size_t)g->L * g->DL * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
  }
  for (int i = 0; i < g->N; i++) {
    if (poses) for (int k = 0; k < g->PS; k++) poses[(size_t)i * g->PS + k] = g->h_X[(size_t)i * g->SR + k];
    if (vels) for (int k = 0; k < g->D; k++) vels[(size_t)i * g->D + k] = g->h_X[(size_t)i * g->SR + g->PS + k];
  }
  if (landmarks && g->L) std::copy(g->h_land.begin(), g->h_land.begin() + (size_t)g->L * g->DL, landmarks);
  return GPB_OK;
}

int gpb_graph_set_shard(gpb_graph* g, int rank, int world, int ext_left, int ext_right) {
  CHECK_OPEN(g);
  if (world < 1 || rank < 0 || rank >= world) return fail(GPB_ERR_ARG, "gpb_graph_set_shard: bad rank/world");
  if ((rank == 0 && ext_left) || (rank == world - 1 && ext_right)) return fail(GPB_ERR_ARG, "gpb_graph_set_shard: the first shard has no left neighbour, the last no right neighbour");
  if (world > 1 && ((rank > 0 && !ext_left) || (rank < world - 1 && !ext_right))) return fail(GPB_ERR_ARG, "gpb_graph_set_shard: interior cuts need their separators");
  g->rank = rank; g->world = world; g->extL = ext_left ? 1 : 0; g->extR = ext_right ? 1 : 0;
  return GPB_OK;
}
int gpb_graph_set_top_map(gpb_graph* g, int n_real, int ntop_global, int n_pinned, const int* pinned_local, const int* pinned_gtop) {
  CHECK_OPEN(g);
  if (n_real < 2 || n_real > g->N || ntop_global < n_pinned || n_pinned < 0) return fail(GPB_ERR_ARG, "gpb_graph_set_top_map: bad sizes");
  for (int k = 0; k < n_pinned; k++) {
    if (pinned_local[k] < 0 || pinned_local[k] >= g->N || (k && pinned_local[k] <= pinned_local[k - 1])) return fail(GPB_ERR_ARG, "gpb_graph_set_top_map: local indices must be ascending and in range");
    if (pinned_gtop[k] < 0 || pinned_gtop[k] >= ntop_global) return fail(GPB_ERR_ARG, "gpb_graph_set_top_map: global top index out of range");
  }
  g->n_real = n_real; g->map_ntop = ntop_global;
  g->map_local.assign(pinned_local, pinned_local + n_pinned); g->map_gtop.assign(pinned_gtop, pinned_gtop + n_pinned);
  return GPB_OK;
}
int gpb_set_allreduce(gpb_graph* g, gpb_allreduce_fn fn, void* ctx) {
  if (!g) return fail(GPB_ERR_ARG, "null graph");
  g->allreduce = fn; g->allreduce_ctx = ctx;
  return GPB_OK;
}

int gpb_set_segment_length(gpb_graph* g, int level0, int upper) {
  CHECK_OPEN(g);
  if ((level0 && level0 < 2) || (upper && upper < 2)) return fail(GPB_ERR_ARG, "segment length must be >= 2");
  g->M0 = level0; g->Mup = upper;
  return GPB_OK;
}

}  // extern "C"
template <int BS, int W> static int occ_fwd() {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_fwd<BS, W>, (W < 32 ? 32 : W), 0) != cudaSuccess) { cudaGetLastError(); nb = 4; }
  return nb < 1 ? 1 : nb;
}
template <int BS, int W> static int occ_bwd() {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_bwd<BS, W>, (W < 32 ? 32 : W), 0) != cudaSuccess) { cudaGetLastError(); nb = 4; }
  return nb < 1 ? 1 : nb;
}
static int bwd_blocks_per_sm(int bs, int W) {
  if (bs == 12) return W == 16 ? occ_bwd<12, 16>() : W == 32 ? occ_bwd<12, 32>() : W == 64 ? occ_bwd<12, 64>() : occ_bwd<12, 128>();
  return W == 16 ? occ_bwd<6, 16>() : W == 32 ? occ_bwd<6, 32>() : W == 64 ? occ_bwd<6, 64>() : occ_bwd<6, 128>();
}
static int fwd_blocks_per_sm(int bs, int W, bool fuse_l0 = false, int panel0_occ = 4, bool panelw = false) {
  if (bs == 12 && W == 64) {
    int nb = 0;
    if (panelw) { if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_panel_w<12>, 64, 0) != cudaSuccess) { cudaGetLastError(); nb = 4; } return nb < 1 ? 1 : nb; }
    if (fuse_l0) { if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_level0_ws<12>, 160, 0) != cudaSuccess) { cudaGetLastError(); nb = 4; } return nb < 1 ? 1 : nb; }
    int nb0 = 0;
    if ((panel0_occ == 4 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb0, k_panel0<12, 4>, 128, 0) : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb0, k_panel0<12, 5>, 128, 0)) != cudaSuccess) { cudaGetLastError(); nb0 = 4; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_panel4<12>, 128, 0) != cudaSuccess) { cudaGetLastError(); nb = 4; }
    nb = std::min(nb, nb0);   // one resident wave must hold for either level-0 panel kernel
    return nb < 1 ? 1 : nb;
  }
  if (bs == 12) return W == 16 ? occ_fwd<12, 16>() : W == 32 ? occ_fwd<12, 32>() : W == 64 ? occ_fwd<12, 64>() : occ_fwd<12, 128>();
  return W == 16 ? occ_fwd<6, 16>() : W == 32 ? occ_fwd<6, 32>() : W == 64 ? occ_fwd<6, 64>() : occ_fwd<6, 128>();
}
extern "C" {
// ===================================================================== finalize: build the device-resident graph
int gpb_graph_finalize(gpb_graph* g, int device) {
  CHECK_OPEN(g);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(GPB_ERR_CUDA, "gpb_graph_finalize: no CUDA device available (this engine has no CPU fallback)"); }
  if (device < 0 || device >= ndev) return fail(GPB_ERR_ARG, "gpb_graph_finalize: bad device index");
  if (g->Rq.empty()) return fail(GPB_ERR_STATE, "gpb_graph_finalize: no Qc model registered");
  if (g->failed) return fail(GPB_ERR_STATE, "gpb_graph_finalize: an earlier finalize of this graph failed; destroy it");
  // every shape check comes before the first CUDA resource is created: a refused graph holds nothing and may be finalized again
  if (2 * g->D + g->L * g->DL + 1 > 128) return fail(GPB_ERR_UNSUPPORTED, "gpb_graph_finalize: landmark border wider than 128 - 2D - 1 columns (38 3-D / 60 2-D landmarks) is not supported by this build");
  if ((g->n_real ? g->n_real : g->N) - (g->extL ? 1 : 0) - (g->extR ? 1 : 0) < 0) return fail(GPB_ERR_ARG, "shard too small for its external separators");
  {
    bool any_closure = false;
    for (const Extra& e : g->extras) any_closure |= e.closure != 0;
    if ((any_closure || (g->n_real && g->n_real < g->N)) && g->world > 1 && g->map_ntop < 0) return fail(GPB_ERR_STATE, "gpb_graph_finalize: a sharded graph with loop closures needs gpb_graph_set_top_map");
  }
  struct FailGuard { gpb_graph* g; ~FailGuard() { if (!g->finalized) g->failed = true; } } fail_guard{g};  // anything failing from here on poisons the graph
  CUDA_TRY(cudaSetDevice(device));
  g->device = device;
  CUDA_TRY(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&g->stream2, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&g->ev_join, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&g->ev_join2, cudaEventDisableTiming));
  g->chunk_lin = getenv("GPB_CHUNK") != nullptr;
  CUDA_TRY(cudaFuncSetAttribute(k_small_solve<256, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_solve_smem(SMALL_SOLVE_MAX)));
  CUDA_TRY(cudaFuncSetAttribute(k_small_solve<256, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_solve_smem(SMALL_SOLVE_MAX)));
  CUDA_TRY(cudaFuncSetAttribute(k_small_solve<256, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_solve_smem(95)));
  CUDA_TRY(cudaFuncSetAttribute(k_small_solve<256, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_solve_smem(143)));
  const int D = g->D, bs = g->bs, DL = g->DL;
  g->nb = g->L * DL; g->w = bs + g->nb + 1;
  g->W = g->w <= 16 ? 16 : g->w <= 32 ? 32 : g->w <= 64 ? 64 : 128;   // panel width class: 64 runs the SE(3) production kernels, 128 the generic one-kernel sweep
  g->ngp = 0; for (double h : g->dt) if (h > 0) g->ngp++;
  g->NFp = (g->nint + AB_TF - 1) / AB_TF * AB_TF;  // whole tiles of the [A|b] layout (ab_off)
  // ---- sort extras by interval (stable), assign row offsets
  g->sorted = g->extras;
  std::stable_sort(g->sorted.begin(), g->sorted.end(), [](const Extra& a, const Extra& b) { return a.interval < b.interval; });
  g->NX = (int)g->sorted.size();
  g->sorted_of_order.assign(g->NX, 0);
  std::vector<int> xkind(g->NX), xsa(g->NX), xsb(g->NX), xl(g->NX), xrow(g->NX), rowoff(g->nint + 1, 0), rowland;
  std::vector<double> xprm((size_t)g->NX * XP_STRIDE);
  int nrows = 0;
  for (int k = 0; k < g->NX; k++) {
    const Extra& e = g->sorted[k];
    g->sorted_of_order[e.order] = k;
    xkind[k] = e.kind; xsa[k] = e.sa; xsb[k] = e.sb; xl[k] = e.l; xrow[k] = nrows;
    std::copy(e.prm, e.prm + XP_STRIDE, xprm.begin() + (size_t)k * XP_STRIDE);
    for (int r = 0; r < e.m; r++) rowland.push_back(e.l);
    nrows += e.m;
    if (e.interval < g->nint) rowoff[e.interval + 1] += e.m;  // loop-closure rows (interval == nint) sit behind every interval's range
  }
  for (int t = 0; t < g->nint; t++) rowoff[t + 1] += rowoff[t];
  std::vector<int> bsoff(g->N + 1, 0), bsrow, bsside;
  {
    std::vector<std::vector<std::pair<int, int>>> per(g->N);
    for (int k = 0; k < g->NX; k++) {
      const Extra& e = g->sorted[k];
      if (e.l < 0) continue;
      for (int r = 0; r < e.m; r++) {
        if (e.sa >= 0) per[e.interval].push_back({xrow[k] + r, 0});
        if (e.sb >= 0) per[e.interval + 1].push_back({xrow[k] + r, 1});
      }
    }
    for (int i = 0; i < g->N; i++) { bsoff[i + 1] = bsoff[i] + (int)per[i].size(); for (auto& pr : per[i]) { bsrow.push_back(pr.first); bsside.push_back(pr.second); } }
  }
  std::vector<int> listA, listB, listC;
  for (int k = 0; k < g->NX; k++) (xkind[k] == X_INTERP_RANGE || xkind[k] == X_INTERP_ATTITUDE ? listA : (xkind[k] == X_INTERP_GPS || xkind[k] == X_INTERP_PROJECTION || xkind[k] == X_INTERP_GPS_VW ? listC : listB)).push_back(k);
  g->nA = (int)listA.size(); g->nB = (int)listB.size(); g->nC = (int)listC.size();
  // ---- loop closures: endpoint states (pinned separators), per-endpoint and per-pair row lists
  std::vector<char> pin(g->N, 0);
  std::vector<int> epstate, epoff, eprow, epside, clos;
  for (int k = 0; k < g->NX; k++) if (g->sorted[k].closure) clos.push_back(k);
  g->nclos = (int)clos.size();
  const int n_real = g->n_real ? g->n_real : g->N;
  if ((g->nclos || n_real < g->N) && g->world > 1 && g->map_ntop < 0) return fail(GPB_ERR_STATE, "gpb_graph_finalize: a sharded graph with loop closures needs gpb_graph_set_top_map");
  for (int k : clos) { pin[g->sorted[k].sa] = 1; pin[g->sorted[k].sb] = 1; }
  if (g->extL) pin[0] = 1;
  if (g->extR) pin[n_real - 1] = 1;
  for (int i = n_real; i < g->N; i++) pin[i] = 1;
  if (g->map_ntop >= 0) {
    std::vector<char> mapped(g->N, 0);
    for (int s_ : g->map_local) { mapped[s_] = 1; pin[s_] = 1; }
    for (int i = 0; i < g->N; i++) if (pin[i] && !mapped[i]) return fail(GPB_ERR_ARG, "gpb_graph_set_top_map: a pinned state (external separator, loop-closure endpoint or ghost) has no global top index");
  }
  g->pinL = pin[0]; g->pinR = pin[g->N - 1];
  {
    std::vector<std::vector<std::pair<int, int>>> per(g->N);
    for (int k : clos) { per[g->sorted[k].sa].push_back({xrow[k], 0}); per[g->sorted[k].sb].push_back({xrow[k], 1}); }
    epoff.push_back(0);
    for (int i = 0; i < g->N; i++) if (!per[i].empty()) {
      epstate.push_back(i);
      for (auto& pr : per[i]) { eprow.push_back(pr.first); epside.push_back(pr.second); }
      epoff.push_back((int)eprow.size());
    }
    g->nep = (int)epstate.size();
  }
  g->NXR = nrows; g->NXRp = (nrows + 31) & ~31; g->h_xrow = xrow; g->ncolsX = 2 * bs + DL + 1;
  // rows per landmark
  std::vector<int> lmoff(g->L + 1, 0), lmrows;
  for (int r = 0; r < nrows; r++) if (rowland[r] >= 0) lmoff[rowland[r] + 1]++;
  for (int l = 0; l < g->L; l++) lmoff[l + 1] += lmoff[l];
  lmrows.assign(lmoff[g->L], 0);
  { std::vector<int> fill(lmoff.begin(), lmoff.end() - 1); for (int r = 0; r < nrows; r++) if (rowland[r] >= 0) lmrows[fill[rowland[r]]++] = r; }
  // ---- device buffers
  int rc;
  std::vector<double> rq((size_t)g->Rq.size() * D * D);
  for (size_t m = 0; m < g->Rq.size(); m++) std::copy(g->Rq[m].begin(), g->Rq[m].end(), rq.begin() + m * D * D);
  if ((rc = dev_alloc(g, &g->d_X, (size_t)(g->N + 1) * g->SR))) return rc;
  if ((rc = dev_alloc(g, &g->d_Xt, (size_t)(g->N + 1) * g->SR))) return rc;
  if ((rc = dev_alloc(g, &g->d_land, (size_t)g->L * std::max(DL, 1)))) return rc;
  if ((rc = dev_alloc(g, &g->d_landt, (size_t)g->L * std::max(DL, 1)))) return rc;
  if ((rc = dev_upload(g, &g->d_dt, g->dt))) return rc;
  if ((rc = dev_upload(g, &g->d_qc, g->gp_qc))) return rc;
  if ((rc = dev_upload(g, &g->d_Rq, rq))) return rc;
  for (int b = 0; b < 2; b++) {
    if ((rc = dev_alloc(g, &g->d_AB[b], (size_t)(4 * D + 1) * D * g->NFp * 2))) return rc;
    CUDA_TRY(cudaMemset(g->d_AB[b], 0, (size_t)(4 * D + 1) * D * g->NFp * 2 * sizeof(double)));  // intervals without a GP prior are never written: they stay zero
    if ((rc = dev_alloc(g, &g->d_XR[b], (size_t)g->ncolsX * std::max(g->NXRp, 32)))) return rc;
    CUDA_TRY(cudaMemset(g->d_XR[b], 0, (size_t)g->ncolsX * std::max(g->NXRp, 32) * sizeof(double)));
  }
  if ((rc = dev_upload(g, &g->d_xkind, xkind))) return rc;
  if ((rc = dev_upload(g, &g->d_xsa, xsa))) return rc;
  if ((rc = dev_upload(g, &g->d_xsb, xsb))) return rc;
  if ((rc = dev_upload(g, &g->d_xl, xl))) return rc;
  if ((rc = dev_upload(g, &g->d_xrow, xrow))) return rc;
  if ((rc = dev_upload(g, &g->d_xprm, xprm))) return rc;
  if ((rc = dev_upload(g, &g->d_bsoff, bsoff))) return rc;
  if ((rc = dev_upload(g, &g->d_bsrow, bsrow))) return rc;
  if ((rc = dev_upload(g, &g->d_bsside, bsside))) return rc;
  g->nbent = (int)bsrow.size();
  if ((rc = dev_alloc(g, &g->d_bent, (size_t)std::max(g->nbent, 1) * 16))) return rc;
  if ((rc = dev_upload(g, &g->d_listA, listA))) return rc;
  if ((rc = dev_upload(g, &g->d_listB, listB))) return rc;
  if ((rc = dev_upload(g, &g->d_listC, listC))) return rc;
  if ((rc = dev_upload(g, &g->d_rowoff, rowoff))) return rc;
  if ((rc = dev_upload(g, &g->d_rowland, rowland))) return rc;
  if ((rc = dev_upload(g, &g->d_lmoff, lmoff))) return rc;
  if ((rc = dev_upload(g, &g->d_lmrows, lmrows))) return rc;
  if ((rc = dev_alloc(g, &g->d_HREC, (size_t)g->N * (2 * bs * bs + bs)))) return rc;
  g->nerrpart = (g->nint + 127) / 128 + 2 * ((g->NX + 127) / 128) + (g->NX + 3) / 4 + (g->N + 127) / 128 + 16;
  if ((rc = dev_alloc(g, &g->d_errpart, (size_t)2 * g->nerrpart))) return rc;
  if ((rc = dev_alloc(g, &g->d_scal, 8))) return rc;
  if ((rc = dev_alloc(g, &g->d_flag, 1))) return rc;
  if ((rc = dev_alloc(g, &g->d_lambda, 1))) return rc;
  CUDA_TRY(cudaMemset(g->d_lambda, 0, sizeof(double)));
  CUDA_TRY(cudaMemset(g->d_flag, 0, sizeof(int)));
  const int centries = g->nb * g->nb + g->nb;
  if ((rc = dev_alloc(g, &g->d_Cbase, (size_t)centries))) return rc;
  CUDA_TRY(cudaMemset(g->d_Cbase, 0, std::max(centries, 1) * sizeof(double)));
  if ((rc = dev_alloc(g, &g->d_xlm, (size_t)std::max(g->nb, 1)))) return rc;
  // ---- elimination levels
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  g->sms = sms;
  // segment lengths: explicit setting > environment (tuning aid) > defaults
  const char* em0 = getenv("GPB_M0"); const char* emu = getenv("GPB_MUP");
  // A/B switches (tuning aids read once here; every non-default arm is a measured alternative recorded in DESIGN.md §7):
  g->split_levels = getenv("GPB_SPLIT_LEVELS") != nullptr;   // spine and panel as two launches on the upper levels too
  g->old_assemble = getenv("GPB_OLD_ASSEMBLE") != nullptr;   // thread-per-tile assembly instead of the DMMA kernel
  g->generic_fwd = getenv("GPB_GENERIC_FWD") != nullptr;     // one-kernel generic forward sweep (k_fwd<12,64>)
  g->fuse_l0 = getenv("GPB_FUSE_L0") != nullptr;             // level 0 as ONE warp-specialised kernel (spine warp + panel warps per CTA)
  g->old_bwd = getenv("GPB_OLD_BWD") != nullptr;             // back-substitution from a stored Y (k_bwd) instead of re-eliminating the rhs (k_bwd2)
  g->dense_panel = getenv("GPB_DENSE_PANEL") != nullptr || g->old_bwd;   // k_panel4 on all 64 columns (the Y-reading k_bwd needs its Y layout)
  g->panelw = getenv("GPB_PANELW") != nullptr && !g->dense_panel;        // two-warp register-resident panel kernel
  if (const char* ev = getenv("GPB_PANEL0_OCC")) g->panel0_occ = atoi(ev) == 5 ? 5 : 4;   // resident CTAs per SM of k_panel0
  g->no_tiny = getenv("GPB_NO_TINY_SOLVE") != nullptr;       // reduced system always through the multi-CTA dense solver
  g->fma_syrk = getenv("GPB_FMA_SYRK") != nullptr;           // dense trailing update on FP64 FMAs instead of the tensor pipe
  if (const char* ev = getenv("GPB_ASM_OCC")) { const int v = atoi(ev); if (v >= 3 && v <= 7) g->asm_occ = v; }   // resident CTAs per SM of k_assemble_mma
  g->asm_persist = getenv("GPB_ASM_PERSIST") != nullptr;     // persistent double-buffered assembly kernel
  // reduced-system solver in shared memory: 0 blocked (default) / 1 register-blocked per column (GPB_OLD_TINY) / 2 off
  g->tiny_mode = g->no_tiny ? 2 : (getenv("GPB_OLD_TINY") != nullptr ? 1 : 0);
  g->qc_diag = 1;
  for (const auto& R : g->Rq) for (int c = 0; c < D; c++) for (int r = 0; r < D; r++) if (r != c && R[r + c * D] != 0.0) g->qc_diag = 0;
  g->lin_variant = g->qc_diag ? 1 : 0;
  if (const char* ev = getenv("GPB_LIN_VARIANT")) {  // A/B switch: 0 dense Rq, 1 diagonal Rq, +2: three CTAs per SM (<= 168 registers)
    const int v = atoi(ev);
    if (v >= 0 && v <= 3 && (!(v & 1) || g->qc_diag)) g->lin_variant = v;
  }
  int M0 = g->M0 ? g->M0 : (em0 ? std::max(2, atoi(em0)) : (g->nb > 0 ? 32 : 16));
  g->thread_chain = bs == 6 && g->nb == 0 && !g->generic_fwd && !g->old_bwd && getenv("GPB_NO_THREAD_CHAIN") == nullptr;
  // one thread per segment: about one resident wave of 256 threads per SM, segments of 8 .. 64 states
  if (g->thread_chain && !g->M0 && !em0) M0 = std::min(64, std::max(8, (g->N + sms * 256 - 1) / (sms * 256)));
  // upper levels: 8 states per segment; 6 on short chains (shards of a few 10k states), where one more level of shorter
  // segments wins (measured on 12.5k / 25k / 50k / 100k states: -3 %, -2 %, 0, +1 %; gpurun_out/r1z_small_sweep2.jsonl)
  // (thread-per-segment chains: every upper level is a latency-bound pass of M states per thread - 4 measured best on C4, 1M states)
  const int Mup = g->Mup ? g->Mup : (emu ? std::max(2, atoi(emu)) : (g->thread_chain ? 4 : (g->N <= 40000 ? 6 : 8)));
  if (!g->M0 && !em0 && bs == 12 && g->W == 64) {
    // the panel kernel runs one resident wave of persistent CTAs, each walking ceil(nseg / slots) segments of M0 (+ a closing
    // separator) states one after the other: pick the segment length that minimises that serial depth (whole rounds - a
    // 100k-state chain on 148 x 5 slots wants 46, not 32); ties go to the longer segment (fewer separators for the next level)
    const int slots = sms * fwd_blocks_per_sm(bs, g->W, g->fuse_l0, g->panel0_occ, g->panelw), m0 = g->N - g->pinL - g->pinR;
    long long best = -1;
    for (int M = 12; M <= 63; M++) {
      const int nseg = (m0 > 0 ? (m0 - 1) / M : 0) + 1;
      const long long cost = (long long)((nseg + slots - 1) / slots) * (M + 2);
      if (best < 0 || cost <= best) { best = cost; M0 = M; }
    }
  }
  const int fstride = g->fstride = 2 * bs * bs + (g->old_bwd ? bs * g->w : 0);
  // Level recursion.  Each level's chain is cut at its interior separators: every pinned state (never eliminated: it stays a
  // separator at every level and ends up in the top system) and, inside every run of g ordinary states between two cuts, every
  // M-th state ((g - 1) / M of them).  The separators (plus the pinned chain ends) form the next level's chain; when no
  // ordinary separator is left, what remains is the top level: the pinned states only.
  std::vector<int> top_state;  // original state index of each top-level chain entry
  {
    std::vector<char> cpin = pin;                 // pinned flag per entry of the current chain
    std::vector<int> corig(g->N);                 // original state index per entry
    for (int i = 0; i < g->N; i++) corig[i] = i;
    int lev = 0;
    while (true) {
      const int n = (int)cpin.size();
      if (n - g->pinL - g->pinR < 0) return fail(GPB_ERR_ARG, "shard too small for its external separators");
      Level L;
      L.n = n; L.M = lev == 0 ? M0 : Mup;
      std::vector<int> sep;
      bool ordinary_sep = false;
      int a = g->pinL ? 0 : -1;                    // position of the last cut
      const int end = g->pinR ? n - 1 : n;         // position of the closing cut
      for (int pos = a + 1; pos <= end; pos++) {
        if (pos == end || cpin[pos]) {
          const int gap = pos - a - 1;
          const int ns = gap > 0 ? (gap - 1) / L.M : 0;
          for (int t = 0; t < ns; t++) { sep.push_back(a + (t + 1) * L.M); ordinary_sep = true; }
          if (pos < end) sep.push_back(pos);
          a = pos;
        }
      }
      std::sort(sep.begin(), sep.end());
      L.S = (int)sep.size(); L.nseg = L.S + 1;
      if (lev == 0 && bs == 12 && g->W == 64) {
        // k_panel0: per segment, the landmarks in order of first appearance (then the ones it never meets), and per interior
        // state the number of 8-column tiles [spike 12 | rhs 1 | 3 per landmark seen so far] holds
        constexpr int LMAX = 17;
        std::vector<unsigned char> lorder((size_t)L.nseg * LMAX, 0), ntile(n, 2);
        std::vector<int> rank(std::max(g->L, 1));
        for (int s_ = 0; s_ < L.nseg; s_++) {
          const int p_ = s_ > 0 ? sep[s_ - 1] : (g->pinL ? 0 : -1), q_ = s_ < L.S ? sep[s_] : (g->pinR ? n - 1 : -1);
          const int i0_ = p_ + 1, i1_ = q_ >= 0 ? q_ - 1 : n - 1;
          std::fill(rank.begin(), rank.end(), -1);
          int seen = 0;
          auto visit = [&](int i) { for (int e = bsoff[i]; e < bsoff[i + 1]; e++) { const int l = rowland[bsrow[e]]; if (rank[l] < 0) { rank[l] = seen; lorder[(size_t)s_ * LMAX + seen] = (unsigned char)l; seen++; } } };
          for (int i = i0_; i <= i1_; i++) { visit(i); ntile[i] = (unsigned char)((BS_PANEL_C0 + 3 * seen + 7) / 8); }
          if (q_ >= 0) visit(q_);
          for (int l = 0; l < g->L; l++) if (rank[l] < 0) { lorder[(size_t)s_ * LMAX + seen] = (unsigned char)l; seen++; }
        }
        if ((rc = dev_upload(g, &g->d_lorder, lorder))) return rc;
        if ((rc = dev_upload(g, &g->d_ntile, ntile))) return rc;
      }
      L.ncta = std::min(L.nseg, sms * fwd_blocks_per_sm(bs, g->W, g->fuse_l0 && lev == 0, g->panel0_occ, g->panelw && lev == 0));  // persistent CTAs: one resident wave
      L.ncta_bwd = std::min(L.nseg, sms * bwd_blocks_per_sm(bs, g->W));
      if ((rc = dev_upload(g, &L.d_sep, sep))) return rc;
      if ((rc = dev_alloc(g, &L.frec, (size_t)n * fstride))) return rc;
      if ((rc = dev_alloc(g, &L.xsol, (size_t)n * bs))) return rc;
      if (g->nb) { if ((rc = dev_alloc(g, &L.cseg, (size_t)L.ncta * centries))) return rc; }
      if (lev > 0) {
        if ((rc = dev_alloc(g, &L.rec, (size_t)n * (3 * bs * bs + 2 * bs)))) return rc;
        if ((rc = dev_alloc(g, &L.brec, (size_t)n * 2 * bs * std::max(g->nb, 1)))) return rc;
      }
      g->levels.push_back(L);
      // next chain: [pinned first] + separators + [pinned last]
      std::vector<char> npin; std::vector<int> norig;
      if (g->pinL) { npin.push_back(1); norig.push_back(corig[0]); }
      for (int sp : sep) { npin.push_back(cpin[sp]); norig.push_back(corig[sp]); }
      if (g->pinR) { npin.push_back(1); norig.push_back(corig[n - 1]); }
      const int n_next = (int)npin.size();
      if (!ordinary_sep) {
        if (n_next > 0) {  // storage for the Schur complement on the pinned states
          Level T; T.top = true; T.n = n_next;
          if ((rc = dev_alloc(g, &T.xsol, (size_t)n_next * bs))) return rc;
          if ((rc = dev_alloc(g, &T.rec, (size_t)n_next * (3 * bs * bs + 2 * bs)))) return rc;
          if ((rc = dev_alloc(g, &T.brec, (size_t)n_next * 2 * bs * std::max(g->nb, 1)))) return rc;
          g->levels.push_back(T);
          top_state = norig;
        }
        break;
      }
      cpin.swap(npin); corig.swap(norig); lev++;
    }
  }
  if ((rc = dev_alloc(g, &g->d_Cpart, (size_t)64 * std::max(centries, 1)))) return rc;
  CUDA_TRY(cudaMemset(g->d_Cpart, 0, (size_t)64 * std::max(centries, 1) * sizeof(double)));
  if ((rc = dev_alloc(g, &g->d_Csum, (size_t)std::max(centries, 1)))) return rc;
  // ---- the reduced (top) system: global top indices of this graph's pinned states
  g->P = (int)top_state.size();
  {
    std::vector<int> gtop(g->P);
    if (g->map_ntop >= 0) {  // explicit map (sharded graph with loop closures)
      g->ntop = g->map_ntop;
      for (int k = 0; k < g->P; k++) {
        const auto it = std::lower_bound(g->map_local.begin(), g->map_local.end(), top_state[k]);
        gtop[k] = g->map_gtop[it - g->map_local.begin()];
      }
    } else if (g->world > 1) {  // sharded: global separator r-1 is this rank's halo, r its own last state
      g->ntop = g->world - 1;
      int k = 0;
      if (g->extL) gtop[k++] = g->rank - 1;
      if (g->extR) gtop[k++] = g->rank;
    } else {
      g->ntop = g->P;
      for (int k = 0; k < g->P; k++) gtop[k] = k;
    }
    g->R = g->ntop * bs + g->nb;
    if ((rc = dev_upload(g, &g->d_gtop, gtop))) return rc;
    // unique endpoint pairs of the loop closures, in top indices
    std::vector<int> top_of(g->N, -1);
    for (int k = 0; k < g->P; k++) top_of[top_state[k]] = gtop[k];
    std::vector<std::pair<std::pair<int, int>, int>> pr;  // ((top a, top b), first row)
    for (int k : clos) pr.push_back({{top_of[g->sorted[k].sa], top_of[g->sorted[k].sb]}, xrow[k]});
    std::sort(pr.begin(), pr.end());
    std::vector<int> pair_a, pair_b, pairoff(1, 0), pairrow;
    for (size_t t = 0; t < pr.size(); t++) {
      if (pr[t].first.first < 0 || pr[t].first.second < 0) return fail(GPB_ERR_STATE, "internal: loop-closure endpoint missing from the top level");
      if (t == 0 || pr[t].first != pr[t - 1].first) { if (t) pairoff.push_back((int)pairrow.size()); pair_a.push_back(pr[t].first.first); pair_b.push_back(pr[t].first.second); }
      pairrow.push_back(pr[t].second);
    }
    if (!pr.empty()) pairoff.push_back((int)pairrow.size());
    g->npair = (int)pair_a.size();
    if ((rc = dev_upload(g, &g->d_epstate, epstate))) return rc;
    if ((rc = dev_upload(g, &g->d_epoff, epoff))) return rc;
    if ((rc = dev_upload(g, &g->d_eprow, eprow))) return rc;
    if ((rc = dev_upload(g, &g->d_epside, epside))) return rc;
    if ((rc = dev_upload(g, &g->d_pair_a, pair_a))) return rc;
    if ((rc = dev_upload(g, &g->d_pair_b, pair_b))) return rc;
    if ((rc = dev_upload(g, &g->d_pairoff, pairoff))) return rc;
    if ((rc = dev_upload(g, &g->d_pairrow, pairrow))) return rc;
  }
  if (g->world > 1 || g->P > 0) {
    if ((rc = dev_alloc(g, &g->d_topbuf, (size_t)(g->R + 1) * g->R + 4))) return rc;
    if ((rc = dev_alloc(g, &g->d_topx, (size_t)std::max(g->R, 1)))) return rc;
  }
  // page-lock the host staging so the H2D / D2H copies of the values run at full PCIe rate
  if (cudaHostRegister(g->h_X.data(), g->h_X.size() * sizeof(double), cudaHostRegisterDefault) == cudaSuccess) {
    g->pinned = true;
    if (g->L) cudaHostRegister(g->h_land.data(), g->h_land.size() * sizeof(double), cudaHostRegisterDefault);
  } else cudaGetLastError();
  g->finalized = true;
  rc = upload_values(g);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return GPB_OK;
}

}  // extern "C"

// ===================================================================== launches
#define CHECK_READY(g) do { if (!(g)) return fail(GPB_ERR_ARG, "null graph"); if (!(g)->finalized) return fail(GPB_ERR_STATE, "graph not finalized"); CUDA_TRY(cudaSetDevice((g)->device)); } while (0)

// k_lin_gp for SE(3) states: VW kernel class, or the body-velocity prior in its dense-Rq / diagonal-Rq instantiation
template <int NT> static void launch_lin_gp_pose3(gpb_graph* g, const double* X, double* AB, int wantJ) {
  const int nb1 = (g->nint + NT - 1) / NT;
  const size_t smem = (size_t)(NT + 1) * g->SR * sizeof(double);
  if (g->vw) { k_lin_gp<G_POSE3VW, NT><<<nb1, NT, smem, g->stream>>>(X, g->d_dt, g->d_qc, g->d_Rq, AB, g->d_errpart, g->nint, g->NFp, wantJ); return; }
  switch (g->lin_variant) {
    case 1: k_lin_gp<G_POSE3, NT, true, 1><<<nb1, NT, smem, g->stream>>>(X, g->d_dt, g->d_qc, g->d_Rq, AB, g->d_errpart, g->nint, g->NFp, wantJ); break;
    case 2: k_lin_gp<G_POSE3, NT, false, 3><<<nb1, NT, smem, g->stream>>>(X, g->d_dt, g->d_qc, g->d_Rq, AB, g->d_errpart, g->nint, g->NFp, wantJ); break;
    case 3: k_lin_gp<G_POSE3, NT, true, 3><<<nb1, NT, smem, g->stream>>>(X, g->d_dt, g->d_qc, g->d_Rq, AB, g->d_errpart, g->nint, g->NFp, wantJ); break;
    default: k_lin_gp<G_POSE3, NT, false, 1><<<nb1, NT, smem, g->stream>>>(X, g->d_dt, g->d_qc, g->d_Rq, AB, g->d_errpart, g->nint, g->NFp, wantJ); break;
  }
}

template <int G> static int launch_linearize(gpb_graph* g, const double* X, const double* land, int buf, int wantJ) {
  constexpr int NT = 128, SR = GroupTraits<G>::PS + GroupTraits<G>::D;
  const int nb1 = (g->nint + NT - 1) / NT, nbA = (g->nA + NT - 1) / NT, nbB = (int)(((size_t)g->nB * 32 + NT - 1) / NT);  // generic factors: one warp each
  const size_t smem = (size_t)(NT + 1) * SR * sizeof(double);
  // every measurement / prior / between factor runs on a forked stream beside the GP-prior kernel: the generic factors first
  // (few, a warp each), then the interpolated ones, whose CTAs fill the SMs the prior kernel's last partial wave leaves idle
  const bool fork = nbB > 0 || nbA > 0;
  if (fork) {
    CUDA_TRY(cudaEventRecord(g->ev_fork, g->stream));
    CUDA_TRY(cudaStreamWaitEvent(g->stream2, g->ev_fork, 0));
    if (nbB > 0) {
      k_lin_extra<G, 1, NT><<<nbB, NT, 0, g->stream2>>>(g->d_listB, g->nB, X, land, g->d_xkind, g->d_xsa, g->d_xsb, g->d_xl, g->d_xrow, g->d_xprm, g->d_XR[buf],
                                                         g->d_errpart + nb1 + nbA, g->NX, g->NXRp, wantJ);
      g->launches++;
    }
  }
  if constexpr (G == G_POSE3) launch_lin_gp_pose3<NT>(g, X, g->d_AB[buf], wantJ);
  else k_lin_gp<G, NT><<<nb1, NT, smem, g->stream>>>(X, g->d_dt, g->d_qc, g->d_Rq, g->d_AB[buf], g->d_errpart, g->nint, g->NFp, wantJ);
  g->launches++;
  if (nbA > 0) {
    k_lin_extra<G, 0, NT><<<nbA, NT, 0, g->stream2>>>(g->d_listA, g->nA, X, land, g->d_xkind, g->d_xsa, g->d_xsb, g->d_xl, g->d_xrow, g->d_xprm, g->d_XR[buf],
                                                       g->d_errpart + nb1, g->NX, g->NXRp, wantJ);
    g->launches++;
  }
  if (fork) CUDA_TRY(cudaEventRecord(g->ev_join, g->stream2));
  const int nbC = (g->nC + NT - 1) / NT;
  if constexpr (G == G_POSE3) {
    if (nbC > 0) {  // GPS / projection factors
      k_lin_extra<G, 2, NT><<<nbC, NT, 0, g->stream>>>(g->d_listC, g->nC, X, land, g->d_xkind, g->d_xsa, g->d_xsb, g->d_xl, g->d_xrow, g->d_xprm, g->d_XR[buf],
                                                        g->d_errpart + nb1 + nbA + nbB, g->NX, g->NXRp, wantJ);
      g->launches++;
    }
  }
  if (fork) CUDA_TRY(cudaStreamWaitEvent(g->stream, g->ev_join, 0));
  k_sum_partials<<<1, 256, 0, g->stream>>>(g->d_errpart, nb1 + nbA + nbB + (G == G_POSE3 ? nbC : 0), g->d_scal, 0);
  g->launches++;
  CUDA_TRY(cudaGetLastError());
  return GPB_OK;
}
static int linearize_dispatch(gpb_graph* g, const double* X, const double* land, int buf, int wantJ) {
  switch (g->group) {
    case GPB_POSE3: return launch_linearize<G_POSE3>(g, X, land, buf, wantJ);
    case GPB_POSE2: return launch_linearize<G_POSE2>(g, X, land, buf, wantJ);
    case GPB_ROT3: return launch_linearize<G_ROT3>(g, X, land, buf, wantJ);
    default: return launch_linearize<G_LINEAR>(g, X, land, buf, wantJ);
  }
}

// SE(3) assembly on the tensor pipe: `nblk` tiles of 8 states from tile `blk0`; resident CTAs per SM = g->asm_occ (register cap 96 / 80 / 72)
static void launch_assemble_mma(gpb_graph* g, int buf, int blk0, int nblk) {
  const double* xr = g->NX ? g->d_XR[buf] : nullptr;
  if (g->asm_persist && blk0 == 0) {  // persistent, double-buffered: one resident wave (two staging buffers: 43.5 KB per CTA)
    const int occ = g->asm_occ > 5 ? 5 : g->asm_occ, grid = std::min(nblk, g->sms * occ);
    if (occ == 5) k_assemble_mma_p<5><<<grid, 128, 0, g->stream>>>(g->d_AB[buf], xr, g->d_rowoff, g->d_HREC, g->N, g->NFp, g->NXRp, g->ncolsX - 1, nblk);
    else if (occ == 4) k_assemble_mma_p<4><<<grid, 128, 0, g->stream>>>(g->d_AB[buf], xr, g->d_rowoff, g->d_HREC, g->N, g->NFp, g->NXRp, g->ncolsX - 1, nblk);
    else k_assemble_mma_p<3><<<grid, 128, 0, g->stream>>>(g->d_AB[buf], xr, g->d_rowoff, g->d_HREC, g->N, g->NFp, g->NXRp, g->ncolsX - 1, nblk);
    return;
  }
  if (g->asm_occ == 5) k_assemble_mma<5><<<nblk, 128, 0, g->stream>>>(g->d_AB[buf], xr, g->d_rowoff, g->d_HREC, g->N, g->NFp, g->NXRp, g->ncolsX - 1, blk0);
  else if (g->asm_occ == 7) k_assemble_mma<7><<<nblk, 128, 0, g->stream>>>(g->d_AB[buf], xr, g->d_rowoff, g->d_HREC, g->N, g->NFp, g->NXRp, g->ncolsX - 1, blk0);
  else k_assemble_mma<6><<<nblk, 128, 0, g->stream>>>(g->d_AB[buf], xr, g->d_rowoff, g->d_HREC, g->N, g->NFp, g->NXRp, g->ncolsX - 1, blk0);
}
template <int G> static int launch_assemble(gpb_graph* g, int buf) {
  constexpr int NT = 128, bs = 2 * GroupTraits<G>::D, TILES = bs == 12 ? 4 : 1;
  const int states_per_cta = 32 * (NT / (32 * TILES));
  const int nblk = (g->N + states_per_cta - 1) / states_per_cta;
  // the landmark block and the packed border entries depend on the measurement rows only: they run on the side stream beside
  // the state-record assembly and are joined before the solve
  const bool fork = g->nb > 0;
  if (fork) {
    CUDA_TRY(cudaEventRecord(g->ev_fork, g->stream));
    CUDA_TRY(cudaStreamWaitEvent(g->stream2, g->ev_fork, 0));
    CUDA_TRY(cudaMemsetAsync(g->d_Cbase, 0, (size_t)(g->nb * g->nb + g->nb) * sizeof(double), g->stream2));
    k_landmark_base<512><<<g->L, 512, 0, g->stream2>>>(g->d_XR[buf], g->d_lmoff, g->d_lmrows, g->NXRp, 2 * bs, g->DL, g->nb, g->d_Cbase);
    g->launches++;
    if (g->nbent > 0) {  // consumed by the level-0 panel kernel (64-column panels) and by the back-substitution
      k_border_pack<<<(g->nbent * 16 + 255) / 256, 256, 0, g->stream2>>>(g->d_XR[buf], g->d_bsrow, g->d_bsside, g->d_rowland, g->nbent, bs, g->DL, g->NXRp, g->d_bent);
      g->launches++;
    }
    CUDA_TRY(cudaEventRecord(g->ev_join, g->stream2));
  }
  if (G == G_POSE3 && !g->old_assemble) launch_assemble_mma(g, buf, 0, (g->N + 7) / 8);
  else k_assemble<G, NT><<<nblk, NT, 0, g->stream>>>(g->d_AB[buf], g->d_dt, g->NX ? g->d_XR[buf] : nullptr, g->d_rowoff, g->d_HREC, g->N, g->NFp, g->NXRp);
  g->launches++;
  if (g->nep) {  // loop closures: diagonal blocks / rhs of their endpoint states
    k_assemble_closures<<<g->nep, 64, 0, g->stream>>>(g->d_XR[buf], g->NXRp, bs, GroupTraits<G>::D, g->ncolsX - 1, g->d_epstate, g->d_epoff, g->d_eprow, g->d_epside, g->d_HREC);
    g->launches++;
  }
  if (fork) CUDA_TRY(cudaStreamWaitEvent(g->stream, g->ev_join, 0));
  CUDA_TRY(cudaGetLastError());
  return GPB_OK;
}
static int assemble_dispatch(gpb_graph* g, int buf) {
  switch (g->group) {
    case GPB_POSE3: return launch_assemble<G_POSE3>(g, buf);
    case GPB_POSE2: return launch_assemble<G_POSE2>(g, buf);
    case GPB_ROT3: return launch_assemble<G_ROT3>(g, buf);
    default: return launch_assemble<G_LINEAR>(g, buf);
  }
}

// Linearise at (X, land) into buffer `buf` AND assemble its normal equations, as one pipelined stage (SE(3), body-velocity priors,
// tensor-pipe assembly): the [A|b] of the GP priors is produced in chunks of tiles and each chunk is assembled two chunks later, while
// it is still in L2 - the 240 MB of [A|b] are written once (HBM write-back) but not read back from HBM.  The other factors run on the
// side stream as in launch_linearize and are joined before the first assembly launch (their rows enter the state records).
// Measured on C3: 1.20 ms per iteration against 1.13 ms for the two whole-graph stages (sixteen small launches and their tails cost
// more than the L2 hits save) - so this is an A/B switch (GPB_CHUNK), and the default is linearize_dispatch + assemble_dispatch.
static int linearize_assemble(gpb_graph* g, const double* X, const double* land, int buf) {
  constexpr int NT = 128, NCH = 8, LA = 2;
  const int nb1 = (g->nint + NT - 1) / NT;
  const bool chunked = g->chunk_lin && g->group == GPB_POSE3 && !g->vw && !g->old_assemble && nb1 >= 8 * NCH;
  int rc;
  if (!chunked) {
    if ((rc = linearize_dispatch(g, X, land, buf, 1))) return rc;
    if ((rc = assemble_dispatch(g, buf))) return rc;
    g->assembled = true;
    return GPB_OK;
  }
  const int nbA = (g->nA + NT - 1) / NT, nbB = (int)(((size_t)g->nB * 32 + NT - 1) / NT), nbC = (g->nC + NT - 1) / NT;
  const int SR = g->SR;
  const bool fork = nbA > 0 || nbB > 0 || g->nb > 0;
  if (fork) {
    CUDA_TRY(cudaEventRecord(g->ev_fork, g->stream));
    CUDA_TRY(cudaStreamWaitEvent(g->stream2, g->ev_fork, 0));
    if (nbB > 0) { k_lin_extra<G_POSE3, 1, NT><<<nbB, NT, 0, g->stream2>>>(g->d_listB, g->nB, X, land, g->d_xkind, g->d_xsa, g->d_xsb, g->d_xl, g->d_xrow, g->d_xprm, g->d_XR[buf], g->d_errpart + nb1 + nbA, g->NX, g->NXRp, 1); g->launches++; }
    if (nbA > 0) { k_lin_extra<G_POSE3, 0, NT><<<nbA, NT, 0, g->stream2>>>(g->d_listA, g->nA, X, land, g->d_xkind, g->d_xsa, g->d_xsb, g->d_xl, g->d_xrow, g->d_xprm, g->d_XR[buf], g->d_errpart + nb1, g->NX, g->NXRp, 1); g->launches++; }
    CUDA_TRY(cudaEventRecord(g->ev_join, g->stream2));
    if (g->nb) {  // landmark block / packed border entries: need the measurement rows only
      CUDA_TRY(cudaMemsetAsync(g->d_Cbase, 0, (size_t)(g->nb * g->nb + g->nb) * sizeof(double), g->stream2));
      k_landmark_base<512><<<g->L, 512, 0, g->stream2>>>(g->d_XR[buf], g->d_lmoff, g->d_lmrows, g->NXRp, 2 * g->bs, g->DL, g->nb, g->d_Cbase); g->launches++;
      if (g->nbent > 0) { k_border_pack<<<(g->nbent * 16 + 255) / 256, 256, 0, g->stream2>>>(g->d_XR[buf], g->d_bsrow, g->d_bsside, g->d_rowland, g->nbent, g->bs, g->DL, g->NXRp, g->d_bent); g->launches++; }
      CUDA_TRY(cudaEventRecord(g->ev_join2, g->stream2));
    }
  }
  if (nbC > 0) { k_lin_extra<G_POSE3, 2, NT><<<nbC, NT, 0, g->stream>>>(g->d_listC, g->nC, X, land, g->d_xkind, g->d_xsa, g->d_xsb, g->d_xl, g->d_xrow, g->d_xprm, g->d_XR[buf], g->d_errpart + nb1 + nbA + nbB, g->NX, g->NXRp, 1); g->launches++; }
  const int per = (nb1 + NCH - 1) / NCH;                    // linearise tiles (128 factors) per chunk
  const int asm_total = (g->N + 7) / 8;                     // assembly tiles (8 states)
  auto lin_chunk = [&](int c) {
    const int b0 = c * per, b1 = std::min(nb1, b0 + per);
    if (b0 >= b1) return;
    const int f0 = b0 * NT, nf = std::min(g->nint, b1 * NT) - f0;
    const size_t smem = (size_t)(NT + 1) * SR * sizeof(double);
    double* AB = g->d_AB[buf] + ab_off(0, f0, (4 * 6 + 1) * 6);
    if (g->lin_variant == 1) k_lin_gp<G_POSE3, NT, true, 1><<<b1 - b0, NT, smem, g->stream>>>(X + (size_t)f0 * SR, g->d_dt + f0, g->d_qc + f0, g->d_Rq, AB, g->d_errpart + b0, nf, g->NFp, 1);
    else if (g->lin_variant == 2) k_lin_gp<G_POSE3, NT, false, 3><<<b1 - b0, NT, smem, g->stream>>>(X + (size_t)f0 * SR, g->d_dt + f0, g->d_qc + f0, g->d_Rq, AB, g->d_errpart + b0, nf, g->NFp, 1);
    else if (g->lin_variant == 3) k_lin_gp<G_POSE3, NT, true, 3><<<b1 - b0, NT, smem, g->stream>>>(X + (size_t)f0 * SR, g->d_dt + f0, g->d_qc + f0, g->d_Rq, AB, g->d_errpart + b0, nf, g->NFp, 1);
    else k_lin_gp<G_POSE3, NT, false, 1><<<b1 - b0, NT, smem, g->stream>>>(X + (size_t)f0 * SR, g->d_dt + f0, g->d_qc + f0, g->d_Rq, AB, g->d_errpart + b0, nf, g->NFp, 1);
    g->launches++;
  };
  auto asm_chunk = [&](int c) {
    // states whose two priors (intervals i-1, i) lie in chunks <= c: assembly tiles [16 b0, 16 b1); the last chunk takes the rest
    const int b0 = c * per, b1 = std::min(nb1, b0 + per);
    const int t0 = std::min(asm_total, 16 * b0), t1 = (c == NCH - 1 || b1 >= nb1) ? asm_total : std::min(asm_total, 16 * b1);
    if (t0 >= t1) return;
    launch_assemble_mma(g, buf, t0, t1 - t0);
    g->launches++;
  };
  for (int c = 0; c < LA; c++) lin_chunk(c);
  if (fork) CUDA_TRY(cudaStreamWaitEvent(g->stream, g->ev_join, 0));   // the measurement rows are complete
  for (int c = 0; c < NCH; c++) { asm_chunk(c); if (c + LA < NCH) lin_chunk(c + LA); }
  k_sum_partials<<<1, 256, 0, g->stream>>>(g->d_errpart, nb1 + nbA + nbB + nbC, g->d_scal, 0); g->launches++;
  if (g->nep) { k_assemble_closures<<<g->nep, 64, 0, g->stream>>>(g->d_XR[buf], g->NXRp, g->bs, 6, g->ncolsX - 1, g->d_epstate, g->d_epoff, g->d_eprow, g->d_epside, g->d_HREC); g->launches++; }
  if (fork && g->nb) CUDA_TRY(cudaStreamWaitEvent(g->stream, g->ev_join2, 0));
  CUDA_TRY(cudaGetLastError());
  g->assembled = true;
  return GPB_OK;
}

template <int BS, int W> static void launch_fwd(const FwdArgs& a, int ncta, cudaStream_t s) { k_fwd<BS, W><<<ncta, (W < 32 ? 32 : W), 0, s>>>(a); }
template <int BS, int W> static void launch_bwd(const BwdArgs& a, int ncta, cudaStream_t s) { k_bwd<BS, W><<<ncta, (W < 32 ? 32 : W), 0, s>>>(a); }
template <int BS> static void fwd_w(int W, const FwdArgs& a, int ncta, cudaStream_t s) { if (W == 16) launch_fwd<BS, 16>(a, ncta, s); else if (W == 32) launch_fwd<BS, 32>(a, ncta, s); else if (W == 64) launch_fwd<BS, 64>(a, ncta, s); else launch_fwd<BS, 128>(a, ncta, s); }
template <int BS> static void bwd_w(int W, const BwdArgs& a, int ncta, cudaStream_t s) { if (W == 16) launch_bwd<BS, 16>(a, ncta, s); else if (W == 32) launch_bwd<BS, 32>(a, ncta, s); else if (W == 64) launch_bwd<BS, 64>(a, ncta, s); else launch_bwd<BS, 128>(a, ncta, s); }

static int launch_fwd_level(gpb_graph* g, int buf, double lambda, int lev, int parts = 3) {
  const int bs = g->bs, nb = g->nb, fstride = g->fstride, centries = nb * nb + nb;
  const int nlev = (int)g->levels.size();
  Level& L = g->levels[lev];
  FwdArgs a;
  a.n = L.n; a.M = L.M; a.S = L.S; a.nseg = L.nseg; a.first_level = lev == 0; a.extL = g->pinL; a.extR = g->pinR; a.sep = L.d_sep;
  a.lamL = (g->pinL && !g->extL) ? 1 : 0;  // a pinned first state that is not a neighbour's halo is damped here
  a.nreal = g->n_real ? g->n_real : g->N;   // ghost entries beyond are damped by their owner rank
  a.rec = lev == 0 ? g->d_HREC : L.rec; a.brec = L.brec;
  a.XR = g->d_XR[buf]; a.bsoff = g->d_bsoff; a.bsrow = g->d_bsrow; a.bsside = g->d_bsside; a.rowland = g->d_rowland; a.bent = g->d_bent; a.NXRp = g->NXRp; a.nb = nb; a.DL = std::max(g->DL, 1);
  a.lambda_ptr = g->d_lambda;
  a.rec_out = lev + 1 < nlev ? g->levels[lev + 1].rec : nullptr; a.brec_out = lev + 1 < nlev ? g->levels[lev + 1].brec : nullptr;
  a.frec = L.frec; a.fstride = fstride; a.cseg = L.cseg; a.flag = g->d_flag; a.store_y = g->old_bwd ? 1 : 0;
  if (g->thread_chain) {
    if (lev == 0) k_fwd6t<true><<<(L.nseg + 63) / 64, 64, 0, g->stream>>>(a); else k_fwd6t<false><<<(L.nseg + 63) / 64, 64, 0, g->stream>>>(a);
    g->launches++;
  } else if (bs == 12 && g->W == 64 && !g->generic_fwd) {
    // spine first (warp per segment: the latency-bound 12x12 recurrence wants many independent warps), then the tensor-pipe panel
    const int spine_ctas = std::min(L.nseg, 16 * g->sms);
    if (lev == 0 && parts == 3 && g->fuse_l0) {
      k_level0_ws<12><<<L.ncta, 160, 0, g->stream>>>(a); g->launches++;
    } else if (lev > 0 && parts == 3 && L.nseg <= 2 * g->sms && !g->split_levels) {
      // small level: spine and panel pipelined inside one kernel (every CTA resident at once)
      k_level_ws<12><<<L.nseg, 160, 0, g->stream>>>(a); g->launches++;
    } else {
      if (parts & 1) { if (lev == 0) k_spine<12, true><<<spine_ctas, 32, 0, g->stream>>>(a); else k_spine<12, false><<<spine_ctas, 32, 0, g->stream>>>(a); g->launches++; }
      if (parts & 2) {
        if (lev == 0 && g->panelw) k_panel_w<12><<<L.ncta, 64, 0, g->stream>>>(a, g->d_lorder, g->d_ntile);
        else if (lev == 0 && !g->dense_panel) { if (g->panel0_occ == 4) k_panel0<12, 4><<<L.ncta, 128, 0, g->stream>>>(a, g->d_lorder, g->d_ntile); else k_panel0<12, 5><<<L.ncta, 128, 0, g->stream>>>(a, g->d_lorder, g->d_ntile); }
        else k_panel4<12><<<L.ncta, 128, 0, g->stream>>>(a);
        g->launches++;
      }
    }
  } else {
    if (bs == 12) fwd_w<12>(g->W, a, L.ncta, g->stream); else fwd_w<6>(g->W, a, L.ncta, g->stream);
    g->launches++;
  }
  (void)centries;  // the per-CTA landmark blocks of every level are summed once, in top_pack (k_cseg_reduce_all)
  return GPB_OK;
}

// number of elimination levels (excluding the storage-only top level of a sharded graph)
static int num_elim_levels(const gpb_graph* g) { return (int)g->levels.size() - (g->levels.back().top ? 1 : 0); }

static int solve_forward(gpb_graph* g, int buf, double lambda) {
  int rc;
  const int nel = num_elim_levels(g);
  for (int lev = 0; lev < nel; lev++) if ((rc = launch_fwd_level(g, buf, lambda, lev))) return rc;
  return GPB_OK;
}
static int solve_backward(gpb_graph* g) {
  const int bs = g->bs, nb = g->nb, fstride = g->fstride;
  const int nel = num_elim_levels(g), nlev = (int)g->levels.size();
  for (int lev = nel - 1; lev >= 0; lev--) {
    Level& L = g->levels[lev];
    BwdArgs b;
    b.n = L.n; b.M = L.M; b.S = L.S; b.nseg = L.nseg; b.nb = nb; b.extL = g->pinL; b.extR = g->pinR; b.sep = L.d_sep; b.frec = L.frec; b.fstride = fstride;
    b.xup = lev + 1 < nlev ? g->levels[lev + 1].xsol : nullptr; b.xl = g->d_xlm; b.xsol = L.xsol;
    b.first_level = lev == 0; b.DL = std::max(g->DL, 1); b.rec = lev == 0 ? g->d_HREC : L.rec; b.brec = L.brec; b.bsoff = g->d_bsoff; b.bent = g->d_bent;
    if (g->old_bwd) { if (bs == 12) bwd_w<12>(g->W, b, L.ncta_bwd, g->stream); else bwd_w<6>(g->W, b, L.ncta_bwd, g->stream); }
    else if (g->thread_chain) {
      if (lev == 0) k_bwd6t<true><<<(L.nseg + 127) / 128, 128, 0, g->stream>>>(b); else k_bwd6t<false><<<(L.nseg + 127) / 128, 128, 0, g->stream>>>(b);
    } else {
      // one warp per segment, four per CTA; up to 16 resident warps per SM
      // level 0: one warp per segment, four per CTA; above: a CTA per segment (dense border blocks: the warps share the rhs pass)
      if (lev == 0) {
        const int nblk = std::min((L.nseg + 3) / 4, g->sms * 4);
        if (bs == 12) k_bwd2<12, false><<<nblk, 128, 0, g->stream>>>(b); else k_bwd2<6, false><<<nblk, 128, 0, g->stream>>>(b);
      } else {
        const int nblk = std::min(L.nseg, g->sms * 4);
        if (bs == 12) k_bwd2<12, true><<<nblk, 128, 0, g->stream>>>(b); else k_bwd2<6, true><<<nblk, 128, 0, g->stream>>>(b);
      }
    }
    g->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return GPB_OK;
}
// in-place sum over ranks of a small device buffer through the registered callback (stream-ordered on both sides)
static int dist_allreduce(gpb_graph* g, double* dbuf, long long count) {
  if (g->nccl) {  // the engine's own communicator: enqueued on the engine stream (and captured with it)
    const int rc = nccl_rt::AllReduce(dbuf, dbuf, (size_t)count, /*ncclDouble*/ 8, /*ncclSum*/ 0, g->nccl, g->stream);
    if (rc != 0) return fail(GPB_ERR_CUDA, std::string("ncclAllReduce: ") + nccl_rt::GetErrorString(rc));
    g->n_allreduce++;
    return GPB_OK;
  }
  if (!g->allreduce) return fail(GPB_ERR_STATE, "sharded graph: no all-reduce registered (gpb_graph_init_nccl or gpb_set_allreduce)");
  if (g->allreduce(g->allreduce_ctx, dbuf, count, (void*)g->stream) != 0) return fail(GPB_ERR_CUDA, "all-reduce callback failed");
  g->n_allreduce++;
  return GPB_OK;
}
// Dense solve of the reduced system in d_topbuf ((R+1) x R, row R = rhs) into d_topx.
static int solve_top_dense(gpb_graph* g) {
  const int R = g->R, ld = R + 1, loff = g->ntop * g->bs;
  if (R <= SMALL_SOLVE_MAX && !g->force_blocked) {
    launch_small_solve(g->stream, g->d_topbuf, ld, g->d_topbuf + R, ld, R, loff, g->d_lambda, g->d_topx, g->d_flag, 3, g->tiny_mode);
    g->launches++;
    return GPB_OK;
  }
  for (int j0 = 0; j0 < R; j0 += DNB) {
    const int n = std::min(DNB, R - j0), below = R + 1 - (j0 + n);  // rows below the diagonal tile, rhs row included
    k_dense_diag<<<1, 256, 0, g->stream>>>(g->d_topbuf, ld, R, j0, loff, g->d_lambda, g->d_flag);
    k_dense_trsm<<<(below + 127) / 128, 128, 0, g->stream>>>(g->d_topbuf, ld, R, j0);
    g->launches += 2;
    if (j0 + n < R) {
      const int nt = (below + DNB - 1) / DNB;
      if (g->fma_syrk) k_dense_syrk<<<dim3(nt, nt), 256, 0, g->stream>>>(g->d_topbuf, ld, R, j0);
      else k_dense_syrk_mma<<<dim3(nt, nt), 128, 0, g->stream>>>(g->d_topbuf, ld, R, j0);
      g->launches++;
    }
  }
  for (int j0 = ((R - 1) / DNB) * DNB; j0 >= 0; j0 -= DNB) {
    k_dense_bwd<<<std::max(1, (j0 + 255) / 256), 256, 0, g->stream>>>(g->d_topbuf, ld, R, j0, g->d_topx);
    g->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return GPB_OK;
}
// The part of the solve between the two sweeps.  No pinned states and a single GPU: the landmark system alone.  Otherwise the
// Schur complement on {pinned states, landmarks} is packed into the reduced system (top_pack) -> (sharded graphs) ONE
// all-reduce -> dense solve, redundantly on every rank, and scatter to the top-level and landmark solutions (top_finish).
// err_local / async: see solve_system_dist.
static bool top_is_landmarks_only(const gpb_graph* g) { return g->world == 1 && g->P == 0; }
static int top_pack(gpb_graph* g, int buf, double err_local, bool async) {
  const int nb = g->nb, centries = nb * nb + nb, nel = num_elim_levels(g), R = g->R;
  if (nb) {
    CsegLevels lv; lv.n = 0;
    for (int v = 0; v < nel && lv.n < 24; v++) { lv.ptr[lv.n] = g->levels[v].cseg; lv.ncta[lv.n] = g->levels[v].ncta; lv.n++; }
    if (nel > 24) return fail(GPB_ERR_UNSUPPORTED, "more than 24 elimination levels (raise the upper segment length)");
    k_cseg_reduce_all<<<dim3((centries + 127) / 128, 64), 128, 0, g->stream>>>(lv, centries, 64, g->d_Cpart);
    k_cseg_final<<<(centries + 63) / 64, 64, 0, g->stream>>>(g->d_Cbase, g->d_Cpart, 64, centries, g->d_Csum);
    g->launches += 2;
  }
  if (top_is_landmarks_only(g)) return GPB_OK;
  const long long total = (long long)(R + 1) * R + 4;
  CUDA_TRY(cudaMemsetAsync(g->d_topbuf, 0, (size_t)total * sizeof(double), g->stream));
  PackArgs pa;
  pa.bs = g->bs; pa.nb = nb; pa.R = R; pa.ntop = g->ntop; pa.P = g->P; pa.gtop = g->d_gtop;
  pa.rec = g->levels.back().rec; pa.brec = g->levels.back().brec; pa.Csum = g->d_Csum; pa.err_local = err_local; pa.err_ptr = async ? g->d_scal : nullptr; pa.flag = g->d_flag; pa.buf = g->d_topbuf;
  k_pack_top<<<g->P + 1, 128, 0, g->stream>>>(pa);
  g->launches++;
  if (g->npair) {
    k_pack_closures<<<g->npair, 64, 0, g->stream>>>(g->d_XR[buf], g->NXRp, g->bs, g->D, R + 1, g->d_pair_a, g->d_pair_b, g->d_pairoff, g->d_pairrow, g->d_topbuf);
    g->launches++;
  }
  return GPB_OK;
}
static int top_finish(gpb_graph* g, double* sc_out /*[4] or null*/) {
  int rc;
  const int nb = g->nb, R = g->R;
  if (top_is_landmarks_only(g)) {
    if (nb) { launch_small_solve(g->stream, g->d_Csum, nb, g->d_Csum + (size_t)nb * nb, 1, nb, 0, g->d_lambda, g->d_xlm, g->d_flag, 2, g->tiny_mode); g->launches++; }
    return GPB_OK;
  }
  if (sc_out) CUDA_TRY(cudaMemcpyAsync(sc_out, g->d_topbuf + (size_t)(R + 1) * R, 4 * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  if ((rc = solve_top_dense(g))) return rc;
  const int nthr = std::max(g->P * g->bs, nb);
  k_top_scatter<<<(nthr + 127) / 128, 128, 0, g->stream>>>(g->d_topx, g->bs, nb, g->ntop, g->P, g->d_gtop, g->levels.back().xsol, g->d_xlm);
  g->launches++;
  return GPB_OK;
}
static int solve_top(gpb_graph* g, int buf, double err_local, bool async, double* sc_out) {
  int rc;
  if ((rc = top_pack(g, buf, err_local, async))) return rc;
  if (g->world > 1 && (rc = dist_allreduce(g, g->d_topbuf, (long long)(g->R + 1) * g->R + 4))) return rc;
  return top_finish(g, sc_out);
}
// sharded solve: local elimination down to the external separators -> pack -> ONE all-reduce -> redundant dense solve ->
// local back-substitution.  global_err_out: sum over ranks of err_local (the error at the current linearisation point).
// async: nothing is read back and the host is not synchronised (plain Gauss-Newton with a fixed iteration count); the local error
// of the current point is then taken from d_scal[0], where the linearise that produced this point left it.
static int solve_system_dist(gpb_graph* g, int buf, double lambda, double err_local, double* global_err_out, int* flag_out, bool async = false) {
  int rc;
  k_set_scalar<<<1, 1, 0, g->stream>>>(g->d_lambda, lambda);
  if ((rc = solve_forward(g, buf, lambda))) return rc;
  double sc[4] = {0, 0, 0, 0};
  if ((rc = solve_top(g, buf, err_local, async, async ? nullptr : sc))) return rc;
  if ((rc = solve_backward(g))) return rc;
  if (async) return GPB_OK;
  int flag = 0;
  CUDA_TRY(cudaMemcpyAsync(&flag, g->d_flag, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  if (global_err_out) *global_err_out = sc[0];
  if (flag_out) *flag_out = (sc[1] != 0.0 || flag != 0) ? 1 : 0;
  return GPB_OK;
}
// global sums of up to 4 host scalars (one tiny all-reduce; used outside the per-iteration hot loop and by LM trials)
static int dist_sum_scalars(gpb_graph* g, double* v4) {
  double* d = g->d_topbuf;  // reuse the head of the exchange buffer
  CUDA_TRY(cudaMemcpyAsync(d, v4, 4 * sizeof(double), cudaMemcpyHostToDevice, g->stream));
  int rc = dist_allreduce(g, d, 4);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(v4, d, 4 * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return GPB_OK;
}
// Solve (H + lambda I) delta = g with the current HREC / XR[buf]; delta lands in levels[0].xsol and d_xlm.
static int solve_system(gpb_graph* g, int buf, double lambda) {
  int rc;
  if (g->world > 1) return solve_system_dist(g, buf, lambda, 0.0, nullptr, nullptr);
  k_set_scalar<<<1, 1, 0, g->stream>>>(g->d_lambda, lambda);
  if ((rc = solve_forward(g, buf, lambda))) return rc;
  if ((rc = solve_top(g, buf, 0.0, false, nullptr))) return rc;
  return solve_backward(g);
}

template <int G> static int launch_retract(gpb_graph* g) {
  constexpr int NT = 128;
  const int nblk = (g->N + NT - 1) / NT;
  double* part = g->d_errpart + g->nerrpart;
  k_retract<G, NT><<<nblk, NT, 0, g->stream>>>(g->d_X, g->levels[0].xsol, g->d_HREC, g->d_Xt, part, part + nblk, g->N, g->extL ? 1 : 0, g->n_real ? g->n_real : g->N);
  k_sum_partials<<<1, 256, 0, g->stream>>>(part, nblk, g->d_scal, 1);
  k_sum_partials<<<1, 256, 0, g->stream>>>(part + nblk, nblk, g->d_scal, 2);
  g->launches += 3;
  if (g->nb) { k_retract_land<<<1, 256, 0, g->stream>>>(g->d_land, g->d_xlm, g->d_Cbase + (size_t)g->nb * g->nb, g->d_landt, g->nb, g->d_scal, g->rank == 0 ? 1 : 0); g->launches++; }
  CUDA_TRY(cudaGetLastError());
  return GPB_OK;
}
static int retract_dispatch(gpb_graph* g) {
  switch (g->group) {
    case GPB_POSE3: return launch_retract<G_POSE3>(g);
    case GPB_POSE2: return launch_retract<G_POSE2>(g);
    case GPB_ROT3: return launch_retract<G_ROT3>(g);
    default: return launch_retract<G_LINEAR>(g);
  }
}

static int read_scalars(gpb_graph* g, double* out3, int* flag) {
  double h[3];
  CUDA_TRY(cudaMemcpyAsync(h, g->d_scal, 3 * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaMemcpyAsync(flag, g->d_flag, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  out3[0] = h[0]; out3[1] = h[1]; out3[2] = h[2];
  return GPB_OK;
}

extern "C" {

int gpb_linearize(gpb_graph* g, double* error_out) {
  CHECK_READY(g);
  int rc = linearize_dispatch(g, g->d_X, g->d_land, g->cur, 1);
  if (rc) return rc;
  double s[3]; int flag;
  if ((rc = read_scalars(g, s, &flag))) return rc;
  g->cur_error_local = s[0];
  if (g->world > 1) { double v[4] = {s[0], 0, 0, 0}; if ((rc = dist_sum_scalars(g, v))) return rc; s[0] = v[0]; }
  g->cur_error = s[0]; g->linearized = true; g->assembled = false;
  if (error_out) *error_out = s[0];
  return GPB_OK;
}

int gpb_error(gpb_graph* g, double* error_out) {
  CHECK_READY(g);
  // cheap path: residuals only (the `else` branches of evaluateError, gp/GaussianProcessPriorPose3.h:73-74)
  int rc = linearize_dispatch(g, g->d_X, g->d_land, g->cur, 0);
  if (rc) return rc;
  double s[3]; int flag;
  if ((rc = read_scalars(g, s, &flag))) return rc;
  if (g->world > 1) { double v[4] = {s[0], 0, 0, 0}; if ((rc = dist_sum_scalars(g, v))) return rc; s[0] = v[0]; }
  if (error_out) *error_out = s[0];
  return GPB_OK;
}

static int ensure_assembled(gpb_graph* g) {
  int rc;
  if (!g->linearized && (rc = gpb_linearize(g, nullptr))) return rc;
  if (!g->assembled) { if ((rc = assemble_dispatch(g, g->cur))) return rc; g->assembled = true; }
  return GPB_OK;
}

int gpb_solve_delta(gpb_graph* g, double lambda, double* delta_states, double* delta_landmarks) {
  CHECK_READY(g);
  int rc = ensure_assembled(g);
  if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(g->d_flag, 0, sizeof(int), g->stream));
  if ((rc = solve_system(g, g->cur, lambda))) return rc;
  int flag = 0;
  CUDA_TRY(cudaMemcpyAsync(&flag, g->d_flag, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
  if (delta_states) CUDA_TRY(cudaMemcpyAsync(delta_states, g->levels[0].xsol, (size_t)g->N * g->bs * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  if (delta_landmarks && g->nb) CUDA_TRY(cudaMemcpyAsync(delta_landmarks, g->d_xlm, (size_t)g->nb * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  if (flag) return fail(GPB_ERR_NUMERIC, "gpb_solve_delta: system is not positive definite (indeterminate linear system)");
  return GPB_OK;
}

}  // extern "C"

// ===================================================================== GN / LM loop
namespace {
struct EventPair {  // timing events, destroyed on every exit path
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ~EventPair() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};
// stream capture that always ends: the captured graph is destroyed, the executable graph survives only on success
struct Capture {
  cudaStream_t s; bool open = false;
  explicit Capture(cudaStream_t s_) : s(s_) {}
  int begin(cudaStreamCaptureMode mode) { CUDA_TRY(cudaStreamBeginCapture(s, mode)); open = true; return GPB_OK; }
  int end(int body_rc, cudaGraphExec_t* exec) {
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(s, &graph);
    open = false;
    if (body_rc) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return body_rc; }
    if (ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return fail(GPB_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce)); }
    const cudaError_t ci = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ci != cudaSuccess) return fail(GPB_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ci));
    return GPB_OK;
  }
  ~Capture() { if (open) { cudaGraph_t graph = nullptr; cudaStreamEndCapture(s, &graph); if (graph) cudaGraphDestroy(graph); cudaGetLastError(); } }
};
}  // namespace

// One asynchronous Gauss-Newton iteration (no damping, no host round trip): assemble -> eliminate -> [sharded: ONE all-reduce of
// the boundary Schur system, which also carries the error of the current point] -> reduced solve -> back-substitute -> retract ->
// linearise at the new point, then the buffers swap.  A failed factorisation anywhere raises the sticky device flag, which the
// caller checks after its last iteration.  The launch sequence is fixed per buffer parity and is replayed as ONE CUDA graph
// (single GPU, or sharded with the engine's own NCCL communicator: the collective is captured with the kernels), or as two graphs
// around a caller-supplied all-reduce callback (gpb_set_allreduce).
static int gn_iteration_async(gpb_graph* g, bool error_only = false) {
  int r = GPB_OK;
  const int par = g->cur;
  const int var = error_only ? 1 : 0;   // graph variant: 1 = the new point's error only (residual pass without Jacobians: a batch step's values are replaced before the next solve)
  const bool dist = g->world > 1;
  const long long top_count = (long long)(g->R + 1) * g->R + 4;
  // the normal equations of the current point are assembled at the END of the previous iteration (linearize_assemble: the [A|b]
  // chunks are consumed while they are still in L2); only the first iteration of a run finds them missing
  if (!g->assembled) { if ((r = assemble_dispatch(g, par))) return r; g->assembled = true; }
  auto first_half = [&]() -> int {
    int rr = GPB_OK;
    k_set_scalar<<<1, 1, 0, g->stream>>>(g->d_lambda, 0.0); g->launches++;
    if (!rr) rr = solve_forward(g, par, 0.0);
    if (!rr) rr = top_pack(g, par, 0.0, true);
    return rr;
  };
  auto second_half = [&]() -> int {
    int rr = top_finish(g, nullptr);
    if (!rr) rr = solve_backward(g);
    if (!rr) rr = retract_dispatch(g);
    if (!rr) rr = error_only ? linearize_dispatch(g, g->d_Xt, g->d_landt, 1 - par, 0) : linearize_assemble(g, g->d_Xt, g->d_landt, 1 - par);
    return rr;
  };
  if (!dist || g->nccl) {
    if (!g->gn_graph[par + 2 * var]) {
      const int l0 = g->launches;
      Capture cap(g->stream);
      if ((r = cap.begin(dist ? cudaStreamCaptureModeRelaxed : cudaStreamCaptureModeThreadLocal))) return r;
      r = first_half();
      if (!r && dist) r = dist_allreduce(g, g->d_topbuf, top_count);
      if (!r) r = second_half();
      if ((r = cap.end(r, &g->gn_graph[par + 2 * var]))) return r;
      g->gn_graph_launches[par + 2 * var] = g->launches - l0;
      g->launches = l0;
      if (dist) g->n_allreduce--;  // counted per replay below
    }
    CUDA_TRY(cudaGraphLaunch(g->gn_graph[par + 2 * var], g->stream));
    g->launches += g->gn_graph_launches[par + 2 * var];
    if (dist) g->n_allreduce++;
  } else {
    if (!g->dist_graph[par][0]) {
      const int l0 = g->launches;
      for (int half = 0; half < 2; half++) {
        Capture cap(g->stream);
        if ((r = cap.begin(cudaStreamCaptureModeThreadLocal))) return r;
        r = half == 0 ? first_half() : second_half();
        if ((r = cap.end(r, &g->dist_graph[par][half]))) return r;
      }
      g->dist_graph_launches[par] = g->launches - l0;
      g->launches = l0;
    }
    CUDA_TRY(cudaGraphLaunch(g->dist_graph[par][0], g->stream));
    if ((r = dist_allreduce(g, g->d_topbuf, top_count))) return r;
    CUDA_TRY(cudaGraphLaunch(g->dist_graph[par][1], g->stream));
    g->launches += g->dist_graph_launches[par];
  }
  std::swap(g->d_X, g->d_Xt); std::swap(g->d_land, g->d_landt); g->cur = 1 - g->cur;
  g->assembled = !error_only; g->linearized = !error_only;
  return GPB_OK;
}

extern "C" int gpb_optimize(gpb_graph* g, const gpb_params* params, int n_iter, gpb_stats* st) {
  CHECK_READY(g);
  gpb_params p;
  if (params) p = *params; else gpb_default_params(&p, 1);
  EventPair ev;
  CUDA_TRY(cudaEventCreate(&ev.e0)); CUDA_TRY(cudaEventCreate(&ev.e1));
  CUDA_TRY(cudaEventRecord(ev.e0, g->stream));
  g->launches = 0; g->n_allreduce = 0;
  int rc;
  if (!g->linearized && (rc = gpb_linearize(g, nullptr))) return rc;
  const double error_initial = g->cur_error;
  double error = error_initial, lambda = p.lambda_initial;
  int iterations = 0, status = 0;
  const bool dist = g->world > 1;
  // Plain Gauss-Newton with a fixed iteration count needs nothing from the device between iterations: it runs asynchronously
  // (gn_iteration_async), on one GPU as on a sharded graph - there with exactly ONE all-reduce per iteration (the boundary Schur
  // system, which also carries the error of the current point).  LM trials and convergence-tested runs need the trial point's
  // (global) error before the next decision: one host sync per trial and, sharded, a second 4-double all-reduce for it.
  const bool need_trial_error = p.use_lm || n_iter <= 0;
  const bool async_gn = !need_trial_error;
  if (async_gn) CUDA_TRY(cudaMemsetAsync(g->d_flag, 0, sizeof(int), g->stream));
  auto one_iteration = [&]() -> int {
    int r;
    if (async_gn) { if ((r = gn_iteration_async(g))) return r; iterations++; return GPB_OK; }
    if (!g->assembled) { if ((r = assemble_dispatch(g, g->cur))) return r; g->assembled = true; }
    while (true) {
      CUDA_TRY(cudaMemsetAsync(g->d_flag, 0, sizeof(int), g->stream));
      int flag = 0;
      const double lam = p.use_lm ? lambda : 0.0;
      if (dist) {
        double gerr = 0;
        if ((r = solve_system_dist(g, g->cur, lam, g->cur_error_local, &gerr, &flag))) return r;
        error = gerr;  // exact global error of the current point
      } else {
        // single GPU: solve -> retract -> linearise(trial) is a fixed launch sequence per buffer parity; replay it as one CUDA
        // graph (about 45 short kernels; the launch gaps were ~10 % of the iteration).  lambda is read from device memory.
        const int par = g->cur;
        if (!g->iter_graph[par]) {
          const int l0 = g->launches;
          Capture cap(g->stream);
          if ((r = cap.begin(cudaStreamCaptureModeThreadLocal))) return r;
          k_clear_flag<<<1, 1, 0, g->stream>>>(g->d_flag);
          r = solve_forward(g, par, lam);
          if (!r) r = solve_top(g, par, 0.0, false, nullptr);
          if (!r) r = solve_backward(g);
          if (!r) r = retract_dispatch(g);
          if (!r) r = linearize_dispatch(g, g->d_Xt, g->d_landt, 1 - par, 1);
          if ((r = cap.end(r, &g->iter_graph[par]))) return r;
          g->iter_graph_launches[par] = g->launches - l0 + 1;
          g->launches = l0;
        }
        k_set_scalar<<<1, 1, 0, g->stream>>>(g->d_lambda, lam);
        CUDA_TRY(cudaGraphLaunch(g->iter_graph[par], g->stream));
        g->launches += g->iter_graph_launches[par] + 1;
      }
      if (dist) {
        if ((r = retract_dispatch(g))) return r;
        // linearise at the trial point into the other buffer: gives the trial error and, if accepted, the next iteration's [A|b]
        if ((r = linearize_dispatch(g, g->d_Xt, g->d_landt, 1 - g->cur, 1))) return r;
      }
      double s[3]; int lflag;
      if ((r = read_scalars(g, s, &lflag))) return r;
      flag |= lflag;
      const double newErrorLocal = s[0];
      if (dist) {
        double v[4] = {s[0], s[1], s[2], (double)flag};
        if ((r = dist_sum_scalars(g, v))) return r;
        s[0] = v[0]; s[1] = v[1]; s[2] = v[2]; flag = v[3] != 0.0;
      }
      const double newError = s[0];
      bool success = false, stop = false;
      if (!p.use_lm) {
        if (flag) { status = 1; return fail(GPB_ERR_NUMERIC, "GaussNewton: indeterminate linear system"); }
        success = true;
      } else if (!flag) {
        // (H + lambda I) d = g  =>  error - linearised_error(d) = g.d - 0.5 d^T H d = 0.5 g.d + 0.5 lambda |d|^2
        const double linearizedCostChange = 0.5 * s[1] + 0.5 * lambda * s[2];
        if (linearizedCostChange >= 0) {
          const double costChange = error - newError;
          double modelFidelity = 0;
          if (linearizedCostChange > 1e-20) modelFidelity = costChange / linearizedCostChange;
          success = modelFidelity > p.min_model_fidelity;
          if (std::fabs(costChange) < p.rel_tol * error) stop = true;
        }
      }
      if (success) {
        std::swap(g->d_X, g->d_Xt); std::swap(g->d_land, g->d_landt); g->cur = 1 - g->cur;
        error = newError; g->cur_error = error; g->cur_error_local = newErrorLocal; g->assembled = false; g->linearized = true;
        if (p.use_lm) lambda = std::max(p.lambda_lower, lambda / p.lambda_factor);
        break;
      } else if (!stop) {
        lambda *= p.lambda_factor;
        if (lambda >= p.lambda_upper) { status = 1; break; }
      } else break;
    }
    iterations++;
    return GPB_OK;
  };
  if (n_iter > 0) {
    for (int k = 0; k < n_iter && !status; k++) if ((rc = one_iteration())) return rc;
  } else if (!(error <= p.err_tol)) {
    do {
      const double currentError = error;
      if ((rc = one_iteration())) return rc;
      if (status) break;
      if (error <= p.err_tol) break;
      const double absDec = currentError - error, relDec = absDec / currentError;
      if ((p.rel_tol && relDec <= p.rel_tol) || absDec <= p.abs_tol) break;
    } while (iterations < p.max_iterations);
  }
  CUDA_TRY(cudaEventRecord(ev.e1, g->stream));
  CUDA_TRY(cudaEventSynchronize(ev.e1));
  if (async_gn) {  // the exact final error and the failure flag, outside the timed loop (sharded: one 4-double all-reduce)
    double s[3]; int lflag = 0;
    if ((rc = read_scalars(g, s, &lflag))) return rc;
    if (iterations) g->cur_error_local = s[0];
    double v[4] = {g->cur_error_local, 0, 0, (double)lflag};
    if (dist && (rc = dist_sum_scalars(g, v))) return rc;
    error = v[0]; g->cur_error = error;
    if (v[3] != 0.0) return fail(GPB_ERR_NUMERIC, "GaussNewton: indeterminate linear system");
  }
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->iterations = iterations; st->error_initial = error_initial; st->error_final = error; st->lambda = lambda; st->total_ms = ms; st->status = status;
  }
  return GPB_OK;
}

extern "C" {
// ---- the engine's own NCCL communicator (sharded graphs)
int gpb_nccl_unique_id(unsigned char* id128) {
  if (!id128) return fail(GPB_ERR_ARG, "gpb_nccl_unique_id: null argument");
  if (const char* why = nccl_rt::load()) return fail(GPB_ERR_UNSUPPORTED, std::string("gpb_nccl_unique_id: ") + why);
  nccl_rt::UniqueId id;
  const int rc = nccl_rt::GetUniqueId(&id);
  if (rc != 0) return fail(GPB_ERR_CUDA, std::string("ncclGetUniqueId: ") + nccl_rt::GetErrorString(rc));
  std::memcpy(id128, id.internal, 128);
  return GPB_OK;
}
int gpb_graph_init_nccl(gpb_graph* g, const unsigned char* id128, int rank, int world) {
  CHECK_READY(g);
  if (!id128 || world != g->world || rank != g->rank) return fail(GPB_ERR_ARG, "gpb_graph_init_nccl: rank / world must match gpb_graph_set_shard");
  if (g->nccl) return fail(GPB_ERR_STATE, "gpb_graph_init_nccl: communicator already created");
  if (const char* why = nccl_rt::load()) return fail(GPB_ERR_UNSUPPORTED, std::string("gpb_graph_init_nccl: ") + why);
  nccl_rt::UniqueId id;
  std::memcpy(id.internal, id128, 128);
  int rc = nccl_rt::CommInitRank(&g->nccl, world, id, rank);
  if (rc != 0) { g->nccl = nullptr; return fail(GPB_ERR_CUDA, std::string("ncclCommInitRank: ") + nccl_rt::GetErrorString(rc)); }
  // one eager all-reduce of the exchange buffer: NCCL finishes its lazy set-up (channels, buffers) outside any stream capture
  if (g->d_topbuf) {
    const long long count = (long long)(g->R + 1) * g->R + 4;
    CUDA_TRY(cudaMemsetAsync(g->d_topbuf, 0, (size_t)count * sizeof(double), g->stream));
    if ((rc = dist_allreduce(g, g->d_topbuf, count))) return rc;
    CUDA_TRY(cudaStreamSynchronize(g->stream));
  }
  return GPB_OK;
}

// ---- pipelined batch interface: K independent (values in -> one Gauss-Newton iteration -> values out) steps on the resident graph.
// The host->device copy of step k+1 and the device->host copy of step k-1 run on their own streams while step k computes
// (double-buffered device staging; all host buffers should be page-locked, gpb_alloc_host).
int gpb_optimize_batch(gpb_graph* g, int K, const double* const* poses_in, const double* const* vels_in, const double* const* land_in,
                       double* const* poses_out, double* const* vels_out, double* const* land_out, double* errors_out, gpb_stats* st) {
  CHECK_READY(g);
  if (K < 1 || !poses_in || !vels_in || !poses_out || !vels_out) return fail(GPB_ERR_ARG, "gpb_optimize_batch: bad arguments");
  if (g->L && (!land_in || !land_out)) return fail(GPB_ERR_ARG, "gpb_optimize_batch: the graph has landmarks: land_in / land_out required");
  const size_t np = (size_t)g->N * g->PS, nv = (size_t)g->N * g->D, nl = (size_t)g->L * g->DL;
  int rc;
  if (!g->s_h2d) {
    for (int b = 0; b < 2; b++) {
      if ((rc = dev_alloc(g, &g->d_stage_in[b], np + nv))) return rc;
      if ((rc = dev_alloc(g, &g->d_stage_out[b], np + nv))) return rc;
      CUDA_TRY(cudaEventCreateWithFlags(&g->ev_in_ready[b], cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&g->ev_in_free[b], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&g->ev_out_ready[b], cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&g->ev_out_free[b], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&g->s_d2h, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&g->s_h2d, cudaStreamNonBlocking));
  }
  double* herr = nullptr;
  CUDA_TRY(cudaHostAlloc((void**)&herr, (size_t)K * sizeof(double), cudaHostAllocDefault));
  struct Free { double* p; ~Free() { cudaFreeHost(p); } } free_herr{herr};
  const int nblk = (int)std::min<size_t>((np + nv + 255) / 256, (size_t)g->sms * 8);
  g->launches = 0; g->n_allreduce = 0;
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  const auto t0 = std::chrono::steady_clock::now();
  CUDA_TRY(cudaMemsetAsync(g->d_flag, 0, sizeof(int), g->stream));
  for (int k = 0; k < K; k++) {
    const int b = k & 1;
    if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(g->s_h2d, g->ev_in_free[b], 0));
    CUDA_TRY(cudaMemcpyAsync(g->d_stage_in[b], poses_in[k], np * sizeof(double), cudaMemcpyHostToDevice, g->s_h2d));
    CUDA_TRY(cudaMemcpyAsync(g->d_stage_in[b] + np, vels_in[k], nv * sizeof(double), cudaMemcpyHostToDevice, g->s_h2d));
    CUDA_TRY(cudaEventRecord(g->ev_in_ready[b], g->s_h2d));
    CUDA_TRY(cudaStreamWaitEvent(g->stream, g->ev_in_ready[b], 0));
    k_pack_values<<<nblk, 256, 0, g->stream>>>(g->d_stage_in[b], g->d_stage_in[b] + np, g->d_X, g->N, g->PS, g->D, 0); g->launches++;
    CUDA_TRY(cudaEventRecord(g->ev_in_free[b], g->stream));
    if (nl) CUDA_TRY(cudaMemcpyAsync(g->d_land, land_in[k], nl * sizeof(double), cudaMemcpyHostToDevice, g->stream));
    if ((rc = linearize_assemble(g, g->d_X, g->d_land, g->cur))) return rc;
    g->linearized = true; g->assembled = true;
    if ((rc = gn_iteration_async(g, /*error_only=*/(g->world == 1 || g->nccl != nullptr)))) return rc;
    if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(g->stream, g->ev_out_free[b], 0));
    k_pack_values<<<nblk, 256, 0, g->stream>>>(g->d_stage_out[b], g->d_stage_out[b] + np, g->d_X, g->N, g->PS, g->D, 1); g->launches++;
    CUDA_TRY(cudaMemcpyAsync(herr + k, g->d_scal, sizeof(double), cudaMemcpyDeviceToHost, g->stream));  // local error at the new point
    if (nl) CUDA_TRY(cudaMemcpyAsync(land_out[k], g->d_land, nl * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaEventRecord(g->ev_out_ready[b], g->stream));
    CUDA_TRY(cudaStreamWaitEvent(g->s_d2h, g->ev_out_ready[b], 0));
    CUDA_TRY(cudaMemcpyAsync(poses_out[k], g->d_stage_out[b], np * sizeof(double), cudaMemcpyDeviceToHost, g->s_d2h));
    CUDA_TRY(cudaMemcpyAsync(vels_out[k], g->d_stage_out[b] + np, nv * sizeof(double), cudaMemcpyDeviceToHost, g->s_d2h));
    CUDA_TRY(cudaEventRecord(g->ev_out_free[b], g->s_d2h));
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->s_d2h));
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  int flag = 0;
  CUDA_TRY(cudaMemcpy(&flag, g->d_flag, sizeof(int), cudaMemcpyDeviceToHost));
  g->cur_error_local = herr[K - 1];
  double v[4] = {herr[K - 1], 0, 0, (double)flag};
  if (g->world > 1 && (rc = dist_sum_scalars(g, v))) return rc;
  g->cur_error = v[0];
  if (errors_out) for (int k = 0; k < K; k++) errors_out[k] = herr[k];  // this rank's share of the error after step k
  if (st) { std::memset(st, 0, sizeof(*st)); st->iterations = K; st->error_final = v[0]; st->total_ms = ms; st->status = v[3] != 0.0; }
  if (v[3] != 0.0) return fail(GPB_ERR_NUMERIC, "GaussNewton: indeterminate linear system");
  return GPB_OK;
}

// testing aid: the reduced-system solver alone.  Solves (A + lambda * diag[loff..R)) x = b on `device` with the shared-memory
// solver (R <= SMALL_SOLVE_MAX and !force_blocked) or the blocked multi-CTA Cholesky.  A: R x R column-major, symmetric.
int gpb_debug_dense_solve(int device, int R, const double* A, const double* b, double lambda, int loff, int force_blocked, double* x_out) {
  if (R < 1 || !A || !b || !x_out) return fail(GPB_ERR_ARG, "gpb_debug_dense_solve: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(GPB_ERR_CUDA, "gpb_debug_dense_solve: no CUDA device available"); }
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaFuncSetAttribute(k_small_solve<256, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_solve_smem(SMALL_SOLVE_MAX)));
  CUDA_TRY(cudaFuncSetAttribute(k_small_solve<256, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_solve_smem(SMALL_SOLVE_MAX)));
  CUDA_TRY(cudaFuncSetAttribute(k_small_solve<256, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_solve_smem(95)));
  CUDA_TRY(cudaFuncSetAttribute(k_small_solve<256, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_solve_smem(143)));
  gpb_graph g;
  g.fma_syrk = getenv("GPB_FMA_SYRK") != nullptr;
  g.R = R; g.bs = 1; g.ntop = loff; g.force_blocked = force_blocked == 1; g.no_tiny = force_blocked == 2; g.tiny_mode = force_blocked == 2 ? 2 : (force_blocked == 3 ? 1 : 0);  // 2: plain per-column loops, 3: register-blocked per-column kernel, 0: blocked factorisation
  CUDA_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  int rc = GPB_OK;
  std::vector<double> T((size_t)(R + 1) * R + 4, 0.0);
  for (int c = 0; c < R; c++) { for (int r = 0; r < R; r++) T[r + (size_t)c * (R + 1)] = A[r + (size_t)c * R]; T[R + (size_t)c * (R + 1)] = b[c]; }
  if (!(rc = dev_upload(&g, &g.d_topbuf, T)) && !(rc = dev_alloc(&g, &g.d_topx, (size_t)R)) && !(rc = dev_alloc(&g, &g.d_lambda, 1)) && !(rc = dev_alloc(&g, &g.d_flag, 1))) {
    cudaMemcpy(g.d_lambda, &lambda, sizeof(double), cudaMemcpyHostToDevice);
    cudaMemset(g.d_flag, 0, sizeof(int));
    rc = solve_top_dense(&g);
    int flag = 0;
    if (!rc) {
      cudaStreamSynchronize(g.stream);
      cudaMemcpy(x_out, g.d_topx, (size_t)R * sizeof(double), cudaMemcpyDeviceToHost);
      cudaMemcpy(&flag, g.d_flag, sizeof(int), cudaMemcpyDeviceToHost);
      const cudaError_t ce = cudaGetLastError();
      if (ce != cudaSuccess) rc = fail(GPB_ERR_CUDA, std::string("gpb_debug_dense_solve: ") + cudaGetErrorString(ce));
      else if (flag) rc = fail(GPB_ERR_NUMERIC, "gpb_debug_dense_solve: matrix is not positive definite");
    }
  }
  for (void* q : g.allocs) cudaFree(q);
  cudaStreamDestroy(g.stream);
  return rc;
}

// profiling aid: FP64 tensor-pipe (mma.sync m8n8k4 / DMMA) issue-rate peak of the device, TFLOP/s, best of 5 launches - the
// roofline denominator for k_panel4 (MEASURED_PEAKS.json holds HBM and bf16 peaks only)
__global__ void __launch_bounds__(128) k_dmma_peak(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int k = 0; k < 8; k++) { c[k][0] = 0.0; c[k][1] = 0.0; }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) dmma884(c[k][0], c[k][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += c[k][0] + c[k][1];
  if (s == 123.456) out[0] = s;
}
int gpb_debug_dmma_peak(int device, double* tflops_out) {
  if (!tflops_out) return fail(GPB_ERR_ARG, "gpb_debug_dmma_peak: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(GPB_ERR_CUDA, "gpb_debug_dmma_peak: no CUDA device available"); }
  CUDA_TRY(cudaSetDevice(device));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  double* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
  const int iters = 20000, grid = sms * 8;
  double best = 0.0;
  for (int r = 0; r < 6; r++) {
    CUDA_TRY(cudaEventRecord(e0, 0));
    k_dmma_peak<<<grid, 128>>>(d, iters);
    CUDA_TRY(cudaEventRecord(e1, 0));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = (double)grid * 4 * iters * 8 * 512.0 / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  *tflops_out = best;
  return GPB_OK;
}

// profiling aid: dependent-issue latencies (clocks per operation in a chain of dependent operations, one warp on an idle SM) of the
// instructions the latency-bound solver kernels are made of: [0] DFMA, [1] DMUL, [2] DMMA m8n8k4 accumulating into the same
// fragment, [3] 64-bit __shfl_sync, [4] shared-memory load -> address of the next load, [5] rsqrt_pos (MUFU.RSQ64H + its
// third-order correction), [6] DADD, [7] global (L2-resident) load -> address of the next load
__global__ void k_latency(double* out, int iters, const int* chase) {
  __shared__ int sidx[256];
  __shared__ double res[8];
  for (int k = threadIdx.x; k < 256; k += blockDim.x) sidx[k] = (k + 1) & 255;
  __syncthreads();
  double x = 1.0 + 1e-9 * threadIdx.x, y = 0.999999, z = 1e-12;
  long long t0, t1;
  t0 = clock64();
  for (int i = 0; i < iters; i++) x = fma(x, y, z);
  t1 = clock64(); if (threadIdx.x == 0) res[0] = double(t1 - t0) / iters;
  t0 = clock64();
  for (int i = 0; i < iters; i++) x = x * y;
  t1 = clock64(); if (threadIdx.x == 0) res[1] = double(t1 - t0) / iters;
  double c0 = x, c1 = y;
  t0 = clock64();
  for (int i = 0; i < iters; i++) dmma884(c0, c1, z, y);
  t1 = clock64(); if (threadIdx.x == 0) res[2] = double(t1 - t0) / iters;
  x += c0 + c1;
  t0 = clock64();
  for (int i = 0; i < iters; i++) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  t1 = clock64(); if (threadIdx.x == 0) res[3] = double(t1 - t0) / iters;
  int j = threadIdx.x;
  t0 = clock64();
  for (int i = 0; i < iters; i++) j = sidx[j];
  t1 = clock64(); if (threadIdx.x == 0) res[4] = double(t1 - t0) / iters;
  x = fabs(x) + 1.0 + j * 1e-30;
  t0 = clock64();
  for (int i = 0; i < iters; i++) x = rsqrt_pos(x) + 1.0;
  t1 = clock64(); if (threadIdx.x == 0) res[5] = double(t1 - t0) / iters;   // includes one DADD
  t0 = clock64();
  for (int i = 0; i < iters; i++) x = x + y;
  t1 = clock64(); if (threadIdx.x == 0) res[6] = double(t1 - t0) / iters;
  int g = threadIdx.x;
  t0 = clock64();
  for (int i = 0; i < iters; i++) g = chase[g];
  t1 = clock64(); if (threadIdx.x == 0) res[7] = double(t1 - t0) / iters;
  __syncthreads();
  if (threadIdx.x < 8) out[threadIdx.x] = res[threadIdx.x];
  if (x == 123.456 || g == -7) out[8] = x;
}
int gpb_debug_latency(int device, double* out8) {
  if (!out8) return fail(GPB_ERR_ARG, "gpb_debug_latency: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(GPB_ERR_CUDA, "gpb_debug_latency: no CUDA device available"); }
  CUDA_TRY(cudaSetDevice(device));
  double* d = nullptr; int* ch = nullptr;
  CUDA_TRY(cudaMalloc(&d, 9 * sizeof(double)));
  std::vector<int> h(4096);
  for (int k = 0; k < 4096; k++) h[k] = (k * 33 + 17) & 4095;
  CUDA_TRY(cudaMalloc(&ch, h.size() * sizeof(int)));
  CUDA_TRY(cudaMemcpy(ch, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
  for (int rep = 0; rep < 2; rep++) k_latency<<<1, 32>>>(d, 2000, ch);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out8, d, 8 * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(d); cudaFree(ch);
  return GPB_OK;
}

// profiling aid: what the [A|b] store pattern of k_lin_gp<SE(3)> costs with no arithmetic in front of it - the floor the
// linearise kernel can reach with this layout.  mode 0: cudaMemsetAsync of the same bytes; 1: one thread per factor, 150
// 128-bit stores at ((c 6 + rp) NFp + f) 16 (k_lin_gp's pattern); 2: the same with st.global.cs (streaming) stores;
// 3: mode 1 with 256 threads per CTA.  Two buffers alternate (each larger than L2), best of 6.
__global__ void k_store_pattern(double* AB, int nint, int NFp, double seed, int CS) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nint) return;
  // CS 2: the tiled layout (ab_off) - consecutive row pairs 2 KB apart inside the CTA's own tile; else row pairs NFp * 16 B apart
  char* const abase = reinterpret_cast<char*>(AB) + (CS == 2 ? ab_off(0, f, 150) * 8 : (size_t)f * 16);
  const unsigned strideB = CS == 2 ? (unsigned)(AB_TF * 16) : (unsigned)NFp * 16u;
  double v = seed + f;
#pragma unroll 25
  for (int k = 0; k < 150; k++) {
    double2 d = make_double2(v, v + 0.5); v += 1.0;
    double2* p = reinterpret_cast<double2*>(abase + (size_t)((unsigned)k) * strideB);
    if (CS == 1) __stcs(p, d); else *p = d;
  }
}
int gpb_debug_store_peak(int device, int mode, int n_factors, double* us_out) {
  if (!us_out || n_factors < 1 || mode < 0 || mode > 4) return fail(GPB_ERR_ARG, "gpb_debug_store_peak: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(GPB_ERR_CUDA, "gpb_debug_store_peak: no CUDA device available"); }
  CUDA_TRY(cudaSetDevice(device));
  const int NFp = (n_factors + AB_TF - 1) / AB_TF * AB_TF;
  const size_t bytes = (size_t)150 * NFp * 16;
  double* buf[2] = {nullptr, nullptr};
  CUDA_TRY(cudaMalloc(&buf[0], bytes)); CUDA_TRY(cudaMalloc(&buf[1], bytes));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
  double best = 1e30;
  for (int r = 0; r < 7; r++) {
    double* b = buf[r & 1];
    CUDA_TRY(cudaEventRecord(e0, 0));
    const int nt = mode == 3 ? 256 : 128;
    if (mode == 0) CUDA_TRY(cudaMemsetAsync(b, 0, bytes, 0));
    else if (mode == 2) k_store_pattern<<<(n_factors + nt - 1) / nt, nt>>>(b, n_factors, NFp, (double)r, 1);
    else if (mode == 4) k_store_pattern<<<(n_factors + nt - 1) / nt, nt>>>(b, n_factors, NFp, (double)r, 2);
    else k_store_pattern<<<(n_factors + nt - 1) / nt, nt>>>(b, n_factors, NFp, (double)r, 0);
    CUDA_TRY(cudaEventRecord(e1, 0));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 && ms * 1e3 < best) best = ms * 1e3;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf[0]); cudaFree(buf[1]);
  CUDA_TRY(cudaGetLastError());
  *us_out = best;
  return GPB_OK;
}

int gpb_kernel_launches_last_optimize(gpb_graph* g) { return g ? g->launches : 0; }
int gpb_allreduces_last_optimize(gpb_graph* g) { return g ? g->n_allreduce : 0; }
// plain cudaMemcpy (kind: 1 host->device, 2 device->host) for callers that implement gpb_allreduce_fn without a CUDA binding of their own
int gpb_stream_synchronize(void* cuda_stream) {
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  return GPB_OK;
}
int gpb_memcpy(void* dst, const void* src, long long bytes, int kind) {
  CUDA_TRY(cudaMemcpy(dst, src, (size_t)bytes, kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost));
  return GPB_OK;
}

// ===================================================================== parity copy-outs
int gpb_get_linearized_factor(gpb_graph* g, int kind, int idx, double* A_out, double* b_out, int* dims_out) {
  CHECK_READY(g);
  if (!g->linearized) return fail(GPB_ERR_STATE, "call gpb_linearize first");
  const int D = g->D, bs = g->bs, DL = g->DL;
  for (int v = 0; v < 5; v++) dims_out[v] = 0;
  if (kind == 0) {
    if (idx < 0 || idx >= g->nint || !(g->dt[idx] > 0)) return fail(GPB_ERR_ARG, "no GP prior on that interval");
    const int m = bs, ncol = 4 * D + 1;
    std::vector<double> col(2);
    for (int c = 0; c < ncol; c++)
      for (int rp = 0; rp < D; rp++) {
        CUDA_TRY(cudaMemcpy(col.data(), g->d_AB[g->cur] + ab_off(c * D + rp, idx, (4 * D + 1) * D), 2 * sizeof(double), cudaMemcpyDeviceToHost));
        for (int t = 0; t < 2; t++) { const int r = 2 * rp + t; if (c < 4 * D) A_out[(size_t)c * m + r] = col[t]; else b_out[r] = col[t]; }
      }
    for (int v = 0; v < 4; v++) dims_out[v] = D;
    return m;
  }
  if (idx < 0 || idx >= g->NX) return fail(GPB_ERR_ARG, "factor index out of range");
  const int k = g->sorted_of_order[idx];
  const Extra& e = g->sorted[k];
  const int m = e.m, row0 = g->h_xrow[k];
  std::vector<double> rows((size_t)g->ncolsX * m);
  for (int c = 0; c < g->ncolsX; c++) CUDA_TRY(cudaMemcpy(rows.data() + (size_t)c * m, g->d_XR[g->cur] + (size_t)c * g->NXRp + row0, m * sizeof(double), cudaMemcpyDeviceToHost));
  // variable order of the reference's factor
  int nv = 0, o = 0;
  auto emit = [&](int col0, int d) { for (int c = 0; c < d; c++) for (int r = 0; r < m; r++) A_out[o++] = rows[(size_t)(col0 + c) * m + r]; dims_out[nv++] = d; };
  const int offs = e.sa >= 0 ? 0 : bs;
  switch (e.kind) {
    case X_INTERP_RANGE: emit(0, D); emit(D, D); emit(bs, D); emit(bs + D, D); emit(2 * bs, DL); break;
    case X_INTERP_ATTITUDE: case X_INTERP_GPS: case X_INTERP_GPS_VW: emit(0, D); emit(D, D); emit(bs, D); emit(bs + D, D); break;
    case X_INTERP_PROJECTION: emit(0, D); emit(D, D); emit(bs, D); emit(bs + D, D); emit(2 * bs, DL); break;
    case X_PRIOR_POSE: emit(offs, D); break;
    case X_PRIOR_VEL: emit(offs + D, D); break;
    case X_PRIOR_LANDMARK: emit(2 * bs, DL); break;
    case X_BETWEEN: if (e.prm[17] != 0.0) { emit(bs, D); emit(0, D); } else { emit(0, D); emit(bs, D); } break;
    case X_ODOMETRY_2D: emit(0, D); emit(bs, D); break;
    case X_RANGE_2D: case X_RANGE_BEARING_2D: emit(offs, D); emit(2 * bs, DL); break;
  }
  for (int r = 0; r < m; r++) b_out[r] = rows[(size_t)(g->ncolsX - 1) * m + r];
  return m;
}

int gpb_get_normal_equations(gpb_graph* g, double* H, double* rhs, int n) {
  CHECK_READY(g);
  int rc = ensure_assembled(g);
  if (rc) return rc;
  const int bs = g->bs, DL = g->DL, nb = g->nb, REC = 2 * bs * bs + bs;
  if (n != g->N * bs + nb) return fail(GPB_ERR_ARG, "gpb_get_normal_equations: wrong system dimension");
  std::vector<double> rec((size_t)g->N * REC), XR((size_t)g->ncolsX * g->NXRp, 0.0), Cb((size_t)nb * nb + nb, 0.0);
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  CUDA_TRY(cudaMemcpy(rec.data(), g->d_HREC, rec.size() * sizeof(double), cudaMemcpyDeviceToHost));
  if (g->NX) CUDA_TRY(cudaMemcpy(XR.data(), g->d_XR[g->cur], XR.size() * sizeof(double), cudaMemcpyDeviceToHost));
  if (nb) CUDA_TRY(cudaMemcpy(Cb.data(), g->d_Cbase, Cb.size() * sizeof(double), cudaMemcpyDeviceToHost));
  std::fill(H, H + (size_t)n * n, 0.0); std::fill(rhs, rhs + n, 0.0);
  for (int i = 0; i < g->N; i++) {
    const double* r = rec.data() + (size_t)i * REC;
    for (int c = 0; c < bs; c++) for (int rr = 0; rr < bs; rr++) H[(size_t)(i * bs + rr) + (size_t)(i * bs + c) * n] = r[rr + c * bs];
    if (i + 1 < g->N) for (int c = 0; c < bs; c++) for (int rr = 0; rr < bs; rr++) {
      const double v = r[bs * bs + rr + c * bs];
      H[(size_t)((i + 1) * bs + rr) + (size_t)(i * bs + c) * n] = v; H[(size_t)(i * bs + c) + (size_t)((i + 1) * bs + rr) * n] = v;
    }
    for (int rr = 0; rr < bs; rr++) rhs[i * bs + rr] = r[2 * bs * bs + rr];
  }
  // border from the extra rows (host side, parity only)
  int row = 0;
  for (int k = 0; k < g->NX; k++) {
    const Extra& e = g->sorted[k];
    for (int r = 0; r < e.m; r++, row++) {
      if (e.l < 0) continue;
      const int lo = g->N * bs + e.l * DL;
      for (int d = 0; d < DL; d++) {
        const double lv = XR[(size_t)(2 * bs + d) * g->NXRp + row];
        for (int side = 0; side < 2; side++) {
          const int s = e.interval + side;
          if (s >= g->N) continue;
          for (int c = 0; c < bs; c++) { const double v = XR[(size_t)(side * bs + c) * g->NXRp + row] * lv; H[(size_t)(s * bs + c) + (size_t)(lo + d) * n] += v; H[(size_t)(lo + d) + (size_t)(s * bs + c) * n] += v; }
        }
      }
    }
  }
  for (int c = 0; c < nb; c++) { for (int r = 0; r < nb; r++) H[(size_t)(g->N * bs + r) + (size_t)(g->N * bs + c) * n] = Cb[r + (size_t)c * nb]; rhs[g->N * bs + c] = Cb[(size_t)nb * nb + c]; }
  // loop closures: their diagonal shares are already in the records (k_assemble_closures); add the cross blocks H(j, i) = A_j^T A_i
  for (int k = 0; k < g->NX; k++) {
    const Extra& e = g->sorted[k];
    if (!e.closure) continue;
    for (int q = 0; q < e.m; q++) {
      const int rw = g->h_xrow[k] + q;
      for (int r = 0; r < bs; r++) for (int c = 0; c < bs; c++) {
        const double v = XR[(size_t)(bs + r) * g->NXRp + rw] * XR[(size_t)c * g->NXRp + rw];
        H[(size_t)(e.sb * bs + r) + (size_t)(e.sa * bs + c) * n] += v; H[(size_t)(e.sa * bs + c) + (size_t)(e.sb * bs + r) * n] += v;
      }
    }
  }
  return GPB_OK;
}

int gpb_get_sizes(gpb_graph* g, gpb_sizes* s) {
  if (!g || !s) return fail(GPB_ERR_ARG, "null argument");
  const int D = g->D, bs = g->bs;
  const double state_bytes = 8.0 * g->SR, land_bytes = 8.0 * g->DL;
  // LINEARISE_BYTES = sum_states value_bytes + sum_factors (param_bytes + 8 m (sum_k d_k + 1))
  double lin = g->N * state_bytes + g->L * land_bytes + g->ngp * (8.0 + 8.0 * bs * (4 * D + 1));
  for (const Extra& e : g->extras) {
    int cols = 0; double prm = 0;
    switch (e.kind) {
      case X_INTERP_RANGE: cols = 4 * D + g->DL; prm = 44; break;
      case X_INTERP_ATTITUDE: cols = 4 * D; prm = 80; break;
      case X_INTERP_GPS: case X_INTERP_GPS_VW: cols = 4 * D; prm = 8.0 * (2 + 3 + 9) + 20; break;
      case X_INTERP_PROJECTION: cols = 4 * D + g->DL; prm = 8.0 * (2 + 2 + 4 + 5) + 20; break;
      case X_PRIOR_POSE: cols = D; prm = 8.0 * (g->PS + D * D); break;
      case X_PRIOR_VEL: cols = D; prm = 8.0 * (D + D * D); break;
      case X_PRIOR_LANDMARK: cols = g->DL; prm = 8.0 * (g->DL + g->DL * g->DL); break;
      case X_BETWEEN: cols = 2 * D; prm = 8.0 * (g->PS + D * D); break;
      case X_ODOMETRY_2D: cols = 6; prm = 8.0 * 12; break;
      case X_RANGE_2D: cols = D + 2; prm = 16; break;
      case X_RANGE_BEARING_2D: cols = D + 2; prm = 48; break;
    }
    lin += prm + 8.0 * e.m * (cols + 1);
  }
  s->linearise_bytes = lin;
  s->fused_bytes = g->N * (state_bytes + 8.0 * (2 * bs * bs + bs));
  s->solve_bytes = 2.0 * g->N * 8.0 * (2 * bs * bs + bs);
  s->hbm_bytes = (double)g->hbm_bytes;
  s->n_gp = g->ngp; s->n_extra = (int)g->extras.size(); s->n_rows = g->NXR; s->border_dim = g->nb; s->levels = (int)g->levels.size();
  return GPB_OK;
}


// Average device milliseconds of one stage over `reps` launches (CUDA events on the engine's own stream).
int gpb_time_stage(gpb_graph* g, int stage, int reps, double* ms_out) {
  CHECK_READY(g);
  if (reps < 1 || !ms_out) return fail(GPB_ERR_ARG, "gpb_time_stage: bad arguments");
  int rc = ensure_assembled(g);
  if (rc) return rc;
  if ((rc = solve_system(g, g->cur, 0.0))) return rc;  // valid delta for the retract stage; also warms up
  EventPair evp;
  CUDA_TRY(cudaEventCreate(&evp.e0)); CUDA_TRY(cudaEventCreate(&evp.e1));
  const cudaEvent_t e0 = evp.e0, e1 = evp.e1;
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  const int other = 1 - g->cur;
  constexpr int NT = 128;
  const int nb1 = (g->nint + NT - 1) / NT;
  auto gp_only = [&](auto tag) {
    constexpr int G = decltype(tag)::value; constexpr int SR = GroupTraits<G>::PS + GroupTraits<G>::D;
    if constexpr (G == G_POSE3) launch_lin_gp_pose3<NT>(g, g->d_X, g->d_AB[other], 1);
    else k_lin_gp<G, NT><<<nb1, NT, (size_t)(NT + 1) * SR * sizeof(double), g->stream>>>(g->d_X, g->d_dt, g->d_qc, g->d_Rq, g->d_AB[other], g->d_errpart, g->nint, g->NFp, 1);
  };
  const int nbA = (g->nA + NT - 1) / NT, nbB = (int)(((size_t)g->nB * 32 + NT - 1) / NT);  // generic factors: one warp each
  auto extra_only = [&](auto tag) {
    constexpr int G = decltype(tag)::value;
    if (nbA) k_lin_extra<G, 0, NT><<<nbA, NT, 0, g->stream>>>(g->d_listA, g->nA, g->d_X, g->d_land, g->d_xkind, g->d_xsa, g->d_xsb, g->d_xl, g->d_xrow, g->d_xprm, g->d_XR[other], g->d_errpart + nb1, g->NX, g->NXRp, 1);
    if (nbB) k_lin_extra<G, 1, NT><<<nbB, NT, 0, g->stream>>>(g->d_listB, g->nB, g->d_X, g->d_land, g->d_xkind, g->d_xsa, g->d_xsb, g->d_xl, g->d_xrow, g->d_xprm, g->d_XR[other], g->d_errpart + nb1 + nbA, g->NX, g->NXRp, 1);
    if constexpr (G == G_POSE3) { if (g->nC) k_lin_extra<G, 2, NT><<<(g->nC + NT - 1) / NT, NT, 0, g->stream>>>(g->d_listC, g->nC, g->d_X, g->d_land, g->d_xkind, g->d_xsa, g->d_xsb, g->d_xl, g->d_xrow, g->d_xprm, g->d_XR[other], g->d_errpart + nb1 + nbA + nbB, g->NX, g->NXRp, 1); }
  };
  auto by_group = [&](auto fn) {
    switch (g->group) {
      case GPB_POSE3: fn(std::integral_constant<int, G_POSE3>{}); break;
      case GPB_POSE2: fn(std::integral_constant<int, G_POSE2>{}); break;
      case GPB_ROT3: fn(std::integral_constant<int, G_ROT3>{}); break;
      default: fn(std::integral_constant<int, G_LINEAR>{}); break;
    }
  };
  CUDA_TRY(cudaEventRecord(e0, g->stream));
  for (int r = 0; r < reps; r++) {
    switch (stage) {
      case 0: by_group(gp_only); break;
      case 1: by_group(extra_only); break;
      case 2: if ((rc = assemble_dispatch(g, g->cur))) return rc; break;
      case 3: if ((rc = solve_system(g, g->cur, 0.0))) return rc; break;
      case 4: if ((rc = retract_dispatch(g))) return rc; break;
      case 5: if ((rc = launch_fwd_level(g, g->cur, 0.0, 0))) return rc; break;
      case 6: if ((rc = launch_fwd_level(g, g->cur, 0.0, 0, 1))) return rc; break;  // spine only (SE(3), 64-column panel)
      case 7: if ((rc = launch_fwd_level(g, g->cur, 0.0, 0, 2))) return rc; break;  // panel only
      case 8: if ((rc = solve_backward(g))) return rc; break;
      case 9: if ((rc = linearize_dispatch(g, g->d_X, g->d_land, other, 1))) return rc; break;  // the whole linearise as the iteration runs it: GP priors and the other factors on their two streams, error reduction
      default: return fail(GPB_ERR_ARG, "gpb_time_stage: unknown stage");
    }
  }
  CUDA_TRY(cudaEventRecord(e1, g->stream));
  CUDA_TRY(cudaEventSynchronize(e1));
  CUDA_TRY(cudaGetLastError());
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  *ms_out = ms / reps;
  // stage 5 leaves the upper levels untouched but consistent; restore a complete solve so later calls see a valid state
  if (stage >= 5) { if ((rc = solve_system(g, g->cur, 0.0))) return rc; CUDA_TRY(cudaStreamSynchronize(g->stream)); }
  return GPB_OK;
}


// Single-factor compatibility path behind NoiseModelFactor::evaluateError: builds a throw-away 2-state graph with unit noise,
// runs the SAME batched kernels on it and un-whitens the result.  Thread-safe (no shared state); slow by design (allocations).
int gpb_eval_factor(int group, int kind, const double* x1, const double* v1, const double* x2, const double* v2, const double* landmark,
                    const double* prm, double* e_out, double* H_out, int* dims_out) {
  if (group < 0 || group > GPB_POSE3VW || !x1 || !prm || !e_out || !dims_out) return fail(GPB_ERR_ARG, "gpb_eval_factor: bad arguments");
  gpb_graph* g = gpb_graph_create(group, 3, 2, 1);
  if (!g) return GPB_ERR_ARG;
  const int D = g->D, PS = g->PS, DL = g->DL;
  auto done = [&](int rc) { gpb_graph_destroy(g); return rc; };
  std::vector<double> I(D * D, 0.0), I2(36, 0.0);
  for (int k = 0; k < D; k++) I[k + k * D] = 1.0;
  int rc = gpb_add_qc_model(g, I.data());
  if (rc < 0) return done(rc);
  const int zero = 0; const double one = 1.0;
  const double dt = prm[0], tau = prm[1];
  auto eye = [&](int m) { std::fill(I2.begin(), I2.end(), 0.0); for (int k = 0; k < m; k++) I2[k + k * m] = 1.0; return I2.data(); };
  switch (kind) {
    case 0: rc = gpb_add_gp_prior(g, 1, &zero, &dt, 0); break;
    case X_INTERP_RANGE: rc = gpb_add_interp_range(g, 1, &zero, &zero, &prm[2], &one, &dt, &tau, 0, prm[16] != 0.0 ? prm + 4 : nullptr); break;
    case X_INTERP_ATTITUDE: rc = gpb_add_interp_attitude(g, 1, &zero, &dt, &tau, 0, prm + 4, prm + 7, &one); break;
    case X_INTERP_GPS: rc = gpb_add_interp_gps(g, 1, &zero, prm + 40, eye(3), &dt, &tau, 0, prm[16] != 0.0 ? prm + 4 : nullptr); break;
    case X_INTERP_PROJECTION: rc = gpb_add_interp_projection(g, 1, &zero, &zero, prm + 40, eye(2), &dt, &tau, 0, prm + 43, prm[16] != 0.0 ? prm + 4 : nullptr); break;
    case X_PRIOR_POSE: rc = gpb_add_prior_pose(g, 0, prm + 4, eye(D)); break;
    case X_PRIOR_VEL: rc = gpb_add_prior_vel(g, 0, prm + 4, eye(D)); break;
    case X_PRIOR_LANDMARK: rc = gpb_add_prior_landmark(g, 0, prm + 4, eye(DL)); break;
    case X_BETWEEN: rc = gpb_add_between(g, 0, 1, prm + 4, eye(D)); break;
    case X_RANGE_2D: rc = gpb_add_range_2d(g, 0, 0, prm[2], 1.0); break;
    case X_RANGE_BEARING_2D: rc = gpb_add_range_bearing_2d(g, 0, 0, prm[2], prm[3], eye(2)); break;
    case X_ODOMETRY_2D: rc = gpb_add_odometry_2d(g, 0, 1, prm + 4, eye(3)); break;
    default: return done(fail(GPB_ERR_ARG, "gpb_eval_factor: unknown factor kind"));
  }
  if (rc < 0) return done(rc);
  std::vector<double> P(2 * PS, 0.0), V(2 * D, 0.0), Lm(std::max(DL, 1), 0.0);
  std::copy(x1, x1 + PS, P.begin());
  if (x2) std::copy(x2, x2 + PS, P.begin() + PS); else std::copy(x1, x1 + PS, P.begin() + PS);
  if (v1) std::copy(v1, v1 + D, V.begin());
  if (v2) std::copy(v2, v2 + D, V.begin() + D);
  if (landmark && DL) std::copy(landmark, landmark + DL, Lm.begin());
  if ((rc = gpb_set_values(g, P.data(), V.data(), Lm.data())) < 0) return done(rc);
  if ((rc = gpb_graph_finalize(g, 0)) < 0) return done(rc);
  if ((rc = gpb_linearize(g, nullptr)) < 0) return done(rc);
  double A[12 * 6 * 5], b[12];
  const int m = gpb_get_linearized_factor(g, kind == 0 ? 0 : 1, 0, A, b, dims_out);
  if (m < 0) return done(m);
  int ncols = 0;
  for (int v = 0; v < 5; v++) ncols += dims_out[v];
  if (kind == 0) {  // un-whiten the GP prior: rows [top; bot] = (U (x) I)^-1 [top'; bot']
    const GpWhiten w = gp_whiten(dt);
    auto unw = [&](double* col) { for (int k = 0; k < D; k++) { const double bot = col[D + k] / w.u22; col[D + k] = bot; col[k] = (col[k] - w.u12 * bot) / w.u11; } };
    for (int c = 0; c < ncols; c++) unw(A + (size_t)c * m);
    unw(b);
  }
  for (int r = 0; r < m; r++) e_out[r] = -b[r];
  if (H_out) std::copy(A, A + (size_t)m * ncols, H_out);
  return done(m);
}

}  // extern "C"

// ---- interpolatePose as a query (GaussianProcessInterpolator*::interpolatePose): slow-path helpers, plain allocations per call
template <int G> static void launch_interp_query(const double* X, const int* ia, const int* ib, const double* dt, const double* tau, int n, double* poses, double* H,
                                                 cudaStream_t stream) {
  k_interp_query<G><<<(n + 127) / 128, 128, 0, stream>>>(X, ia, ib, dt, tau, n, poses, H);
}
static int interp_query_device(int group, bool vw, const double* dX, const int* dia, const int* dib, const double* ddt, const double* dtau, int n, double* dposes, double* dH,
                               cudaStream_t stream) {
  if (group == GPB_POSE3 && vw) launch_interp_query<G_POSE3VW>(dX, dia, dib, ddt, dtau, n, dposes, dH, stream);
  else if (group == GPB_POSE3) launch_interp_query<G_POSE3>(dX, dia, dib, ddt, dtau, n, dposes, dH, stream);
  else if (group == GPB_POSE2) launch_interp_query<G_POSE2>(dX, dia, dib, ddt, dtau, n, dposes, dH, stream);
  else if (group == GPB_ROT3) launch_interp_query<G_ROT3>(dX, dia, dib, ddt, dtau, n, dposes, dH, stream);
  else launch_interp_query<G_LINEAR>(dX, dia, dib, ddt, dtau, n, dposes, dH, stream);
  CUDA_TRY(cudaGetLastError());
  return GPB_OK;
}
namespace {
struct DevBufs {  // frees whatever was allocated, on every exit path
  std::vector<void*> p;
  ~DevBufs() { for (void* q : p) cudaFree(q); }
  template <class T> int up(T** d, const T* h, size_t count) {
    void* q = nullptr;
    CUDA_TRY(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    p.push_back(q); *d = (T*)q;
    if (h && count) CUDA_TRY(cudaMemcpy(q, h, count * sizeof(T), cudaMemcpyHostToDevice));
    return GPB_OK;
  }
};
}  // namespace

extern "C" {

int gpb_interpolate_poses(int group, int device, int n, const double* x1, const double* v1, const double* x2, const double* v2, const double* delta_t, const double* tau,
                          double* poses_out, double* H_out) {
  if (group < 0 || group > GPB_POSE3VW || n < 1 || !x1 || !v1 || !x2 || !v2 || !delta_t || !tau || !poses_out) return fail(GPB_ERR_ARG, "gpb_interpolate_poses: bad arguments");
  for (int k = 0; k < n; k++) if (!(delta_t[k] > 0.0)) return fail(GPB_ERR_ARG, "gpb_interpolate_poses: delta_t must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(GPB_ERR_CUDA, "gpb_interpolate_poses: no CUDA device available (this engine has no CPU fallback)"); }
  CUDA_TRY(cudaSetDevice(device));
  const bool vw = group == GPB_POSE3VW;
  const int grp = vw ? GPB_POSE3 : group;
  const int D = grp == GPB_POSE3 ? 6 : 3, PS = pose_storage(grp, D), SR = PS + D;
  std::vector<double> X((size_t)2 * n * SR);
  std::vector<int> ia(n), ib(n);
  for (int k = 0; k < n; k++) {
    double* a = X.data() + (size_t)(2 * k) * SR; double* b = a + SR;
    std::copy(x1 + (size_t)k * PS, x1 + (size_t)(k + 1) * PS, a); std::copy(v1 + (size_t)k * D, v1 + (size_t)(k + 1) * D, a + PS);
    std::copy(x2 + (size_t)k * PS, x2 + (size_t)(k + 1) * PS, b); std::copy(v2 + (size_t)k * D, v2 + (size_t)(k + 1) * D, b + PS);
    ia[k] = 2 * k; ib[k] = 2 * k + 1;
  }
  DevBufs B; int rc;
  double *dX, *ddt, *dtau, *dP, *dH = nullptr; int *dia, *dib;
  if ((rc = B.up(&dX, X.data(), X.size())) || (rc = B.up(&dia, ia.data(), (size_t)n)) || (rc = B.up(&dib, ib.data(), (size_t)n)) || (rc = B.up(&ddt, delta_t, (size_t)n)) ||
      (rc = B.up(&dtau, tau, (size_t)n)) || (rc = B.up(&dP, (const double*)nullptr, (size_t)n * PS))) return rc;
  if (H_out && (rc = B.up(&dH, (const double*)nullptr, (size_t)n * 4 * D * D))) return rc;
  if ((rc = interp_query_device(grp, vw, dX, dia, dib, ddt, dtau, n, dP, dH, 0))) return rc;
  CUDA_TRY(cudaMemcpy(poses_out, dP, (size_t)n * PS * sizeof(double), cudaMemcpyDeviceToHost));
  if (H_out) CUDA_TRY(cudaMemcpy(H_out, dH, (size_t)n * 4 * D * D * sizeof(double), cudaMemcpyDeviceToHost));
  return GPB_OK;
}

int gpb_interpolate_velocities(int group, int device, int n, int dim, const double* x1, const double* v1, const double* x2, const double* v2, const double* delta_t, const double* tau,
                               double* vels_out, double* H_out) {
  if (n < 1 || !x1 || !v1 || !x2 || !v2 || !delta_t || !tau || !vels_out) return fail(GPB_ERR_ARG, "gpb_interpolate_velocities: bad arguments");
  // the reference implements interpolateVelocity for the Linear interpolator only; the Lie-group ones are declared and never defined
  // (gp/GaussianProcessInterpolatorPose3.h:118-123)
  if (group != GPB_LINEAR) return fail(GPB_ERR_UNSUPPORTED, "gpb_interpolate_velocities: GaussianProcessInterpolatorLinear only (the reference defines no other interpolateVelocity)");
  if (dim < 1 || dim > 16) return fail(GPB_ERR_ARG, "gpb_interpolate_velocities: dim out of range");
  for (int k = 0; k < n; k++) if (!(delta_t[k] > 0.0)) return fail(GPB_ERR_ARG, "gpb_interpolate_velocities: delta_t must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(GPB_ERR_CUDA, "gpb_interpolate_velocities: no CUDA device available (this engine has no CPU fallback)"); }
  CUDA_TRY(cudaSetDevice(device));
  DevBufs B; int rc;
  double *d1, *dv1, *d2, *dv2, *ddt, *dtau, *dV, *dH = nullptr;
  const size_t nd = (size_t)n * dim;
  if ((rc = B.up(&d1, x1, nd)) || (rc = B.up(&dv1, v1, nd)) || (rc = B.up(&d2, x2, nd)) || (rc = B.up(&dv2, v2, nd)) || (rc = B.up(&ddt, delta_t, (size_t)n)) ||
      (rc = B.up(&dtau, tau, (size_t)n)) || (rc = B.up(&dV, (const double*)nullptr, nd))) return rc;
  if (H_out && (rc = B.up(&dH, (const double*)nullptr, (size_t)n * 4))) return rc;
  k_interp_velocity_linear<<<(n + 127) / 128, 128>>>(d1, dv1, d2, dv2, ddt, dtau, n, dim, dV, dH);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(vels_out, dV, nd * sizeof(double), cudaMemcpyDeviceToHost));
  if (H_out) CUDA_TRY(cudaMemcpy(H_out, dH, (size_t)n * 4 * sizeof(double), cudaMemcpyDeviceToHost));
  return GPB_OK;
}

int gpb_graph_interpolate(gpb_graph* g, int n, const int* interval, const double* tau, double* poses_out) {
  CHECK_READY(g);
  if (n < 1 || !interval || !tau || !poses_out) return fail(GPB_ERR_ARG, "gpb_graph_interpolate: bad arguments");
  std::vector<int> ia(n), ib(n); std::vector<double> dts(n);
  for (int k = 0; k < n; k++) {
    if (interval[k] < 0 || interval[k] >= g->nint) return fail(GPB_ERR_ARG, "gpb_graph_interpolate: interval out of range");
    if (!(g->dt[interval[k]] > 0.0)) return fail(GPB_ERR_ARG, "gpb_graph_interpolate: no GP prior on that interval (its delta_t is unknown)");
    ia[k] = interval[k]; ib[k] = interval[k] + 1; dts[k] = g->dt[interval[k]];
  }
  DevBufs B; int rc;
  double *ddt, *dtau, *dP; int *dia, *dib;
  if ((rc = B.up(&dia, ia.data(), (size_t)n)) || (rc = B.up(&dib, ib.data(), (size_t)n)) || (rc = B.up(&ddt, dts.data(), (size_t)n)) || (rc = B.up(&dtau, tau, (size_t)n)) ||
      (rc = B.up(&dP, (const double*)nullptr, (size_t)n * g->PS))) return rc;
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  if ((rc = interp_query_device(g->group, g->vw != 0, g->d_X, dia, dib, ddt, dtau, n, dP, nullptr, g->stream))) return rc;
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  CUDA_TRY(cudaMemcpy(poses_out, dP, (size_t)n * g->PS * sizeof(double), cudaMemcpyDeviceToHost));
  return GPB_OK;
}

}  // extern "C"
