/* gpb.h — C ABI of the B200-native sparse-GP factor-graph linearise-and-solve engine.
 *
 * This is the drop-in boundary for gtrll/gpslam's hot path.  gpslam has no FFI
 * of its own — its plugin API is GTSAM's factor interface — so each entry point below names the
 * reference/GTSAM interface it stands in for.  Plain pointers and sizes only; all matrices are
 * column-major doubles (Eigen's layout); every call returns 0 on success or a negative status,
 * with gpb_last_error() giving the message.  No CPU fallback exists: without a CUDA device (or
 * with the CUDA library missing) gpb_graph_finalize() and everything after it fails loudly.
 *
 * Wire layouts (shared with the oracle and the kernels):
 *   Pose3 = 12 doubles [R column-major (9) | t (3)],  Rot3 = 9 (R column-major),
 *   Pose2 = (x, y, theta),  Linear<D> = D doubles;  velocities = tangent vectors (D doubles);
 *   Pose3 tangent order (omega, v) [gp/Pose3utils.cpp:116], Pose2 tangent (vx, vy, omega).
 *   Landmarks: Point3 for Pose3 graphs, Point2 for Pose2 / Linear<3> graphs.
 */
#ifndef GPB_H
#define GPB_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpb_graph gpb_graph;

/* state manifold of the trajectory; selects the factor family
 * (gp/GaussianProcessPrior{Pose3,Pose2,Rot3,Linear}.h and the matching interpolators) */
/* GPB_POSE3VW: the reference's Pose3 "VW" family (gp/GaussianProcessPriorPose3VW.h, gp/GaussianProcessInterpolatorPose3VW.h,
 * slam/GPInterpolatedGPSFactorPose3VW.h): states (Pose3 x, Vector3 v, Vector3 w) with v, w in the world frame.  Wire layout of a
 * velocity: [v(3) | w(3)]; tangent order of a state: [pose(6) | v(3) | w(3)].  On such a graph gpb_add_gp_prior adds
 * GaussianProcessPriorPose3VW(x_i, v_i, w_i, x_{i+1}, v_{i+1}, w_{i+1}, delta_t, Qc) and gpb_add_interp_gps adds
 * GPInterpolatedGPSFactorPose3VW; PriorFactor / BetweenFactor work as on GPB_POSE3 (gpb_add_prior_vel: one 6x6 sqrt information
 * over [v | w], i.e. the reference's two PriorFactor<Vector3> as a block-diagonal); range / projection factors have no VW
 * variant in the reference and are rejected. */
enum { GPB_POSE3 = 0, GPB_POSE2 = 1, GPB_ROT3 = 2, GPB_LINEAR = 3, GPB_POSE3VW = 4 };

enum { GPB_OK = 0, GPB_ERR_ARG = -1, GPB_ERR_CUDA = -2, GPB_ERR_STATE = -3, GPB_ERR_UNSUPPORTED = -4, GPB_ERR_NUMERIC = -5 };

/* LevenbergMarquardtParams / GaussNewtonParams (GTSAM defaults) */
typedef struct gpb_params {
  int max_iterations;        /* 100 */
  double rel_tol, abs_tol, err_tol; /* 1e-5, 1e-5, 0 */
  double lambda_initial, lambda_factor, lambda_upper, lambda_lower; /* 1e-5, 10, 1e5, 0 */
  double min_model_fidelity; /* 1e-3 */
  int use_lm;                /* 1 = LevenbergMarquardtOptimizer, 0 = GaussNewtonOptimizer */
} gpb_params;

typedef struct gpb_stats {
  int iterations;
  double error_initial, error_final, lambda;
  double linearize_ms, assemble_ms, solve_ms, update_ms, total_ms; /* device time (CUDA events) */
  int status;
} gpb_stats;

const char* gpb_last_error(void);
int gpb_device_count(void);
void gpb_default_params(gpb_params* p, int use_lm);

/* -- graph construction: stands in for NonlinearFactorGraph::add + Values::insert
 *    (matlab/PlazaPose2.m:90-204).  dim is used by GPB_LINEAR only (supported: 3). */
gpb_graph* gpb_graph_create(int group, int dim, int n_states, int n_landmarks);
/* the group the graph was created with (GPB_POSE3VW for a VW graph) */
int gpb_graph_group(const gpb_graph* g);
void gpb_graph_destroy(gpb_graph* g);

/* getQc (gp/GPutils.cpp:16-20): registers a D x D process-noise covariance Qc, returns its id */
int gpb_add_qc_model(gpb_graph* g, const double* Qc);

/* GaussianProcessPrior{Pose3,Pose2,Rot3,Linear}(x_i, v_i, x_{i+1}, v_{i+1}, delta_t, Qc)
 * (gp/GaussianProcessPriorPose3.h:43-49): n factors on intervals i[k] -> i[k]+1 */
int gpb_add_gp_prior(gpb_graph* g, int n, const int* i, const double* delta_t, int qc);

/* GPInterpolatedRangeFactor{Pose3,Pose2,2DLinear}(z, meas_model(sigma), Qc, x_i, v_i, x_{i+1}, v_{i+1}, l, delta_t, tau
 * [, body_P_sensor])  (slam/GPInterpolatedRangeFactorPose3.h:46-54).  body_P_sensor: one pose shared by the n factors, or NULL. */
int gpb_add_interp_range(gpb_graph* g, int n, const int* i, const int* l, const double* z, const double* sigma,
                         const double* delta_t, const double* tau, int qc, const double* body_P_sensor);

/* GPInterpolatedAttitudeFactorRot3(x_i, v_i, x_{i+1}, v_{i+1}, delta_t, tau, Qc, meas_model(sigma), nZ, bRef)
 * (slam/GPInterpolatedAttitudeFactorRot3.h:44-51).  nZ, bRef: 3 doubles per factor. */
int gpb_add_interp_attitude(gpb_graph* g, int n, const int* i, const double* delta_t, const double* tau, int qc,
                            const double* nZ, const double* bRef, const double* sigma);

/* GPInterpolatedGPSFactorPose3(measured_point3, meas_model, Qc, x_i, v_i, x_{i+1}, v_{i+1}, delta_t, tau [, body_P_sensor])
 * (slam/GPInterpolatedGPSFactorPose3.h:47-56).  measured: 3 doubles per factor; sqrt_info: upper-triangular 3 x 3 R of the
 * measurement model, shared by the n factors; body_P_sensor: one pose or NULL.  Pose3 trajectories only. */
int gpb_add_interp_gps(gpb_graph* g, int n, const int* i, const double* measured, const double* sqrt_info, const double* delta_t, const double* tau, int qc,
                       const double* body_P_sensor);

/* GPInterpolatedProjectionFactorPose3<Cal3_S2>(measured_point2, cam_model, Qc, x_i, v_i, x_{i+1}, v_{i+1}, l, delta_t, tau, K
 * [, body_P_sensor])  (slam/GPInterpolatedProjectionFactorPose3.h:60-75).  measured: 2 doubles per factor; sqrt_info: 2 x 2 R;
 * K = (fx, fy, s, u0, v0).  A landmark behind the camera follows the reference's no-throw path (:123-138): zero Jacobians and
 * the residual (2 fx, 2 fx); throwCheirality is not offered. */
int gpb_add_interp_projection(gpb_graph* g, int n, const int* i, const int* l, const double* measured, const double* sqrt_info, const double* delta_t,
                              const double* tau, int qc, const double* K, const double* body_P_sensor);

/* gtsam::PriorFactor<Pose|Vector|Point>: value in wire layout, sqrt_info = upper-triangular R (d x d) */
int gpb_add_prior_pose(gpb_graph* g, int i, const double* value, const double* sqrt_info);
int gpb_add_prior_vel(gpb_graph* g, int i, const double* value, const double* sqrt_info);
int gpb_add_prior_landmark(gpb_graph* g, int l, const double* value, const double* sqrt_info);

/* gtsam::BetweenFactor<Pose>(x_i, x_j, measured, model): odometry when |i-j| == 1; otherwise a loop closure - both states
 * become pinned separators of the elimination and join the dense reduced system (single-GPU graphs) */
int gpb_add_between(gpb_graph* g, int i, int j, const double* measured, const double* sqrt_info);

/* plain 2-way factors of gpslam/slam (Linear<3> states unless noted):
 * RangeFactor2DLinear (slam/RangeFactor2DLinear.h:43-56) / RangeFactorPose2 (slam/RangeFactorPose2.h:15, Pose2 graphs),
 * RangeBearingFactor2DLinear (slam/RangeBearingFactor2DLinear.h:47-84), OdometryFactor2DLinear (slam/OdometryFactor2DLinear.h:50-75) */
int gpb_add_range_2d(gpb_graph* g, int i, int l, double z, double sigma);
int gpb_add_range_bearing_2d(gpb_graph* g, int i, int l, double range, double bearing, const double* sqrt_info);
int gpb_add_odometry_2d(gpb_graph* g, int i, int j, const double* measured, const double* sqrt_info);

/* Single-factor compatibility path: NoiseModelFactorN::evaluateError(x..., H...) of ONE factor (e.g.
 * gp/GaussianProcessPriorPose3.h:60-65, slam/GPInterpolatedRangeFactorPose3.h:64-69): unwhitened residual e (m doubles) and,
 * when H_out != NULL, the Jacobian blocks concatenated column-major in the factor's variable order (dims_out[5] = their widths).
 * Evaluated by the same device code as the batched path on a throw-away 2-state graph; thread-safe, not fast.
 * kind: GPB_F_*.  prm[48]: [0] delta_t [1] tau [2] range / z [3] bearing [4..15] body_P_sensor | nZ(3),bRef(3) | prior value |
 * measured (wire layout) [16] has_sensor [40..42] GPS point / image point [43..47] Cal3_S2 (fx, fy, s, u0, v0).
 * x2/v2/landmark may be NULL when the factor does not use them.  Returns m or a status. */
enum { GPB_F_GP_PRIOR = 0, GPB_F_INTERP_RANGE = 1, GPB_F_INTERP_ATTITUDE = 2, GPB_F_PRIOR_POSE = 3, GPB_F_PRIOR_VEL = 4, GPB_F_PRIOR_LANDMARK = 5,
       GPB_F_BETWEEN = 6, GPB_F_RANGE_2D = 7, GPB_F_RANGE_BEARING_2D = 8, GPB_F_ODOMETRY_2D = 9, GPB_F_INTERP_GPS = 10, GPB_F_INTERP_PROJECTION = 11 };
int gpb_eval_factor(int group, int kind, const double* x1, const double* v1, const double* x2, const double* v2, const double* landmark,
                    const double* prm, double* e_out, double* H_out, int* dims_out);

/* GaussianProcessInterpolator{Pose3,Pose3VW,Pose2,Rot3,Linear}(Qc, delta_t, tau).interpolatePose(pose1, vel1, pose2, vel2, H1..H4)
 * (gp/GaussianProcessInterpolatorPose3.h:57-105, ...Pose3VW.h:58-108, ...Pose2.h:56-89, ...Rot3.h:56-86, ...Linear.h:70-90) for n
 * independent queries.  Host buffers in the wire layouts of gpb_set_values; poses_out [n x pose_storage]; H_out (or NULL)
 * [n][4][D x D column-major] = Hint1..Hint4 (GPB_POSE3VW: Hint2 = [H2 | H3] over [v | w], Hint4 = [H5 | H6]).  Lambda and Psi do not
 * depend on Qc, so no Qc is passed.  tau may lie outside [0, delta_t].  Slow path (allocations per call). */
int gpb_interpolate_poses(int group, int device, int n, const double* x1, const double* v1, const double* x2, const double* v2, const double* delta_t, const double* tau,
                          double* poses_out, double* H_out);
/* GaussianProcessInterpolatorLinear<dim>(Qc, delta_t, tau).interpolateVelocity(pose1, vel1, pose2, vel2, H1..H4)
 * (gp/GaussianProcessInterpolatorLinear.h:106-126) for n queries: vels_out [n x dim]; H_out (or NULL) [n][4] - H1..H4 are scalar
 * multiples of the identity (Lambda21, Lambda22, Psi21, Psi22), returned as those four scalars.  group must be GPB_LINEAR: the
 * reference declares interpolateVelocity for its Lie-group interpolators but never defines it. */
int gpb_interpolate_velocities(int group, int device, int n, int dim, const double* x1, const double* v1, const double* x2, const double* v2,
                               const double* delta_t, const double* tau, double* vels_out, double* H_out);
/* The pose of the graph's CURRENT estimate at time tau[k] into interval[k] (between states interval[k] and interval[k] + 1, which
 * must carry a GP prior: its delta_t is used) - dense trajectory output after gpb_optimize, what the reference's scripts do with
 * interpolatePose on the optimised Values.  poses_out [n x pose_storage]. */
int gpb_graph_interpolate(gpb_graph* g, int n, const int* interval, const double* tau, double* poses_out);

/* Values::insert / Values::at : host buffers, [n_states x pose_storage], [n_states x D], [n_landmarks x DL] */
int gpb_set_values(gpb_graph* g, const double* poses, const double* vels, const double* landmarks);
int gpb_get_values(gpb_graph* g, double* poses, double* vels, double* landmarks);
/* page-locked host memory for the value arrays: gpb_set_values / gpb_get_values then move them at PCIe rate (any host pointer
 * is accepted; pageable memory goes through the driver's staging copies) */
int gpb_alloc_host(void** ptr, long long bytes);
int gpb_free_host(void* ptr);

/* Freezes the graph: sorts factors by interval, builds the device-resident SoA layout, allocates
 * every solver buffer on `device`.  Must be called once after the last gpb_add_*. */
int gpb_graph_finalize(gpb_graph* g, int device);

/* NonlinearFactorGraph::error(values): 0.5 * sum |R e|^2 at the current values */
int gpb_error(gpb_graph* g, double* error_out);

/* NonlinearFactorGraph::linearize(values): batched evaluation of every factor's whitened
 * [A|b] (the GaussianFactorGraph payload) into device memory; returns the graph error. */
int gpb_linearize(gpb_graph* g, double* error_out);

/* Copy-out of one factor's whitened JacobianFactor after gpb_linearize (parity checks):
 * kind 0 = GP prior of interval idx, kind 1 = idx-th measurement/prior/between factor in insertion order.
 * A_out: concatenated column-major blocks (m x d_v) in the factor's variable order; b_out: m.  Returns m (>0) or a status. */
int gpb_get_linearized_factor(gpb_graph* g, int kind, int idx, double* A_out, double* b_out, int* dims_out /*[5]*/);

/* Dense copy-out of the assembled normal equations H = A^T A, rhs = A^T b in the variable order
 * [x_0 v_0 x_1 v_1 ... | landmarks] (small graphs only, parity checks).  n must equal the system dimension. */
int gpb_get_normal_equations(gpb_graph* g, double* H_out, double* rhs_out, int n);

/* GaussNewtonOptimizer::iterate / LevenbergMarquardtOptimizer::iterate (matlab/PlazaPose2.m:225):
 * n_iter > 0 runs exactly n_iter iterations; n_iter <= 0 runs NonlinearOptimizer::optimize() to convergence. */
int gpb_optimize(gpb_graph* g, const gpb_params* params, int n_iter, gpb_stats* stats);

/* Solve the current linearisation once: delta = argmin |A d - b|^2 + lambda |d|^2, copied to the host
 * as [n_states x 2D] and [n_landmarks x DL] (parity checks of the block solver). */
int gpb_solve_delta(gpb_graph* g, double lambda, double* delta_states, double* delta_landmarks);

/* -- trajectory sharding across the GPUs of one node.  Each process owns one contiguous segment of the chain
 * and builds only that sub-graph (local state indices).  ext_left: local state 0 is the last state of the left neighbour
 * (halo copy; a separator of the global reduced system); ext_right: the local last state is this rank's boundary separator.
 * Call before gpb_graph_finalize.  Landmarks are replicated; factors attached to interval (t, t+1) belong to the rank
 * whose local chain contains both states. */
int gpb_graph_set_shard(gpb_graph* g, int rank, int world, int ext_left, int ext_right);

/* Sharded graphs WITH loop closures.  A closure (i, j) is evaluated by the rank that owns state min(i, j); a remote endpoint is
 * carried as a GHOST: an extra chain entry after the shard's own states (create the graph with n_real + n_ghost states; no
 * factor but the closures touches a ghost; its values are set like any state's and must equal the owner's).  Every rank applies
 * the same reduced-system solution to its ghosts, so replicas stay bit-identical without any exchange of values.
 * pinned_local (ascending): every local chain entry that belongs to the global reduced system - the halo (0) and the shard's
 * last own state when they are external separators, every loop-closure endpoint owned by this shard (whether or not the
 * closure is evaluated here), every ghost; pinned_gtop: their indices in the global list of top states (ntop_global long:
 * shard boundaries and closure endpoints in trajectory order). */
int gpb_graph_set_top_map(gpb_graph* g, int n_real, int ntop_global, int n_pinned, const int* pinned_local, const int* pinned_gtop);

/* In-place SUM all-reduce of `count` doubles at device pointer `buf` over all ranks, STREAM-ORDERED on `cuda_stream` (the
 * engine's cudaStream_t): it must consume the buffer after the work already enqueued on that stream and produce the sums
 * before work enqueued on it later; it need not be complete on return (NCCL enqueued on that stream is the intended
 * implementation - the engine never synchronises the host around the call, so a plain Gauss-Newton run stays asynchronous).
 * An implementation that works on the host instead calls gpb_stream_synchronize(cuda_stream) first and returns when done.
 * The engine calls it exactly once per Gauss-Newton iteration, on the packed boundary Schur-complement system (which also
 * carries the error scalar). */
typedef int (*gpb_allreduce_fn)(void* ctx, double* device_buf, long long count, void* cuda_stream);
int gpb_set_allreduce(gpb_graph* g, gpb_allreduce_fn fn, void* ctx);

/* The engine's own all-reduce: an NCCL communicator over the ranks of the sharded graph, created from a ncclUniqueId (128 bytes)
 * that rank 0 obtains with gpb_nccl_unique_id and hands to the other ranks by whatever means the host program has (the Python
 * host uses one torch.distributed broadcast).  libnccl.so.2 is bound at run time (the copy the process already holds, else the
 * system one; GPB_NCCL_LIB overrides).  With it the boundary all-reduce is enqueued by the engine on its own stream and a whole
 * Gauss-Newton iteration - collective included - is replayed as ONE CUDA graph.  Collective call: every rank, after
 * gpb_graph_finalize.  Takes precedence over gpb_set_allreduce. */
int gpb_nccl_unique_id(unsigned char* id128_out);
int gpb_graph_init_nccl(gpb_graph* g, const unsigned char* id128, int rank, int world);

/* Pipelined batch of K independent steps on the resident graph (same factors, new values each step), each step =
 * Values::insert (host -> device) + GaussNewtonOptimizer::iterate() once + Values::at (device -> host), the usage of
 * matlab/PlazaPose2.m:224-233 with values crossing the boundary every step.  Step k+1's host->device copy and step k-1's
 * device->host copy run on their own streams while step k computes.  poses_in[k] / vels_in[k] / land_in[k] and the *_out[k]
 * are host buffers in the layouts of gpb_set_values (page-locked ones from gpb_alloc_host for full overlap; an out buffer must
 * not alias an in buffer of a later step).  errors_out (or NULL) [K]: this rank's graph error after each step.  Gauss-Newton only. */
int gpb_optimize_batch(gpb_graph* g, int K, const double* const* poses_in, const double* const* vels_in, const double* const* land_in,
                       double* const* poses_out, double* const* vels_out, double* const* land_out, double* errors_out, gpb_stats* stats);

/* solver tuning: segment length per elimination level (>= 2); 0 keeps the default */
int gpb_set_segment_length(gpb_graph* g, int level0, int upper_levels);

/* bytes of HBM held by the graph, and the algorithmic byte counts of SURVEY.md §8(d) */
typedef struct gpb_sizes { double hbm_bytes, linearise_bytes, fused_bytes, solve_bytes; int n_gp, n_extra, n_rows, border_dim, levels; } gpb_sizes;
int gpb_get_sizes(gpb_graph* g, gpb_sizes* s);

/* profiling aid: average device milliseconds (CUDA events on the engine's stream) of one stage of the hot path over `reps`
 * launches at the current values.  stage: 0 batched GP-prior linearise kernel, 1 linearise of the other factors, 2 assembly,
 * 3 whole block solve (all levels, both sweeps), 4 retract, 5 level-0 forward elimination only, 6 / 7 its spine / panel kernel
 * alone (SE(3) graphs with a 64-column panel), 8 back-substitution (all levels), 9 the whole linearise as an iteration runs it
 * (GP priors and the other factors on their two streams + the error reduction). */
int gpb_time_stage(gpb_graph* g, int stage, int reps, double* ms_out);

/* names and average device milliseconds of the kernels timed during the last gpb_optimize (profiling aid) */
int gpb_kernel_launches_last_optimize(gpb_graph* g);
/* plain cudaMemcpy (kind 1: host->device, 2: device->host), for gpb_allreduce_fn implementations without their own CUDA binding */
int gpb_memcpy(void* dst, const void* src, long long bytes, int kind);
/* cudaStreamSynchronize for the same callers */
int gpb_stream_synchronize(void* cuda_stream);
/* profiling aid: measured FP64 tensor-pipe (mma.sync m8n8k4, SASS DMMA) peak of the device in TFLOP/s - the roofline
 * denominator of the panel kernel */
int gpb_debug_dmma_peak(int device, double* tflops_out);
/* profiling aid: dependent-issue latency, in SM clocks per operation, of the instructions the latency-bound solver kernels chain:
 * out[0] DFMA, [1] DMUL, [2] DMMA m8n8k4 (same accumulator), [3] 64-bit shuffle, [4] shared-memory load, [5] rsqrt (MUFU seed +
 * third-order correction + 1 add), [6] DADD, [7] global load served by L2 */
int gpb_debug_latency(int device, double* out8);
/* profiling aid: microseconds to write n_factors SE(3) [A|b] records (2400 B each) in k_lin_gp's store pattern with no arithmetic
 * in front (mode 4: the tiled layout the engine uses; 1: untiled SoA, row pairs NFp*16 B apart; 2: 1 with streaming stores; 3: 1 with
 * 256-thread CTAs) or with cudaMemsetAsync (mode 0) - the floor of the layout */
int gpb_debug_store_peak(int device, int mode, int n_factors, double* us_out);
/* testing aid: the reduced-system solver alone - (A + lambda * diag[loff..R)) x = b, A symmetric R x R column-major; the
 * shared-memory single-CTA solver up to R = 160 (blocked factorisation in 8-column steps with tensor-pipe trailing updates;
 * force_blocked == 2 / 3: the per-column kernels with plain loops / a register-blocked trailing update), the blocked multi-CTA
 * Cholesky beyond (or when force_blocked == 1) */
int gpb_debug_dense_solve(int device, int R, const double* A, const double* b, double lambda, int loff, int force_blocked, double* x_out);
/* all-reduce calls issued by the last gpb_optimize (sharded graphs) */
int gpb_allreduces_last_optimize(gpb_graph* g);

#ifdef __cplusplus
}
#endif
#endif /* GPB_H */
