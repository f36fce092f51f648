// Type-checks include/gpslam_b200/gtsam_adapter.h against the API stubs of tests/cpp/gtsam_stub (GTSAM itself is not available in
// the build container).  It builds a small Pose2 graph the way matlab/PlazaPose2.m does and calls optimizeOnB200; without a GPU the
// call must fail loudly at gpb_graph_finalize (exit code 2), which also shows the whole lowering ran.
#include <cmath>
#include <cstdio>

#include "gpslam_b200/gtsam_adapter.h"

#ifndef GPSLAM_B200_HAVE_GTSAM
#error "the adapter's __has_include guard did not see the stub headers"
#endif

using namespace gtsam;

static SharedNoiseModel gaussian(int n, double sigma) {
  auto g = std::make_shared<noiseModel::Gaussian>();
  g->R_ = Matrix(n, n); g->cov_ = Matrix(n, n);
  for (int k = 0; k < n; k++) { g->R_(k, k) = 1.0 / sigma; g->cov_(k, k) = sigma * sigma; }
  return g;
}
static SharedNoiseModel gpNoise(int D, double qc, double dt) {  // Gaussian::Covariance(calcQ(Qc, dt)), gp/GPutils.h:24-30
  auto g = std::make_shared<noiseModel::Gaussian>();
  g->cov_ = Matrix(2 * D, 2 * D); g->R_ = Matrix(2 * D, 2 * D);
  for (int k = 0; k < D; k++) { g->cov_(k, k) = dt * dt * dt / 3 * qc; g->cov_(k, D + k) = g->cov_(D + k, k) = dt * dt / 2 * qc; g->cov_(D + k, D + k) = dt * qc; }
  return g;
}

int main() {
  NonlinearFactorGraph graph;
  Values init;
  const double dt = 0.1;
  for (int i = 1; i <= 3; i++) { init.insert(Symbol('x', i), Pose2(0.1 * i, 0, 0)); init.insert(Symbol('v', i), Vector3()); }
  init.insert(Symbol('l', 0), Point2(2, 3));
  graph.push_back(std::make_shared<PriorFactor<Pose2>>(Symbol('x', 1), Pose2(0, 0, 0), gaussian(3, 1.0)));
  graph.push_back(std::make_shared<PriorFactor<Point2>>(Symbol('l', 0), Point2(2, 3), gaussian(2, 1.0)));
  for (int i = 1; i < 3; i++) {
    graph.push_back(std::make_shared<gpslam::GaussianProcessPriorPose2>(Symbol('x', i), Symbol('v', i), Symbol('x', i + 1), Symbol('v', i + 1), dt, gpNoise(3, 0.01, dt)));
    graph.push_back(std::make_shared<BetweenFactor<Pose2>>(Symbol('x', i), Symbol('x', i + 1), Pose2(0.1, 0, 0), gaussian(3, 1e-3)));
    graph.push_back(std::make_shared<gpslam::GPInterpolatedRangeFactorPose2>(3.5, gaussian(1, 0.5), gpNoise(3, 0.01, dt), Symbol('x', i), Symbol('v', i), Symbol('x', i + 1),
                                                                             Symbol('v', i + 1), Symbol('l', 0), dt, 0.05));
  }
  // the prior's (delta_t, Qc) come back out of its noise model
  std::vector<double> Qc;
  const double got = gpslam_b200::adapter::priorDeltaT(gpNoise(3, 0.01, dt), 3, Qc);
  if (std::fabs(got - dt) > 1e-12 || std::fabs(Qc[0] - 0.01) > 1e-12 || std::fabs(Qc[4] - 0.01) > 1e-12 || Qc[1] != 0.0) { std::printf("priorDeltaT wrong: %g %g\n", got, Qc[0]); return 1; }
  try {
    gpb_stats st;
    const Values result = gpslam_b200::optimizeOnB200(graph, init, true, 0, &st);
    std::printf("optimised: %d iterations, error %g, x3 = (%g, %g, %g)\n", st.iterations, st.error_final, result.at<Pose2>(Symbol('x', 3)).x(), result.at<Pose2>(Symbol('x', 3)).y(),
                result.at<Pose2>(Symbol('x', 3)).theta());
  } catch (const std::exception& e) { std::printf("exception: %s\n", e.what()); return 2; }
  return 0;
}
