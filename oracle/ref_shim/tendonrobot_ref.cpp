// TendonRobot::shape / home_shape / is_valid of the reference, compiled from its own text
// (test infrastructure only).
//
// tendon/TendonRobot.cpp as a whole needs Boost.odeint, levmar, cpptoml, csv, spline and FCL.  This
// translation unit compiles, unmodified:
//   tendon/TendonRobot.h, TendonResult.h, TendonSpecs.h, BackboneSpecs.h, get_r_info.h ...   as they are
//   tendon/{get_r_info,tendon_deriv,solve_initial_bending}.cpp, collision/collision_primitives.cpp   as they are
// and these runs of definitions, cut out of the reference by anchors at build time (oracle/Makefile,
// _ref/gen/, deleted after the build):
//   tr_A.inc   TendonRobot.cpp from `namespace E = Eigen;` up to (not including) tension_shape_unopt:
//              t_range, get_stiffness_matrices, simpsons, PointForces::calc_point_forces, random_state,
//              home_shape, **tension_shape**                                   (TendonRobot.cpp:56-500)
//   tr_B.inc   TendonRobot::is_valid, TendonRobot::collides_self               (:955-974)
//   tr_rot.inc TendonResult::rotate_z                                          (TendonResult.cpp:13-18)
//   tr_deg.inc poly_degree, TendonSpecs::r_degree / theta_degree               (TendonSpecs.cpp:8-30)
//   tr_jac.inc tip_control::Jacobian                                          (tip-control/tip_control.cpp:243-265)
//   sc_*.inc   Sphere / Capsule / CapsuleSequence, the inline collides overloads, collides_self
//              (as for libselfcol_ref.so)
// against TWO stand-ins: Eigen (eigen_standin/) and Boost.odeint (odeint_standin/): read their headers for
// what that does and does not pin.  AbstractValidityChecker::is_valid_shape (AbstractValidityChecker.cpp:
// 99-114, an OMPL class) is represented by its three ingredients evaluated separately in trref_flags.
#include <tendon/TendonRobot.h>
#include <tendon/get_r_info.h>
#include <tendon/solve_initial_bending.h>
#include <tendon/tendon_deriv.h>
#include <collision/collision_primitives.h>
#include <util/macros.h>
#include <util/poly.h>
#include <util/vector_ops.h>

#include <Eigen/Core>
#include <Eigen/Dense>
#include <boost/numeric/odeint.hpp>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <iterator>
#include <random>
#include <stdexcept>
#include <tuple>
#include <vector>

namespace collision {
#include "sc_sphere.inc"
#include "sc_capsule.inc"
#include "sc_capseq.inc"
#include "sc_collides.inc"
#include "sc_self.inc"
}  // namespace collision

#include "tr_A.inc"   // opens namespace tendon { and leaves it open
#include "tr_B.inc"
#include "tr_rot.inc"
#include "tr_deg.inc"
}  // namespace tendon

// tip_control::Jacobian (tip-control/tip_control.cpp:243-265), cut out by anchors (tr_jac.inc): the
// forward-difference tip Jacobian of the resolved-rate controllers -- note its `float dist`
namespace tip_control {
namespace E = Eigen;
#include "tr_jac.inc"
}  // namespace tip_control

namespace {

tendon::TendonRobot make_robot(const double *hdr, int N, int Nc, int Nd, const double *C, const double *D,
                               const double *lim /* [N][3]: max_tension, min_length, max_length */) {
  tendon::TendonRobot rb;
  rb.r = hdr[0];
  rb.specs.L = hdr[1]; rb.specs.dL = hdr[2]; rb.specs.ro = hdr[3]; rb.specs.ri = hdr[4];
  rb.specs.E = hdr[5]; rb.specs.nu = hdr[6];
  rb.residual_threshold = hdr[7];
  rb.enable_rotation = hdr[8] != 0.0;
  rb.enable_retraction = hdr[9] != 0.0;
  rb.tendons.resize((size_t)N);
  for (int j = 0; j < N; j++) {
    rb.tendons[j].C.resize(Nc);
    rb.tendons[j].D.resize(Nd);
    for (int i = 0; i < Nc; i++) rb.tendons[j].C[i] = C[j * Nc + i];
    for (int i = 0; i < Nd; i++) rb.tendons[j].D[i] = D[j * Nd + i];
    rb.tendons[j].max_tension = lim[3 * j];
    rb.tendons[j].min_length = lim[3 * j + 1];
    rb.tendons[j].max_length = lim[3 * j + 2];
  }
  return rb;
}

}  // namespace

extern "C" {

// TendonRobot::shape(state).  Returns the number of points (or -1 if it exceeds cap).  misc: L, then
// u_i, u_f, v_i, v_f (3 each), then converged (0/1); L_i: [N]; R column-major 9 per point.
int trref_shape(const double *hdr, int N, int Nc, int Nd, const double *C, const double *D,
                const double *lim, const double *state, int cap, double *t, double *p, double *R,
                double *L_i, double *misc) {
  tendon::TendonRobot rb = make_robot(hdr, N, Nc, Nd, C, D, lim);
  std::vector<double> st(state, state + rb.state_size());
  tendon::TendonResult res = rb.shape(st);
  const int n = (int)res.t.size();
  if (n > cap) return -1;
  for (int i = 0; i < n; i++) {
    t[i] = res.t[i];
    for (int k = 0; k < 3; k++) p[3 * i + k] = res.p[i][k];
    for (int k = 0; k < 9; k++) R[9 * i + k] = res.R[i].data()[k];
  }
  for (int j = 0; j < N; j++) L_i[j] = res.L_i[j];
  misc[0] = res.L;
  for (int k = 0; k < 3; k++) {
    misc[1 + k] = res.u_i[k]; misc[4 + k] = res.u_f[k]; misc[7 + k] = res.v_i[k]; misc[10 + k] = res.v_f[k];
  }
  misc[13] = res.converged ? 1.0 : 0.0;
  return n;
}

// TendonRobot::shape over a batch, the OpenMP loop of apps/estimate_length_discretization.cpp:62-71 (one
// robot, `#pragma omp parallel for` over configurations, every TendonResult kept until the loop ends like
// the app keeps them).  Timing baseline of bench.py (--impl reference / cpu_baseline); tips[n][3] and
// npts[n] are written so the work cannot be optimised away and so the caller can spot-check the results.
void trref_shape_batch(const double *hdr, int N, int Nc, int Nd, const double *C, const double *D,
                       const double *lim, const double *states, long long n, int nthreads, double *tips,
                       int *npts) {
  const tendon::TendonRobot rb = make_robot(hdr, N, Nc, Nd, C, D, lim);
  const size_t S = rb.state_size();
  std::vector<tendon::TendonResult> results((size_t)n);
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (long long i = 0; i < n; i++) {
    std::vector<double> st(states + i * S, states + (i + 1) * S);
    results[(size_t)i] = rb.shape(st);
  }
  for (long long i = 0; i < n; i++) {
    const auto &res = results[(size_t)i];
    npts[i] = (int)res.p.size();
    for (int k = 0; k < 3; k++) tips[3 * i + k] = res.p.empty() ? 0.0 : res.p.back()[k];
  }
}

// tip_control::Jacobian(robot, ps, dist, tau): J is written row-major 3 x S (J[j * S + i] = J(j, i))
void trref_tip_jacobian(const double *hdr, int N, int Nc, int Nd, const double *C, const double *D,
                        const double *lim, const double *state, const double *ps, float dist, double *J) {
  tendon::TendonRobot rb = make_robot(hdr, N, Nc, Nd, C, D, lim);
  const size_t S = rb.state_size();
  std::vector<double> st(state, state + S);
  Eigen::MatrixXd Jm = tip_control::Jacobian(rb, Eigen::Vector3d(ps[0], ps[1], ps[2]), dist, st);
  for (size_t j = 0; j < 3; j++)
    for (size_t i = 0; i < S; i++) J[j * S + i] = Jm(j, i);
}

// home_shape(state).L_i
void trref_home_lengths(const double *hdr, int N, int Nc, int Nd, const double *C, const double *D,
                        const double *lim, const double *state, double *L_i) {
  tendon::TendonRobot rb = make_robot(hdr, N, Nc, Nd, C, D, lim);
  std::vector<double> st(state, state + rb.state_size());
  tendon::TendonResult h = rb.home_shape(st);
  for (int j = 0; j < N; j++) L_i[j] = h.L_i[j];
}

// the three ingredients of AbstractValidityChecker::is_valid_shape (AbstractValidityChecker.cpp:99-114),
// each from the reference's own code: bit 0 !converged, bit 1 !is_within_length_limits(calc_dl(...)),
// bit 2 collides_self
unsigned trref_flags(const double *hdr, int N, int Nc, int Nd, const double *C, const double *D,
                     const double *lim, const double *state) {
  tendon::TendonRobot rb = make_robot(hdr, N, Nc, Nd, C, D, lim);
  std::vector<double> st(state, state + rb.state_size());
  tendon::TendonResult fk = rb.shape(st), home = rb.home_shape(st);
  unsigned f = 0;
  if (!fk.converged || !home.converged) f |= 1u;
  if (!rb.is_within_length_limits(rb.calc_dl(home.L_i, fk.L_i))) f |= 2u;
  if (rb.collides_self(fk)) f |= 4u;
  return f;
}

}  // extern "C"
