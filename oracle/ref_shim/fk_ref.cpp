// C wrapper around four UNMODIFIED reference source files (test infrastructure only):
//     tendon/get_r_info.cpp            get_r_info2                      
//     tendon/tendon_deriv.cpp          tendon_deriv, tendon_deriv_unopt (row 3)
//     tendon/solve_initial_bending.cpp solve_initial_bending            (row 2)
//     collision/collision_primitives.{h,cpp}  closest_st_segment, segment_aabox_intersect (rows 6, 8)
// compiled where they lie under /root/reference/cpp/src against the Eigen STAND-IN in
// eigen_standin/Eigen/Core (read its header: it pins the reference's expression structure, not
// Eigen's rounding).  oracle/Makefile links this file with them into oracle/_ref/libfk_ref.so.
//
// What is NOT the reference here (and is marked so below): the stiffness matrices (restated from
// tendon/TendonRobot.cpp:105-148, which cannot be compiled: Boost.odeint, levmar, cpptoml, FCL) and
// the classic-RK4 driver fkref_shape (tendon/TendonRobot.cpp:458-462 calls odeint, not vendored).
#include <tendon/TendonSpecs.h>
#include <tendon/get_r_info.h>
#include <tendon/solve_initial_bending.h>
#include <tendon/tendon_deriv.h>
#include <collision/collision_primitives.h>

#include <cmath>
#include <cstdint>
#include <tuple>
#include <vector>

namespace {

using V3 = Eigen::Vector3d;
using M3 = Eigen::Matrix3d;

std::vector<tendon::TendonSpecs> make_tendons(int N, int Nc, int Nd, const double *C, const double *D) {
  std::vector<tendon::TendonSpecs> t(static_cast<size_t>(N));
  for (int j = 0; j < N; j++) {
    t[j].C.resize(Nc);
    t[j].D.resize(Nd);
    for (int i = 0; i < Nc; i++) t[j].C[i] = C[j * Nc + i];
    for (int i = 0; i < Nd; i++) t[j].D[i] = D[j * Nd + i];
  }
  return t;
}

// restated: tendon/TendonRobot.cpp:105-148 (get_stiffness_matrices)
struct Stiff { M3 K_bt, K_se, K_bt_inv, K_se_inv; };
Stiff stiffness(double ro, double ri, double E, double nu) {
  double ro2 = ro * ro, ri2 = ri * ri;
  double I = (1.0 / 4.0) * M_PI * (ro2 * ro2 - ri2 * ri2);
  double Ar = M_PI * (ro2 - ri2);
  double J = 2 * I;
  double Gmod = E / (2 * (1 + nu));
  Stiff s;
  s.K_bt = M3::Zero(); s.K_se = M3::Zero(); s.K_bt_inv = M3::Zero(); s.K_se_inv = M3::Zero();
  s.K_bt(0, 0) = E * I; s.K_bt(1, 1) = E * I; s.K_bt(2, 2) = J * Gmod;
  s.K_bt_inv(0, 0) = 1 / (E * I); s.K_bt_inv(1, 1) = 1 / (E * I); s.K_bt_inv(2, 2) = 1 / (J * Gmod);
  s.K_se(0, 0) = Gmod * Ar; s.K_se(1, 1) = Gmod * Ar; s.K_se(2, 2) = E * Ar;
  s.K_se_inv(0, 0) = 1 / (Gmod * Ar); s.K_se_inv(1, 1) = 1 / (Gmod * Ar); s.K_se_inv(2, 2) = 1 / (E * Ar);
  return s;
}

}  // namespace

extern "C" {

// reference get_r_info2 at arclength t; outputs row-major [N][3]
void fkref_r_info(int N, int Nc, int Nd, const double *C, const double *D, double t,
                  double *r, double *r_dot, double *r_ddot) {
  auto tendons = make_tendons(N, Nc, Nd, C, D);
  tendon::rInfo info;
  tendon::rInfoCache cache;
  info.resize(N);
  cache.resize(tendons);
  tendon::get_r_info2(tendons, t, info, cache);
  for (int j = 0; j < N; j++)
    for (int k = 0; k < 3; k++) {
      r[3 * j + k] = info.r[j][k];
      r_dot[3 * j + k] = info.r_dot[j][k];
      r_ddot[3 * j + k] = info.r_ddot[j][k];
    }
}

// reference tendon_deriv (unopt != 0: tendon_deriv_unopt).  x, dxdt: 19+N doubles.
void fkref_deriv(int N, int Nc, int Nd, const double *C, const double *D, const double *tau,
                 double ro, double ri, double E, double nu, const double *x, double t, double *dxdt,
                 int unopt) {
  auto tendons = make_tendons(N, Nc, Nd, C, D);
  std::vector<double> tv(tau, tau + N);
  Stiff K = stiffness(ro, ri, E, nu);
  tendon::State xs(x, x + 19 + N), dx(19 + N, 0.0);
  if (unopt) {
    tendon::tendon_deriv_unopt(xs, dx, t, tendons, tv, K.K_bt, K.K_se);
  } else {
    tendon::rInfo info;
    tendon::rInfoCache cache;
    info.resize(N);
    cache.resize(tendons);
    tendon::tendon_deriv(xs, dx, t, tendons, tv, K.K_bt, K.K_se, info, cache);
  }
  for (int i = 0; i < 19 + N; i++) dxdt[i] = dx[i];
}

// reference solve_initial_bending with the arguments of its call site tendon/TendonRobot.cpp:398-408
// (guesses (0,0,1)/(0,0,0), 1000 iterations, 1e-9 relative steps).  Returns iters.
int fkref_initial_bending(int N, int Nc, int Nd, const double *C, const double *D, const double *tau,
                          double ro, double ri, double E, double nu, double residual_threshold,
                          double s_start, double *v0, double *u0) {
  auto tendons = make_tendons(N, Nc, Nd, C, D);
  std::vector<double> tv(tau, tau + N);
  Stiff K = stiffness(ro, ri, E, nu);
  tendon::rInfo info;
  tendon::rInfoCache cache;
  info.resize(N);
  cache.resize(tendons);
  const V3 v_guess(0, 0, 1), u_guess(0, 0, 0);
  auto [v, u, iters] = tendon::solve_initial_bending(v_guess, u_guess, tendons, tv, K.K_bt, K.K_se,
                                                     K.K_bt_inv, K.K_se_inv, 1000, residual_threshold,
                                                     1e-9, 1e-9, s_start, info, cache);
  for (int k = 0; k < 3; k++) { v0[k] = v[k]; u0[k] = u[k]; }
  return iters;
}

// NOT the reference's integrator: a plain classic-RK4 walk over the caller's time grid `times`
// (h = min(dL, t_next - t) while t_next - t > eps, the documented integrate_times rule) whose every
// derivative is the reference's tendon_deriv and whose initial condition is the reference's
// solve_initial_bending.  states: [nt][19+N] row-major, states[0] = initial state.  Returns steps.
int fkref_shape(int N, int Nc, int Nd, const double *C, const double *D, const double *tau,
                double ro, double ri, double E, double nu, double residual_threshold, double dL,
                const double *times, int nt, double *states) {
  auto tendons = make_tendons(N, Nc, Nd, C, D);
  std::vector<double> tv(tau, tau + N);
  Stiff K = stiffness(ro, ri, E, nu);
  tendon::rInfo info;
  tendon::rInfoCache cache;
  info.resize(N);
  cache.resize(tendons);
  const V3 v_guess(0, 0, 1), u_guess(0, 0, 0);
  auto [v0, u0, iters] = tendon::solve_initial_bending(v_guess, u_guess, tendons, tv, K.K_bt, K.K_se,
                                                       K.K_bt_inv, K.K_se_inv, 1000, residual_threshold,
                                                       1e-9, 1e-9, times[0], info, cache);
  (void)iters;
  const int n = 19 + N;
  tendon::State x(n, 0.0), k1(n), k2(n), k3(n), k4(n), xt(n);
  x[3] = x[7] = x[11] = 1;
  for (int k = 0; k < 3; k++) { x[12 + k] = v0[k]; x[15 + k] = u0[k]; }
  auto f = [&](const tendon::State &a, tendon::State &d, double t) {
    tendon::tendon_deriv(a, d, t, tendons, tv, K.K_bt, K.K_se, info, cache);
  };
  for (int i = 0; i < n; i++) states[i] = x[i];
  int steps = 0;
  for (int s = 1; s < nt; s++) {
    double t = times[s - 1];
    const double t_next = times[s];
    while (t_next - t > 1e-15 * std::fmax(std::fabs(t), std::fabs(t_next)) && t_next - t > 0) {
      double h = std::fmin(dL, t_next - t);
      f(x, k1, t);
      for (int i = 0; i < n; i++) xt[i] = x[i] + 0.5 * h * k1[i];
      f(xt, k2, t + 0.5 * h);
      for (int i = 0; i < n; i++) xt[i] = x[i] + 0.5 * h * k2[i];
      f(xt, k3, t + 0.5 * h);
      for (int i = 0; i < n; i++) xt[i] = x[i] + h * k3[i];
      f(xt, k4, t + h);
      for (int i = 0; i < n; i++) x[i] += h / 6.0 * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
      t += h;
      steps++;
    }
    for (int i = 0; i < n; i++) states[s * n + i] = x[i];
  }
  return steps;
}

void fkref_closest_st_segment(const double *A, const double *B, const double *C, const double *D,
                              double *s, double *t) {
  auto [ss, tt] = collision::closest_st_segment(V3(A), V3(B), V3(C), V3(D));
  *s = ss;
  *t = tt;
}

int fkref_segment_aabox_intersect(const double *A, const double *B, const double *C, const double *D) {
  return collision::segment_aabox_intersect(V3(A), V3(B), V3(C), V3(D)) ? 1 : 0;
}

double fkref_closest_t_segment(const double *a, const double *b, const double *p) {
  return collision::closest_t_segment(V3(a), V3(b), V3(p));
}

}  // extern "C"
