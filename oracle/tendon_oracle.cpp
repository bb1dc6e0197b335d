/*
 * tendon_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * From-scratch restatement of the reference's FK / voxelise / voxel-check path,
 * following the cited reference lines operation by operation (own 3-vector and
 * 3x3 helpers in place of Eigen, own RK4 in place of Boost.odeint, own
 * interpolation / segment count in place of OMPL).  See tendon_oracle.h for the
 * parity status (partly pinned by pieces of the reference compiled into oracle/_ref;
 * "parity unpinned by the reference" for the rest).
 *
 * Citations are relative to /root/reference/cpp/src/.
 */
#include "tendon_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <tuple>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------
// tiny fixed-size linear algebra, evaluation order chosen to match Eigen's
// coefficient-based fixed-size products:  c(i,j) = (a(i,0)b(0,j)+a(i,1)b(1,j))+a(i,2)b(2,j)
// ---------------------------------------------------------------------------
struct V3 {
  double v[3];
  double &operator[](int i) { return v[i]; }
  const double &operator[](int i) const { return v[i]; }
};
struct M3 {
  double m[3][3];
};

inline V3 mk(double a, double b, double c) { return V3{{a, b, c}}; }
inline V3 operator+(const V3 &a, const V3 &b) { return mk(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline V3 operator-(const V3 &a, const V3 &b) { return mk(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline V3 operator-(const V3 &a) { return mk(-a[0], -a[1], -a[2]); }
inline V3 operator*(double s, const V3 &a) { return mk(s * a[0], s * a[1], s * a[2]); }
inline V3 operator*(const V3 &a, double s) { return mk(a[0] * s, a[1] * s, a[2] * s); }
inline V3 operator/(const V3 &a, double s) { return mk(a[0] / s, a[1] / s, a[2] / s); }
inline double dot(const V3 &a, const V3 &b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline double sqnorm(const V3 &a) { return dot(a, a); }
inline double norm(const V3 &a) { return std::sqrt(sqnorm(a)); }
inline V3 cross(const V3 &a, const V3 &b) {
  return mk(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
// Eigen normalized(): n / sqrt(z) if z > 0 else n
inline V3 normalized(const V3 &a) {
  double z = sqnorm(a);
  if (z > 0.0) return a / std::sqrt(z);
  return a;
}
inline V3 cabs(const V3 &a) { return mk(std::fabs(a[0]), std::fabs(a[1]), std::fabs(a[2])); }

inline M3 zero3() {
  M3 r;
  std::memset(&r, 0, sizeof(r));
  return r;
}
// util/vector_ops.h:53-59
inline M3 hat(const V3 &u) {
  M3 r;
  r.m[0][0] = 0;     r.m[0][1] = -u[2]; r.m[0][2] = u[1];
  r.m[1][0] = u[2];  r.m[1][1] = 0;     r.m[1][2] = -u[0];
  r.m[2][0] = -u[1]; r.m[2][1] = u[0];  r.m[2][2] = 0;
  return r;
}
inline M3 operator*(const M3 &a, const M3 &b) {
  M3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      c.m[i][j] = (a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j];
  return c;
}
inline V3 operator*(const M3 &a, const V3 &b) {
  V3 c;
  for (int i = 0; i < 3; i++) c[i] = (a.m[i][0] * b[0] + a.m[i][1] * b[1]) + a.m[i][2] * b[2];
  return c;
}
inline M3 operator*(double s, const M3 &a) {
  M3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c.m[i][j] = s * a.m[i][j];
  return c;
}
inline M3 operator/(const M3 &a, double s) {
  M3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c.m[i][j] = a.m[i][j] / s;
  return c;
}
inline M3 operator+(const M3 &a, const M3 &b) {
  M3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c.m[i][j] = a.m[i][j] + b.m[i][j];
  return c;
}
inline M3 operator-(const M3 &a, const M3 &b) {
  M3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c.m[i][j] = a.m[i][j] - b.m[i][j];
  return c;
}
inline M3 operator-(const M3 &a) {
  M3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c.m[i][j] = -a.m[i][j];
  return c;
}
// Eigen 3x3 inverse: cofactors / determinant (Eigen/src/LU/InverseImpl.h, restated
// from documented behaviour: adjugate times 1/det, det expanded along column 0)
inline double cof(const M3 &a, int i, int j) {
  int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return a.m[i1][j1] * a.m[i2][j2] - a.m[i1][j2] * a.m[i2][j1];
}
inline M3 inverse(const M3 &a) {
  double c00 = cof(a, 0, 0), c10 = cof(a, 1, 0), c20 = cof(a, 2, 0);
  double det = (c00 * a.m[0][0] + c10 * a.m[1][0]) + c20 * a.m[2][0];
  double invdet = 1.0 / det;
  M3 r;
  // result(i,j) = cofactor(j,i) * invdet
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = cof(a, j, i) * invdet;
  return r;
}

// ---------------------------------------------------------------------------
// stiffness -- tendon/TendonRobot.cpp:105-148
// ---------------------------------------------------------------------------
struct Stiff {
  M3 K_bt, K_se, K_bt_inv, K_se_inv;
};
Stiff get_stiffness(const orc_robot &rb) {
  double ro2 = rb.ro * rb.ro, ri2 = rb.ri * rb.ri;
  double I = (1.0 / 4.0) * M_PI * (ro2 * ro2 - ri2 * ri2);
  double Ar = M_PI * (ro2 - ri2);
  double J = 2 * I;
  double Gmod = rb.E / (2 * (1 + rb.nu));
  Stiff s;
  s.K_bt = zero3(); s.K_se = zero3(); s.K_bt_inv = zero3(); s.K_se_inv = zero3();
  s.K_bt.m[0][0] = rb.E * I; s.K_bt.m[1][1] = rb.E * I; s.K_bt.m[2][2] = J * Gmod;
  s.K_bt_inv.m[0][0] = 1 / (rb.E * I); s.K_bt_inv.m[1][1] = 1 / (rb.E * I);
  s.K_bt_inv.m[2][2] = 1 / (J * Gmod);
  s.K_se.m[0][0] = Gmod * Ar; s.K_se.m[1][1] = Gmod * Ar; s.K_se.m[2][2] = rb.E * Ar;
  s.K_se_inv.m[0][0] = 1 / (Gmod * Ar); s.K_se_inv.m[1][1] = 1 / (Gmod * Ar);
  s.K_se_inv.m[2][2] = 1 / (rb.E * Ar);
  return s;
}

// ---------------------------------------------------------------------------
// routing -- tendon/get_r_info.cpp:17-40 (get_poly_vecs), :105-144 (get_r_info2)
// ---------------------------------------------------------------------------
struct RInfo {
  V3 r[ORC_MAX_TENDONS], rd[ORC_MAX_TENDONS], rdd[ORC_MAX_TENDONS];
};
void get_r_info2(const orc_robot &rb, double t, RInfo &info) {
  const int Nt = rb.n_tendons, Na = rb.n_c, Nm = rb.n_d;
  const int Ns = std::max(Na, Nm);
  double S[ORC_MAX_COEF], Sd[ORC_MAX_COEF], Sdd[ORC_MAX_COEF];
  S[0] = 1; Sd[0] = 0; Sdd[0] = 0;
  if (Ns >= 2) { S[1] = t; Sd[1] = 1; Sdd[1] = 0; }
  for (int i = 2; i < Ns; i++) {
    S[i] = t * S[i - 1];
    Sd[i] = i * S[i - 1];
    Sdd[i] = i * (i - 1) * S[i - 2];
  }
  for (int j = 0; j < Nt; j++) {
    const double *C = rb.C + j * ORC_MAX_COEF, *D = rb.D + j * ORC_MAX_COEF;
    double C_a = 0, C_ad = 0, C_add = 0, D_m = 0, D_md = 0, D_mdd = 0;
    for (int i = 0; i < Na; i++) { C_a += C[i] * S[i]; C_ad += C[i] * Sd[i]; C_add += C[i] * Sdd[i]; }
    for (int i = 0; i < Nm; i++) { D_m += D[i] * S[i]; D_md += D[i] * Sd[i]; D_mdd += D[i] * Sdd[i]; }
    double sa = std::sin(C_a), ca = std::cos(C_a);
    info.r[j] = D_m * mk(sa, ca, 0);
    info.rd[j] = D_md * mk(sa, ca, 0) + D_m * mk(ca * C_ad, -sa * C_ad, 0);
    info.rdd[j] = ((D_mdd * mk(sa, ca, 0) + (2 * D_md) * mk(ca * C_ad, -sa * C_ad, 0))
                   - D_m * mk(sa * C_ad * C_ad, ca * C_ad * C_ad, 0))
                  + D_m * mk(ca * C_add, -sa * C_add, 0);
  }
}

// ---------------------------------------------------------------------------
// derivative -- tendon/tendon_deriv.cpp:60-87 (linsubsolve2), :95-178
// state x = [p(3), R(9, column-major), v(3), u(3), L, L_i(N)]
// ---------------------------------------------------------------------------
struct DerivBlocks {
  M3 Ktl, G, B, Kbr;  // [K_se + A, G; B, K_bt + H]
  V3 d, c;
};
void deriv_blocks(const orc_robot &rb, const Stiff &Ks, const double *tau, const double *x,
                  double t, DerivBlocks &blk, double *si_dot, M3 &R, V3 &v, V3 &u) {
  const int Nt = rb.n_tendons;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R.m[i][j] = x[3 + i + 3 * j];
  v = mk(x[12], x[13], x[14]);
  u = mk(x[15], x[16], x[17]);
  M3 vhat = hat(v), uhat = hat(u);
  RInfo rs;
  get_r_info2(rb, t, rs);
  M3 A = zero3(), B = zero3(), G = zero3(), H = zero3();
  V3 a = mk(0, 0, 0), b = mk(0, 0, 0);
  for (int j = 0; j < Nt; j++) {
    M3 rhat = hat(rs.r[j]);
    V3 pi_dot_b = (uhat * rs.r[j] + rs.rd[j]) + v;
    M3 ph = hat(pi_dot_b);
    si_dot[j] = norm(pi_dot_b);
    M3 Ai = (((-tau[j]) * ph) * ph) / (si_dot[j] * si_dot[j] * si_dot[j]);
    M3 Bi = rhat * Ai;
    M3 Gi = (-Ai) * rhat;
    M3 Hi = (-Bi) * rhat;
    V3 ai = Ai * ((uhat * pi_dot_b + uhat * rs.rd[j]) + rs.rdd[j]);
    V3 bi = rhat * ai;
    A = A + Ai; B = B + Bi; G = G + Gi; H = H + Hi;
    a = a + ai; b = b + bi;
  }
  V3 vmv = v - mk(0, 0, 1);
  blk.c = (((-uhat) * Ks.K_bt) * u - ((vhat * Ks.K_se) * vmv)) - b;
  blk.d = (((-uhat) * Ks.K_se) * vmv) - a;
  blk.Ktl = Ks.K_se + A;
  blk.G = G;
  blk.B = B;
  blk.Kbr = Ks.K_bt + H;
}

void finish_deriv(const orc_robot &rb, const M3 &R, const V3 &v, const V3 &u, const V3 &v_dot,
                  const V3 &u_dot, const double *si_dot, double *dxdt) {
  V3 p_dot = R * v;
  M3 R_dot = R * hat(u);
  for (int i = 0; i < 3; i++) dxdt[i] = p_dot[i];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) dxdt[3 + i + 3 * j] = R_dot.m[i][j];
  for (int i = 0; i < 3; i++) dxdt[12 + i] = v_dot[i];
  for (int i = 0; i < 3; i++) dxdt[15 + i] = u_dot[i];
  dxdt[18] = norm(v);
  for (int j = 0; j < rb.n_tendons; j++) dxdt[19 + j] = si_dot[j];
}

void tendon_deriv(const orc_robot &rb, const Stiff &Ks, const double *tau, const double *x,
                  double t, double *dxdt) {
  DerivBlocks k;
  double si_dot[ORC_MAX_TENDONS];
  M3 R;
  V3 v, u;
  deriv_blocks(rb, Ks, tau, x, t, k, si_dot, R, v, u);
  // linsubsolve2(A=Ktl, B=G, C=B, D=Kbr, a=d, b=c): blockwise inverse
  M3 Ai = inverse(k.Ktl);
  M3 Gs = k.Kbr - (k.B * Ai) * k.G;
  M3 Gi = inverse(Gs);
  M3 AiB = Ai * k.G;
  M3 CAi = k.B * Ai;
  M3 M00 = Ai + (AiB * Gi) * CAi;
  M3 M03 = (-AiB) * Gi;
  M3 M30 = (-Gi) * CAi;
  M3 M33 = Gi;
  V3 v_dot, u_dot;
  for (int i = 0; i < 3; i++) {
    v_dot[i] = ((((M00.m[i][0] * k.d[0] + M00.m[i][1] * k.d[1]) + M00.m[i][2] * k.d[2])
                 + M03.m[i][0] * k.c[0]) + M03.m[i][1] * k.c[1]) + M03.m[i][2] * k.c[2];
    u_dot[i] = ((((M30.m[i][0] * k.d[0] + M30.m[i][1] * k.d[1]) + M30.m[i][2] * k.d[2])
                 + M33.m[i][0] * k.c[0]) + M33.m[i][1] * k.c[1]) + M33.m[i][2] * k.c[2];
  }
  finish_deriv(rb, R, v, u, v_dot, u_dot, si_dot, dxdt);
}

// alternative linear solve: dense 6x6 Gaussian elimination with partial pivoting
void tendon_deriv_alt(const orc_robot &rb, const Stiff &Ks, const double *tau, const double *x,
                      double t, double *dxdt) {
  DerivBlocks k;
  double si_dot[ORC_MAX_TENDONS];
  M3 R;
  V3 v, u;
  deriv_blocks(rb, Ks, tau, x, t, k, si_dot, R, v, u);
  double M[6][7];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      M[i][j] = k.Ktl.m[i][j];
      M[i][j + 3] = k.G.m[i][j];
      M[i + 3][j] = k.B.m[i][j];
      M[i + 3][j + 3] = k.Kbr.m[i][j];
    }
  for (int i = 0; i < 3; i++) { M[i][6] = k.d[i]; M[i + 3][6] = k.c[i]; }
  for (int c = 0; c < 6; c++) {
    int piv = c;
    for (int r = c + 1; r < 6; r++)
      if (std::fabs(M[r][c]) > std::fabs(M[piv][c])) piv = r;
    if (piv != c)
      for (int j = 0; j < 7; j++) std::swap(M[c][j], M[piv][j]);
    for (int r = c + 1; r < 6; r++) {
      double f = M[r][c] / M[c][c];
      for (int j = c; j < 7; j++) M[r][j] -= f * M[c][j];
    }
  }
  double sol[6];
  for (int r = 5; r >= 0; r--) {
    double s = M[r][6];
    for (int j = r + 1; j < 6; j++) s -= M[r][j] * sol[j];
    sol[r] = s / M[r][r];
  }
  finish_deriv(rb, R, v, u, mk(sol[0], sol[1], sol[2]), mk(sol[3], sol[4], sol[5]), si_dot, dxdt);
}

// ---------------------------------------------------------------------------
// t grid -- util/vector_ops.h:67-75 (range), tendon/TendonRobot.cpp:69-84 (t_range)
// ---------------------------------------------------------------------------
std::vector<double> t_range(double start, double end, double dt) {
  std::vector<double> t;
  for (double p = start; p <= end - (dt / 2); p += dt) t.push_back(p);
  t.push_back(end);
  for (auto &val : t) val = end - (val - start);
  std::reverse(t.begin(), t.end());
  return t;
}

// ---------------------------------------------------------------------------
// initial condition -- tendon/solve_initial_bending.cpp:15-73
// ---------------------------------------------------------------------------
int solve_initial_bending(const orc_robot &rb, const Stiff &Ks, const double *tau, double s_start,
                          int iter_max, double residual_threshold, double dv_threshold,
                          double du_threshold, V3 &v, V3 &u) {
  v = mk(0, 0, 1);
  u = mk(0, 0, 0);
  const int Nt = rb.n_tendons;
  RInfo info;
  get_r_info2(rb, s_start, info);
  M3 rhat[ORC_MAX_TENDONS];
  for (int k = 0; k < Nt; k++) rhat[k] = hat(info.r[k]);
  int iters = 0;
  for (iters = 0; iters < iter_max; ++iters) {
    M3 uhat = hat(u);
    V3 Ft = mk(0, 0, 0), Lt = mk(0, 0, 0);
    for (int k = 0; k < Nt; ++k) {
      V3 pi_dot_unit = normalized((uhat * info.r[k] + info.rd[k]) + v);
      Ft = Ft - tau[k] * pi_dot_unit;
      Lt = Lt - (tau[k] * rhat[k]) * pi_dot_unit;
    }
    V3 n = Ks.K_se * (v - mk(0, 0, 1));
    V3 m = Ks.K_bt * u;
    double residual = std::sqrt(sqnorm(n - Ft) + sqnorm(m - Lt));
    if (residual < residual_threshold) break;
    V3 v_new = Ks.K_se_inv * Ft + mk(0, 0, 1);
    V3 u_new = Ks.K_bt_inv * Lt;
    if (norm(v_new - v) < dv_threshold * norm(v) && norm(u_new - u) < du_threshold * norm(u)) break;
    v = v_new;
    u = u_new;
  }
  return iters;
}

// PointForces::calc_point_forces residual -- tendon/TendonRobot.cpp:188-217
double base_residual(const orc_robot &rb, const Stiff &Ks, const double *tau, const M3 &R,
                     const V3 &u, const V3 &v, const RInfo &rs) {
  V3 n = (R * Ks.K_se) * (v - mk(0, 0, 1));
  V3 m = (R * Ks.K_bt) * u;
  V3 F_t = mk(0, 0, 0), L_t = mk(0, 0, 0);
  for (int i = 0; i < rb.n_tendons; i++) {
    V3 pdot_unit = normalized(R * ((cross(u, rs.r[i]) + rs.rd[i]) + v));
    V3 F_ti = (-tau[i]) * pdot_unit;
    V3 L_ti = cross(R * rs.r[i], F_ti);
    F_t = F_t + F_ti;
    L_t = L_t + L_ti;
  }
  V3 F_e = n - F_t, L_e = m - L_t;
  return std::sqrt(sqnorm(F_e) + sqnorm(L_e));
}

// ---------------------------------------------------------------------------
// tension_shape -- tendon/TendonRobot.cpp:325-500; integrate_times(runge_kutta4)
// restated from Boost.odeint's documented behaviour (integrate_times.hpp:
// observe at each time, then step with min(dt, t_next - t) while
// t_next - t > eps; runge_kutta4 = classic tableau as a generic RK).
// ---------------------------------------------------------------------------
struct Shape {
  std::vector<double> t;
  std::vector<V3> p;
  std::vector<M3> R;
  orc_fk_out out;
};

void rk4_step(const orc_robot &rb, const Stiff &Ks, const double *tau, double *x, double t,
              double dt, int n) {
  double k1[19 + ORC_MAX_TENDONS], k2[19 + ORC_MAX_TENDONS], k3[19 + ORC_MAX_TENDONS],
      k4[19 + ORC_MAX_TENDONS], xt[19 + ORC_MAX_TENDONS];
  const double a = dt * 0.5;
  tendon_deriv(rb, Ks, tau, x, t, k1);
  for (int i = 0; i < n; i++) xt[i] = x[i] + a * k1[i];
  tendon_deriv(rb, Ks, tau, xt, t + 0.5 * dt, k2);
  for (int i = 0; i < n; i++) xt[i] = x[i] + a * k2[i];
  tendon_deriv(rb, Ks, tau, xt, t + 0.5 * dt, k3);
  for (int i = 0; i < n; i++) xt[i] = x[i] + dt * k3[i];
  tendon_deriv(rb, Ks, tau, xt, t + dt, k4);
  const double b1 = dt * (1.0 / 6.0), b2 = dt * (1.0 / 3.0);
  for (int i = 0; i < n; i++) x[i] = (((x[i] + b1 * k1[i]) + b2 * k2[i]) + b2 * k3[i]) + b1 * k4[i];
}

void tension_shape(const orc_robot &rb, const double *tau, double s_start, Shape &res) {
  const int N = rb.n_tendons;
  std::memset(&res.out, 0, sizeof(res.out));
  res.t.clear(); res.p.clear(); res.R.clear();
  res.out.converged = 1;
  if (s_start > rb.L) s_start = rb.L;  // :359 (s_start < 0 is NOT clamped)
  M3 I3 = zero3();
  I3.m[0][0] = I3.m[1][1] = I3.m[2][2] = 1;
  if (s_start == rb.L) {  // :361-372
    res.t.push_back(s_start);
    res.p.push_back(mk(0, 0, 0));
    res.R.push_back(I3);
    res.out.L = 0;
    res.out.v_i[2] = 1; res.out.v_f[2] = 1;
    res.out.npts = 1;
    return;
  }
  const Stiff Ks = get_stiffness(rb);
  V3 v0, u0;
  int iters = solve_initial_bending(rb, Ks, tau, s_start, 1000, rb.residual_threshold, 1e-9, 1e-9,
                                    v0, u0);
  res.out.iters = iters;
  const int n = 19 + N;
  double x[19 + ORC_MAX_TENDONS];
  for (int i = 0; i < n; i++) x[i] = 0;
  x[3] = 1; x[7] = 1; x[11] = 1;
  for (int i = 0; i < 3; i++) { x[12 + i] = v0[i]; x[15 + i] = u0[i]; res.out.v_i[i] = v0[i]; res.out.u_i[i] = u0[i]; }

  res.t = t_range(s_start, rb.L, rb.dL);
  const double dt = rb.dL;
  const double eps = std::numeric_limits<double>::epsilon();
  int nsteps = 0;
  size_t it = 0;
  double current_dt = dt;
  while (true) {
    double current_time = res.t[it++];
    // observer: copy p and R
    res.p.push_back(mk(x[0], x[1], x[2]));
    M3 Rm;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Rm.m[i][j] = x[3 + i + 3 * j];
    res.R.push_back(Rm);
    if (it == res.t.size()) break;
    // less_with_sign(t1, t2, dt>0) := (t2 - t1) > eps
    while ((res.t[it] - current_time) > eps) {
      current_dt = std::min(dt, res.t[it] - current_time);
      rk4_step(rb, Ks, tau, x, current_time, current_dt, n);
      ++nsteps;
      current_time += current_dt;
      current_dt = std::max(dt, current_dt);
    }
  }
  res.out.nsteps = nsteps;
  res.out.npts = (int)res.t.size();
  res.out.L = x[18];
  for (int j = 0; j < N; j++) res.out.L_i[j] = x[19 + j];
  for (int i = 0; i < 3; i++) { res.out.v_f[i] = x[12 + i]; res.out.u_f[i] = x[15 + i]; }
  RInfo rs;
  get_r_info2(rb, s_start, rs);
  double resid = base_residual(rb, Ks, tau, res.R.front(), u0, v0, rs);
  res.out.converged = (resid <= rb.residual_threshold) ? 1 : 0;
}

// TendonResult::rotate_z -- tendon/TendonResult.cpp:13-18 (AngleAxis about z)
void rotate_z(Shape &s, double theta) {
  double c = std::cos(theta), sn = std::sin(theta);
  M3 rot = zero3();
  rot.m[0][0] = c; rot.m[0][1] = -sn; rot.m[1][0] = sn; rot.m[1][1] = c; rot.m[2][2] = 1;
  for (auto &p : s.p) p = rot * p;
  for (auto &R : s.R) R = rot * R;
}

// TendonRobot::shape(state) -- tendon/TendonRobot.h:105-131
void robot_shape(const orc_robot &rb, const double *state, Shape &res) {
  const int N = rb.n_tendons;
  double rotate = rb.enable_rotation ? state[N] : 0.0;
  double retract = rb.enable_retraction ? state[orc_state_size(&rb) - 1] : 0.0;
  tension_shape(rb, state, retract, res);
  if (rb.enable_rotation) rotate_z(res, rotate);
}

// TendonSpecs::is_straight / is_helix -- tendon/TendonSpecs.cpp:17-30
int poly_degree(const double *coef, int n) {
  if (n == 0) return 0;
  for (int i = n - 1; i > 0; i--)
    if (std::fabs(coef[i]) > 0.0) return i;
  return 0;
}

// home_shape L_i -- tendon/TendonRobot.cpp:249-314 (closed forms only; the general
// branch calls simpsons() which reads out of bounds in the reference)
void home_lengths(const orc_robot &rb, double s_start, double *L_i) {
  if (s_start < 0.0) s_start = 0.0;
  if (s_start > rb.L) s_start = rb.L;
  if (s_start == rb.L) {
    for (int j = 0; j < rb.n_tendons; j++) L_i[j] = 0.0;
    return;
  }
  double Lres = rb.L - s_start;
  for (int j = 0; j < rb.n_tendons; j++) {
    const double *C = rb.C + j * ORC_MAX_COEF, *D = rb.D + j * ORC_MAX_COEF;
    int rdeg = poly_degree(D, rb.n_d), tdeg = poly_degree(C, rb.n_c);
    if (rdeg == 0 && tdeg == 0) {
      L_i[j] = Lres;
    } else if (rdeg == 0 && tdeg == 1) {
      double d0 = D[0], c1 = C[1];
      L_i[j] = Lres * std::sqrt(1 + d0 * d0 * c1 * c1);
    } else {
      L_i[j] = std::numeric_limits<double>::quiet_NaN();  // unsupported (reference UB)
    }
  }
}

// ---------------------------------------------------------------------------
// self collision -- collision/collision_primitives.cpp:10-102, collision.hxx:56-108,
// collision/collision.cpp:6-46
// ---------------------------------------------------------------------------
inline double bound01(double t) { return std::max(0.0, std::min(1.0, t)); }

std::pair<double, double> closest_st_segment(const V3 &A, const V3 &B, const V3 &C, const V3 &D) {
  const double eps = std::numeric_limits<double>::epsilon();
  const double eps_squared = eps * eps;
  double s = 0.0, t = 0.0;
  const V3 AB = B - A, CD = D - C;
  const double a = dot(AB, AB), c = dot(CD, CD);
  auto closest_AB_s = [&](const V3 &P) { return (a <= eps_squared) ? 0.0 : dot(AB, P - A) / a; };
  auto closest_CD_t = [&](const V3 &P) { return (c <= eps_squared) ? 0.0 : dot(CD, P - C) / c; };
  if (a <= eps_squared) return {0.0, bound01(closest_CD_t(A))};
  if (c <= eps_squared) return {bound01(closest_AB_s(C)), 0.0};
  const V3 AC = C - A;
  const double b = dot(AB, CD), d = dot(AC, AB), e = dot(AC, CD);
  const double denom = std::max(0.0, a * c - b * b);
  if (denom <= eps_squared) {
    t = closest_CD_t(A);
    if (0.0 <= t && t <= 1.0) return {0.0, t};
    t = closest_CD_t(B);
    if (0.0 <= t && t <= 1.0) return {1.0, t};
    s = closest_AB_s(C);
    if (0.0 <= s && s <= 1.0) return {s, 0.0};
    const V3 AD = D - A, BC = C - B, BD = D - B;
    const double ac2 = dot(AC, AC), ad2 = dot(AD, AD), bc2 = dot(BC, BC), bd2 = dot(BD, BD);
    if (ac2 <= ad2 && ac2 <= bc2 && ac2 <= bd2) return {0.0, 0.0};
    if (ad2 <= bc2 && ad2 <= bd2) return {0.0, 1.0};
    if (bc2 <= bd2) return {1.0, 0.0};
    return {1.0, 1.0};
  }
  s = (c * d - b * e) / denom;
  t = (b * d - a * e) / denom;
  if (0.0 <= t && t <= 1.0) return {bound01(s), t};
  if (t < 0.0) return {bound01(-c / a), 0.0};
  return {bound01((b - c) / a), 1.0};
}

inline V3 interp_pt(const V3 &a, const V3 &b, double t) { return a + (b - a) * t; }

bool capsules_collide(const V3 &a0, const V3 &a1, const V3 &b0, const V3 &b1, double r1,
                      double r2) {
  auto [s, t] = closest_st_segment(a0, a1, b0, b1);
  V3 c1 = interp_pt(a0, a1, s), c2 = interp_pt(b0, b1, t);
  V3 diff = c1 - c2;
  double rr = r1 + r2;
  return dot(diff, diff) <= (rr * rr);
}

bool collides_self(const V3 *pts, int N, double r) {
  double dist_to_consider = 3.0 * r;
  if (N <= 2) return false;
  std::vector<double> acc(N);
  double dist = 0.0;
  V3 prev = pts[0];
  for (int i = 0; i < N; i++) {
    dist += norm(pts[i] - prev);
    acc[i] = dist;
    prev = pts[i];
  }
  for (int a = 0; a + 3 < N; ++a) {
    for (int b = a + 2; b < N - 1; ++b) {
      if (acc[b] - acc[a + 1] < dist_to_consider) continue;
      if (capsules_collide(pts[a], pts[a + 1], pts[b], pts[b + 1], r, r)) return true;
    }
  }
  return false;
}

uint32_t validity_flags(const orc_robot &rb, const double *state, const orc_fk_out &fk,
                        const V3 *p) {
  uint32_t f = 0;
  if (!fk.converged) f |= ORC_FLAG_NONCONVERGED;
  double home[ORC_MAX_TENDONS];
  double retract = rb.enable_retraction ? state[orc_state_size(&rb) - 1] : 0.0;
  home_lengths(rb, retract, home);
  for (int i = 0; i < rb.n_tendons; i++) {
    double dl = home[i] - fk.L_i[i];
    if (dl < rb.min_length[i] || rb.max_length[i] < dl) f |= ORC_FLAG_LENGTH_LIMIT;
  }
  if (collides_self(p, fk.npts, rb.r)) f |= ORC_FLAG_SELF_COLLISION;
  return f;
}

}  // namespace

// ===========================================================================
// voxel octree -- collision/detail/TreeNode.{h,hxx}, collision/VoxelOctree.cpp
// ===========================================================================
struct OrcNode {
  OrcNode *ch[8];
  uint64_t bits;
  OrcNode() : bits(0) {
    for (auto &c : ch) c = nullptr;
  }
};

struct orc_octree {
  int Ng;
  double xmin, xmax, ymin, ymax, zmin, zmax, dx, dy, dz;
  double inv_rot[9];
  OrcNode *root;
};

namespace {

void node_free(OrcNode *n) {
  if (!n) return;
  for (auto c : n->ch) node_free(c);
  delete n;
}
OrcNode *node_copy(const OrcNode *n) {
  if (!n) return nullptr;
  OrcNode *r = new OrcNode();
  r->bits = n->bits;
  for (int i = 0; i < 8; i++) r->ch[i] = node_copy(n->ch[i]);
  return r;
}
// Nbt = blocks per axis of this node; leaf when Nbt == 1
inline int child_idx(int bx, int by, int bz, int c) { return (bz / c) + 2 * (by / c) + 4 * (bx / c); }

bool node_is_empty(const OrcNode *n, int Nbt) {
  if (Nbt == 1) return !n->bits;
  for (auto c : n->ch)
    if (c) return false;
  return true;
}
uint64_t node_block(const OrcNode *n, int Nbt, int bx, int by, int bz) {
  if (Nbt == 1) return n->bits;
  int c = Nbt / 2;
  const OrcNode *child = n->ch[child_idx(bx, by, bz, c)];
  if (child) return node_block(child, c, bx % c, by % c, bz % c);
  return 0;
}
void node_set_block(OrcNode *n, int Nbt, int bx, int by, int bz, uint64_t value) {
  if (Nbt == 1) { n->bits = value; return; }
  int c = Nbt / 2;
  OrcNode *&child = n->ch[child_idx(bx, by, bz, c)];
  if (!child && value) child = new OrcNode();
  if (child) {
    node_set_block(child, c, bx % c, by % c, bz % c, value);
    if (!value && node_is_empty(child, c)) { node_free(child); child = nullptr; }
  }
}
uint64_t node_union_block(OrcNode *n, int Nbt, int bx, int by, int bz, uint64_t value) {
  if (Nbt == 1) { uint64_t prev = n->bits; n->bits |= value; return prev; }
  int c = Nbt / 2;
  OrcNode *&child = n->ch[child_idx(bx, by, bz, c)];
  if (!child) child = new OrcNode();
  return node_union_block(child, c, bx % c, by % c, bz % c, value);
}
void node_union_tree(OrcNode *a, const OrcNode *b, int Nbt) {
  if (Nbt == 1) { a->bits |= b->bits; return; }
  for (int i = 0; i < 8; i++) {
    if (a->ch[i] && b->ch[i]) node_union_tree(a->ch[i], b->ch[i], Nbt / 2);
    else if (!a->ch[i] && b->ch[i]) a->ch[i] = node_copy(b->ch[i]);
  }
}
bool node_collides(const OrcNode *a, const OrcNode *b, int Nbt) {
  if (Nbt == 1) return (a->bits & b->bits) != 0;
  for (int i = 0; i < 8; i++)
    if (a->ch[i] && b->ch[i] && node_collides(a->ch[i], b->ch[i], Nbt / 2)) return true;
  return false;
}
int64_t node_nblocks(const OrcNode *n, int Nbt) {
  if (Nbt == 1) return 1;
  int64_t s = 0;
  for (auto c : n->ch)
    if (c) s += node_nblocks(c, Nbt / 2);
  return s;
}
template <typename F>
void node_visit_leaves(const OrcNode *n, int Nbt, int ox, int oy, int oz, const F &f) {
  if (Nbt == 1) { f(ox, oy, oz, n->bits); return; }
  int c = Nbt / 2;
  for (int bx = 0; bx < Nbt; bx += c)
    for (int by = 0; by < Nbt; by += c)
      for (int bz = 0; bz < Nbt; bz += c) {
        const OrcNode *child = n->ch[child_idx(bx, by, bz, c)];
        if (child) node_visit_leaves(child, c, ox + bx, oy + by, oz + bz, f);
      }
}

inline uint64_t bitmask(int x, int y, int z) { return uint64_t(1) << (x * 16 + y * 4 + z); }

// VoxelOctree::set_cell(ix,iy,iz,true) -- collision/VoxelOctree.cpp:262-272
inline void set_cell(orc_octree *t, int ix, int iy, int iz) {
  uint64_t mask = bitmask(ix % 4, iy % 4, iz % 4);
  node_union_block(t->root, t->Ng / 4, ix / 4, iy / 4, iz / 4, mask);
}

// collision/collision_primitives.h:62-85
bool segment_aabox_intersect(const V3 &A, const V3 &B, const V3 &C, const V3 &D) {
  const V3 AB = B - A;
  const double len = norm(AB) / 2;
  const V3 U = AB / (2 * len);
  const V3 Uabs = cabs(U);
  const V3 P = (A + B) / 2 - (D + C) / 2;
  const V3 ext = cabs(D - C) / 2;
  const V3 UxP = cabs(cross(U, P));
  const V3 Pabs = cabs(P);
  bool intersects = Pabs[0] > ext[0] + len * Uabs[0] || Pabs[1] > ext[1] + len * Uabs[1] ||
                    Pabs[2] > ext[2] + len * Uabs[2] ||
                    UxP[0] > ext[1] * Uabs[2] + ext[2] * Uabs[1] ||
                    UxP[1] > ext[2] * Uabs[0] + ext[0] * Uabs[2] ||
                    UxP[2] > ext[0] * Uabs[1] + ext[1] * Uabs[0];
  return !intersects;
}

// VoxelOctree::add_line -- collision/VoxelOctree.cpp:325-426 (reproduced literally,
// including the "index times metric cell size" initial-error quirk and the overshoot)
void add_line(orc_octree *t, const V3 &a, const V3 &b) {
  const V3 ll = mk(t->xmin, t->ymin, t->zmin), ur = mk(t->xmax, t->ymax, t->zmax);
  if (!segment_aabox_intersect(a, b, ll, ur)) return;
  const V3 nvpm = mk(1 / t->dx, 1 / t->dy, 1 / t->dz);
  const V3 A = mk((a[0] - ll[0]) * nvpm[0], (a[1] - ll[1]) * nvpm[1], (a[2] - ll[2]) * nvpm[2]);
  const V3 B = mk((b[0] - ll[0]) * nvpm[0], (b[1] - ll[1]) * nvpm[1], (b[2] - ll[2]) * nvpm[2]);
  const int Axi = int(A[0]) - (A[0] < 0), Ayi = int(A[1]) - (A[1] < 0), Azi = int(A[2]) - (A[2] < 0);
  const int Bxi = int(B[0]) - (B[0] < 0), Byi = int(B[1]) - (B[1] < 0), Bzi = int(B[2]) - (B[2] < 0);
  const int N = t->Ng;
  auto idx_is_in = [N](int x) { return 0 <= x && x < N; };
  auto voxel_is_in = [&](int x, int y, int z) { return idx_is_in(x) && idx_is_in(y) && idx_is_in(z); };
  bool entered = voxel_is_in(Axi, Ayi, Azi);
  if (entered) set_cell(t, Axi, Ayi, Azi);
  if (voxel_is_in(Bxi, Byi, Bzi)) set_cell(t, Bxi, Byi, Bzi);
  const V3 U = normalized(B - A);
  const int step_x = 1 - 2 * (U[0] < 0), step_y = 1 - 2 * (U[1] < 0), step_z = 1 - 2 * (U[2] < 0);
  const double ex = std::fabs(A[0] - (Axi + step_x) * t->dx);
  const double ey = std::fabs(A[1] - (Ayi + step_y) * t->dy);
  const double ez = std::fabs(A[2] - (Azi + step_z) * t->dz);
  const V3 Uabs = cabs(U);
  const double threshold = 1e-10;
  const double tx_delta = (Uabs[0] > threshold) ? 1 / Uabs[0] : 1 / threshold;
  const double ty_delta = (Uabs[1] > threshold) ? 1 / Uabs[1] : 1 / threshold;
  const double tz_delta = (Uabs[2] > threshold) ? 1 / Uabs[2] : 1 / threshold;
  double tx = std::fabs(ex * tx_delta), ty = std::fabs(ey * ty_delta), tz = std::fabs(ez * tz_delta);
  int xi = Axi, yi = Ayi, zi = Azi;
  while (step_x * (Bxi - xi) >= 0 && step_y * (Byi - yi) >= 0 && step_z * (Bzi - zi) >= 0) {
    const bool tx_is_min = (tx < ty) && (tx < tz);
    const bool ty_is_min = !(tx < ty) && (ty < tz);
    if (tx_is_min) {
      xi += step_x;
      if (entered && !idx_is_in(xi)) break;
      tx += tx_delta;
    } else if (ty_is_min) {
      yi += step_y;
      if (entered && !idx_is_in(yi)) break;
      ty += ty_delta;
    } else {
      zi += step_z;
      if (entered && !idx_is_in(zi)) break;
      tz += tz_delta;
    }
    if (!entered && voxel_is_in(xi, yi, zi)) entered = true;
    if (entered) set_cell(t, xi, yi, zi);
  }
}

inline V3 rotate_point(const double *inv_rot, const V3 &p) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = inv_rot[3 * i + j];
  return r * p;
}

// find_cell -- collision/VoxelOctree.cpp:309-317 + domain_check :1511-1521
bool find_cell(const orc_octree *t, const V3 &p, long *c) {
  if (p[0] < t->xmin || t->xmax < p[0]) return false;
  if (p[1] < t->ymin || t->ymax < p[1]) return false;
  if (p[2] < t->zmin || t->zmax < p[2]) return false;
  c[0] = (long)size_t((p[0] - t->xmin) / t->dx);
  c[1] = (long)size_t((p[1] - t->ymin) / t->dy);
  c[2] = (long)size_t((p[2] - t->zmin) / t->dz);
  return true;
}

orc_octree *octree_from_grid(const orc_grid *g) {
  orc_octree *t = new orc_octree();
  t->Ng = g->Ng;
  t->xmin = g->lim[0]; t->xmax = g->lim[1];
  t->ymin = g->lim[2]; t->ymax = g->lim[3];
  t->zmin = g->lim[4]; t->zmax = g->lim[5];
  // set_xlim -- collision/VoxelOctree.cpp:152-177
  t->dx = (t->xmax - t->xmin) / g->Ng;
  t->dy = (t->ymax - t->ymin) / g->Ng;
  t->dz = (t->zmax - t->zmin) / g->Ng;
  std::memcpy(t->inv_rot, g->inv_rot, sizeof(t->inv_rot));
  t->root = new OrcNode();
  return t;
}

// util/angles.h:13-33 (canonical_angle) is only used by the non-OMPL interp; the
// planner path uses OMPL's compound interpolate, restated in orc_interpolate.

}  // namespace

// ===========================================================================
// C API
// ===========================================================================
extern "C" {

int orc_state_size(const orc_robot *rb) {
  return rb->n_tendons + (rb->enable_rotation ? 1 : 0) + (rb->enable_retraction ? 1 : 0);
}

int orc_t_range(double s, double L, double dL, double *out, int cap) {
  auto t = t_range(s, L, dL);
  if ((int)t.size() > cap) return -1;
  for (size_t i = 0; i < t.size(); i++) out[i] = t[i];
  return (int)t.size();
}

void orc_routing(const orc_robot *rb, double t, double *r, double *rd, double *rdd) {
  RInfo info;
  get_r_info2(*rb, t, info);
  for (int j = 0; j < rb->n_tendons; j++)
    for (int k = 0; k < 3; k++) {
      r[3 * j + k] = info.r[j][k];
      rd[3 * j + k] = info.rd[j][k];
      rdd[3 * j + k] = info.rdd[j][k];
    }
}

void orc_tendon_deriv(const orc_robot *rb, const double *tau, const double *x, double t,
                      double *dxdt) {
  tendon_deriv(*rb, get_stiffness(*rb), tau, x, t, dxdt);
}
void orc_tendon_deriv_alt(const orc_robot *rb, const double *tau, const double *x, double t,
                          double *dxdt) {
  tendon_deriv_alt(*rb, get_stiffness(*rb), tau, x, t, dxdt);
}

int orc_shape(const orc_robot *rb, const double *state, int cap_pts, double *t, double *p,
              double *R, orc_fk_out *out) {
  if (rb->n_tendons < 0 || rb->n_tendons > ORC_MAX_TENDONS) return -2;
  Shape s;
  robot_shape(*rb, state, s);
  if (out) *out = s.out;
  int n = (int)s.t.size();
  if (n > cap_pts) return -1;
  for (int i = 0; i < n; i++) {
    if (t) t[i] = s.t[i];
    if (p) for (int k = 0; k < 3; k++) p[3 * i + k] = s.p[i][k];
    if (R)  // column-major like Eigen::Matrix3d storage
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R[9 * i + r + 3 * c] = s.R[i].m[r][c];
  }
  return n;
}

// ---- finite-difference tip Jacobians, sequential like the reference --------
namespace {
// robot.forward_kinematics(state).back()
void fk_tip(const orc_robot &rb, const double *state, double *tip) {
  Shape s;
  robot_shape(rb, state, s);
  for (int k = 0; k < 3; k++) tip[k] = s.p.back()[k];
}
// fk_wrap of tip-control/tip_control.cpp:92-122 (what levmar differentiates)
void fk_wrap(const orc_robot &rb, const double *p, int m, double *x) {
  if (rb.enable_retraction && p[m - 1] > rb.L) {
    x[0] = 0.0; x[1] = 0.0; x[2] = rb.L - p[m - 1];
    return;
  }
  fk_tip(rb, p, x);
}
}  // namespace

void orc_tip_jacobian(const orc_robot *rb, const double *state, int mode, double delta, double *tip,
                      double *J) {
  const int m = orc_state_size(rb), n = 3;
  std::vector<double> p(state, state + m);
  double hx[3], hxx[3], hxm[3], hxp[3];
  if (mode == 0) {
    // tip_control::Jacobian, tip-control/tip_control.cpp:243-265 (ps = fk(state).back())
    fk_tip(*rb, p.data(), hx);
    for (int i = 0; i < m; i++) {
      std::vector<double> tau2 = p;
      tau2[i] = p[i] + delta;
      fk_tip(*rb, tau2.data(), hxx);
      for (int j = 0; j < 3; j++) J[j * m + i] = (hxx[j] - hx[j]) / delta;
    }
  } else if (mode == 1) {
    // dlevmar_fdif_forw_jac_approx, 3rdparty/levmar-2.6/misc_core.c:137-172
    fk_wrap(*rb, p.data(), m, hx);
    for (int j = 0; j < m; ++j) {
      double d = 1E-04 * p[j];
      d = std::fabs(d);
      if (d < delta) d = delta;
      double tmp = p[j];
      p[j] += d;
      fk_wrap(*rb, p.data(), m, hxx);
      p[j] = tmp;
      d = 1.0 / d;
      for (int i = 0; i < n; ++i) J[i * m + j] = (hxx[i] - hx[i]) * d;
    }
  } else {
    // dlevmar_fdif_cent_jac_approx, misc_core.c:175-211
    fk_wrap(*rb, p.data(), m, hx);
    for (int j = 0; j < m; ++j) {
      double d = 1E-04 * p[j];
      d = std::fabs(d);
      if (d < delta) d = delta;
      double tmp = p[j];
      p[j] -= d;
      fk_wrap(*rb, p.data(), m, hxm);
      p[j] = tmp + d;
      fk_wrap(*rb, p.data(), m, hxp);
      p[j] = tmp;
      d = 0.5 / d;
      for (int i = 0; i < n; ++i) J[i * m + j] = (hxp[i] - hxm[i]) * d;
    }
  }
  if (tip) for (int k = 0; k < 3; k++) tip[k] = hx[k];
}

void orc_home_lengths(const orc_robot *rb, double s_start, double *L_i) {
  home_lengths(*rb, s_start, L_i);
}

int orc_collides_self(const double *p, int npts, double r) {
  std::vector<V3> pts(npts);
  for (int i = 0; i < npts; i++) pts[i] = mk(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
  return collides_self(pts.data(), npts, r) ? 1 : 0;
}

void orc_closest_st_segment(const double *A, const double *B, const double *C, const double *D,
                            double *s, double *t) {
  auto st = closest_st_segment(mk(A[0], A[1], A[2]), mk(B[0], B[1], B[2]), mk(C[0], C[1], C[2]),
                               mk(D[0], D[1], D[2]));
  *s = st.first;
  *t = st.second;
}

int orc_segment_aabox_intersect(const double *A, const double *B, const double *C, const double *D) {
  return segment_aabox_intersect(mk(A[0], A[1], A[2]), mk(B[0], B[1], B[2]), mk(C[0], C[1], C[2]),
                                 mk(D[0], D[1], D[2])) ? 1 : 0;
}

uint32_t orc_validity_flags(const orc_robot *rb, const double *state, const orc_fk_out *fk,
                            const double *p) {
  std::vector<V3> pts(fk->npts);
  for (int i = 0; i < fk->npts; i++) pts[i] = mk(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
  return validity_flags(*rb, state, *fk, pts.data());
}

void orc_fk_batch(const orc_robot *rb, const double *states, int64_t n, int cap_pts, double *p,
                  int32_t *npts, double *L_i, double *tip, uint32_t *flags, int32_t *iters,
                  int32_t *nsteps, int nthreads) {
  const int S = orc_state_size(rb);
  const int N = rb->n_tendons;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16)
  for (int64_t i = 0; i < n; i++) {
    Shape s;
    robot_shape(*rb, states + i * S, s);
    int np = std::min((int)s.p.size(), cap_pts);
    if (p)
      for (int k = 0; k < np; k++)
        for (int c = 0; c < 3; c++) p[(i * cap_pts + k) * 3 + c] = s.p[k][c];
    if (npts) npts[i] = (int)s.p.size();
    if (L_i) for (int j = 0; j < N; j++) L_i[i * N + j] = s.out.L_i[j];
    if (tip) for (int c = 0; c < 3; c++) tip[i * 3 + c] = s.p.back()[c];
    if (flags) flags[i] = validity_flags(*rb, states + i * S, s.out, s.p.data());
    if (iters) iters[i] = s.out.iters;
    if (nsteps) nsteps[i] = s.out.nsteps;
  }
}

// ---- octree ----
orc_octree *orc_octree_new(const orc_grid *g) { return octree_from_grid(g); }
orc_octree *orc_octree_copy(const orc_octree *t) {
  orc_octree *r = new orc_octree(*t);
  r->root = node_copy(t->root);
  return r;
}
void orc_octree_free(orc_octree *t) {
  if (!t) return;
  node_free(t->root);
  delete t;
}
void orc_octree_clear(orc_octree *t) {
  node_free(t->root);
  t->root = new OrcNode();
}
uint64_t orc_octree_block(const orc_octree *t, int bx, int by, int bz) {
  return node_block(t->root, t->Ng / 4, bx, by, bz);
}
void orc_octree_set_block(orc_octree *t, int bx, int by, int bz, uint64_t v) {
  node_set_block(t->root, t->Ng / 4, bx, by, bz, v);
}
uint64_t orc_octree_union_block(orc_octree *t, int bx, int by, int bz, uint64_t v) {
  if (v) return node_union_block(t->root, t->Ng / 4, bx, by, bz, v);  // VoxelOctree.cpp:228-238
  return orc_octree_block(t, bx, by, bz);
}
int64_t orc_octree_nblocks(const orc_octree *t) {
  return node_nblocks(t->root, t->Ng / 4);
}
int64_t orc_octree_ncells(const orc_octree *t) {
  int64_t count = 0;
  node_visit_leaves(t->root, t->Ng / 4, 0, 0, 0,
                    [&](int, int, int, uint64_t b) { count += __builtin_popcountll(b); });
  return count;
}
void orc_octree_add_line(orc_octree *t, const double *a, const double *b) {
  add_line(t, mk(a[0], a[1], a[2]), mk(b[0], b[1], b[2]));
}
void orc_octree_add_piecewise_line(orc_octree *t, const double *pts, int npts) {
  for (int i = 1; i < npts; i++)
    add_line(t, mk(pts[3 * i - 3], pts[3 * i - 2], pts[3 * i - 1]),
             mk(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
}
void orc_octree_add_voxels(orc_octree *t, const orc_octree *other) {
  node_union_tree(t->root, other->root, t->Ng / 4);
}
int orc_octree_collides(const orc_octree *a, const orc_octree *b) {
  if (a->Ng != b->Ng) return -1;  // check_dims would throw std::invalid_argument
  return node_collides(a->root, b->root, a->Ng / 4) ? 1 : 0;
}
int64_t orc_octree_export(const orc_octree *t, int64_t cap, uint8_t *bxyz, uint64_t *bits) {
  int64_t n = 0;
  node_visit_leaves(t->root, t->Ng / 4, 0, 0, 0, [&](int bx, int by, int bz, uint64_t b) {
    if (n < cap) {
      if (bxyz) { bxyz[3 * n] = (uint8_t)bx; bxyz[3 * n + 1] = (uint8_t)by; bxyz[3 * n + 2] = (uint8_t)bz; }
      if (bits) bits[n] = b;
    }
    n++;
  });
  return n;
}
int orc_find_cell(const orc_grid *g, const double *p, int64_t *cell) {
  orc_octree *t = octree_from_grid(g);
  long c[3] = {0, 0, 0};
  bool ok = find_cell(t, mk(p[0], p[1], p[2]), c);
  cell[0] = c[0]; cell[1] = c[1]; cell[2] = c[2];
  orc_octree_free(t);
  return ok ? 0 : 1;
}

// add_sphere / add_capsule -- collision/VoxelOctree.cpp:434-515 (used only to build
// synthetic obstacle environments for tests and benches)
}  // extern "C"
static void nearest_block_idx(const orc_octree *t, double x, double y, double z, int *b) {
  int ix = (int)((x - t->xmin) / t->dx), iy = (int)((y - t->ymin) / t->dy),
      iz = (int)((z - t->zmin) / t->dz);
  int Nb = t->Ng / 4;
  b[0] = std::min(Nb - 1, std::max(0, ix / 4));
  b[1] = std::min(Nb - 1, std::max(0, iy / 4));
  b[2] = std::min(Nb - 1, std::max(0, iz / 4));
}
static void add_point(orc_octree *t, const V3 &p) {
  if (!(t->xmin <= p[0] && p[0] <= t->xmax && t->ymin <= p[1] && p[1] <= t->ymax &&
        t->zmin <= p[2] && p[2] <= t->zmax))
    return;
  int ix = (int)((p[0] - t->xmin) / t->dx), iy = (int)((p[1] - t->ymin) / t->dy),
      iz = (int)((p[2] - t->zmin) / t->dz);
  ix = std::min(t->Ng - 1, std::max(0, ix));
  iy = std::min(t->Ng - 1, std::max(0, iy));
  iz = std::min(t->Ng - 1, std::max(0, iz));
  set_cell(t, ix, iy, iz);
}
template <typename Pred>
static void add_region(orc_octree *t, const V3 &ll, const V3 &tr, const Pred &inside) {
  int bl[3], bt[3];
  nearest_block_idx(t, ll[0], ll[1], ll[2], bl);
  nearest_block_idx(t, tr[0], tr[1], tr[2], bt);
  for (int bx = bl[0]; bx <= bt[0]; bx++)
    for (int by = bl[1]; by <= bt[1]; by++)
      for (int bz = bl[2]; bz <= bt[2]; bz++) {
        uint64_t b = 0;
        for (int i = 0; i < 4; i++)
          for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) {
              V3 c = mk(t->xmin + t->dx * ((bx << 2) + i + 0.5), t->ymin + t->dy * ((by << 2) + j + 0.5),
                        t->zmin + t->dz * ((bz << 2) + k + 0.5));
              if (inside(c)) b |= bitmask(i, j, k);
            }
        if (b) node_union_block(t->root, t->Ng / 4, bx, by, bz, b);
      }
}
extern "C" {
// VoxelOctree::add(Point) -> add_point (VoxelOctree.cpp:319-323): the cell of a point inside the (inclusive) limits
void orc_octree_add_point(orc_octree *t, const double *p) { add_point(t, mk(p[0], p[1], p[2])); }
void orc_octree_add_sphere(orc_octree *t, const double *c, double r) {
  V3 ctr = mk(c[0], c[1], c[2]);
  add_point(t, ctr);
  add_region(t, ctr - mk(r, r, r), ctr + mk(r, r, r), [&](const V3 &p) {
    V3 d = ctr - p;
    return dot(d, d) <= r * r;
  });
}
void orc_octree_add_capsule(orc_octree *t, const double *pa, const double *pb, double r) {
  V3 a = mk(pa[0], pa[1], pa[2]), b = mk(pb[0], pb[1], pb[2]);
  add_point(t, a);
  add_point(t, b);
  V3 ll = mk(std::min(a[0], b[0]) - r, std::min(a[1], b[1]) - r, std::min(a[2], b[2]) - r);
  V3 tr = mk(std::max(a[0], b[0]) + r, std::max(a[1], b[1]) + r, std::max(a[2], b[2]) + r);
  add_region(t, ll, tr, [&](const V3 &p) {
    // closest_t_segment -- collision_primitives.h:33-49
    const double eps = std::numeric_limits<double>::epsilon();
    V3 diff = b - a;
    double d2 = dot(diff, diff);
    double tt = (d2 <= eps * eps) ? 0.0 : dot(diff, p - a) / d2;
    tt = std::max(0.0, std::min(1.0, tt));
    V3 closest = interp_pt(a, b, tt);
    V3 d = closest - p;
    return dot(d, d) <= r * r;
  });
}

// ---- environment preparation ----
// dilate_one_impl / dilate_6neighbor / dilate_27neighbor / dilate_sphere --
// collision/VoxelOctree.cpp:693-952, restated literally: depth-limited DFS inside the 3x3x3 block
// neighbourhood (12^3 cells), at most four dilation steps per pass over a snapshot of the tree.
}  // extern "C"
namespace {
typedef void (*NeighborFn)(int x, int y, int z, int out[27][3], int *count);
void neighbors6(int x, int y, int z, int out[27][3], int *count) {
  const int d[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
  for (int i = 0; i < 6; i++) { out[i][0] = x + d[i][0]; out[i][1] = y + d[i][1]; out[i][2] = z + d[i][2]; }
  *count = 6;
}
void neighbors27(int x, int y, int z, int out[27][3], int *count) {
  // the reference's list (VoxelOctree.cpp:790-818) names (x+1,y+1,z+1) twice and never
  // (x-1,y+1,z+1): reproduced as written
  const int d[27][3] = {
      {0, 0, 0},  {-1, 0, 0},  {1, 0, 0},  {0, -1, 0},  {-1, -1, 0},  {1, -1, 0},  {0, 1, 0},  {-1, 1, 0},  {1, 1, 0},
      {0, 0, -1}, {-1, 0, -1}, {1, 0, -1}, {0, -1, -1}, {-1, -1, -1}, {1, -1, -1}, {0, 1, -1}, {-1, 1, -1}, {1, 1, -1},
      {0, 0, 1},  {-1, 0, 1},  {1, 0, 1},  {0, -1, 1},  {-1, -1, 1},  {1, -1, 1},  {0, 1, 1},  {1, 1, 1},   {1, 1, 1}};
  for (int i = 0; i < 27; i++) { out[i][0] = x + d[i][0]; out[i][1] = y + d[i][1]; out[i][2] = z + d[i][2]; }
  *count = 27;
}
struct DepthSet {
  uint8_t depths[12][12][12];
  uint64_t voxels[3][3][3];
  bool in(int x, int y, int z, int d) const { return d == 0 || d <= depths[x][y][z]; }
  bool add(int x, int y, int z, int d) {
    if (!in(x, y, z, d)) {
      depths[x][y][z] = (uint8_t)d;
      voxels[x / 4][y / 4][z / 4] |= bitmask(x % 4, y % 4, z % 4);
      return true;
    }
    return false;
  }
};
void dfs_visit(DepthSet &vs, NeighborFn nb, int x, int y, int z, int d) {
  if (!vs.add(x, y, z, d)) return;
  int out[27][3], cnt = 0;
  nb(x, y, z, out, &cnt);
  for (int i = 0; i < cnt; i++) dfs_visit(vs, nb, out[i][0], out[i][1], out[i][2], d - 1);
}
void dilate_one_impl(orc_octree *t, NeighborFn nb, int num) {
  if (num <= 0) return;
  const int Nb = t->Ng / 4;
  for (; num > 0; num -= 4) {
    OrcNode *copy = node_copy(t->root);
    node_visit_leaves(copy, Nb, 0, 0, 0, [&](int bx, int by, int bz, uint64_t old_b) {
      const int n = std::min(num, 4);
      DepthSet vs;
      std::memset(&vs, 0, sizeof(vs));
      for (int x = 0; x < 4; ++x)
        for (int y = 0; y < 4; ++y)
          for (int z = 0; z < 4; ++z)
            if (old_b & bitmask(x, y, z)) dfs_visit(vs, nb, x + 4, y + 4, z + 4, n + 1);
      for (int nx = 0; nx < 3; ++nx) {
        const int bxi = bx + nx - 1;
        if (bxi < 0 || bxi >= Nb) continue;
        for (int ny = 0; ny < 3; ++ny) {
          const int byi = by + ny - 1;
          if (byi < 0 || byi >= Nb) continue;
          for (int nz = 0; nz < 3; ++nz) {
            const int bzi = bz + nz - 1;
            if (bzi < 0 || bzi >= Nb) continue;
            if (vs.voxels[nx][ny][nz]) node_union_block(t->root, Nb, bxi, byi, bzi, vs.voxels[nx][ny][nz]);
          }
        }
      }
    });
    node_free(copy);
  }
}
// remove_interior_6neighbor / _27neighbor -- collision/VoxelOctree.cpp:533-689
void remove_interior(orc_octree *t, bool diagonal) {
  const int Nb = t->Ng / 4;
  OrcNode *copy = node_copy(t->root);
  const uint64_t full = ~uint64_t(0);
  node_visit_leaves(copy, Nb, 0, 0, 0, [&](int bx, int by, int bz, uint64_t old_b) {
    uint64_t nbh[3][3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) {
          const int x = bx - 1 + i, y = by - 1 + j, z = bz - 1 + k;
          if (x < 0 || Nb - 1 < x || y < 0 || Nb - 1 < y || z < 0 || Nb - 1 < z) nbh[i][j][k] = full;
          else if (i == 1 && j == 1 && k == 1) nbh[i][j][k] = old_b;
          else nbh[i][j][k] = node_block(copy, Nb, x, y, z);
        }
    auto neighbor = [&](int ix, int iy, int iz) -> bool {
      int nx = 1, ny = 1, nz = 1;
      if (ix == -1) { ix = 3; nx = 0; } else if (ix == 4) { ix = 0; nx = 2; }
      if (iy == -1) { iy = 3; ny = 0; } else if (iy == 4) { iy = 0; ny = 2; }
      if (iz == -1) { iz = 3; nz = 0; } else if (iz == 4) { iz = 0; nz = 2; }
      return (nbh[nx][ny][nz] & bitmask(ix, iy, iz)) != 0;
    };
    uint64_t new_b = old_b;
    for (int ix = 0; ix < 4; ix++)
      for (int iy = 0; iy < 4; iy++)
        for (int iz = 0; iz < 4; iz++) {
          bool interior = true;
          if (diagonal) {
            for (int i = -1; i <= 1 && interior; ++i)
              for (int j = -1; j <= 1 && interior; ++j)
                for (int k = -1; k <= 1 && interior; ++k)
                  if (!neighbor(ix + i, iy + j, iz + k)) interior = false;
          } else {
            interior = neighbor(ix, iy, iz) && neighbor(ix - 1, iy, iz) && neighbor(ix + 1, iy, iz) &&
                       neighbor(ix, iy - 1, iz) && neighbor(ix, iy + 1, iz) && neighbor(ix, iy, iz - 1) &&
                       neighbor(ix, iy, iz + 1);
          }
          if (interior) new_b &= ~bitmask(ix, iy, iz);
        }
    node_set_block(t->root, Nb, bx, by, bz, new_b);
  });
  node_free(copy);
}
}  // namespace
extern "C" {
void orc_octree_dilate_6neighbor(orc_octree *t, int num) { dilate_one_impl(t, neighbors6, num); }
void orc_octree_dilate_27neighbor(orc_octree *t, int num) { dilate_one_impl(t, neighbors27, num); }
void orc_octree_dilate_sphere(orc_octree *t, double r) {  // VoxelOctree.cpp:949-951 ("attempt #3")
  dilate_one_impl(t, neighbors6, int(std::round(r / std::min(t->dx, std::min(t->dy, t->dz)))));
}
void orc_octree_remove_interior(orc_octree *t, int keep_diagonal) { remove_interior(t, keep_diagonal != 0); }

// ---- OMPL-side restatement ----
// Problem.cpp:101-163: tension RealVector (weight 1), rotation SO2, retraction RealVector.
// OMPL 1.5 (documented behaviour): StateSpace::validSegmentCount =
//   factor(1) * ceil(distance / longestValidSegment), longestValidSegment =
//   maximumExtent * fraction; CompoundStateSpace = max over subspaces.
uint32_t orc_valid_segment_count(const orc_robot *rb, const orc_space *sp, const double *a,
                                 const double *b) {
  const int N = rb->n_tendons;
  double ext2 = 0;
  for (int i = 0; i < N; i++) ext2 += rb->max_tension[i] * rb->max_tension[i];
  const double tendon_extent = std::sqrt(ext2);
  const double len_t = tendon_extent * (sp->min_tension_change / tendon_extent);
  double d2 = 0;
  for (int i = 0; i < N; i++) d2 += (a[i] - b[i]) * (a[i] - b[i]);
  unsigned sc = (unsigned)std::ceil(std::sqrt(d2) / len_t);
  int idx = N;
  if (rb->enable_rotation) {
    const double len_r = M_PI * (sp->min_rotation_change / (2 * M_PI));
    double d = std::fabs(a[idx] - b[idx]);
    d = (d > M_PI) ? 2.0 * M_PI - d : d;
    sc = std::max(sc, (unsigned)std::ceil(d / len_r));
    idx++;
  }
  if (rb->enable_retraction) {
    const double len_s = rb->L * std::min(0.01, sp->min_retraction_change / rb->L);
    double d = std::sqrt((a[idx] - b[idx]) * (a[idx] - b[idx]));
    sc = std::max(sc, (unsigned)std::ceil(d / len_s));
  }
  return sc;
}

// OMPL compound interpolate: RealVector linear; SO2 shortest arc with wrap
void orc_interpolate(const orc_robot *rb, const double *a, const double *b, double t,
                     double *out) {
  const int N = rb->n_tendons;
  for (int i = 0; i < N; i++) out[i] = a[i] + (b[i] - a[i]) * t;
  int idx = N;
  if (rb->enable_rotation) {
    double diff = b[idx] - a[idx];
    if (std::fabs(diff) <= M_PI) {
      out[idx] = a[idx] + diff * t;
    } else {
      if (diff > 0.0) diff = 2.0 * M_PI - diff;
      else diff = -2.0 * M_PI - diff;
      double v = a[idx] - diff * t;
      if (v > M_PI) v -= 2.0 * M_PI;
      else if (v < -M_PI) v += 2.0 * M_PI;
      out[idx] = v;
    }
    idx++;
  }
  if (rb->enable_retraction) out[idx] = a[idx] + (b[idx] - a[idx]) * t;
}

void orc_voxelize_shape(const orc_grid *g, const double *p, int npts, orc_octree *out) {
  (void)g;
  V3 prev = mk(0, 0, 0);
  for (int i = 0; i < npts; i++) {
    V3 q = rotate_point(out->inv_rot, mk(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
    if (i > 0) add_line(out, prev, q);
    prev = q;
  }
}

// VoxelEnvironment::voxelize_valid_backbone_motion -- motion-planning/VoxelEnvironment.cpp:207-444
// driven as VoxelBackboneMotionValidator::generic_voxelize (.cpp:19-74) drives it.
void orc_voxelize_edge(const orc_robot *rb, const orc_grid *g, const orc_space *sp,
                       const double *a, const double *b, const orc_octree *env,
                       orc_octree *voxels, orc_edge_out *info) {
  const int S = orc_state_size(rb);
  std::memset(info, 0, sizeof(*info));
  const unsigned nseg = orc_valid_segment_count(rb, sp, a, b);
  const double rel_threshold = 1.0 / double(nseg);

  struct Sample {
    double t;
    std::vector<V3> shape;  // rotated points
    bool is_valid;
  };
  std::vector<Sample> fks;
  double first_invalid_t = 10.0;
  bool ood = false;

  auto add_fk = [&](double t, const double *config) {
    size_t i = fks.size();
    Shape s;
    robot_shape(*rb, config, s);
    bool is_valid = (validity_flags(*rb, config, s.out, s.p.data()) == 0);
    if (is_valid && env) {  // voxelize_until_invalid: _vc->collides(shape)
      orc_octree *tmp = octree_from_grid(g);
      std::vector<double> flat(3 * s.p.size());
      for (size_t k = 0; k < s.p.size(); k++)
        for (int c = 0; c < 3; c++) flat[3 * k + c] = s.p[k][c];
      orc_voxelize_shape(g, flat.data(), (int)s.p.size(), tmp);
      if (orc_octree_collides(env, tmp) == 1) is_valid = false;
      orc_octree_free(tmp);
    }
    if (!is_valid && t < first_invalid_t) first_invalid_t = t;
    for (auto &p : s.p) p = rotate_point(g->inv_rot, p);
    fks.push_back(Sample{t, std::move(s.p), is_valid});
    return i;
  };
  std::vector<double> current(a, a + S);
  auto interpolate = [&](double t) {
    orc_interpolate(rb, a, b, t, current.data());
    return add_fk(t, current.data());
  };

  std::pair<size_t, size_t> motion;
  motion.first = add_fk(0.0, a);
  motion.second = add_fk(1.0, b);

  auto should_subdivide = [&](size_t ia, size_t ib) {
    const Sample &sa = fks[ia], &sb = fks[ib];
    if (!sa.is_valid) return false;
    if (sa.shape.size() + 1 < sb.shape.size() || sa.shape.size() > sb.shape.size() + 1) return true;
    int P = int(std::min(sa.shape.size(), sb.shape.size()));
    for (int i = P - 1; i >= 0; i--) {
      long s[3], e[3];
      if (!find_cell(voxels, sa.shape[i], s) || !find_cell(voxels, sb.shape[i], e)) {
        ood = true;  // reference: std::domain_error escapes
        return false;
      }
      long dx = std::labs(s[0] - e[0]), dy = std::labs(s[1] - e[1]), dz = std::labs(s[2] - e[2]);
      if (dx > 1 || dy > 1 || dz > 1) return true;
    }
    return false;
  };

  std::vector<std::pair<size_t, size_t>> frontier;  // used as a stack
  if (should_subdivide(motion.first, motion.second)) frontier.push_back(motion);
  while (!frontier.empty()) {
    auto interval = frontier.back();
    frontier.pop_back();
    const double t_a = fks[interval.first].t, t_b = fks[interval.second].t;
    if ((t_b - t_a) <= rel_threshold) continue;
    if (first_invalid_t <= t_a) continue;
    double mid_interp = (t_a + t_b) / 2;
    size_t mid_idx = interpolate(mid_interp);
    if (should_subdivide(mid_idx, interval.second)) frontier.emplace_back(mid_idx, interval.second);
    if (should_subdivide(interval.first, mid_idx)) frontier.emplace_back(interval.first, mid_idx);
  }

  double last_valid_t = 0.0;
  size_t last_valid_idx = 0;
  for (size_t i = fks.size(); i-- > 0;) {
    Sample &s = fks[i];
    if (s.t < first_invalid_t) {
      for (size_t k = 1; k < s.shape.size(); k++) add_line(voxels, s.shape[k - 1], s.shape[k]);
      if (last_valid_t < s.t) { last_valid_t = s.t; last_valid_idx = i; }
    }
  }
  const double t = fks[last_valid_idx].t;
  orc_interpolate(rb, a, b, t, current.data());
  info->is_fully_valid = (5.0 < first_invalid_t) ? 1 : 0;
  info->nsamples = (int32_t)fks.size();
  info->out_of_domain = ood ? 1 : 0;
  info->t = t;
  for (int i = 0; i < S; i++) info->last_valid[i] = current[i];
}

// ---- set store + batch drivers ----
}  // extern "C"

struct orc_setstore {
  orc_grid grid;
  std::vector<orc_octree *> sets;
};

extern "C" {

orc_setstore *orc_setstore_new(const orc_grid *g, int64_t n) {
  orc_setstore *s = new orc_setstore();
  s->grid = *g;
  s->sets.resize(n);
  for (auto &t : s->sets) t = octree_from_grid(g);
  return s;
}
void orc_setstore_free(orc_setstore *s) {
  if (!s) return;
  for (auto t : s->sets) orc_octree_free(t);
  delete s;
}
int64_t orc_setstore_size(const orc_setstore *s) { return (int64_t)s->sets.size(); }
orc_octree *orc_setstore_get(orc_setstore *s, int64_t i) { return s->sets[i]; }
int64_t orc_setstore_total_blocks(const orc_setstore *s) {
  int64_t n = 0;
  for (auto t : s->sets) n += orc_octree_export(t, 0, nullptr, nullptr);
  return n;
}
uint32_t orc_morton_key(int bx, int by, int bz, int Nb) {
  uint32_t key = 0;
  for (int l = 0; (1 << l) < Nb; l++)
    key |= (uint32_t((bx >> l) & 1) << (3 * l + 2)) | (uint32_t((by >> l) & 1) << (3 * l + 1)) |
           (uint32_t((bz >> l) & 1) << (3 * l));
  return key;
}
void orc_setstore_export(const orc_setstore *s, uint64_t *offsets, uint32_t *keys,
                         uint64_t *bits) {
  uint64_t off = 0;
  const int Nb = s->grid.Ng / 4;
  for (size_t i = 0; i < s->sets.size(); i++) {
    offsets[i] = off;
    node_visit_leaves(s->sets[i]->root, Nb, 0, 0, 0, [&](int bx, int by, int bz, uint64_t b) {
      keys[off] = orc_morton_key(bx, by, bz, Nb);
      bits[off] = b;
      off++;
    });
  }
  offsets[s->sets.size()] = off;
}

// VoxelCachedLazyPRM::voxelizeVertex over all vertices (VoxelCachedLazyPRM.cpp:1704-1712,2803-2837)
void orc_voxelize_vertices_batch(const orc_robot *rb, const orc_grid *g, const double *states,
                                 int64_t n, orc_setstore *out, uint32_t *flags, int nthreads) {
  const int S = orc_state_size(rb);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 100)
  for (int64_t i = 0; i < n; i++) {
    Shape s;
    robot_shape(*rb, states + i * S, s);
    uint32_t f = validity_flags(*rb, states + i * S, s.out, s.p.data());
    if (flags) flags[i] = f;
    orc_octree_clear(out->sets[i]);
    if (f == 0) {
      std::vector<double> flat(3 * s.p.size());
      for (size_t k = 0; k < s.p.size(); k++)
        for (int c = 0; c < 3; c++) flat[3 * k + c] = s.p[k][c];
      orc_voxelize_shape(g, flat.data(), (int)s.p.size(), out->sets[i]);
    }
  }
}

// VoxelCachedLazyPRM::voxelizeEdge over all edges (VoxelCachedLazyPRM.cpp:1520-1542,2879-2902)
void orc_voxelize_edges_batch(const orc_robot *rb, const orc_grid *g, const orc_space *sp,
                              const double *a, const double *b, int64_t n, orc_setstore *out,
                              uint32_t *flags, double *t_last, int32_t *nsamples,
                              int nthreads) {
  const int S = orc_state_size(rb);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
  for (int64_t i = 0; i < n; i++) {
    orc_edge_out info;
    orc_octree_clear(out->sets[i]);
    orc_voxelize_edge(rb, g, sp, a + i * S, b + i * S, nullptr, out->sets[i], &info);
    uint32_t f = 0;
    if (!info.is_fully_valid) f |= ORC_FLAG_PARTIAL;
    if (info.out_of_domain) f |= ORC_FLAG_OUT_OF_DOMAIN;
    if (flags) flags[i] = f;
    if (t_last) t_last[i] = info.t;
    if (nsamples) nsamples[i] = info.nsamples;
  }
}

void orc_check_sets_batch(const orc_setstore *s, const orc_octree *env, int64_t begin,
                          int64_t end, uint8_t *verdict, int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 100)
  for (int64_t i = begin; i < end; i++)
    verdict[i - begin] = (uint8_t)(node_collides(env->root, s->sets[i]->root, env->Ng / 4) ? 1 : 0);
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
