// capsule_pair.h -- arithmetic of the exact stage of collision::collides_self(CapsuleSequence)
// (collision/collision.cpp:6-46): the capsule-pair test with closest_st_segment (collision_primitives.cpp:10-102,
// collision.hxx:65-68,102-108), the per-point segment length and the index gap below which the reference's own
// arc-length rule skips a pair.  Shared by the device kernel (selfcol.cu, built with -fmad=false) and a host harness
// (tests/cpp/test_capsule_pair_host.cpp, built with -ffp-contract=off) that checks them against the oracle without a
// GPU.  Every floating-point operation is in the reference's order.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define CP_HD static __device__
#define CP_HDI static __device__ __forceinline__
#else
#define CP_HD inline
#define CP_HDI inline
#endif

constexpr int SC_CHUNK = 8;

struct P3 {
  double x, y, z;
};
CP_HDI P3 sub3(const P3 &a, const P3 &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
CP_HDI double dot3(const P3 &a, const P3 &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
CP_HDI double bound01(double t) { return fmax(0.0, fmin(1.0, t)); }

// closest_st_segment -- collision/collision_primitives.cpp:10-102
CP_HD void closest_st(const P3 &A, const P3 &B, const P3 &C, const P3 &D, double &s, double &t) {
  const double eps = 2.220446049250313e-16;
  const double eps2 = eps * eps;
  const P3 AB = sub3(B, A), CD = sub3(D, C);
  const double a = dot3(AB, AB), c = dot3(CD, CD);
  if (a <= eps2) {
    s = 0.0;
    t = (c <= eps2) ? 0.0 : bound01(dot3(CD, sub3(A, C)) / c);
    return;
  }
  if (c <= eps2) {
    s = bound01(dot3(AB, sub3(C, A)) / a);
    t = 0.0;
    return;
  }
  const P3 AC = sub3(C, A);
  const double b = dot3(AB, CD), d = dot3(AC, AB), e = dot3(AC, CD);
  const double denom = fmax(0.0, a * c - b * b);
  if (denom <= eps2) {
    double tt = dot3(CD, sub3(A, C)) / c;
    if (0.0 <= tt && tt <= 1.0) { s = 0.0; t = tt; return; }
    tt = dot3(CD, sub3(B, C)) / c;
    if (0.0 <= tt && tt <= 1.0) { s = 1.0; t = tt; return; }
    double ss = dot3(AB, sub3(C, A)) / a;
    if (0.0 <= ss && ss <= 1.0) { s = ss; t = 0.0; return; }
    const P3 AD = sub3(D, A), BC = sub3(C, B), BD = sub3(D, B);
    const double ac2 = dot3(AC, AC), ad2 = dot3(AD, AD), bc2 = dot3(BC, BC), bd2 = dot3(BD, BD);
    if (ac2 <= ad2 && ac2 <= bc2 && ac2 <= bd2) { s = 0.0; t = 0.0; return; }
    if (ad2 <= bc2 && ad2 <= bd2) { s = 0.0; t = 1.0; return; }
    if (bc2 <= bd2) { s = 1.0; t = 0.0; return; }
    s = 1.0; t = 1.0;
    return;
  }
  const double ss = (c * d - b * e) / denom;
  const double tt = (b * d - a * e) / denom;
  if (0.0 <= tt && tt <= 1.0) { s = bound01(ss); t = tt; return; }
  if (tt < 0.0) { s = bound01(-c / a); t = 0.0; return; }
  s = bound01((b - c) / a);
  t = 1.0;
}

// collides(Capsule, Capsule) -- collision/collision.hxx:102-108
CP_HD bool capsules_collide(const P3 &a0, const P3 &a1, const P3 &b0, const P3 &b1, double rr) {
  double s, t;
  closest_st(a0, a1, b0, b1, s, t);
  const P3 dA = sub3(a1, a0), dB = sub3(b1, b0);
  const P3 c1 = {a0.x + dA.x * s, a0.y + dA.y * s, a0.z + dA.z * s};
  const P3 c2 = {b0.x + dB.x * t, b0.y + dB.y * t, b0.z + dB.z * t};
  const P3 diff = sub3(c1, c2);
  return dot3(diff, diff) <= (rr * rr);
}


// segment length into point i (the reference's per-point distance, collision.cpp:22-29); 0 for i = 0
CP_HDI double sc_segment_len(const double *px, int i) {
  double len = 0.0;
  if (i > 0) {
    const double dx = px[3 * i] - px[3 * i - 3], dy = px[3 * i + 1] - px[3 * i - 2],
                 dz = px[3 * i + 2] - px[3 * i - 1];
    len = sqrt((dx * dx + dy * dy) + dz * dz);
  }
  return len;
}

// Index pre-filter, exact: acc[b] - acc[a+1] is a sum of (b - a - 1) segment lengths, so it is
// below 3r whenever (b - a - 1) * maxlen is (with a 1e-9 safety factor for the rounding of the
// running sum): such pairs are skipped by the reference's own rule (collision.cpp:37-39).
// Returns the smallest b - a - 1 that can survive the rule.
CP_HDI int sc_min_gap(double maxlen, double dist_to_consider) {
  const double safe = dist_to_consider * (1.0 - 1e-9);
  return (maxlen > 0.0) ? (int)fmin(1e6, floor(safe / maxlen)) : 1000000;
}
