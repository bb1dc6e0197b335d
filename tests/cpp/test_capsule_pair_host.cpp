// Host check of the arithmetic the exact stage of the device self-collision test runs
// (interactive-rate-tendons_b200/csrc/capsule_pair.h: closest_st, capsules_collide, sc_segment_len, sc_min_gap -- the
// file selfcol.cu includes) against the oracle's closest_st_segment / collides_self, which are pinned by the
// reference's own text (collision/collision.cpp:6-46, collision_primitives.cpp:10-102).
//  (1) closest_st on segment pairs of every kind -- random, zero-length, parallel, collinear and overlapping,
//      nearly parallel, touching end points: (s, t) bit-equal to the oracle's;
//  (2) whole backbones: the kernel's decision procedure composed serially from the shared functions -- bounding
//      spheres over chunks of 8 capsules, the chunk-pair test with the index-gap rule, the reference's loop bounds
//      and arc-length rule over the sequential running sum, the capsule test (the chunk pruning itself is restated
//      here from selfcol.cu's self_collision_kernel, statement by statement) -- against orc_collides_self on random
//      walks, arcs, spirals, hairpins at 2r +- ulps and corners around the 3r rule: same verdict on every shape, i.e.
//      the pruning never drops a pair the reference would have found.
//  (3) the FP32 filter in front of the exact stage (self_collision_filter_kernel: the turning-angle early-out and the
//      single-precision midpoint test with outward margins), restated here statement by statement with the host's
//      sqrtf / atan2f (they differ from the device's by ulps; the margins are 1e-5 .. 1e-3 relative): a backbone it
//      lets go never collides by the oracle.
// Compile with -ffp-contract=off (the device file is built with -fmad=false).  The oracle is linked as the checker
// only.  Built and run by tests/test_abi_and_host.py.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "../../interactive-rate-tendons_b200/csrc/capsule_pair.h"
#include "../../oracle/tendon_oracle.h"

static std::mt19937_64 gen(20220803);
static double U(double a, double b) { return std::uniform_real_distribution<double>(a, b)(gen); }
static double Nrm() { return std::normal_distribution<double>(0.0, 1.0)(gen); }

// ---- (1) closest_st ------------------------------------------------------------------------------------------
static long long check_closest(long long n, long long *n_parallel) {
  long long bad = 0;
  for (long long i = 0; i < n; i++) {
    double A[3], B[3], C[3], D[3];
    const double sc = std::pow(10.0, U(-4, 0));
    for (int c = 0; c < 3; c++) { A[c] = U(-1, 1) * sc; B[c] = A[c] + U(-1, 1) * sc; C[c] = U(-1, 1) * sc; D[c] = C[c] + U(-1, 1) * sc; }
    const int kind = (int)(i % 12);
    const double k1 = U(-2, 2), k2 = U(-2, 2);
    switch (kind) {
      case 1: for (int c = 0; c < 3; c++) B[c] = A[c]; break;                                    // AB has zero length
      case 2: for (int c = 0; c < 3; c++) D[c] = C[c]; break;                                    // CD has zero length
      case 3: for (int c = 0; c < 3; c++) { B[c] = A[c]; D[c] = C[c]; } break;                   // both
      case 4: for (int c = 0; c < 3; c++) D[c] = C[c] + k1 * (B[c] - A[c]); break;               // parallel
      case 5: for (int c = 0; c < 3; c++) { C[c] = A[c] + k1 * (B[c] - A[c]); D[c] = A[c] + k2 * (B[c] - A[c]); } break;   // collinear
      case 6: for (int c = 0; c < 3; c++) D[c] = C[c] + k1 * (B[c] - A[c]) + 1e-9 * sc * Nrm(); break;   // nearly parallel
      case 7: for (int c = 0; c < 3; c++) C[c] = B[c]; break;                                    // share an end point
      case 8: for (int c = 0; c < 3; c++) { C[c] = A[c]; D[c] = B[c]; } break;                   // the same segment
      case 9: A[2] = B[2] = 0; C[2] = D[2] = 0.03; break;                                        // two planes
      case 10: for (int c = 0; c < 3; c++) B[c] = A[c] + 1e-17 * Nrm(); break;                   // length below eps
      default: break;
    }
    if (kind >= 4 && kind <= 6) ++*n_parallel;
    double s0, t0, s1, t1;
    orc_closest_st_segment(A, B, C, D, &s0, &t0);
    closest_st(P3{A[0], A[1], A[2]}, P3{B[0], B[1], B[2]}, P3{C[0], C[1], C[2]}, P3{D[0], D[1], D[2]}, s1, t1);
    if (!(s0 == s1 && t0 == t1)) bad++;
  }
  return bad;
}

// ---- (2) the kernel's decision for one backbone, composed serially -----------------------------------------
// (selfcol.cu: self_collision_kernel; lanes become loops, the any / ballot votes become ORs)
static bool kernel_decision(const double *px, int N, double r, long long *pairs_tested) {
  if (N <= 2) return false;
  const double dist_to_consider = 3.0 * r, rr = r + r;
  std::vector<double> acc((size_t)N);
  double maxlen = 0.0;
  for (int i = 0; i < N; i++) { acc[i] = sc_segment_len(px, i); maxlen = fmax(maxlen, acc[i]); }
  const int min_gap = sc_min_gap(maxlen, dist_to_consider);
  const int ncap = N - 1, nchunk = (ncap + SC_CHUNK - 1) / SC_CHUNK;
  std::vector<double> ch((size_t)4 * nchunk);
  for (int c = 0; c < nchunk; c++) {
    const int i0 = c * SC_CHUNK, i1 = i0 + SC_CHUNK < ncap ? i0 + SC_CHUNK : ncap;
    const int im = (i0 + i1) >> 1;
    const double cx = px[3 * im], cy = px[3 * im + 1], cz = px[3 * im + 2];
    double rad2 = 0.0;
    for (int i = i0; i <= i1; i++) {
      const double dx = px[3 * i] - cx, dy = px[3 * i + 1] - cy, dz = px[3 * i + 2] - cz;
      rad2 = fmax(rad2, (dx * dx + dy * dy) + dz * dz);
    }
    ch[4 * c] = cx; ch[4 * c + 1] = cy; ch[4 * c + 2] = cz; ch[4 * c + 3] = sqrt(rad2);
  }
  bool have_acc = false;
  for (int pair = 0; pair < nchunk * nchunk; pair++) {
    const int ca = pair / nchunk, cb = pair - ca * nchunk;
    const int max_gap = (cb + 1) * SC_CHUNK - 1 - ca * SC_CHUNK - 1;
    bool near = false;
    if (cb >= ca && max_gap >= min_gap) {
      const double dx = ch[4 * ca] - ch[4 * cb], dy = ch[4 * ca + 1] - ch[4 * cb + 1], dz = ch[4 * ca + 2] - ch[4 * cb + 2];
      const double reach = (ch[4 * ca + 3] + ch[4 * cb + 3] + rr) * (1.0 + 1e-9) + 1e-12;
      near = ((dx * dx + dy * dy) + dz * dz) <= reach * reach;
    }
    if (!near) continue;
    if (!have_acc) {
      double dsum = 0.0;
      for (int i = 0; i < N; i++) { dsum += acc[i]; acc[i] = dsum; }
      have_acc = true;
    }
    for (int k = 0; k < SC_CHUNK * SC_CHUNK; k++) {
      const int a = ca * SC_CHUNK + (k >> 3), b = cb * SC_CHUNK + (k & 7);
      if (a < N - 3 && b >= a + 2 && b < N - 1) {
        if (!(acc[b] - acc[a + 1] < dist_to_consider)) {
          ++*pairs_tested;
          const P3 A0 = {px[3 * a], px[3 * a + 1], px[3 * a + 2]}, A1 = {px[3 * a + 3], px[3 * a + 4], px[3 * a + 5]};
          const P3 B0 = {px[3 * b], px[3 * b + 1], px[3 * b + 2]}, B1 = {px[3 * b + 3], px[3 * b + 4], px[3 * b + 5]};
          if (capsules_collide(A0, A1, B0, B1, rr)) return true;
        }
      }
    }
  }
  return false;
}

// ---- (3) the FP32 filter, restated serially (selfcol.cu: self_collision_filter_kernel) ------------------------
// returns 0 = left at the turning-angle bound, 1 = no candidate pair, 2 = goes on to the exact stage
static int filter_decision(const double *src, int N, double r) {
  if (N <= 3) return 1;
  const float rr = (float)(2.0 * r) * 1.00001f + 1e-6f;
  const int ncap = N - 1;
  std::vector<float> sx((size_t)ncap), sy((size_t)ncap), sz((size_t)ncap), sw((size_t)ncap);
  float maxhl = 0.0f, tsum = 0.0f, tmax = 0.0f;
  for (int i = 0; i < ncap; i++) {
    const double ax = src[3 * i], ay = src[3 * i + 1], az = src[3 * i + 2];
    const double bx = src[3 * i + 3], by = src[3 * i + 4], bz = src[3 * i + 5];
    const float dx = (float)(bx - ax), dy = (float)(by - ay), dz = (float)(bz - az);
    const float hl = 0.5f * sqrtf(dx * dx + dy * dy + dz * dz) * 1.00001f + 1e-9f;
    sx[i] = (float)(0.5 * (ax + bx)); sy[i] = (float)(0.5 * (ay + by)); sz[i] = (float)(0.5 * (az + bz)); sw[i] = hl;
    maxhl = fmaxf(maxhl, hl);
    if (i + 1 < ncap) {
      const float ex = (float)(src[3 * i + 6] - bx), ey = (float)(src[3 * i + 7] - by), ez = (float)(src[3 * i + 8] - bz);
      const float cx = dy * ez - dz * ey, cy = dz * ex - dx * ez, cz = dx * ey - dy * ex;
      const float th = atan2f(sqrtf(cx * cx + cy * cy + cz * cz), dx * ex + dy * ey + dz * ez) * 1.001f + 1e-6f;
      tsum += th;
      tmax = fmaxf(tmax, th);
    }
  }
  if (0.5f * tsum + tmax < 0.83f) return 0;
  const float safe = (float)(3.0 * r) * 0.9999f;
  const int min_gap = (maxhl > 0.0f) ? (int)fminf(1e6f, floorf(safe / (2.0f * maxhl))) : 1000000;
  const int first = 1 + (min_gap < 1 ? 1 : min_gap);
  for (int a = 0; a < ncap - first; a++) {
    if (a >= N - 3) continue;
    for (int b = a + first; b < ncap; b++) {
      const float dx = sx[a] - sx[b], dy = sy[a] - sy[b], dz = sz[a] - sz[b];
      const float reach = rr + sw[a] + sw[b];
      if (fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= reach * reach) return 2;
    }
  }
  return 1;
}

using Shape = std::vector<double>;   // xyz per point
static void rigid(Shape &s) {
  double q[3][3];
  for (auto &row : q) for (double &v : row) v = Nrm();
  for (int i = 0; i < 3; i++) {      // Gram-Schmidt
    for (int j = 0; j < i; j++) {
      double d = 0;
      for (int c = 0; c < 3; c++) d += q[i][c] * q[j][c];
      for (int c = 0; c < 3; c++) q[i][c] -= d * q[j][c];
    }
    double n = 0;
    for (int c = 0; c < 3; c++) n += q[i][c] * q[i][c];
    n = std::sqrt(n);
    for (int c = 0; c < 3; c++) q[i][c] /= n;
  }
  const double off[3] = {U(-0.15, 0.15), U(-0.15, 0.15), U(-0.15, 0.15)};
  for (size_t k = 0; k < s.size() / 3; k++) {
    const double x = s[3 * k], y = s[3 * k + 1], z = s[3 * k + 2];
    for (int c = 0; c < 3; c++) s[3 * k + c] = q[c][0] * x + q[c][1] * y + q[c][2] * z + off[c];
  }
}

int main() {
  long long n_par = 0;
  const long long n_pairs = 6000000;
  const long long bad_st = check_closest(n_pairs, &n_par);
  std::printf("closest_st: %lld segment pairs (%lld parallel / collinear / nearly parallel), %lld differ\n", n_pairs, n_par, bad_st);

  const double r = 0.015, dl = 0.005;
  std::vector<Shape> shapes;
  auto put = [&](Shape s, bool rotate) { if (rotate) rigid(s); shapes.push_back(std::move(s)); };
  // hairpins: two parallel runs D apart, D = 2r +- ulps and +- small relative steps
  for (int j : {-4, -3, -2, -1, 0, 1, 2, 3, 4, -1000, 1000, -1000000, 1000000})
    for (int nleg : {10, 14, 20, 30})
      for (int rot = 0; rot < 4; rot++) {
        const double D = j == 0 ? std::nextafter(2 * r, 1.0) : 2 * r * (1 + j * 2.2e-16);
        Shape s;
        for (int i = 0; i < nleg; i++) { s.push_back(i * dl); s.push_back(0); s.push_back(0); }
        for (int k = 1; k < 6; k++) {
          s.push_back((nleg - 1) * dl + 0.5 * D * std::sin(M_PI * k / 6)); s.push_back(0.5 * D * (1 - std::cos(M_PI * k / 6))); s.push_back(0);
        }
        for (int i = 0; i < nleg; i++) { s.push_back((nleg - 1 - i) * dl); s.push_back(D); s.push_back(0); }
        put(s, rot > 0);
      }
  // V shapes around the 3r arc rule and the 2r distance
  for (int ti = 0; ti < 29; ti++)
    for (int n1 : {3, 5, 8, 9, 10, 12, 17})
      for (double scale : {1.0, 1.0 - 1e-12, 1.0 + 1e-12, 0.9, 1.1}) {
        const double theta = 0.3 + 2.8 * ti / 28.0;
        const double d2[3] = {std::cos(M_PI - theta), std::sin(M_PI - theta), 0.0};
        Shape s;
        for (int i = 0; i < n1; i++) { s.push_back(-(n1 - i) * dl * scale); s.push_back(0); s.push_back(0); }
        s.push_back(0); s.push_back(0); s.push_back(0);
        for (int i = 0; i < n1; i++) for (int c = 0; c < 3; c++) s.push_back(d2[c] * (i + 1) * dl * scale);
        put(s, true);
      }
  // arcs and spirals
  for (int ti = 0; ti < 60; ti++)
    for (int n : {20, 40, 67, 120})
      for (double grow : {0.0, 1.0}) {
        const double turn = 1.2 + 13.0 * ti / 59.0, L = n * dl;
        Shape s = {0, 0, 0};
        double ang = 0, x = 0, y = 0, z = 0;
        for (int k = 0; k < n; k++) {
          x += std::cos(ang) * dl; y += std::sin(ang) * dl; z += 0.02 * dl;
          s.push_back(x); s.push_back(y); s.push_back(z);
          ang += turn / L * (1 + grow * (k * dl) / L) / (1 + grow / 2) * dl;
        }
        put(s, true);
      }
  // random walks with bounded turning per step: many true collisions, many near misses
  // ... and gentle ones (every third): backbones like the robot's, most of which the filter lets go
  for (int w = 0; w < 90000; w++) {
    const int n = 8 + (int)(gen() % 120);
    double d[3] = {Nrm(), Nrm(), Nrm()};
    Shape s = {0, 0, 0};
    double p[3] = {0, 0, 0};
    const double wob = (w % 3 == 2) ? U(0.002, 0.06) : U(0.05, 0.6);
    for (int k = 0; k < n - 1; k++) {
      double nn = 0;
      for (int c = 0; c < 3; c++) { d[c] += Nrm() * wob; }
      for (int c = 0; c < 3; c++) nn += d[c] * d[c];
      nn = std::sqrt(nn);
      const double step = dl * U(0.6, 1.2);
      for (int c = 0; c < 3; c++) { d[c] /= nn; p[c] += d[c] * step; s.push_back(p[c]); }
    }
    put(s, true);
  }
  // degenerate: repeated points, very short shapes
  put(Shape(30, 0.0), false);
  put(Shape{0, 0, 0, dl, 0, 0}, false);
  put(Shape{0, 0, 0, dl, 0, 0, 0, 0, 0, dl, 0, 0, 0, 0, 0, dl, 0, 0}, false);

  long long bad = 0, hits = 0, pairs = 0, dropped = 0, left[3] = {0, 0, 0};
  for (auto &s : shapes) {
    const int N = (int)(s.size() / 3);
    const bool want = orc_collides_self(s.data(), N, r) == 1;
    const bool got = kernel_decision(s.data(), N, r, &pairs);
    const int f = filter_decision(s.data(), N, r);
    hits += want;
    left[f]++;
    if (want != got) bad++;
    if (want && f != 2) dropped++;
  }
  std::printf("FP32 filter: %lld backbones left at the turning bound, %lld without a candidate pair, %lld to the exact stage; "
              "%lld true hits dropped\n", left[0], left[1], left[2], dropped);
  std::printf("collides_self: %zu backbones (%lld collide), %lld capsule pairs reached the exact test, %lld verdicts differ\n",
              shapes.size(), hits, pairs, bad);
  const bool mix = hits * 10 > (long long)shapes.size() && hits * 10 < 9 * (long long)shapes.size();
  // the filter must be exercised on both sides: it lets a good share of the collision-free backbones go
  const bool selective = left[0] > 1000 && left[1] > 1000 && left[2] >= hits;
  if (bad_st || bad || dropped || !selective || !mix || n_par < 1000000) {
    std::printf("FAILED\n");
    return 1;
  }
  std::printf("capsule pair ok\n");
  return 0;
}
