// Host check of the per-block arithmetic the device kernel env_add_primitives_kernel runs
// (interactive-rate-tendons_b200/csrc/env_prims.h) against the oracle's add_point / add_sphere / add_capsule,
// which are pinned by the reference's own text.  Compile with -ffp-contract=off.  The oracle is linked as the
// checker only.  Built and run by tests/test_abi_and_host.py.
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../interactive-rate-tendons_b200/csrc/env_prims.h"
#include "../../oracle/tendon_oracle.h"

static uint32_t morton(int bx, int by, int bz, int Nb) { return orc_morton_key(bx, by, bz, Nb); }

static int run(int Ng, const double *lim, unsigned seed, bool huge) {
  orc_grid g;
  std::memset(&g, 0, sizeof(g));
  g.Ng = Ng;
  std::memcpy(g.lim, lim, sizeof(g.lim));
  g.inv_rot[0] = g.inv_rot[4] = g.inv_rot[8] = 1;
  const int Nb = Ng / 4;
  double lo[3], hi[3], d[3], ext[3];
  for (int a = 0; a < 3; a++) {
    lo[a] = lim[2 * a]; hi[a] = lim[2 * a + 1];
    d[a] = (hi[a] - lo[a]) / Ng;
    ext[a] = hi[a] - lo[a];
  }
  std::mt19937_64 gen(seed);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  const double radii[6] = {0.001, 0.01, 0.05, 0.12, 0.02, huge ? 2.0 : 0.2};  // huge: a sphere beyond the grid
  double emin = ext[0] < ext[1] ? ext[0] : ext[1];
  emin = emin < ext[2] ? emin : ext[2];
  std::vector<double> prims, points;
  orc_octree *want = orc_octree_new(&g);
  for (int k = 0; k < 90; k++) {
    double a[3], b[3];
    for (int c = 0; c < 3; c++) {
      a[c] = lo[c] + ext[c] * (-0.3 + 1.6 * U(gen));
      b[c] = (k % 9 == 2) ? a[c] : a[c] + ext[c] * (U(gen) - 0.5);
    }
    double r = emin * radii[(int)(U(gen) * 6) % 6];
    if (k % 3 == 0) {
      orc_octree_add_point(want, a);
      points.insert(points.end(), a, a + 3);
    } else if (k % 3 == 1) {
      orc_octree_add_sphere(want, a, r);
      const double p[8] = {a[0], a[1], a[2], 0, 0, 0, r, 0.0};
      prims.insert(prims.end(), p, p + 8);
      points.insert(points.end(), a, a + 3);   // add_sphere starts with add_point(centre)
    } else {
      if (r > 0.2 * emin) r = 0.2 * emin;
      orc_octree_add_capsule(want, a, b, r);
      const double p[8] = {a[0], a[1], a[2], b[0], b[1], b[2], r, 1.0};
      prims.insert(prims.end(), p, p + 8);
      points.insert(points.end(), a, a + 3);
      points.insert(points.end(), b, b + 3);
    }
  }
  // limits and corners through add_point
  const double extra[4][3] = {{lo[0], lo[1], lo[2]}, {hi[0], hi[1], hi[2]},
                              {lo[0] + 3 * d[0], lo[1] + 3 * d[1], lo[2] + 3 * d[2]}, {hi[0] + 1e-12, hi[1], hi[2]}};
  for (auto &p : extra) { orc_octree_add_point(want, p); points.insert(points.end(), p, p + 3); }

  std::vector<uint64_t> got((size_t)Nb * Nb * Nb, 0);
  for (int bx = 0; bx < Nb; bx++)
    for (int by = 0; by < Nb; by++)
      for (int bz = 0; bz < Nb; bz++) {
        uint64_t acc = 0;
        for (size_t i = 0; i < prims.size(); i += 8) acc |= ep_block_bits(lo, d, bx, by, bz, &prims[i]);
        got[morton(bx, by, bz, Nb)] |= acc;
      }
  for (size_t i = 0; i < points.size(); i += 3) {
    int c[3];
    if (ep_point_cell(lo, hi, d, Ng, &points[i], c))
      got[morton(c[0] >> 2, c[1] >> 2, c[2] >> 2, Nb)] |= 1ull << ((c[0] & 3) * 16 + (c[1] & 3) * 4 + (c[2] & 3));
  }
  long long flips = 0, cells = 0;
  for (int bx = 0; bx < Nb; bx++)
    for (int by = 0; by < Nb; by++)
      for (int bz = 0; bz < Nb; bz++) {
        const uint64_t w = orc_octree_block(want, bx, by, bz), h = got[morton(bx, by, bz, Nb)];
        flips += __builtin_popcountll(w ^ h);
        cells += __builtin_popcountll(w);
      }
  orc_octree_free(want);
  std::printf("Ng=%d: %lld of %lld cells occupied, %lld flips\n", Ng, cells, (long long)Ng * Ng * Ng, flips);
  return flips == 0 && cells > 0 && (huge || cells < (long long)Ng * Ng * Ng) ? 0 : 1;
}

int main() {
  const double l1[6] = {0, 1, 0, 1, 0, 1}, l2[6] = {-0.3, 0.2, -0.1, 0.4, 0.0, 0.25}, l3[6] = {-0.21, 0.21, -0.21, 0.21, -0.21, 0.21};
  int bad = run(16, l1, 1, false) + run(64, l2, 2, false) + run(128, l3, 3, false) + run(32, l3, 4, true);
  std::printf(bad ? "FAILED\n" : "env primitives ok\n");
  return bad;
}
