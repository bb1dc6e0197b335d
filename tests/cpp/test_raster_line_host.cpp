// Host check of the voxel traversal the device rasteriser runs (interactive-rate-tendons_b200/csrc/raster_line.h:
// add_line with its division-free decisions, the hand-over to the literal code on close calls, path-order emission)
// against the oracle's add_line, which is pinned bit-exactly by the reference's own VoxelOctree::add_line text.
// Millions of segments of every kind on several grids -- far more than the GPU tests can afford -- and the count of
// segments that took the hand-over path, so that both paths are known to be exercised.  Compile with
// -ffp-contract=off (the device file is built with -fmad=false).  The oracle is linked as the checker only.
// Built and run by tests/test_abi_and_host.py.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <vector>

static long long n_fast = 0, n_literal = 0;
#define RL_ON_FAST_PATH_DONE ++n_fast;
#define RL_ON_LITERAL_PATH ++n_literal;
#include "../../interactive-rate-tendons_b200/csrc/raster_line.h"
#include "../../oracle/tendon_oracle.h"

struct Grid {
  double lo[3], hi[3], d[3], inv_d[3];
  int Ng;
};

// same cell -> (block, bit) rule as the device sinks (VoxelOctree::bitmask, VoxelOctree.cpp:1501-1503)
struct MapSink {
  std::map<uint32_t, uint64_t> &m;
  long long cells = 0;
  explicit MapSink(std::map<uint32_t, uint64_t> &mm) : m(mm) {}
  void cell(int ix, int iy, int iz) {
    const uint32_t key = ((uint32_t)(ix >> 2) << 16) | ((uint32_t)(iy >> 2) << 8) | (uint32_t)(iz >> 2);
    m[key] |= 1ull << ((ix & 3) * 16 + (iy & 3) * 4 + (iz & 3));
    cells++;
  }
  void finish() {}
};

static long long run(int Ng, const double *lim, unsigned seed, int nseg, long long *total_cells) {
  orc_grid og;
  std::memset(&og, 0, sizeof(og));
  og.Ng = Ng;
  std::memcpy(og.lim, lim, sizeof(og.lim));
  og.inv_rot[0] = og.inv_rot[4] = og.inv_rot[8] = 1;
  Grid g;
  g.Ng = Ng;
  double ext[3];
  for (int a = 0; a < 3; a++) {   // make_grid_dev (csrc/ctx.cu): VoxelOctree.cpp:152-177, :338
    g.lo[a] = lim[2 * a];
    g.hi[a] = lim[2 * a + 1];
    g.d[a] = (lim[2 * a + 1] - lim[2 * a]) / Ng;
    g.inv_d[a] = 1 / g.d[a];
    ext[a] = g.hi[a] - g.lo[a];
  }
  std::mt19937_64 gen(seed);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::normal_distribution<double> N01(0.0, 1.0);
  orc_octree *want = orc_octree_new(&og);
  long long bad = 0;
  std::vector<uint8_t> xyz;
  std::vector<uint64_t> bits;
  const int group = 8;   // segments per comparison
  for (int s0 = 0; s0 < nseg; s0 += group) {
    orc_octree_clear(want);
    std::map<uint32_t, uint64_t> got;
    for (int s = s0; s < s0 + group; s++) {
      double a[3], b[3];
      const int kind = s % 8;
      for (int c = 0; c < 3; c++) a[c] = g.lo[c] + ext[c] * (-0.1 + 1.2 * U(gen));
      switch (kind) {
        case 0:   // short, like a backbone segment (about one cell)
        case 1:
          for (int c = 0; c < 3; c++) b[c] = a[c] + N01(gen) * 1.2 * g.d[c];
          break;
        case 2:   // long, may cross or miss the grid
          for (int c = 0; c < 3; c++) b[c] = g.lo[c] + ext[c] * (-0.25 + 1.5 * U(gen));
          break;
        case 3: {   // axis aligned
          for (int c = 0; c < 3; c++) b[c] = a[c];
          const int c = (int)(U(gen) * 3) % 3;
          b[c] += ext[c] * (U(gen) - 0.5) * 0.3;
          break;
        }
        case 4:   // zero length
          for (int c = 0; c < 3; c++) b[c] = a[c];
          break;
        case 5:   // both ends on cell faces / corners: equal ray parameters, the close-call hand-over
          for (int c = 0; c < 3; c++) {
            a[c] = g.lo[c] + g.d[c] * (double)(int)(U(gen) * Ng);
            b[c] = a[c] + g.d[c] * (double)((int)(U(gen) * 9) - 4);
          }
          break;
        case 6: {   // near-degenerate direction components (the reference's 1e-10 threshold)
          const double tiny[3] = {1e-12, 3e-11, 2e-10};
          for (int c = 0; c < 3; c++) b[c] = a[c] + (U(gen) - 0.5) * ext[c] * 0.1;
          const int c = (int)(U(gen) * 3) % 3;
          b[c] = a[c] + (U(gen) - 0.5) * ext[c] * tiny[(int)(U(gen) * 3) % 3];
          break;
        }
        default:   // exact diagonals through cell corners
          for (int c = 0; c < 3; c++) a[c] = g.lo[c] + g.d[c] * ((double)(int)(U(gen) * Ng) + 0.5);
          {
            const int n = (int)(U(gen) * 12) - 6;
            for (int c = 0; c < 3; c++) b[c] = a[c] + g.d[c] * n * ((U(gen) < 0.5) ? 1 : -1);
          }
          break;
      }
      orc_octree_add_line(want, a, b);
      MapSink sink(got);
      const D3 A = {a[0], a[1], a[2]}, B = {b[0], b[1], b[2]};
      add_line(g, sink, A, B);
      *total_cells += sink.cells;
    }
    const int64_t n = orc_octree_export(want, 0, nullptr, nullptr);
    xyz.resize((size_t)(n > 0 ? n : 1) * 3);
    bits.resize((size_t)(n > 0 ? n : 1));
    orc_octree_export(want, n, xyz.data(), bits.data());
    bool same = (int64_t)got.size() == n;
    for (int64_t i = 0; same && i < n; i++) {
      const uint32_t key = ((uint32_t)xyz[3 * i] << 16) | ((uint32_t)xyz[3 * i + 1] << 8) | (uint32_t)xyz[3 * i + 2];
      auto it = got.find(key);
      same = it != got.end() && it->second == bits[(size_t)i];
    }
    if (!same) bad++;
  }
  orc_octree_free(want);
  return bad;
}

// find_cell (reciprocal multiply, exact division next to a cell face) and the point-pair event of should_subdivide
// against the oracle's find_cell (VoxelOctree.cpp:309-317, domain_check :1511-1521): points anywhere, on cell
// faces +- a few ulp, on and just beyond the limits; pairs near and far
static long long run_cells(int Ng, const double *lim, unsigned seed, int npts, long long *n_exact_div) {
  orc_grid og;
  std::memset(&og, 0, sizeof(og));
  og.Ng = Ng;
  std::memcpy(og.lim, lim, sizeof(og.lim));
  og.inv_rot[0] = og.inv_rot[4] = og.inv_rot[8] = 1;
  Grid g;
  g.Ng = Ng;
  double ext[3];
  for (int a = 0; a < 3; a++) {
    g.lo[a] = lim[2 * a];
    g.hi[a] = lim[2 * a + 1];
    g.d[a] = (lim[2 * a + 1] - lim[2 * a]) / Ng;
    g.inv_d[a] = 1 / g.d[a];
    ext[a] = g.hi[a] - g.lo[a];
  }
  std::mt19937_64 gen(seed);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  long long bad = 0;
  auto make = [&](double *p, int kind) {
    for (int c = 0; c < 3; c++) {
      switch (kind) {
        case 0: p[c] = g.lo[c] + ext[c] * (-0.02 + 1.04 * U(gen)); break;                       // anywhere
        case 1: {                                                                               // on a face +- ulps
          double v = g.lo[c] + g.d[c] * (double)(int)(U(gen) * (Ng + 1));
          const int k = (int)(U(gen) * 7) - 3;
          for (int i = 0; i < (k < 0 ? -k : k); i++) v = std::nextafter(v, k < 0 ? -1e300 : 1e300);
          p[c] = v;
          break;
        }
        case 2: p[c] = g.lo[c] + g.d[c] * ((double)(int)(U(gen) * Ng) + (U(gen) < 0.5 ? 1e-10 : 1 - 1e-10)); break;
        default: {                                                                              // the limits
          const double pick[4] = {g.lo[c], g.hi[c], std::nextafter(g.hi[c], 1e300), std::nextafter(g.lo[c], -1e300)};
          p[c] = (U(gen) < 0.7) ? g.lo[c] + ext[c] * U(gen) : pick[(int)(U(gen) * 4) % 4];
        }
      }
    }
  };
  for (int i = 0; i < npts; i++) {
    double a[3], b[3];
    make(a, i % 4);
    if (i % 3 == 0) make(b, (i / 4) % 4);
    else for (int c = 0; c < 3; c++) b[c] = a[c] + g.d[c] * (U(gen) - 0.5) * ((i % 3 == 1) ? 1.9 : 5.0);
    const D3 A = {a[0], a[1], a[2]}, B = {b[0], b[1], b[2]};
    long long ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
    int64_t wa[3], wb[3];
    const bool oka = find_cell(g, A, ca), okb = find_cell(g, B, cb);
    const int ea = orc_find_cell(&og, a, wa), eb = orc_find_cell(&og, b, wb);
    if (oka != (ea == 0) || okb != (eb == 0)) { bad++; continue; }
    if (oka && (ca[0] != wa[0] || ca[1] != wa[1] || ca[2] != wa[2])) bad++;
    if (okb && (cb[0] != wb[0] || cb[1] != wb[1] || cb[2] != wb[2])) bad++;
    int want = 2;
    if (ea == 0 && eb == 0) {
      const long long dx = std::llabs(wa[0] - wb[0]), dy = std::llabs(wa[1] - wb[1]), dz = std::llabs(wa[2] - wb[2]);
      want = (dx > 1 || dy > 1 || dz > 1) ? 1 : 0;
    }
    if (pair_event_core(g, A, B) != want) bad++;
    if (i % 4 == 1 || i % 4 == 2) ++*n_exact_div;
  }
  return bad;
}

int main() {
  const double lung[6] = {-0.21, 0.21, -0.21, 0.21, -0.21, 0.21};
  const double skew[6] = {-0.1, 0.3, 0.05, 0.25, -0.4, 0.0};
  const double unit[6] = {0, 1, 0, 1, 0, 1};
  struct { int Ng; const double *lim; unsigned seed; int nseg; } cases[] = {
      {128, lung, 1, 1200000}, {128, skew, 2, 600000}, {32, unit, 3, 400000}, {256, lung, 4, 400000}};
  long long all_bad = 0;
  for (auto &c : cases) {
    long long cells = 0;
    const long long bad = run(c.Ng, c.lim, c.seed, c.nseg, &cells);
    std::printf("Ng %3d: %d segments, %lld cells emitted, %lld groups differ\n", c.Ng, c.nseg, cells, bad);
    all_bad += bad;
  }
  {
    long long near_face = 0, bad = 0;
    bad += run_cells(128, lung, 11, 3000000, &near_face);
    bad += run_cells(128, skew, 12, 1500000, &near_face);
    bad += run_cells(512, unit, 13, 1500000, &near_face);
    std::printf("find_cell / pair events: 6000000 point pairs (%lld with a point within ulps of a cell face), %lld differ\n",
                near_face, bad);
    all_bad += bad;
  }
  std::printf("division-free path %lld segments, literal path %lld (close calls, near-degenerate directions, "
              "zero length)\n", n_fast, n_literal);
  if (all_bad || n_fast < 1000000 || n_literal < 100000) {
    std::printf("FAILED\n");
    return 1;
  }
  std::printf("raster line ok\n");
  return 0;
}
