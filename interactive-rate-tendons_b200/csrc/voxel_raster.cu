// voxel_raster.cu -- K2 `swept_voxel_raster`: backbone polyline -> sparse voxel block sets,
// for roadmap vertices (one FK sample per set) and edges (adaptive swept volume).
//
// Replaces, behind the same results:
//   VoxelOctree::add_line / add_piecewise_line      collision/VoxelOctree.cpp:325-432
//   segment_aabox_intersect                         collision/collision_primitives.h:62-85
//   VoxelBackboneValidityChecker::voxelize_impl     motion-planning/VoxelBackboneValidityChecker.h:49-57
//   VoxelEnvironment::voxelize_valid_backbone_motion motion-planning/VoxelEnvironment.cpp:207-444
//   VoxelBackboneMotionValidator::generic_voxelize  motion-planning/VoxelBackboneMotionValidator.cpp:41-74
//
// B200-first structure (not a translation of the reference's per-edge LIFO loop):
//   * the bisection of ALL edges of a chunk runs level-synchronously: every round interpolates the
//     midpoints of all open intervals, runs ONE batched K1 launch (+ validity epilogue) over them,
//     then one warp per interval evaluates should_subdivide (VoxelEnvironment.cpp:304-341).  The
//     sample set this produces below the first invalid t equals the reference's depth-first
//     order exactly (DESIGN.md "level-synchronous bisection" has the argument).
//   * rasterisation: one warp per set; lanes take polyline segments, run the reference's voxel
//     traversal literally (same operation order, this file is compiled with -fmad=false so no
//     contraction changes a cell index) and OR cell bits into a shared-memory hash keyed by the
//     Morton block key; the table is then rank-sorted by key, which is the reference's
//     visit_leaves order, and written as {u32 key, u64 bits} records (8 consecutive keys = one
//     512-bit 2x2x2 super-block).  Two passes (count, exclusive scan, emit) give an exact CSR
//     without per-set capacity slots.
#include <cmath>
#include <cstring>
#include <vector>

#include <chrono>

#include "common.cuh"

namespace {

// IRT_B200_TRACE=1: per-phase wall-clock of the K2 host pipeline (synchronising; debugging only)
struct Trace {
  bool on;
  cudaStream_t st;
  std::chrono::steady_clock::time_point t0;
  explicit Trace(cudaStream_t s) : on(getenv("IRT_B200_TRACE") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
  void point(const char *name, long long count = -1) {
    if (!on) return;
    cudaStreamSynchronize(st);
    auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[irt trace] %-28s %9.3f ms", name, std::chrono::duration<double, std::milli>(t1 - t0).count());
    if (count >= 0) std::fprintf(stderr, "  (%lld)", count);
    std::fprintf(stderr, "\n");
    t0 = std::chrono::steady_clock::now();
  }
};

constexpr int RS_WARPS = 4;
constexpr int RS_HLOG_SMALL = 8;     // main pass: 256 hash entries per warp (a set rarely exceeds ~150 blocks)
constexpr int RS_HLOG_BIG = 12;      // fallback pass for the sets that overflow it: 4096 entries, 1 warp per CTA
constexpr int RS_SLOT = 1 << RS_HLOG_SMALL;   // slot capacity of the main pass
constexpr int RS_BIGSLOT = 1 << RS_HLOG_BIG;  // slot capacity of the fallback pass
constexpr int RS_MAX_OVF = 2048;     // fallback slots per raster launch
constexpr uint32_t RS_EMPTY = 0xffffffffu;
constexpr int RS_MAXS = 128;          // samples of one set handled by the flat segment walk
constexpr int rs_warp_bytes(int hlog) { return (1 << hlog) * 18 + 16 + RS_MAXS * 8; }
constexpr uint32_t INVALID_MASK = IRT_FLAG_NONCONVERGED | IRT_FLAG_LENGTH_LIMIT |
                                  IRT_FLAG_SELF_COLLISION | IRT_FLAG_BAD_STATE | IRT_FLAG_ENV_COLLISION;

struct D3 {
  double x, y, z;
};

__device__ __forceinline__ uint32_t spread3(uint32_t x) {
  x = (x | (x << 16)) & 0x030000FFu;
  x = (x | (x << 8)) & 0x0300F00Fu;
  x = (x | (x << 4)) & 0x030C30C3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}
// x is the most significant bit of each 3-bit group: child index 4*x + 2*y + z (TreeNode.h:66-68)
__device__ __forceinline__ uint32_t morton_key(uint32_t bx, uint32_t by, uint32_t bz) {
  return (spread3(bx) << 2) | (spread3(by) << 1) | spread3(bz);
}

__device__ __forceinline__ D3 rotate_pt(const GridDev &g, const double *p) {
  if (g.identity_rot) return {p[0], p[1], p[2]};
  // Eigen 3x3 * 3x1: (m0*b0 + m1*b1) + m2*b2
  return {(g.inv_rot[0] * p[0] + g.inv_rot[1] * p[1]) + g.inv_rot[2] * p[2],
          (g.inv_rot[3] * p[0] + g.inv_rot[4] * p[1]) + g.inv_rot[5] * p[2],
          (g.inv_rot[6] * p[0] + g.inv_rot[7] * p[1]) + g.inv_rot[8] * p[2]};
}

struct WarpHash {
  uint32_t *keys;   // [H]
  unsigned long long *bits;  // [H]
  uint16_t *list;   // [H] slots in insertion order (dense list of the occupied entries)
  uint32_t *count;  // number of occupied entries
  uint32_t *overflow;
  int hlog;         // H = 1 << hlog
};

__device__ __forceinline__ void hash_insert(const WarpHash &h, uint32_t key, unsigned long long mask) {
  const uint32_t H = 1u << h.hlog;
  if (*h.overflow) return;  // this set will be redone with the big table
  uint32_t slot = (key * 2654435761u) >> (32 - h.hlog);
  for (uint32_t probe = 0; probe < H; probe++) {
    const uint32_t old = atomicCAS(&h.keys[slot], RS_EMPTY, key);
    if (old == RS_EMPTY) {
      const uint32_t idx = atomicAdd(h.count, 1u);
      h.list[idx] = (uint16_t)slot;
      if (idx >= (H >> 2) * 3) *h.overflow = 1u;  // keep the load factor below 3/4
    }
    if (old == RS_EMPTY || old == key) {
      atomicOr(&h.bits[slot], mask);
      return;
    }
    slot = (slot + 1) & (H - 1);
  }
  *h.overflow = 1u;
}

// VoxelOctree::set_cell(ix,iy,iz,true) -- collision/VoxelOctree.cpp:262-272,1501-1503.
// Consecutive cells of a segment mostly fall into the same 4x4x4 block: bits are collected in
// registers and flushed to the shared hash (one Morton key + one atomicOr) when the block changes.
struct BlockAcc {
  int bx = -1, by = 0, bz = 0;
  unsigned long long mask = 0ull;
};
__device__ __forceinline__ void flush_block(const WarpHash &h, BlockAcc &acc) {
  if (acc.mask) hash_insert(h, morton_key((uint32_t)acc.bx, (uint32_t)acc.by, (uint32_t)acc.bz), acc.mask);
  acc.mask = 0ull;
}
__device__ __forceinline__ void set_cell(const WarpHash &h, BlockAcc &acc, int ix, int iy, int iz) {
  const int bx = ix >> 2, by = iy >> 2, bz = iz >> 2;
  if (bx != acc.bx || by != acc.by || bz != acc.bz) {
    flush_block(h, acc);
    acc.bx = bx; acc.by = by; acc.bz = bz;
  }
  acc.mask |= 1ull << ((ix & 3) * 16 + (iy & 3) * 4 + (iz & 3));
}

// cell sinks of the voxel traversal: HashSink builds a voxel set, EnvSink only asks whether any
// traversed cell is occupied in the environment (AbstractVoxelValidityChecker::collides of a shape)
struct HashSink {
  const WarpHash &h;
  BlockAcc acc;
  __device__ __forceinline__ explicit HashSink(const WarpHash &hh) : h(hh) {}
  __device__ __forceinline__ void cell(int ix, int iy, int iz) { set_cell(h, acc, ix, iy, iz); }
  __device__ __forceinline__ void finish() { flush_block(h, acc); }
};
struct EnvSink {
  const uint64_t *env;
  const uint32_t *occ;
  bool hit = false;
  __device__ __forceinline__ EnvSink(const uint64_t *e, const uint32_t *o) : env(e), occ(o) {}
  __device__ __forceinline__ void cell(int ix, int iy, int iz) {
    const uint32_t key = morton_key((uint32_t)ix >> 2, (uint32_t)iy >> 2, (uint32_t)iz >> 2);
    if ((occ[key >> 5] >> (key & 31)) & 1u)
      if ((env[key] >> ((ix & 3) * 16 + (iy & 3) * 4 + (iz & 3))) & 1ull) hit = true;
  }
  __device__ __forceinline__ void finish() {}
};

// collision/collision_primitives.h:62-85, literal operation order
__device__ bool segment_aabox_intersect(const D3 &A, const D3 &B, const D3 &C, const D3 &D) {
  const D3 AB = {B.x - A.x, B.y - A.y, B.z - A.z};
  const double len = sqrt((AB.x * AB.x + AB.y * AB.y) + AB.z * AB.z) / 2;
  const double l2 = 2 * len;
  const D3 U = {AB.x / l2, AB.y / l2, AB.z / l2};
  const D3 Uabs = {fabs(U.x), fabs(U.y), fabs(U.z)};
  const D3 P = {(A.x + B.x) / 2 - (D.x + C.x) / 2, (A.y + B.y) / 2 - (D.y + C.y) / 2,
                (A.z + B.z) / 2 - (D.z + C.z) / 2};
  const D3 ext = {fabs(D.x - C.x) / 2, fabs(D.y - C.y) / 2, fabs(D.z - C.z) / 2};
  const D3 UxP = {fabs(U.y * P.z - U.z * P.y), fabs(U.z * P.x - U.x * P.z), fabs(U.x * P.y - U.y * P.x)};
  const D3 Pabs = {fabs(P.x), fabs(P.y), fabs(P.z)};
  const bool separated = Pabs.x > ext.x + len * Uabs.x || Pabs.y > ext.y + len * Uabs.y ||
                         Pabs.z > ext.z + len * Uabs.z ||
                         UxP.x > ext.y * Uabs.z + ext.z * Uabs.y ||
                         UxP.y > ext.z * Uabs.x + ext.x * Uabs.z ||
                         UxP.z > ext.x * Uabs.y + ext.y * Uabs.x;
  return !separated;
}

// VoxelOctree::add_line -- collision/VoxelOctree.cpp:325-426, reproduced literally including the
// "voxel index times metric cell size" initial error (:371-373) and the overshoot past B (:423-424).
template <typename Sink>
__device__ void add_line(const GridDev &g, Sink &sink, const D3 &a, const D3 &b) {
  const D3 ll = {g.lo[0], g.lo[1], g.lo[2]}, ur = {g.hi[0], g.hi[1], g.hi[2]};
  // A segment whose endpoints both lie inside the grid box (by a safety margin of a ten-thousandth
  // of a cell) intersects it: the reference's separating-axis test cannot report otherwise, so it
  // is only evaluated for segments that touch or leave the box.
  const double mx = 1e-4 * g.d[0], my = 1e-4 * g.d[1], mz = 1e-4 * g.d[2];
  const bool inside = a.x > ll.x + mx && a.x < ur.x - mx && b.x > ll.x + mx && b.x < ur.x - mx &&
                      a.y > ll.y + my && a.y < ur.y - my && b.y > ll.y + my && b.y < ur.y - my &&
                      a.z > ll.z + mz && a.z < ur.z - mz && b.z > ll.z + mz && b.z < ur.z - mz;
  if (!inside && !segment_aabox_intersect(a, b, ll, ur)) return;
  const D3 A = {(a.x - ll.x) * g.inv_d[0], (a.y - ll.y) * g.inv_d[1], (a.z - ll.z) * g.inv_d[2]};
  const D3 B = {(b.x - ll.x) * g.inv_d[0], (b.y - ll.y) * g.inv_d[1], (b.z - ll.z) * g.inv_d[2]};
  const int Axi = (int)A.x - (A.x < 0), Ayi = (int)A.y - (A.y < 0), Azi = (int)A.z - (A.z < 0);
  const int Bxi = (int)B.x - (B.x < 0), Byi = (int)B.y - (B.y < 0), Bzi = (int)B.z - (B.z < 0);
  const int N = g.Ng;
#define IDX_IN(v) (0 <= (v) && (v) < N)
#define VOX_IN(x, y, z) (IDX_IN(x) && IDX_IN(y) && IDX_IN(z))
  bool entered = VOX_IN(Axi, Ayi, Azi);
  if (VOX_IN(Bxi, Byi, Bzi)) sink.cell(Bxi, Byi, Bzi);
  if (entered) sink.cell(Axi, Ayi, Azi);
  D3 U = {B.x - A.x, B.y - A.y, B.z - A.z};
  {
    const double z = (U.x * U.x + U.y * U.y) + U.z * U.z;  // Eigen normalized()
    if (z > 0.0) {
      const double n = sqrt(z);
      U.x /= n; U.y /= n; U.z /= n;
    }
  }
  const int step_x = 1 - 2 * (U.x < 0), step_y = 1 - 2 * (U.y < 0), step_z = 1 - 2 * (U.z < 0);
  const double ex = fabs(A.x - (Axi + step_x) * g.d[0]);
  const double ey = fabs(A.y - (Ayi + step_y) * g.d[1]);
  const double ez = fabs(A.z - (Azi + step_z) * g.d[2]);
  const double ux = fabs(U.x), uy = fabs(U.y), uz = fabs(U.z);
  const double threshold = 1e-10;
  const double tx_delta = (ux > threshold) ? 1 / ux : 1 / threshold;
  const double ty_delta = (uy > threshold) ? 1 / uy : 1 / threshold;
  const double tz_delta = (uz > threshold) ? 1 / uz : 1 / threshold;
  double tx = fabs(ex * tx_delta), ty = fabs(ey * ty_delta), tz = fabs(ez * tz_delta);
  int xi = Axi, yi = Ayi, zi = Azi;
  while (step_x * (Bxi - xi) >= 0 && step_y * (Byi - yi) >= 0 && step_z * (Bzi - zi) >= 0) {
    const bool tx_is_min = (tx < ty) && (tx < tz);
    const bool ty_is_min = !(tx < ty) && (ty < tz);
    if (tx_is_min) {
      xi += step_x;
      if (entered && !IDX_IN(xi)) break;
      tx += tx_delta;
    } else if (ty_is_min) {
      yi += step_y;
      if (entered && !IDX_IN(yi)) break;
      ty += ty_delta;
    } else {
      zi += step_z;
      if (entered && !IDX_IN(zi)) break;
      tz += tz_delta;
    }
    if (!entered && VOX_IN(xi, yi, zi)) entered = true;
    if (entered) sink.cell(xi, yi, zi);
  }
  sink.finish();
#undef IDX_IN
#undef VOX_IN
}

// One warp per set.  A set is a linked list of FK samples (set_head / sample_next); a sample is
// included iff t < tlimit[set] (edges) -- vertices have a single sample and no limit.
// Output: the set's occupied leaf blocks, key-sorted, in its slot (slot_keys / slot_bits),
// counts[set], optional t_last / nsamples.  A gather kernel then packs the slots into the CSR at
// the scanned offsets.  Per-set cost is proportional to the number of occupied blocks: occupied
// hash slots are kept in an append list, so neither the sort nor the clean-up scans the table.
// HLOG = RS_HLOG_SMALL: main pass over all sets; a set that outgrows the small table is appended
// to ovf_list and left for the HLOG = RS_HLOG_BIG pass, which works through that list (count on
// the device) with one warp per CTA and writes into the big slots.
template <int HLOG, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
swept_voxel_raster_kernel(const GridDev g, const double *__restrict__ pts,
                          const int32_t *__restrict__ npts, int cap_pts,
                          const int32_t *__restrict__ set_head, const int32_t *__restrict__ sample_next,
                          const double *__restrict__ sample_t, const double *__restrict__ tlimit,
                          int64_t nsets, uint32_t *__restrict__ counts, double *__restrict__ t_last,
                          int32_t *__restrict__ nsamples, uint32_t *__restrict__ slot_keys,
                          uint64_t *__restrict__ slot_bits, uint32_t *__restrict__ set_flags,
                          int32_t *__restrict__ ovf_list, int32_t *__restrict__ ovf_count,
                          int32_t *__restrict__ ovf_slot) {
  constexpr int H = 1 << HLOG;
  constexpr bool BIG = (HLOG == RS_HLOG_BIG);
  extern __shared__ unsigned long long rs_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per warp: bits u64[H] | keys u32[H] | dense keys u32[H] | list u16[H] | count, overflow | sample list
  unsigned char *wbase = reinterpret_cast<unsigned char *>(rs_smem) + (size_t)warp * rs_warp_bytes(HLOG);
  WarpHash h;
  h.hlog = HLOG;
  h.bits = reinterpret_cast<unsigned long long *>(wbase);
  h.keys = reinterpret_cast<uint32_t *>(wbase + H * 8);
  uint32_t *dk = reinterpret_cast<uint32_t *>(wbase + H * 12);
  h.list = reinterpret_cast<uint16_t *>(wbase + H * 16);
  h.count = reinterpret_cast<uint32_t *>(wbase + H * 18);
  h.overflow = h.count + 1;
  int32_t *ls = reinterpret_cast<int32_t *>(wbase + H * 18 + 16);  // sample ids
  int32_t *lc = ls + RS_MAXS;                                      // first flat segment index
  for (int i = lane; i < H; i += 32) { h.keys[i] = RS_EMPTY; h.bits[i] = 0ull; }
  if (lane == 0) { *h.count = 0u; *h.overflow = 0u; }
  __syncwarp();

  const int64_t n_work = BIG ? (int64_t)min(*ovf_count, RS_MAX_OVF) : nsets;
  for (int64_t w = (int64_t)blockIdx.x * WARPS + warp; w < n_work; w += (int64_t)gridDim.x * WARPS) {
    const int64_t set = BIG ? (int64_t)ovf_list[w] : w;
    const double lim = tlimit ? tlimit[set] : 0.0;
    double tl = 0.0;
    int ns = 0, nsm = 0, total = 0;
    // pass 1: list the included samples with their cumulative segment counts (flat work list)
    for (int smp = set_head[set]; smp >= 0; smp = sample_next ? sample_next[smp] : -1) {
      ns++;
      if (tlimit) {
        const double t = sample_t[smp];
        if (!(t < lim)) continue;  // VoxelEnvironment.cpp:409
        if (tl < t) tl = t;        // :419-422 last valid t
      }
      const int P = npts[smp];
      if (P < 2) continue;         // add_piecewise_line of fewer than 2 points adds nothing
      if (nsm < RS_MAXS) {
        if (lane == 0) { ls[nsm] = smp; lc[nsm] = total; }
        nsm++;
        total += P - 1;
      } else {  // very long sample lists: the remainder goes sample by sample
        const double *sp = pts + (int64_t)smp * cap_pts * 3;
        for (int i = 1 + lane; i < P; i += 32) {
          HashSink sink(h);
          add_line(g, sink, rotate_pt(g, sp + 3 * (i - 1)), rotate_pt(g, sp + 3 * i));
        }
      }
    }
    __syncwarp();
    // pass 2: lanes take consecutive segments of the concatenated polylines
    for (int f = lane; f < total; f += 32) {
      int k = 0;
      for (int step = RS_MAXS >> 1; step > 0; step >>= 1)
        if (k + step < nsm && lc[k + step] <= f) k += step;
      const int i = f - lc[k] + 1;  // segment (i-1, i) of sample ls[k]
      const double *sp = pts + (int64_t)ls[k] * cap_pts * 3;
      HashSink sink(h);
      add_line(g, sink, rotate_pt(g, sp + 3 * (i - 1)), rotate_pt(g, sp + 3 * i));
    }
    __syncwarp();
    const bool over = *h.overflow != 0u;
    const int used = (int)min(*h.count, (uint32_t)H);
    int cnt = used;
    int64_t slot_id = set;
    if (!BIG && over) {
      // defer to the big-table pass (or give up with a flag when its slots are exhausted)
      cnt = 0;
      if (lane == 0) {
        const int32_t idx = atomicAdd(ovf_count, 1);
        if (idx < RS_MAX_OVF) ovf_list[idx] = (int32_t)set;
        else if (set_flags) set_flags[set] |= IRT_FLAG_CAPACITY;
      }
    } else {
      if (BIG) {
        slot_id = w;
        if (over && lane == 0 && set_flags) set_flags[set] |= IRT_FLAG_CAPACITY;
      }
      for (int i = lane; i < cnt; i += 32) dk[i] = h.keys[h.list[i]];
      __syncwarp();
      // rank sort by key == the reference's visit_leaves order
      uint32_t *sk = slot_keys + slot_id * H;
      uint64_t *sb = slot_bits + slot_id * H;
      for (int i = lane; i < cnt; i += 32) {
        const uint32_t k = dk[i];
        int rank = 0;
        for (int j = 0; j < cnt; j++) rank += (dk[j] < k);
        sk[rank] = k;
        sb[rank] = h.bits[h.list[i]];
      }
    }
    if (lane == 0) {
      counts[set] = (uint32_t)cnt;
      if (BIG) ovf_slot[set] = (int32_t)w + 1;
      if (t_last) t_last[set] = tl;
      if (nsamples) nsamples[set] = ns;
    }
    __syncwarp();
    for (int i = lane; i < used; i += 32) {  // clean only what was used
      const int sl = h.list[i];
      h.keys[sl] = RS_EMPTY;
      h.bits[sl] = 0ull;
    }
    __syncwarp();
    if (lane == 0) { *h.count = 0u; *h.overflow = 0u; }
    __syncwarp();
  }
}

// voxelize_until_invalid: a sample is also invalid when its own backbone voxels hit the
// environment (_vc->collides(shape), VoxelBackboneMotionValidator.cpp:83-91).  One warp per
// sample; no set is built, the traversal just probes the environment grid.
__global__ void sample_env_collision_kernel(const GridDev g, const double *__restrict__ pts,
                                            const int32_t *__restrict__ npts, int cap_pts, int64_t n,
                                            const uint64_t *__restrict__ env, const uint32_t *__restrict__ occ,
                                            uint32_t *__restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int64_t smp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (smp >= n) return;
  if (flags[smp] & INVALID_MASK) return;  // is_valid_shape short-circuits the collision check
  const int P = npts[smp];
  const double *sp = pts + smp * (int64_t)cap_pts * 3;
  EnvSink sink(env, occ);
  for (int i = 1 + lane; i < P; i += 32) add_line(g, sink, rotate_pt(g, sp + 3 * (i - 1)), rotate_pt(g, sp + 3 * i));
  if (__any_sync(0xffffffffu, sink.hit) && lane == 0) flags[smp] |= IRT_FLAG_ENV_COLLISION;
}

// pack the slots into the CSR: one warp per set, coalesced copies
__global__ void raster_gather_kernel(const uint32_t *__restrict__ slot_keys, const uint64_t *__restrict__ slot_bits,
                                     const uint32_t *__restrict__ big_keys, const uint64_t *__restrict__ big_bits,
                                     const int32_t *__restrict__ ovf_slot,
                                     const uint32_t *__restrict__ counts, const uint64_t *__restrict__ offsets,
                                     int64_t nsets, uint32_t *__restrict__ out_keys, uint64_t *__restrict__ out_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t set = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (set >= nsets) return;
  const uint32_t n = counts[set];
  const uint64_t base = offsets[set];
  const int32_t ov = ovf_slot[set];
  const uint32_t *sk = ov ? big_keys + (int64_t)(ov - 1) * RS_BIGSLOT : slot_keys + set * RS_SLOT;
  const uint64_t *sb = ov ? big_bits + (int64_t)(ov - 1) * RS_BIGSLOT : slot_bits + set * RS_SLOT;
  for (uint32_t i = lane; i < n; i += 32) {
    out_keys[base + i] = sk[i];
    out_bits[base + i] = sb[i];
  }
}

// ---- exclusive scan: uint32 counts -> uint64 offsets (n+1 entries) ------------------------
constexpr int SCAN_T = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_T * SCAN_ITEMS;

__global__ void scan_tile_sums_kernel(const uint32_t *__restrict__ in, int64_t n,
                                      uint64_t *__restrict__ tile_sums) {
  __shared__ uint64_t sh[SCAN_T / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  uint64_t s = 0;
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const int64_t i = base + (int64_t)k * SCAN_T + threadIdx.x;
    if (i < n) s += in[i];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = 0;
    for (int w = 0; w < SCAN_T / 32; w++) t += sh[w];
    tile_sums[blockIdx.x] = t;
  }
}

__global__ void scan_tile_offsets_kernel(uint64_t *tile_sums, int64_t ntiles, uint64_t *total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint64_t acc = 0;
    for (int64_t i = 0; i < ntiles; i++) {
      const uint64_t c = tile_sums[i];
      tile_sums[i] = acc;
      acc += c;
    }
    *total = acc;
  }
}

__global__ void scan_apply_kernel(const uint32_t *__restrict__ in, int64_t n,
                                  const uint64_t *__restrict__ tile_sums,
                                  const uint64_t *__restrict__ total, uint64_t base_off,
                                  uint64_t *__restrict__ out) {
  // thread t owns SCAN_ITEMS consecutive items of the tile
  __shared__ uint64_t sh[SCAN_T];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint64_t s = 0;
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    s += v[k];
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < SCAN_T; o <<= 1) {  // Hillis-Steele inclusive scan of the thread sums
    uint64_t add = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
    __syncthreads();
    sh[threadIdx.x] += add;
    __syncthreads();
  }
  uint64_t acc = base_off + tile_sums[blockIdx.x] + sh[threadIdx.x] - s;
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = acc;
    acc += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = base_off + *total;
}

// offsets[0..n] = base_off + exclusive scan of counts[0..n); returns the chunk total on the host
int exclusive_scan(irt_ctx *ctx, const uint32_t *d_counts, int64_t n, uint64_t base_off,
                   uint64_t *d_offsets, uint64_t *d_tmp /* ntiles + 1 */, uint64_t *h_total,
                   cudaStream_t st) {
  const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  scan_tile_sums_kernel<<<(unsigned)ntiles, SCAN_T, 0, st>>>(d_counts, n, d_tmp);
  scan_tile_offsets_kernel<<<1, 32, 0, st>>>(d_tmp, ntiles, d_tmp + ntiles);
  scan_apply_kernel<<<(unsigned)ntiles, SCAN_T, 0, st>>>(d_counts, n, d_tmp, d_tmp + ntiles, base_off, d_offsets);
  ctx->launches.fetch_add(3);
  IRT_CUDA(ctx, cudaGetLastError());
  IRT_CUDA(ctx, cudaMemcpyAsync(h_total, d_tmp + ntiles, 8, cudaMemcpyDeviceToHost, st));
  IRT_CUDA(ctx, cudaStreamSynchronize(st));
  return IRT_OK;
}

// ---- vertex mode helpers --------------------------------------------------------------------
__global__ void vertex_heads_kernel(const uint32_t *__restrict__ flags, int64_t n,
                                    int32_t *__restrict__ heads) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) heads[i] = (flags[i] & INVALID_MASK) ? -1 : (int32_t)i;
}

// ---- edge mode: level-synchronous bisection -----------------------------------------------------
struct Interval {
  int32_t edge, ia, ib;
};
struct Pending {
  int32_t edge, ia, im, ib;
};

struct EdgePool {
  // per edge
  const double *a, *b;        // [E][S]
  const double *thr;          // [E] rel_threshold
  unsigned long long *first_invalid;  // [E] double bits (positive -> uint order == double order)
  int32_t *head;              // [E]
  uint32_t *eflags;           // [E]
  // per sample
  int32_t *s_edge, *s_next, *s_npts;
  uint32_t *s_flags;
  double *s_t, *s_state, *s_p;
  int32_t cap_samples;
  int S, N, enable_rotation, enable_retraction, cap_pts;
};

__global__ void edge_init_kernel(EdgePool P, int32_t E) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int32_t i0 = 2 * e, i1 = 2 * e + 1;
  P.s_edge[i0] = e; P.s_edge[i1] = e;
  P.s_t[i0] = 0.0; P.s_t[i1] = 1.0;
  for (int k = 0; k < P.S; k++) {
    P.s_state[(int64_t)i0 * P.S + k] = P.a[(int64_t)e * P.S + k];
    P.s_state[(int64_t)i1 * P.S + k] = P.b[(int64_t)e * P.S + k];
  }
  P.s_next[i0] = -1;
  P.s_next[i1] = i0;
  P.head[e] = i1;
  P.first_invalid[e] = (unsigned long long)__double_as_longlong(10.0);  // VoxelEnvironment.cpp:261
  P.eflags[e] = 0u;
}

// indexed form: the endpoints of edge e are roadmap vertices whose FK already exists in a vertex
// pool; one warp per endpoint sample copies state, shape, point count and validity flags.
__global__ void edge_init_indexed_kernel(EdgePool P, int32_t E, const int64_t *__restrict__ pairs,
                                         const double *__restrict__ vstates, const double *__restrict__ vp,
                                         const int32_t *__restrict__ vnpts, const uint32_t *__restrict__ vflags,
                                         double *__restrict__ a_out, double *__restrict__ b_out) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= 2 * (int64_t)E) return;
  const int32_t e = (int32_t)(w >> 1), side = (int32_t)(w & 1);
  const int64_t v = pairs[2 * (int64_t)e + side];
  const int32_t smp = 2 * e + side;
  const int np = vnpts[v];
  const double *src = vp + v * (int64_t)P.cap_pts * 3;
  double *dst = P.s_p + (int64_t)smp * P.cap_pts * 3;
  for (int i = lane; i < np * 3; i += 32) dst[i] = src[i];
  double *eo = (side ? b_out : a_out) + (int64_t)e * P.S;
  for (int k = lane; k < P.S; k += 32) {
    const double x = vstates[v * P.S + k];
    P.s_state[(int64_t)smp * P.S + k] = x;
    eo[k] = x;
  }
  if (lane == 0) {
    P.s_edge[smp] = e;
    P.s_t[smp] = side ? 1.0 : 0.0;
    P.s_npts[smp] = np;
    P.s_flags[smp] = vflags[v];
    P.s_next[smp] = side ? (smp - 1) : -1;
    if (side) {
      P.head[e] = smp;
      P.first_invalid[e] = (unsigned long long)__double_as_longlong(10.0);
      P.eflags[e] = 0u;
    }
  }
}

// rel_threshold = 1 / validSegmentCount(a, b) (VoxelBackboneMotionValidator.cpp:55-56), the same
// arithmetic as irt_valid_segment_count (this file is compiled without FMA contraction)
__global__ void edge_threshold_kernel(EdgePool P, int32_t E, const double *__restrict__ a,
                                      const double *__restrict__ b, double len_t, double len_r,
                                      double len_s, double *__restrict__ thr) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const double pi = 3.14159265358979323846;
  const double *ea = a + (int64_t)e * P.S, *eb = b + (int64_t)e * P.S;
  double d2 = 0;
  for (int i = 0; i < P.N; i++) d2 += (ea[i] - eb[i]) * (ea[i] - eb[i]);
  unsigned sc = (unsigned)ceil(sqrt(d2) / len_t);
  int idx = P.N;
  if (P.enable_rotation) {
    double d = fabs(ea[idx] - eb[idx]);
    d = (d > pi) ? 2.0 * pi - d : d;
    const unsigned c = (unsigned)ceil(d / len_r);
    if (c > sc) sc = c;
    idx++;
  }
  if (P.enable_retraction) {
    const double d = sqrt((ea[idx] - eb[idx]) * (ea[idx] - eb[idx]));
    const unsigned c = (unsigned)ceil(d / len_s);
    if (c > sc) sc = c;
  }
  thr[e] = 1.0 / double(sc);
}

// after FK: invalid samples lower first_invalid_t of their edge (VoxelEnvironment.cpp:266-268)
__global__ void edge_mark_kernel(EdgePool P, int32_t lo, int32_t hi) {
  const int32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  if (P.s_flags[i] & INVALID_MASK)
    atomicMin(&P.first_invalid[P.s_edge[i]], (unsigned long long)__double_as_longlong(P.s_t[i]));
}

// OMPL compound interpolate restated (RealVector linear, SO2 shortest arc + wrap), the `interp`
// lambda of VoxelBackboneMotionValidator.cpp:58-66
__device__ void interpolate_state(const EdgePool &P, const double *a, const double *b, double t,
                                  double *out) {
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < P.N; i++) out[i] = a[i] + (b[i] - a[i]) * t;
  int idx = P.N;
  if (P.enable_rotation) {
    double diff = b[idx] - a[idx];
    if (fabs(diff) <= pi) {
      out[idx] = a[idx] + diff * t;
    } else {
      if (diff > 0.0) diff = 2.0 * pi - diff;
      else diff = -2.0 * pi - diff;
      double v = a[idx] - diff * t;
      if (v > pi) v -= 2.0 * pi;
      else if (v < -pi) v += 2.0 * pi;
      out[idx] = v;
    }
    idx++;
  }
  if (P.enable_retraction) out[idx] = a[idx] + (b[idx] - a[idx]) * t;
}

// one thread per open interval: VoxelEnvironment.cpp:369-384
__global__ void edge_split_kernel(EdgePool P, const Interval *__restrict__ cur, int32_t ncur,
                                  int32_t *__restrict__ n_samples, Pending *__restrict__ pend,
                                  int32_t *__restrict__ n_pend) {
  const int32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= ncur) return;
  const Interval iv = cur[q];
  const double ta = P.s_t[iv.ia], tb = P.s_t[iv.ib];
  if ((tb - ta) <= P.thr[iv.edge]) return;
  const double fi = __longlong_as_double((long long)P.first_invalid[iv.edge]);
  if (fi <= ta) return;
  const int32_t m = atomicAdd(n_samples, 1);
  if (m >= P.cap_samples) {
    atomicOr(&P.eflags[iv.edge], IRT_FLAG_CAPACITY);
    return;
  }
  const double tm = (ta + tb) / 2;
  P.s_edge[m] = iv.edge;
  P.s_t[m] = tm;
  interpolate_state(P, P.a + (int64_t)iv.edge * P.S, P.b + (int64_t)iv.edge * P.S, tm,
                    P.s_state + (int64_t)m * P.S);
  P.s_next[m] = atomicExch(&P.head[iv.edge], m);
  const int32_t k = atomicAdd(n_pend, 1);
  pend[k] = Pending{iv.edge, iv.ia, m, iv.ib};
}

// find_cell -- collision/VoxelOctree.cpp:309-317 (+domain_check :1511-1521); false = domain error
__device__ __forceinline__ bool find_cell(const GridDev &g, const D3 &p, long long *c) {
  if (p.x < g.lo[0] || g.hi[0] < p.x) return false;
  if (p.y < g.lo[1] || g.hi[1] < p.y) return false;
  if (p.z < g.lo[2] || g.hi[2] < p.z) return false;
  c[0] = (long long)((p.x - g.lo[0]) / g.d[0]);
  c[1] = (long long)((p.y - g.lo[1]) / g.d[1]);
  c[2] = (long long)((p.z - g.lo[2]) / g.d[2]);
  return true;
}

// should_subdivide(a, b) -- VoxelEnvironment.cpp:304-341, warp-cooperative (lanes over points,
// scanned from the tip like the reference so the first event found is the same one)
__device__ bool should_subdivide(const GridDev &g, const EdgePool &P, int32_t ia, int32_t ib,
                                 int32_t edge, int lane) {
  if (P.s_flags[ia] & INVALID_MASK) return false;
  const int na = P.s_npts[ia], nb = P.s_npts[ib];
  if (na + 1 < nb || na > nb + 1) return true;
  const int Pn = min(na, nb);
  const double *pa = P.s_p + (int64_t)ia * P.cap_pts * 3, *pb = P.s_p + (int64_t)ib * P.cap_pts * 3;
  for (int top = Pn - 1; top >= 0; top -= 32) {
    const int i = top - lane;
    int ev = 0;  // 1 = far apart, 2 = domain error
    if (i >= 0) {
      long long s[3], e[3];
      const D3 qa = rotate_pt(g, pa + 3 * i), qb = rotate_pt(g, pb + 3 * i);
      if (!find_cell(g, qa, s) || !find_cell(g, qb, e)) {
        ev = 2;
      } else {
        const long long dx = llabs(s[0] - e[0]), dy = llabs(s[1] - e[1]), dz = llabs(s[2] - e[2]);
        if (dx > 1 || dy > 1 || dz > 1) ev = 1;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, ev != 0);
    if (m) {
      const int first = __ffs(m) - 1;  // lane 0 holds the highest index
      const int fev = __shfl_sync(0xffffffffu, ev, first);
      if (fev == 2) {
        if (lane == 0) atomicOr(&P.eflags[edge], IRT_FLAG_OUT_OF_DOMAIN);
        return false;
      }
      return true;
    }
  }
  return false;
}

// one warp per candidate: round 0 tests (2e, 2e+1); later rounds test both halves of a bisected
// interval (VoxelEnvironment.cpp:350-353,386-397)
__global__ void edge_subdivide_kernel(const GridDev g, EdgePool P, const Pending *__restrict__ pend,
                                      int32_t npend, int32_t E_round0, Interval *__restrict__ next,
                                      int32_t *__restrict__ n_next, int32_t cap_next) {
  const int lane = threadIdx.x & 31;
  const int32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int32_t total = pend ? npend : E_round0;
  if (w >= total) return;
  if (!pend) {
    const int32_t e = w;
    if (should_subdivide(g, P, 2 * e, 2 * e + 1, e, lane) && lane == 0) {
      const int32_t k = atomicAdd(n_next, 1);
      if (k < cap_next) next[k] = Interval{e, 2 * e, 2 * e + 1};
      else atomicOr(&P.eflags[e], IRT_FLAG_CAPACITY);
    }
    return;
  }
  const Pending pd = pend[w];
  const bool distal = should_subdivide(g, P, pd.im, pd.ib, pd.edge, lane);
  const bool proximal = should_subdivide(g, P, pd.ia, pd.im, pd.edge, lane);
  if (lane == 0) {
    if (distal) {
      const int32_t k = atomicAdd(n_next, 1);
      if (k < cap_next) next[k] = Interval{pd.edge, pd.im, pd.ib};
      else atomicOr(&P.eflags[pd.edge], IRT_FLAG_CAPACITY);
    }
    if (proximal) {
      const int32_t k = atomicAdd(n_next, 1);
      if (k < cap_next) next[k] = Interval{pd.edge, pd.ia, pd.im};
      else atomicOr(&P.eflags[pd.edge], IRT_FLAG_CAPACITY);
    }
  }
}

__global__ void edge_finish_kernel(EdgePool P, int32_t E, double *__restrict__ tlimit,
                                   uint32_t *__restrict__ flags_out) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const double fi = __longlong_as_double((long long)P.first_invalid[e]);
  tlimit[e] = fi;
  uint32_t f = P.eflags[e];
  if (!(5.0 < fi)) f |= IRT_FLAG_PARTIAL;  // VoxelEnvironment.cpp:438
  flags_out[e] = f;
}

// bump allocator over the context's grow-only K2 arena.  Layouts are described once by a
// lambda and run twice: a dry run to size the arena, then the real carve-up.
struct Arena {
  char *base = nullptr;
  size_t cap = 0, used = 0;
  bool dry = true;
  template <typename T>
  bool alloc(T **out, size_t count) {
    const size_t bytes = ((count ? count : 1) * sizeof(T) + 255) & ~(size_t)255;
    if (dry) { *out = nullptr; used += bytes; return true; }
    if (used + bytes > cap) return false;
    *out = reinterpret_cast<T *>(base + used);
    used += bytes;
    return true;
  }
};
template <typename F>
int arena_layout(irt_ctx *ctx, const F &layout) {
  Arena dry;
  layout(dry);
  char *base = (char *)ctx_arena(ctx, dry.used);
  if (!base) return irt_fail(ctx, IRT_ERR_CUDA, "K2 arena allocation of %zu bytes failed", dry.used);
  Arena real;
  real.base = base; real.cap = dry.used; real.dry = false;
  if (!layout(real)) return irt_fail(ctx, IRT_ERR_CUDA, "K2 arena layout failed");
  return IRT_OK;
}

constexpr int64_t RS_CHUNK_SETS = 262144;  // sets rasterised per launch (slot scratch = 3 KiB per set)

struct RasterScratch {
  uint32_t *slot_keys = nullptr, *big_keys = nullptr;
  uint64_t *slot_bits = nullptr, *big_bits = nullptr;
  uint32_t *counts = nullptr;
  uint64_t *scan_tmp = nullptr;
  int32_t *ovf_list = nullptr, *ovf_count = nullptr, *ovf_slot = nullptr;
  template <typename A>
  bool layout(A &a, int64_t chunk) {
    const int64_t ntiles = (chunk + SCAN_TILE - 1) / SCAN_TILE;
    return a.alloc(&slot_keys, (size_t)chunk * RS_SLOT) && a.alloc(&slot_bits, (size_t)chunk * RS_SLOT) &&
           a.alloc(&big_keys, (size_t)RS_MAX_OVF * RS_BIGSLOT) && a.alloc(&big_bits, (size_t)RS_MAX_OVF * RS_BIGSLOT) &&
           a.alloc(&counts, (size_t)chunk) && a.alloc(&scan_tmp, (size_t)ntiles + 2) &&
           a.alloc(&ovf_list, (size_t)RS_MAX_OVF) && a.alloc(&ovf_count, 64) && a.alloc(&ovf_slot, (size_t)chunk);
  }
};

// Rasterise sets [0, nsets) of one sample pool and APPEND them to the store: the sets become
// store sets [set_base, set_base + nsets) and their leaves start at leaf_base.
int raster_append(irt_ctx *ctx, const GridDev &g, const double *d_pts, const int32_t *d_npts,
                  int cap_pts, const int32_t *d_heads, const int32_t *d_next, const double *d_st,
                  const double *d_tlimit, int64_t nsets, double *d_tlast, int32_t *d_nsamples,
                  uint32_t *d_setflags, RasterScratch &rs, irt_setstore *store, int64_t set_base,
                  uint64_t leaf_base, uint64_t *leaf_total_out, cudaStream_t st) {
  auto k_small = swept_voxel_raster_kernel<RS_HLOG_SMALL, RS_WARPS>;
  auto k_big = swept_voxel_raster_kernel<RS_HLOG_BIG, 1>;
  const int smem_small = RS_WARPS * rs_warp_bytes(RS_HLOG_SMALL), smem_big = rs_warp_bytes(RS_HLOG_BIG);
  IRT_CUDA(ctx, cudaFuncSetAttribute(k_big, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_big));
  uint64_t running = leaf_base;
  for (int64_t c0 = 0; c0 < nsets; c0 += RS_CHUNK_SETS) {
    const int64_t m = (nsets - c0 < RS_CHUNK_SETS) ? (nsets - c0) : RS_CHUNK_SETS;
    int64_t blocks = (m + RS_WARPS - 1) / RS_WARPS;
    const int64_t max_blocks = (int64_t)ctx->sm_count * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    IRT_CUDA(ctx, cudaMemsetAsync(rs.ovf_count, 0, 4, st));
    IRT_CUDA(ctx, cudaMemsetAsync(rs.ovf_slot, 0, (size_t)m * 4, st));
    k_small<<<(unsigned)blocks, RS_WARPS * 32, smem_small, st>>>(
        g, d_pts, d_npts, cap_pts, d_heads + c0, d_next, d_st, d_tlimit ? d_tlimit + c0 : nullptr, m, rs.counts,
        d_tlast ? d_tlast + c0 : nullptr, d_nsamples ? d_nsamples + c0 : nullptr, rs.slot_keys, rs.slot_bits,
        d_setflags ? d_setflags + c0 : nullptr, rs.ovf_list, rs.ovf_count, rs.ovf_slot);
    IRT_LAUNCHED(ctx);
    // sets that outgrew the small table (normally none): big-table pass over the device-side list
    k_big<<<(unsigned)ctx->sm_count, 32, smem_big, st>>>(
        g, d_pts, d_npts, cap_pts, d_heads + c0, d_next, d_st, d_tlimit ? d_tlimit + c0 : nullptr, m, rs.counts,
        d_tlast ? d_tlast + c0 : nullptr, d_nsamples ? d_nsamples + c0 : nullptr, rs.big_keys, rs.big_bits,
        d_setflags ? d_setflags + c0 : nullptr, rs.ovf_list, rs.ovf_count, rs.ovf_slot);
    IRT_LAUNCHED(ctx);
    IRT_CUDA(ctx, cudaGetLastError());
    uint64_t total = 0;
    int rc = exclusive_scan(ctx, rs.counts, m, running, store->d_offsets + set_base + c0, rs.scan_tmp, &total, st);
    if (rc) return rc;
    rc = setstore_grow_blocks(ctx, store, (int64_t)(running + total), (int64_t)running, st);
    if (rc) return rc;
    const int T = 256;
    raster_gather_kernel<<<(unsigned)((m * 32 + T - 1) / T), T, 0, st>>>(
        rs.slot_keys, rs.slot_bits, rs.big_keys, rs.big_bits, rs.ovf_slot, rs.counts,
        store->d_offsets + set_base + c0, m, store->d_keys, store->d_bits);
    IRT_LAUNCHED(ctx);
    IRT_CUDA(ctx, cudaGetLastError());
    running += total;
  }
  *leaf_total_out = running - leaf_base;
  return IRT_OK;
}

}  // namespace

// ===============================================================================================
// C ABI
// ===============================================================================================
extern "C" {

uint32_t irt_valid_segment_count(const irt_robot_desc *rb, const irt_space *sp, const double *a,
                                 const double *b) {
  // OMPL 1.5 documented behaviour: StateSpace::validSegmentCount = ceil(distance /
  // longestValidSegment), longestValidSegment = maximumExtent * fraction (fractions set in
  // motion-planning/Problem.cpp:118-144); CompoundStateSpace takes the max over subspaces.
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
    unsigned c = (unsigned)std::ceil(d / len_r);
    if (c > sc) sc = c;
    idx++;
  }
  if (rb->enable_retraction) {
    const double len_s = rb->L * std::fmin(0.01, sp->min_retraction_change / rb->L);
    double d = std::sqrt((a[idx] - b[idx]) * (a[idx] - b[idx]));
    unsigned c = (unsigned)std::ceil(d / len_s);
    if (c > sc) sc = c;
  }
  return sc;
}

static int store_begin(irt_ctx *ctx, irt_setstore *store, int64_t n, int64_t est_blocks, cudaStream_t st) {
  int rc = setstore_reserve(ctx, store, n, est_blocks);
  if (rc) return rc;
  if (n == 0) IRT_CUDA(ctx, cudaMemsetAsync(store->d_offsets, 0, 8, st));
  return IRT_OK;
}

static int check_dl_vs_grid(irt_ctx *ctx, const irt_robot *rb, const GridDev &g) {
  // VoxelBackboneValidityChecker ctor: dL must not exceed the largest voxel side (.h:37-45)
  const double dmax = std::fmax(g.d[0], std::fmax(g.d[1], g.d[2]));
  if (rb->desc.dL > dmax)
    return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT,
                    "robot.specs.dL is larger than expected by VoxelBackboneValidityChecker (%g > %g)",
                    rb->desc.dL, dmax);
  return IRT_OK;
}

int irt_voxelize_vertices(irt_ctx *ctx, const irt_robot *rb, const double *states, int state_size,
                          int64_t n, irt_setstore *store, uint32_t *flags, double *tips) {
  if (!ctx || !rb || !store || n < 0 || (n > 0 && !states)) return IRT_ERR_INVALID_ARGUMENT;
  if (state_size != rb->state_size)
    return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "State is not the right size (%d != %d)",
                    state_size, rb->state_size);
  const GridDev &g = store->gd;
  int rc = check_dl_vs_grid(ctx, rb, g);
  if (rc) return rc;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  rc = store_begin(ctx, store, n, n * 24, st);
  if (rc) return rc;
  if (n == 0) return setstore_finalize(ctx, store, 0, 0, st);
  const int cap = rb->max_points, S = state_size;
  const int64_t chunk = std::min<int64_t>(n, 262144);
  double *d_states = nullptr, *d_p = nullptr, *d_tip = nullptr;
  int32_t *d_npts = nullptr, *d_heads = nullptr;
  uint32_t *d_flags = nullptr;
  RasterScratch rs;
  rc = arena_layout(ctx, [&](Arena &a) {
    return a.alloc(&d_states, (size_t)chunk * S) && a.alloc(&d_p, (size_t)chunk * cap * 3) &&
           a.alloc(&d_tip, (size_t)chunk * 3) && a.alloc(&d_npts, (size_t)chunk) &&
           a.alloc(&d_heads, (size_t)chunk) && a.alloc(&d_flags, (size_t)chunk) &&
           rs.layout(a, std::min<int64_t>(chunk, RS_CHUNK_SETS));
  });
  if (rc) return rc;
  uint64_t running = 0;
  const int T = 256;
  for (int64_t off = 0; off < n; off += chunk) {
    const int64_t m = std::min<int64_t>(chunk, n - off);
    IRT_CUDA(ctx, cudaMemcpyAsync(d_states, states + off * S, (size_t)m * S * 8, cudaMemcpyHostToDevice, st));
    irt_fk_outputs o;
    std::memset(&o, 0, sizeof(o));
    o.p = d_p; o.npts = d_npts; o.flags = d_flags; o.tip = d_tip;
    rc = fk_launch(ctx, rb, d_states, m, cap, o, nullptr, st);
    if (rc) return rc;
    rc = self_collision_launch(ctx, rb, d_p, d_npts, m, cap, d_flags, st);
    if (rc) return rc;
    vertex_heads_kernel<<<(unsigned)((m + T - 1) / T), T, 0, st>>>(d_flags, m, d_heads);
    IRT_LAUNCHED(ctx);
    uint64_t total = 0;
    rc = raster_append(ctx, g, d_p, d_npts, cap, d_heads, nullptr, nullptr, nullptr, m, nullptr, nullptr,
                       d_flags, rs, store, off, running, &total, st);
    if (rc) return rc;
    running += total;
    if (flags) IRT_CUDA(ctx, cudaMemcpyAsync(flags + off, d_flags, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    if (tips) IRT_CUDA(ctx, cudaMemcpyAsync(tips + off * 3, d_tip, (size_t)m * 24, cudaMemcpyDeviceToHost, st));
    IRT_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return setstore_finalize(ctx, store, n, (int64_t)running, st);
}

int irt_voxelize_shapes(irt_ctx *ctx, const double *p, const int32_t *npts, int cap_pts, int64_t n,
                        irt_setstore *store) {
  if (!ctx || !store || n < 0 || cap_pts < 1 || (n > 0 && (!p || !npts))) return IRT_ERR_INVALID_ARGUMENT;
  for (int64_t i = 0; i < n; i++)
    if (npts[i] < 0 || npts[i] > cap_pts)
      return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "npts[%lld] out of range", (long long)i);
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const GridDev &g = store->gd;
  int rc = store_begin(ctx, store, n, n * 24, st);
  if (rc) return rc;
  if (n == 0) return setstore_finalize(ctx, store, 0, 0, st);
  const int64_t chunk = std::min<int64_t>(n, 262144);
  double *d_p = nullptr;
  int32_t *d_npts = nullptr, *d_heads = nullptr;
  uint32_t *d_flags = nullptr;
  RasterScratch rs;
  rc = arena_layout(ctx, [&](Arena &a) {
    return a.alloc(&d_p, (size_t)chunk * cap_pts * 3) && a.alloc(&d_npts, (size_t)chunk) &&
           a.alloc(&d_heads, (size_t)chunk) && a.alloc(&d_flags, (size_t)chunk) &&
           rs.layout(a, std::min<int64_t>(chunk, RS_CHUNK_SETS));
  });
  if (rc) return rc;
  uint64_t running = 0;
  const int T = 256;
  for (int64_t off = 0; off < n; off += chunk) {
    const int64_t m = std::min<int64_t>(chunk, n - off);
    IRT_CUDA(ctx, cudaMemcpyAsync(d_p, p + off * cap_pts * 3, (size_t)m * cap_pts * 24, cudaMemcpyHostToDevice, st));
    IRT_CUDA(ctx, cudaMemcpyAsync(d_npts, npts + off, (size_t)m * 4, cudaMemcpyHostToDevice, st));
    IRT_CUDA(ctx, cudaMemsetAsync(d_flags, 0, (size_t)m * 4, st));
    vertex_heads_kernel<<<(unsigned)((m + T - 1) / T), T, 0, st>>>(d_flags, m, d_heads);
    IRT_LAUNCHED(ctx);
    uint64_t total = 0;
    rc = raster_append(ctx, g, d_p, d_npts, cap_pts, d_heads, nullptr, nullptr, nullptr, m, nullptr, nullptr,
                       d_flags, rs, store, off, running, &total, st);
    if (rc) return rc;
    running += total;
    IRT_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return setstore_finalize(ctx, store, n, (int64_t)running, st);
}

}  // extern "C"

// a, b: per-edge endpoint states (host) -- or, indexed form: vstates[nv][S] + pairs[n][2]
static int voxelize_edges_core(irt_ctx *ctx, const irt_robot *rb, const irt_space *space, const double *a,
                               const double *b, const double *vstates, int64_t nv, const int64_t *pairs,
                               int state_size, int64_t n, const irt_env *env, irt_setstore *store,
                               uint32_t *flags, double *t_last, int32_t *nsamples) {
  const bool indexed = pairs != nullptr;
  if (env && store && env->grid.Ng != store->grid.Ng)
    return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "voxel dimension mismatch (%d != %d)", env->grid.Ng,
                    store->grid.Ng);
  if (!ctx || !rb || !space || !store || n < 0) return IRT_ERR_INVALID_ARGUMENT;
  if (!indexed && n > 0 && (!a || !b)) return IRT_ERR_INVALID_ARGUMENT;
  if (indexed && (nv < 0 || (nv > 0 && !vstates))) return IRT_ERR_INVALID_ARGUMENT;
  if (indexed)
    for (int64_t i = 0; i < 2 * n; i++)
      if (pairs[i] < 0 || pairs[i] >= nv)
        return irt_fail(ctx, IRT_ERR_OUT_OF_RANGE, "edge endpoint %lld outside [0,%lld)", (long long)pairs[i], (long long)nv);
  if (state_size != rb->state_size)
    return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "State is not the right size (%d != %d)",
                    state_size, rb->state_size);
  const GridDev &g = store->gd;
  int rc = check_dl_vs_grid(ctx, rb, g);
  if (rc) return rc;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int S = state_size, cap = rb->max_points;
  if (n > (int64_t)0x3fffffff) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "too many edges");
  rc = store_begin(ctx, store, n, n * 32, st);
  if (rc) return rc;
  if (n == 0) return setstore_finalize(ctx, store, 0, 0, st);
  Trace tr(st);

  // chunk sizing: a sample costs cap*24 + S*8 + 72 bytes; budget ~6 GiB or 40% of free memory
  size_t free_b = 0, total_b = 0;
  IRT_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  free_b += ctx->arena_bytes;  // the arena is ours to reuse
  const size_t per_sample = (size_t)cap * 24 + (size_t)S * 8 + 72;
  size_t budget = (size_t)16 << 30;
  if (budget > free_b * 2 / 5) budget = free_b * 2 / 5;
  int64_t cap_samples = (int64_t)(budget / per_sample);
  if (cap_samples > 0x3fffffff) cap_samples = 0x3fffffff;
  // typical roadmap edges need 4-15 FK samples; a chunk whose pool overflows is split and redone
  int64_t chunk = cap_samples / 12;
  if (chunk < 256) { chunk = 256; if (cap_samples < chunk * 8) cap_samples = chunk * 8; }
  if (chunk > n) {
    chunk = n;
    cap_samples = std::min<int64_t>(cap_samples, std::max<int64_t>(chunk * 64, 4096));
  }

  // longest valid segment lengths of the three subspaces (Problem.cpp:118-144), as in
  // irt_valid_segment_count; the per-edge count itself is evaluated on the device
  double len_t, len_r, len_s;
  {
    double ext2 = 0;
    for (int i = 0; i < rb->desc.n_tendons; i++) ext2 += rb->desc.max_tension[i] * rb->desc.max_tension[i];
    const double tendon_extent = std::sqrt(ext2);
    len_t = tendon_extent * (space->min_tension_change / tendon_extent);
    len_r = M_PI * (space->min_rotation_change / (2 * M_PI));
    len_s = rb->desc.L * std::fmin(0.01, space->min_retraction_change / rb->desc.L);
  }

  EdgePool P;
  std::memset(&P, 0, sizeof(P));
  double *d_a = nullptr, *d_b = nullptr, *d_thr = nullptr, *d_tlimit = nullptr, *d_tlast = nullptr;
  int32_t *d_nsamp_set = nullptr, *d_counters = nullptr;
  uint32_t *d_flags_out = nullptr;
  Interval *d_q0 = nullptr, *d_q1 = nullptr;
  Pending *d_pend = nullptr;
  RasterScratch rs;
  // indexed form: FK of every roadmap vertex once; edges gather their endpoint shapes from it
  double *d_vstates = nullptr, *d_vp = nullptr;
  int32_t *d_vnpts = nullptr;
  uint32_t *d_vflags = nullptr;
  int64_t *d_pairs = nullptr;
  rc = arena_layout(ctx, [&](Arena &A) {
    if (indexed &&
        !(A.alloc(&d_vstates, (size_t)nv * S) && A.alloc(&d_vp, (size_t)nv * cap * 3) &&
          A.alloc(&d_vnpts, (size_t)nv) && A.alloc(&d_vflags, (size_t)nv) && A.alloc(&d_pairs, (size_t)chunk * 2)))
      return false;
    return A.alloc(&d_a, (size_t)chunk * S) && A.alloc(&d_b, (size_t)chunk * S) &&
           A.alloc(&d_thr, (size_t)chunk) && A.alloc(&d_tlimit, (size_t)chunk) &&
           A.alloc(&d_tlast, (size_t)chunk) && A.alloc(&d_nsamp_set, (size_t)chunk) &&
           A.alloc(&d_flags_out, (size_t)chunk) && A.alloc(&P.first_invalid, (size_t)chunk) &&
           A.alloc(&P.head, (size_t)chunk) && A.alloc(&P.eflags, (size_t)chunk) &&
           A.alloc(&P.s_edge, (size_t)cap_samples) && A.alloc(&P.s_next, (size_t)cap_samples) &&
           A.alloc(&P.s_npts, (size_t)cap_samples) && A.alloc(&P.s_flags, (size_t)cap_samples) &&
           A.alloc(&P.s_t, (size_t)cap_samples) && A.alloc(&P.s_state, (size_t)cap_samples * S) &&
           A.alloc(&P.s_p, (size_t)cap_samples * cap * 3) && A.alloc(&d_q0, (size_t)cap_samples) &&
           A.alloc(&d_q1, (size_t)cap_samples) && A.alloc(&d_pend, (size_t)cap_samples) &&
           A.alloc(&d_counters, 8) && rs.layout(A, std::min<int64_t>(chunk, RS_CHUNK_SETS));
  });
  if (rc) return rc;
  tr.point("thr + pool layout", cap_samples);
  P.a = d_a; P.b = d_b; P.thr = d_thr;
  P.cap_samples = (int32_t)cap_samples;
  P.S = S; P.N = rb->desc.n_tendons; P.cap_pts = cap;
  P.enable_rotation = rb->desc.enable_rotation ? 1 : 0;
  P.enable_retraction = rb->desc.enable_retraction ? 1 : 0;

  uint64_t grand_total = 0;
  const int T = 256;
  if (indexed && nv > 0) {
    IRT_CUDA(ctx, cudaMemcpyAsync(d_vstates, vstates, (size_t)nv * S * 8, cudaMemcpyHostToDevice, st));
    for (int64_t v0 = 0; v0 < nv; v0 += 1048576) {
      const int64_t m = std::min<int64_t>(1048576, nv - v0);
      irt_fk_outputs o;
      std::memset(&o, 0, sizeof(o));
      o.p = d_vp + v0 * cap * 3; o.npts = d_vnpts + v0; o.flags = d_vflags + v0;
      rc = fk_launch(ctx, rb, d_vstates + v0 * S, m, cap, o, nullptr, st);
      if (rc) return rc;
      rc = self_collision_launch(ctx, rb, o.p, o.npts, m, cap, o.flags, st);
      if (rc) return rc;
      if (env) {
        sample_env_collision_kernel<<<(unsigned)((m * 32 + T - 1) / T), T, 0, st>>>(
            g, o.p, o.npts, cap, m, env->d_blocks, env->d_occ, o.flags);
        IRT_LAUNCHED(ctx);
      }
    }
    tr.point("vertex fk", nv);
  }
  // work list of edge ranges in index order; a range whose sample pool overflows is split in two
  std::vector<std::pair<int64_t, int64_t>> work;
  for (int64_t off = n - ((n - 1) % chunk + 1); off >= 0; off -= chunk)
    work.emplace_back(off, std::min<int64_t>(chunk, n - off));
  while (!work.empty()) {
    const int64_t off = work.back().first;
    const int32_t E = (int32_t)work.back().second;
    work.pop_back();
    bool pool_overflow = false;
    IRT_CUDA(ctx, cudaMemsetAsync(d_counters, 0, 32, st));
    if (indexed) {
      IRT_CUDA(ctx, cudaMemcpyAsync(d_pairs, pairs + 2 * off, (size_t)E * 16, cudaMemcpyHostToDevice, st));
      const int64_t threads = (int64_t)E * 2 * 32;
      edge_init_indexed_kernel<<<(unsigned)((threads + T - 1) / T), T, 0, st>>>(
          P, E, d_pairs, d_vstates, d_vp, d_vnpts, d_vflags, d_a, d_b);
    } else {
      IRT_CUDA(ctx, cudaMemcpyAsync(d_a, a + off * S, (size_t)E * S * 8, cudaMemcpyHostToDevice, st));
      IRT_CUDA(ctx, cudaMemcpyAsync(d_b, b + off * S, (size_t)E * S * 8, cudaMemcpyHostToDevice, st));
      edge_init_kernel<<<(E + T - 1) / T, T, 0, st>>>(P, E);
    }
    IRT_LAUNCHED(ctx);
    edge_threshold_kernel<<<(E + T - 1) / T, T, 0, st>>>(P, E, d_a, d_b, len_t, len_r, len_s, d_thr);
    IRT_LAUNCHED(ctx);
    int32_t n_samples = 2 * E;
    IRT_CUDA(ctx, cudaMemcpyAsync(d_counters, &n_samples, 4, cudaMemcpyHostToDevice, st));

    auto run_fk = [&](int32_t lo, int32_t hi) -> int {
      if (hi <= lo) return IRT_OK;
      irt_fk_outputs o;
      std::memset(&o, 0, sizeof(o));
      o.p = P.s_p + (int64_t)lo * cap * 3;
      o.npts = P.s_npts + lo;
      o.flags = P.s_flags + lo;
      int r = fk_launch(ctx, rb, P.s_state + (int64_t)lo * S, hi - lo, cap, o, nullptr, st);
      if (r) return r;
      r = self_collision_launch(ctx, rb, o.p, o.npts, hi - lo, cap, o.flags, st);
      if (r) return r;
      if (env) {
        sample_env_collision_kernel<<<(unsigned)(((int64_t)(hi - lo) * 32 + T - 1) / T), T, 0, st>>>(
            g, o.p, o.npts, cap, hi - lo, env->d_blocks, env->d_occ, o.flags);
        IRT_LAUNCHED(ctx);
      }
      edge_mark_kernel<<<(hi - lo + T - 1) / T, T, 0, st>>>(P, lo, hi);
      IRT_LAUNCHED(ctx);
      return IRT_OK;
    };
    tr.point("h2d + init");
    if (indexed) {  // endpoint shapes came from the vertex pool: only mark the invalid ones
      edge_mark_kernel<<<(n_samples + T - 1) / T, T, 0, st>>>(P, 0, n_samples);
      IRT_LAUNCHED(ctx);
    } else {
      rc = run_fk(0, n_samples);
      if (rc) return rc;
    }
    tr.point("round 0 fk", n_samples);
    Interval *cur = d_q0, *nxt = d_q1;
    int32_t *n_cur = d_counters + 2, *n_nxt = d_counters + 3;
    {
      const int64_t threads = (int64_t)E * 32;
      edge_subdivide_kernel<<<(unsigned)((threads + T - 1) / T), T, 0, st>>>(
          g, P, nullptr, 0, E, cur, n_cur, (int32_t)cap_samples);
      IRT_LAUNCHED(ctx);
    }
    for (int round = 0; round < 64; round++) {
      int32_t h_cnt[4];
      IRT_CUDA(ctx, cudaMemcpyAsync(h_cnt, d_counters, 16, cudaMemcpyDeviceToHost, st));
      IRT_CUDA(ctx, cudaStreamSynchronize(st));
      const int32_t ncur_raw = *(cur == d_q0 ? &h_cnt[2] : &h_cnt[3]);
      if (ncur_raw > (int32_t)cap_samples) pool_overflow = true;
      const int32_t ncur = std::min(ncur_raw, (int32_t)cap_samples);
      if (ncur == 0) break;
      const int32_t s_lo = std::min(h_cnt[0], (int32_t)cap_samples);
      IRT_CUDA(ctx, cudaMemsetAsync(d_counters + 1, 0, 4, st));  // n_pend
      IRT_CUDA(ctx, cudaMemsetAsync(n_nxt, 0, 4, st));
      edge_split_kernel<<<(ncur + T - 1) / T, T, 0, st>>>(P, cur, ncur, d_counters, d_pend, d_counters + 1);
      IRT_LAUNCHED(ctx);
      IRT_CUDA(ctx, cudaMemcpyAsync(h_cnt, d_counters, 8, cudaMemcpyDeviceToHost, st));
      IRT_CUDA(ctx, cudaStreamSynchronize(st));
      const int32_t s_hi = std::min(h_cnt[0], (int32_t)cap_samples);
      const int32_t npend = h_cnt[1];
      if (h_cnt[0] > (int32_t)cap_samples) {  // keep the counter in range
        int32_t capv = (int32_t)cap_samples;
        IRT_CUDA(ctx, cudaMemcpyAsync(d_counters, &capv, 4, cudaMemcpyHostToDevice, st));
        pool_overflow = true;
      }
      rc = run_fk(s_lo, s_hi);
      if (rc) return rc;
      if (npend > 0) {
        const int64_t threads = (int64_t)npend * 32;
        edge_subdivide_kernel<<<(unsigned)((threads + T - 1) / T), T, 0, st>>>(
            g, P, d_pend, npend, 0, nxt, n_nxt, (int32_t)cap_samples);
        IRT_LAUNCHED(ctx);
      }
      std::swap(cur, nxt);
      std::swap(n_cur, n_nxt);
      tr.point("bisection round", s_hi - s_lo);
    }
    if (pool_overflow && E > 256) {  // redo this range as two halves (nothing was appended yet)
      const int64_t h1 = E / 2;
      work.emplace_back(off + h1, E - h1);
      work.emplace_back(off, h1);
      tr.point("pool overflow -> split");
      continue;
    }
    edge_finish_kernel<<<(E + T - 1) / T, T, 0, st>>>(P, E, d_tlimit, d_flags_out);
    IRT_LAUNCHED(ctx);
    // rasterise every sample below the first invalid t (VoxelEnvironment.cpp:406-422)
    uint64_t total = 0;
    rc = raster_append(ctx, g, P.s_p, P.s_npts, cap, P.head, P.s_next, P.s_t, d_tlimit, E, d_tlast,
                       d_nsamp_set, d_flags_out, rs, store, off, grand_total, &total, st);
    if (rc) return rc;
    tr.point("raster + scan + gather", (long long)total);
    if (flags) IRT_CUDA(ctx, cudaMemcpyAsync(flags + off, d_flags_out, (size_t)E * 4, cudaMemcpyDeviceToHost, st));
    if (t_last) IRT_CUDA(ctx, cudaMemcpyAsync(t_last + off, d_tlast, (size_t)E * 8, cudaMemcpyDeviceToHost, st));
    if (nsamples) IRT_CUDA(ctx, cudaMemcpyAsync(nsamples + off, d_nsamp_set, (size_t)E * 4, cudaMemcpyDeviceToHost, st));
    IRT_CUDA(ctx, cudaStreamSynchronize(st));
    grand_total += total;
    tr.point("chunk d2h");
  }
  rc = setstore_finalize(ctx, store, n, (int64_t)grand_total, st);
  tr.point("finalize");
  return rc;
}

extern "C" {

int irt_voxelize_edges(irt_ctx *ctx, const irt_robot *rb, const irt_space *space, const double *a,
                       const double *b, int state_size, int64_t n, irt_setstore *store,
                       uint32_t *flags, double *t_last, int32_t *nsamples) {
  return voxelize_edges_core(ctx, rb, space, a, b, nullptr, 0, nullptr, state_size, n, nullptr, store, flags,
                             t_last, nsamples);
}

int irt_voxelize_edges_until_invalid(irt_ctx *ctx, const irt_robot *rb, const irt_space *space,
                                     const double *a, const double *b, int state_size, int64_t n,
                                     const irt_env *env, irt_setstore *store, uint32_t *flags,
                                     double *t_last, int32_t *nsamples) {
  if (!env) return IRT_ERR_INVALID_ARGUMENT;
  return voxelize_edges_core(ctx, rb, space, a, b, nullptr, 0, nullptr, state_size, n, env, store, flags,
                             t_last, nsamples);
}

int irt_voxelize_edges_indexed(irt_ctx *ctx, const irt_robot *rb, const irt_space *space,
                               const double *vertex_states, int state_size, int64_t n_vertices,
                               const int64_t *pairs, int64_t n_edges, irt_setstore *store,
                               uint32_t *flags, double *t_last, int32_t *nsamples) {
  if (n_edges > 0 && !pairs) return IRT_ERR_INVALID_ARGUMENT;
  static const int64_t dummy[2] = {0, 0};
  return voxelize_edges_core(ctx, rb, space, nullptr, nullptr, vertex_states, n_vertices,
                             pairs ? pairs : dummy, state_size, n_edges, nullptr, store, flags, t_last, nsamples);
}

}  // extern "C"
