// env_prep.cu -- environment preparation on the device: morphology (below) and the voxelisation
// of Environment's points / spheres / capsules (Environment::voxelize, at the end of this file).
//
// Replaces VoxelOctree::dilate_6neighbor / dilate_27neighbor / dilate_sphere
// (collision/VoxelOctree.cpp:693-952) and remove_interior_6neighbor / _27neighbor (:533-689) on the
// dense Morton-ordered environment grid of voxel_check.cu.
//
// The reference dilates by a depth-limited DFS over the neighbour graph, at most four steps per
// pass; the cells it reaches are exactly the cells within `num` graph steps of an occupied cell,
// so the device path applies `num` single-step dilations (ping-pong between the grid and a
// scratch copy).  A shortest path between two in-grid cells never has to leave their bounding
// box, so clipping every step at the grid boundary gives the same set as the reference's
// clip-at-the-end.  The 27-neighbour list of the reference names (x+1,y+1,z+1) twice and never
// (x-1,y+1,z+1); that asymmetric 25-offset neighbourhood is reproduced.
//
// One thread per 4x4x4 leaf block: it loads the 3x3x3 block neighbourhood once and derives every
// shifted copy with per-axis shift/mask pairs, so a step is ~27 loads + ~400 integer ops per
// block and the whole 128^3 grid is one 2M-thread launch.
#include <vector>

#include "common.cuh"
#include "env_prims.h"

namespace {

__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__device__ __forceinline__ uint32_t compact3(uint32_t v) {
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030c30c3u;
  v = (v | (v >> 4)) & 0x0300f00fu;
  v = (v | (v >> 8)) & 0x030000ffu;
  v = (v | (v >> 16)) & 0x3ffu;
  return v;
}

// result bit (x,y,z) = occupancy of cell (x,y,z+dz) given the block itself (c) and its z-1 / z+1
// block neighbours (m / p); bit = x*16 + y*4 + z (VoxelOctree.h:318-329)
__device__ __forceinline__ uint64_t shift_z(uint64_t m, uint64_t c, uint64_t p, int dz) {
  if (dz > 0) return ((c >> 1) & 0x7777777777777777ull) | ((p & 0x1111111111111111ull) << 3);
  if (dz < 0) return ((c << 1) & 0xeeeeeeeeeeeeeeeeull) | ((m & 0x8888888888888888ull) >> 3);
  return c;
}
__device__ __forceinline__ uint64_t shift_y(uint64_t m, uint64_t c, uint64_t p, int dy) {
  if (dy > 0) return ((c >> 4) & 0x0fff0fff0fff0fffull) | ((p & 0x000f000f000f000full) << 12);
  if (dy < 0) return ((c << 4) & 0xfff0fff0fff0fff0ull) | ((m & 0xf000f000f000f000ull) >> 12);
  return c;
}
__device__ __forceinline__ uint64_t shift_x(uint64_t m, uint64_t c, uint64_t p, int dx) {
  if (dx > 0) return (c >> 16) | (p << 48);
  if (dx < 0) return (c << 16) | (m >> 48);
  return c;
}

// offsets: bit ((dx+1)*9 + (dy+1)*3 + (dz+1)) set -> the cell at +(dx,dy,dz) takes part.
// ERODE=false: dst = OR over offsets of src(c + d)            (out-of-grid = empty)
// ERODE=true : dst = src & ~(AND over offsets of src(c + d))  (out-of-grid = full)
template <bool ERODE>
__global__ void __launch_bounds__(128)
env_morph_kernel(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst, int Nb,
                 uint32_t offsets) {
  const uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
  if (key >= (uint32_t)Nb * Nb * Nb) return;
  const int bx = (int)compact3(key >> 2), by = (int)compact3(key >> 1), bz = (int)compact3(key);
  const uint64_t self = src[key];
  if (ERODE && self == 0ull) { dst[key] = 0ull; return; }
  const uint64_t outside = ERODE ? ~0ull : 0ull;
  uint64_t nb[3][3][3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int x = bx + i - 1;
    const uint32_t kx = spread3((uint32_t)x) << 2;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int y = by + j - 1;
      const uint32_t ky = spread3((uint32_t)y) << 1;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int z = bz + k - 1;
        const bool in = x >= 0 && x < Nb && y >= 0 && y < Nb && z >= 0 && z < Nb;
        nb[i][j][k] = (i == 1 && j == 1 && k == 1) ? self
                      : in                           ? src[kx | ky | spread3((uint32_t)z)]
                                                     : outside;
      }
    }
  }
  uint64_t acc = ERODE ? ~0ull : 0ull;
#pragma unroll
  for (int dz = -1; dz <= 1; dz++) {
    if (((offsets >> (dz + 1)) & 0x01249249u) == 0u) continue;  // no offset with this dz
    uint64_t zs[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) zs[i][j] = shift_z(nb[i][j][0], nb[i][j][1], nb[i][j][2], dz);
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
      uint64_t ys[3];
#pragma unroll
      for (int i = 0; i < 3; i++) ys[i] = shift_y(zs[i][0], zs[i][1], zs[i][2], dy);
#pragma unroll
      for (int dx = -1; dx <= 1; dx++) {
        if (!((offsets >> ((dx + 1) * 9 + (dy + 1) * 3 + (dz + 1))) & 1u)) continue;
        const uint64_t v = shift_x(ys[0], ys[1], ys[2], dx);
        acc = ERODE ? (acc & v) : (acc | v);
      }
    }
  }
  dst[key] = ERODE ? (self & ~acc) : acc;
}

constexpr uint32_t off_bit(int dx, int dy, int dz) {
  return 1u << ((dx + 1) * 9 + (dy + 1) * 3 + (dz + 1));
}
constexpr uint32_t OFF_6 = off_bit(0, 0, 0) | off_bit(-1, 0, 0) | off_bit(1, 0, 0) | off_bit(0, -1, 0) |
                           off_bit(0, 1, 0) | off_bit(0, 0, -1) | off_bit(0, 0, 1);
constexpr uint32_t OFF_27 = (1u << 27) - 1u;
// dilate_27neighbor reaches c from s when c - s is in the reference's list, i.e. every offset
// but (-1,+1,+1); the kernel gathers, so it reads src(c + d) for every d but (+1,-1,-1)
constexpr uint32_t OFF_27_DILATE_GATHER = OFF_27 & ~off_bit(1, -1, -1);

int env_morph(irt_ctx *ctx, irt_env *env, bool erode, uint32_t offsets, int num) {
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  if (num <= 0) return IRT_OK;
  const size_t bytes = (size_t)env->n_blocks_total * 8;
  uint64_t *tmp = (uint64_t *)ctx_scratch(ctx, bytes);
  if (!tmp) return irt_fail(ctx, IRT_ERR_CUDA, "scratch alloc failed");
  cudaStream_t st = ctx->stream;
  const int T = 128;
  const unsigned nblk = (unsigned)((env->n_blocks_total + T - 1) / T);
  uint64_t *a = env->d_blocks, *b = tmp;
  for (int it = 0; it < num; it++) {
    if (erode) env_morph_kernel<true><<<nblk, T, 0, st>>>(a, b, env->gd.Nb, offsets);
    else env_morph_kernel<false><<<nblk, T, 0, st>>>(a, b, env->gd.Nb, offsets);
    IRT_LAUNCHED(ctx);
    uint64_t *t = a; a = b; b = t;
  }
  IRT_CUDA(ctx, cudaGetLastError());
  if (a != env->d_blocks)
    IRT_CUDA(ctx, cudaMemcpyAsync(env->d_blocks, a, bytes, cudaMemcpyDeviceToDevice, st));
  int rc = env_rebuild_occ(ctx, env, st);
  if (rc) return rc;
  IRT_CUDA(ctx, cudaStreamSynchronize(st));
  return IRT_OK;
}

// ---- Environment::voxelize's primitives (motion-planning/Environment.cpp:62-74) -------------------------
// One thread per 4x4x4 leaf block ORs the bits of every object into its own block (no atomics); the objects
// are staged through shared memory in chunks.  The per-block arithmetic is env_prims.h (checked against the
// oracle on the host).  add_point() of the centres / end points / plain points is a second, tiny kernel.
constexpr int EP_CHUNK = 64;

__global__ void __launch_bounds__(128)
env_add_primitives_kernel(uint64_t *__restrict__ blocks, const GridDev gd, const double *__restrict__ prims,
                          int64_t n_prims) {
  __shared__ double sh[EP_CHUNK * EP_PRIM_DOUBLES];
  const uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = key < (uint32_t)gd.Nb * gd.Nb * gd.Nb;
  const int bx = (int)compact3(key >> 2), by = (int)compact3(key >> 1), bz = (int)compact3(key);
  uint64_t acc = 0ull;
  for (int64_t base = 0; base < n_prims; base += EP_CHUNK) {
    const int m = (int)((n_prims - base < EP_CHUNK) ? (n_prims - base) : EP_CHUNK);
    __syncthreads();
    for (int i = threadIdx.x; i < m * EP_PRIM_DOUBLES; i += blockDim.x) sh[i] = prims[base * EP_PRIM_DOUBLES + i];
    __syncthreads();
    if (live)
      for (int i = 0; i < m; i++) acc |= ep_block_bits(gd.lo, gd.d, bx, by, bz, sh + i * EP_PRIM_DOUBLES);
  }
  if (live && acc) blocks[key] |= acc;
}

__global__ void env_add_points_kernel(uint64_t *__restrict__ blocks, const GridDev gd,
                                      const double *__restrict__ pts, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  int c[3];
  if (!ep_point_cell(gd.lo, gd.hi, gd.d, gd.Ng, p, c)) return;
  const uint32_t key = (spread3((uint32_t)(c[0] >> 2)) << 2) | (spread3((uint32_t)(c[1] >> 2)) << 1) |
                       spread3((uint32_t)(c[2] >> 2));
  atomicOr((unsigned long long *)&blocks[key], 1ull << ((c[0] & 3) * 16 + (c[1] & 3) * 4 + (c[2] & 3)));
}

}  // namespace

extern "C" {

int irt_env_add_primitives(irt_ctx *ctx, irt_env *env, const double *points, int64_t n_points,
                           const double *spheres, int64_t n_spheres, const double *capsules,
                           int64_t n_capsules, int clear_first) {
  if (!ctx || !env || n_points < 0 || n_spheres < 0 || n_capsules < 0 || (n_points && !points) ||
      (n_spheres && !spheres) || (n_capsules && !capsules))
    return IRT_ERR_INVALID_ARGUMENT;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (clear_first) IRT_CUDA(ctx, cudaMemsetAsync(env->d_blocks, 0, (size_t)env->n_blocks_total * 8, st));
  // host staging: objects as 8 doubles each; add_point() inputs = plain points, sphere centres, capsule ends
  const int64_t n_prims = n_spheres + n_capsules, n_pts = n_points + n_spheres + 2 * n_capsules;
  if (n_prims + n_pts > 0) {
    std::vector<double> h((size_t)n_prims * EP_PRIM_DOUBLES + (size_t)n_pts * 3);
    double *hp = h.data(), *hq = h.data() + (size_t)n_prims * EP_PRIM_DOUBLES;
    for (int64_t i = 0; i < n_points; i++, hq += 3) { hq[0] = points[3 * i]; hq[1] = points[3 * i + 1]; hq[2] = points[3 * i + 2]; }
    for (int64_t i = 0; i < n_spheres; i++, hp += EP_PRIM_DOUBLES, hq += 3) {
      const double *s = spheres + 4 * i;
      hp[0] = s[0]; hp[1] = s[1]; hp[2] = s[2]; hp[3] = hp[4] = hp[5] = 0.0; hp[6] = s[3]; hp[7] = 0.0;
      hq[0] = s[0]; hq[1] = s[1]; hq[2] = s[2];
    }
    for (int64_t i = 0; i < n_capsules; i++, hp += EP_PRIM_DOUBLES, hq += 6) {
      const double *c = capsules + 7 * i;
      for (int k = 0; k < 6; k++) hp[k] = hq[k] = c[k];
      hp[6] = c[6]; hp[7] = 1.0;
    }
    double *d = (double *)ctx_scratch(ctx, h.size() * sizeof(double));
    if (!d) return irt_fail(ctx, IRT_ERR_CUDA, "scratch alloc failed");
    IRT_CUDA(ctx, cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    IRT_CUDA(ctx, cudaStreamSynchronize(st));  // h is pageable and goes out of scope
    if (n_prims > 0) {
      const int T = 128;
      env_add_primitives_kernel<<<(unsigned)((env->n_blocks_total + T - 1) / T), T, 0, st>>>(env->d_blocks, env->gd,
                                                                                            d, n_prims);
      IRT_LAUNCHED(ctx);
    }
    if (n_pts > 0) {
      const int T = 256;
      env_add_points_kernel<<<(unsigned)((n_pts + T - 1) / T), T, 0, st>>>(
          env->d_blocks, env->gd, d + (size_t)n_prims * EP_PRIM_DOUBLES, n_pts);
      IRT_LAUNCHED(ctx);
    }
    IRT_CUDA(ctx, cudaGetLastError());
  }
  int rc = env_rebuild_occ(ctx, env, st);
  if (rc) return rc;
  IRT_CUDA(ctx, cudaStreamSynchronize(st));
  return IRT_OK;
}

int irt_env_dilate(irt_ctx *ctx, irt_env *env, int num, int use_diagonal) {
  if (!ctx || !env) return IRT_ERR_INVALID_ARGUMENT;
  return env_morph(ctx, env, false, use_diagonal ? OFF_27_DILATE_GATHER : OFF_6, num);
}

int irt_env_dilate_sphere(irt_ctx *ctx, irt_env *env, double r) {
  if (!ctx || !env) return IRT_ERR_INVALID_ARGUMENT;
  const double dmin = fmin(env->gd.d[0], fmin(env->gd.d[1], env->gd.d[2]));
  return env_morph(ctx, env, false, OFF_6, (int)round(r / dmin));
}

int irt_env_remove_interior(irt_ctx *ctx, irt_env *env, int keep_diagonal) {
  if (!ctx || !env) return IRT_ERR_INVALID_ARGUMENT;
  return env_morph(ctx, env, true, keep_diagonal ? OFF_27 : OFF_6, 1);
}

int irt_env_download(irt_ctx *ctx, const irt_env *env, uint64_t *blocks) {
  if (!ctx || !env || !blocks) return IRT_ERR_INVALID_ARGUMENT;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  IRT_CUDA(ctx, cudaMemcpyAsync(blocks, env->d_blocks, (size_t)env->n_blocks_total * 8,
                                cudaMemcpyDeviceToHost, ctx->stream));
  IRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return IRT_OK;
}

}  // extern "C"
