// voxel_check.cu -- environment grid, cached set store and K3 `voxel_and_popc`.
//
// Replaces VoxelOctree::collides(const VoxelOctree&) (collision/VoxelOctree.cpp:973-978 ->
// TreeNode<N>::collides, collision/detail/TreeNode.hxx:165-174, leaf test `_tree & other._tree`
// :268) for every cached vertex/edge voxel set against the environment.
//
// Data layout in HBM (B200-first, not the reference's pointer octree):
//   environment  dense uint64[Nb^3] leaf blocks indexed by Morton key (x-major octant order =
//                the reference's child order, TreeNode.h:66-68), so 8 consecutive entries are one
//                512-bit 2x2x2 super-block (one 64-byte line); plus a 1-bit-per-leaf occupancy
//                bitmap (4 KiB at 128^3) that the kernel stages in shared memory: the bitmap plays
//                the role of the octree's "null child" early-out.
//   set store    CSR, structure-of-arrays: keys uint32[nb], bits uint64[nb], offsets uint64[n+1];
//                12 bytes per occupied leaf block = the algorithmic traffic of SURVEY 8(d).
// K3 streams keys/bits with 128-bit loads, fully coalesced, flat over all leaves of the range
// (no per-set divergence); a leaf that intersects the environment locates its set through a
// per-128-leaf "first set" table and ORs one bit into the verdict bitmask.
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

constexpr int K3_THREADS = 256;
constexpr int K3_CHUNK = 128;  // leaves per entry of the chunk -> first-set table

__global__ void env_build_occ_kernel(const uint64_t *__restrict__ blocks, int64_t nblk,
                                     uint32_t *__restrict__ occ) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nwords = (nblk + 31) / 32;
  if (w >= nwords) return;
  uint32_t m = 0;
  for (int b = 0; b < 32; b++) {
    const int64_t i = w * 32 + b;
    if (i < nblk && blocks[i] != 0) m |= 1u << b;
  }
  occ[w] = m;
}

__global__ void env_scatter_sparse_kernel(const uint8_t *__restrict__ bxyz,
                                          const uint64_t *__restrict__ bits, int64_t n, int levels,
                                          uint64_t *__restrict__ blocks) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t bx = bxyz[3 * i], by = bxyz[3 * i + 1], bz = bxyz[3 * i + 2];
  uint32_t key = 0;
  for (int l = 0; l < levels; l++)
    key |= (((bx >> l) & 1u) << (3 * l + 2)) | (((by >> l) & 1u) << (3 * l + 1)) | (((bz >> l) & 1u) << (3 * l));
  blocks[key] = bits[i];
}

__global__ void count_nonzero_kernel(const uint64_t *__restrict__ blocks, int64_t n,
                                     unsigned long long *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool nz = (i < n) && blocks[i] != 0;
  const unsigned m = __ballot_sync(0xffffffffu, nz);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

// chunk c (leaves [c*128, c*128+128)) -> id of the set containing leaf c*128
__global__ void build_chunk_table_kernel(const uint64_t *__restrict__ offsets, int64_t n_sets,
                                         int64_t n_chunks, uint32_t *__restrict__ chunk_set) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sets) return;
  const uint64_t lo = offsets[s], hi = offsets[s + 1];
  if (hi == lo) return;
  // chunk starts c*128 with lo <= c*128 < hi
  int64_t c0 = (int64_t)((lo + K3_CHUNK - 1) / K3_CHUNK);
  for (int64_t c = c0; c < n_chunks && (uint64_t)c * K3_CHUNK < hi; c++) chunk_set[c] = (uint32_t)s;
}

// K3.  One thread per 4 consecutive leaves per iteration (uint4 keys + 2 x ulonglong2 bits).
template <bool OCC_SMEM, bool STATS>
__global__ void __launch_bounds__(K3_THREADS)
voxel_and_popc_kernel(const uint32_t *__restrict__ keys, const uint64_t *__restrict__ bits,
                      const uint64_t *__restrict__ offsets, const uint32_t *__restrict__ chunk_set,
                      const uint64_t *__restrict__ env, const uint32_t *__restrict__ occ,
                      int occ_words, int64_t leaf_begin, int64_t leaf_end, int64_t set_begin,
                      int64_t set_end, uint32_t *__restrict__ verdict,
                      unsigned long long *__restrict__ stats) {
  extern __shared__ uint32_t s_occ[];
  if (OCC_SMEM) {
    for (int i = threadIdx.x; i < occ_words; i += blockDim.x) s_occ[i] = occ[i];
    __syncthreads();
  }
  const uint32_t *occp = OCC_SMEM ? s_occ : occ;
  const int64_t q_begin = leaf_begin >> 2, q_end = (leaf_end + 3) >> 2;  // quads of 4 leaves
  unsigned long long vox = 0, hits = 0;
  for (int64_t q = q_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < q_end;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j0 = q << 2;
    uint32_t k[4];
    uint64_t b[4];
    if (j0 >= leaf_begin && j0 + 4 <= leaf_end) {
      const uint4 kk = __ldcs(reinterpret_cast<const uint4 *>(keys + j0));
      const ulonglong2 b01 = __ldcs(reinterpret_cast<const ulonglong2 *>(bits + j0));
      const ulonglong2 b23 = __ldcs(reinterpret_cast<const ulonglong2 *>(bits + j0 + 2));
      k[0] = kk.x; k[1] = kk.y; k[2] = kk.z; k[3] = kk.w;
      b[0] = b01.x; b[1] = b01.y; b[2] = b23.x; b[3] = b23.y;
    } else {  // ragged head / tail of the range
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int64_t j = j0 + e;
        const bool in = (j >= leaf_begin && j < leaf_end);
        k[e] = in ? keys[j] : 0u;
        b[e] = in ? bits[j] : 0ull;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; e++) {
      if (b[e] == 0ull) continue;
      if (!((occp[k[e] >> 5] >> (k[e] & 31)) & 1u)) continue;  // empty environment leaf
      const uint64_t x = b[e] & env[k[e]];
      if (x == 0ull) continue;
      if (STATS) { vox += (unsigned long long)__popcll(x); hits++; }
      // locate the set of leaf j: chunk table, then a short forward scan of the offsets
      const int64_t j = j0 + e;
      int64_t s = (int64_t)chunk_set[j / K3_CHUNK];
      if (s < set_begin) s = set_begin;
      while (s + 1 < set_end && offsets[s + 1] <= (uint64_t)j) s++;
      const int64_t rel = s - set_begin;
      const uint32_t bit = 1u << (rel & 31);
      if (!(verdict[rel >> 5] & bit)) atomicOr(&verdict[rel >> 5], bit);
    }
  }
  if (STATS) {
    for (int o = 16; o > 0; o >>= 1) {
      vox += __shfl_down_sync(0xffffffffu, vox, o);
      hits += __shfl_down_sync(0xffffffffu, hits, o);
    }
    if ((threadIdx.x & 31) == 0 && (vox | hits)) {
      atomicAdd(&stats[0], vox);
      atomicAdd(&stats[1], hits);
    }
  }
}

int ensure_store_capacity(irt_ctx *ctx, irt_setstore *s, int64_t n_sets, int64_t n_blocks) {
  if ((size_t)(n_sets + 1) > s->cap_sets) {
    if (s->d_offsets) cudaFree(s->d_offsets);
    s->d_offsets = nullptr;
    s->cap_sets = 0;
    IRT_CUDA(ctx, cudaMalloc(&s->d_offsets, (size_t)(n_sets + 1) * 8));
    s->cap_sets = (size_t)(n_sets + 1);
  }
  if ((size_t)n_blocks + 4 > s->cap_blocks) {
    if (s->d_keys) cudaFree(s->d_keys);
    if (s->d_bits) cudaFree(s->d_bits);
    s->d_keys = nullptr; s->d_bits = nullptr; s->cap_blocks = 0;
    IRT_CUDA(ctx, cudaMalloc(&s->d_keys, ((size_t)n_blocks + 4) * 4));
    IRT_CUDA(ctx, cudaMalloc(&s->d_bits, ((size_t)n_blocks + 4) * 8));
    s->cap_blocks = (size_t)n_blocks + 4;
  }
  return IRT_OK;
}

}  // namespace

// used by voxel_raster.cu after it has filled the store on the device
int setstore_reserve(irt_ctx *ctx, irt_setstore *s, int64_t n_sets, int64_t n_blocks) {
  return ensure_store_capacity(ctx, s, n_sets, n_blocks);
}

int setstore_finalize(irt_ctx *ctx, irt_setstore *s, int64_t n_sets, int64_t n_blocks,
                      cudaStream_t st) {
  s->n_sets = n_sets;
  s->n_blocks = n_blocks;
  irt_setstore &a = *s;
  const int64_t n_chunks = (n_blocks + K3_CHUNK - 1) / K3_CHUNK + 1;
  if ((size_t)n_chunks > a.cap_chunks) {
    if (a.d_chunk_set) cudaFree(a.d_chunk_set);
    a.d_chunk_set = nullptr;
    a.cap_chunks = 0;
    IRT_CUDA(ctx, cudaMalloc(&a.d_chunk_set, (size_t)n_chunks * 4));
    a.cap_chunks = (size_t)n_chunks;
  }
  IRT_CUDA(ctx, cudaMemsetAsync(a.d_chunk_set, 0, (size_t)n_chunks * 4, st));
  if (n_sets > 0) {
    const int T = 256;
    build_chunk_table_kernel<<<(unsigned)((n_sets + T - 1) / T), T, 0, st>>>(s->d_offsets, n_sets,
                                                                           n_chunks, a.d_chunk_set);
    IRT_LAUNCHED(ctx);
    IRT_CUDA(ctx, cudaGetLastError());
  }
  return IRT_OK;
}

extern "C" {

// ---- environment -------------------------------------------------------------------------
int irt_env_create(irt_ctx *ctx, const irt_grid *grid, irt_env **out) {
  if (!ctx || !out) return IRT_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int rc = grid_check(ctx, grid);
  if (rc) return rc;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  irt_env *e = new irt_env();
  e->ctx = ctx;
  e->grid = *grid;
  e->gd = make_grid_dev(*grid);
  const int64_t nblk = (int64_t)e->gd.Nb * e->gd.Nb * e->gd.Nb;
  e->n_blocks_total = nblk;
  const int64_t nwords = (nblk + 31) / 32;
  if (cudaMalloc(&e->d_blocks, (size_t)nblk * 8) != cudaSuccess ||
      cudaMalloc(&e->d_occ, (size_t)nwords * 4) != cudaSuccess) {
    irt_env_destroy(e);
    return irt_fail(ctx, IRT_ERR_CUDA, "environment allocation failed");
  }
  cudaMemsetAsync(e->d_blocks, 0, (size_t)nblk * 8, ctx->stream);
  cudaMemsetAsync(e->d_occ, 0, (size_t)nwords * 4, ctx->stream);
  IRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out = e;
  return IRT_OK;
}

void irt_env_destroy(irt_env *env) {
  if (!env) return;
  cudaSetDevice(env->ctx->device);
  if (env->d_blocks) cudaFree(env->d_blocks);
  if (env->d_occ) cudaFree(env->d_occ);
  delete env;
}

static int env_rebuild_occ(irt_ctx *ctx, irt_env *env, cudaStream_t st) {
  const int64_t nwords = (env->n_blocks_total + 31) / 32;
  const int T = 256;
  env_build_occ_kernel<<<(unsigned)((nwords + T - 1) / T), T, 0, st>>>(env->d_blocks,
                                                                      env->n_blocks_total, env->d_occ);
  IRT_LAUNCHED(ctx);
  IRT_CUDA(ctx, cudaGetLastError());
  return IRT_OK;
}

int irt_env_update_dev(irt_ctx *ctx, irt_env *env, const uint64_t *d_blocks, void *stream) {
  if (!ctx || !env || !d_blocks) return IRT_ERR_INVALID_ARGUMENT;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  IRT_CUDA(ctx, cudaMemcpyAsync(env->d_blocks, d_blocks, (size_t)env->n_blocks_total * 8,
                                cudaMemcpyDeviceToDevice, st));
  return env_rebuild_occ(ctx, env, st);
}

int irt_env_update(irt_ctx *ctx, irt_env *env, const uint64_t *blocks) {
  if (!ctx || !env || !blocks) return IRT_ERR_INVALID_ARGUMENT;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  IRT_CUDA(ctx, cudaMemcpyAsync(env->d_blocks, blocks, (size_t)env->n_blocks_total * 8,
                                cudaMemcpyHostToDevice, ctx->stream));
  int rc = env_rebuild_occ(ctx, env, ctx->stream);
  if (rc) return rc;
  IRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return IRT_OK;
}

int irt_env_update_sparse(irt_ctx *ctx, irt_env *env, const uint8_t *bxyz, const uint64_t *bits,
                          int64_t nblocks) {
  if (!ctx || !env || nblocks < 0 || (nblocks > 0 && (!bxyz || !bits)))
    return IRT_ERR_INVALID_ARGUMENT;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  for (int64_t i = 0; i < nblocks; i++)
    if (bxyz[3 * i] >= env->gd.Nb || bxyz[3 * i + 1] >= env->gd.Nb || bxyz[3 * i + 2] >= env->gd.Nb)
      return irt_fail(ctx, IRT_ERR_OUT_OF_RANGE, "block index outside the grid");
  IRT_CUDA(ctx, cudaMemsetAsync(env->d_blocks, 0, (size_t)env->n_blocks_total * 8, st));
  if (nblocks > 0) {
    char *scr = (char *)ctx_scratch(ctx, (size_t)nblocks * 11 + 64);
    if (!scr) return irt_fail(ctx, IRT_ERR_CUDA, "scratch alloc failed");
    uint64_t *d_bits = (uint64_t *)scr;
    uint8_t *d_xyz = (uint8_t *)(scr + (size_t)nblocks * 8);
    IRT_CUDA(ctx, cudaMemcpyAsync(d_bits, bits, (size_t)nblocks * 8, cudaMemcpyHostToDevice, st));
    IRT_CUDA(ctx, cudaMemcpyAsync(d_xyz, bxyz, (size_t)nblocks * 3, cudaMemcpyHostToDevice, st));
    const int T = 256;
    env_scatter_sparse_kernel<<<(unsigned)((nblocks + T - 1) / T), T, 0, st>>>(
        d_xyz, d_bits, nblocks, env->gd.levels, env->d_blocks);
    IRT_LAUNCHED(ctx);
  }
  int rc = env_rebuild_occ(ctx, env, st);
  if (rc) return rc;
  IRT_CUDA(ctx, cudaStreamSynchronize(st));
  return IRT_OK;
}

int64_t irt_env_nblocks(irt_ctx *ctx, const irt_env *env) {
  if (!ctx || !env) return -1;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return -1;
  unsigned long long *d = (unsigned long long *)ctx_scratch(ctx, 64);
  if (!d) return -1;
  cudaMemsetAsync(d, 0, 8, ctx->stream);
  const int T = 256;
  count_nonzero_kernel<<<(unsigned)((env->n_blocks_total + T - 1) / T), T, 0, ctx->stream>>>(
      env->d_blocks, env->n_blocks_total, d);
  IRT_LAUNCHED(ctx);
  unsigned long long h = 0;
  if (cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
  return (int64_t)h;
}

// ---- set store ---------------------------------------------------------------------------
int irt_setstore_create(irt_ctx *ctx, const irt_grid *grid, irt_setstore **out) {
  if (!ctx || !out) return IRT_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int rc = grid_check(ctx, grid);
  if (rc) return rc;
  irt_setstore *s = new irt_setstore();
  s->ctx = ctx;
  s->grid = *grid;
  s->gd = make_grid_dev(*grid);
  *out = s;
  return IRT_OK;
}

void irt_setstore_destroy(irt_setstore *s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  if (s->d_chunk_set) cudaFree(s->d_chunk_set);
  if (s->d_offsets) cudaFree(s->d_offsets);
  if (s->d_keys) cudaFree(s->d_keys);
  if (s->d_bits) cudaFree(s->d_bits);
  delete s;
}

int64_t irt_setstore_num_sets(const irt_setstore *s) { return s ? s->n_sets : -1; }
int64_t irt_setstore_num_blocks(const irt_setstore *s) { return s ? s->n_blocks : -1; }

int irt_setstore_import(irt_ctx *ctx, irt_setstore *s, int64_t n_sets, const uint64_t *offsets,
                        const uint32_t *keys, const uint64_t *bits) {
  if (!ctx || !s || n_sets < 0 || !offsets) return IRT_ERR_INVALID_ARGUMENT;
  if (offsets[0] != 0) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "offsets[0] must be 0");
  const int64_t nb = (int64_t)offsets[n_sets];
  if (nb > 0 && (!keys || !bits)) return IRT_ERR_INVALID_ARGUMENT;
  const uint32_t nkeys = (uint32_t)((int64_t)s->gd.Nb * s->gd.Nb * s->gd.Nb);
  for (int64_t i = 0; i < n_sets; i++)
    if (offsets[i + 1] < offsets[i]) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "offsets not monotone");
  for (int64_t i = 0; i < nb; i++)
    if (keys[i] >= nkeys) return irt_fail(ctx, IRT_ERR_OUT_OF_RANGE, "block key outside the grid");
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_store_capacity(ctx, s, n_sets, nb);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  IRT_CUDA(ctx, cudaMemcpyAsync(s->d_offsets, offsets, (size_t)(n_sets + 1) * 8, cudaMemcpyHostToDevice, st));
  if (nb > 0) {
    IRT_CUDA(ctx, cudaMemcpyAsync(s->d_keys, keys, (size_t)nb * 4, cudaMemcpyHostToDevice, st));
    IRT_CUDA(ctx, cudaMemcpyAsync(s->d_bits, bits, (size_t)nb * 8, cudaMemcpyHostToDevice, st));
  }
  rc = setstore_finalize(ctx, s, n_sets, nb, st);
  if (rc) return rc;
  IRT_CUDA(ctx, cudaStreamSynchronize(st));
  return IRT_OK;
}

int irt_setstore_export(irt_ctx *ctx, const irt_setstore *s, uint64_t *offsets, uint32_t *keys,
                        uint64_t *bits) {
  if (!ctx || !s) return IRT_ERR_INVALID_ARGUMENT;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (offsets) {
    if (s->d_offsets)
      IRT_CUDA(ctx, cudaMemcpyAsync(offsets, s->d_offsets, (size_t)(s->n_sets + 1) * 8, cudaMemcpyDeviceToHost, st));
    else
      offsets[0] = 0;
  }
  if (keys && s->n_blocks > 0)
    IRT_CUDA(ctx, cudaMemcpyAsync(keys, s->d_keys, (size_t)s->n_blocks * 4, cudaMemcpyDeviceToHost, st));
  if (bits && s->n_blocks > 0)
    IRT_CUDA(ctx, cudaMemcpyAsync(bits, s->d_bits, (size_t)s->n_blocks * 8, cudaMemcpyDeviceToHost, st));
  IRT_CUDA(ctx, cudaStreamSynchronize(st));
  return IRT_OK;
}

int irt_setstore_device_ptrs(const irt_setstore *s, const uint64_t **d_offsets,
                             const uint32_t **d_keys, const uint64_t **d_bits) {
  if (!s) return IRT_ERR_INVALID_ARGUMENT;
  if (d_offsets) *d_offsets = s->d_offsets;
  if (d_keys) *d_keys = s->d_keys;
  if (d_bits) *d_bits = s->d_bits;
  return IRT_OK;
}

// ---- K3 ----------------------------------------------------------------------------------
static int check_sets_impl(irt_ctx *ctx, const irt_setstore *store, const irt_env *env,
                           int64_t begin, int64_t end, uint32_t *d_verdict,
                           unsigned long long *d_stats, cudaStream_t st) {
  if (!ctx || !store || !env || !d_verdict) return IRT_ERR_INVALID_ARGUMENT;
  if (store->grid.Ng != env->grid.Ng)  // check_dims -- collision/VoxelOctree.cpp:46-53
    return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "voxel dimension mismatch (%d != %d)",
                    store->grid.Ng, env->grid.Ng);
  if (begin < 0 || end < begin || end > store->n_sets)
    return irt_fail(ctx, IRT_ERR_OUT_OF_RANGE, "set range [%lld,%lld) outside [0,%lld)",
                    (long long)begin, (long long)end, (long long)store->n_sets);
  const int64_t n = end - begin;
  if (n == 0) return IRT_OK;
  IRT_CUDA(ctx, cudaMemsetAsync(d_verdict, 0, (size_t)((n + 31) / 32) * 4, st));
  // leaf range of the set range (known on the host for a whole-store sweep, else two 8-byte reads)
  uint64_t lr[2] = {0, (uint64_t)store->n_blocks};
  if (begin != 0 || end != store->n_sets) {
    IRT_CUDA(ctx, cudaMemcpyAsync(&lr[0], store->d_offsets + begin, 8, cudaMemcpyDeviceToHost, st));
    IRT_CUDA(ctx, cudaMemcpyAsync(&lr[1], store->d_offsets + end, 8, cudaMemcpyDeviceToHost, st));
    IRT_CUDA(ctx, cudaStreamSynchronize(st));
  }
  const int64_t leaf_begin = (int64_t)lr[0], leaf_end = (int64_t)lr[1];
  if (leaf_end == leaf_begin) return IRT_OK;
  const irt_setstore &a = *store;
  const int occ_words = (int)((env->n_blocks_total + 31) / 32);
  const bool occ_smem = (size_t)occ_words * 4 <= 64 * 1024;
  const int64_t quads = ((leaf_end + 3) >> 2) - (leaf_begin >> 2);
  int64_t blocks = (quads + K3_THREADS - 1) / K3_THREADS;
  const int64_t max_blocks = (int64_t)ctx->sm_count * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > max_blocks) blocks = max_blocks;
  const size_t smem = occ_smem ? (size_t)occ_words * 4 : 0;
#define K3_LAUNCH(OS, ST)                                                                          \
  do {                                                                                             \
    auto kfn = voxel_and_popc_kernel<OS, ST>;                                                      \
    if (smem > 48 * 1024)                                                                          \
      IRT_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kfn<<<(unsigned)blocks, K3_THREADS, smem, st>>>(                                               \
        store->d_keys, store->d_bits, store->d_offsets, a.d_chunk_set, env->d_blocks, env->d_occ,  \
        occ_words, leaf_begin, leaf_end, begin, end, d_verdict, d_stats);                          \
  } while (0)
  if (occ_smem) {
    if (d_stats) K3_LAUNCH(true, true); else K3_LAUNCH(true, false);
  } else {
    if (d_stats) K3_LAUNCH(false, true); else K3_LAUNCH(false, false);
  }
#undef K3_LAUNCH
  IRT_LAUNCHED(ctx);
  IRT_CUDA(ctx, cudaGetLastError());
  return IRT_OK;
}

int irt_check_sets_dev(irt_ctx *ctx, const irt_setstore *store, const irt_env *env,
                       int64_t begin, int64_t end, uint32_t *d_verdict_words, void *stream) {
  if (!ctx) return IRT_ERR_INVALID_ARGUMENT;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  return check_sets_impl(ctx, store, env, begin, end, d_verdict_words, nullptr, st);
}

int irt_check_sets(irt_ctx *ctx, const irt_setstore *store, const irt_env *env, int64_t begin,
                   int64_t end, uint32_t *verdict_words) {
  if (!ctx || !verdict_words) return IRT_ERR_INVALID_ARGUMENT;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = end - begin;
  if (n <= 0) return (n == 0) ? IRT_OK : IRT_ERR_OUT_OF_RANGE;
  const size_t words = (size_t)((n + 31) / 32);
  uint32_t *d_v = (uint32_t *)ctx_scratch(ctx, words * 4 + 64);
  if (!d_v) return irt_fail(ctx, IRT_ERR_CUDA, "scratch alloc failed");
  int rc = check_sets_impl(ctx, store, env, begin, end, d_v, nullptr, ctx->stream);
  if (rc) return rc;
  IRT_CUDA(ctx, cudaMemcpyAsync(verdict_words, d_v, words * 4, cudaMemcpyDeviceToHost, ctx->stream));
  IRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return IRT_OK;
}

// colliding-voxel popcount over the range: stats[0] = sum popc(bits & env), stats[1] = #leaves hit
int irt_check_sets_popcount(irt_ctx *ctx, const irt_setstore *store, const irt_env *env,
                            int64_t begin, int64_t end, uint64_t *stats) {
  if (!ctx || !stats) return IRT_ERR_INVALID_ARGUMENT;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = end - begin;
  stats[0] = stats[1] = 0;
  if (n <= 0) return (n == 0) ? IRT_OK : IRT_ERR_OUT_OF_RANGE;
  const size_t words = (size_t)((n + 31) / 32);
  char *scr = (char *)ctx_scratch(ctx, words * 4 + 128);
  if (!scr) return irt_fail(ctx, IRT_ERR_CUDA, "scratch alloc failed");
  unsigned long long *d_stats = (unsigned long long *)scr;
  uint32_t *d_v = (uint32_t *)(scr + 64);
  IRT_CUDA(ctx, cudaMemsetAsync(d_stats, 0, 16, ctx->stream));
  int rc = check_sets_impl(ctx, store, env, begin, end, d_v, d_stats, ctx->stream);
  if (rc) return rc;
  IRT_CUDA(ctx, cudaMemcpyAsync(stats, d_stats, 16, cudaMemcpyDeviceToHost, ctx->stream));
  IRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return IRT_OK;
}

int64_t irt_check_sets_algorithmic_bytes(const irt_setstore *store, int64_t begin, int64_t end) {
  if (!store || begin < 0 || end < begin || end > store->n_sets) return -1;
  if (end == begin) return 0;
  uint64_t lr[2] = {0, 0};
  if (cudaSetDevice(store->ctx->device) != cudaSuccess) return -1;
  if (cudaMemcpy(&lr[0], store->d_offsets + begin, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (cudaMemcpy(&lr[1], store->d_offsets + end, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  const int64_t nb = (int64_t)(lr[1] - lr[0]), n = end - begin;
  const int64_t Nb = store->gd.Nb;
  return 12 * nb + 8 * n + 8 * Nb * Nb * Nb + (n + 7) / 8;
}

}  // extern "C"
