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
// K3: one CTA per tile of 256 consecutive sets = 8 uint32 verdict words.  The tile's leaves are
// one contiguous CSR range that the CTA streams with 128-bit loads; a leaf whose `bits & env[key]`
// (behind the occupancy bitmap) is non-zero sets one bit in a shared-memory hit bitmap, and each
// thread then tests the bits of the one set it owns; the word is one __ballot_sync per warp: no
// global atomics, no memset, no search, no per-set divergence (details at the kernel).
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

constexpr int K3_THREADS = 256;

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

// K3 `voxel_and_popc`
//
// One CTA per tile of K3_THREADS consecutive sets (= 8 verdict words).  The tile's leaves are one
// contiguous CSR range that all 256 threads stream as 16-byte aligned quads (uint4 of keys +
// 2 x ulonglong2 of bits per thread and step, two steps in flight, ld.global.cs).  A leaf whose
// `bits & env[key]` is non-zero sets ONE BIT at its tile-relative position in a shared-memory hit
// bitmap (atomicOr, only on a hit); after the stream thread t -- which owns set s0+t and holds its
// leaf range [lo,hi) -- tests the bits of its range (1-2 words for a typical 20-60 leaf set), and a
// __ballot_sync per warp is the verdict word.  No shuffles, no search and no per-set divergence in the
// streaming loop; tiles with more leaves than the bitmap holds are streamed in chunks.  The bitmap is
// double-buffered by tile parity so a chunk costs two __syncthreads.
constexpr int K3_CHUNK_LEAVES = 32768;                      // hit-bitmap capacity per buffer (4 KiB)
constexpr int K3_BM_WORDS = K3_CHUNK_LEAVES / 32 + 2;       // +1 straddle word, +1 pad
constexpr int K3_SMEM_FIXED_WORDS = 3 * K3_BM_WORDS + (3 * K3_BM_WORDS & 1) + 2 * 2 * (K3_THREADS + 2);

// predicated 8-byte read-only load: the four environment look-ups of a quad are issued back to back
// (no branch per look-up), so their latencies overlap; lanes whose leaf is in an empty environment
// block issue nothing
__device__ __forceinline__ uint64_t ld_env_if(const uint64_t *p, uint32_t pred) {
  uint64_t v;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tmov.u64 %0, 0;\n\t@p ld.global.nc.u64 %0, [%2];\n\t}"
      : "=l"(v) : "r"(pred), "l"(p));
  return v;
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// streaming 16-byte loads (evict-first); volatile keeps the issue order of the software pipeline
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ ulonglong2 ld_stream(const ulonglong2 *p) {
  ulonglong2 v;
  asm volatile("ld.global.cs.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  return v;
}

template <bool OCC_SMEM, bool STATS>
__global__ void __launch_bounds__(K3_THREADS, 4)
voxel_and_popc_kernel(const uint32_t *__restrict__ keys, const uint64_t *__restrict__ bits,
                      const uint64_t *__restrict__ offsets, const uint64_t *__restrict__ env,
                      const uint32_t *__restrict__ occ, int occ_words,
                      int64_t set_begin, int64_t set_end, uint32_t *__restrict__ verdict,
                      unsigned long long *__restrict__ stats) {
  // [0, 3*K3_BM_WORDS): three hit bitmaps (rotating, so a chunk costs ONE __syncthreads: the buffer
  // cleared during chunk k was last read two barriers ago); then two tiles of set offsets
  // (cp.async prefetch of the next tile's CSR offsets while this tile streams); then the
  // occupancy bitmap (OCC_SMEM)
  extern __shared__ __align__(16) uint32_t s_mem[];
  constexpr int OFF_OFF = 3 * K3_BM_WORDS + (3 * K3_BM_WORDS & 1);          // 8-byte aligned
  constexpr int OCC_OFF = OFF_OFF + 2 * 2 * (K3_THREADS + 2);
  uint64_t *s_off = reinterpret_cast<uint64_t *>(s_mem + OFF_OFF);          // [2][K3_THREADS + 2]
  const int tid = threadIdx.x;
  const int64_t n_sets = set_end - set_begin;
  const int64_t ntiles = (n_sets + K3_THREADS - 1) / K3_THREADS;
  const int64_t nwords_out = (n_sets + 31) >> 5;
  unsigned long long vox = 0, nhit = 0;

  // asynchronous copy of the offsets of tile t into s_off[buf]: entry i = offsets[s0 + min(i, n_valid)]
  auto prefetch_offsets = [&](int64_t t, int buf) {
    const int64_t s0 = set_begin + t * K3_THREADS;
    const int64_t n_valid = min((int64_t)K3_THREADS, set_end - s0);
    uint64_t *dst = s_off + buf * (K3_THREADS + 2);
    cp_async8(dst + tid, offsets + s0 + min((int64_t)tid, n_valid));
    if (tid == 0) cp_async8(dst + K3_THREADS, offsets + s0 + n_valid);
  };

  if ((int64_t)blockIdx.x < ntiles) prefetch_offsets(blockIdx.x, 0);
  if (OCC_SMEM) {
    for (int i = tid; i < occ_words; i += K3_THREADS) s_mem[OCC_OFF + i] = occ[i];
  }
  for (int i = tid; i < K3_BM_WORDS; i += K3_THREADS) s_mem[i] = 0;
  cp_async_wait_all();
  __syncthreads();
  int hb = 0, mb = 0;   // current hit bitmap / offsets buffer

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint64_t *my_off = s_off + mb * (K3_THREADS + 2);
    const uint64_t t_lo = my_off[0], t_hi = my_off[K3_THREADS];
    const uint64_t lo = my_off[tid], hi = my_off[tid + 1];
    if (tile + gridDim.x < ntiles) prefetch_offsets(tile + gridDim.x, mb ^ 1);
    mb ^= 1;
    bool own = false;
    if (t_lo == t_hi) {   // a tile of empty sets: no chunk below, but the prefetch still needs its barrier
      cp_async_wait_all();
      __syncthreads();
    }

    for (uint64_t c_lo = t_lo; c_lo < t_hi; c_lo += K3_CHUNK_LEAVES) {   // block-uniform
      const uint64_t c_hi = min(c_lo + (uint64_t)K3_CHUNK_LEAVES, t_hi);
      const uint32_t c_n = (uint32_t)(c_hi - c_lo);
      const int hit_off = hb * K3_BM_WORDS;
      hb = (hb == 2) ? 0 : hb + 1;
      const int next_hit_off = hb * K3_BM_WORDS;

      const int64_t q_begin = (int64_t)(c_lo >> 2);
      const int nq = (int)((int64_t)((c_hi + 3) >> 2) - q_begin);          // <= 8193 quads
      const int32_t head = (int32_t)((q_begin << 2) - (int64_t)c_lo);      // in (-4, 0]

      auto process = [&](int32_t r0, const uint4 &kk, const ulonglong2 &b01, const ulonglong2 &b23) {
        const uint32_t k[4] = {kk.x, kk.y, kk.z, kk.w};
        const uint64_t b[4] = {b01.x, b01.y, b23.x, b23.y};
        uint32_t o[4];
        uint64_t ev[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {   // occupied environment leaf?
          const uint32_t w = OCC_SMEM ? s_mem[OCC_OFF + (k[e] >> 5)] : __ldg(occ + (k[e] >> 5));
          o[e] = (w >> (k[e] & 31)) & 1u;
        }
#pragma unroll
        for (int e = 0; e < 4; e++) ev[e] = ld_env_if(env + k[e], o[e]);
        uint32_t hitmask = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const uint64_t x = b[e] & ev[e];
          if (x != 0ull) hitmask |= 1u << e;
          if (STATS) {
            if (x != 0ull && (uint32_t)(r0 + e) < c_n) { vox += (unsigned long long)__popcll(x); nhit++; }
          }
        }
        if (hitmask) {
          // clip leaves of the neighbouring chunks / tiles (only the first and last quad can have any)
          if (r0 < 0) { hitmask >>= -r0; r0 = 0; }
          if ((uint32_t)r0 + 4u > c_n) hitmask &= (c_n > (uint32_t)r0) ? (0xFu >> ((uint32_t)r0 + 4u - c_n)) : 0u;
          const uint64_t m = (uint64_t)hitmask << (r0 & 31);
          if ((uint32_t)m) atomicOr(&s_mem[hit_off + (r0 >> 5)], (uint32_t)m);
          if ((uint32_t)(m >> 32)) atomicOr(&s_mem[hit_off + (r0 >> 5) + 1], (uint32_t)(m >> 32));
        }
      };

      // flat walk over the chunk's leaves: software-pipelined ping-pong, the loads of a thread's next
      // quad are issued before its current quad is processed (two quads per thread in flight)
      const uint4 *kp = reinterpret_cast<const uint4 *>(keys) + q_begin + tid;
      const ulonglong2 *bp = reinterpret_cast<const ulonglong2 *>(bits) + 2 * (q_begin + tid);
      int32_t ra = head + 4 * tid;
      uint4 ka = make_uint4(0, 0, 0, 0), kb = ka;
      ulonglong2 a01 = make_ulonglong2(0, 0), a23 = a01, b01 = a01, b23 = a01;
      if (tid < nq) { ka = ld_stream(kp); a01 = ld_stream(bp); a23 = ld_stream(bp + 1); }
      for (int i = tid; i < nq; i += 2 * K3_THREADS) {
        const bool vb = i + K3_THREADS < nq, vc = i + 2 * K3_THREADS < nq;
        if (vb) {
          kb = ld_stream(kp + K3_THREADS);
          b01 = ld_stream(bp + 2 * K3_THREADS);
          b23 = ld_stream(bp + 2 * K3_THREADS + 1);
        }
        process(ra, ka, a01, a23);
        if (vc) {
          ka = ld_stream(kp + 2 * K3_THREADS);
          a01 = ld_stream(bp + 4 * K3_THREADS);
          a23 = ld_stream(bp + 4 * K3_THREADS + 1);
        }
        if (vb) process(ra + 4 * K3_THREADS, kb, b01, b23);
        kp += 2 * K3_THREADS;
        bp += 4 * K3_THREADS;
        ra += 8 * K3_THREADS;
      }
      for (int i = tid; i < K3_BM_WORDS; i += K3_THREADS) s_mem[next_hit_off + i] = 0;
      cp_async_wait_all();
      __syncthreads();

      // does any hit bit fall into this thread's set range, clipped to the chunk?
      const uint64_t a64 = max(lo, c_lo), b64 = min(hi, c_hi);
      if (a64 < b64 && !own) {
        const uint32_t a = (uint32_t)(a64 - c_lo), b = (uint32_t)(b64 - c_lo);   // [a, b), b > a
        const uint32_t w0 = a >> 5, w1 = (b - 1) >> 5;
        for (uint32_t w = w0; w <= w1; w++) {
          uint32_t m = s_mem[hit_off + w];
          if (w == w0) m &= 0xffffffffu << (a & 31);
          if (w == w1) m &= 0xffffffffu >> (31 - ((b - 1) & 31));
          if (m) { own = true; break; }
        }
      }
    }
    const unsigned word = __ballot_sync(0xffffffffu, own);
    const int64_t wi = tile * (K3_THREADS / 32) + (tid >> 5);
    if ((tid & 31) == 0 && wi < nwords_out) verdict[wi] = word;
  }
  if (STATS) {
    for (int o = 16; o > 0; o >>= 1) {
      vox += __shfl_down_sync(0xffffffffu, vox, o);
      nhit += __shfl_down_sync(0xffffffffu, nhit, o);
    }
    if ((tid & 31) == 0 && (vox | nhit)) {
      atomicAdd(&stats[0], vox);
      atomicAdd(&stats[1], nhit);
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

// content-preserving growth of the keys/bits arrays (amortised x1.5)
int setstore_grow_blocks(irt_ctx *ctx, irt_setstore *s, int64_t need_blocks, int64_t keep_blocks,
                         cudaStream_t st) {
  if ((size_t)need_blocks + 4 <= s->cap_blocks) return IRT_OK;
  size_t cap = s->cap_blocks + s->cap_blocks / 2;
  if (cap < (size_t)need_blocks + 4) cap = (size_t)need_blocks + 4;
  uint32_t *nk = nullptr;
  uint64_t *nb = nullptr;
  IRT_CUDA(ctx, cudaMalloc(&nk, cap * 4));
  IRT_CUDA(ctx, cudaMalloc(&nb, cap * 8));
  if (keep_blocks > 0) {
    IRT_CUDA(ctx, cudaMemcpyAsync(nk, s->d_keys, (size_t)keep_blocks * 4, cudaMemcpyDeviceToDevice, st));
    IRT_CUDA(ctx, cudaMemcpyAsync(nb, s->d_bits, (size_t)keep_blocks * 8, cudaMemcpyDeviceToDevice, st));
    IRT_CUDA(ctx, cudaStreamSynchronize(st));
  }
  if (s->d_keys) cudaFree(s->d_keys);
  if (s->d_bits) cudaFree(s->d_bits);
  s->d_keys = nk;
  s->d_bits = nb;
  s->cap_blocks = cap;
  return IRT_OK;
}

int setstore_finalize(irt_ctx *ctx, irt_setstore *s, int64_t n_sets, int64_t n_blocks,
                      cudaStream_t st) {
  // K3 reads whole 16-byte quads: the (clipped) leaves past the end of the store must hold valid keys
  if (s->d_keys) IRT_CUDA(ctx, cudaMemsetAsync(s->d_keys + n_blocks, 0, 4 * sizeof(uint32_t), st));
  s->n_sets = n_sets;
  s->n_blocks = n_blocks;
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

}  // extern "C"
int env_rebuild_occ(irt_ctx *ctx, irt_env *env, cudaStream_t st) {
  const int64_t nwords = (env->n_blocks_total + 31) / 32;
  const int T = 256;
  env_build_occ_kernel<<<(unsigned)((nwords + T - 1) / T), T, 0, st>>>(env->d_blocks,
                                                                      env->n_blocks_total, env->d_occ);
  IRT_LAUNCHED(ctx);
  IRT_CUDA(ctx, cudaGetLastError());
  return IRT_OK;
}
extern "C" {

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
  if (store->n_blocks == 0) {
    IRT_CUDA(ctx, cudaMemsetAsync(d_verdict, 0, (size_t)((n + 31) / 32) * 4, st));
    return IRT_OK;
  }
  const int occ_words = (int)((env->n_blocks_total + 31) / 32);
  const bool occ_smem = (size_t)occ_words * 4 <= 64 * 1024;
  int64_t blocks = (n + K3_THREADS - 1) / K3_THREADS;     // one CTA per tile of 256 sets ...
  const int64_t max_blocks = (int64_t)ctx->sm_count * 8;  // ... persistent over 8 CTAs per SM
  if (blocks > max_blocks) blocks = max_blocks;
  const size_t smem = (size_t)K3_SMEM_FIXED_WORDS * 4 + (occ_smem ? (size_t)occ_words * 4 : 0);
#define K3_LAUNCH(OS, ST)                                                                          \
  do {                                                                                             \
    auto kfn = voxel_and_popc_kernel<OS, ST>;                                                      \
    if (smem > 48 * 1024)                                                                          \
      IRT_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kfn<<<(unsigned)blocks, K3_THREADS, smem, st>>>(store->d_keys, store->d_bits, store->d_offsets, \
                                                    env->d_blocks, env->d_occ, occ_words, begin,   \
                                                    end, d_verdict, d_stats);                      \
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
