// selfcol.cu -- validity epilogue of K1: collision::collides_self(CapsuleSequence)
// (compiled with -fmad=false: the exact stage repeats the reference's operations one by one, and a contracted
// multiply-add moves a capsule distance by an ulp -- enough to flip a pair that sits exactly at 2r; the FP32
// filter spells its fused operations out)
// (collision/collision.cpp:6-46) with closest_st_segment (collision_primitives.cpp:10-102)
// and the capsule/sphere tests of collision.hxx:65-68,102-108.
//
// One warp per shape.  The O(P^2) capsule-pair loop of the reference is pruned by an exact
// (conservative) two-level test: chunks of 8 consecutive capsules get a bounding sphere; only
// chunk pairs whose spheres come within 2r (+margin) have their 64 capsule pairs examined
// with the reference's arithmetic.  The verdict (any pair collides) is unchanged.
#include "common.cuh"
#include "capsule_pair.h"

namespace {

constexpr int SC_WARPS = 4;

// Stage 1 -- conservative FP32 pair filter, one warp per shape.  Every capsule pair (a, b) that can
// survive the reference's 3r arc-length rule (exact index bound from the longest segment:
// acc[b] - acc[a+1] <= (b - a - 1) * maxlen) is tested with
//     |mid_a - mid_b| <= 2r + len_a/2 + len_b/2        (a lower bound of the capsule distance)
// in single precision with outward margins.  A shape with no such pair cannot self-collide; the
// others (rare: the backbone has to curl back on itself) go to stage 2.  Rows a and R-1-a are
// given to the same lane so that lanes carry equal work.
constexpr int SC_FILTER_WARPS = 8;

// LPS = lanes per shape: 32, or 16 when a shape has so few capsules that half a warp covers its pair rows
// (P = 41: 15 row pairs) -- then a warp filters two shapes at once.  All warp-level operations use the group's
// own lane mask, so the two halves may run different trip counts.
template <int LPS>
__global__ void __launch_bounds__(SC_FILTER_WARPS * 32)
self_collision_filter_kernel(const double *__restrict__ p, const int32_t *__restrict__ npts, int64_t n_host,
                             const int32_t *__restrict__ range, int cap_pts, double r,
                             int32_t *__restrict__ cand, int32_t *__restrict__ n_cand,
                             const int64_t *__restrict__ row_off) {
  extern __shared__ float fsm[];
  constexpr int GPW = 32 / LPS;                                  // groups (shapes) per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / LPS, gl = lane % LPS;
  const unsigned gmask = (LPS == 32) ? 0xffffffffu : (0xffffu << (16 * grp));
  float4 *seg = reinterpret_cast<float4 *>(fsm) + (size_t)(warp * GPW + grp) * cap_pts;  // (mid.x, mid.y, mid.z, half length)
  const float rr = (float)(2.0 * r) * 1.00001f + 1e-6f;
  int64_t lo = 0, n = n_host;
  if (range) { lo = range[0]; n = (int64_t)range[1] - lo; }   // rows [lo, hi) with the bounds in device memory

  for (int64_t w = ((int64_t)blockIdx.x * SC_FILTER_WARPS + warp) * GPW + grp; w < n;
       w += (int64_t)gridDim.x * SC_FILTER_WARPS * GPW) {
    const int64_t shape = lo + w;
    const int N = npts[shape];
    __syncwarp(gmask);
    if (N <= 3) continue;  // collision.cpp:14 and the loop bounds a < N-3
    const double *src = p + (row_off ? row_off[shape] : shape * (int64_t)cap_pts) * 3;   // packed or dense rows
    const int ncap = N - 1;
    float maxhl = 0.0f, tsum = 0.0f, tmax = 0.0f;
    for (int i = gl; i < ncap; i += LPS) {
      const double ax = src[3 * i], ay = src[3 * i + 1], az = src[3 * i + 2];
      const double bx = src[3 * i + 3], by = src[3 * i + 4], bz = src[3 * i + 5];
      const float dx = (float)(bx - ax), dy = (float)(by - ay), dz = (float)(bz - az);
      const float hl = 0.5f * sqrtf(dx * dx + dy * dy + dz * dz) * 1.00001f + 1e-9f;
      seg[i] = make_float4((float)(0.5 * (ax + bx)), (float)(0.5 * (ay + by)), (float)(0.5 * (az + bz)), hl);
      maxhl = fmaxf(maxhl, hl);
      if (i + 1 < ncap) {   // turning angle between capsule i and capsule i + 1 (over-estimated)
        const float ex = (float)(src[3 * i + 6] - bx), ey = (float)(src[3 * i + 7] - by), ez = (float)(src[3 * i + 8] - bz);
        const float cx = dy * ez - dz * ey, cy = dz * ex - dx * ez, cz = dx * ey - dy * ex;
        const float th = atan2f(sqrtf(cx * cx + cy * cy + cz * cz), dx * ex + dy * ey + dz * ez) * 1.001f + 1e-6f;
        tsum += th;
        tmax = fmaxf(tmax, th);
      }
    }
    for (int o = LPS / 2; o > 0; o >>= 1) {
      maxhl = fmaxf(maxhl, __shfl_xor_sync(gmask, maxhl, o));
      tsum += __shfl_xor_sync(gmask, tsum, o);
      tmax = fmaxf(tmax, __shfl_xor_sync(gmask, tmax, o));
    }
    __syncwarp(gmask);
    // A backbone that turns too little cannot touch itself: between any two points x, y on capsules a < b every
    // segment direction lies within phi = turning(a..b) / 2 + (largest single turn) of one of them (the one where
    // the running turn passes half), so |x - y| >= (arc length between them) * cos(phi); pairs the reference does
    // not skip have arc length >= 3r (collision.cpp:37-39), and a collision needs |x - y| <= 2r: impossible while
    // cos(phi) > 2/3.  tsum over-estimates every turning(a..b).
    if (0.5f * tsum + tmax < 0.83f) continue;   // acos(2/3) = 0.8411
    // smallest index gap b - a - 1 that can reach 3r of arc length (maxhl over-estimates len/2)
    const float safe = (float)(3.0 * r) * 0.9999f;
    const int min_gap = (maxhl > 0.0f) ? (int)fminf(1e6f, floorf(safe / (2.0f * maxhl))) : 1000000;
    const int first = 1 + (min_gap < 1 ? 1 : min_gap);  // b >= a + first (and the reference's b >= a + 2)
    const int R = ncap - first;                         // rows a = 0 .. R-1 have at least one b
    bool found = false;
    for (int k = gl; k < (R + 1) / 2; k += LPS) {       // rows k and R-1-k: R + 1 pair tests per lane
      for (int half = 0; half < 2; half++) {
        const int a = half ? (R - 1 - k) : k;
        if (half && a == k) break;
        if (a >= N - 3) continue;
        const float4 sa = seg[a];
        for (int b = a + first; b < ncap; b++) {
          const float4 sb = seg[b];
          const float dx = sa.x - sb.x, dy = sa.y - sb.y, dz = sa.z - sb.z;
          const float reach = rr + sa.w + sb.w;
          if (fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= reach * reach) found = true;
        }
      }
    }
    if (__any_sync(gmask, found) && gl == 0) cand[atomicAdd(n_cand, 1)] = (int32_t)w;
  }
}

// Stage 2 -- exact test, one warp per candidate shape (all shapes when cand == nullptr).
__global__ void __launch_bounds__(SC_WARPS * 32)
self_collision_kernel(const double *__restrict__ p, const int32_t *__restrict__ npts, int64_t n,
                      const int32_t *__restrict__ range, int cap_pts, double r, uint32_t *__restrict__ flags,
                      const int32_t *__restrict__ cand, const int32_t *__restrict__ n_cand,
                      const int64_t *__restrict__ row_off) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per warp: xyz[cap*3], acc[cap], chunk centre xyz + radius [4 * nchunk_max]
  const int nchunk_max = (cap_pts + SC_CHUNK - 1) / SC_CHUNK;
  double *base = sm + (size_t)warp * ((size_t)cap_pts * 4 + (size_t)nchunk_max * 4);
  double *px = base, *acc = base + (size_t)cap_pts * 3, *ch = acc + cap_pts;
  const double dist_to_consider = 3.0 * r;
  const double rr = r + r;

  const int64_t lo = range ? range[0] : 0;
  if (range) n = (int64_t)range[1] - lo;
  const int64_t n_work = cand ? (int64_t)*n_cand : n;
  for (int64_t w = (int64_t)blockIdx.x * SC_WARPS + warp; w < n_work; w += (int64_t)gridDim.x * SC_WARPS) {
    const int64_t shape = lo + (cand ? (int64_t)cand[w] : w);
    const int N = npts[shape];
    __syncwarp();
    if (N <= 2) continue;  // collision.cpp:14
    const double *src = p + (row_off ? row_off[shape] : shape * (int64_t)cap_pts) * 3;   // packed or dense rows
    for (int i = lane; i < N * 3; i += 32) px[i] = src[i];
    __syncwarp();
    // segment lengths (the reference's per-point distance, collision.cpp:22-29); acc[] is only
    // turned into the running sum (same sequential order as the reference) if it is ever needed
    double maxlen = 0.0;
    for (int i = lane; i < N; i += 32) {
      const double len = sc_segment_len(px, i);
      acc[i] = len;
      maxlen = fmax(maxlen, len);
    }
    for (int o = 16; o > 0; o >>= 1) maxlen = fmax(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    // min_gap = smallest b - a - 1 that can survive the reference's arc-length rule (capsule_pair.h)
    const int min_gap = sc_min_gap(maxlen, dist_to_consider);
    __syncwarp();
    // chunk bounding spheres over capsules [c*8, c*8+8) i.e. points [c*8, min(c*8+8, N-1)]
    const int ncap = N - 1;
    const int nchunk = (ncap + SC_CHUNK - 1) / SC_CHUNK;
    for (int c = lane; c < nchunk; c += 32) {
      const int i0 = c * SC_CHUNK, i1 = min(i0 + SC_CHUNK, ncap);  // points i0..i1 inclusive
      const int im = (i0 + i1) >> 1;
      const double cx = px[3 * im], cy = px[3 * im + 1], cz = px[3 * im + 2];
      double rad2 = 0.0;
      for (int i = i0; i <= i1; i++) {
        const double dx = px[3 * i] - cx, dy = px[3 * i + 1] - cy, dz = px[3 * i + 2] - cz;
        rad2 = fmax(rad2, (dx * dx + dy * dy) + dz * dz);
      }
      ch[4 * c] = cx; ch[4 * c + 1] = cy; ch[4 * c + 2] = cz; ch[4 * c + 3] = sqrt(rad2);
    }
    __syncwarp();
    bool hit = false;
    bool have_acc = false;
    const int npairs = nchunk * nchunk;
    for (int base_pair = 0; base_pair < npairs && !hit; base_pair += 32) {
      const int pair = base_pair + lane;
      bool near = false;
      int ca = 0, cb = 0;
      if (pair < npairs) {
        ca = pair / nchunk; cb = pair - ca * nchunk;
        // largest index gap b - a - 1 inside this chunk pair
        const int max_gap = (cb + 1) * SC_CHUNK - 1 - ca * SC_CHUNK - 1;
        if (cb >= ca && max_gap >= min_gap) {
          const double dx = ch[4 * ca] - ch[4 * cb], dy = ch[4 * ca + 1] - ch[4 * cb + 1],
                       dz = ch[4 * ca + 2] - ch[4 * cb + 2];
          const double reach = (ch[4 * ca + 3] + ch[4 * cb + 3] + rr) * (1.0 + 1e-9) + 1e-12;
          near = ((dx * dx + dy * dy) + dz * dz) <= reach * reach;
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, near);
      if (m && !have_acc) {  // rare: some distant chunks come close -> build the exact running sum
        if (lane == 0) {
          double dsum = 0.0;
          for (int i = 0; i < N; i++) { dsum += acc[i]; acc[i] = dsum; }
        }
        have_acc = true;
        __syncwarp();
      }
      while (m && !hit) {
        const int src_lane = __ffs(m) - 1;
        m &= m - 1;
        const int a0 = __shfl_sync(0xffffffffu, ca, src_lane) * SC_CHUNK;
        const int b0 = __shfl_sync(0xffffffffu, cb, src_lane) * SC_CHUNK;
        bool h = false;
        for (int k = lane; k < SC_CHUNK * SC_CHUNK; k += 32) {
          const int a = a0 + (k >> 3), b = b0 + (k & 7);
          // loop bounds of collision.cpp:35-36: a < N-3, a+2 <= b < N-1
          if (a < N - 3 && b >= a + 2 && b < N - 1) {
            if (!(acc[b] - acc[a + 1] < dist_to_consider)) {
              const P3 A0 = {px[3 * a], px[3 * a + 1], px[3 * a + 2]};
              const P3 A1 = {px[3 * a + 3], px[3 * a + 4], px[3 * a + 5]};
              const P3 B0 = {px[3 * b], px[3 * b + 1], px[3 * b + 2]};
              const P3 B1 = {px[3 * b + 3], px[3 * b + 4], px[3 * b + 5]};
              if (capsules_collide(A0, A1, B0, B1, rr)) h = true;
            }
          }
        }
        hit = __any_sync(0xffffffffu, h);
      }
    }
    if (hit && lane == 0) flags[shape] |= IRT_FLAG_SELF_COLLISION;
  }
}

}  // namespace

size_t selfcol_work_bytes(int64_t n) { return 256 + (((size_t)n * 4 + 255) & ~(size_t)255); }

static int selfcol_run(irt_ctx *ctx, double radius, size_t scratch_offset, const double *d_p,
                       const int32_t *d_npts, int64_t n, int cap_pts, uint32_t *d_flags,
                       cudaStream_t st, const int64_t *d_row_off, const int32_t *d_range, void *work);

int self_collision_launch(irt_ctx *ctx, const irt_robot *rb, const double *d_p,
                          const int32_t *d_npts, int64_t n, int cap_pts, uint32_t *d_flags,
                          cudaStream_t st, const int64_t *d_row_off, const int32_t *d_range, void *work) {
  // without private scratch: behind the bucket permutation area of the context scratch
  return selfcol_run(ctx, rb->dev.r, work ? 0 : fk_work_bytes(rb, n), d_p, d_npts, n, cap_pts, d_flags, st, d_row_off,
                     d_range, work);
}

static int selfcol_run(irt_ctx *ctx, double radius, size_t scratch_offset, const double *d_p,
                       const int32_t *d_npts, int64_t n, int cap_pts, uint32_t *d_flags,
                       cudaStream_t st, const int64_t *d_row_off, const int32_t *d_range, void *work) {
  if (n <= 0) return IRT_OK;
  if (n > 0x7fffffffLL) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "batch too large");
  if (cap_pts > IRT_CAP_PTS_MAX) return irt_fail(ctx, IRT_ERR_CAPACITY, "cap_pts=%d too large", cap_pts);
  const int nchunk_max = (cap_pts + SC_CHUNK - 1) / SC_CHUNK;
  size_t smem = (size_t)SC_WARPS * ((size_t)cap_pts * 4 + (size_t)nchunk_max * 4) * sizeof(double);
  if (smem > 200 * 1024) return irt_fail(ctx, IRT_ERR_CAPACITY, "cap_pts=%d too large", cap_pts);
  IRT_CUDA(ctx, cudaFuncSetAttribute(self_collision_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  char *scr = (char *)work;
  if (!scr) {
    scr = (char *)ctx_scratch(ctx, scratch_offset + selfcol_work_bytes(n));
    if (!scr) return irt_fail(ctx, IRT_ERR_CUDA, "scratch alloc failed");
    scr += scratch_offset;
  }
  int32_t *d_ncand = (int32_t *)scr;
  int32_t *d_cand = (int32_t *)(scr + 256);
  IRT_CUDA(ctx, cudaMemsetAsync(d_ncand, 0, 4, st));
  {
    // half a warp per shape while a shape's pair rows fit 16 lanes (up to ~48 points at the usual r / dL)
    const bool half = cap_pts <= 48;
    const int gpw = half ? 2 : 1;
    int64_t fb = (n + SC_FILTER_WARPS * gpw - 1) / (SC_FILTER_WARPS * gpw);
    const int64_t fmax_blocks = (int64_t)ctx->sm_count * 16;
    if (fb > fmax_blocks) fb = fmax_blocks;
    const size_t fsmem = (size_t)SC_FILTER_WARPS * gpw * cap_pts * sizeof(float4);
    auto kf = half ? self_collision_filter_kernel<16> : self_collision_filter_kernel<32>;
    if (fsmem > 48 * 1024)
      IRT_CUDA(ctx, cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
    kf<<<(unsigned)fb, SC_FILTER_WARPS * 32, fsmem, st>>>(d_p, d_npts, n, d_range, cap_pts, radius, d_cand,
                                                          d_ncand, d_row_off);
  }
  IRT_LAUNCHED(ctx);
  // stage 2 runs over however many candidates stage 1 found (count stays on the device)
  int64_t blocks = (n + SC_WARPS - 1) / SC_WARPS;
  const int64_t max_blocks = (int64_t)ctx->sm_count * 4;
  if (blocks > max_blocks) blocks = max_blocks;
  self_collision_kernel<<<(unsigned)blocks, SC_WARPS * 32, smem, st>>>(d_p, d_npts, n, d_range, cap_pts, radius,
                                                                      d_flags, d_cand, d_ncand, d_row_off);
  IRT_LAUNCHED(ctx);
  IRT_CUDA(ctx, cudaGetLastError());
  return IRT_OK;
}

// collision::collides_self(CapsuleSequence{points, r}) (collision/collision.cpp:6-46; TendonRobot::collides_self,
// tendon/TendonRobot.cpp:955-974) for n given backbones: p[n][cap_pts][3], npts[n] (host); collides[i] = 0 / 1
extern "C" int irt_self_collision_shapes(irt_ctx *ctx, const double *p, const int32_t *npts, int cap_pts, int64_t n,
                                         double r, uint8_t *collides) {
  if (!ctx || n < 0 || cap_pts < 1 || (n > 0 && (!p || !npts || !collides))) return IRT_ERR_INVALID_ARGUMENT;
  if (!(r >= 0.0)) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "negative radius");
  for (int64_t i = 0; i < n; i++)
    if (npts[i] < 0 || npts[i] > cap_pts)
      return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "npts[%lld] out of range", (long long)i);
  if (n == 0) return IRT_OK;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t b_p = ((size_t)n * cap_pts * 24 + 255) & ~(size_t)255, b_n = ((size_t)n * 4 + 255) & ~(size_t)255;
  char *io = (char *)ctx_io(ctx, b_p + 2 * b_n);
  if (!io) return irt_fail(ctx, IRT_ERR_CUDA, "device staging allocation failed");
  double *d_p = (double *)io;
  int32_t *d_npts = (int32_t *)(io + b_p);
  uint32_t *d_flags = (uint32_t *)(io + b_p + b_n);
  IRT_CUDA(ctx, cudaMemcpyAsync(d_p, p, (size_t)n * cap_pts * 24, cudaMemcpyHostToDevice, st));
  IRT_CUDA(ctx, cudaMemcpyAsync(d_npts, npts, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  IRT_CUDA(ctx, cudaMemsetAsync(d_flags, 0, (size_t)n * 4, st));
  int rc = selfcol_run(ctx, r, 0, d_p, d_npts, n, cap_pts, d_flags, st, nullptr, nullptr, nullptr);
  if (rc) return rc;
  std::vector<uint32_t> h((size_t)n);
  IRT_CUDA(ctx, cudaMemcpyAsync(h.data(), d_flags, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  IRT_CUDA(ctx, cudaStreamSynchronize(st));
  for (int64_t i = 0; i < n; i++) collides[i] = (h[(size_t)i] & IRT_FLAG_SELF_COLLISION) ? 1 : 0;
  return IRT_OK;
}
