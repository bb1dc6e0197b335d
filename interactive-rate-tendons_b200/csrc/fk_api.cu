// fk_api.cu -- C-ABI entry points of the FK path (host-pointer and device-pointer forms).
#include <cstring>

#include "common.cuh"

namespace {

struct DevBuf {
  void *p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  bool alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1) == cudaSuccess; }
};

// off[0] = 0, off[i + 1] = cnt[0] + ... + cnt[i] for one chunk (m <= ~100k row counts): one CTA walks the array
// in tiles of 1024 (coalesced loads), scans a tile with warp shuffles and carries the running total
__global__ void __launch_bounds__(1024) row_offsets_kernel(const int64_t *__restrict__ cnt, int64_t *__restrict__ off,
                                                           int64_t m) {
  __shared__ int64_t wsum[32];
  __shared__ int64_t carry_s;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) { off[0] = 0; carry_s = 0; }
  __syncthreads();
  for (int64_t base = 0; base < m; base += 1024) {
    const int64_t i = base + t;
    int64_t v = (i < m) ? cnt[i] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {          // inclusive scan inside the warp
      const int64_t o = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += o;
    }
    if (lane == 31) wsum[w] = v;
    __syncthreads();
    if (w == 0) {                               // scan of the 32 warp totals
      int64_t s = wsum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int64_t o = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += o;
      }
      wsum[lane] = s;
    }
    __syncthreads();
    const int64_t carry = carry_s;
    const int64_t incl = carry + v + (w ? wsum[w - 1] : 0);
    if (i < m) off[i + 1] = incl;
    __syncthreads();
    if (t == 1023) carry_s = incl;
    __syncthreads();
  }
}

__global__ void copy_i64_kernel(int64_t *__restrict__ dst, const int64_t *__restrict__ src, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

int check_fk_args(irt_ctx *ctx, const irt_robot *rb, const void *states, int state_size, int64_t n,
                  int cap_pts, const irt_fk_outputs *out) {
  if (!ctx || !rb || !out) return IRT_ERR_INVALID_ARGUMENT;
  if (n < 0) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "negative batch size");
  if (n > 0 && !states) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "null states");
  if (state_size != rb->state_size)  // TendonRobot.h:107-109 std::invalid_argument
    return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "State is not the right size (%d != %d)",
                    state_size, rb->state_size);
  // flags also need it: the self-collision test reads points staged at stride cap_pts
  if ((out->p || out->R || out->t || out->flags) && cap_pts < rb->max_points)
    return irt_fail(ctx, IRT_ERR_CAPACITY, "cap_pts=%d < max_points=%d", cap_pts, rb->max_points);
  return IRT_OK;
}

}  // namespace

extern "C" {

int irt_fk_batch_dev(irt_ctx *ctx, const irt_robot *rb, const double *d_states, int state_size,
                     int64_t n, int cap_pts, const irt_fk_outputs *d_out, void *stream) {
  int rc = check_fk_args(ctx, rb, d_states, state_size, n, cap_pts, d_out);
  if (rc) return rc;
  if (n == 0) return IRT_OK;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  irt_fk_outputs o = *d_out;
  if (o.flags && !(o.p && o.npts))
    return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT,
                    "flags output needs p and npts outputs (self-collision reads the points)");
  rc = fk_launch(ctx, rb, d_states, n, cap_pts, o, st);
  if (rc) return rc;
  if (o.flags) rc = self_collision_launch(ctx, rb, o.p, o.npts, n, cap_pts, o.flags, st);
  return rc;
}

int irt_fk_batch(irt_ctx *ctx, const irt_robot *rb, const double *states, int state_size,
                 int64_t n, int cap_pts, const irt_fk_outputs *out) {
  int rc = check_fk_args(ctx, rb, states, state_size, n, cap_pts, out);
  if (rc) return rc;
  if (n == 0) return IRT_OK;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
  const int N = rb->desc.n_tendons;
  // Chunked, double-buffered pipeline: the D2H copy of chunk c (copy stream) overlaps the
  // kernels of chunk c+1 (compute stream).  With pinned host buffers the copies are truly
  // asynchronous; with pageable buffers the runtime stages them and the overlap is partial.
  int64_t chunk = 96 * 1024;
  if (chunk > n) chunk = n;
  const bool need_p = out->p || out->flags;
  struct Stage {
    double *states, *p, *R, *t, *L, *Li, *tip, *uv;
    int32_t *npts, *iters, *nsteps;
    uint32_t *flags;
    cudaEvent_t computed, copied;
  } stage[2];
  std::memset(stage, 0, sizeof(stage));
  const int nstage = (n > chunk) ? 2 : 1;
  {  // carve both stages out of the context's grow-only staging buffer (no per-call cudaMalloc)
    auto carve = [&](char *base, size_t *total) {
      size_t used = 0;
      auto take = [&](size_t bytes) -> char * {
        char *ptr = base ? base + used : nullptr;
        used += (bytes + 255) & ~(size_t)255;
        return ptr;
      };
      for (int b = 0; b < nstage; b++) {
        Stage &s = stage[b];
        s.states = (double *)take((size_t)chunk * state_size * 8);
        s.p = need_p ? (double *)take((size_t)chunk * cap_pts * 24) : nullptr;
        s.R = out->R ? (double *)take((size_t)chunk * cap_pts * 72) : nullptr;
        s.t = out->t ? (double *)take((size_t)chunk * cap_pts * 8) : nullptr;
        s.npts = (int32_t *)take((size_t)chunk * 4);
        s.L = out->L ? (double *)take((size_t)chunk * 8) : nullptr;
        s.Li = out->L_i ? (double *)take((size_t)chunk * N * 8) : nullptr;
        s.tip = out->tip ? (double *)take((size_t)chunk * 24) : nullptr;
        s.uv = out->uv ? (double *)take((size_t)chunk * 96) : nullptr;
        s.flags = out->flags ? (uint32_t *)take((size_t)chunk * 4) : nullptr;
        s.iters = out->iters ? (int32_t *)take((size_t)chunk * 4) : nullptr;
        s.nsteps = out->nsteps ? (int32_t *)take((size_t)chunk * 4) : nullptr;
        s.computed = ctx->ev_computed[b];
        s.copied = ctx->ev_copied[b];
      }
      *total = used;
    };
    size_t total = 0;
    carve(nullptr, &total);
    char *base = (char *)ctx_io(ctx, total);
    if (!base) return irt_fail(ctx, IRT_ERR_CUDA, "device staging allocation of %zu bytes failed", total);
    carve(base, &total);
  }
  int64_t c = 0;
  for (int64_t off = 0; off < n; off += chunk, c++) {
    Stage &s = stage[c % nstage];
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    if (c >= nstage) IRT_CUDA(ctx, cudaStreamWaitEvent(st, s.copied, 0));  // buffers free again
    IRT_CUDA(ctx, cudaMemcpyAsync(s.states, states + off * state_size, (size_t)m * state_size * 8,
                                  cudaMemcpyHostToDevice, st));
    irt_fk_outputs o;
    std::memset(&o, 0, sizeof(o));
    o.p = s.p; o.R = s.R; o.t = s.t; o.npts = s.npts; o.L = s.L; o.L_i = s.Li;
    o.tip = s.tip; o.uv = s.uv; o.flags = s.flags; o.iters = s.iters; o.nsteps = s.nsteps;
    // rows beyond npts[i] are returned as zeros (deterministic padding)
    if (o.p) IRT_CUDA(ctx, cudaMemsetAsync(o.p, 0, (size_t)m * cap_pts * 24, st));
    if (o.R) IRT_CUDA(ctx, cudaMemsetAsync(o.R, 0, (size_t)m * cap_pts * 72, st));
    if (o.t) IRT_CUDA(ctx, cudaMemsetAsync(o.t, 0, (size_t)m * cap_pts * 8, st));
    rc = fk_launch(ctx, rb, s.states, m, cap_pts, o, st);
    if (rc) return rc;
    if (o.flags) {
      rc = self_collision_launch(ctx, rb, o.p, o.npts, m, cap_pts, o.flags, st);
      if (rc) return rc;
    }
    IRT_CUDA(ctx, cudaEventRecord(s.computed, st));
    IRT_CUDA(ctx, cudaStreamWaitEvent(cs, s.computed, 0));
#define D2H(dst, src, bytes_per)                                                               \
  if (dst) IRT_CUDA(ctx, cudaMemcpyAsync((char *)(dst) + (size_t)off * (bytes_per), (src),      \
                                         (size_t)m * (bytes_per), cudaMemcpyDeviceToHost, cs))
    D2H(out->p, s.p, (size_t)cap_pts * 24);
    D2H(out->R, s.R, (size_t)cap_pts * 72);
    D2H(out->t, s.t, (size_t)cap_pts * 8);
    D2H(out->npts, s.npts, 4);
    D2H(out->L, s.L, 8);
    D2H(out->L_i, s.Li, (size_t)N * 8);
    D2H(out->tip, s.tip, 24);
    D2H(out->uv, s.uv, 96);
    D2H(out->flags, s.flags, 4);
    D2H(out->iters, s.iters, 4);
    D2H(out->nsteps, s.nsteps, 4);
#undef D2H
    IRT_CUDA(ctx, cudaEventRecord(s.copied, cs));
  }
  IRT_CUDA(ctx, cudaStreamSynchronize(st));
  IRT_CUDA(ctx, cudaStreamSynchronize(cs));
  return IRT_OK;
}

// Packed form of irt_fk_batch: the per-point outputs p / R / t hold only the rows that exist
// (configuration i occupies rows [row_offsets[i], row_offsets[i+1])), like the reference's
// std::vector<Point> per shape (TendonResult.h:17-28) laid end to end.  With retraction the mean shape
// has ~3/4 of max_points rows, and the D2H copy of p is what bounds the host-pointer path (PCIe).
// Same chunked, double-buffered pipeline; the row offsets of a chunk are known before its FK kernel
// runs (count + scan take microseconds), so the host learns the chunk's size without waiting for it.
int irt_fk_batch_packed(irt_ctx *ctx, const irt_robot *rb, const double *states, int state_size,
                        int64_t n, const irt_fk_outputs *out, int64_t cap_rows, int64_t *row_offsets) {
  if (!row_offsets) return IRT_ERR_INVALID_ARGUMENT;
  int rc = check_fk_args(ctx, rb, states, state_size, n, rb ? rb->max_points : 0, out);
  if (rc) return rc;
  row_offsets[0] = 0;
  if (n == 0) return IRT_OK;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
  const int N = rb->desc.n_tendons, cap_pts = rb->max_points;
  int64_t chunk = 96 * 1024;
  if (chunk > n) chunk = n;
  const bool need_p = out->p || out->flags;
  struct Stage {
    double *states, *p, *R, *t, *L, *Li, *tip, *uv;
    int64_t *cnt, *off;
    int32_t *npts, *iters, *nsteps;
    uint32_t *flags;
    int64_t *h_off;  // pinned host copy of off[0..m]
    cudaEvent_t computed, copied, offsets;
  } stage[2];
  std::memset(stage, 0, sizeof(stage));
  const int nstage = (n > chunk) ? 2 : 1;
  {
    auto carve = [&](char *base, size_t *total) {
      size_t used = 0;
      auto take = [&](size_t bytes) -> char * {
        char *ptr = base ? base + used : nullptr;
        used += (bytes + 255) & ~(size_t)255;
        return ptr;
      };
      for (int b = 0; b < nstage; b++) {
        Stage &s = stage[b];
        s.states = (double *)take((size_t)chunk * state_size * 8);
        s.cnt = (int64_t *)take((size_t)chunk * 8);
        s.off = (int64_t *)take((size_t)(chunk + 1) * 8);
        s.p = need_p ? (double *)take((size_t)chunk * cap_pts * 24) : nullptr;
        s.R = out->R ? (double *)take((size_t)chunk * cap_pts * 72) : nullptr;
        s.t = out->t ? (double *)take((size_t)chunk * cap_pts * 8) : nullptr;
        s.npts = (int32_t *)take((size_t)chunk * 4);
        s.L = out->L ? (double *)take((size_t)chunk * 8) : nullptr;
        s.Li = out->L_i ? (double *)take((size_t)chunk * N * 8) : nullptr;
        s.tip = out->tip ? (double *)take((size_t)chunk * 24) : nullptr;
        s.uv = out->uv ? (double *)take((size_t)chunk * 96) : nullptr;
        s.flags = out->flags ? (uint32_t *)take((size_t)chunk * 4) : nullptr;
        s.iters = out->iters ? (int32_t *)take((size_t)chunk * 4) : nullptr;
        s.nsteps = out->nsteps ? (int32_t *)take((size_t)chunk * 4) : nullptr;
        s.computed = ctx->ev_computed[b];
        s.copied = ctx->ev_copied[b];
        s.offsets = ctx->ev_offsets[b];
      }
      *total = used;
    };
    size_t total = 0;
    carve(nullptr, &total);
    char *base = (char *)ctx_io(ctx, total);
    if (!base) return irt_fail(ctx, IRT_ERR_CUDA, "device staging allocation of %zu bytes failed", total);
    carve(base, &total);
    int64_t *pin = (int64_t *)ctx_pinned(ctx, (size_t)nstage * (chunk + 1) * 8);
    if (!pin) return irt_fail(ctx, IRT_ERR_CUDA, "pinned staging allocation failed");
    for (int b = 0; b < nstage; b++) stage[b].h_off = pin + (size_t)b * (chunk + 1);
  }
  int64_t row_base = 0, c = 0;
  for (int64_t off = 0; off < n; off += chunk, c++) {
    Stage &s = stage[c % nstage];
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    if (c >= nstage) IRT_CUDA(ctx, cudaStreamWaitEvent(st, s.copied, 0));  // buffers free again
    IRT_CUDA(ctx, cudaMemcpyAsync(s.states, states + off * state_size, (size_t)m * state_size * 8,
                                  cudaMemcpyHostToDevice, st));
    rc = fk_row_counts(ctx, rb, s.states, m, s.cnt, st);
    if (rc) return rc;
    row_offsets_kernel<<<1, 1024, 0, st>>>(s.cnt, s.off, m);
    IRT_LAUNCHED(ctx);
    // the offsets go to the host through SM stores into page-locked memory, not through the D2H copy
    // engine: a small cudaMemcpyAsync would queue behind the previous chunk's large p copy and hold back
    // this chunk's FK kernel (same stream) until that copy is done
    copy_i64_kernel<<<(unsigned)((m + 1 + 255) / 256), 256, 0, st>>>(s.h_off, s.off, m + 1);
    IRT_LAUNCHED(ctx);
    IRT_CUDA(ctx, cudaEventRecord(s.offsets, st));
    irt_fk_outputs o;
    std::memset(&o, 0, sizeof(o));
    o.p = s.p; o.R = s.R; o.t = s.t; o.npts = s.npts; o.L = s.L; o.L_i = s.Li;
    o.tip = s.tip; o.uv = s.uv; o.flags = s.flags; o.iters = s.iters; o.nsteps = s.nsteps;
    rc = fk_launch(ctx, rb, s.states, m, cap_pts, o, st, s.off);
    if (rc) return rc;
    if (o.flags) {
      rc = self_collision_launch(ctx, rb, o.p, o.npts, m, cap_pts, o.flags, st, s.off);
      if (rc) return rc;
    }
    IRT_CUDA(ctx, cudaEventRecord(s.computed, st));
    // the chunk's size: available as soon as count + scan are done (the FK kernel is still running)
    IRT_CUDA(ctx, cudaEventSynchronize(s.offsets));
    const int64_t rows = s.h_off[m];
    if (row_base + rows > cap_rows) {
      cudaStreamSynchronize(st);
      cudaStreamSynchronize(cs);
      return irt_fail(ctx, IRT_ERR_CAPACITY, "packed outputs need more than cap_rows=%lld rows",
                      (long long)cap_rows);
    }
    for (int64_t i = 0; i < m; i++) row_offsets[off + i + 1] = row_base + s.h_off[i + 1];
    IRT_CUDA(ctx, cudaStreamWaitEvent(cs, s.computed, 0));
#define D2H_ROWS(dst, src, bytes_per)                                                              \
  if (dst && rows > 0) IRT_CUDA(ctx, cudaMemcpyAsync((char *)(dst) + (size_t)row_base * (bytes_per), (src), \
                                                     (size_t)rows * (bytes_per), cudaMemcpyDeviceToHost, cs))
#define D2H(dst, src, bytes_per)                                                               \
  if (dst) IRT_CUDA(ctx, cudaMemcpyAsync((char *)(dst) + (size_t)off * (bytes_per), (src),      \
                                         (size_t)m * (bytes_per), cudaMemcpyDeviceToHost, cs))
    D2H_ROWS(out->p, s.p, 24);
    D2H_ROWS(out->R, s.R, 72);
    D2H_ROWS(out->t, s.t, 8);
    D2H(out->npts, s.npts, 4);
    D2H(out->L, s.L, 8);
    D2H(out->L_i, s.Li, (size_t)N * 8);
    D2H(out->tip, s.tip, 24);
    D2H(out->uv, s.uv, 96);
    D2H(out->flags, s.flags, 4);
    D2H(out->iters, s.iters, 4);
    D2H(out->nsteps, s.nsteps, 4);
#undef D2H
#undef D2H_ROWS
    IRT_CUDA(ctx, cudaEventRecord(s.copied, cs));
    row_base += rows;
  }
  IRT_CUDA(ctx, cudaStreamSynchronize(st));
  IRT_CUDA(ctx, cudaStreamSynchronize(cs));
  return IRT_OK;
}

// TendonRobot::home_shape(state).L_i closed forms (tendon/TendonRobot.cpp:249-314); trivial
// arithmetic on the host side of the boundary (no kernel needed).
int irt_home_lengths_batch(irt_ctx *ctx, const irt_robot *rb, const double *states,
                           int state_size, int64_t n, double *L_i) {
  if (!ctx || !rb || !L_i || (n > 0 && !states)) return IRT_ERR_INVALID_ARGUMENT;
  if (state_size != rb->state_size)
    return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "State is not the right size");
  const int N = rb->desc.n_tendons;
  for (int j = 0; j < N; j++)
    if (rb->dev.home_factor[j] != rb->dev.home_factor[j])
      return irt_fail(ctx, IRT_ERR_UNSUPPORTED,
                      "home length of tendon %d needs the reference's (out-of-bounds) Simpson rule", j);
  for (int64_t i = 0; i < n; i++) {
    double s = rb->desc.enable_retraction ? states[i * state_size + state_size - 1] : 0.0;
    if (s < 0.0) s = 0.0;
    if (s > rb->desc.L) s = rb->desc.L;
    const double Lres = (s == rb->desc.L) ? 0.0 : rb->desc.L - s;
    for (int j = 0; j < N; j++) L_i[i * N + j] = Lres * rb->dev.home_factor[j];
  }
  return IRT_OK;
}

}  // extern "C"
