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

int check_fk_args(irt_ctx *ctx, const irt_robot *rb, const void *states, int state_size, int64_t n,
                  int cap_pts, const irt_fk_outputs *out) {
  if (!ctx || !rb || !out) return IRT_ERR_INVALID_ARGUMENT;
  if (n < 0) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "negative batch size");
  if (n > 0 && !states) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "null states");
  if (state_size != rb->state_size)  // TendonRobot.h:107-109 std::invalid_argument
    return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "State is not the right size (%d != %d)",
                    state_size, rb->state_size);
  if ((out->p || out->R || out->t) && cap_pts < rb->max_points)
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
  rc = fk_launch(ctx, rb, d_states, n, cap_pts, o, nullptr, st);
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
  cudaStream_t st = ctx->stream;
  const int N = rb->desc.n_tendons;
  // chunk so that device staging stays bounded (<= ~2 GiB of points per chunk)
  const int64_t per_cfg = (int64_t)cap_pts * (3 + (out->R ? 9 : 0) + (out->t ? 1 : 0)) * 8 + 256;
  int64_t chunk = (int64_t)(2048LL << 20) / per_cfg;
  if (chunk < 1024) chunk = 1024;
  if (chunk > n) chunk = n;
  const bool need_p = out->p || out->flags;
  DevBuf d_states, d_p, d_R, d_t, d_npts, d_L, d_Li, d_tip, d_uv, d_flags, d_iters, d_nsteps;
  bool ok = d_states.alloc((size_t)chunk * state_size * 8);
  if (need_p) ok = ok && d_p.alloc((size_t)chunk * cap_pts * 24);
  if (out->R) ok = ok && d_R.alloc((size_t)chunk * cap_pts * 72);
  if (out->t) ok = ok && d_t.alloc((size_t)chunk * cap_pts * 8);
  ok = ok && d_npts.alloc((size_t)chunk * 4);
  if (out->L) ok = ok && d_L.alloc((size_t)chunk * 8);
  if (out->L_i) ok = ok && d_Li.alloc((size_t)chunk * N * 8);
  if (out->tip) ok = ok && d_tip.alloc((size_t)chunk * 24);
  if (out->uv) ok = ok && d_uv.alloc((size_t)chunk * 96);
  if (out->flags) ok = ok && d_flags.alloc((size_t)chunk * 4);
  if (out->iters) ok = ok && d_iters.alloc((size_t)chunk * 4);
  if (out->nsteps) ok = ok && d_nsteps.alloc((size_t)chunk * 4);
  if (!ok) return irt_fail(ctx, IRT_ERR_CUDA, "device staging allocation failed");

  for (int64_t off = 0; off < n; off += chunk) {
    const int64_t m = (n - off < chunk) ? (n - off) : chunk;
    IRT_CUDA(ctx, cudaMemcpyAsync(d_states.p, states + off * state_size, (size_t)m * state_size * 8,
                                  cudaMemcpyHostToDevice, st));
    irt_fk_outputs o;
    std::memset(&o, 0, sizeof(o));
    o.p = (double *)d_p.p; o.R = (double *)d_R.p; o.t = (double *)d_t.p;
    o.npts = (int32_t *)d_npts.p; o.L = (double *)d_L.p; o.L_i = (double *)d_Li.p;
    o.tip = (double *)d_tip.p; o.uv = (double *)d_uv.p; o.flags = (uint32_t *)d_flags.p;
    o.iters = (int32_t *)d_iters.p; o.nsteps = (int32_t *)d_nsteps.p;
    // rows beyond npts[i] are returned as zeros (deterministic padding)
    if (o.p) IRT_CUDA(ctx, cudaMemsetAsync(o.p, 0, (size_t)m * cap_pts * 24, st));
    if (o.R) IRT_CUDA(ctx, cudaMemsetAsync(o.R, 0, (size_t)m * cap_pts * 72, st));
    if (o.t) IRT_CUDA(ctx, cudaMemsetAsync(o.t, 0, (size_t)m * cap_pts * 8, st));
    rc = fk_launch(ctx, rb, (const double *)d_states.p, m, cap_pts, o, nullptr, st);
    if (rc) return rc;
    if (o.flags) {
      rc = self_collision_launch(ctx, rb, o.p, o.npts, m, cap_pts, o.flags, st);
      if (rc) return rc;
    }
#define D2H(dst, src, bytes_per)                                                               \
  if (dst) IRT_CUDA(ctx, cudaMemcpyAsync((char *)(dst) + (size_t)off * (bytes_per), (src),      \
                                         (size_t)m * (bytes_per), cudaMemcpyDeviceToHost, st))
    D2H(out->p, d_p.p, (size_t)cap_pts * 24);
    D2H(out->R, d_R.p, (size_t)cap_pts * 72);
    D2H(out->t, d_t.p, (size_t)cap_pts * 8);
    D2H(out->npts, d_npts.p, 4);
    D2H(out->L, d_L.p, 8);
    D2H(out->L_i, d_Li.p, (size_t)N * 8);
    D2H(out->tip, d_tip.p, 24);
    D2H(out->uv, d_uv.p, 96);
    D2H(out->flags, d_flags.p, 4);
    D2H(out->iters, d_iters.p, 4);
    D2H(out->nsteps, d_nsteps.p, 4);
#undef D2H
    IRT_CUDA(ctx, cudaStreamSynchronize(st));
  }
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
