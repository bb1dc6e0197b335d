// jacobian.cu -- batched finite-difference tip Jacobians.
//
// The reference's IK / tip controllers differentiate the tip position by finite differences, one FK
// per perturbed parameter and per call:
//   * tip_control::Jacobian (tip-control/tip_control.cpp:243-265): forward difference with a fixed
//     step `dist` over every state entry, J(j,i) = (fk(state + dist e_i).back()[j] - ps[j]) / dist;
//   * levmar's dlevmar_bc_dif behind tip_control::inverse_kinematics_impl (tip_control.cpp:34-153),
//     which evaluates the wrapper `fk_wrap` (:92-122: a retraction beyond L returns (0,0,L-s))
//     with levmar-2.6 misc_core.c:137-211: d = max(|1e-4 p_j|, delta),
//     forward  jac[i*m+j] = (f(p + d e_j)[i] - f(p)[i]) * (1/d),
//     central  p_j = tmp - d -> hxm;  p_j = tmp + d -> hxp;  jac[i*m+j] = (hxp[i] - hxm[i]) * (0.5/d)
//     (the reference asks for the central form: levmar_opt[4] = -delta, :85).
// roadmapIk (motion-planning/VoxelCachedLazyPRM.cpp:3095-3205) runs that IK from k roadmap neighbours,
// one after the other.  Here the n x (S+1) or n x (2S+1) perturbed states of a whole batch of seeds are
// generated on the device, go through ONE K1 launch (tip output only) and are differenced by one small
// kernel: a Jacobian batch costs one FK batch.
#include "common.cuh"

namespace {

// evaluation e of seed i: e = 0 the seed itself; forward modes: e = 1 + j -> +d on entry j;
// central mode: e = 1 + 2 j -> -d, e = 2 + 2 j -> +d
__device__ __forceinline__ double fd_step(int mode, double pj, double delta) {
  if (mode == IRT_JAC_FORWARD_FIXED) return delta;
  double d = 1E-04 * pj;   // levmar misc_core.c:155-158 / :193-196
  d = fabs(d);
  if (d < delta) d = delta;
  return d;
}

__global__ void jac_perturb_kernel(const double *__restrict__ states, int S, int64_t n, int mode,
                                   double delta, int evals, double *__restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (seed, eval)
  if (idx >= n * evals) return;
  const int64_t i = idx / evals;
  const int e = (int)(idx - i * evals);
  const double *src = states + i * S;
  double *dst = out + idx * S;
  for (int k = 0; k < S; k++) dst[k] = src[k];
  if (e == 0) return;
  if (mode == IRT_JAC_LEVMAR_CENTRAL) {
    const int j = (e - 1) >> 1;
    const double tmp = src[j], d = fd_step(mode, tmp, delta);
    dst[j] = ((e - 1) & 1) ? tmp + d : tmp - d;
  } else {
    const int j = e - 1;
    const double tmp = src[j], d = fd_step(mode, tmp, delta);
    dst[j] = tmp + d;
  }
}

// fk_wrap of tip_control.cpp:92-122 for the levmar modes: a retraction beyond L is not evaluated,
// the value is (0, 0, L - s)
__device__ __forceinline__ void wrapped_tip(const double *tips, const double *pert, int64_t row, int S,
                                            int mode, int retract, double L, double (&x)[3]) {
  x[0] = tips[row * 3]; x[1] = tips[row * 3 + 1]; x[2] = tips[row * 3 + 2];
  if (mode != IRT_JAC_FORWARD_FIXED && retract) {
    const double s = pert[row * S + S - 1];
    if (s > L) { x[0] = 0.0; x[1] = 0.0; x[2] = L - s; }
  }
}

__global__ void jac_diff_kernel(const double *__restrict__ states, const double *__restrict__ pert,
                                const double *__restrict__ tips, int S, int64_t n, int mode,
                                double delta, int evals, int retract, double L,
                                double *__restrict__ tip_out, double *__restrict__ J) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (seed, column)
  if (idx >= n * S) return;
  const int64_t i = idx / S;
  const int j = (int)(idx - i * S);
  const int64_t base = i * evals;
  double f0[3];
  wrapped_tip(tips, pert, base, S, mode, retract, L, f0);
  if (j == 0 && tip_out) { tip_out[i * 3] = f0[0]; tip_out[i * 3 + 1] = f0[1]; tip_out[i * 3 + 2] = f0[2]; }
  const double d = fd_step(mode, states[i * S + j], delta);
  double a[3], b[3];
  if (mode == IRT_JAC_LEVMAR_CENTRAL) {
    wrapped_tip(tips, pert, base + 1 + 2 * j, S, mode, retract, L, a);   // hxm
    wrapped_tip(tips, pert, base + 2 + 2 * j, S, mode, retract, L, b);   // hxp
    const double w = 0.5 / d;
    for (int r = 0; r < 3; r++) J[(i * 3 + r) * S + j] = (b[r] - a[r]) * w;
  } else {
    wrapped_tip(tips, pert, base + 1 + j, S, mode, retract, L, b);
    if (mode == IRT_JAC_FORWARD_FIXED) {
      for (int r = 0; r < 3; r++) J[(i * 3 + r) * S + j] = (b[r] - f0[r]) / d;
    } else {
      const double w = 1.0 / d;
      for (int r = 0; r < 3; r++) J[(i * 3 + r) * S + j] = (b[r] - f0[r]) * w;
    }
  }
}

}  // namespace

extern "C" {

int irt_fk_tip_jacobian_batch_dev(irt_ctx *ctx, const irt_robot *rb, const double *d_states,
                                  int state_size, int64_t n, int mode, double delta,
                                  double *d_tips, double *d_J, void *stream) {
  if (!ctx || !rb || !d_J) return IRT_ERR_INVALID_ARGUMENT;
  if (n < 0) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "negative batch size");
  if (n > 0 && !d_states) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "null states");
  if (state_size != rb->state_size)
    return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "State is not the right size (%d != %d)",
                    state_size, rb->state_size);
  if (mode != IRT_JAC_FORWARD_FIXED && mode != IRT_JAC_LEVMAR_FORWARD && mode != IRT_JAC_LEVMAR_CENTRAL)
    return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "unknown Jacobian mode %d", mode);
  if (!(delta > 0.0)) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "finite-difference step must be > 0");
  if (n == 0) return IRT_OK;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  const int S = state_size;
  const int evals = 1 + ((mode == IRT_JAC_LEVMAR_CENTRAL) ? 2 * S : S);
  const int64_t m = n * evals;
  const size_t bytes_states = (size_t)m * S * 8, bytes_tips = (size_t)m * 24;
  // the context's arena (K2's sample pool, idle here): fk_launch itself uses ctx_scratch for bucketing
  char *scr = (char *)ctx_arena(ctx, bytes_states + bytes_tips + 512);
  if (!scr) return irt_fail(ctx, IRT_ERR_CUDA, "arena allocation failed");
  double *pert = (double *)scr;
  double *tips = (double *)(scr + ((bytes_states + 255) & ~(size_t)255));
  const int T = 256;
  jac_perturb_kernel<<<(unsigned)((m + T - 1) / T), T, 0, st>>>(d_states, S, n, mode, delta, evals, pert);
  IRT_LAUNCHED(ctx);
  irt_fk_outputs o{};
  o.tip = tips;
  int rc = fk_launch(ctx, rb, pert, m, rb->max_points, o, st);
  if (rc) return rc;
  jac_diff_kernel<<<(unsigned)((n * S + T - 1) / T), T, 0, st>>>(
      d_states, pert, tips, S, n, mode, delta, evals, rb->desc.enable_retraction, rb->desc.L, d_tips, d_J);
  IRT_LAUNCHED(ctx);
  IRT_CUDA(ctx, cudaGetLastError());
  return IRT_OK;
}

int irt_fk_tip_jacobian_batch(irt_ctx *ctx, const irt_robot *rb, const double *states, int state_size,
                              int64_t n, int mode, double delta, double *tips, double *J) {
  if (!ctx || !rb || !J) return IRT_ERR_INVALID_ARGUMENT;
  if (n < 0) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "negative batch size");
  if (n > 0 && !states) return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "null states");
  if (state_size != rb->state_size)
    return irt_fail(ctx, IRT_ERR_INVALID_ARGUMENT, "State is not the right size (%d != %d)",
                    state_size, rb->state_size);
  if (n == 0) return IRT_OK;
  IRT_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int S = state_size;
  const size_t b_st = (size_t)n * S * 8, b_tip = (size_t)n * 24, b_J = (size_t)n * 3 * S * 8;
  char *io = (char *)ctx_io(ctx, b_st + b_tip + b_J + 1024);
  if (!io) return irt_fail(ctx, IRT_ERR_CUDA, "device staging allocation failed");
  double *d_states = (double *)io;
  double *d_tips = (double *)(io + ((b_st + 255) & ~(size_t)255));
  double *d_J = (double *)((char *)d_tips + ((b_tip + 255) & ~(size_t)255));
  IRT_CUDA(ctx, cudaMemcpyAsync(d_states, states, b_st, cudaMemcpyHostToDevice, st));
  int rc = irt_fk_tip_jacobian_batch_dev(ctx, rb, d_states, S, n, mode, delta, d_tips, d_J, st);
  if (rc) return rc;
  if (tips) IRT_CUDA(ctx, cudaMemcpyAsync(tips, d_tips, b_tip, cudaMemcpyDeviceToHost, st));
  IRT_CUDA(ctx, cudaMemcpyAsync(J, d_J, b_J, cudaMemcpyDeviceToHost, st));
  IRT_CUDA(ctx, cudaStreamSynchronize(st));
  return IRT_OK;
}

}  // extern "C"
