// TEST INFRASTRUCTURE ONLY -- never linked into the product, never shipped.
//
// A stand-in for the C-ABI entry points (include/irt_b200.h) that the C++ host mirror
// (interactive-rate-tendons_b200/host/irt_host.hpp) goes through, answered by the CPU oracle, so that the
// HOST LOGIC above the boundary -- layout conversions (column-major R, Morton keys, CSR <-> block maps),
// status -> exception translation, home_shape / validators / PartialVoxelization assembly, createRoadmap's
// rejection rounds, connection strategy, edge removal and validity bookkeeping -- can be tested where there
// is no GPU (`pytest -m "not gpu"`, tests/test_abi_and_host.py).  The SAME test source
// (tests/cpp/test_host_mirror.cpp) is linked against the real libirt_b200.so on the GPU box
// (tests/test_gpu_host_cpp.py).  The product library has no CPU path: irt_ctx_create fails without a device.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/irt_b200.h"
#include "../../oracle/tendon_oracle.h"

static_assert(sizeof(orc_robot) == sizeof(irt_robot_desc), "POD mirrors must match");
static_assert(sizeof(orc_grid) == sizeof(irt_grid), "POD mirrors must match");
static_assert(sizeof(orc_space) == sizeof(irt_space), "POD mirrors must match");

struct irt_ctx { std::string err; };
struct irt_robot { orc_robot rb; };
struct irt_env { orc_grid g; orc_octree *t; };
struct irt_setstore { orc_grid g; orc_setstore *s; };

static void store_reset(irt_setstore *st, int64_t n) {
  if (st->s) orc_setstore_free(st->s);
  st->s = orc_setstore_new(&st->g, n);
}

extern "C" {

const char *irt_status_string(int status) { return status == IRT_OK ? "ok" : "error (stand-in)"; }
const char *irt_last_error(const irt_ctx *ctx) { return ctx->err.c_str(); }
int irt_ctx_create(int, irt_ctx **out) { *out = new irt_ctx; return IRT_OK; }
void irt_ctx_destroy(irt_ctx *ctx) { delete ctx; }

int irt_robot_create(irt_ctx *, const irt_robot_desc *desc, irt_robot **out) {
  *out = new irt_robot;
  std::memcpy(&(*out)->rb, desc, sizeof(orc_robot));
  return IRT_OK;
}
void irt_robot_destroy(irt_robot *rb) { delete rb; }

uint32_t irt_morton_key(int bx, int by, int bz, int Nb) { return orc_morton_key(bx, by, bz, Nb); }
void irt_morton_decode(uint32_t key, int Nb, int *bx, int *by, int *bz) {
  *bx = *by = *bz = 0;
  for (int l = 0; (1 << l) < Nb; l++) {
    *bx |= ((key >> (3 * l + 2)) & 1) << l;
    *by |= ((key >> (3 * l + 1)) & 1) << l;
    *bz |= ((key >> (3 * l)) & 1) << l;
  }
}

int irt_env_create(irt_ctx *, const irt_grid *grid, irt_env **out) {
  *out = new irt_env;
  std::memcpy(&(*out)->g, grid, sizeof(orc_grid));
  (*out)->t = orc_octree_new(&(*out)->g);
  return IRT_OK;
}
void irt_env_destroy(irt_env *env) {
  if (!env) return;
  orc_octree_free(env->t);
  delete env;
}
int irt_env_update_sparse(irt_ctx *, irt_env *env, const uint8_t *bxyz, const uint64_t *bits, int64_t n) {
  orc_octree_clear(env->t);
  for (int64_t i = 0; i < n; i++) orc_octree_set_block(env->t, bxyz[3 * i], bxyz[3 * i + 1], bxyz[3 * i + 2], bits[i]);
  return IRT_OK;
}

int irt_setstore_create(irt_ctx *, const irt_grid *grid, irt_setstore **out) {
  *out = new irt_setstore;
  std::memcpy(&(*out)->g, grid, sizeof(orc_grid));
  (*out)->s = nullptr;
  return IRT_OK;
}
void irt_setstore_destroy(irt_setstore *s) {
  if (!s) return;
  if (s->s) orc_setstore_free(s->s);
  delete s;
}
int64_t irt_setstore_num_sets(const irt_setstore *s) { return s->s ? orc_setstore_size(s->s) : 0; }

int irt_voxelize_vertices(irt_ctx *, const irt_robot *rb, const double *states, int state_size, int64_t n,
                          irt_setstore *store, uint32_t *flags, double *tips) {
  if (state_size != orc_state_size(&rb->rb)) return IRT_ERR_INVALID_ARGUMENT;
  store_reset(store, n);
  orc_voxelize_vertices_batch(&rb->rb, &store->g, states, n, store->s, flags, orc_max_threads());
  if (tips) {
    std::vector<double> t(1024), p(3 * 1024);
    for (int64_t i = 0; i < n; i++) {
      orc_fk_out fo;
      int m = orc_shape(&rb->rb, states + i * state_size, 1024, t.data(), p.data(), nullptr, &fo);
      for (int c = 0; c < 3; c++) tips[3 * i + c] = m > 0 ? p[3 * (m - 1) + c] : 0.0;
    }
  }
  return IRT_OK;
}

int irt_voxelize_edges(irt_ctx *, const irt_robot *rb, const irt_space *space, const double *a, const double *b,
                       int state_size, int64_t n, irt_setstore *store, uint32_t *flags, double *t_last,
                       int32_t *nsamples) {
  if (state_size != orc_state_size(&rb->rb)) return IRT_ERR_INVALID_ARGUMENT;
  store_reset(store, n);
  orc_space sp;
  std::memcpy(&sp, space, sizeof(sp));
  orc_voxelize_edges_batch(&rb->rb, &store->g, &sp, a, b, n, store->s, flags, t_last, nsamples, orc_max_threads());
  return IRT_OK;
}

int irt_robot_max_points(const irt_robot *rb) {
  std::vector<double> t(1 << 16);
  return orc_t_range(0.0, rb->rb.L, rb->rb.dL, t.data(), (int)t.size());
}

int irt_fk_batch(irt_ctx *ctx, const irt_robot *rb, const double *states, int state_size, int64_t n, int cap_pts,
                 const irt_fk_outputs *o) {
  if (state_size != orc_state_size(&rb->rb)) { ctx->err = "State is not the right size"; return IRT_ERR_INVALID_ARGUMENT; }
  if (cap_pts < irt_robot_max_points(rb)) return IRT_ERR_CAPACITY;
  const int N = rb->rb.n_tendons;
  std::vector<double> t(cap_pts), p(3 * cap_pts), R(9 * cap_pts);
  for (int64_t i = 0; i < n; i++) {
    orc_fk_out fo;
    const int m = orc_shape(&rb->rb, states + i * state_size, cap_pts, t.data(), p.data(), R.data(), &fo);
    if (o->p) std::memcpy(o->p + i * cap_pts * 3, p.data(), sizeof(double) * 3 * m);
    if (o->R) std::memcpy(o->R + i * cap_pts * 9, R.data(), sizeof(double) * 9 * m);
    if (o->t) std::memcpy(o->t + i * cap_pts, t.data(), sizeof(double) * m);
    if (o->npts) o->npts[i] = m;
    if (o->L) o->L[i] = fo.L;
    if (o->L_i) std::memcpy(o->L_i + i * N, fo.L_i, sizeof(double) * N);
    if (o->tip) for (int c = 0; c < 3; c++) o->tip[3 * i + c] = m > 0 ? p[3 * (m - 1) + c] : 0.0;
    if (o->uv) {
      std::memcpy(o->uv + 12 * i, fo.u_i, 24); std::memcpy(o->uv + 12 * i + 3, fo.u_f, 24);
      std::memcpy(o->uv + 12 * i + 6, fo.v_i, 24); std::memcpy(o->uv + 12 * i + 9, fo.v_f, 24);
    }
    if (o->flags) o->flags[i] = orc_validity_flags(&rb->rb, states + i * state_size, &fo, p.data());
    if (o->iters) o->iters[i] = fo.iters;
    if (o->nsteps) o->nsteps[i] = fo.nsteps;
  }
  return IRT_OK;
}

int irt_fk_tip_jacobian_batch(irt_ctx *, const irt_robot *rb, const double *states, int state_size, int64_t n,
                              int mode, double delta, double *tips, double *J) {
  if (state_size != orc_state_size(&rb->rb)) return IRT_ERR_INVALID_ARGUMENT;
  for (int64_t i = 0; i < n; i++) {
    double tip[3];
    orc_tip_jacobian(&rb->rb, states + i * state_size, mode, delta, tip, J + i * 3 * state_size);
    if (tips) std::memcpy(tips + 3 * i, tip, 24);
  }
  return IRT_OK;
}

int irt_home_lengths_batch(irt_ctx *, const irt_robot *rb, const double *states, int state_size, int64_t n,
                           double *L_i) {
  if (state_size != orc_state_size(&rb->rb)) return IRT_ERR_INVALID_ARGUMENT;
  for (int64_t i = 0; i < n; i++)
    orc_home_lengths(&rb->rb, rb->rb.enable_retraction ? states[i * state_size + state_size - 1] : 0.0,
                     L_i + i * rb->rb.n_tendons);
  return IRT_OK;
}

uint32_t irt_valid_segment_count(const irt_robot_desc *desc, const irt_space *space, const double *a,
                                 const double *b) {
  orc_robot rb;
  orc_space sp;
  std::memcpy(&rb, desc, sizeof(rb));
  std::memcpy(&sp, space, sizeof(sp));
  return orc_valid_segment_count(&rb, &sp, a, b);
}

int irt_voxelize_shapes(irt_ctx *, const double *p, const int32_t *npts, int cap_pts, int64_t n,
                        irt_setstore *store) {
  store_reset(store, n);
  for (int64_t i = 0; i < n; i++)
    orc_voxelize_shape(&store->g, p + i * cap_pts * 3, npts[i], orc_setstore_get(store->s, i));
  return IRT_OK;
}

int irt_voxelize_edges_until_invalid(irt_ctx *, const irt_robot *rb, const irt_space *space, const double *a,
                                     const double *b, int state_size, int64_t n, const irt_env *env,
                                     irt_setstore *store, uint32_t *flags, double *t_last, int32_t *nsamples) {
  if (state_size != orc_state_size(&rb->rb)) return IRT_ERR_INVALID_ARGUMENT;
  store_reset(store, n);
  orc_space sp;
  std::memcpy(&sp, space, sizeof(sp));
  for (int64_t i = 0; i < n; i++) {
    orc_edge_out info;
    orc_voxelize_edge(&rb->rb, &store->g, &sp, a + i * state_size, b + i * state_size, env->t,
                      orc_setstore_get(store->s, i), &info);
    if (flags) flags[i] = (info.is_fully_valid ? 0u : IRT_FLAG_PARTIAL) | (info.out_of_domain ? IRT_FLAG_OUT_OF_DOMAIN : 0u);
    if (t_last) t_last[i] = info.t;
    if (nsamples) nsamples[i] = info.nsamples;
  }
  return IRT_OK;
}

int64_t irt_setstore_num_blocks(const irt_setstore *s) { return s->s ? orc_setstore_total_blocks(s->s) : 0; }
int irt_setstore_import(irt_ctx *, irt_setstore *s, int64_t n_sets, const uint64_t *offsets, const uint32_t *keys,
                        const uint64_t *bits) {
  store_reset(s, n_sets);
  for (int64_t i = 0; i < n_sets; i++)
    for (uint64_t j = offsets[i]; j < offsets[i + 1]; j++) {
      int bx, by, bz;
      irt_morton_decode(keys[j], s->g.Ng / 4, &bx, &by, &bz);
      orc_octree_set_block(orc_setstore_get(s->s, i), bx, by, bz, bits[j]);
    }
  return IRT_OK;
}
int irt_setstore_export(irt_ctx *, const irt_setstore *s, uint64_t *offsets, uint32_t *keys, uint64_t *bits) {
  if (!s->s) { offsets[0] = 0; return IRT_OK; }
  orc_setstore_export(s->s, offsets, keys, bits);
  return IRT_OK;
}

int irt_env_add_primitives(irt_ctx *, irt_env *env, const double *points, int64_t n_points, const double *spheres,
                           int64_t n_spheres, const double *capsules, int64_t n_capsules, int clear_first) {
  if (clear_first) orc_octree_clear(env->t);
  for (int64_t i = 0; i < n_points; i++) orc_octree_add_point(env->t, points + 3 * i);
  for (int64_t i = 0; i < n_spheres; i++) orc_octree_add_sphere(env->t, spheres + 4 * i, spheres[4 * i + 3]);
  for (int64_t i = 0; i < n_capsules; i++)
    orc_octree_add_capsule(env->t, capsules + 7 * i, capsules + 7 * i + 3, capsules[7 * i + 6]);
  return IRT_OK;
}
int irt_env_dilate(irt_ctx *, irt_env *env, int num, int use_diagonal) {
  if (use_diagonal) orc_octree_dilate_27neighbor(env->t, num); else orc_octree_dilate_6neighbor(env->t, num);
  return IRT_OK;
}
int irt_env_dilate_sphere(irt_ctx *, irt_env *env, double r) { orc_octree_dilate_sphere(env->t, r); return IRT_OK; }
int irt_env_remove_interior(irt_ctx *, irt_env *env, int keep_diagonal) {
  orc_octree_remove_interior(env->t, keep_diagonal);
  return IRT_OK;
}
int irt_env_download(irt_ctx *, const irt_env *env, uint64_t *blocks) {
  const int Nb = env->g.Ng / 4;
  const int64_t nb = orc_octree_nblocks(env->t);
  std::memset(blocks, 0, sizeof(uint64_t) * (size_t)Nb * Nb * Nb);
  std::vector<uint8_t> xyz(3 * nb + 3);
  std::vector<uint64_t> bits(nb + 1);
  orc_octree_export(env->t, nb, xyz.data(), bits.data());
  for (int64_t i = 0; i < nb; i++) blocks[orc_morton_key(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], Nb)] = bits[i];
  return IRT_OK;
}

int irt_check_sets(irt_ctx *ctx, const irt_setstore *store, const irt_env *env, int64_t begin, int64_t end,
                   uint32_t *verdict_words) {
  if (store->g.Ng != env->g.Ng) { ctx->err = "voxel dimension mismatch"; return IRT_ERR_INVALID_ARGUMENT; }
  std::vector<uint8_t> v(end - begin + 1);
  orc_check_sets_batch(store->s, env->t, begin, end, v.data(), orc_max_threads());
  for (int64_t w = 0; w < (end - begin + 31) / 32; w++) verdict_words[w] = 0;
  for (int64_t i = 0; i < end - begin; i++)
    if (v[i]) verdict_words[i >> 5] |= 1u << (i & 31);
  return IRT_OK;
}

}  // extern "C"
