// TEST INFRASTRUCTURE ONLY -- never linked into the product, never shipped.
//
// A stand-in for the handful of C-ABI entry points (include/irt_b200.h) that the C++ host mirror's
// VoxelCachedLazyPRM::createRoadmap goes through, answered by the CPU oracle, so that the HOST LOGIC above
// the boundary (rejection sampling rounds, connection strategy, edge removal, validity bookkeeping, growing
// a roadmap) can be tested where there is no GPU (`pytest -m "not gpu"`, tests/test_abi_and_host.py).
// The same checks (tests/cpp/create_roadmap_checks.hpp) run against the real libirt_b200.so on the GPU box
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

int irt_check_sets(irt_ctx *, const irt_setstore *store, const irt_env *env, int64_t begin, int64_t end,
                   uint32_t *verdict_words) {
  if (store->g.Ng != env->g.Ng) return IRT_ERR_INVALID_ARGUMENT;
  std::vector<uint8_t> v(end - begin + 1);
  orc_check_sets_batch(store->s, env->t, begin, end, v.data(), orc_max_threads());
  for (int64_t w = 0; w < (end - begin + 31) / 32; w++) verdict_words[w] = 0;
  for (int64_t i = 0; i < end - begin; i++)
    if (v[i]) verdict_words[i >> 5] |= 1u << (i & 31);
  return IRT_OK;
}

}  // extern "C"
