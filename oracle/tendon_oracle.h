/*
 * tendon_oracle.h -- C API of the CPU ORACLE.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It is a from-scratch CPU
 * restatement of the reference algorithm (Kuntz-Lab/interactive-rate-tendons)
 * for the FK / voxelise / voxel-check hot path.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load
 * it.  The product path (interactive-rate-tendons_b200/) never links, imports or
 * calls anything in this directory.
 *
 * PARITY STATUS.  The reference ships no tests, fixtures or golden vectors and
 * cannot be compiled as a whole in this image (needs Eigen3, Boost.odeint, OMPL,
 * FCL, ITK ...).  The pieces of it that DO compile here pin this oracle
 * (oracle/_ref, oracle/ref.py, tests/test_oracle_vs_reference.py):
 *   - collision/detail/TreeNode.h as is: octree storage, set algebra, collides,
 *     visit_leaves order;
 *   - tendon_deriv.cpp, solve_initial_bending.cpp, get_r_info.cpp and
 *     collision_primitives.{h,cpp}, unmodified, against a stand-in for the Eigen
 *     subset they use (pins formulas / operand order / control flow, not Eigen's
 *     rounding): routing bit-exact, derivative identical, same fixed-point
 *     iteration counts;
 *   - the hot-path core of VoxelOctree.{h,cpp} (add_line, find_cell, cells,
 *     dilate_*, remove_interior_*), collides_self (collision.cpp) and
 *     VoxelEnvironment::voxelize_valid_backbone_motion, cut out of the reference's
 *     files by function anchors at build time and compiled unmodified: voxel
 *     sets, verdicts, swept volumes, FK call counts identical;
 *   - TendonRobot.h as is plus tension_shape / home_shape / calc_point_forces /
 *     is_valid cut out of TendonRobot.cpp, with a second stand-in for the two
 *     Boost.odeint facilities they use: grids, points, frames, lengths and all
 *     validity flags identical;
 *   - the reference's vendored levmar-2.6: finite-difference Jacobian rule.
 * "Parity unpinned by the reference" still holds for third-party arithmetic that
 * is not under /root/reference: Eigen's rounding, Boost.odeint's stepping rule
 * (restated here and in the stand-in) and OMPL 1.5 validSegmentCount /
 * interpolate.  Those are pinned by
 *   (1) analytic known-answer tests (tests/test_oracle_kat.py) and
 *   (2) an independent numpy + mpmath restatement (oracle/fk_second_opinion.py).
 *
 * All file:line citations are relative to /root/reference/cpp/src/.
 */
#ifndef TENDON_ORACLE_H
#define TENDON_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_TENDONS 12
#define ORC_MAX_COEF 8

/* per-item flag word (same bit meaning as include/irt_b200.h, restated) */
#define ORC_FLAG_NONCONVERGED 1u
#define ORC_FLAG_LENGTH_LIMIT 2u
#define ORC_FLAG_SELF_COLLISION 4u
#define ORC_FLAG_OUT_OF_DOMAIN 8u
#define ORC_FLAG_PARTIAL 16u

/* tendon/TendonRobot.h:52-58, tendon/BackboneSpecs.h:15-21, tendon/TendonSpecs.h:24-30 */
typedef struct orc_robot {
  double r;                  /* robot radius */
  double L, dL, ro, ri, E, nu;
  double residual_threshold;
  int32_t n_tendons;
  int32_t n_c;               /* #theta coefficients (all tendons share tendon 0's size) */
  int32_t n_d;               /* #rho coefficients */
  int32_t enable_rotation;
  int32_t enable_retraction;
  int32_t _pad;
  double C[ORC_MAX_TENDONS * ORC_MAX_COEF]; /* row-major [tendon][coef] */
  double D[ORC_MAX_TENDONS * ORC_MAX_COEF];
  double max_tension[ORC_MAX_TENDONS];
  double min_length[ORC_MAX_TENDONS];
  double max_length[ORC_MAX_TENDONS];
} orc_robot;

/* collision/VoxelOctree.h:310-329 limits + motion-planning/VoxelEnvironment.h inv_rotation */
typedef struct orc_grid {
  int32_t Ng;                /* voxels per axis: 4..512, power of two */
  int32_t _pad;
  double lim[6];             /* xmin,xmax,ymin,ymax,zmin,zmax */
  double inv_rot[9];         /* row-major 3x3 applied to points before voxelising */
} orc_grid;

/* motion-planning/Problem.h:59-63 */
typedef struct orc_space {
  double min_tension_change;
  double min_rotation_change;
  double min_retraction_change;
} orc_space;

typedef struct orc_fk_out {
  int32_t npts;
  int32_t converged;
  int32_t iters;             /* fixed-point iterations used */
  int32_t nsteps;            /* RK4 steps taken */
  double L;
  double L_i[ORC_MAX_TENDONS];
  double u_i[3], u_f[3], v_i[3], v_f[3];
} orc_fk_out;

int orc_state_size(const orc_robot *rb);
int orc_t_range(double s, double L, double dL, double *out, int cap);
void orc_routing(const orc_robot *rb, double t, double *r, double *rd, double *rdd); /* [N][3] each */
void orc_tendon_deriv(const orc_robot *rb, const double *tau, const double *x, double t,
                      double *dxdt);
/* same ODE with the dense 6x6 solved by Gaussian elimination w/ partial pivoting:
 * stands in for tendon_deriv_unopt / linsubsolve1 (tendon_deriv.cpp:40-56,180-261) */
void orc_tendon_deriv_alt(const orc_robot *rb, const double *tau, const double *x, double t,
                          double *dxdt);

/* TendonRobot::shape(state)  (tendon/TendonRobot.h:105-131).  t,p,R may be NULL.
 * returns npts, or -1 if cap_pts too small, -2 on bad sizes. */
int orc_shape(const orc_robot *rb, const double *state, int cap_pts, double *t, double *p,
              double *R, orc_fk_out *out);
/* finite-difference tip Jacobian J[3][m] (levmar's jac[i*m+j] layout): mode 0 tip_control::Jacobian
 * (tip-control/tip_control.cpp:243-265), 1/2 levmar-2.6 forward/central rule (misc_core.c:137-211)
 * applied to fk_wrap (tip_control.cpp:92-122) */
void orc_tip_jacobian(const orc_robot *rb, const double *state, int mode, double delta, double *tip,
                      double *J);
void orc_home_lengths(const orc_robot *rb, double s_start, double *L_i);
int orc_collides_self(const double *p, int npts, double r);
void orc_closest_st_segment(const double *A, const double *B, const double *C, const double *D,
                            double *s, double *t);
/* validity flag word for one computed shape (AbstractValidityChecker.cpp:99-114),
 * every test evaluated independently */
/* collision/collision_primitives.h:62-85 */
int orc_segment_aabox_intersect(const double *A, const double *B, const double *C, const double *D);
uint32_t orc_validity_flags(const orc_robot *rb, const double *state, const orc_fk_out *fk,
                            const double *p);

/* batched FK, OpenMP over configs (apps/estimate_length_discretization.cpp:62-71) */
void orc_fk_batch(const orc_robot *rb, const double *states, int64_t n, int cap_pts, double *p,
                  int32_t *npts, double *L_i, double *tip, uint32_t *flags, int32_t *iters,
                  int32_t *nsteps, int nthreads);

/* ---- voxels ---- */
typedef struct orc_octree orc_octree;
orc_octree *orc_octree_new(const orc_grid *g);
orc_octree *orc_octree_copy(const orc_octree *t);
void orc_octree_free(orc_octree *t);
void orc_octree_clear(orc_octree *t);
uint64_t orc_octree_block(const orc_octree *t, int bx, int by, int bz);
void orc_octree_set_block(orc_octree *t, int bx, int by, int bz, uint64_t v);
uint64_t orc_octree_union_block(orc_octree *t, int bx, int by, int bz, uint64_t v);
int64_t orc_octree_nblocks(const orc_octree *t);
int64_t orc_octree_ncells(const orc_octree *t);
void orc_octree_add_line(orc_octree *t, const double *a, const double *b);
void orc_octree_add_piecewise_line(orc_octree *t, const double *pts, int npts);
void orc_octree_add_voxels(orc_octree *t, const orc_octree *other);
int orc_octree_collides(const orc_octree *a, const orc_octree *b);
/* visit_leaves order (TreeNode.hxx:177-190); returns count; arrays may be NULL */
int64_t orc_octree_export(const orc_octree *t, int64_t cap, uint8_t *bxyz, uint64_t *bits);
/* find_cell (VoxelOctree.cpp:309-317): returns 0, or 1 for domain_error */
int orc_find_cell(const orc_grid *g, const double *p, int64_t *cell);
void orc_octree_add_point(orc_octree *t, const double *p);
void orc_octree_add_sphere(orc_octree *t, const double *c, double r);
/* environment preparation: collision/VoxelOctree.cpp:533-689 (remove_interior: keep_diagonal != 0
 * is the 27-neighbour variant) and :693-952 (dilate) */
void orc_octree_dilate_6neighbor(orc_octree *t, int num);
void orc_octree_dilate_27neighbor(orc_octree *t, int num);
void orc_octree_dilate_sphere(orc_octree *t, double r);
void orc_octree_remove_interior(orc_octree *t, int keep_diagonal);
void orc_octree_add_capsule(orc_octree *t, const double *a, const double *b, double r);

/* OMPL-side restatement (Problem.cpp:101-163, VoxelBackboneMotionValidator.cpp:55-66) */
uint32_t orc_valid_segment_count(const orc_robot *rb, const orc_space *sp, const double *a,
                                 const double *b);
void orc_interpolate(const orc_robot *rb, const double *a, const double *b, double t,
                     double *out);

/* VoxelBackboneValidityChecker::voxelize_impl (VoxelBackboneValidityChecker.h:49-57) */
void orc_voxelize_shape(const orc_grid *g, const double *p, int npts, orc_octree *out);

typedef struct orc_edge_out {
  int32_t is_fully_valid;
  int32_t nsamples;          /* FK samples evaluated */
  int32_t out_of_domain;     /* find_cell would have thrown */
  int32_t _pad;
  double t;                  /* last valid t */
  double last_valid[ORC_MAX_TENDONS + 2];
} orc_edge_out;
/* VoxelBackboneMotionValidator::voxelize_impl -> voxelize_valid_backbone_motion
 * (VoxelEnvironment.cpp:207-444).  env==NULL -> "voxelize" (self-validity only);
 * env!=NULL -> "voxelize_until_invalid". */
void orc_voxelize_edge(const orc_robot *rb, const orc_grid *g, const orc_space *sp,
                       const double *a, const double *b, const orc_octree *env,
                       orc_octree *out, orc_edge_out *info);

/* ---- batch drivers over cached sets (VoxelCachedLazyPRM.cpp:1520-1542,1584-1591) ---- */
typedef struct orc_setstore orc_setstore;
orc_setstore *orc_setstore_new(const orc_grid *g, int64_t n);
void orc_setstore_free(orc_setstore *s);
int64_t orc_setstore_size(const orc_setstore *s);
orc_octree *orc_setstore_get(orc_setstore *s, int64_t i);
int64_t orc_setstore_total_blocks(const orc_setstore *s);
/* CSR export: offsets[n+1]; per block morton key (x-major octant order = visit_leaves order) */
void orc_setstore_export(const orc_setstore *s, uint64_t *offsets, uint32_t *keys,
                         uint64_t *bits);
void orc_voxelize_vertices_batch(const orc_robot *rb, const orc_grid *g, const double *states,
                                 int64_t n, orc_setstore *out, uint32_t *flags, int nthreads);
void orc_voxelize_edges_batch(const orc_robot *rb, const orc_grid *g, const orc_space *sp,
                              const double *a, const double *b, int64_t n, orc_setstore *out,
                              uint32_t *flags, double *t_last, int32_t *nsamples,
                              int nthreads);
/* computeVertexValidity / computeEdgeValidity with warm caches (VoxelCachedLazyPRM.cpp:2607-2631):
 * verdict[i] = 1 if set i collides with env */
void orc_check_sets_batch(const orc_setstore *s, const orc_octree *env, int64_t begin,
                          int64_t end, uint8_t *verdict, int nthreads);
uint32_t orc_morton_key(int bx, int by, int bz, int Nb);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
