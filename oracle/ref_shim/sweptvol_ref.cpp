// The reference's swept-volume driver, compiled from its own text (test infrastructure only).
//
// motion-planning/VoxelEnvironment.cpp needs cpptoml and ITK, but the hot function does not:
//   VoxelEnvironment::voxelize_valid_backbone_motion (VoxelEnvironment.cpp:207-444) -- the LIFO
//   bisection with should_subdivide, first_invalid_t pruning and the final "t < first_invalid_t" union --
//   and rotate_point / rotate_points (:125-131).
// oracle/Makefile cuts those definitions (and the VoxelOctree core, as for libvoxeloctree_ref.so) out of
// the reference by anchors into oracle/_ref/gen/ (deleted after the build).  tendon/TendonResult.h is
// included as is.  What is hand-written here is only the DECLARATION of struct VoxelEnvironment, reduced
// to the members those definitions touch (VoxelEnvironment.h:36-143: the three std::function aliases,
// inv_rotation, rotate_point(s), PartialVoxelization, the method's signature) -- the reference's struct
// also declares file / ITK / toml members.  FK, validity and interpolation come in as callbacks, exactly
// as the reference passes them (VoxelBackboneMotionValidator.cpp:41-74).
#include <collision/Point.h>
#include <collision/collision_primitives.h>
#include <collision/detail/TreeNode.h>
#include <tendon/TendonResult.h>
#include <util/macros.h>

#include <algorithm>
#include <bitset>
#include <cmath>
#include <cstdint>
#include <deque>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <memory>
#include <queue>
#include <set>
#include <sstream>
#include <stack>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <variant>
#include <vector>

namespace collision {
#include "vo_class.inc"
}  // namespace collision
#include "vo_A.inc"  // opens namespace collision
#include "vo_B.inc"
#include "vo_C.inc"
}  // namespace collision

namespace E = Eigen;

namespace motion_planning {

struct VoxelEnvironment {  // declaration subset, see the header comment
  using InterpFunc = std::function<void(const std::vector<double>&, const std::vector<double>&,
                                        double, std::vector<double>&)>;
  using FkFunc    = std::function<tendon::TendonResult(const std::vector<double>)>;
  using ValidFunc = std::function<bool(const std::vector<double>&, const tendon::TendonResult&)>;
  Eigen::Matrix3d inv_rotation = Eigen::Matrix3d::Identity();
  Eigen::Vector3d rotate_point(const Eigen::Vector3d &point) const;
  void rotate_points(std::vector<Eigen::Vector3d> &points) const;
  struct PartialVoxelization {
    bool is_fully_valid;
    double t;
    std::vector<double> last_valid;
    std::vector<collision::Point> last_backbone;
    collision::VoxelOctree voxels {4};
  };
  PartialVoxelization voxelize_valid_backbone_motion(
      const collision::VoxelOctree &reference,
      const InterpFunc &interp,
      const FkFunc &fk,
      const ValidFunc &checker,
      const std::vector<double> &start,
      const std::vector<double> &end,
      double rel_threshold = 1e-5) const;
};

#include "ve_rot.inc"
#include "ve_motion.inc"

}  // namespace motion_planning

extern "C" {

typedef void (*veref_interp_cb)(const double *a, const double *b, int S, double t, double *out);
typedef int (*veref_fk_cb)(const double *state, int S, double *p, int cap);          // returns npts
typedef int (*veref_valid_cb)(const double *state, int S, const double *p, int npts);

// Returns 0, 1 for std::domain_error (a backbone point outside the grid in should_subdivide) or 2 for any
// other exception.  leaves: visit_leaves records {bx,by,bz,bits}; *n_leaves in: capacity, out: count.
int veref_voxelize_valid_backbone_motion(uint64_t Ng, const double *lim, const double *inv_rot /* row-major */,
                                         const double *a, const double *b, int S, double rel_threshold,
                                         int cap_pts, veref_interp_cb interp, veref_fk_cb fk,
                                         veref_valid_cb valid, int *is_fully_valid, double *t_last,
                                         double *last_valid, int *n_fk, uint64_t *leaves,
                                         uint64_t *n_leaves) {
  try {
    collision::VoxelOctree grid(Ng);
    grid.set_xlim(lim[0], lim[1]);
    grid.set_ylim(lim[2], lim[3]);
    grid.set_zlim(lim[4], lim[5]);
    motion_planning::VoxelEnvironment env;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) env.inv_rotation(r, c) = inv_rot[3 * r + c];
    int calls = 0;
    auto interp_f = [&](const std::vector<double> &s, const std::vector<double> &e, double t,
                        std::vector<double> &out) {
      out.resize(s.size());
      interp(s.data(), e.data(), (int)s.size(), t, out.data());
    };
    auto fk_f = [&](const std::vector<double> state) {
      tendon::TendonResult res;
      std::vector<double> p((size_t)cap_pts * 3);
      const int n = fk(state.data(), (int)state.size(), p.data(), cap_pts);
      for (int i = 0; i < n; i++) res.p.emplace_back(p.data() + 3 * i);
      calls++;
      return res;
    };
    auto valid_f = [&](const std::vector<double> &state, const tendon::TendonResult &shape) {
      std::vector<double> p(shape.p.size() * 3);
      for (size_t i = 0; i < shape.p.size(); i++)
        for (int k = 0; k < 3; k++) p[3 * i + k] = shape.p[i][k];
      return valid(state.data(), (int)state.size(), p.data(), (int)shape.p.size()) != 0;
    };
    std::vector<double> va(a, a + S), vb(b, b + S);
    auto ans = env.voxelize_valid_backbone_motion(grid, interp_f, fk_f, valid_f, va, vb, rel_threshold);
    *is_fully_valid = ans.is_fully_valid ? 1 : 0;
    *t_last = ans.t;
    for (int i = 0; i < S; i++) last_valid[i] = ans.last_valid[i];
    *n_fk = calls;
    uint64_t n = 0;
    const uint64_t cap = *n_leaves;
    ans.voxels.visit_leaves([&](size_t bx, size_t by, size_t bz, uint64_t bits) {
      if (n < cap) { leaves[4 * n] = bx; leaves[4 * n + 1] = by; leaves[4 * n + 2] = bz; leaves[4 * n + 3] = bits; }
      n++;
    });
    *n_leaves = n;
    return 0;
  } catch (const std::domain_error &) {
    return 1;
  } catch (...) {
    return 2;
  }
}

}  // extern "C"
