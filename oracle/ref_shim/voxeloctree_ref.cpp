// The reference's collision::VoxelOctree core, compiled from its own text (test infrastructure only).
//
// collision/VoxelOctree.cpp as a whole needs OMPL, ITK, FCL, cpptoml and nlohmann-json, but the part
// of it on the hot path does not.  oracle/Makefile cuts these line ranges out of the reference's
// files -- by function-name anchors, with awk, into oracle/_ref/gen/*.inc, nothing is copied into
// the repository -- and this file compiles them unmodified:
//   vo_class.inc  class VoxelOctree { ... } from collision/VoxelOctree.h:68-330, minus the member
//                 declarations that mention ITK / Mesh / Sphere / Capsule / json / toml / file IO
//   vo_A.inc      VoxelOctree.cpp from `#define my_assert` up to (not including) add_sphere:
//                 constructor, limits, block / cell accessors, nearest_cell / find_cell, add_point,
//                 **add_line**, add_piecewise_line                                (cpp:36-432)
//   vo_P.inc      add_sphere, add_capsule (cpp:434-515) with struct Sphere / Capsule and the inline collides()
//                 overloads of collision.hxx:55-108 (this library only; the sweptvol / rmp extractions drop them)
//   vo_B.inc      add_voxels ... visit_modify_voxels: remove_interior_6/27neighbor, dilate_6/27neighbor,
//                 dilate_sphere, collides, remove/intersect, the visitors         (cpp:517-1074)
//   vo_C.inc      bitmask, is_in_domain, domain_check                            (cpp:1499-1521)
// Point arithmetic goes through the Eigen stand-in (eigen_standin/Eigen/Core: pins the reference's
// expressions, not Eigen's rounding; add_line uses -, cwiseProduct, normalized, cwiseAbs only).
// The octree storage underneath is the reference's real TreeNode.h.
#include <collision/Point.h>
#include <collision/collision_primitives.h>
#include <collision/detail/TreeNode.h>
#include <util/macros.h>

#include <algorithm>
#include <bitset>
#include <cmath>
#include <cstdint>
#include <functional>
#include <iomanip>
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
#ifdef VOREF_WITH_PRIMITIVES
// struct Sphere / Capsule (truncated at their first cpptoml / fcl member) and the inline collides() overloads
// for Point / Sphere / Capsule (collision.hxx:55-108), as in selfcol_ref.cpp
#include "sc_sphere.inc"
#include "sc_capsule.inc"
#include "sc_collides.inc"
#endif
#include "vo_class.inc"
}  // namespace collision

#include "vo_A.inc"  // opens namespace collision { and leaves it open, like the file it comes from
#ifdef VOREF_WITH_PRIMITIVES
#include "vo_P.inc"  // add_sphere, add_capsule (cpp:434-515): Environment::voxelize's primitives
#endif
#include "vo_B.inc"
#include "vo_C.inc"
}  // namespace collision (opened inside vo_A.inc)

using collision::VoxelOctree;
using collision::Point;

extern "C" {

void *voref_new(uint64_t Ng, const double *lim) {
  try {
    VoxelOctree *t = new VoxelOctree(Ng);
    t->set_xlim(lim[0], lim[1]);
    t->set_ylim(lim[2], lim[3]);
    t->set_zlim(lim[4], lim[5]);
    return t;
  } catch (...) {
    return nullptr;
  }
}
void voref_free(void *h) { delete static_cast<VoxelOctree *>(h); }
void *voref_copy(const void *h) { return new VoxelOctree(*static_cast<const VoxelOctree *>(h)); }
uint64_t voref_nblocks(const void *h) { return static_cast<const VoxelOctree *>(h)->nblocks(); }
uint64_t voref_ncells(const void *h) { return static_cast<const VoxelOctree *>(h)->ncells(); }
uint64_t voref_block(const void *h, uint64_t x, uint64_t y, uint64_t z) {
  return static_cast<const VoxelOctree *>(h)->block(x, y, z);
}
void voref_set_block(void *h, uint64_t x, uint64_t y, uint64_t z, uint64_t v) {
  static_cast<VoxelOctree *>(h)->set_block(x, y, z, v);
}
uint64_t voref_union_block(void *h, uint64_t x, uint64_t y, uint64_t z, uint64_t v) {
  return static_cast<VoxelOctree *>(h)->union_block(x, y, z, v);
}
int voref_set_cell(void *h, uint64_t x, uint64_t y, uint64_t z) {
  return static_cast<VoxelOctree *>(h)->set_cell(x, y, z) ? 1 : 0;
}
void voref_add_line(void *h, const double *a, const double *b) {
  static_cast<VoxelOctree *>(h)->add_line(Point(a), Point(b));
}
void voref_add_piecewise_line(void *h, const double *pts, int n) {
  std::vector<Point> line;
  for (int i = 0; i < n; i++) line.emplace_back(pts + 3 * i);
  static_cast<VoxelOctree *>(h)->add_piecewise_line(line);
}
void voref_add_voxels(void *h, const void *o) {
  static_cast<VoxelOctree *>(h)->add_voxels(*static_cast<const VoxelOctree *>(o));
}
// 0 = ok, 1 = std::domain_error (point outside the grid), like the oracle's orc_find_cell
int voref_find_cell(const void *h, const double *p, int64_t *cell) {
  try {
    auto [x, y, z] = static_cast<const VoxelOctree *>(h)->find_cell(Point(p));
    cell[0] = (int64_t)x; cell[1] = (int64_t)y; cell[2] = (int64_t)z;
    return 0;
  } catch (const std::domain_error &) {
    return 1;
  }
}
void voref_nearest_cell(const void *h, const double *p, int64_t *cell) {
  auto [x, y, z] = static_cast<const VoxelOctree *>(h)->nearest_cell(Point(p));
  cell[0] = (int64_t)x; cell[1] = (int64_t)y; cell[2] = (int64_t)z;
}
int voref_collides(const void *h, const void *o) {
  try {
    return static_cast<const VoxelOctree *>(h)->collides(*static_cast<const VoxelOctree *>(o)) ? 1 : 0;
  } catch (const std::invalid_argument &) {
    return -1;  // voxel dimension mismatch (check_dims, VoxelOctree.cpp:46-53)
  }
}
void voref_dilate(void *h, int num, int use_diagonal) {
  static_cast<VoxelOctree *>(h)->dilate(num, use_diagonal != 0);
}
void voref_dilate_sphere(void *h, double r) { static_cast<VoxelOctree *>(h)->dilate_sphere(r); }
void voref_remove_interior(void *h, int keep_diagonal) {
  static_cast<VoxelOctree *>(h)->remove_interior(keep_diagonal != 0);
}
#ifdef VOREF_WITH_PRIMITIVES
void voref_add_point(void *h, const double *p) { static_cast<VoxelOctree *>(h)->add(Point(p)); }
void voref_add_sphere(void *h, const double *c, double r) {
  static_cast<VoxelOctree *>(h)->add(collision::Sphere{Point(c), r});
}
void voref_add_capsule(void *h, const double *a, const double *b, double r) {
  static_cast<VoxelOctree *>(h)->add(collision::Capsule{Point(a), Point(b), r});
}
#endif
uint64_t voref_bitmask(int x, int y, int z) { return VoxelOctree::bitmask(x, y, z); }
// visit_leaves order; writes min(n, cap) records {bx,by,bz,bits} and returns n
uint64_t voref_leaves(const void *h, uint64_t *out, uint64_t cap) {
  uint64_t n = 0;
  static_cast<const VoxelOctree *>(h)->visit_leaves(
      [&](size_t bx, size_t by, size_t bz, uint64_t b) {
        if (n < cap) { out[4 * n] = bx; out[4 * n + 1] = by; out[4 * n + 2] = bz; out[4 * n + 3] = b; }
        n++;
      });
  return n;
}

}  // extern "C"
