// The reference's binary roadmap format (.rmp), compiled from its own text (test infrastructure only).
//
// motion-planning/VoxelCachedLazyPRM.cpp needs OMPL and Boost.Graph, but its .rmp reader and writer do not:
// serialize_inner, binary_write / binary_read, BinaryIFStream / BinaryISStream / BinaryOFStream,
// LazyRmpParser and RmpStreamer (VoxelCachedLazyPRM.cpp:635-1114) are cut out by anchors at build time
// (oracle/Makefile -> _ref/gen/rmp_core.inc, deleted after the build) and compiled unmodified;
// motion-planning/io/RoadmapParser.h and RoadmapWriter.h are included as they are; VoxelOctree is the
// extraction of libvoxeloctree_ref.so.  Hand-written here, because the originals live in OMPL or in the
// OMPL-derived planner class: the three OMPL log macros (no-ops), ompl::Exception (a runtime_error) and the
// two type aliases VoxelCachedLazyPRM::VoxelPtr / TipPosition (VoxelCachedLazyPRM.h:141, :151).
#include <collision/Point.h>
#include <collision/collision_primitives.h>
#include <collision/detail/TreeNode.h>
#include <util/macros.h>

#include <algorithm>
#include <array>
#include <bitset>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iomanip>
#include <limits>
#include <memory>
#include <optional>
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

#include <motion-planning/io/RoadmapParser.h>
#include <motion-planning/io/RoadmapWriter.h>

#define OMPL_DEBUG(...) ((void)0)
#define OMPL_INFORM(...) ((void)0)
#define OMPL_WARN(...) ((void)0)
namespace ompl {
struct Exception : std::runtime_error {
  explicit Exception(const std::string &what) : std::runtime_error(what) {}
  Exception(const std::string &prefix, const std::string &what) : std::runtime_error(prefix + ": " + what) {}
};
}  // namespace ompl

using motion_planning::io::ParsedType;
using motion_planning::io::ParsedVertex;
using motion_planning::io::ParsedEdge;
using motion_planning::io::RoadmapParser;
using motion_planning::io::RoadmapWriter;

namespace motion_planning {
struct VoxelCachedLazyPRM {  // the two aliases the streamer's signatures use
  using VoxelPtr = std::shared_ptr<collision::VoxelOctree>;
  using TipPosition = std::optional<Eigen::Vector3d>;
};
namespace {
#include "rmp_core.inc"
}  // namespace
}  // namespace motion_planning

namespace mp = motion_planning;

namespace {
struct Writer {
  mp::BinaryOFStream out;
  std::unique_ptr<mp::RmpStreamer<mp::BinaryOFStream>> s;
  Writer(const char *path, uint32_t nv, uint32_t ne) : out(path) {
    s.reset(new mp::RmpStreamer<mp::BinaryOFStream>(out, nv, ne));
  }
};
struct Reader {
  mp::BinaryIFStream in;
  std::unique_ptr<mp::LazyRmpParser<mp::BinaryIFStream>> p;
  explicit Reader(const char *path) : in(path) { p.reset(new mp::LazyRmpParser<mp::BinaryIFStream>(in)); }
};
std::shared_ptr<collision::VoxelOctree> make_vox(uint64_t Ng, const double *lim, uint64_t nb, const uint8_t *bxyz,
                                                 const uint64_t *bits) {
  auto v = std::make_shared<collision::VoxelOctree>(Ng);
  v->set_xlim(lim[0], lim[1]); v->set_ylim(lim[2], lim[3]); v->set_zlim(lim[4], lim[5]);
  for (uint64_t i = 0; i < nb; i++) v->set_block(bxyz[3 * i], bxyz[3 * i + 1], bxyz[3 * i + 2], bits[i]);
  return v;
}
}  // namespace

extern "C" {

void *rmpref_writer_open(const char *path, uint32_t nv, uint32_t ne) {
  try { return new Writer(path, nv, ne); } catch (...) { return nullptr; }
}
int rmpref_write_reference(void *h, uint64_t Ng, const double *lim) {
  try {
    collision::VoxelOctree ref(Ng);
    ref.set_xlim(lim[0], lim[1]); ref.set_ylim(lim[2], lim[3]); ref.set_zlim(lim[4], lim[5]);
    static_cast<Writer *>(h)->s->write_reference_voxels(ref);
    return 0;
  } catch (...) { return 1; }
}
int rmpref_write_vertex(void *h, uint32_t index, const double *state, int S, int has_tip, const double *tip,
                        int has_vox, uint64_t Ng, const double *lim, uint64_t nb, const uint8_t *bxyz,
                        const uint64_t *bits) {
  try {
    std::optional<Eigen::Vector3d> t;
    if (has_tip) t = Eigen::Vector3d(tip[0], tip[1], tip[2]);
    static_cast<Writer *>(h)->s->write_vertex(index, std::vector<double>(state, state + S), t,
                                             has_vox ? make_vox(Ng, lim, nb, bxyz, bits) : nullptr);
    return 0;
  } catch (...) { return 1; }
}
int rmpref_write_edge(void *h, uint32_t src, uint32_t dst, double w, int has_vox, uint64_t Ng, const double *lim,
                      uint64_t nb, const uint8_t *bxyz, const uint64_t *bits) {
  try {
    static_cast<Writer *>(h)->s->write_edge(src, dst, w, has_vox ? make_vox(Ng, lim, nb, bxyz, bits) : nullptr);
    return 0;
  } catch (...) { return 1; }
}
void rmpref_writer_close(void *h) {
  Writer *w = static_cast<Writer *>(h);
  w->s.reset();   // the streamer's destructor finishes the header of an empty roadmap
  delete w;       // (BinaryOFStream never closes its FILE; flush through a fresh handle below)
  fflush(nullptr);
}

void *rmpref_reader_open(const char *path) {
  try { return new Reader(path); } catch (...) { return nullptr; }
}
void rmpref_reader_close(void *h) { delete static_cast<Reader *>(h); }
// returns ParsedType (1 vertex, 2 edge, 3 done, -1 error); fills hdr = {index|source, target, state size,
// has_tip, has_voxels, n_blocks}, vals = {weight, tip xyz}
int rmpref_next(void *h, uint32_t *hdr, double *vals, double *state, int cap_state, uint64_t *leaves,
                uint64_t cap_leaves) {
  try {
    auto &p = *static_cast<Reader *>(h)->p;
    ParsedType t = p.next();
    if (t == ParsedType::DONE) return 3;
    p.populate_voxels();
    std::shared_ptr<collision::VoxelOctree> vox;
    if (t == ParsedType::VERTEX) {
      auto v = p.current_vertex();
      hdr[0] = v.index; hdr[1] = 0; hdr[2] = (uint32_t)v.state.size(); hdr[3] = v.tip_pos ? 1 : 0;
      if ((int)v.state.size() > cap_state) return -1;
      for (size_t i = 0; i < v.state.size(); i++) state[i] = v.state[i];
      vals[0] = 0;
      for (int k = 0; k < 3; k++) vals[1 + k] = v.tip_pos ? (*v.tip_pos)[k] : 0.0;
      vox = v.voxels;
    } else {
      auto e = p.current_edge();
      hdr[0] = e.source; hdr[1] = e.target; hdr[2] = 0; hdr[3] = 0;
      vals[0] = e.weight;
      vox = e.voxels;
    }
    hdr[4] = vox ? 1 : 0;
    uint64_t n = 0;
    if (vox)
      vox->visit_leaves([&](size_t bx, size_t by, size_t bz, uint64_t b) {
        if (n < cap_leaves) { leaves[4 * n] = bx; leaves[4 * n + 1] = by; leaves[4 * n + 2] = bz; leaves[4 * n + 3] = b; }
        n++;
      });
    hdr[5] = (uint32_t)n;
    return t == ParsedType::VERTEX ? 1 : 2;
  } catch (...) { return -1; }
}

}  // extern "C"
