// C wrapper around the REFERENCE's own octree template (test infrastructure only).
//
// This file contains no octree logic: it instantiates collision::detail::TreeNode<N> from
// /root/reference/cpp/src/collision/detail/TreeNode.h (+ TreeNode.hxx), which needs nothing
// but the C++ standard library, and exposes it to ctypes so tests can pin oracle/ against the
// reference's real storage / set algebra / collides() / visit_leaves() order
// (reference call sites collision/VoxelOctree.cpp:46-53,973-978 and
// VoxelOctree.h:310-329, which dispatch over the same sizes with std::variant).
//
// Built by oracle/Makefile into oracle/_ref/libtreenode_ref.so, only when /root/reference exists.
#include <collision/detail/TreeNode.h>

#include <cstdint>
#include <cstddef>
#include <vector>

namespace {

struct AnyTree {
  virtual ~AnyTree() {}
  virtual size_t Ng() const = 0;
  virtual AnyTree *clone() const = 0;
  virtual size_t nblocks() const = 0;
  virtual int is_empty() const = 0;
  virtual uint64_t block(size_t, size_t, size_t) const = 0;
  virtual void set_block(size_t, size_t, size_t, uint64_t) = 0;
  virtual uint64_t union_block(size_t, size_t, size_t, uint64_t) = 0;
  virtual uint64_t intersect_block(size_t, size_t, size_t, uint64_t) = 0;
  virtual int union_tree(const AnyTree &) = 0;
  virtual int intersect_tree(const AnyTree &) = 0;
  virtual int remove_tree(const AnyTree &) = 0;
  virtual int collides(const AnyTree &) const = 0;
  virtual int equals(const AnyTree &) const = 0;
  virtual void leaves(std::vector<uint64_t> &out) const = 0;
};

template <size_t N> struct Tree final : AnyTree {
  collision::detail::TreeNode<N> t;
  size_t Ng() const override { return N; }
  AnyTree *clone() const override { return new Tree<N>(*this); }
  size_t nblocks() const override { return t.nblocks(); }
  int is_empty() const override { return t.is_empty(); }
  uint64_t block(size_t x, size_t y, size_t z) const override { return t.block(x, y, z); }
  void set_block(size_t x, size_t y, size_t z, uint64_t v) override { t.set_block(x, y, z, v); }
  uint64_t union_block(size_t x, size_t y, size_t z, uint64_t v) override {
    return t.union_block(x, y, z, v);
  }
  uint64_t intersect_block(size_t x, size_t y, size_t z, uint64_t v) override {
    return t.intersect_block(x, y, z, v);
  }
  const Tree<N> *same(const AnyTree &o) const {
    return o.Ng() == N ? static_cast<const Tree<N> *>(&o) : nullptr;
  }
  int union_tree(const AnyTree &o) override {
    auto p = same(o); if (!p) return -1; t.union_tree(p->t); return 0;
  }
  int intersect_tree(const AnyTree &o) override {
    auto p = same(o); if (!p) return -1; t.intersect_tree(p->t); return 0;
  }
  int remove_tree(const AnyTree &o) override {
    auto p = same(o); if (!p) return -1; t.remove_tree(p->t); return 0;
  }
  int collides(const AnyTree &o) const override {
    auto p = same(o); if (!p) return -1; return t.collides(p->t) ? 1 : 0;
  }
  int equals(const AnyTree &o) const override {
    auto p = same(o); if (!p) return -1; return (t == p->t) ? 1 : 0;
  }
  void leaves(std::vector<uint64_t> &out) const override {
    t.visit_leaves([&out](size_t bx, size_t by, size_t bz, uint64_t b) {
      out.push_back(bx); out.push_back(by); out.push_back(bz); out.push_back(b);
    });
  }
};

AnyTree *make(size_t Ng) {
  switch (Ng) {
    case 4: return new Tree<4>();
    case 8: return new Tree<8>();
    case 16: return new Tree<16>();
    case 32: return new Tree<32>();
    case 64: return new Tree<64>();
    case 128: return new Tree<128>();
    case 256: return new Tree<256>();
    case 512: return new Tree<512>();
  }
  return nullptr;
}

}  // namespace

extern "C" {

void *tnref_new(uint64_t Ng) { return make(Ng); }
void tnref_free(void *h) { delete static_cast<AnyTree *>(h); }
void *tnref_clone(const void *h) { return static_cast<const AnyTree *>(h)->clone(); }
uint64_t tnref_nblocks(const void *h) { return static_cast<const AnyTree *>(h)->nblocks(); }
int tnref_is_empty(const void *h) { return static_cast<const AnyTree *>(h)->is_empty(); }
uint64_t tnref_block(const void *h, uint64_t x, uint64_t y, uint64_t z) {
  return static_cast<const AnyTree *>(h)->block(x, y, z);
}
void tnref_set_block(void *h, uint64_t x, uint64_t y, uint64_t z, uint64_t v) {
  static_cast<AnyTree *>(h)->set_block(x, y, z, v);
}
uint64_t tnref_union_block(void *h, uint64_t x, uint64_t y, uint64_t z, uint64_t v) {
  return static_cast<AnyTree *>(h)->union_block(x, y, z, v);
}
uint64_t tnref_intersect_block(void *h, uint64_t x, uint64_t y, uint64_t z, uint64_t v) {
  return static_cast<AnyTree *>(h)->intersect_block(x, y, z, v);
}
int tnref_union_tree(void *h, const void *o) {
  return static_cast<AnyTree *>(h)->union_tree(*static_cast<const AnyTree *>(o));
}
int tnref_intersect_tree(void *h, const void *o) {
  return static_cast<AnyTree *>(h)->intersect_tree(*static_cast<const AnyTree *>(o));
}
int tnref_remove_tree(void *h, const void *o) {
  return static_cast<AnyTree *>(h)->remove_tree(*static_cast<const AnyTree *>(o));
}
int tnref_collides(const void *h, const void *o) {
  return static_cast<const AnyTree *>(h)->collides(*static_cast<const AnyTree *>(o));
}
int tnref_equals(const void *h, const void *o) {
  return static_cast<const AnyTree *>(h)->equals(*static_cast<const AnyTree *>(o));
}
// visit_leaves order; writes min(n, cap) records {bx,by,bz,bits} and returns n.
uint64_t tnref_leaves(const void *h, uint64_t *out, uint64_t cap) {
  std::vector<uint64_t> v;
  static_cast<const AnyTree *>(h)->leaves(v);
  uint64_t n = v.size() / 4;
  for (uint64_t i = 0; i < n && i < cap; i++)
    for (int k = 0; k < 4; k++) out[4 * i + k] = v[4 * i + k];
  return n;
}
// Batch form of the loop at VoxelCachedLazyPRM.cpp:1584-1591 for timing / verdict parity:
// sets are given as CSR {off[n+1], bx,by,bz (u8), bits}; each is rebuilt as a TreeNode and
// tested with collides() against env.  verdict[i] in {0,1}.
int tnref_check_csr(const void *env, uint64_t n, const uint64_t *off, const uint8_t *bx,
                    const uint8_t *by, const uint8_t *bz, const uint64_t *bits, uint8_t *verdict) {
  const AnyTree *e = static_cast<const AnyTree *>(env);
  for (uint64_t i = 0; i < n; i++) {
    AnyTree *s = make(e->Ng());
    if (!s) return -1;
    for (uint64_t k = off[i]; k < off[i + 1]; k++) s->set_block(bx[k], by[k], bz[k], bits[k]);
    verdict[i] = (uint8_t)e->collides(*s);
    delete s;
  }
  return 0;
}

// Timing form of the same loop: the sets are built once (untimed by the caller) and kept as trees,
// as the reference keeps them in its vertex/edge properties (VoxelCachedLazyPRM.h:141,165-179);
// tnref_sets_check then is the `#pragma omp parallel for` of VoxelCachedLazyPRM.cpp:1584-1591 with
// nothing but TreeNode::collides inside.
struct SetArray { std::vector<AnyTree *> sets; };

void *tnref_sets_build(uint64_t Ng, uint64_t n, const uint64_t *off, const uint8_t *bx,
                       const uint8_t *by, const uint8_t *bz, const uint64_t *bits) {
  SetArray *a = new SetArray();
  a->sets.resize(n, nullptr);
  for (uint64_t i = 0; i < n; i++) {
    AnyTree *s = make(Ng);
    if (!s) { delete a; return nullptr; }
    for (uint64_t k = off[i]; k < off[i + 1]; k++) s->set_block(bx[k], by[k], bz[k], bits[k]);
    a->sets[i] = s;
  }
  return a;
}
void tnref_sets_free(void *h) {
  SetArray *a = static_cast<SetArray *>(h);
  if (!a) return;
  for (AnyTree *s : a->sets) delete s;
  delete a;
}
int tnref_sets_check(const void *env, const void *h, uint8_t *verdict, int nthreads) {
  const AnyTree *e = static_cast<const AnyTree *>(env);
  const SetArray *a = static_cast<const SetArray *>(h);
  const int64_t n = (int64_t)a->sets.size();
  (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t i = 0; i < n; i++) verdict[i] = (uint8_t)e->collides(*a->sets[i]);
  return 0;
}

}  // extern "C"
