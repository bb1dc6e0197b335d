// rmp_io.cpp -- reader/writer of the reference's binary roadmap format `.rmp`
// (LazyRmpParser / RmpStreamer, motion-planning/VoxelCachedLazyPRM.cpp:862-1114; block records
// as serialize_inner :636-657 writes them).  Host-side IO: it turns the per-item block lists
// {u8 bx, u8 by, u8 bz, u64 bits} into the CSR (offsets, Morton keys, bits) of the device set
// store and back, so reference-built roadmaps load straight into irt_setstore_import.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/irt_b200.h"

namespace {

// Counts in the file are untrusted: every read is checked against the bytes that are left, so a truncated
// or corrupt file ends in IRT_ERR_INVALID_ARGUMENT, never in an out-of-memory allocation or a crash.
struct Reader {
  std::FILE *f;
  uint64_t left = 0;  // bytes of the file not consumed yet
  bool ok = true;
  template <typename T>
  bool get(T *dst, size_t count = 1) {
    if (!ok) return false;
    if (!fits(count, sizeof(T))) return ok = false;
    if (count && std::fread(dst, sizeof(T), count, f) != count) return ok = false;
    left -= (uint64_t)count * sizeof(T);
    return true;
  }
  bool fits(uint64_t count, uint64_t size) const { return count <= left / (size ? size : 1); }
};

template <typename T>
T *dup(const std::vector<T> &v) {
  T *p = (T *)std::malloc((v.size() ? v.size() : 1) * sizeof(T));
  if (p && !v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

bool read_blocks(Reader &in, int Nb, std::vector<uint32_t> &keys, std::vector<uint64_t> &bits) {
  uint32_t nblocks = 0;
  if (!in.get(&nblocks)) return false;
  if (!in.fits(nblocks, 11)) return in.ok = false;
  std::vector<unsigned char> buf((size_t)nblocks * 11);
  if (!in.get(buf.data(), buf.size())) return false;
  for (uint32_t i = 0; i < nblocks; i++) {
    const unsigned char *r = &buf[(size_t)i * 11];
    uint64_t v;
    std::memcpy(&v, r + 3, 8);
    if (r[0] >= Nb || r[1] >= Nb || r[2] >= Nb) return in.ok = false;
    keys.push_back(irt_morton_key(r[0], r[1], r[2], Nb));
    bits.push_back(v);
  }
  return true;
}

bool write_blocks(std::FILE *f, int Nb, const uint32_t *keys, const uint64_t *bits, uint64_t lo, uint64_t hi) {
  const uint32_t nblocks = (uint32_t)(hi - lo);
  if (std::fwrite(&nblocks, 4, 1, f) != 1) return false;
  std::vector<unsigned char> buf((size_t)nblocks * 11);
  for (uint64_t j = lo; j < hi; j++) {
    int bx, by, bz;
    irt_morton_decode(keys[j], Nb, &bx, &by, &bz);
    unsigned char *r = &buf[(size_t)(j - lo) * 11];
    r[0] = (unsigned char)bx; r[1] = (unsigned char)by; r[2] = (unsigned char)bz;
    std::memcpy(r + 3, &bits[j], 8);
  }
  return buf.empty() || std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
}

}  // namespace

extern "C" {

void irt_rmp_free(irt_rmp *r) {
  if (!r) return;
  std::free(r->v_index); std::free(r->v_state); std::free(r->v_has_tip); std::free(r->v_tip);
  std::free(r->v_has_vox); std::free(r->v_off); std::free(r->v_keys); std::free(r->v_bits);
  std::free(r->e_src); std::free(r->e_dst); std::free(r->e_weight); std::free(r->e_has_vox);
  std::free(r->e_off); std::free(r->e_keys); std::free(r->e_bits);
  std::free(r);
}

static int rmp_read_impl(std::FILE *f, irt_rmp **out) {
  Reader in{f};
  if (std::fseek(f, 0, SEEK_END) != 0) return IRT_ERR_INVALID_ARGUMENT;
  const long size = std::ftell(f);
  if (size < 0 || std::fseek(f, 0, SEEK_SET) != 0) return IRT_ERR_INVALID_ARGUMENT;
  in.left = (uint64_t)size;
  irt_rmp *r = (irt_rmp *)std::calloc(1, sizeof(irt_rmp));
  if (!r) return IRT_ERR_CAPACITY;
  struct Guard {  // frees the partial result if a container throws on the way
    irt_rmp *r;
    ~Guard() { if (r) irt_rmp_free(r); }
  } guard{r};
  uint8_t has_vox = 0;
  in.get(&r->n_verts); in.get(&r->n_edges); in.get(&has_vox);
  r->has_voxels = has_vox ? 1 : 0;
  if (has_vox) {
    uint8_t nb = 0;
    in.get(&nb); in.get(r->lims, 6);
    r->Nb = nb;
    // VoxelOctree(Ng): Ng = 4 Nb must be a power of two in [4, 512] (collision/VoxelOctree.cpp:83-116)
    if (nb < 1 || nb > 128 || (nb & (nb - 1)) != 0) in.ok = false;
  }
  std::vector<uint32_t> vidx, vkeys, esrc, edst, ekeys;
  std::vector<double> vstate, vtip, ew;
  std::vector<uint8_t> vhastip, vhasvox, ehasvox;
  std::vector<uint64_t> voff{0}, vbits, eoff{0}, ebits;
  r->state_size = -1;
  bool ok = in.ok;
  for (uint32_t i = 0; ok && i < r->n_verts; i++) {
    uint32_t idx = 0, cnt = 0;
    ok = in.get(&idx) && in.get(&cnt);
    if (!ok) break;
    if (r->state_size < 0) r->state_size = (int32_t)cnt;
    if ((int32_t)cnt != r->state_size || !in.fits(cnt, sizeof(double))) { ok = false; break; }
    std::vector<double> st(cnt);
    uint8_t has_tip = 0;
    double tip[3] = {0, 0, 0};
    ok = in.get(st.data(), cnt) && in.get(&has_tip);
    if (ok && has_tip) ok = in.get(tip, 3);
    uint8_t hv = 0;
    if (ok && has_vox) {
      ok = in.get(&hv);
      if (ok && hv) ok = read_blocks(in, r->Nb, vkeys, vbits);
    }
    vidx.push_back(idx);
    vstate.insert(vstate.end(), st.begin(), st.end());
    vhastip.push_back(has_tip);
    vtip.insert(vtip.end(), tip, tip + 3);
    vhasvox.push_back(hv);
    voff.push_back(vkeys.size());
  }
  for (uint32_t i = 0; ok && i < r->n_edges; i++) {
    uint32_t s = 0, t = 0;
    double w = 0;
    ok = in.get(&s) && in.get(&t) && in.get(&w);
    uint8_t hv = 0;
    if (ok && has_vox) {
      ok = in.get(&hv);
      if (ok && hv) ok = read_blocks(in, r->Nb, ekeys, ebits);
    }
    esrc.push_back(s); edst.push_back(t); ew.push_back(w); ehasvox.push_back(hv);
    eoff.push_back(ekeys.size());
  }
  if (r->state_size < 0) r->state_size = 0;
  r->v_index = dup(vidx); r->v_state = dup(vstate); r->v_has_tip = dup(vhastip); r->v_tip = dup(vtip);
  r->v_has_vox = dup(vhasvox); r->v_off = dup(voff); r->v_keys = dup(vkeys); r->v_bits = dup(vbits);
  r->e_src = dup(esrc); r->e_dst = dup(edst); r->e_weight = dup(ew); r->e_has_vox = dup(ehasvox);
  r->e_off = dup(eoff); r->e_keys = dup(ekeys); r->e_bits = dup(ebits);
  if (!ok || !in.ok) return IRT_ERR_INVALID_ARGUMENT;  // truncated or malformed file (guard frees r)
  if (!r->v_index || !r->v_state || !r->v_has_tip || !r->v_tip || !r->v_has_vox || !r->v_off || !r->v_keys ||
      !r->v_bits || !r->e_src || !r->e_dst || !r->e_weight || !r->e_has_vox || !r->e_off || !r->e_keys ||
      !r->e_bits)
    return IRT_ERR_CAPACITY;
  guard.r = nullptr;
  *out = r;
  return IRT_OK;
}

int irt_rmp_read(const char *path, irt_rmp **out) {
  if (!path || !out) return IRT_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  std::FILE *f = std::fopen(path, "rb");
  if (!f) return IRT_ERR_INVALID_ARGUMENT;
  int rc;
  try {
    rc = rmp_read_impl(f, out);
  } catch (...) {  // std::bad_alloc of a container: the C ABI never throws
    rc = IRT_ERR_CAPACITY;
  }
  std::fclose(f);
  return rc;
}

int irt_rmp_write(const char *path, const irt_rmp *r) {
  if (!path || !r) return IRT_ERR_INVALID_ARGUMENT;
  if (r->has_voxels && (r->Nb < 1 || r->Nb > 128)) return IRT_ERR_INVALID_ARGUMENT;  // u8 Nb, Ng <= 512
  if (r->state_size < 0) return IRT_ERR_INVALID_ARGUMENT;
  if (r->n_verts && (!r->v_index || (r->state_size && !r->v_state))) return IRT_ERR_INVALID_ARGUMENT;
  if (r->n_edges && (!r->e_src || !r->e_dst || !r->e_weight)) return IRT_ERR_INVALID_ARGUMENT;
  for (uint32_t i = 0; i < r->n_verts; i++) {
    if (r->v_has_tip && r->v_has_tip[i] && !r->v_tip) return IRT_ERR_INVALID_ARGUMENT;
    if (r->has_voxels && r->v_has_vox && r->v_has_vox[i] &&
        (!r->v_off || !r->v_keys || !r->v_bits || r->v_off[i + 1] < r->v_off[i] ||
         r->v_off[i + 1] - r->v_off[i] > 0xffffffffull))
      return IRT_ERR_INVALID_ARGUMENT;
  }
  for (uint32_t i = 0; i < r->n_edges; i++)
    if (r->has_voxels && r->e_has_vox && r->e_has_vox[i] &&
        (!r->e_off || !r->e_keys || !r->e_bits || r->e_off[i + 1] < r->e_off[i] ||
         r->e_off[i + 1] - r->e_off[i] > 0xffffffffull))
      return IRT_ERR_INVALID_ARGUMENT;
  std::FILE *f = std::fopen(path, "wb");
  if (!f) return IRT_ERR_INVALID_ARGUMENT;
  bool ok = std::fwrite(&r->n_verts, 4, 1, f) == 1 && std::fwrite(&r->n_edges, 4, 1, f) == 1;
  const uint8_t has_vox = r->has_voxels ? 1 : 0;
  ok = ok && std::fwrite(&has_vox, 1, 1, f) == 1;
  if (ok && has_vox) {
    const uint8_t nb = (uint8_t)r->Nb;
    ok = std::fwrite(&nb, 1, 1, f) == 1 && std::fwrite(r->lims, 8, 6, f) == 6;
  }
  for (uint32_t i = 0; ok && i < r->n_verts; i++) {
    const uint32_t cnt = (uint32_t)r->state_size;
    ok = std::fwrite(&r->v_index[i], 4, 1, f) == 1 && std::fwrite(&cnt, 4, 1, f) == 1 &&
         (cnt == 0 || std::fwrite(r->v_state + (size_t)i * cnt, 8, cnt, f) == cnt);
    const uint8_t ht = r->v_has_tip ? r->v_has_tip[i] : 0;
    ok = ok && std::fwrite(&ht, 1, 1, f) == 1;
    if (ok && ht) ok = std::fwrite(r->v_tip + (size_t)i * 3, 8, 3, f) == 3;
    if (ok && has_vox) {
      const uint8_t hv = r->v_has_vox ? r->v_has_vox[i] : 0;
      ok = std::fwrite(&hv, 1, 1, f) == 1;
      if (ok && hv) ok = write_blocks(f, r->Nb, r->v_keys, r->v_bits, r->v_off[i], r->v_off[i + 1]);
    }
  }
  for (uint32_t i = 0; ok && i < r->n_edges; i++) {
    ok = std::fwrite(&r->e_src[i], 4, 1, f) == 1 && std::fwrite(&r->e_dst[i], 4, 1, f) == 1 &&
         std::fwrite(&r->e_weight[i], 8, 1, f) == 1;
    if (ok && has_vox) {
      const uint8_t hv = r->e_has_vox ? r->e_has_vox[i] : 0;
      ok = std::fwrite(&hv, 1, 1, f) == 1;
      if (ok && hv) ok = write_blocks(f, r->Nb, r->e_keys, r->e_bits, r->e_off[i], r->e_off[i + 1]);
    }
  }
  ok = (std::fclose(f) == 0) && ok;
  return ok ? IRT_OK : IRT_ERR_INVALID_ARGUMENT;
}

}  // extern "C"
