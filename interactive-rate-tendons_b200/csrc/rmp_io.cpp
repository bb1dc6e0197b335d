// rmp_io.cpp -- reader/writer of the reference's binary roadmap format `.rmp`
// (LazyRmpParser / RmpStreamer, motion-planning/VoxelCachedLazyPRM.cpp:862-1114; block records
// as serialize_inner :636-657 writes them).  Host-side IO: it turns the per-item block lists
// {u8 bx, u8 by, u8 bz, u64 bits} into the CSR (offsets, Morton keys, bits) of the device set
// store and back, so reference-built roadmaps load straight into irt_setstore_import.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/irt_b200.h"

namespace {

// Counts in the file are untrusted: every read is checked against the bytes that are left, so a truncated
// or corrupt file ends in IRT_ERR_INVALID_ARGUMENT, never in an out-of-memory allocation or a crash.
// The file is parsed from memory (mmap) in two passes: the first walks the records to validate them and to
// size the result exactly, the second fills arrays allocated once (a roadmap file is gigabytes of 11-byte
// block records; per-field stdio calls and growing containers cost more than the parse).
struct Cursor {
  const unsigned char *cur, *end;
  bool ok = true;
  uint64_t left() const { return (uint64_t)(end - cur); }
  bool fits(uint64_t count, uint64_t size) const { return count <= left() / (size ? size : 1); }
  template <typename T>
  bool get(T *dst, size_t count = 1) {
    if (!ok) return false;
    if (!fits(count, sizeof(T))) return ok = false;
    std::memcpy(dst, cur, count * sizeof(T));
    cur += count * sizeof(T);
    return true;
  }
  bool skip(uint64_t bytes) {
    if (!ok) return false;
    if (bytes > left()) return ok = false;
    cur += bytes;
    return true;
  }
};

template <typename T>
T *alloc(size_t n) { return (T *)std::malloc((n ? n : 1) * sizeof(T)); }

// Morton key <-> block coordinates through small tables (a roadmap file holds hundreds of millions of block
// records; irt_morton_key / irt_morton_decode loop over the tree levels).  spread[a][c] = the bits of coordinate
// c on axis a at their key positions; the decode tables take the key in two pieces.
struct MortonTables {
  uint32_t spread[3][256];
  uint8_t lo[2048][4], hi[1024][4];  // key bits 0..10 / 11..20 -> the coordinate bits they carry (x, y, z, pad)
  explicit MortonTables(int Nb) {
    for (int c = 0; c < 256; c++) {
      spread[0][c] = c < Nb ? irt_morton_key(c, 0, 0, Nb) : 0;
      spread[1][c] = c < Nb ? irt_morton_key(0, c, 0, Nb) : 0;
      spread[2][c] = c < Nb ? irt_morton_key(0, 0, c, Nb) : 0;
    }
    auto fill = [](uint8_t (*t)[4], int n, int shift) {
      for (int v = 0; v < n; v++) {
        int b[3];
        irt_morton_decode((uint32_t)v << shift, 128, &b[0], &b[1], &b[2]);  // the bit layout does not depend on Nb
        t[v][0] = (uint8_t)b[0]; t[v][1] = (uint8_t)b[1]; t[v][2] = (uint8_t)b[2]; t[v][3] = 0;
      }
    };
    fill(lo, 2048, 0);
    fill(hi, 1024, 11);
  }
  uint32_t key(unsigned bx, unsigned by, unsigned bz) const { return spread[0][bx] | spread[1][by] | spread[2][bz]; }
  void decode(uint32_t k, unsigned char *xyz) const {  // k < Nb^3 <= 2^21
    const uint8_t *a = lo[k & 2047u], *b = hi[(k >> 11) & 1023u];
    xyz[0] = a[0] | b[0]; xyz[1] = a[1] | b[1]; xyz[2] = a[2] | b[2];
  }
};

// one block list: u32 count, then count x {u8 bx, u8 by, u8 bz, u64 bits}.  keys == nullptr: validate and
// count only (first pass).
bool read_blocks(Cursor &in, int Nb, const MortonTables &mt, uint32_t *keys, uint64_t *bits, uint64_t *count) {
  uint32_t nblocks = 0;
  if (!in.get(&nblocks)) return false;
  if (!in.fits(nblocks, 11)) return in.ok = false;
  const unsigned char *r = in.cur;
  if (!keys) {
    for (uint32_t i = 0; i < nblocks; i++, r += 11)
      if (r[0] >= Nb || r[1] >= Nb || r[2] >= Nb) return in.ok = false;
  } else {
    for (uint32_t i = 0; i < nblocks; i++, r += 11) {
      keys[i] = mt.key(r[0], r[1], r[2]);
      std::memcpy(&bits[i], r + 3, 8);
    }
  }
  in.cur = r;
  *count = nblocks;
  return true;
}

// output goes through one growing buffer that is flushed in large writes
struct Sink {
  std::FILE *f = nullptr;
  std::vector<unsigned char> buf;
  bool ok = true;
  unsigned char *grow(size_t n) {
    if (buf.size() + n > (4u << 20)) flush();
    const size_t at = buf.size();
    buf.resize(at + n);
    return buf.data() + at;
  }
  template <typename T>
  void put(const T *src, size_t count = 1) { std::memcpy(grow(count * sizeof(T)), src, count * sizeof(T)); }
  void flush() {
    if (ok && !buf.empty() && std::fwrite(buf.data(), 1, buf.size(), f) != buf.size()) ok = false;
    buf.clear();
  }
};

bool write_blocks(Sink &out, const MortonTables &mt, uint32_t nkeys, const uint32_t *keys, const uint64_t *bits,
                  uint64_t lo, uint64_t hi) {
  const uint32_t nblocks = (uint32_t)(hi - lo);
  out.put(&nblocks);
  for (uint64_t j = lo; j < hi;) {   // in runs, so that one grow() serves many records
    const uint64_t m = (hi - j < 4096) ? hi - j : 4096;
    unsigned char *r = out.grow((size_t)m * 11);
    for (uint64_t k = 0; k < m; k++, r += 11) {
      if (keys[j + k] >= nkeys) return false;  // a key outside the grid has no {bx, by, bz}
      mt.decode(keys[j + k], r);
      std::memcpy(r + 3, &bits[j + k], 8);
    }
    j += m;
  }
  return out.ok;
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

// pass == 0: validate + count into r->n_*/state_size and the totals; pass == 1: fill the allocated arrays
static bool rmp_walk(Cursor in, irt_rmp *r, const MortonTables &mt, bool fill, uint64_t *v_blocks, uint64_t *e_blocks) {
  const bool has_vox = r->has_voxels != 0;
  uint64_t nvb = 0, neb = 0;
  int32_t S = fill ? r->state_size : -1;
  for (uint32_t i = 0; i < r->n_verts; i++) {
    uint32_t idx = 0, cnt = 0;
    if (!in.get(&idx) || !in.get(&cnt)) return false;
    if (S < 0) {
      if (cnt > 0x7fffffffu) return false;
      S = (int32_t)cnt;
    }
    if ((int64_t)cnt != (int64_t)S || !in.fits(cnt, sizeof(double))) return false;  // one state size per file
    uint8_t has_tip = 0, hv = 0;
    double tip[3] = {0, 0, 0};
    if (fill) {
      r->v_index[i] = idx;
      in.get(r->v_state + (size_t)i * S, cnt);
    } else {
      in.skip((uint64_t)cnt * sizeof(double));
    }
    if (!in.get(&has_tip)) return false;
    if (has_tip && !in.get(tip, 3)) return false;
    uint64_t nb = 0;
    if (has_vox) {
      if (!in.get(&hv)) return false;
      if (hv && !read_blocks(in, r->Nb, mt, fill ? r->v_keys + nvb : nullptr, fill ? r->v_bits + nvb : nullptr, &nb))
        return false;
    }
    nvb += nb;
    if (fill) {
      r->v_has_tip[i] = has_tip;
      std::memcpy(r->v_tip + (size_t)i * 3, tip, sizeof(tip));
      r->v_has_vox[i] = hv;
      r->v_off[i + 1] = nvb;
    }
  }
  for (uint32_t i = 0; i < r->n_edges; i++) {
    uint32_t s = 0, t = 0;
    double w = 0;
    uint8_t hv = 0;
    if (!in.get(&s) || !in.get(&t) || !in.get(&w)) return false;
    uint64_t nb = 0;
    if (has_vox) {
      if (!in.get(&hv)) return false;
      if (hv && !read_blocks(in, r->Nb, mt, fill ? r->e_keys + neb : nullptr, fill ? r->e_bits + neb : nullptr, &nb))
        return false;
    }
    neb += nb;
    if (fill) {
      r->e_src[i] = s; r->e_dst[i] = t; r->e_weight[i] = w; r->e_has_vox[i] = hv;
      r->e_off[i + 1] = neb;
    }
  }
  if (!fill) r->state_size = S < 0 ? 0 : S;
  *v_blocks = nvb;
  *e_blocks = neb;
  return in.ok;
}

static int rmp_parse(const unsigned char *data, uint64_t size, irt_rmp **out) {
  Cursor in;
  in.cur = data;
  in.end = data + size;
  irt_rmp *r = (irt_rmp *)std::calloc(1, sizeof(irt_rmp));
  if (!r) return IRT_ERR_CAPACITY;
  struct Guard {  // frees the partial result on every early return
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
  if (!in.ok) return IRT_ERR_INVALID_ARGUMENT;
  // a record is at least 9 (vertex) / 16 (edge) bytes: counts beyond what the file can hold are corrupt
  if (!in.fits(r->n_verts, 9) || !in.fits(r->n_edges, 16)) return IRT_ERR_INVALID_ARGUMENT;
  const MortonTables mt(has_vox ? r->Nb : 1);
  uint64_t nvb = 0, neb = 0;
  if (!rmp_walk(in, r, mt, false, &nvb, &neb)) return IRT_ERR_INVALID_ARGUMENT;  // truncated or malformed file
  const size_t nv = r->n_verts, ne = r->n_edges, S = (size_t)r->state_size;
  r->v_index = alloc<uint32_t>(nv); r->v_state = alloc<double>(nv * S); r->v_has_tip = alloc<uint8_t>(nv);
  r->v_tip = alloc<double>(nv * 3); r->v_has_vox = alloc<uint8_t>(nv); r->v_off = alloc<uint64_t>(nv + 1);
  r->v_keys = alloc<uint32_t>(nvb); r->v_bits = alloc<uint64_t>(nvb);
  r->e_src = alloc<uint32_t>(ne); r->e_dst = alloc<uint32_t>(ne); r->e_weight = alloc<double>(ne);
  r->e_has_vox = alloc<uint8_t>(ne); r->e_off = alloc<uint64_t>(ne + 1);
  r->e_keys = alloc<uint32_t>(neb); r->e_bits = alloc<uint64_t>(neb);
  if (!r->v_index || !r->v_state || !r->v_has_tip || !r->v_tip || !r->v_has_vox || !r->v_off || !r->v_keys ||
      !r->v_bits || !r->e_src || !r->e_dst || !r->e_weight || !r->e_has_vox || !r->e_off || !r->e_keys ||
      !r->e_bits)
    return IRT_ERR_CAPACITY;
  r->v_off[0] = r->e_off[0] = 0;
  if (!rmp_walk(in, r, mt, true, &nvb, &neb)) return IRT_ERR_INVALID_ARGUMENT;
  guard.r = nullptr;
  *out = r;
  return IRT_OK;
}

int irt_rmp_read(const char *path, irt_rmp **out) {
  if (!path || !out) return IRT_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  const int fd = ::open(path, O_RDONLY);
  if (fd < 0) return IRT_ERR_INVALID_ARGUMENT;
  struct stat sb;
  if (::fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { ::close(fd); return IRT_ERR_INVALID_ARGUMENT; }
  const uint64_t size = (uint64_t)sb.st_size;
  void *map = size ? ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
  ::close(fd);
  if (size && map == MAP_FAILED) return IRT_ERR_CAPACITY;
  if (size) ::madvise(map, size, MADV_SEQUENTIAL);
  int rc;
  try {
    rc = rmp_parse((const unsigned char *)map, size, out);
  } catch (...) {  // the C ABI never throws
    rc = IRT_ERR_CAPACITY;
  }
  if (size) ::munmap(map, size);
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
  const MortonTables mt(r->has_voxels ? r->Nb : 1);
  const uint32_t nkeys = r->has_voxels ? (uint32_t)r->Nb * (uint32_t)r->Nb * (uint32_t)r->Nb : 1u;
  Sink out;
  out.f = f;
  bool ok = true;
  try {
    out.put(&r->n_verts); out.put(&r->n_edges);
    const uint8_t has_vox = r->has_voxels ? 1 : 0;
    out.put(&has_vox);
    if (has_vox) {
      const uint8_t nb = (uint8_t)r->Nb;
      out.put(&nb); out.put(r->lims, 6);
    }
    const uint32_t cnt = (uint32_t)r->state_size;
    for (uint32_t i = 0; ok && i < r->n_verts; i++) {
      out.put(&r->v_index[i]); out.put(&cnt);
      if (cnt) out.put(r->v_state + (size_t)i * cnt, cnt);
      const uint8_t ht = r->v_has_tip ? r->v_has_tip[i] : 0;
      out.put(&ht);
      if (ht) out.put(r->v_tip + (size_t)i * 3, 3);
      if (has_vox) {
        const uint8_t hv = r->v_has_vox ? r->v_has_vox[i] : 0;
        out.put(&hv);
        if (hv) ok = write_blocks(out, mt, nkeys, r->v_keys, r->v_bits, r->v_off[i], r->v_off[i + 1]);
      }
    }
    for (uint32_t i = 0; ok && i < r->n_edges; i++) {
      out.put(&r->e_src[i]); out.put(&r->e_dst[i]); out.put(&r->e_weight[i]);
      if (has_vox) {
        const uint8_t hv = r->e_has_vox ? r->e_has_vox[i] : 0;
        out.put(&hv);
        if (hv) ok = write_blocks(out, mt, nkeys, r->e_keys, r->e_bits, r->e_off[i], r->e_off[i + 1]);
      }
    }
    out.flush();
    ok = ok && out.ok;
  } catch (...) {  // std::bad_alloc of the buffer: the C ABI never throws
    ok = false;
  }
  ok = (std::fclose(f) == 0) && ok;
  return ok ? IRT_OK : IRT_ERR_INVALID_ARGUMENT;
}

}  // extern "C"
