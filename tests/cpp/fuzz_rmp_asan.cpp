// Sanitizer fuzzing of the .rmp reader / writer (interactive-rate-tendons_b200/csrc/rmp_io.cpp is compiled into
// this executable with -fsanitize=address,undefined by tests/test_rmp_io.py): a synthetic roadmap is written,
// then thousands of mutated copies (byte flips, truncations, splices) are read back.  Every read must return
// a status and, when it succeeds, a structure that can be written again -- with no sanitizer report.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/irt_b200.h"

// the two host helpers of the library that rmp_io.cpp uses (they live in a CUDA file there)
extern "C" uint32_t irt_morton_key(int bx, int by, int bz, int Nb) {
  uint32_t key = 0;
  for (int l = 0; (1 << l) < Nb; l++)
    key |= (uint32_t((bx >> l) & 1) << (3 * l + 2)) | (uint32_t((by >> l) & 1) << (3 * l + 1)) | (uint32_t((bz >> l) & 1) << (3 * l));
  return key;
}
extern "C" void irt_morton_decode(uint32_t key, int Nb, int *bx, int *by, int *bz) {
  *bx = *by = *bz = 0;
  for (int l = 0; (1 << l) < Nb; l++) {
    *bx |= ((key >> (3 * l + 2)) & 1) << l;
    *by |= ((key >> (3 * l + 1)) & 1) << l;
    *bz |= ((key >> (3 * l)) & 1) << l;
  }
}

static std::vector<unsigned char> slurp(const std::string &path) {
  std::vector<unsigned char> b;
  if (FILE *f = std::fopen(path.c_str(), "rb")) {
    unsigned char buf[4096];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) b.insert(b.end(), buf, buf + n);
    std::fclose(f);
  }
  return b;
}
static void spit(const std::string &path, const std::vector<unsigned char> &b) {
  FILE *f = std::fopen(path.c_str(), "wb");
  if (!b.empty()) std::fwrite(b.data(), 1, b.size(), f);
  std::fclose(f);
}

int main(int argc, char **argv) {
  const std::string dir = argc > 1 ? argv[1] : "/tmp";
  const int iters = argc > 2 ? std::atoi(argv[2]) : 3000;
  const std::string good = dir + "/fuzz_good.rmp", bad = dir + "/fuzz_bad.rmp", again = dir + "/fuzz_again.rmp";
  std::mt19937_64 gen(20220801);
  // a small roadmap: 7 vertices (S = 5), 9 edges, 32^3 grid
  const uint32_t nv = 7, ne = 9;
  const int S = 5, Nb = 8;
  std::vector<uint32_t> vidx(nv), vkeys, esrc(ne), edst(ne), ekeys;
  std::vector<double> vstate(nv * S), vtip(nv * 3), ew(ne);
  std::vector<uint8_t> vht(nv), vhv(nv), ehv(ne);
  std::vector<uint64_t> voff{0}, vbits, eoff{0}, ebits;
  auto fill = [&](std::vector<uint32_t> &keys, std::vector<uint64_t> &bits, std::vector<uint64_t> &off, uint8_t has) {
    if (has) {
      uint32_t k = gen() % 40;
      for (int j = 0, n = 1 + (int)(gen() % 6); j < n; j++) { keys.push_back(k); bits.push_back(gen() | 1); k += 1 + gen() % 50; }
    }
    off.push_back(keys.size());
  };
  for (uint32_t i = 0; i < nv; i++) {
    vidx[i] = i;
    for (int k = 0; k < S; k++) vstate[i * S + k] = (double)(gen() % 1000) / 50.0;
    vht[i] = i % 3 != 0;
    for (int k = 0; k < 3; k++) vtip[3 * i + k] = (double)(gen() % 1000) / 5000.0;
    vhv[i] = i % 4 != 1;
    fill(vkeys, vbits, voff, vhv[i]);
  }
  for (uint32_t i = 0; i < ne; i++) {
    esrc[i] = gen() % nv; edst[i] = gen() % nv; ew[i] = (double)(gen() % 1000) / 10.0;
    ehv[i] = i % 5 != 2;
    fill(ekeys, ebits, eoff, ehv[i]);
  }
  irt_rmp r;
  std::memset(&r, 0, sizeof(r));
  r.n_verts = nv; r.n_edges = ne; r.has_voxels = 1; r.Nb = Nb; r.state_size = S;
  for (int k = 0; k < 6; k++) r.lims[k] = (k % 2) ? 0.21 : -0.21;
  r.v_index = vidx.data(); r.v_state = vstate.data(); r.v_has_tip = vht.data(); r.v_tip = vtip.data();
  r.v_has_vox = vhv.data(); r.v_off = voff.data(); r.v_keys = vkeys.data(); r.v_bits = vbits.data();
  r.e_src = esrc.data(); r.e_dst = edst.data(); r.e_weight = ew.data(); r.e_has_vox = ehv.data();
  r.e_off = eoff.data(); r.e_keys = ekeys.data(); r.e_bits = ebits.data();
  if (irt_rmp_write(good.c_str(), &r) != IRT_OK) { std::printf("cannot write the seed file\n"); return 1; }
  const std::vector<unsigned char> seed = slurp(good);
  irt_rmp *back = nullptr;
  if (irt_rmp_read(good.c_str(), &back) != IRT_OK || back->n_verts != nv || back->e_off[ne] != eoff[ne]) {
    std::printf("seed file does not read back\n");
    return 1;
  }
  irt_rmp_free(back);

  int parsed = 0, rejected = 0, rewritten = 0;
  for (int it = 0; it < iters; it++) {
    std::vector<unsigned char> b = seed;
    switch (it % 4) {
      case 0:  // byte flips, anywhere
        for (int k = 0, n = 1 + (int)(gen() % 4); k < n; k++) b[gen() % b.size()] = (unsigned char)gen();
        break;
      case 1:  // byte flips in the header and the first records
        for (int k = 0, n = 1 + (int)(gen() % 3); k < n; k++) b[gen() % 120] = (unsigned char)gen();
        break;
      case 2:  // truncation (+ a flip)
        b.resize(gen() % b.size());
        if (!b.empty() && gen() % 2) b[gen() % b.size()] = (unsigned char)gen();
        break;
      default: {  // splice: a run of bytes copied over another place, or garbage appended
        const size_t n = 1 + gen() % 24, from = gen() % (b.size() - n), to = gen() % (b.size() - n);
        std::memmove(&b[to], &b[from], n);
        if (gen() % 3 == 0) b.insert(b.end(), 16, (unsigned char)gen());
      }
    }
    spit(bad, b);
    irt_rmp *m = nullptr;
    const int rc = irt_rmp_read(bad.c_str(), &m);
    if (rc != IRT_OK) {
      if (m != nullptr) { std::printf("error status with a non-null result\n"); return 1; }
      rejected++;
      continue;
    }
    parsed++;
    // whatever was accepted is a consistent structure: it can be written again and read back identically
    if (irt_rmp_write(again.c_str(), m) == IRT_OK) {
      irt_rmp *m2 = nullptr;
      if (irt_rmp_read(again.c_str(), &m2) != IRT_OK || m2->n_verts != m->n_verts || m2->n_edges != m->n_edges ||
          m2->v_off[m2->n_verts] != m->v_off[m->n_verts] || m2->e_off[m2->n_edges] != m->e_off[m->n_edges]) {
        std::printf("accepted file does not survive a write / read round trip (iteration %d)\n", it);
        return 1;
      }
      irt_rmp_free(m2);
      rewritten++;
    }
    irt_rmp_free(m);
  }
  std::printf("rmp fuzz ok: %d mutated files, %d parsed (%d rewritten), %d rejected\n", iters, parsed, rewritten, rejected);
  return parsed > 0 && rejected > 0 ? 0 : 1;
}
