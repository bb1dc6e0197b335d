"""`.rmp` roadmap files: the native reader/writer against a byte-level restatement of
the reference's RmpStreamer layout written here with struct.pack
(motion-planning/VoxelCachedLazyPRM.cpp:636-657, 862-1114)."""
import struct

import numpy as np
import pytest


def _pack_blocks(wl, keys, bits, Nb):
    bx, by, bz = wl.morton_decode(np.asarray(keys, dtype=np.uint32), Nb)
    out = struct.pack("<I", len(keys))
    for x, y, z, b in zip(bx.tolist(), by.tolist(), bz.tolist(), np.asarray(bits).tolist()):
        out += struct.pack("<BBBQ", x, y, z, b)
    return out


def _reference_style_bytes(wl, d):
    """what RmpStreamer writes: u32 nV, u32 nE, bool has_voxels, [u8 Nb, 6 f64], items..."""
    Nb = d["Ng"] // 4
    out = struct.pack("<II?", d["n_verts"], d["n_edges"], d["has_voxels"])
    if d["has_voxels"]:
        out += struct.pack("<B6d", Nb, *d["lims"])
    for i in range(d["n_verts"]):
        st = d["v_state"][i]
        out += struct.pack("<II", int(d["v_index"][i]), len(st)) + struct.pack("<%dd" % len(st), *st)
        out += struct.pack("<?", bool(d["v_has_tip"][i]))
        if d["v_has_tip"][i]:
            out += struct.pack("<3d", *d["v_tip"][i])
        if d["has_voxels"]:
            out += struct.pack("<?", bool(d["v_has_vox"][i]))
            if d["v_has_vox"][i]:
                lo, hi = int(d["v_off"][i]), int(d["v_off"][i + 1])
                out += _pack_blocks(wl, d["v_keys"][lo:hi], d["v_bits"][lo:hi], Nb)
    for i in range(d["n_edges"]):
        out += struct.pack("<IId", int(d["e_src"][i]), int(d["e_dst"][i]), float(d["e_weight"][i]))
        if d["has_voxels"]:
            out += struct.pack("<?", bool(d["e_has_vox"][i]))
            if d["e_has_vox"][i]:
                lo, hi = int(d["e_off"][i]), int(d["e_off"][i + 1])
                out += _pack_blocks(wl, d["e_keys"][lo:hi], d["e_bits"][lo:hi], Nb)
    return out


def _roadmap(orc, wl, n=24):
    spec = wl.robot_b(0.003)
    g = wl.workspace_grid(spec)
    rb = orc.robot(spec)
    grid = orc.grid(g["Ng"], g["lim"])
    st = wl.sample_states(spec, n, stream=401)
    vs, vflags = orc.voxelize_vertices_batch(rb, grid, st)
    pairs = wl.knn_edges(spec, st, k=2)
    es, einfo = orc.voxelize_edges_batch(rb, grid, orc.space(), st[pairs[:, 0]], st[pairs[:, 1]])
    voff, vkeys, vbits = vs.export()
    eoff, ekeys, ebits = es.export()
    ref = orc.fk_batch(rb, st, 68, want_p=False)
    d = dict(n_verts=n, n_edges=len(pairs), has_voxels=True, Ng=g["Ng"], lims=g["lim"],
             v_index=np.arange(n, dtype=np.uint32), v_state=st, v_has_tip=np.arange(n) % 5 != 0, v_tip=ref["tip"],
             v_has_vox=(np.diff(voff.astype(np.int64)) > 0), v_off=voff, v_keys=vkeys, v_bits=vbits,
             e_src=pairs[:, 0].astype(np.uint32), e_dst=pairs[:, 1].astype(np.uint32),
             e_weight=np.linalg.norm(st[pairs[:, 0]] - st[pairs[:, 1]], axis=1),
             e_has_vox=(np.diff(eoff.astype(np.int64)) > 0), e_off=eoff, e_keys=ekeys, e_bits=ebits)
    return spec, g, d


def test_rmp_roundtrip_and_reference_layout(tmp_path, orc, wl):
    import irt_b200
    spec, g, d = _roadmap(orc, wl)
    ref_bytes = _reference_style_bytes(wl, d)
    ref_file = tmp_path / "ref.rmp"
    ref_file.write_bytes(ref_bytes)
    r = irt_b200.read_rmp(str(ref_file))
    assert r["n_verts"] == d["n_verts"] and r["n_edges"] == d["n_edges"] and r["Ng"] == 128
    assert np.allclose(r["lims"], d["lims"], atol=0)
    for k in ("v_index", "v_state", "v_has_tip", "v_has_vox", "v_off", "v_keys", "v_bits",
              "e_src", "e_dst", "e_weight", "e_has_vox", "e_off", "e_keys", "e_bits"):
        assert np.array_equal(r[k], np.asarray(d[k])), k
    assert np.array_equal(r["v_tip"][d["v_has_tip"]], d["v_tip"][d["v_has_tip"]])
    out_file = tmp_path / "out.rmp"
    irt_b200.write_rmp(str(out_file), r)
    assert out_file.read_bytes() == ref_bytes          # byte-identical to the reference layout
    # no-voxel roadmap
    d2 = dict(d, has_voxels=False)
    f2 = tmp_path / "novox.rmp"
    f2.write_bytes(_reference_style_bytes(wl, d2))
    r2 = irt_b200.read_rmp(str(f2))
    assert not r2["has_voxels"] and r2["v_keys"].size == 0 and np.array_equal(r2["v_state"], d["v_state"])
    # truncated file is rejected
    bad = tmp_path / "bad.rmp"
    bad.write_bytes(ref_bytes[:len(ref_bytes) // 2])
    with pytest.raises(irt_b200.IrtError):
        irt_b200.read_rmp(str(bad))


@pytest.mark.gpu
def test_rmp_loads_into_device_store(tmp_path, orc, wl):
    """a reference-style .rmp goes file -> CSR -> device store -> verdicts, and a GPU-built roadmap
    is written back byte-identically"""
    import irt_b200 as irt
    spec, g, d = _roadmap(orc, wl, n=64)
    f = tmp_path / "ref.rmp"
    f.write_bytes(_reference_style_bytes(wl, d))
    r = irt.read_rmp(str(f))
    ctx = irt.Context(0)
    grid = irt.make_grid(r["Ng"], r["lims"])
    es = irt.SetStore(ctx, grid)
    es.import_csr(r["e_off"], r["e_keys"], r["e_bits"])
    env_blocks = wl.dense_to_morton_blocks(wl.lung_like_env_dense(spec, g))
    env = irt.Env(ctx, grid)
    env.update(env_blocks)
    got = es.check(env)
    x = d["e_bits"] & env_blocks[d["e_keys"]]
    want = np.add.reduceat((x != 0).astype(np.int64), d["e_off"][:-1].astype(np.int64)) > 0
    want[np.diff(d["e_off"].astype(np.int64)) == 0] = False
    assert np.array_equal(got, want)
    # GPU-built caches written in the reference's format are byte-identical to the oracle-built file
    rb = irt.Robot(ctx, spec)
    vs = irt.SetStore(ctx, grid)
    vs.voxelize_vertices(rb, r["v_state"])
    es2 = irt.SetStore(ctx, grid)
    es2.voxelize_edges(rb, irt.make_space(), r["v_state"][r["e_src"]], r["v_state"][r["e_dst"]])
    voff, vkeys, vbits = vs.export_csr()
    eoff, ekeys, ebits = es2.export_csr()
    out = dict(r, v_off=voff, v_keys=vkeys, v_bits=vbits, e_off=eoff, e_keys=ekeys, e_bits=ebits)
    g_file = tmp_path / "gpu.rmp"
    irt.write_rmp(str(g_file), out)
    assert g_file.read_bytes() == f.read_bytes()


def test_rmp_against_reference_reader_and_writer(tmp_path, orc, wl):
    """The native .rmp writer / reader against the REFERENCE'S OWN RmpStreamer / LazyRmpParser
    (VoxelCachedLazyPRM.cpp:635-1114, compiled from their own text into oracle/_ref/librmp_ref.so):
    same roadmap -> byte-identical files; each side parses the other's file."""
    from oracle import ref
    if not ref.RefRmp.available():
        pytest.skip("oracle/_ref/librmp_ref.so not built (no /root/reference here)")
    import irt_b200
    for has_voxels in (True, False):
        spec, g, d = _roadmap(orc, wl)
        d = dict(d, has_voxels=has_voxels)
        Nb = d["Ng"] // 4
        dd = dict(d)
        for pre in ("v", "e"):
            bx, by, bz = wl.morton_decode(np.asarray(d[pre + "_keys"], dtype=np.uint32), Nb)
            dd[pre + "_bxyz"] = np.stack([bx, by, bz], axis=1).astype(np.uint8)
        ref_file, own_file = tmp_path / ("ref%d.rmp" % has_voxels), tmp_path / ("own%d.rmp" % has_voxels)
        ref.RefRmp.write(str(ref_file), dd)
        irt_b200.write_rmp(str(own_file), d)
        assert own_file.read_bytes() == ref_file.read_bytes()
        assert ref_file.read_bytes() == _reference_style_bytes(wl, d)     # and the struct.pack restatement
        # the reference's parser on the native file
        items = ref.RefRmp.read(str(own_file))
        assert len(items) == d["n_verts"] + d["n_edges"]
        # the reference's binary_read(std::optional<T>&) (VoxelCachedLazyPRM.cpp:715-724) leaves the optional
        # untouched when the file says "absent", and LazyRmpParser::next() (:899-903) reuses one _current_vertex:
        # a vertex written without a tip comes back with the LAST tip read before it.  The native reader
        # reports what the file holds (v_has_tip); the expectation here follows the reference's carry-over.
        last_tip = None
        for i in range(d["n_verts"]):
            kind, idx, st, tip, lv = items[i]
            assert kind == "vertex" and idx == d["v_index"][i] and np.array_equal(st, d["v_state"][i])
            if d["v_has_tip"][i]:
                last_tip = np.asarray(d["v_tip"][i], dtype=np.float64)
            assert (tip is not None) == (last_tip is not None)
            if tip is not None:
                assert np.array_equal(tip, last_tip)
            lo, hi = int(d["v_off"][i]), int(d["v_off"][i + 1])
            if has_voxels and d["v_has_vox"][i]:
                assert np.array_equal(lv[:, :3].astype(np.uint8), dd["v_bxyz"][lo:hi])
                assert np.array_equal(lv[:, 3], d["v_bits"][lo:hi])
            else:
                assert lv is None
        for i in range(d["n_edges"]):
            kind, src, dst, w, lv = items[d["n_verts"] + i]
            assert kind == "edge" and (src, dst, w) == (d["e_src"][i], d["e_dst"][i], d["e_weight"][i])
            lo, hi = int(d["e_off"][i]), int(d["e_off"][i + 1])
            if has_voxels and d["e_has_vox"][i]:
                assert np.array_equal(lv[:, 3], d["e_bits"][lo:hi])
            else:
                assert lv is None
        # the native reader on the reference's file
        r = irt_b200.read_rmp(str(ref_file))
        assert r["n_verts"] == d["n_verts"] and bool(r["has_voxels"]) == has_voxels
        assert np.array_equal(r["v_state"], d["v_state"]) and np.array_equal(r["e_weight"], d["e_weight"])
        if has_voxels:
            assert np.array_equal(r["v_keys"], d["v_keys"]) and np.array_equal(r["e_bits"], d["e_bits"])


def test_rmp_reader_rejects_truncated_and_corrupt_files(tmp_path, orc, wl):
    """Counts inside a .rmp file are untrusted input: a file cut anywhere, or with a count far beyond the bytes
    that follow, must come back as an error status -- no crash, no attempt to allocate what the count claims."""
    import irt_b200
    spec, g, d = _roadmap(orc, wl, n=6)
    good = tmp_path / "good.rmp"
    irt_b200.write_rmp(str(good), d)
    blob = good.read_bytes()
    assert irt_b200.read_rmp(str(good))["n_verts"] == 6
    bad = tmp_path / "bad.rmp"
    # every proper prefix (every byte in the header and the first records, then a stride through the rest)
    cuts = list(range(0, 400)) + list(range(400, len(blob), 37)) + [len(blob) - 1]
    for cut in cuts:
        bad.write_bytes(blob[:cut])
        with pytest.raises(irt_b200.IrtError):
            irt_b200.read_rmp(str(bad))
    # trailing garbage is not read (the reader stops after n_verts + n_edges records, like LazyRmpParser)
    bad.write_bytes(blob + b"\x00" * 7)
    assert irt_b200.read_rmp(str(bad))["n_edges"] == d["n_edges"]

    def patched(offset, fmt, value):
        b = bytearray(blob)
        struct.pack_into(fmt, b, offset, value)
        bad.write_bytes(bytes(b))
        with pytest.raises(irt_b200.IrtError) as e:
            irt_b200.read_rmp(str(bad))
        return e.value.status

    hdr = 4 + 4 + 1 + 1 + 48            # nV, nE, has_voxels, Nb, 6 limits
    assert patched(0, "<I", 0xFFFFFFFF) == irt_b200.IRT_ERR_INVALID_ARGUMENT      # 4e9 vertices, 6 in the file
    assert patched(4, "<I", 0xFFFFFFF0) == irt_b200.IRT_ERR_INVALID_ARGUMENT      # 4e9 edges
    assert patched(9, "<B", 0) == irt_b200.IRT_ERR_INVALID_ARGUMENT               # Nb = 0
    assert patched(9, "<B", 48) == irt_b200.IRT_ERR_INVALID_ARGUMENT              # Ng = 192: not a power of two
    assert patched(hdr + 4, "<I", 0x7FFFFFFF) == irt_b200.IRT_ERR_INVALID_ARGUMENT  # state size of vertex 0: 2^31 doubles
    # the first vertex with voxels: patch its block count to 4e9 and a block coordinate beyond Nb
    S = d["v_state"].shape[1]
    off = hdr
    i = 0
    while not d["v_has_vox"][i]:
        off += 8 + 8 * S + 1 + (24 if d["v_has_tip"][i] else 0) + 1
        i += 1
    off += 8 + 8 * S + 1 + (24 if d["v_has_tip"][i] else 0) + 1
    assert patched(off, "<I", 0xFFFFFFFF) == irt_b200.IRT_ERR_INVALID_ARGUMENT
    assert patched(off + 4, "<B", 200) == irt_b200.IRT_ERR_INVALID_ARGUMENT       # bx = 200 >= Nb = 32


def test_rmp_writer_rejects_inconsistent_descriptors(tmp_path, orc, wl):
    import ctypes as C
    import irt_b200
    spec, g, d = _roadmap(orc, wl, n=4)
    out = str(tmp_path / "x.rmp")
    with pytest.raises(irt_b200.IrtError):
        irt_b200.write_rmp(out, dict(d, Ng=4 * 200))                              # Nb does not fit the u8 of the format
    L = irt_b200.lib()
    r = irt_b200.Rmp()
    r.n_verts, r.state_size = 3, -1
    assert L.irt_rmp_write(out.encode(), C.byref(r)) == irt_b200.IRT_ERR_INVALID_ARGUMENT
    r.state_size = 7                                                               # vertices declared, no arrays
    assert L.irt_rmp_write(out.encode(), C.byref(r)) == irt_b200.IRT_ERR_INVALID_ARGUMENT
    assert L.irt_rmp_read(out.encode(), None) == irt_b200.IRT_ERR_INVALID_ARGUMENT
    p = C.POINTER(irt_b200.Rmp)()
    assert L.irt_rmp_read(str(tmp_path / "missing.rmp").encode(), C.byref(p)) == irt_b200.IRT_ERR_INVALID_ARGUMENT


def test_rmp_reader_survives_random_corruption(tmp_path, orc, wl):
    """byte-level fuzzing of a valid file: whatever is flipped, the reader returns a roadmap or an error status
    (and what it returns is internally consistent) -- it never crashes and never allocates by a corrupt count"""
    import irt_b200
    spec, g, d = _roadmap(orc, wl, n=5)
    good = tmp_path / "good.rmp"
    irt_b200.write_rmp(str(good), d)
    blob = np.frombuffer(good.read_bytes(), dtype=np.uint8)
    rng = np.random.default_rng(2022)
    bad = tmp_path / "fuzz.rmp"
    parsed = failed = 0
    for it in range(400):
        b = blob.copy()
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(0, len(b) if it % 2 else min(len(b), 200)))     # every other round: header region
            b[pos] = rng.integers(0, 256)
        bad.write_bytes(b.tobytes())
        try:
            r = irt_b200.read_rmp(str(bad))
        except irt_b200.IrtError:
            failed += 1
            continue
        parsed += 1
        assert len(r["v_off"]) == r["n_verts"] + 1 and len(r["e_off"]) == r["n_edges"] + 1
        assert len(r["v_keys"]) == int(r["v_off"][-1]) == len(r["v_bits"])
        assert len(r["e_keys"]) == int(r["e_off"][-1]) == len(r["e_bits"])
        if r["has_voxels"] and len(r["v_keys"]):
            assert int(r["v_keys"].max()) < (r["Ng"] // 4) ** 3
    assert parsed > 0 and failed > 0


def test_rmp_io_under_sanitizers(tmp_path):
    """csrc/rmp_io.cpp compiled with -fsanitize=address,undefined into a fuzz harness (tests/cpp/fuzz_rmp_asan.cpp):
    3000 byte-flipped / truncated / spliced copies of a valid roadmap are read; no sanitizer report, every accepted
    file survives a write / read round trip"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fuzz_rmp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                           "-Wall", "-Wextra", "-Werror", os.path.join(root, "tests", "cpp", "fuzz_rmp_asan.cpp"),
                           os.path.join(root, "interactive-rate-tendons_b200", "csrc", "rmp_io.cpp"), "-o", exe])
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1")
    env.pop("LD_PRELOAD", None)
    out = subprocess.run([exe, str(tmp_path), "3000"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "rmp fuzz ok" in out.stdout and "runtime error" not in out.stderr
