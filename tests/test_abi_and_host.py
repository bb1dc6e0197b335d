"""CPU-side tests: the C-ABI library loads and exports every declared symbol, the host logic
(sharding, verdict gather/assembly, Morton keys, workloads) and the N>1 path over gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def irt():
    import __graft_entry__ as ge
    import irt_b200
    if not os.path.exists(irt_b200.LIB_PATH):
        ge.build()
    return irt_b200


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "irt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(irt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(irt):
    declared = _declared_symbols()
    assert len(declared) >= 35
    assert sorted(irt.ABI_SYMBOLS) == declared, "python binding list out of sync with the header"
    lib = C.CDLL(irt.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libirt_b200.so does not export %s" % name
    assert irt.lib().irt_abi_version() == 1
    assert b"no CPU fallback" in irt.lib().irt_status_string(irt.IRT_ERR_NO_DEVICE)


def test_no_cpu_fallback_without_device(irt):
    """on a box without a GPU the product path must fail loudly, not compute on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(irt.IrtError) as ei:
        irt.Context(0)
    assert ei.value.status == irt.IRT_ERR_NO_DEVICE


def test_library_has_sm100a_kernels(irt):
    out = subprocess.run(["cuobjdump", "-lelf", irt.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "interactive-rate-tendons_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "tendon_oracle" not in txt and "liboracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f


def test_morton_key_matches_reference_child_order(irt, orc):
    L = irt.lib()
    rng = np.random.default_rng(1)
    for Nb in (1, 2, 8, 32, 128):
        for _ in range(50):
            bx, by, bz = (int(v) for v in rng.integers(0, Nb, 3))
            k = L.irt_morton_key(bx, by, bz, Nb)
            assert k == orc.morton_key(bx, by, bz, Nb)
            x, y, z = C.c_int(), C.c_int(), C.c_int()
            L.irt_morton_decode(k, Nb, C.byref(x), C.byref(y), C.byref(z))
            assert (x.value, y.value, z.value) == (bx, by, bz)
    # child index = bz/c + 2 by/c + 4 bx/c (collision/detail/TreeNode.h:66-68)
    assert L.irt_morton_key(1, 0, 0, 2) == 4 and L.irt_morton_key(0, 1, 0, 2) == 2 and L.irt_morton_key(0, 0, 1, 2) == 1


def test_valid_segment_count_host(irt, orc, wl):
    for spec in (wl.robot_a(), wl.robot_b(), wl.robot_b(rotation=True)):
        st = wl.sample_states(spec, 40, stream=5)
        desc = irt.robot_desc(spec)
        sp = irt.make_space()
        for i in range(0, 40, 2):
            a, b = np.ascontiguousarray(st[i]), np.ascontiguousarray(st[i + 1])
            got = irt.lib().irt_valid_segment_count(C.byref(desc), C.byref(sp), a.ctypes.data, b.ctypes.data)
            assert got == orc.valid_segment_count(orc.robot(spec), orc.space(), a, b)


def test_shard_ranges_cover_and_align(irt):
    for n in (0, 1, 63, 64, 65, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = irt.shard_range(n, r, world)
                assert lo == prev and lo <= hi <= n
                assert lo % 64 == 0 or lo == n
                prev = hi
            assert prev == n


def test_verdict_pack_assemble_roundtrip(irt):
    from irt_b200.roadmap import assemble_verdicts, shard_words
    rng = np.random.default_rng(2)
    for n, world in ((1000, 1), (1000, 2), (129, 4), (5, 8)):
        truth = rng.random(n) < 0.3
        w = shard_words(n, world)
        allw = np.zeros(world * w, dtype=np.uint32)
        for r in range(world):
            lo, hi = irt.shard_range(n, r, world)
            bits = np.zeros(w * 32, dtype=np.uint8)
            bits[:hi - lo] = truth[lo:hi]
            allw[r * w:(r + 1) * w] = np.packbits(bits, bitorder="little").view(np.uint32)
        assert np.array_equal(assemble_verdicts(allw, n, world), truth)


def test_workloads_are_seeded_and_shaped(wl):
    a = wl.sample_states(wl.robot_b(), 100, stream=3)
    b = wl.sample_states(wl.robot_b(), 100, stream=3)
    assert np.array_equal(a, b) and a.shape == (100, 7)
    assert np.all(a[:, :6] >= 0) and np.all(a[:, :6] <= 20) and np.all(a[:, 6] >= 0) and np.all(a[:, 6] <= 0.2)
    e = wl.knn_edges(wl.robot_b(), a, k=3)
    assert e.shape[1] == 2 and np.all(e[:, 0] < e[:, 1]) and len(np.unique(e, axis=0)) == len(e)
    occ = np.zeros((8, 8, 8), dtype=bool)
    occ[5, 2, 7] = True
    blocks = wl.dense_to_morton_blocks(occ)
    k = int(wl.morton_key(1, 0, 1, 2))
    assert blocks[k] == np.uint64(1) << np.uint64(1 * 16 + 2 * 4 + 3) and np.count_nonzero(blocks) == 1
    bx, by, bz = wl.morton_decode(np.array([k], dtype=np.uint32), 2)
    assert (int(bx[0]), int(by[0]), int(bz[0])) == (1, 0, 1)


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import irt_b200
from irt_b200.roadmap import gather_verdict_words, assemble_verdicts, shard_words
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
n = 1000
truth = np.random.default_rng(4).random(n) < 0.4      # same on every rank
lo, hi = irt_b200.shard_range(n, rank, world)
w = shard_words(n, world)
bits = np.zeros(w * 32, dtype=np.uint8)
bits[:hi - lo] = truth[lo:hi]                          # this rank only knows its shard
local = torch.from_numpy(np.packbits(bits, bitorder="little").view(np.int32).copy())
allw = gather_verdict_words(local, dist)
got = assemble_verdicts(allw.numpy().view(np.uint32), n, world)
assert np.array_equal(got, truth), "rank %%d: gathered verdicts differ" %% rank
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


def test_two_rank_gloo_verdict_gather(tmp_path):
    """world_size-2 run of the N>1 host path (shard -> pack -> all_gather -> assemble) on CPU"""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_cpp_host_mirror_compiles_and_links(tmp_path, irt, orc):
    """the whole C++ host mirror test links against the C-ABI library here; it must then refuse to run
    without a device (no CPU fallback) -- the run itself is tests/test_gpu_host_cpp.py"""
    exe = str(tmp_path / "test_host_mirror")
    pkg = os.path.dirname(irt.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O0", "-Wall", "-Wextra", "-Werror",
                           os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"), "-o", exe,
                           "-L" + pkg, "-lirt_b200", "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                           "-Wl,-rpath," + pkg, "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"])
    import torch
    if not torch.cuda.is_available():
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert out.returncode != 0 and "no CUDA device" in out.stderr


def test_cpp_host_mirror_host_logic(tmp_path, orc):
    """The C++ host mirror's logic above the boundary, WITHOUT a GPU: the same test source that runs against
    libirt_b200.so on the GPU box (tests/cpp/test_host_mirror.cpp) is linked against a test-only stand-in of
    the C ABI answered by the oracle (tests/cpp/abi_standin_over_oracle.cpp).  Covers the layout conversions
    (column-major R, Morton keys, CSR <-> block maps), status -> exception translation, home_shape, the
    validators' PartialVoxelization assembly, the batch entry points of VoxelCachedLazyPRM incl.
    createRoadmap(N, opt) (rejection rounds, KBounded connection, removal of invalid edges, growing) and the
    batched Jacobians."""
    exe = str(tmp_path / "test_host_mirror_cpu")
    cpp = os.path.join(ROOT, "tests", "cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-DIRT_TEST_OVER_STANDIN",
                           os.path.join(cpp, "test_host_mirror.cpp"),
                           os.path.join(cpp, "abi_standin_over_oracle.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "host mirror ok" in out.stdout and "createRoadmap: 120 vertices" in out.stdout


def test_cpp_roadmap_ik_every_branch(tmp_path, orc):
    """roadmapIk of the C++ host mirror, every branch (accepted, closest valid, stepping back, RMAP_IK_AUTO_ADD connected
    / the closest collision-free connection with and without RMAP_IK_LAZY_ADD, each with and without RMAP_IK_ACCURATE),
    against the reference's sequential loop (VoxelCachedLazyPRM.cpp:3095-3565) restated over the oracle in
    tests/cpp/test_roadmap_ik_host.cpp: the same result, removals, added vertices / edges and validity; and chainedPlan
    (the milestone loop of apps/roadmap_chained_plan.cpp:535-679: addMilestone's lazy connections checked on first use,
    paths valid and shortest by the oracle, no further sweep, no cache rebuild).  Host logic: linked against the
    test-only stand-in of the C ABI."""
    exe = str(tmp_path / "test_roadmap_ik_host")
    cpp = os.path.join(ROOT, "tests", "cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-DIRT_TEST_OVER_STANDIN",
                           os.path.join(cpp, "test_roadmap_ik_host.cpp"),
                           os.path.join(cpp, "abi_standin_over_oracle.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "roadmap ik ok" in out.stdout and "chainedPlan:" in out.stdout
    for kind in ("accepted", "closest_valid", "stepped_back", "connected", "self", "fallback"):
        assert kind in out.stdout


def test_capsule_pair_arithmetic_on_host(tmp_path, orc):
    """The arithmetic of the exact stage of the device self-collision test (csrc/capsule_pair.h: closest_st,
    capsules_collide, segment lengths, the index-gap rule -- the file selfcol.cu includes) compiled for the host with
    -ffp-contract=off: (s, t) of closest_st bit-equal to the oracle's closest_st_segment on 6 M segment pairs of every
    kind, and the kernel's decision procedure composed serially (chunk bounding spheres, chunk-pair pruning, loop bounds,
    arc-length rule, capsule test) equal to the oracle's collides_self on 92 k backbones that sit on its decision
    boundaries (hairpins at 2r +- ulps, corners around the 3r rule, arcs, spirals, random walks); the FP32 filter in
    front of it, restated, never lets a backbone go that the oracle finds colliding."""
    exe = str(tmp_path / "test_capsule_pair_host")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror",
                           os.path.join(ROOT, "tests", "cpp", "test_capsule_pair_host.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "capsule pair ok" in out.stdout and ", 0 differ" in out.stdout and " 0 verdicts differ" in out.stdout
    assert "; 0 true hits dropped" in out.stdout


def test_env_primitive_arithmetic_on_host(tmp_path, orc):
    """The per-leaf-block arithmetic the device kernel env_add_primitives_kernel runs (csrc/env_prims.h: voxel
    centres inside spheres / capsules, add_point cells) compiled for the host with -ffp-contract=off and
    compared with the oracle's add_point / add_sphere / add_capsule on four grids: 0 flips."""
    exe = str(tmp_path / "test_env_prims_host")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror",
                           os.path.join(ROOT, "tests", "cpp", "test_env_prims_host.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "env primitives ok" in out.stdout and out.stdout.count(" 0 flips") == 4


def test_raster_line_traversal_on_host(tmp_path, orc):
    """The voxel traversal the device rasteriser runs (csrc/raster_line.h: add_line with the division-free decisions,
    the hand-over to the literal code, path-order emission -- the file voxel_raster.cu includes) compiled for the
    host with -ffp-contract=off and compared with the oracle's add_line (pinned by the reference's own text) on
    2.6 M segments of every kind on four grids: 0 differences, both paths exercised."""
    exe = str(tmp_path / "test_raster_line_host")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror",
                           os.path.join(ROOT, "tests", "cpp", "test_raster_line_host.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "raster line ok" in out.stdout and out.stdout.count(" 0 groups differ") == 4
