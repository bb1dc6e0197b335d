"""Pins of the CPU oracle against the REFERENCE'S OWN CODE (no GPU needed).

Two layers:
* golden: tests/golden/reference_vectors.npz holds outputs of the reference's unmodified sources
  compiled in the build container (tests/golden/make_reference_vectors.py, oracle/ref.py);
  these tests run anywhere.
* live: when oracle/_ref/ is present (it is built wherever /root/reference exists and travels with
  the repo snapshot), the oracle is compared with the reference pieces directly on more inputs.

What "reference" means here, precisely: collision/detail/TreeNode.h is the real thing (std-only);
tendon_deriv / solve_initial_bending / get_r_info2 / closest_st_segment / segment_aabox_intersect are
the reference's unmodified .cpp files compiled against a stand-in for Eigen's fixed-size algebra
(oracle/ref_shim/eigen_standin/Eigen/Core), so they pin the reference's formulas, operand order and
control flow, not Eigen's rounding.  The RK4 stepping (Boost.odeint), t_range, add_line and the
swept-volume driver remain restated-only.
"""
import os

import numpy as np
import pytest

from oracle import ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz")
live = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (no /root/reference here)")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def robots(wl):
    return {"a005": wl.robot_a(0.005), "b003": wl.robot_b(0.003),
            "b005rot": wl.robot_b(0.005, rotation=True)}


# ------------------------------------------------------------------ golden (reference outputs)
@pytest.mark.parametrize("name", ["a005", "b003", "b005rot"])
def test_golden_routing_bit_exact(orc, gold, robots, name):
    rb = orc.robot(robots[name])
    for t, want in zip(gold[name + "_t"], gold[name + "_rinfo"]):
        got = np.stack(orc.routing(rb, float(t)))
        assert np.array_equal(got, want)          # get_r_info.cpp:105-144, bit for bit


@pytest.mark.parametrize("name", ["a005", "b003", "b005rot"])
def test_golden_tendon_deriv(orc, gold, robots, name):
    rb = orc.robot(robots[name])
    worst = 0.0
    for t, tau, x, want in zip(gold[name + "_t"], gold[name + "_tau"], gold[name + "_x"],
                               gold[name + "_dxdt"]):
        got = orc.deriv(rb, tau, x, float(t))
        worst = max(worst, np.abs(got - want).max() / max(1.0, np.abs(want).max()))
    assert worst <= 1e-13, worst                 # tendon_deriv.cpp:95-178 (observed: 0)


@pytest.mark.parametrize("name", ["a005", "b003", "b005rot"])
def test_golden_initial_bending(orc, gold, robots, name):
    spec = robots[name]
    rb = orc.robot(spec)
    N = len(spec["C"])
    for tau, s0, v0, u0, it in zip(gold[name + "_tau"], gold[name + "_s0"], gold[name + "_v0"],
                                   gold[name + "_u0"], gold[name + "_iters"]):
        state = list(tau) + ([0.0] if spec.get("enable_rotation") else []) + \
            ([float(s0)] if spec.get("enable_retraction") else [])
        assert len(state) == orc.state_size(rb) and len(tau) == N
        s = orc.shape(rb, state)
        assert s["iters"] == int(it)             # solve_initial_bending.cpp:15-73: same trip count
        assert np.allclose(s["v_i"], v0, rtol=0, atol=1e-15)
        assert np.allclose(s["u_i"], u0, rtol=1e-14, atol=1e-13)


def test_golden_closest_st_segment(orc, gold):
    for seg, want in zip(gold["segs"], gold["segs_st"]):
        got = orc.closest_st_segment(*seg)
        assert got == (want[0], want[1])         # collision_primitives.cpp:10-102, all branches


def test_golden_segment_aabox(orc, gold):
    hits = np.array([orc.segment_aabox_intersect(*b) for b in gold["boxes"]], dtype=np.uint8)
    assert np.array_equal(hits, gold["boxes_hit"]) and 0 < hits.sum() < len(hits)


def test_golden_octree_edit_script(orc, gold):
    """set_block / union_block replayed on the oracle's octree give the reference TreeNode's leaves
    in the reference's visit_leaves order (TreeNode.hxx:75-190)."""
    g = orc.grid(32, [0, 1, 0, 1, 0, 1])
    t = orc.octree(g)
    for op, bx, by, bz, val in gold["tree_script"].tolist():
        (t.set_block, t.union_block)[op](bx, by, bz, val)
    bxyz, bits = t.export()
    assert np.array_equal(bxyz, gold["tree_leaves_xyz"])
    assert np.array_equal(bits, gold["tree_leaves_bits"])
    assert t.nblocks() == len(bits)


def test_golden_file_matches_generator(gold):
    keys = set(gold.files)
    assert {"segs", "segs_st", "boxes", "tree_script", "a005_dxdt", "b003_rinfo"} <= keys


# ------------------------------------------------------------------ live (oracle/_ref present)
@live
@pytest.mark.parametrize("name", ["a005", "b003", "b005rot"])
def test_live_shape_against_reference_deriv(orc, wl, robots, name):
    """Whole shapes: RK4 over the oracle's own t grid with the reference's tendon_deriv and
    solve_initial_bending inside, vs the oracle's shape().  Positions within 1e-13 L."""
    spec = robots[name]
    rb = orc.robot(spec)
    rf = ref.RefFK(spec)
    N = rf.N
    states = wl.sample_states(spec, 40, stream=11)
    worst = 0.0
    for s in states:
        sh = orc.shape(rb, s)
        if len(sh["t"]) < 2:
            continue
        st, nsteps = rf.shape_states(s[:N], sh["t"])
        assert nsteps == sh["nsteps"]
        rot = s[N] if spec.get("enable_rotation") else 0.0
        p = st[:, :3]
        if rot:                                   # TendonResult::rotate_z (TendonResult.cpp:13-18)
            c, sn = np.cos(rot), np.sin(rot)
            p = p @ np.array([[c, -sn, 0], [sn, c, 0], [0, 0, 1]]).T
        worst = max(worst, np.abs(p - sh["p"]).max() / spec["L"])
        assert abs(st[-1, 18] - sh["L"]) <= 1e-13
        assert np.allclose(st[-1, 19:], sh["L_i"], rtol=0, atol=1e-13)
    assert worst <= 1e-13, worst


@live
def test_live_unopt_solver_agrees(wl, robots):
    """KAT (vii): the reference's block-elimination solve vs its dense 6x6 solve (stand-in: LU)."""
    rng = np.random.default_rng(5)
    from tests.golden.make_reference_vectors import random_ode_state
    for spec in robots.values():
        rf = ref.RefFK(spec)
        for _ in range(20):
            x = random_ode_state(rng, rf.N)
            tau = rng.uniform(0, 20, rf.N)
            t = rng.uniform(0, spec["L"])
            a, b = rf.deriv(tau, x, t), rf.deriv(tau, x, t, unopt=True)
            assert np.abs(a - b).max() <= 1e-10 * max(1.0, np.abs(a).max())


@live
@pytest.mark.parametrize("Ng", [4, 8, 64, 128, 512])
def test_live_octree_algebra(orc, Ng):
    """Random trees: block(), nblocks(), leaf order, collides() and add_voxels == union_tree."""
    rng = np.random.default_rng(Ng)
    g = orc.grid(Ng, [0, 1, 0, 1, 0, 1])
    Nb = Ng // 4
    for trial in range(6):
        ta, tb = orc.octree(g), orc.octree(g)
        ra, rb_ = ref.RefTree(Ng), ref.RefTree(Ng)
        n = int(rng.integers(0, 40))
        for (to, tr) in ((ta, ra), (tb, rb_)):
            for _ in range(n):
                bx, by, bz = (int(v) for v in rng.integers(0, Nb, 3))
                # sparse bit patterns so that collides() has both outcomes
                val = int(rng.integers(0, 2 ** 63)) & int(rng.integers(0, 2 ** 63)) & int(rng.integers(0, 2 ** 63))
                if rng.random() < 0.5:
                    to.union_block(bx, by, bz, val), tr.union_block(bx, by, bz, val)
                else:
                    to.set_block(bx, by, bz, val), tr.set_block(bx, by, bz, val)
        assert bool(ta.collides(tb)) == ra.collides(rb_)
        assert bool(tb.collides(ta)) == rb_.collides(ra)
        for to, tr in ((ta, ra), (tb, rb_)):
            bxyz, bits = to.export()
            x, y, z, rbits = tr.leaves()
            assert np.array_equal(bxyz, np.stack([x, y, z], axis=1).reshape(-1, 3))
            assert np.array_equal(bits, rbits)
            assert to.nblocks() == tr.nblocks()
        ta.add_voxels(tb)
        ra.union_tree(rb_)
        bxyz, bits = ta.export()
        x, y, z, rbits = ra.leaves()
        assert np.array_equal(bits, rbits) and np.array_equal(bxyz[:, 0], x)


@live
def test_live_verdicts_of_a_voxelised_roadmap(orc, wl):
    """The loop of VoxelCachedLazyPRM.cpp:1584-1591 with the reference's TreeNode::collides on sets the
    oracle voxelised, vs the oracle's check_sets_batch."""
    spec = wl.robot_b(dL=0.003)
    rb = orc.robot(spec)
    g = wl.workspace_grid(spec)
    og = orc.grid(g["Ng"], g["lim"], g["inv_rot"])
    states = wl.sample_states(spec, 300, stream=21)
    store, _ = orc.voxelize_vertices_batch(rb, og, states)
    env_blocks = wl.dense_to_morton_blocks(wl.lung_like_env_dense(spec, g))
    Nb = g["Ng"] // 4
    keys = np.nonzero(env_blocks)[0].astype(np.uint32)
    ex, ey, ez = wl.morton_decode(keys, Nb)
    oenv, renv = orc.octree(og), ref.RefTree(g["Ng"])
    for x, y, z, k in zip(ex.tolist(), ey.tolist(), ez.tolist(), keys.tolist()):
        oenv.set_block(x, y, z, int(env_blocks[k]))
        renv.set_block(x, y, z, int(env_blocks[k]))
    off, skeys, bits = store.export()
    bx, by, bz = wl.morton_decode(skeys, Nb)
    want = renv.check_csr(off, bx, by, bz, bits)
    got = orc.check_sets_batch(store, oenv).astype(bool)
    assert np.array_equal(got, want) and 0 < want.sum() < len(want)


# ------------------------------------------------------------------ levmar (vendored by the reference)
levmar = pytest.mark.skipif(not (ref.available() and ref.RefLevmar.available()),
                            reason="oracle/_ref/liblevmar_ref.so not built")


def _fk_wrap(orc, rb, L):
    """fk_wrap of tip-control/tip_control.cpp:92-122 on the oracle's FK"""
    def f(p):
        if rb.enable_retraction and p[-1] > L:
            return np.array([0.0, 0.0, L - p[-1]])
        return orc.shape(rb, p)["p"][-1]
    return f


@levmar
@pytest.mark.parametrize("name", ["a005", "b003", "b005rot"])
def test_live_levmar_finite_difference_rule(orc, wl, robots, name):
    """orc_tip_jacobian modes 1/2 == levmar-2.6's own dlevmar_fdif_{forw,cent}_jac_approx
    (misc_core.c:137-211) driving the same FK: bit for bit."""
    spec = robots[name]
    rb = orc.robot(spec)
    f = _fk_wrap(orc, rb, spec["L"])
    states = wl.sample_states(spec, 6, stream=33)
    if spec.get("enable_retraction"):
        states[0, -1] = 0.0
        states[1, -1] = spec["L"] - 1e-5
    for s in states:
        for central in (False, True):
            for delta in (1e-6, 1e-4):
                want = ref.RefLevmar.fdif_jac(f, s, 3, delta, central)
                _, got = orc.tip_jacobian(rb, s, 2 if central else 1, delta)
                assert np.array_equal(got, want)


@levmar
def test_live_ik_with_reference_levmar(orc, wl):
    """tip_control::inverse_kinematics (tip_control.cpp:34-153) = dlevmar_bc_dif over fk_wrap with the
    defaults of Controller.h:55-65: reaches the tip of a nearby configuration."""
    spec = wl.robot_b(0.005)
    rb = orc.robot(spec)
    L = spec["L"]
    f = _fk_wrap(orc, rb, L)
    st = wl.sample_states(spec, 4, stream=34)
    lb, ub = np.zeros(7), np.array([20.0] * 6 + [L])       # Bounds::from_robot, tip_control.cpp:160-176
    for s in st:
        goal_state = np.clip(s + np.array([0.8, -0.5, 0.3, 0.6, -0.4, 0.2, 0.004]), lb, ub)
        des = f(goal_state)
        p, info, rc = ref.RefLevmar.bc_dif(f, s, des, lb, ub, 100, [0.1, 1e-9, 1e-8, 1e-8, -1e-6])
        assert rc >= 0 and np.all(p >= lb) and np.all(p <= ub)
        assert np.linalg.norm(f(p) - des) < 2e-4 and info[1] < info[0]


# ------------------------------------------------------------------ VoxelOctree core (reference's own text)
vox = pytest.mark.skipif(not (ref.available() and ref.RefVoxelOctree.available()),
                         reason="oracle/_ref/libvoxeloctree_ref.so not built")


def _same_tree(a, b):
    ea, eb = a.export(), b.export()
    return np.array_equal(ea[0], eb[0]) and np.array_equal(ea[1], eb[1])


def test_golden_add_line(orc, gold):
    """VoxelOctree::add_line (VoxelOctree.cpp:325-426, both quirks included): segments of every kind,
    leaves bit-exact and in visit_leaves order."""
    lim = gold["vo_lim"].tolist()
    g = orc.grid(128, lim)
    segs, off = gold["vo_segs"], gold["vo_leaf_off"]
    for k in range(0, len(segs), 3):
        t = orc.octree(g)
        for a, b in segs[k:k + 3]:
            t.add_line(a, b)
        xyz, bits = t.export()
        lo, hi = off[k // 3], off[k // 3 + 1]
        assert np.array_equal(xyz, gold["vo_leaf_xyz"][lo:hi]) and np.array_equal(bits, gold["vo_leaf_bits"][lo:hi]), k
    assert off[-1] > 300    # the fixture is not trivially empty


def test_golden_find_cell(orc, gold):
    g = orc.grid(128, gold["vo_lim"].tolist())
    for p, want in zip(gold["vo_pts"], gold["vo_find_cell"]):
        got = orc.find_cell(g, p)                      # VoxelOctree.cpp:309-317, domain_check :1511-1521
        assert (got is None and want[0] < 0) or tuple(want) == got
    assert (gold["vo_find_cell"][:, 0] < 0).sum() >= 2   # the out-of-domain probes are in the fixture


@pytest.mark.parametrize("name,op", [("dilate6x1", ("dilate", 1, False)), ("dilate6x3", ("dilate", 3, False)),
                                     ("dilate27x2", ("dilate", 2, True)), ("sphere", ("dilate_sphere", 0.07)),
                                     ("interior27", ("remove_interior", True)),
                                     ("interior6", ("remove_interior", False))])
def test_golden_environment_preparation(orc, gold, name, op):
    """dilate_6/27neighbor, dilate_sphere, remove_interior_6/27neighbor (VoxelOctree.cpp:533-952)"""
    g = orc.grid(32, [0, 1] * 3)
    t = orc.octree(g)
    for x, y, z in gold["vo_env_cells"].tolist():
        t.union_block(x // 4, y // 4, z // 4, 1 << ((x % 4) * 16 + (y % 4) * 4 + (z % 4)))
    if name.startswith("interior"):
        t.dilate(3, True)
    getattr(t, op[0])(*op[1:])
    xyz, bits = t.export()
    assert np.array_equal(xyz, gold["vo_prep_%s_xyz" % name])
    assert np.array_equal(bits, gold["vo_prep_%s_bits" % name])


@vox
def test_live_backbone_voxelisation(orc, wl, robots):
    """VoxelBackboneValidityChecker::voxelize_impl (VoxelBackboneValidityChecker.h:49-57) =
    add_piecewise_line of the shape: oracle vs the reference's own add_line on real backbones."""
    for name in ("a005", "b003"):
        spec = robots[name]
        rb = orc.robot(spec)
        g = wl.workspace_grid(spec)
        og = orc.grid(g["Ng"], g["lim"])
        for s in wl.sample_states(spec, 150, stream=44):
            p = orc.shape(rb, s)["p"]
            t = ref.RefVoxelOctree(g["Ng"], g["lim"])
            t.add_piecewise_line(p)
            assert _same_tree(orc.voxelize_shape(og, p), t)
    assert ref.RefVoxelOctree.bitmask(1, 2, 3) == 1 << (16 + 8 + 3)


@vox
def test_live_lung_environment_preparation(orc, wl):
    """the C4/C5 environment through every preparation step, oracle vs reference"""
    spec = wl.robot_b(0.003)
    g = wl.workspace_grid(spec)
    blocks = wl.dense_to_morton_blocks(wl.lung_like_env_dense(spec, g))
    Nb = g["Ng"] // 4
    keys = np.nonzero(blocks)[0].astype(np.uint32)
    ex, ey, ez = wl.morton_decode(keys, Nb)
    og = orc.grid(g["Ng"], g["lim"])
    for op in (("dilate", 1, False), ("dilate", 2, True), ("dilate_sphere", 0.01),
               ("remove_interior", True), ("remove_interior", False)):
        to, tr = orc.octree(og), ref.RefVoxelOctree(g["Ng"], g["lim"])
        for x, y, z, k in zip(ex.tolist(), ey.tolist(), ez.tolist(), keys.tolist()):
            to.set_block(x, y, z, int(blocks[k]))
            tr.set_block(x, y, z, int(blocks[k]))
        getattr(to, op[0])(*op[1:])
        getattr(tr, op[0])(*op[1:])
        assert _same_tree(to, tr), op
        assert tr.collides(tr) and to.nblocks() == tr.nblocks() and to.ncells() == tr.ncells()


def test_golden_environment_primitives(orc, gold):
    """add(Point) / add_sphere / add_capsule of the reference (VoxelOctree.cpp:319-323, 434-515), committed as
    golden leaves: the oracle reproduces them where /root/reference and oracle/_ref are absent"""
    t = orc.octree(orc.grid(64, gold["vo_prim_lim"].tolist()))
    for o in gold["vo_prim_objs"]:
        if o[7] == 0:
            t.add_point(o[:3])
        elif o[7] == 1:
            t.add_sphere(o[:3], float(o[6]))
        else:
            t.add_capsule(o[:3], o[3:6], float(o[6]))
    xyz, bits = t.export()
    assert np.array_equal(xyz, gold["vo_prim_xyz"]) and np.array_equal(bits, gold["vo_prim_bits"])
    assert len(bits) > 100


@vox
@pytest.mark.skipif(not ref.RefVoxelOctree.has_primitives(), reason="libvoxeloctree_ref.so built without primitives")
@pytest.mark.parametrize("Ng,lim", [(16, [0, 1, 0, 1, 0, 1]), (64, [-0.3, 0.2, -0.1, 0.4, 0.0, 0.25]),
                                     (128, [-0.21, 0.21, -0.21, 0.21, -0.21, 0.21])])
def test_live_environment_primitives(orc, wl, Ng, lim):
    """Environment::voxelize's primitives (motion-planning/Environment.cpp:62-74): VoxelOctree::add(Point),
    add_sphere, add_capsule (VoxelOctree.cpp:319-323, 434-515) with collides(Sphere, Point) /
    collides(Capsule, Point) (collision.hxx:62-84): oracle vs the reference's own text, bit-exact -- objects
    inside, straddling a face, outside the grid, degenerate capsules, radii below a cell and above the grid."""
    rng = np.random.default_rng(1234 + Ng)
    lo, hi = np.array(lim[0::2]), np.array(lim[1::2])
    ext = hi - lo
    og = orc.grid(Ng, lim)
    to, tr = orc.octree(og), ref.RefVoxelOctree(Ng, lim)
    n_obj = 0
    for k in range(60):
        c = lo + ext * rng.uniform(-0.3, 1.3, 3)
        r = float(ext.min() * rng.choice([0.001, 0.01, 0.05, 0.15, 0.4, 2.0]))
        kind = k % 3
        if kind == 0:
            to.add_point(c); tr.add_point(c)
        elif kind == 1:
            to.add_sphere(c, r); tr.add_sphere(c, r)
        else:
            b = c if k % 9 == 2 else c + ext * rng.uniform(-0.5, 0.5, 3)      # a == b: closest_t's eps branch
            to.add_capsule(c, b, min(r, 0.2 * ext.min())); tr.add_capsule(c, b, min(r, 0.2 * ext.min()))
        n_obj += 1
        if k % 10 == 9:
            assert _same_tree(to, tr), (Ng, k)
    # points exactly on the limits (inclusive on both ends, clamped into the last cell) and cell corners
    for p in (lo, hi, (lo + hi) / 2, lo + ext / Ng * 3, hi + 1e-12, lo - 1e-12):
        to.add_point(p); tr.add_point(p)
    assert _same_tree(to, tr) and to.ncells() == tr.ncells() and to.ncells() > 0
    # the lung-like capsule tree the benchmark environment is made of
    if Ng == 128:
        spec = wl.robot_b(0.003)
        to, tr = orc.octree(og), ref.RefVoxelOctree(Ng, lim)
        for a, b, rad in wl.lung_like_capsules(spec):
            to.add_capsule(a, b, rad); tr.add_capsule(a, b, rad)
        assert _same_tree(to, tr) and tr.ncells() > 1000


# ------------------------------------------------------------------ collides_self (reference's own text)
@pytest.mark.skipif(not ref.RefSelfCollision.available(), reason="oracle/_ref/libselfcol_ref.so not built")
def test_live_collides_self(orc, wl):
    """collision::collides_self (collision.cpp:6-46) on real backbones, soft enough to curl back onto
    themselves: oracle vs the reference's own text, both outcomes well represented."""
    n = hit = 0
    for base in (wl.robot_a(0.005), wl.robot_b(0.003)):
        for E in (0.5e6, 2.1e6):
            spec = dict(base)
            spec["E"] = E
            rb = orc.robot(spec)
            for s in wl.sample_states(spec, 250, stream=13):
                p = orc.shape(rb, s)["p"]
                want = ref.RefSelfCollision.collides_self(p, spec["r"])
                assert orc.collides_self(p, spec["r"]) == want
                n += 1
                hit += want
    # degenerate inputs: fewer than three points never collide (collision.cpp:14)
    for k in (0, 1, 2):
        assert not ref.RefSelfCollision.collides_self(np.zeros((k, 3)), 0.015)
        assert not orc.collides_self(np.zeros((k, 3)), 0.015)
    assert 0.1 * n < hit < 0.9 * n


# ------------------------------------------------------------------ swept-volume driver (reference's own text)
@pytest.mark.skipif(not ref.RefSweptVolume.available(), reason="oracle/_ref/libsweptvol_ref.so not built")
@pytest.mark.parametrize("variant", ["b003", "b003rot_rotgrid", "a003_soft"])
def test_live_swept_volume_driver(orc, wl, variant):
    """VoxelEnvironment::voxelize_valid_backbone_motion (VoxelEnvironment.cpp:207-444): the oracle's
    voxelize_edge vs the reference's own text driven by the same FK / validity / interpolation callbacks.
    Leaves, is_fully_valid, t, last_valid and the number of FK calls (LIFO order) must be identical, on
    fully valid AND partially valid edges."""
    inv_rot = None
    if variant == "b003":
        spec = wl.robot_b(0.003)
    elif variant == "b003rot_rotgrid":
        spec = wl.robot_b(0.003, rotation=True)
        c, s = np.cos(0.3), np.sin(0.3)
        inv_rot = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]) @ np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    else:
        spec = wl.robot_a(0.003)
        spec["E"] = 1.0e6                                   # soft: edges run into invalid shapes
    rb = orc.robot(spec)
    g = wl.workspace_grid(spec)
    og = orc.grid(g["Ng"], g["lim"], inv_rot)
    sp = orc.space()
    st = wl.sample_states(spec, 1200, stream=17)
    frac = 0.5 if variant == "a003_soft" else 0.08
    n = partial = 0
    for i in range(0, 1200, 2):
        if n >= 30:
            break
        a = st[i]
        b = a + frac * (st[i + 1] - a)
        if orc.validity_flags(rb, a, orc.shape(rb, a)) != 0:
            continue                                        # the planner only starts edges at valid vertices
        tree, info = orc.voxelize_edge(rb, og, sp, a, b)
        if info["out_of_domain"]:
            continue
        cache = {}

        def fk(s_):
            sh = orc.shape(rb, s_)
            cache[s_.tobytes()] = sh
            return sh["p"]

        def valid(s_, p_):
            return orc.validity_flags(rb, s_, cache[s_.tobytes()]) == 0

        r = ref.RefSweptVolume.voxelize(g["Ng"], g["lim"], inv_rot, a, b,
                                        1.0 / orc.valid_segment_count(rb, sp, a, b), 128, fk, valid,
                                        lambda x, y, t: orc.interpolate(rb, x, y, t))
        assert r["rc"] == 0
        bxyz, bits = tree.export()
        assert np.array_equal(bxyz, r["bxyz"]) and np.array_equal(bits, r["bits"]), i
        assert r["is_fully_valid"] == info["is_fully_valid"] and r["t"] == info["t"], i
        assert r["nsamples"] == info["nsamples"], i
        assert np.array_equal(r["last_valid"], info["last_valid"]), i
        n += 1
        partial += not info["is_fully_valid"]
    assert n >= 10
    if variant == "a003_soft":
        assert partial >= 3, "fixture must contain partially valid edges"


# ------------------------------------------------------------------ TendonRobot::shape (reference's own text)
@pytest.mark.skipif(not ref.RefTendonRobot.available(), reason="oracle/_ref/libtendonrobot_ref.so not built")
@pytest.mark.parametrize("name", ["a005", "b003", "b005rot", "a003soft", "b005tight"])
def test_live_tendon_robot_shape(orc, wl, name):
    """TendonRobot::shape -> tension_shape (TendonRobot.cpp:325-500: t_range, initial condition, RK4 over the
    grid, observer, result assembly, calc_point_forces convergence flag), home_shape lengths (:249-314),
    rotate_z, and the three ingredients of is_valid_shape -- the oracle vs the reference's own text
    (Eigen and Boost.odeint stand-ins).  Grids, points, frames, lengths and flags are compared exactly,
    except rotate_z (Eigen's AngleAxis computes the zz entry as (1 - c) + c)."""
    spec = {"a005": wl.robot_a(0.005), "b003": wl.robot_b(0.003), "b005rot": wl.robot_b(0.005, rotation=True),
            "a003soft": dict(wl.robot_a(0.003), E=0.7e6),
            "b005tight": dict(wl.robot_b(0.005), residual_threshold=1e-13)}[name]
    rb, rr = orc.robot(spec), ref.RefTendonRobot(spec)
    st = wl.sample_states(spec, 200, stream=23)
    if spec.get("enable_retraction"):
        L = spec["L"]
        st[0, -1], st[1, -1], st[2, -1], st[3, -1], st[4, -1] = L, L + 0.01, 0.1995, 0.0, -0.001
    tol = 2e-16 if spec.get("enable_rotation") else 0.0
    seen = 0
    for s in st:
        a, b = orc.shape(rb, s), rr.shape(s)
        assert np.array_equal(a["t"], b["t"])
        for k in ("p", "R"):
            assert np.abs(a[k] - b[k]).max(initial=0.0) <= tol * max(1.0, spec["L"]), k
        assert a["L"] == b["L"] and np.array_equal(a["L_i"], b["L_i"])
        for k in ("u_i", "u_f", "v_i", "v_f"):
            assert np.array_equal(a[k], b[k]), k
        assert a["converged"] == b["converged"]
        fr = rr.flags(s)
        assert (orc.validity_flags(rb, s, a) & 7) == fr
        seen |= fr
        sret = s[-1] if spec.get("enable_retraction") else 0.0
        assert np.array_equal(orc.home_lengths(rb, sret), rr.home_lengths(s))
    if name == "b005tight":
        assert seen & 1, "fixture must contain non-converged shapes"
    if name == "a003soft":
        assert seen & 2 and seen & 4, "fixture must contain length-limit and self-collision cases"


@pytest.mark.parametrize("name", ["a005", "b003", "b005rot", "a003soft"])
def test_golden_tendon_robot_shape(orc, wl, gold, name):
    """whole shapes of the reference's TendonRobot::shape (see test_live_tendon_robot_shape), from the
    committed vectors: runs where oracle/_ref is absent"""
    spec = {"a005": wl.robot_a(0.005), "b003": wl.robot_b(0.003), "b005rot": wl.robot_b(0.005, rotation=True),
            "a003soft": dict(wl.robot_a(0.003), E=0.7e6)}[name]
    rb = orc.robot(spec)
    k = "tr_%s_" % name
    off = gold[k + "off"]
    tol = 2e-16 if spec.get("enable_rotation") else 0.0
    for i, s in enumerate(gold[k + "states"]):
        a = orc.shape(rb, s)
        lo, hi = off[i], off[i + 1]
        assert np.array_equal(a["t"], gold[k + "t"][lo:hi])
        assert np.abs(a["p"] - gold[k + "p"][lo:hi]).max(initial=0.0) <= tol
        assert a["L"] == gold[k + "L"][i] and np.array_equal(a["L_i"], gold[k + "Li"][i])
        assert a["converged"] == bool(gold[k + "conv"][i])
        assert (orc.validity_flags(rb, s, a) & 7) == int(gold[k + "flags"][i])
        sret = s[-1] if spec.get("enable_retraction") else 0.0
        assert np.array_equal(orc.home_lengths(rb, sret), gold[k + "home"][i])


@pytest.mark.skipif(not (ref.RefTendonRobot.available() and ref.RefTendonRobot.release_available()),
                    reason="oracle/_ref/libtendonrobot_ref*.so not built")
def test_reference_batch_loop_is_a_sound_timing_baseline(orc, wl):
    """bench.py's `reference_own_text` figure times trref_shape_batch (the reference's TendonRobot::shape in the
    OpenMP loop of apps/estimate_length_discretization.cpp:62-71).  Its results must be the reference's: node
    counts equal, tips identical between the -O2 build and the oracle, and within 1e-12 L for the -Ofast build
    (which may reassociate)."""
    spec = wl.robot_b(0.005)
    st = wl.sample_states(spec, 300, stream=77)
    rr = ref.RefTendonRobot(spec)
    want = orc.fk_batch(orc.robot(spec), st, 64)
    tips, npts = rr.shape_batch(st, nthreads=2, release=False)
    assert np.array_equal(npts, want["npts"]) and np.array_equal(tips, want["tip"])
    tips_r, npts_r = rr.shape_batch(st, nthreads=2, release=True)
    assert np.array_equal(npts_r, want["npts"]) and np.abs(tips_r - want["tip"]).max() < 1e-12 * spec["L"]


@pytest.mark.skipif(not ref.RefTendonRobot.available(), reason="oracle/_ref/libtendonrobot_ref.so not built")
@pytest.mark.parametrize("name", ["a005", "b003rot"])
def test_live_tip_control_jacobian(orc, wl, name):
    """tip_control::Jacobian (tip-control/tip_control.cpp:243-265), the reference's own text: forward
    differences from the caller's ps with a step that is a C `float` there.  The oracle's mode 0
    (IRT_JAC_FORWARD_FIXED) with delta = double(float(dist)) is bit-exact; with the unrounded double step it
    differs at the 1e-8 level, which is why the mirrors round the step the way the reference's signature does."""
    spec = {"a005": wl.robot_a(0.005), "b003rot": wl.robot_b(0.003, rotation=True)}[name]
    rb, rr = orc.robot(spec), ref.RefTendonRobot(spec)
    worst_unrounded = 0.0
    for st in wl.sample_states(spec, 12, stream=91):
        for dist in (1e-3, 1e-4, 0.01):
            d32 = float(np.float32(dist))
            tip, J = orc.tip_jacobian(rb, st, 0, d32)
            Jr = rr.tip_jacobian(st, tip, dist)
            assert np.array_equal(J, Jr), (name, dist)
            _, Ju = orc.tip_jacobian(rb, st, 0, dist)
            worst_unrounded = max(worst_unrounded, float(np.abs(Ju - Jr).max() / max(1e-300, np.abs(Jr).max())))
    assert 0 < worst_unrounded < 1e-5


@levmar
@pytest.mark.skipif(not ref.RefTendonRobot.ik_available(), reason="oracle/_ref/libik_ref.so not built")
@pytest.mark.parametrize("name", ["b005", "b005rot"])
def test_live_inverse_kinematics_own_text(orc, wl, name):
    """tip_control::inverse_kinematics compiled from the reference's OWN text (inverse_kinematics_impl with its
    levmar options, fk_wrap and its (0, 0, L - s) branch, Bounds::from_robot, canonical_angle; the robot's own
    forward_kinematics underneath) against the driver as the tests restate it (dlevmar_bc_dif over the ORACLE's
    FK with the same options and bounds): same iteration / FK-call counts; solution and levmar info bit for bit
    without rotation, to 1e-6 with it."""
    spec = wl.robot_b(0.005, rotation=(name == "b005rot"))
    rb, rr = orc.robot(spec), ref.RefTendonRobot(spec)
    L, N, S = spec["L"], 6, wl.state_size(spec)
    f = _fk_wrap(orc, rb, L)
    lb, ub = np.zeros(S), np.zeros(S)
    ub[:N] = 20.0
    if spec.get("enable_rotation"):
        lb[N], ub[N] = np.finfo(np.float64).min, np.finfo(np.float64).max      # Bounds::from_robot
    ub[-1] = L
    step = np.zeros(S)
    step[:N] = [0.8, -0.5, 0.3, 0.6, -0.4, 0.2]
    step[-1] = 0.004
    if spec.get("enable_rotation"):
        step[N] = 0.3
    for s0 in wl.sample_states(spec, 4, stream=35):
        des = f(np.clip(s0 + step, np.maximum(lb, -10), np.minimum(ub, 10)))
        p, info, rc = ref.RefLevmar.bc_dif(f, s0, des, lb, ub, 100, [0.1, 1e-9, 1e-8, 1e-8, -1e-6])
        got = rr.inverse_kinematics(s0, des, 100, 0.1, 1e-9, 1e-4, 1e-4, 1e-6)
        assert rc >= 0 and got["iters"] == int(info[5]) and got["num_fk_calls"] == int(info[7])
        want_state = p.copy()
        if not spec.get("enable_rotation"):
            assert np.array_equal(got["info"], info) and np.array_equal(got["state"], want_state)
            assert np.array_equal(got["tip"], orc.shape(rb, p)["p"][-1]) and got["error"] == np.sqrt(info[1])
        else:
            # rotate_z is the one place where the own text and the oracle differ (by <= 2e-16: Eigen's AngleAxis
            # forms the zz entry as (1 - c) + c); a hundred LM iterations carry that to ~1e-8
            want_state[N] = (want_state[N] + np.pi) % (2 * np.pi) - np.pi          # util::canonical_angle
            assert np.allclose(got["info"], info, rtol=1e-6, atol=0) and np.allclose(got["state"], want_state, rtol=1e-6, atol=1e-9)
            assert np.allclose(got["tip"], orc.shape(rb, p)["p"][-1], rtol=0, atol=1e-9 * L)
        assert info[1] < info[0] and got["iters"] > 1      # it did iterate and the tip error went down
