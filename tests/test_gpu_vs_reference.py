"""GPU parity against the REFERENCE'S OWN CODE (oracle/_ref, prebuilt where /root/reference exists
and shipped with the repo snapshot; see oracle/ref.py for exactly what it is).

* K3 verdicts vs the reference's real octree: every GPU-built set is rebuilt as a
  collision::detail::TreeNode<128> and tested with TreeNode::collides (TreeNode.hxx:165-174, :268) --
  the loop of VoxelCachedLazyPRM.cpp:1584-1591.  Bit-exact, 0 flips.
* K1 backbone points vs an RK4 walk whose derivative and initial condition are the reference's
  unmodified tendon_deriv.cpp / solve_initial_bending.cpp / get_r_info.cpp: 1e-9 L.
"""
import numpy as np
import pytest

from oracle import ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref.available(), reason="oracle/_ref was not shipped")]

FK_REL_TOL = 1e-9


@pytest.fixture(scope="module")
def irt():
    import irt_b200
    return irt_b200


@pytest.fixture(scope="module")
def ctx(irt):
    return irt.Context(0)


def _ref_env(wl, Ng, env_blocks):
    Nb = Ng // 4
    keys = np.nonzero(env_blocks)[0].astype(np.uint32)
    ex, ey, ez = wl.morton_decode(keys, Nb)
    renv = ref.RefTree(Ng)
    for x, y, z, k in zip(ex.tolist(), ey.tolist(), ez.tolist(), keys.tolist()):
        renv.set_block(x, y, z, int(env_blocks[k]))
    return renv


def test_k3_verdicts_vs_reference_treenode(irt, ctx, wl):
    spec = wl.robot_b(0.003)
    g = wl.workspace_grid(spec)
    Nb = g["Ng"] // 4
    grid = irt.make_grid(g["Ng"], g["lim"])
    rb = irt.Robot(ctx, spec)
    states = wl.sample_states(spec, 20000, stream=71)
    store = irt.SetStore(ctx, grid)
    store.voxelize_vertices(rb, states)
    env_blocks = wl.dense_to_morton_blocks(wl.lung_like_env_dense(spec, g))
    env = irt.Env(ctx, grid)
    env.update(env_blocks)
    got = store.check(env)
    off, keys, bits = store.export_csr()
    bx, by, bz = wl.morton_decode(keys, Nb)
    want = _ref_env(wl, g["Ng"], env_blocks).check_csr(off, bx, by, bz, bits)
    flips = int(np.count_nonzero(got != want))
    assert flips == 0, "%d verdict flips vs TreeNode::collides" % flips
    assert 0.02 < want.mean() < 0.98


def test_k3_edge_verdicts_vs_reference_treenode(irt, ctx, wl):
    """swept-volume sets of roadmap edges (K2 output) against the reference octree, incl. a tick
    of the replanning loop (changed environment)"""
    spec = wl.robot_b(0.003)
    g = wl.workspace_grid(spec)
    Nb = g["Ng"] // 4
    grid = irt.make_grid(g["Ng"], g["lim"])
    rb = irt.Robot(ctx, spec)
    a = wl.sample_states(spec, 1500, stream=72)
    b = a + 0.05 * (wl.sample_states(spec, 1500, stream=73) - a)
    store = irt.SetStore(ctx, grid)
    store.voxelize_edges(rb, irt.make_space(), a, b)
    off, keys, bits = store.export_csr()
    bx, by, bz = wl.morton_decode(keys, Nb)
    env_blocks = wl.dense_to_morton_blocks(wl.lung_like_env_dense(spec, g))
    env = irt.Env(ctx, grid)
    for tick in range(3):
        if tick:
            rng = np.random.default_rng(tick)
            env_blocks = wl.toggle_blob(env_blocks, g, rng.uniform(-0.1, 0.1, 3), 0.012)
        env.update(env_blocks)
        got = store.check(env)
        want = _ref_env(wl, g["Ng"], env_blocks).check_csr(off, bx, by, bz, bits)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("robot,dL,rot", [("a", 0.005, False), ("b", 0.003, False), ("b", 0.005, True)])
def test_k1_points_vs_reference_derivative(irt, ctx, wl, robot, dL, rot):
    spec = wl.robot_a(dL) if robot == "a" else wl.robot_b(dL, rotation=rot)
    rb = irt.Robot(ctx, spec)
    rf = ref.RefFK(spec)
    N = rf.N
    states = wl.sample_states(spec, 200, stream=74)
    out = rb.shape_batch(states, want=("p", "t", "npts", "L", "L_i", "uv", "iters", "nsteps", "flags"))
    worst = 0.0
    for i, s in enumerate(states):
        n = int(out["npts"][i])
        if n < 2:
            continue
        st, nsteps = rf.shape_states(s[:N], out["t"][i, :n])
        assert nsteps == out["nsteps"][i]
        p = st[:, :3]
        if rot:
            c, sn = np.cos(s[N]), np.sin(s[N])
            p = p @ np.array([[c, -sn, 0], [sn, c, 0], [0, 0, 1]]).T
        worst = max(worst, np.abs(p - out["p"][i, :n]).max() / spec["L"])
        assert abs(st[-1, 18] - out["L"][i]) < FK_REL_TOL * spec["L"]
        assert np.abs(st[-1, 19:] - out["L_i"][i]).max() < FK_REL_TOL * spec["L"]
        s0 = s[-1] if spec.get("enable_retraction") else 0.0
        v0, u0, it = rf.initial_bending(s[:N], s0)
        assert it == out["iters"][i]
        assert np.allclose(out["uv"][i][0:3], u0, rtol=1e-9, atol=1e-10)
        assert np.allclose(out["uv"][i][6:9], v0, rtol=1e-9, atol=1e-10)
    assert worst < FK_REL_TOL, worst


@pytest.mark.skipif(not ref.RefLevmar.available(), reason="oracle/_ref/liblevmar_ref.so was not shipped")
def test_ik_levmar_driven_by_gpu_jacobians(irt, ctx, orc, wl):
    """The IK of the reference end to end ("IK"): its own optimiser, levmar-2.6's
    box-constrained LM, once as the reference runs it -- dlevmar_bc_dif differentiating a CPU FK
    sequentially (tip_control.cpp:124-137) -- and once as dlevmar_bc_der with the GPU supplying tips and
    whole central-difference Jacobians per call (one K1 launch instead of 2S+1 sequential FKs).
    dlevmar_bc_dif IS dlevmar_bc_der with finite-difference wrappers, so the iterates agree."""
    spec = wl.robot_b(0.005)
    rb = irt.Robot(ctx, spec)
    orb = orc.robot(spec)
    L = spec["L"]
    delta = 1e-6

    def f_cpu(p):
        if p[-1] > L:
            return np.array([0.0, 0.0, L - p[-1]])
        return orc.shape(orb, p)["p"][-1]

    def f_gpu(p):
        return rb.tip_jacobian_batch(p[None], mode=irt.JAC_LEVMAR_CENTRAL, delta=delta)[0][0]

    def j_gpu(p):
        return rb.tip_jacobian_batch(p[None], mode=irt.JAC_LEVMAR_CENTRAL, delta=delta)[1][0]

    lb, ub = np.zeros(7), np.array([20.0] * 6 + [L])
    st = wl.sample_states(spec, 5, stream=34)
    for s in st:
        goal_state = np.clip(s + np.array([0.8, -0.5, 0.3, 0.6, -0.4, 0.2, 0.004]), lb, ub)
        des = f_cpu(goal_state)
        p_ref, info_ref, _ = ref.RefLevmar.bc_dif(f_cpu, s, des, lb, ub, 100, [0.1, 1e-9, 1e-8, 1e-8, -delta])
        p_gpu, info_gpu, _ = ref.RefLevmar.bc_der(f_gpu, j_gpu, s, des, lb, ub, 100, [0.1, 1e-9, 1e-8, 1e-8])
        assert info_gpu[5] == info_ref[5] and info_gpu[6] == info_ref[6]      # iterations, stop reason
        assert np.abs(p_gpu - p_ref).max() < 1e-6 * 20.0
        assert np.linalg.norm(f_cpu(p_gpu) - des) < 2e-4


vo_ref = pytest.mark.skipif(not ref.RefVoxelOctree.available(),
                            reason="oracle/_ref/libvoxeloctree_ref.so was not shipped")


@vo_ref
@pytest.mark.parametrize("robot", ["a", "b"])
def test_k2_vertex_sets_vs_reference_add_line(irt, ctx, wl, robot):
    """K2 vertex voxel sets vs VoxelOctree::add_piecewise_line -- the reference's own add_line text
    (VoxelOctree.cpp:325-432) -- applied to the points K1 produced: bit-exact CSR, visit_leaves order."""
    spec = wl.robot_a(0.003) if robot == "a" else wl.robot_b(0.003)
    g = wl.workspace_grid(spec)
    Nb = g["Ng"] // 4
    rb = irt.Robot(ctx, spec)
    states = wl.sample_states(spec, 1500, stream=81)
    fk = rb.shape_batch(states, want=("p", "npts", "flags"))
    store = irt.SetStore(ctx, irt.make_grid(g["Ng"], g["lim"]))
    flags, _ = store.voxelize_vertices(rb, states)
    off, keys, bits = store.export_csr()
    flips = 0
    for i in range(len(states)):
        t = ref.RefVoxelOctree(g["Ng"], g["lim"])
        if flags[i] == 0:   # a vertex the reference would never keep (invalid shape) has an empty set
            t.add_piecewise_line(fk["p"][i, :fk["npts"][i]])
        xyz, rbits = t.export()
        rkeys = wl.morton_key(xyz[:, 0].astype(np.int64), xyz[:, 1].astype(np.int64), xyz[:, 2].astype(np.int64), Nb) \
            if len(xyz) else np.zeros(0, dtype=np.uint32)
        gk, gb = keys[int(off[i]):int(off[i + 1])], bits[int(off[i]):int(off[i + 1])]
        if not (np.array_equal(gk, np.asarray(rkeys, dtype=np.uint32)) and np.array_equal(gb, rbits)):
            flips += 1
    assert flips == 0, "%d vertex sets differ from the reference's add_piecewise_line" % flips


@vo_ref
def test_env_preparation_vs_reference(irt, ctx, wl):
    """irt_env_dilate / dilate_sphere / remove_interior vs the reference's own tree walks
    (VoxelOctree.cpp:533-952) on the C4/C5 environment"""
    spec = wl.robot_b(0.003)
    g = wl.workspace_grid(spec)
    Nb = g["Ng"] // 4
    grid = irt.make_grid(g["Ng"], g["lim"])
    blocks = wl.dense_to_morton_blocks(wl.lung_like_env_dense(spec, g))
    keys = np.nonzero(blocks)[0].astype(np.uint32)
    ex, ey, ez = wl.morton_decode(keys, Nb)
    for op in (("dilate", 1, False), ("dilate", 2, True), ("dilate_sphere", 0.01),
               ("remove_interior", True), ("remove_interior", False)):
        env = irt.Env(ctx, grid)
        env.update(blocks)
        tr = ref.RefVoxelOctree(g["Ng"], g["lim"])
        for x, y, z, k in zip(ex.tolist(), ey.tolist(), ez.tolist(), keys.tolist()):
            tr.set_block(x, y, z, int(blocks[k]))
        getattr(env, op[0])(*op[1:])
        getattr(tr, op[0])(*op[1:])
        got = env.download()
        xyz, rbits = tr.export()
        want = np.zeros_like(got)
        want[wl.morton_key(xyz[:, 0].astype(np.int64), xyz[:, 1].astype(np.int64), xyz[:, 2].astype(np.int64), Nb)] = rbits
        assert np.array_equal(got, want), op


@vo_ref
@pytest.mark.skipif(not ref.RefVoxelOctree.has_primitives(), reason="libvoxeloctree_ref.so built without primitives")
def test_env_voxelize_vs_reference(irt, ctx, wl):
    """irt_env_add_primitives (Environment::voxelize, Environment.cpp:62-74) vs the reference's own
    add(Point) / add_sphere / add_capsule text (VoxelOctree.cpp:319-323, 434-515): the benchmark's capsule tree
    plus points and spheres inside, across and outside the grid; 0 flips"""
    spec = wl.robot_b(0.003)
    g = wl.workspace_grid(spec)
    Nb = g["Ng"] // 4
    rng = np.random.default_rng(99)
    caps = [np.concatenate([a, b, [r]]) for a, b, r in wl.lung_like_capsules(spec)]
    pts = rng.uniform(-0.25, 0.25, (40, 3))
    sph = np.concatenate([rng.uniform(-0.25, 0.25, (30, 3)), rng.choice([0.0005, 0.004, 0.02, 0.06], (30, 1))], axis=1)
    tr = ref.RefVoxelOctree(g["Ng"], g["lim"])
    for p in pts:
        tr.add_point(p)
    for s_ in sph:
        tr.add_sphere(s_[:3], float(s_[3]))
    for c in caps:
        tr.add_capsule(c[:3], c[3:6], float(c[6]))
    env = irt.Env(ctx, irt.make_grid(g["Ng"], g["lim"]))
    env.add_primitives(pts, sph, np.array(caps), clear=True)
    got = env.download()
    xyz, rbits = tr.export()
    want = np.zeros_like(got)
    want[wl.morton_key(xyz[:, 0].astype(np.int64), xyz[:, 1].astype(np.int64), xyz[:, 2].astype(np.int64), Nb)] = rbits
    assert np.array_equal(got, want) and np.count_nonzero(want) > 500


@pytest.mark.skipif(not ref.RefSelfCollision.available(), reason="oracle/_ref/libselfcol_ref.so was not shipped")
def test_self_collision_flags_vs_reference(irt, ctx, wl):
    """IRT_FLAG_SELF_COLLISION of the validity epilogue (FP32 pair filter + exact kernel) vs the
    reference's own collides_self text (collision.cpp:6-46) on the points K1 produced."""
    n = hit = 0
    for base in (wl.robot_a(0.005), wl.robot_b(0.003)):
        for E in (0.5e6, 2.1e6):
            spec = dict(base)
            spec["E"] = E
            rb = irt.Robot(ctx, spec)
            st = wl.sample_states(spec, 600, stream=14)
            out = rb.shape_batch(st, want=("p", "npts", "flags"))
            for i in range(len(st)):
                want = ref.RefSelfCollision.collides_self(out["p"][i, :out["npts"][i]], spec["r"])
                assert bool(out["flags"][i] & irt.FLAG_SELF_COLLISION) == want, (E, i)
                n += 1
                hit += want
    assert 0.1 * n < hit < 0.9 * n


@pytest.mark.skipif(not ref.RefTendonRobot.available(), reason="oracle/_ref/libtendonrobot_ref.so was not shipped")
@pytest.mark.parametrize("name", ["a005", "b003", "b005rot", "a003soft", "b005tight"])
def test_k1_vs_reference_tendon_robot_shape(irt, ctx, wl, name):
    """K1 + validity epilogue vs the reference's own TendonRobot::shape text (tension_shape, home_shape,
    calc_point_forces, collides_self; Eigen and Boost.odeint stand-ins): node counts and grids exact, points /
    lengths within 1e-9 L, NONCONVERGED / LENGTH_LIMIT / SELF_COLLISION flags identical."""
    spec = {"a005": wl.robot_a(0.005), "b003": wl.robot_b(0.003), "b005rot": wl.robot_b(0.005, rotation=True),
            "a003soft": dict(wl.robot_a(0.003), E=0.7e6),
            "b005tight": dict(wl.robot_b(0.005), residual_threshold=1e-13)}[name]
    rb, rr = irt.Robot(ctx, spec), ref.RefTendonRobot(spec)
    L = spec["L"]
    st = wl.sample_states(spec, 400, stream=23)
    if spec.get("enable_retraction"):
        st[0, -1], st[1, -1], st[2, -1], st[3, -1], st[4, -1] = L, L + 0.01, 0.1995, 0.0, -0.001
    out = rb.shape_batch(st, want=("p", "R", "t", "npts", "L", "L_i", "uv", "flags"))
    home = rb.home_lengths(st)
    seen = 0
    for i, s in enumerate(st):
        b = rr.shape(s)
        n = len(b["t"])
        assert out["npts"][i] == n
        assert np.allclose(out["t"][i, :n], b["t"], rtol=0, atol=1e-15)
        assert np.abs(out["p"][i, :n] - b["p"]).max(initial=0.0) < FK_REL_TOL * L
        assert np.abs(out["R"][i, :n] - b["R"]).max(initial=0.0) < 1e-9
        assert abs(out["L"][i] - b["L"]) < FK_REL_TOL * L
        assert np.abs(out["L_i"][i] - b["L_i"]).max() < FK_REL_TOL * L
        fr = rr.flags(s)
        assert (int(out["flags"][i]) & 7) == fr, (i, int(out["flags"][i]), fr)
        seen |= fr
        assert np.abs(home[i] - rr.home_lengths(s)).max() < 1e-15
    if name == "b005tight":
        assert seen & 1
    if name == "a003soft":
        assert seen & 2 and seen & 4


@pytest.mark.skipif(not ref.RefLevmar.available(), reason="oracle/_ref/liblevmar_ref.so was not shipped")
def test_roadmap_ik_lockstep_with_the_reference_levmar(irt, ctx, orc, wl):
    """roadmapIk as a batch (VoxelCachedLazyPRM.cpp:3095-3205): the k seeds' IK problems run the reference's own
    optimiser (levmar-2.6 dlevmar_bc_der, one host thread per seed) while every FK / Jacobian request of the k
    solvers is answered in lockstep by single K1 launches; results and the order of acceptance equal the
    reference's sequential loop (dlevmar_bc_dif over a CPU FK, neighbour by neighbour)."""
    from irt_b200 import roadmap as R
    spec = wl.robot_b(0.003)       # voxel caches at 128^3 need dL <= the voxel size
    g = wl.workspace_grid(spec)
    grid = irt.make_grid(g["Ng"], g["lim"], g["inv_rot"])
    rb = irt.Robot(ctx, spec)
    orb = orc.robot(spec)
    L = spec["L"]
    delta = 1e-6
    prm = R.VoxelCachedLazyPRM(ctx, rb, grid)
    prm.createRoadmap(300, lambda cnt, rnd: wl.sample_states(spec, cnt, stream=950 + rnd),
                      lambda st: wl.knn_edges(spec, st, k=4), opt=R.VoxelizeVertices)
    env = irt.Env(ctx, grid)
    env.add_primitives(spheres=[[0.06, 0.0, 0.12, 0.025]], clear=True)
    prm.setEnvironment(env.download())
    ogrid = orc.grid(g["Ng"], g["lim"])
    oenv = orc.octree(ogrid)
    oenv.add_sphere([0.06, 0.0, 0.12], 0.025)
    lb, ub = np.zeros(7), np.array([20.0] * 6 + [L])
    opts = [0.1, 1e-9, 1e-8, 1e-8]

    def f_cpu(p):
        if p[-1] > L:
            return np.array([0.0, 0.0, L - p[-1]])
        return orc.shape(orb, p)["p"][-1]

    def solver(start, request, fk):
        last = {}

        def both(p):
            key = p.tobytes()
            if last.get("k") != key:
                last["k"], last["v"] = key, fk(p)
            return last["v"]

        p, info, _ = ref.RefLevmar.bc_der(lambda q: both(q)[0], lambda q: both(q)[1], start, request, lb, ub, 100, opts)
        return p

    rng = np.random.default_rng(8)
    accepted = 0
    for trial in range(3):
        goal = np.clip(prm.states[rng.integers(300)] + rng.normal(size=7) * [0.8, 0.8, 0.8, 0.8, 0.8, 0.8, 0.003], lb, ub)
        request = f_cpu(goal)
        got = prm.roadmapIk(request, 5e-4, 4, solver, delta=delta)
        assert got["lockstep_batches"] < got["fk_requests"]
        want = None
        for i, v in enumerate(got["neighbors"]):
            p_ref, info_ref, _ = ref.RefLevmar.bc_dif(f_cpu, prm.states[v].copy(), request, lb, ub, 100, opts + [-delta])
            st_, fl_ = orc.voxelize_vertices_batch(orb, ogrid, p_ref[None])
            ok = fl_[0] == 0 and not orc.check_sets_batch(st_, oenv)[0]
            err = float(np.linalg.norm(f_cpu(p_ref) - request))
            assert bool(got["valid"][i]) == bool(ok)
            assert abs(got["errors"][i] - err) < 2e-5
            if want is None and ok and err < 5e-4:
                want = (i, p_ref)
        if want is not None:
            accepted += 1
            assert got["index"] == want[0] and np.abs(got["controls"] - want[1]).max() < 1e-6 * 20.0
    assert accepted >= 1
