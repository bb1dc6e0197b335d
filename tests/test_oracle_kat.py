"""Pins for the CPU oracle (oracle/tendon_oracle.cpp).

The reference ships no tests or golden vectors, so the oracle is pinned by
analytic known-answer tests and by an independent numpy/mpmath restatement
(oracle/fk_second_opinion.py).  No GPU needed.
"""
import math

import numpy as np
import pytest

from oracle import fk_second_opinion as so


def test_t_range_grid(orc):
    # TendonRobot.cpp:69-84: anchored at L, only the first gap is irregular
    t = orc.t_range(0.0, 0.2, 0.005)
    assert len(t) == 41 and t[0] == 0.0 and t[-1] == 0.2
    assert np.allclose(np.diff(t), 0.005, atol=1e-15)
    t = orc.t_range(0.0, 0.2, 0.003)
    assert len(t) == 68
    assert abs((t[1] - t[0]) - 0.002) < 1e-12
    assert np.allclose(np.diff(t)[1:], 0.003, atol=1e-15)
    t = orc.t_range(0.0131, 0.2, 0.005)
    gaps = np.diff(t)
    assert np.allclose(gaps[1:], 0.005, atol=1e-15) and 0.0025 < gaps[0] <= 0.0075 + 1e-15
    # L - dL/2 < s < L: single point
    assert len(orc.t_range(0.199, 0.2, 0.005)) == 1


def test_kat_zero_tension_is_home(orc, wl):
    # (i) tau = 0 -> p = (0,0,t-s), R = I, L_i = L - s
    rb = orc.robot(wl.robot_a())
    s = orc.shape(rb, [0, 0, 0, 0])
    assert s["converged"] and len(s["t"]) == 41
    assert np.allclose(s["p"][:, :2], 0, atol=0) and np.allclose(s["p"][:, 2], s["t"], atol=1e-15)
    assert np.allclose(s["R"], np.eye(3)[None], atol=1e-15)
    assert abs(s["L"] - 0.2) < 1e-14 and np.allclose(s["L_i"], 0.2, atol=1e-14)
    assert np.allclose(orc.home_lengths(rb, 0.0), 0.2)
    rbb = orc.robot(wl.robot_b())
    sb = orc.shape(rbb, [0] * 6 + [0.05])
    assert np.allclose(sb["p"][:, 2], sb["t"] - 0.05, atol=1e-15)
    # helix home length (TendonRobot.cpp:289-295)
    want = (0.2 - 0.05) * math.sqrt(1 + 0.01 ** 2 * (2 * math.pi / 0.2) ** 2)
    assert np.allclose(orc.home_lengths(rbb, 0.05), want, rtol=1e-15)
    assert np.allclose(sb["L_i"], want, rtol=1e-9)


def test_kat_single_tendon_circular_arc(orc, wl):
    # (ii) one straight tendon: u, v constant along s -> planar circular arc R(s) = exp(s u^)
    spec = wl.robot_a()
    rb = orc.robot(spec)
    s = orc.shape(rb, [5.0, 0, 0, 0])
    assert s["converged"]
    u0, v0 = s["u_i"], s["v_i"]
    assert np.allclose(s["u_f"], u0, rtol=1e-9, atol=1e-12) and np.allclose(s["v_f"], v0, rtol=1e-9)
    # closed form: theta = |u| t ; p(t) = integral R(t) v dt with constant body-frame v, u
    k = np.linalg.norm(u0)
    ax = u0 / k
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    for i in (10, 25, 40):
        t = s["t"][i]
        Rt = np.eye(3) + math.sin(k * t) * K + (1 - math.cos(k * t)) * K @ K
        # integral of exp(tau K k) dtau = I t + (1-cos)/k K + (t - sin/k) K^2
        P = np.eye(3) * t + (1 - math.cos(k * t)) / k * K + (t - math.sin(k * t) / k) * K @ K
        assert np.allclose(s["R"][i], Rt, atol=1e-9)
        assert np.allclose(s["p"][i], P @ v0, atol=1e-9)


def test_kat_opposite_tendons_straight_compressed(orc, wl):
    # (iii) equal tension on opposite straight tendons: straight, compressed, equal L_i
    rb = orc.robot(wl.robot_a())
    s = orc.shape(rb, [8.0, 0, 8.0, 0])
    assert np.allclose(s["p"][:, :2], 0, atol=1e-12)
    assert s["v_i"][2] < 1.0 and s["p"][-1, 2] < 0.2
    assert abs(s["L_i"][0] - s["L_i"][2]) < 1e-12


def test_kat_routing_rotation_convention(orc, wl):
    # (iv) x = rho sin(theta), y = rho cos(theta): adding phi to every theta0 rotates the shape
    # about z by -phi
    spec = wl.robot_b()
    spec["enable_retraction"] = False
    phi = 0.7
    spec2 = dict(spec)
    spec2["C"] = [[c[0] + phi, c[1]] for c in spec["C"]]
    tau = [3.0, 7.5, 1.0, 0.0, 12.0, 4.0]
    a = orc.shape(orc.robot(spec), tau)
    b = orc.shape(orc.robot(spec2), tau)
    c, sn = math.cos(-phi), math.sin(-phi)
    Rz = np.array([[c, -sn, 0], [sn, c, 0], [0, 0, 1]])
    assert np.allclose(b["p"], a["p"] @ Rz.T, atol=1e-10)
    r, _, _ = orc.routing(orc.robot(spec), 0.0)
    assert np.allclose(r[0], [0.0, 0.01, 0.0], atol=1e-18)  # theta = 0 -> +y


def test_kat_rotation_control_equals_rotate_z(orc, wl):
    # (v) shape(tau, rot) == Rz(rot) shape(tau, 0)   (TendonResult.cpp:13-18)
    spec = wl.robot_b(rotation=True)
    rb = orc.robot(spec)
    tau = [3.0, 7.5, 1.0, 0.0, 12.0, 4.0]
    a = orc.shape(rb, tau + [0.0, 0.03])
    b = orc.shape(rb, tau + [1.1, 0.03])
    c, sn = math.cos(1.1), math.sin(1.1)
    Rz = np.array([[c, -sn, 0], [sn, c, 0], [0, 0, 1]])
    assert np.allclose(b["p"], a["p"] @ Rz.T, atol=1e-15)
    assert np.allclose(b["R"], Rz[None] @ a["R"], atol=1e-15)


def test_kat_retraction_edges(orc, wl):
    # (vi) s >= L -> one point at the origin (TendonRobot.cpp:361-372); first-gap rule
    rb = orc.robot(wl.robot_b())
    for s_start in (0.2, 0.25):
        s = orc.shape(rb, [1, 2, 3, 4, 5, 6, s_start])
        assert len(s["t"]) == 1 and np.all(s["p"] == 0) and s["L"] == 0 and s["converged"]
        assert np.all(s["L_i"] == 0) and s["nsteps"] == 0
    s = orc.shape(rb, [1, 2, 3, 4, 5, 6, 0.0131])
    # gap 0.0019 + ... : 0.2 - 0.0131 = 0.1869 = 37*0.005 + 0.0019 -> first gap 0.0069 > dL: 2 steps
    assert len(s["t"]) == 38 and s["nsteps"] == 38
    s = orc.shape(rb, [1, 2, 3, 4, 5, 6, 0.0169])  # first gap 0.0031 < dL: one step
    assert len(s["t"]) == 38 and s["nsteps"] == 37
    s = orc.shape(rb, [1, 2, 3, 4, 5, 6, 0.199])   # K == 0: the grid is {s}
    assert len(s["t"]) == 1 and s["nsteps"] == 0 and s["L"] == 0.0


def test_kat_blockwise_vs_dense_solve(orc, wl):
    # (vii) linsubsolve2 (block inverse) vs a dense 6x6 solve agree to 1e-10
    spec = wl.robot_b()
    rb = orc.robot(spec)
    rng = np.random.default_rng(5)
    for _ in range(20):
        tau = rng.uniform(0, 20, 6)
        x = np.zeros(25)
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        x[3:12] = q.T.reshape(-1)
        x[12:15] = [rng.normal() * 0.01, rng.normal() * 0.01, 1 + rng.normal() * 0.01]
        x[15:18] = rng.normal(size=3) * 3
        t = rng.uniform(0, 0.2)
        d1, d2 = orc.deriv(rb, tau, x, t), orc.deriv(rb, tau, x, t, alt=True)
        assert np.allclose(d1, d2, rtol=1e-10, atol=1e-10)


def test_kat_rk4_order_four(orc, wl):
    # (viii) halving dL divides the tip error by ~16
    tau = [3.0, 7.5, 1.0, 0.0, 12.0, 4.0]
    tips = {}
    for dL in (0.01, 0.005, 0.0025, 0.000625):
        spec = wl.robot_b(dL)
        spec["enable_retraction"] = False
        tips[dL] = orc.shape(orc.robot(spec), tau)["p"][-1]
    e1 = np.linalg.norm(tips[0.01] - tips[0.000625])
    e2 = np.linalg.norm(tips[0.005] - tips[0.000625])
    e3 = np.linalg.norm(tips[0.0025] - tips[0.000625])
    assert 10 < e1 / e2 < 24 and 10 < e2 / e3 < 24


def _cmp_second_opinion(orc, spec, state, mp=None, tol=1e-11):
    rb = orc.robot(spec)
    a = orc.shape(rb, state)
    # impose the oracle's initial condition so the comparison is about the ODE + integrator
    b = so.fk(spec, list(state), mp=mp, v0u0=(a["v_i"], a["u_i"]))
    assert len(b["t"]) == len(a["t"])
    assert np.allclose(np.array([float(x) for x in b["t"]]), a["t"], atol=1e-15)
    p = np.array([[float(c) for c in q] for q in b["p"]])
    assert np.abs(p - a["p"]).max() < tol * spec["L"]
    Li = np.array([float(x) for x in b["L_i"]])
    assert np.abs(Li - a["L_i"]).max() < tol
    return a, b


def test_cross_numpy_second_opinion(orc, wl):
    for spec_fn, n in ((wl.robot_a, 12), (wl.robot_b, 25)):
        spec = spec_fn()
        for st in wl.sample_states(spec, n, stream=11):
            _cmp_second_opinion(orc, spec, st)
    spec = wl.robot_b(rotation=True)
    for st in wl.sample_states(spec, 6, stream=12):
        _cmp_second_opinion(orc, spec, st)


def test_cross_initial_bending_second_opinion(orc, wl):
    # fixed-point iteration: same iterate count and v0,u0 as the independent restatement
    spec = wl.robot_b()
    rb = orc.robot(spec)
    for st in wl.sample_states(spec, 20, stream=13):
        a = orc.shape(rb, st)
        b = so.fk(spec, list(st))
        assert b["iters"] == a["iters"]
        assert np.allclose([float(x) for x in b["v0"]], a["v_i"], rtol=1e-12, atol=1e-14)
        assert np.allclose([float(x) for x in b["u0"]], a["u_i"], rtol=1e-12, atol=1e-12)


def test_cross_mpmath_50_digits(orc, wl):
    mp = pytest.importorskip("mpmath").mp
    mp.dps = 50
    spec = wl.robot_b()
    for st in wl.sample_states(spec, 3, stream=14):
        _cmp_second_opinion(orc, spec, st, mp=mp, tol=2e-12)


# ---------------------------------------------------------------------------------------------
# validity epilogue
# ---------------------------------------------------------------------------------------------
def test_closest_st_segment_cases(orc):
    # crossing segments
    s, t = orc.closest_st_segment([0, 0, 0], [1, 0, 0], [0.5, -1, 1], [0.5, 1, 1])
    assert abs(s - 0.5) < 1e-15 and abs(t - 0.5) < 1e-15
    # degenerate first segment
    s, t = orc.closest_st_segment([0, 0, 0], [0, 0, 0], [1, -1, 0], [1, 3, 0])
    assert s == 0.0 and abs(t - 0.25) < 1e-15
    # parallel, overlapping -> (0, t)
    s, t = orc.closest_st_segment([0, 0, 0], [1, 0, 0], [-1, 1, 0], [3, 1, 0])
    assert s == 0.0 and abs(t - 0.25) < 1e-15
    # parallel, disjoint -> closest endpoints (1,0)
    s, t = orc.closest_st_segment([0, 0, 0], [1, 0, 0], [2, 1, 0], [3, 1, 0])
    assert (s, t) == (1.0, 0.0)
    # t clamped below 0
    s, t = orc.closest_st_segment([0, 0, 0], [1, 0, 0], [0.3, 1, 0], [0.3, 2, 0])
    assert t == 0.0 and 0.0 <= s <= 1.0


def test_collides_self(orc):
    r = 0.015
    line = np.stack([np.zeros(41), np.zeros(41), np.linspace(0, 0.2, 41)], axis=1)
    assert not orc.collides_self(line, r)
    assert not orc.collides_self(line[:2], r)  # N <= 2 (collision.cpp:14)
    # a loop that comes back onto itself: circle of circumference 0.2 -> start meets end
    th = np.linspace(0, 2 * math.pi, 41)
    rad = 0.2 / (2 * math.pi)
    circ = np.stack([rad * (1 - np.cos(th)), np.zeros(41), rad * np.sin(th)], axis=1)
    assert orc.collides_self(circ, r)
    # half circle of the same length: ends are 2*0.2/pi = 0.127 apart -> free
    th = np.linspace(0, math.pi, 41)
    rad = 0.2 / math.pi
    half = np.stack([rad * (1 - np.cos(th)), np.zeros(41), rad * np.sin(th)], axis=1)
    assert not orc.collides_self(half, r)


def test_validity_flags_length_limits(orc, wl):
    spec = wl.robot_a()
    rb = orc.robot(spec)
    from oracle.oracle import FLAG_LENGTH_LIMIT
    s = orc.shape(rb, [20.0, 0, 0, 0])  # strong single pull: opposite tendon lengthens a lot
    f = orc.validity_flags(rb, [20.0, 0, 0, 0], s)
    dl = orc.home_lengths(rb, 0.0) - s["L_i"]
    expect = np.any((dl < -0.015) | (dl > 0.035))
    assert bool(f & FLAG_LENGTH_LIMIT) == bool(expect)
    spec2 = dict(spec)
    spec2["max_length"] = [1e-4] * 4
    rb2 = orc.robot(spec2)
    s2 = orc.shape(rb2, [5.0, 0, 0, 0])
    assert orc.validity_flags(rb2, [5.0, 0, 0, 0], s2) & FLAG_LENGTH_LIMIT


# ---------------------------------------------------------------------------------------------
# voxels
# ---------------------------------------------------------------------------------------------
def _grid(orc, Ng=16, lim=(0, 1.6, 0, 1.6, 0, 1.6)):
    return orc.grid(Ng, lim)


def test_bitmask_and_child_order(orc):
    g = _grid(orc, 16)
    t = orc.octree(g)
    # bit = x*16 + y*4 + z (VoxelOctree.cpp:1501-1503); cell (5, 2, 7) -> block (1,0,1), local (1,2,3)
    t.add_line([0.55, 0.25, 0.75], [0.55, 0.25, 0.75])
    # hand trace of the reference's quirks for a zero-length segment: U = 0 -> every step = +1;
    # e = |A - (Ai + 1) * d| = (4.9, 2.2, 6.7) (voxel coordinate minus a METRIC length,
    # VoxelOctree.cpp:371-373), all deltas 1e10 -> y is the smallest t -> the walk steps to
    # (5, 3, 7) and marks it BEFORE the loop condition fails (:423-424).
    assert t.block(1, 0, 1) == (1 << (1 * 16 + 2 * 4 + 3)) | (1 << (1 * 16 + 3 * 4 + 3))
    assert t.nblocks() == 1 and t.ncells() == 2
    # visit order = child index bz/c + 2 by/c + 4 bx/c recursively -> morton with x most significant
    t2 = orc.octree(g)
    coords = [(3, 3, 3), (0, 0, 1), (0, 1, 0), (1, 0, 0), (2, 0, 0), (0, 0, 2), (0, 2, 0)]
    for c in coords:
        t2.set_block(*c, 1)
    bxyz, _ = t2.export()
    keys = [orc.morton_key(int(x), int(y), int(z), 4) for x, y, z in bxyz]
    assert keys == sorted(keys)
    assert [tuple(b) for b in bxyz.tolist()] == [(0, 0, 1), (0, 1, 0), (1, 0, 0), (0, 0, 2), (0, 2, 0), (2, 0, 0), (3, 3, 3)]
    assert orc.morton_key(1, 0, 0, 4) == 4 and orc.morton_key(0, 1, 0, 4) == 2 and orc.morton_key(0, 0, 1, 4) == 1
    assert orc.morton_key(2, 0, 0, 4) == 32


def test_add_line_hand_traced(orc):
    g = _grid(orc, 16)  # cell size 0.1
    t = orc.octree(g)
    t.add_line([0.05, 0.05, 0.05], [0.45, 0.05, 0.05])  # axis aligned, 5 cells
    cells = t.cells()
    assert {(i, 0, 0) for i in range(5)} <= cells
    # the reference's traversal marks the cell one step beyond B before re-testing (:423-424)
    assert cells - {(i, 0, 0) for i in range(6)} == set() or len(cells) <= 7
    t = orc.octree(g)
    t.add_line([0.35, 0.35, 0.35], [0.35, 0.35, 0.35])  # zero length: cell + one overshoot cell
    assert (3, 3, 3) in t.cells() and len(t.cells()) == 2
    t = orc.octree(g)
    t.add_line([-5, -5, -5], [-4, -4, -4])  # misses the grid box entirely
    assert t.nblocks() == 0 or t.ncells() == 0
    t = orc.octree(g)
    t.add_line([-0.25, 0.05, 0.05], [0.25, 0.05, 0.05])  # enters from outside
    assert (0, 0, 0) in t.cells() and (2, 0, 0) in t.cells()
    t = orc.octree(g)
    t.add_line([0.05, 0.05, 0.05], [0.35, 0.35, 0.35])  # diagonal
    c = t.cells()
    assert (0, 0, 0) in c and (3, 3, 3) in c and len(c) >= 4


def test_add_line_endpoints_always_marked(orc):
    g = _grid(orc, 32, (-0.21, 0.21, -0.21, 0.21, -0.21, 0.21))
    rng = np.random.default_rng(3)
    for _ in range(200):
        a = rng.uniform(-0.2, 0.2, 3)
        b = a + rng.normal(size=3) * 0.004
        t = orc.octree(g)
        t.add_line(a, b)
        ca, cb = orc.find_cell(g, a), orc.find_cell(g, b)
        cells = t.cells()
        assert ca in cells and cb in cells


def test_union_and_collides_vs_dense(orc):
    g = _grid(orc, 32, (0, 1, 0, 1, 0, 1))
    rng = np.random.default_rng(9)
    for trial in range(30):
        a, b = orc.octree(g), orc.octree(g)
        da, db = np.zeros((8, 8, 8), dtype=np.uint64), np.zeros((8, 8, 8), dtype=np.uint64)
        for tree, dense in ((a, da), (b, db)):
            for _ in range(rng.integers(1, 6)):
                bx, by, bz = rng.integers(0, 8, 3)
                v = int(rng.integers(1, 2 ** 63)) if trial % 2 else 1 << int(rng.integers(0, 64))
                tree.union_block(int(bx), int(by), int(bz), v)
                dense[bx, by, bz] |= np.uint64(v)
        assert a.collides(b) == int(np.any(da & db))
        u = a.copy()
        u.add_voxels(b)
        bxyz, bits = u.export()
        dense_u = np.zeros((8, 8, 8), dtype=np.uint64)
        dense_u[bxyz[:, 0], bxyz[:, 1], bxyz[:, 2]] = bits
        assert np.array_equal(dense_u, da | db)
    other = orc.octree(_grid(orc, 16, (0, 1, 0, 1, 0, 1)))
    assert a.collides(other) == -1  # dimension mismatch: reference throws std::invalid_argument


def test_find_cell_domain(orc):
    g = _grid(orc, 16)
    assert orc.find_cell(g, [0.0, 0.0, 0.0]) == (0, 0, 0)
    assert orc.find_cell(g, [0.15, 0.25, 1.55]) == (1, 2, 15)
    assert orc.find_cell(g, [1.6, 0.0, 0.0])[0] == 16  # upper limit is inside the domain check
    assert orc.find_cell(g, [-1e-9, 0.0, 0.0]) is None and orc.find_cell(g, [0, 0, 1.7]) is None


# ---------------------------------------------------------------------------------------------
# swept volume / OMPL-side restatement
# ---------------------------------------------------------------------------------------------
def test_valid_segment_count_table(orc, wl):
    spec = wl.robot_b(rotation=True)
    rb, sp = orc.robot(spec), orc.space()
    a = np.zeros(8)
    b = a.copy()
    assert orc.valid_segment_count(rb, sp, a, b) == 0
    b[0] = 1.0  # tension distance 1.0 / 0.02 = 50
    assert orc.valid_segment_count(rb, sp, a, b) == 50
    b[0] = 1.001
    assert orc.valid_segment_count(rb, sp, a, b) == 51
    b = a.copy(); b[6] = 0.1  # rotation 0.1 rad / 0.005 = 20
    assert orc.valid_segment_count(rb, sp, a, b) == 20
    a2 = a.copy(); a2[6] = -3.1; b = a.copy(); b[6] = 3.1  # shortest arc = 2pi - 6.2
    assert orc.valid_segment_count(rb, sp, a2, b) == math.ceil((2 * math.pi - 6.2) / 0.005)
    b = a.copy(); b[7] = 0.01  # retraction 0.01 / 1e-4 = 100
    assert orc.valid_segment_count(rb, sp, a, b) in (100, 101)
    b[0] = 3.0  # max over subspaces
    assert orc.valid_segment_count(rb, sp, a, b) == 150


def test_interpolate_so2_shortest_arc(orc, wl):
    spec = wl.robot_b(rotation=True)
    rb = orc.robot(spec)
    a = np.zeros(8); b = np.zeros(8)
    a[6], b[6] = 3.0, -3.0  # crosses +-pi the short way
    m = orc.interpolate(rb, a, b, 0.5)
    assert abs(abs(m[6]) - math.pi) < 1e-12
    a[6], b[6] = 0.5, 1.5
    assert abs(orc.interpolate(rb, a, b, 0.25)[6] - 0.75) < 1e-15
    a[0], b[0], a[7], b[7] = 2.0, 4.0, 0.0, 0.1
    m = orc.interpolate(rb, a, b, 0.5)
    assert m[0] == 3.0 and abs(m[7] - 0.05) < 1e-17


def test_edge_swept_volume_properties(orc, wl):
    spec = wl.robot_b(0.003)
    rb = orc.robot(spec)
    g = wl.workspace_grid(spec)
    grid = orc.grid(g["Ng"], g["lim"])
    sp = orc.space()
    st = wl.sample_states(spec, 12, stream=21)
    for i in range(0, 12, 2):
        a, b = st[i], st[i] + 0.15 * (st[i + 1] - st[i])
        tree, info = orc.voxelize_edge(rb, grid, sp, a, b)
        assert info["nsamples"] >= 2 and not info["out_of_domain"]
        if not info["is_fully_valid"]:
            continue
        assert info["t"] == 1.0 and np.allclose(info["last_valid"], b)
        cells = tree.cells()
        # superset of both endpoint vertex sets
        for x in (a, b):
            sh = orc.shape(rb, x)
            assert orc.voxelize_shape(grid, sh["p"]).cells() <= cells
        # never more samples than the full subdivision to rel_threshold allows
        nseg = orc.valid_segment_count(rb, sp, a, b)
        assert info["nsamples"] <= 2 * max(nseg, 1) + 1


def test_edge_identical_endpoints(orc, wl):
    spec = wl.robot_b(0.003)
    rb = orc.robot(spec)
    g = wl.workspace_grid(spec)
    grid = orc.grid(g["Ng"], g["lim"])
    x = wl.sample_states(spec, 1, stream=22)[0]
    tree, info = orc.voxelize_edge(rb, grid, orc.space(), x, x)
    assert info["nsamples"] == 2 and info["is_fully_valid"]
    assert tree.cells() == orc.voxelize_shape(grid, orc.shape(rb, x)["p"]).cells()


def test_edge_until_invalid_stops_at_obstacle(orc, wl):
    spec = wl.robot_a(0.003)
    rb = orc.robot(spec)
    g = wl.workspace_grid(spec)
    grid = orc.grid(g["Ng"], g["lim"])
    env = orc.octree(grid)
    a = np.array([0.0, 0, 0, 0]); b = np.array([12.0, 0, 0, 0])
    tip_b = orc.shape(rb, b)["p"][-1]
    env.add_sphere(tip_b, 0.01)  # obstacle where the motion ends
    tree, info = orc.voxelize_edge(rb, grid, orc.space(), a, b, env=env)
    assert not info["is_fully_valid"] and 0.0 < info["t"] < 1.0
    assert not tree.collides(env)  # only the valid prefix is voxelised
    tree2, info2 = orc.voxelize_edge(rb, grid, orc.space(), a, b)  # plain voxelize ignores env
    assert info2["is_fully_valid"] and tree2.collides(env)


def test_batch_drivers_match_single_calls(orc, wl):
    spec = wl.robot_b(0.003)
    rb = orc.robot(spec)
    g = wl.workspace_grid(spec)
    grid = orc.grid(g["Ng"], g["lim"])
    st = wl.sample_states(spec, 16, stream=23)
    store, flags = orc.voxelize_vertices_batch(rb, grid, st, nthreads=2)
    off, keys, bits = store.export()
    for i in range(16):
        sh = orc.shape(rb, st[i])
        single = orc.voxelize_shape(grid, sh["p"])
        bxyz, b1 = single.export()
        k1 = [orc.morton_key(int(x), int(y), int(z), 32) for x, y, z in bxyz]
        lo, hi = int(off[i]), int(off[i + 1])
        if flags[i] == 0:
            assert list(keys[lo:hi]) == k1 and np.array_equal(bits[lo:hi], b1)
        else:
            assert hi == lo
    env = orc.octree(grid)
    env.add_sphere([0.0, 0.0, 0.1], 0.02)
    v = orc.check_sets_batch(store, env, nthreads=2)
    for i in range(16):
        assert v[i] == store.get(i).collides(env)


# ---- environment preparation (collision/VoxelOctree.cpp:533-952) vs dense numpy morphology ----
_OFF6 = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]
# the reference's 27-neighbour list as written: (+1,+1,+1) twice, (-1,+1,+1) never
_OFF27_REF = [(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)
              if (i, j, k) not in ((0, 0, 0), (-1, 1, 1))]


def _dense_of(tree, Ng):
    d = np.zeros((Ng, Ng, Ng), dtype=bool)
    for c in tree.cells():
        d[c] = True
    return d


def _tree_from_dense(orc, g, dense):
    t = orc.octree(g)
    Ng = dense.shape[0]
    w = (1 << (np.arange(4)[:, None, None] * 16 + np.arange(4)[None, :, None] * 4
               + np.arange(4)[None, None, :]).astype(np.uint64))
    for bx in range(Ng // 4):
        for by in range(Ng // 4):
            for bz in range(Ng // 4):
                blk = dense[4 * bx:4 * bx + 4, 4 * by:4 * by + 4, 4 * bz:4 * bz + 4]
                if blk.any():
                    t.set_block(bx, by, bz, int(np.sum(w[blk], dtype=np.uint64)))
    return t


def _shifted(dense, off, fill):
    """out[c] = dense[c + off], `fill` outside the grid."""
    Ng = dense.shape[0]
    p = np.pad(dense, 1, constant_values=fill)
    i, j, k = off
    return p[1 + i:1 + i + Ng, 1 + j:1 + j + Ng, 1 + k:1 + k + Ng]


def _np_dilate(dense, offs, num):
    for _ in range(num):
        new = dense.copy()
        for d in offs:  # cell c is reached from c - d
            new |= _shifted(dense, (-d[0], -d[1], -d[2]), False)
        dense = new
    return dense


def _np_remove_interior(dense, diagonal):
    offs = [(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)] if diagonal \
        else _OFF6 + [(0, 0, 0)]
    interior = np.ones_like(dense)
    for d in offs:
        interior &= _shifted(dense, d, True)
    return dense & ~interior


def _random_env_dense(rng, Ng, nblob):
    d = np.zeros((Ng, Ng, Ng), dtype=bool)
    ax = np.arange(Ng)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    for _ in range(nblob):
        c = rng.integers(-1, Ng + 1, 3)  # blobs may poke through the grid faces
        r = rng.uniform(0.5, Ng / 5)
        d |= (X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2 <= r * r
    d[rng.integers(0, Ng, 12), rng.integers(0, Ng, 12), rng.integers(0, Ng, 12)] = True
    return d


def test_dilate_single_cell_hand_count(orc):
    g = _grid(orc, 32, (0, 1, 0, 1, 0, 1))
    t = orc.octree(g)
    t.set_block(3, 3, 3, 1 << (1 * 16 + 2 * 4 + 1))  # one interior cell
    a = t.copy(); a.dilate(1)
    assert a.ncells() == 7  # the cell + its six face neighbours
    a = t.copy(); a.dilate(3)
    assert a.ncells() == 63  # L1 ball of radius 3: 1 + sum_{k<=3} (4k^2 + 2)
    a = t.copy(); a.dilate(6)
    assert a.ncells() == 377  # radius 6 crosses the four-steps-per-pass boundary of the reference
    a = t.copy(); a.dilate(1, True)
    assert a.ncells() == 26  # 27 minus the never-named (x-1,y+1,z+1)
    c = (13, 14, 13)
    assert (c[0] - 1, c[1] + 1, c[2] + 1) not in a.cells() and (c[0] + 1, c[1] + 1, c[2] + 1) in a.cells()
    a = t.copy(); a.dilate(0)
    assert a.ncells() == 1
    # corner cell: clipped at the grid faces
    t = orc.octree(g)
    t.set_block(0, 0, 0, 1)
    a = t.copy(); a.dilate(2)
    assert a.ncells() == 10  # |{x+y+z<=2, x,y,z>=0}|


@pytest.mark.parametrize("Ng", [16, 32])
def test_dilate_and_remove_interior_vs_dense_numpy(orc, Ng):
    g = _grid(orc, Ng, (0, 1, 0, 2, 0, 0.5))
    rng = np.random.default_rng(77 + Ng)
    for trial in range(4):
        dense = _random_env_dense(rng, Ng, 3 + trial)
        t = _tree_from_dense(orc, g, dense)
        assert np.array_equal(_dense_of(t, Ng), dense)
        for num in (1, 2, 4, 5, 9):
            a = t.copy(); a.dilate(num)
            assert np.array_equal(_dense_of(a, Ng), _np_dilate(dense, _OFF6, num)), (trial, num)
        for num in (1, 3, 5):
            a = t.copy(); a.dilate(num, True)
            assert np.array_equal(_dense_of(a, Ng), _np_dilate(dense, _OFF27_REF, num)), (trial, num)
        for diag in (False, True):
            a = t.copy(); a.remove_interior(diag)
            assert np.array_equal(_dense_of(a, Ng), _np_remove_interior(dense, diag)), (trial, diag)
    # dilate_sphere(r) = dilate_6neighbor(round(r / min(dx,dy,dz))): dz = 0.5/Ng is the smallest
    a, b = t.copy(), t.copy()
    a.dilate_sphere(2.4 * 0.5 / Ng); b.dilate(2)
    assert a.cells() == b.cells()
    a, b = t.copy(), t.copy()
    a.dilate_sphere(2.6 * 0.5 / Ng); b.dilate(3)
    assert a.cells() == b.cells()


def test_remove_interior_full_grid_and_shell(orc):
    g = _grid(orc, 16, (0, 1, 0, 1, 0, 1))
    t = orc.octree(g)
    for bx in range(4):
        for by in range(4):
            for bz in range(4):
                t.set_block(bx, by, bz, (1 << 64) - 1)
    a = t.copy(); a.remove_interior(True)
    assert a.ncells() == 0 and a.nblocks() == 0  # outside counts as occupied: everything is interior
    t.set_block(0, 0, 0, ((1 << 64) - 1) & ~1)  # open one corner cell
    a = t.copy(); a.remove_interior(False)
    assert a.cells() == {(1, 0, 0), (0, 1, 0), (0, 0, 1)}
    a = t.copy(); a.remove_interior(True)
    assert a.cells() == {(i, j, k) for i in (0, 1) for j in (0, 1) for k in (0, 1)} - {(0, 0, 0)}
