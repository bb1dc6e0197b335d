"""Generate tests/golden/reference_vectors.npz from the REFERENCE's own code (oracle/_ref).

Run in the container that has /root/reference:  python tests/golden/make_reference_vectors.py
The vectors are outputs of the reference's unmodified source files compiled here
(oracle/ref.py explains which, and the Eigen stand-in caveat); they let the pins in
tests/test_oracle_vs_reference.py run where /root/reference and oracle/_ref are absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402
import irt_b200.workloads as wl  # noqa: E402

SEED = 20220801


def robots():
    return {"a005": wl.robot_a(0.005), "b003": wl.robot_b(0.003), "b005rot": wl.robot_b(0.005, rotation=True)}


def random_ode_state(rng, N):
    x = np.zeros(19 + N)
    x[:3] = rng.normal(size=3) * 0.1
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    x[3:12] = q.T.reshape(-1)
    x[12:15] = np.array([0, 0, 1.0]) + rng.normal(size=3) * 0.05
    x[15:18] = rng.normal(size=3) * 5
    x[18] = rng.uniform(0, 0.2)
    x[19:] = rng.uniform(0, 0.2, N)
    return x


def main():
    ref.build()
    assert ref.available() and ref.RefVoxelOctree.available() and ref.RefTendonRobot.available()
    out = {}
    rng = np.random.default_rng(SEED)
    for name, spec in robots().items():
        rf = ref.RefFK(spec)
        N = rf.N
        n = 64
        ts = rng.uniform(0, spec["L"], n)
        taus = rng.uniform(0, 20, (n, N))
        xs = np.stack([random_ode_state(rng, N) for _ in range(n)])
        out[name + "_t"] = ts
        out[name + "_tau"] = taus
        out[name + "_x"] = xs
        out[name + "_rinfo"] = np.stack([np.stack(rf.r_info(t)) for t in ts])           # [n][3][N][3]
        out[name + "_dxdt"] = np.stack([rf.deriv(taus[i], xs[i], ts[i]) for i in range(n)])
        s0 = rng.uniform(0, 0.9 * spec["L"], n) if spec.get("enable_retraction") else np.zeros(n)
        ib = [rf.initial_bending(taus[i], s0[i]) for i in range(n)]
        out[name + "_s0"] = s0
        out[name + "_v0"] = np.stack([b[0] for b in ib])
        out[name + "_u0"] = np.stack([b[1] for b in ib])
        out[name + "_iters"] = np.array([b[2] for b in ib], dtype=np.int32)
    # capsule-pair primitive: random, degenerate and parallel cases
    rf = ref.RefFK(wl.robot_a())
    segs = rng.normal(size=(256, 4, 3))
    segs[:16, 1] = segs[:16, 0]                       # A == B
    segs[16:32, 3] = segs[16:32, 2]                   # C == D
    segs[32:64, 3] = segs[32:64, 2] + (segs[32:64, 1] - segs[32:64, 0]) * rng.uniform(-2, 2, (32, 1))  # parallel
    segs[64:80, 2] = segs[64:80, 0] + 0.3 * (segs[64:80, 1] - segs[64:80, 0])   # collinear, overlapping
    segs[64:80, 3] = segs[64:80, 0] + 1.7 * (segs[64:80, 1] - segs[64:80, 0])
    out["segs"] = segs
    out["segs_st"] = np.array([rf.closest_st_segment(*s) for s in segs])
    boxes = rng.uniform(-1, 1, (256, 4, 3))
    out["boxes"] = boxes
    out["boxes_hit"] = np.array([rf.segment_aabox_intersect(*b) for b in boxes], dtype=np.uint8)
    # octree: a random edit script replayed on the reference's TreeNode<32>; leaves in visit order
    Ng = 32
    t = ref.RefTree(Ng)
    script = []
    for _ in range(400):
        op = int(rng.integers(0, 2))   # set_block / union_block (the two the hot path uses)
        bx, by, bz = (int(v) for v in rng.integers(0, Ng // 4, 3))
        val = int(rng.integers(0, 2 ** 63)) if rng.random() > 0.15 else 0
        script.append((op, bx, by, bz, val))
        (t.set_block, t.union_block)[op](bx, by, bz, val)
    out["tree_script"] = np.array(script, dtype=np.uint64)
    bx, by, bz, bits = t.leaves()
    out["tree_leaves_xyz"] = np.stack([bx, by, bz], axis=1)
    out["tree_leaves_bits"] = bits
    # VoxelOctree::add_line / find_cell / dilate / remove_interior (the reference's own text, see
    # oracle/ref_shim/voxeloctree_ref.cpp): segments of every kind -> leaves in visit_leaves order
    lim = [-0.21, 0.21] * 3
    dxv = 0.42 / 128
    segs2 = []
    for i in range(600):
        kind = i % 6
        a = rng.uniform(-0.25, 0.25, 3)
        if kind == 0:
            b = a + rng.normal(size=3) * 0.004                 # short, like a backbone segment
        elif kind == 1:
            b = rng.uniform(-0.3, 0.3, 3)                      # long, may cross or miss the grid
        elif kind == 2:
            b = a.copy(); b[rng.integers(0, 3)] += rng.uniform(-0.05, 0.05)   # axis aligned
        elif kind == 3:
            b = a.copy()                                       # zero length
        elif kind == 4:
            a = np.round(a / dxv) * dxv; b = a + rng.integers(-3, 4, 3) * dxv   # on cell boundaries
        else:
            a = rng.uniform(-0.2, 0.2, 3); b = a + rng.uniform(-1, 1, 3) * np.array([1e-12, 0.02, 1e-11])
        segs2.append((a, b))
    segs2 = np.array(segs2)
    out["vo_lim"] = np.array(lim)
    out["vo_segs"] = segs2
    leaf_off, leaf_xyz, leaf_bits = [0], [], []
    for k in range(0, len(segs2), 3):                          # three segments per tree
        t = ref.RefVoxelOctree(128, lim)
        for a, b in segs2[k:k + 3]:
            t.add_line(a, b)
        xyz, bits = t.export()
        leaf_xyz.append(xyz); leaf_bits.append(bits); leaf_off.append(leaf_off[-1] + len(bits))
    out["vo_leaf_off"] = np.array(leaf_off)
    out["vo_leaf_xyz"] = np.concatenate(leaf_xyz)
    out["vo_leaf_bits"] = np.concatenate(leaf_bits)
    t = ref.RefVoxelOctree(128, lim)
    pts = rng.uniform(-0.22, 0.22, (500, 3))
    pts[:8] = np.array([[-0.21, 0, 0], [0.21, 0, 0], [0, -0.21, 0.21], [0.2099999, 0, 0], [0, 0, 0],
                        [dxv, 2 * dxv, -3 * dxv], [-0.2100001, 0, 0], [0, 0.2100001, 0]])
    out["vo_pts"] = pts
    out["vo_find_cell"] = np.array([(-1, -1, -1) if t.find_cell(p) is None else t.find_cell(p) for p in pts])
    out["vo_nearest_cell"] = np.array([t.nearest_cell(p) for p in pts])
    # environment preparation on a small random obstacle set (32^3 grid)
    seed_cells = rng.integers(2, 30, (60, 3))
    out["vo_env_cells"] = seed_cells
    prep = {}
    for name, args in (("dilate6x1", ("dilate", 1, False)), ("dilate6x3", ("dilate", 3, False)),
                       ("dilate27x2", ("dilate", 2, True)), ("sphere", ("dilate_sphere", 0.07)),
                       ("interior27", ("remove_interior", True)), ("interior6", ("remove_interior", False))):
        t = ref.RefVoxelOctree(32, [0, 1] * 3)
        for c in seed_cells:
            t.set_cell(*[int(v) for v in c])
        if name.startswith("interior"):
            t.dilate(3, True)                                  # something with an interior to remove
        getattr(t, args[0])(*args[1:])
        xyz, bits = t.export()
        out["vo_prep_%s_xyz" % name] = xyz
        out["vo_prep_%s_bits" % name] = bits
    # TendonRobot::shape of the reference (its own tension_shape text, Eigen + odeint stand-ins): whole shapes
    for name, spec in {"a005": wl.robot_a(0.005), "b003": wl.robot_b(0.003),
                       "b005rot": wl.robot_b(0.005, rotation=True),
                       "a003soft": dict(wl.robot_a(0.003), E=0.7e6)}.items():
        rr = ref.RefTendonRobot(spec)
        st = wl.sample_states(spec, 24, stream=29)
        if spec.get("enable_retraction"):
            L = spec["L"]
            st[0, -1], st[1, -1], st[2, -1], st[3, -1], st[4, -1] = L, L + 0.01, 0.1995, 0.0, -0.001
        shapes = [rr.shape(s_) for s_ in st]
        out["tr_%s_states" % name] = st
        out["tr_%s_off" % name] = np.concatenate([[0], np.cumsum([len(sh["t"]) for sh in shapes])])
        out["tr_%s_t" % name] = np.concatenate([sh["t"] for sh in shapes])
        out["tr_%s_p" % name] = np.concatenate([sh["p"] for sh in shapes])
        out["tr_%s_L" % name] = np.array([sh["L"] for sh in shapes])
        out["tr_%s_Li" % name] = np.stack([sh["L_i"] for sh in shapes])
        out["tr_%s_conv" % name] = np.array([sh["converged"] for sh in shapes])
        out["tr_%s_flags" % name] = np.array([rr.flags(s_) for s_ in st], dtype=np.uint32)
        out["tr_%s_home" % name] = np.stack([rr.home_lengths(s_) for s_ in st])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")
    # Environment::voxelize's primitives: VoxelOctree::add(Point) / add_sphere / add_capsule of the reference
    # (VoxelOctree.cpp:319-323, 434-515) on a non-cubic-celled 64^3 grid; kind 0 point, 1 sphere, 2 capsule
    prng = np.random.default_rng(20220801 + 77)
    plim = [-0.3, 0.2, -0.1, 0.4, 0.0, 0.25]
    plo, pext = np.array(plim[0::2]), np.array(plim[1::2]) - np.array(plim[0::2])
    objs = np.zeros((45, 8))
    t = ref.RefVoxelOctree(64, plim)
    for k in range(len(objs)):
        a = plo + pext * prng.uniform(-0.3, 1.3, 3)
        b = a if k % 9 == 2 else a + pext * prng.uniform(-0.5, 0.5, 3)
        r = float(pext.min() * prng.choice([0.001, 0.01, 0.05, 0.12, 0.02, 0.2]))
        objs[k] = np.concatenate([a, b, [r, k % 3]])
        if k % 3 == 0:
            t.add_point(a)
        elif k % 3 == 1:
            t.add_sphere(a, r)
        else:
            t.add_capsule(a, b, r)
    out["vo_prim_lim"] = np.array(plim)
    out["vo_prim_objs"] = objs
    out["vo_prim_xyz"], out["vo_prim_bits"] = t.export()
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
