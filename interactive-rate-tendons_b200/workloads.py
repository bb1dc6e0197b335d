"""Synthetic workloads of BASELINE.json: benchmark robots, seeded
configuration samplers, roadmap topology and the lung-like obstacle environment.

Everything is generated from numpy's counter-based Philox generator, seed 20220801,
so the CPU oracle and the GPU path are fed identical arrays.  Reference defaults cited
relative to /root/reference/cpp/src/.
"""
import math

import numpy as np

SEED = 20220801


def rng(stream=0):
    return np.random.Generator(np.random.Philox(key=SEED + stream))


def _base(dL):
    # tendon/BackboneSpecs.h:15-21, tendon/TendonRobot.h:53-58
    return dict(r=0.015, L=0.2, dL=dL, ro=0.01, ri=0.0, E=2.1e6, nu=0.3,
                residual_threshold=5e-6, enable_rotation=False, enable_retraction=False)


def robot_a(dL=0.005):
    """Robot A (config C1): 4 straight tendons, no rotation/retraction.
    tendon/TendonSpecs.h:28-30 for the limits."""
    rb = _base(dL)
    rb["C"] = [[k * math.pi / 2] for k in range(4)]
    rb["D"] = [[0.01] for _ in range(4)]
    rb["max_tension"] = [20.0] * 4
    rb["min_length"] = [-0.015] * 4
    rb["max_length"] = [0.035] * 4
    return rb


def robot_b(dL=0.005, rotation=False):
    """Robot B (configs C2-C5): 6 helical tendons, one turn, alternating handedness,
    retraction enabled (optionally rotation too)."""
    rb = _base(dL)
    L = rb["L"]
    rb["C"] = [[k * math.pi / 3, (-1) ** k * 2 * math.pi / L] for k in range(6)]
    rb["D"] = [[0.01] for _ in range(6)]
    rb["max_tension"] = [20.0] * 6
    rb["min_length"] = [-0.015] * 6
    rb["max_length"] = [0.035] * 6
    rb["enable_retraction"] = True
    rb["enable_rotation"] = bool(rotation)
    return rb


def state_size(rb):
    return len(rb["C"]) + int(rb["enable_rotation"]) + int(rb["enable_retraction"])


def sample_states(rb, n, stream=0, like_sphere=True):
    """tau_i ~ U[0,max_tension]; rotation ~ U[-pi,pi]; retraction = L - L*cbrt(U)
    (motion-planning/RetractionSampler.h:53-62) or U[0,L] (tendon/TendonRobot.cpp:240-243)."""
    g = rng(stream)
    N = len(rb["C"])
    cols = [g.uniform(0.0, rb["max_tension"][i], size=n) for i in range(N)]
    if rb["enable_rotation"]:
        cols.append(g.uniform(-math.pi, math.pi, size=n))
    if rb["enable_retraction"]:
        u = g.uniform(0.0, 1.0, size=n)
        cols.append(rb["L"] - rb["L"] * np.cbrt(u) if like_sphere else u * rb["L"])
    return np.ascontiguousarray(np.stack(cols, axis=1))


def workspace_grid(rb, Ng=128, padding=0.05):
    """python/src/voxel_ops.py:130-151: cubic limits +-L*(1+padding), identity rotation."""
    h = rb["L"] * (1.0 + padding)
    return dict(Ng=Ng, lim=[-h, h, -h, h, -h, h], inv_rot=np.eye(3))


def space_weights(rb):
    """motion-planning/Problem.cpp:118-141: compound-space distance weights."""
    ext = math.sqrt(sum(t * t for t in rb["max_tension"]))
    return dict(ext=ext, w_rot=ext / (4 * math.pi), w_ret=2 * ext / rb["L"])


def knn_edges(rb, states, k=10, block=2048):
    """Roadmap topology: undirected k-nearest-neighbour edges under the compound-space
    distance |dtau| + w_rot*|dtheta| + w_ret*|ds| (exact brute force; topology is an input,
    not under test).  Returns int64[m,2] with a<b, deduplicated, sorted."""
    n = states.shape[0]
    N = len(rb["C"])
    w = space_weights(rb)
    tau = states[:, :N]
    idx = N
    rot = ret = None
    if rb["enable_rotation"]:
        rot = states[:, idx]
        idx += 1
    if rb["enable_retraction"]:
        ret = states[:, idx]
    pairs = []
    sq = (tau * tau).sum(1)
    for s in range(0, n, block):
        e = min(n, s + block)
        d2 = sq[s:e, None] + sq[None, :] - 2.0 * tau[s:e] @ tau.T
        d = np.sqrt(np.maximum(d2, 0.0))
        if rot is not None:
            dr = np.abs(rot[s:e, None] - rot[None, :])
            dr = np.where(dr > math.pi, 2 * math.pi - dr, dr)
            d += w["w_rot"] * dr
        if ret is not None:
            d += w["w_ret"] * np.abs(ret[s:e, None] - ret[None, :])
        d[np.arange(e - s), np.arange(s, e)] = np.inf
        kk = min(k, n - 1)
        nb = np.argpartition(d, kk, axis=1)[:, :kk]
        src = np.repeat(np.arange(s, e), kk)
        pairs.append(np.stack([src, nb.reshape(-1)], axis=1))
    pr = np.concatenate(pairs, axis=0)
    pr = np.sort(pr, axis=1)
    pr = np.unique(pr, axis=0)
    return pr.astype(np.int64)


def lung_like_capsules(rb, stream=7, generations=4):
    """Random branching 'airway' tree rooted at the origin along +z.  Returns a list of
    (a[3], b[3], radius).  The obstacle is everything in the workspace ball NOT inside the
    airway (apps/prepare_voxel_env.cpp:49-80,273-305 builds the analogous complement shell)."""
    g = rng(stream)
    L = rb["L"]
    caps = []
    frontier = [(np.zeros(3), np.array([0.0, 0.0, 1.0]), 0.012, 0.07)]
    for gen in range(generations):
        nxt = []
        for (p0, d, rad, ln) in frontier:
            p1 = p0 + d * ln
            caps.append((p0.copy(), p1.copy(), float(rad)))
            for _ in range(2):
                perturb = g.normal(size=3) * 0.7
                nd = d + perturb
                nd /= np.linalg.norm(nd)
                if nd[2] < -0.2:
                    nd[2] = -nd[2]
                nxt.append((p1, nd, max(0.004, rad * 0.8), ln * 0.75))
        frontier = nxt
    return [c for c in caps if np.linalg.norm(c[1]) < 1.2 * L]


def lung_like_env_dense(rb, grid, stream=7, shell_voxels=2, radius_scale=4.0):
    """Dense boolean obstacle array [Ng,Ng,Ng] (x,y,z): a shell of `shell_voxels` voxels
    around the airway tree (distance in (r_airway, r_airway + shell*dx]) clipped to the
    workspace ball, with the base hole left open.  `radius_scale` widens the airways; 4.0 gives
    ~2.6 % occupied leaf blocks at 128^3 and about half of random configurations colliding."""
    Ng = grid["Ng"]
    lim = grid["lim"]
    dx = (lim[1] - lim[0]) / Ng
    ax = lim[0] + (np.arange(Ng) + 0.5) * dx
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    P = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    dist = np.full(P.shape[0], np.inf)
    for (a, b, rad) in lung_like_capsules(rb, stream):
        ab = b - a
        t = np.clip(((P - a) @ ab) / max(ab @ ab, 1e-30), 0.0, 1.0)
        c = a + t[:, None] * ab
        dist = np.minimum(dist, np.linalg.norm(P - c, axis=1) - rad * radius_scale)
    shell = (dist > 0.0) & (dist <= shell_voxels * dx)
    inside_ball = np.linalg.norm(P, axis=1) <= rb["L"] * 1.02
    occ = shell & inside_ball
    return occ.reshape(Ng, Ng, Ng)


def dense_to_morton_blocks(occ):
    """bool[Ng,Ng,Ng] -> uint64[Nb^3] indexed by morton block key (x-major octant order,
    the reference's visit_leaves order, collision/detail/TreeNode.h:66-68); bit = x*16+y*4+z
    (collision/VoxelOctree.cpp:1501-1503)."""
    Ng = occ.shape[0]
    Nb = Ng // 4
    o = occ.reshape(Nb, 4, Nb, 4, Nb, 4).transpose(0, 2, 4, 1, 3, 5).reshape(Nb, Nb, Nb, 64)
    weights = (np.uint64(1) << np.arange(64, dtype=np.uint64))
    blocks = (o.astype(np.uint64) * weights).sum(axis=-1, dtype=np.uint64)
    bx, by, bz = np.meshgrid(np.arange(Nb), np.arange(Nb), np.arange(Nb), indexing="ij")
    key = morton_key(bx, by, bz, Nb)
    out = np.zeros(Nb ** 3, dtype=np.uint64)
    out[key.reshape(-1)] = blocks.reshape(-1)
    return out


def morton_key(bx, by, bz, Nb):
    bx = np.asarray(bx, dtype=np.uint32)
    by = np.asarray(by, dtype=np.uint32)
    bz = np.asarray(bz, dtype=np.uint32)
    key = np.zeros(np.broadcast(bx, by, bz).shape, dtype=np.uint32)
    l = 0
    while (1 << l) < Nb:
        key |= (((bx >> l) & 1) << (3 * l + 2)) | (((by >> l) & 1) << (3 * l + 1)) | (((bz >> l) & 1) << (3 * l))
        l += 1
    return key


def morton_decode(key, Nb):
    key = np.asarray(key, dtype=np.uint32)
    bx = np.zeros(key.shape, dtype=np.uint32)
    by = np.zeros(key.shape, dtype=np.uint32)
    bz = np.zeros(key.shape, dtype=np.uint32)
    l = 0
    while (1 << l) < Nb:
        bx |= ((key >> (3 * l + 2)) & 1) << l
        by |= ((key >> (3 * l + 1)) & 1) << l
        bz |= ((key >> (3 * l)) & 1) << l
        l += 1
    return bx, by, bz


def toggle_blob(env_blocks, grid, center, radius):
    """C5 tick: XOR a spherical blob of voxels into a morton-ordered dense block array."""
    Ng = grid["Ng"]
    Nb = Ng // 4
    lim = grid["lim"]
    dx = (lim[1] - lim[0]) / Ng
    lo = np.maximum(0, np.floor((np.asarray(center) - radius - lim[0]) / dx).astype(int))
    hi = np.minimum(Ng - 1, np.floor((np.asarray(center) + radius - lim[0]) / dx).astype(int))
    out = env_blocks.copy()
    for ix in range(lo[0], hi[0] + 1):
        for iy in range(lo[1], hi[1] + 1):
            for iz in range(lo[2], hi[2] + 1):
                c = lim[0] + (np.array([ix, iy, iz]) + 0.5) * dx
                if np.linalg.norm(c - center) <= radius:
                    k = int(morton_key(ix // 4, iy // 4, iz // 4, Nb))
                    out[k] ^= np.uint64(1) << np.uint64((ix % 4) * 16 + (iy % 4) * 4 + (iz % 4))
    return out
