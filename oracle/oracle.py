"""ctypes binding of the CPU ORACLE (oracle/liboracle*.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.
See oracle/tendon_oracle.h for what is restated and the parity status
(partly pinned by the reference's own code in oracle/_ref, see oracle/ref.py; "parity unpinned by
the reference" for the parts that are restated only).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_TENDONS = 12
MAX_COEF = 8

FLAG_NONCONVERGED = 1
FLAG_LENGTH_LIMIT = 2
FLAG_SELF_COLLISION = 4
FLAG_OUT_OF_DOMAIN = 8
FLAG_PARTIAL = 16


class OrcRobot(C.Structure):
    _fields_ = [
        ("r", C.c_double),
        ("L", C.c_double), ("dL", C.c_double), ("ro", C.c_double), ("ri", C.c_double),
        ("E", C.c_double), ("nu", C.c_double),
        ("residual_threshold", C.c_double),
        ("n_tendons", C.c_int32), ("n_c", C.c_int32), ("n_d", C.c_int32),
        ("enable_rotation", C.c_int32), ("enable_retraction", C.c_int32), ("_pad", C.c_int32),
        ("C", C.c_double * (MAX_TENDONS * MAX_COEF)),
        ("D", C.c_double * (MAX_TENDONS * MAX_COEF)),
        ("max_tension", C.c_double * MAX_TENDONS),
        ("min_length", C.c_double * MAX_TENDONS),
        ("max_length", C.c_double * MAX_TENDONS),
    ]


class OrcGrid(C.Structure):
    _fields_ = [("Ng", C.c_int32), ("_pad", C.c_int32), ("lim", C.c_double * 6),
                ("inv_rot", C.c_double * 9)]


class OrcSpace(C.Structure):
    _fields_ = [("min_tension_change", C.c_double), ("min_rotation_change", C.c_double),
                ("min_retraction_change", C.c_double)]


class OrcFkOut(C.Structure):
    _fields_ = [("npts", C.c_int32), ("converged", C.c_int32), ("iters", C.c_int32),
                ("nsteps", C.c_int32), ("L", C.c_double), ("L_i", C.c_double * MAX_TENDONS),
                ("u_i", C.c_double * 3), ("u_f", C.c_double * 3), ("v_i", C.c_double * 3),
                ("v_f", C.c_double * 3)]


class OrcEdgeOut(C.Structure):
    _fields_ = [("is_fully_valid", C.c_int32), ("nsamples", C.c_int32),
                ("out_of_domain", C.c_int32), ("_pad", C.c_int32), ("t", C.c_double),
                ("last_valid", C.c_double * (MAX_TENDONS + 2))]


def build(force=False):
    """Compile oracle/liboracle.so and liboracle_fast.so (gcc only, no GPU needed)."""
    need = force or not all(os.path.exists(os.path.join(_HERE, n))
                            for n in ("liboracle.so", "liboracle_fast.so", "liboracle_refflags.so"))
    src_m = max(os.path.getmtime(os.path.join(_HERE, n))
                for n in ("tendon_oracle.cpp", "tendon_oracle.h"))
    if not need:
        need = any(os.path.getmtime(os.path.join(_HERE, n)) < src_m
                   for n in ("liboracle.so", "liboracle_fast.so", "liboracle_refflags.so"))
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-s"], env=dict(os.environ, CXX="g++"))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class Oracle:
    """One loaded oracle library ("canonical" or "fast")."""

    def __init__(self, variant="canonical"):
        name = {"canonical": "liboracle.so", "fast": "liboracle_fast.so",
                "refflags": "liboracle_refflags.so"}[variant]
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        self.variant = variant
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_state_size.restype = C.c_int
        L.orc_t_range.restype = C.c_int
        L.orc_t_range.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), C.c_int]
        L.orc_shape.restype = C.c_int
        L.orc_collides_self.restype = C.c_int
        L.orc_collides_self.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double]
        L.orc_validity_flags.restype = C.c_uint32
        L.orc_octree_new.restype = C.c_void_p
        L.orc_octree_copy.restype = C.c_void_p
        L.orc_octree_copy.argtypes = [C.c_void_p]
        L.orc_octree_free.argtypes = [C.c_void_p]
        L.orc_octree_clear.argtypes = [C.c_void_p]
        L.orc_octree_block.restype = C.c_uint64
        L.orc_octree_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_octree_set_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64]
        L.orc_octree_union_block.restype = C.c_uint64
        L.orc_octree_union_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64]
        L.orc_octree_nblocks.restype = C.c_int64
        L.orc_octree_nblocks.argtypes = [C.c_void_p]
        L.orc_octree_ncells.restype = C.c_int64
        L.orc_octree_ncells.argtypes = [C.c_void_p]
        L.orc_octree_add_line.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_octree_add_piecewise_line.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        L.orc_octree_add_voxels.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_octree_collides.restype = C.c_int
        L.orc_octree_collides.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_octree_export.restype = C.c_int64
        L.orc_octree_export.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)]
        L.orc_octree_add_sphere.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_double]
        L.orc_octree_add_capsule.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double]
        L.orc_octree_dilate_6neighbor.argtypes = [C.c_void_p, C.c_int]
        L.orc_octree_dilate_27neighbor.argtypes = [C.c_void_p, C.c_int]
        L.orc_octree_dilate_sphere.argtypes = [C.c_void_p, C.c_double]
        L.orc_octree_remove_interior.argtypes = [C.c_void_p, C.c_int]
        L.orc_find_cell.restype = C.c_int
        L.orc_valid_segment_count.restype = C.c_uint32
        L.orc_voxelize_shape.argtypes = [C.POINTER(OrcGrid), C.POINTER(C.c_double), C.c_int, C.c_void_p]
        L.orc_voxelize_edge.argtypes = [C.POINTER(OrcRobot), C.POINTER(OrcGrid), C.POINTER(OrcSpace),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p,
                                        C.c_void_p, C.POINTER(OrcEdgeOut)]
        L.orc_setstore_new.restype = C.c_void_p
        L.orc_setstore_new.argtypes = [C.POINTER(OrcGrid), C.c_int64]
        L.orc_setstore_free.argtypes = [C.c_void_p]
        L.orc_setstore_size.restype = C.c_int64
        L.orc_setstore_size.argtypes = [C.c_void_p]
        L.orc_setstore_get.restype = C.c_void_p
        L.orc_setstore_get.argtypes = [C.c_void_p, C.c_int64]
        L.orc_setstore_total_blocks.restype = C.c_int64
        L.orc_setstore_total_blocks.argtypes = [C.c_void_p]
        L.orc_setstore_export.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                          C.POINTER(C.c_uint64)]
        L.orc_voxelize_vertices_batch.argtypes = [C.POINTER(OrcRobot), C.POINTER(OrcGrid),
                                                  C.POINTER(C.c_double), C.c_int64, C.c_void_p,
                                                  C.POINTER(C.c_uint32), C.c_int]
        L.orc_voxelize_edges_batch.argtypes = [C.POINTER(OrcRobot), C.POINTER(OrcGrid),
                                               C.POINTER(OrcSpace), C.POINTER(C.c_double),
                                               C.POINTER(C.c_double), C.c_int64, C.c_void_p,
                                               C.POINTER(C.c_uint32), C.POINTER(C.c_double),
                                               C.POINTER(C.c_int32), C.c_int]
        L.orc_check_sets_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                           C.POINTER(C.c_uint8), C.c_int]
        L.orc_morton_key.restype = C.c_uint32
        L.orc_morton_key.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_fk_batch.argtypes = [C.POINTER(OrcRobot), C.POINTER(C.c_double), C.c_int64, C.c_int,
                                   C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                   C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32), C.c_int]

    # ---- struct helpers --------------------------------------------------
    @staticmethod
    def robot(spec):
        """spec: dict with keys r,L,dL,ro,ri,E,nu,residual_threshold,C,D (lists of lists),
        max_tension,min_length,max_length,enable_rotation,enable_retraction."""
        rb = OrcRobot()
        for k in ("r", "L", "dL", "ro", "ri", "E", "nu", "residual_threshold"):
            setattr(rb, k, float(spec[k]))
        Cc, Dd = spec["C"], spec["D"]
        n = len(Cc)
        rb.n_tendons = n
        rb.n_c = len(Cc[0]) if n else 0
        rb.n_d = len(Dd[0]) if n else 0
        assert n <= MAX_TENDONS and rb.n_c <= MAX_COEF and rb.n_d <= MAX_COEF
        for j in range(n):
            for i, c in enumerate(Cc[j]):
                rb.C[j * MAX_COEF + i] = float(c)
            for i, d in enumerate(Dd[j]):
                rb.D[j * MAX_COEF + i] = float(d)
            rb.max_tension[j] = float(spec["max_tension"][j])
            rb.min_length[j] = float(spec["min_length"][j])
            rb.max_length[j] = float(spec["max_length"][j])
        rb.enable_rotation = int(bool(spec.get("enable_rotation", False)))
        rb.enable_retraction = int(bool(spec.get("enable_retraction", False)))
        return rb

    @staticmethod
    def grid(Ng, lim, inv_rot=None):
        g = OrcGrid()
        g.Ng = int(Ng)
        for i, v in enumerate(lim):
            g.lim[i] = float(v)
        R = np.eye(3) if inv_rot is None else np.asarray(inv_rot, dtype=np.float64).reshape(3, 3)
        for i, v in enumerate(R.reshape(-1)):
            g.inv_rot[i] = float(v)
        return g

    @staticmethod
    def space(min_tension_change=0.02, min_rotation_change=0.01, min_retraction_change=0.0001):
        return OrcSpace(min_tension_change, min_rotation_change, min_retraction_change)

    # ---- FK ----------------------------------------------------------------
    def state_size(self, rb):
        return self.lib.orc_state_size(C.byref(rb))

    def t_range(self, s, L, dL, cap=4096):
        out = np.empty(cap)
        n = self.lib.orc_t_range(s, L, dL, _dp(out), cap)
        assert n >= 0
        return out[:n].copy()

    def routing(self, rb, t):
        N = rb.n_tendons
        r, rd, rdd = (np.zeros((N, 3)) for _ in range(3))
        self.lib.orc_routing(C.byref(rb), C.c_double(t), _dp(r), _dp(rd), _dp(rdd))
        return r, rd, rdd

    def deriv(self, rb, tau, x, t, alt=False):
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros_like(x)
        f = self.lib.orc_tendon_deriv_alt if alt else self.lib.orc_tendon_deriv
        f(C.byref(rb), _dp(tau), _dp(x), C.c_double(t), _dp(out))
        return out

    def shape(self, rb, state, cap=1024):
        state = np.ascontiguousarray(state, dtype=np.float64)
        t = np.zeros(cap)
        p = np.zeros((cap, 3))
        R = np.zeros((cap, 9))
        out = OrcFkOut()
        n = self.lib.orc_shape(C.byref(rb), _dp(state), cap, _dp(t), _dp(p), _dp(R), C.byref(out))
        assert n >= 0, n
        N = rb.n_tendons
        return dict(t=t[:n].copy(), p=p[:n].copy(),
                    R=R[:n].reshape(n, 3, 3).transpose(0, 2, 1).copy(),  # col-major -> [i][row][col]
                    L=out.L, L_i=np.array(out.L_i[:N]), u_i=np.array(out.u_i[:]),
                    u_f=np.array(out.u_f[:]), v_i=np.array(out.v_i[:]), v_f=np.array(out.v_f[:]),
                    converged=bool(out.converged), iters=out.iters, nsteps=out.nsteps, _raw=out)

    def tip_jacobian(self, rb, state, mode, delta):
        state = np.ascontiguousarray(state, dtype=np.float64)
        m = self.state_size(rb)
        tip, J = np.zeros(3), np.zeros((3, m))
        self.lib.orc_tip_jacobian(C.byref(rb), _dp(state), C.c_int(mode), C.c_double(delta), _dp(tip), _dp(J))
        return tip, J

    def home_lengths(self, rb, s):
        out = np.zeros(rb.n_tendons)
        self.lib.orc_home_lengths(C.byref(rb), C.c_double(s), _dp(out))
        return out

    def collides_self(self, p, r):
        p = np.ascontiguousarray(p, dtype=np.float64)
        return bool(self.lib.orc_collides_self(_dp(p), len(p), r))

    def closest_st_segment(self, A, B, Cc, D):
        a, b, c, d = (np.ascontiguousarray(v, dtype=np.float64) for v in (A, B, Cc, D))
        s, t = C.c_double(), C.c_double()
        self.lib.orc_closest_st_segment(_dp(a), _dp(b), _dp(c), _dp(d), C.byref(s), C.byref(t))
        return s.value, t.value

    def segment_aabox_intersect(self, A, B, Cc, D):
        a, b, c, d = (np.ascontiguousarray(v, dtype=np.float64) for v in (A, B, Cc, D))
        return bool(self.lib.orc_segment_aabox_intersect(_dp(a), _dp(b), _dp(c), _dp(d)))

    def validity_flags(self, rb, state, shape):
        state = np.ascontiguousarray(state, dtype=np.float64)
        p = np.ascontiguousarray(shape["p"], dtype=np.float64)
        return int(self.lib.orc_validity_flags(C.byref(rb), _dp(state), C.byref(shape["_raw"]), _dp(p)))

    def fk_batch(self, rb, states, cap_pts, nthreads=0, want_p=True):
        states = np.ascontiguousarray(states, dtype=np.float64)
        n = states.shape[0]
        N = rb.n_tendons
        p = np.zeros((n, cap_pts, 3)) if want_p else None
        npts = np.zeros(n, dtype=np.int32)
        L_i = np.zeros((n, N))
        tip = np.zeros((n, 3))
        flags = np.zeros(n, dtype=np.uint32)
        iters = np.zeros(n, dtype=np.int32)
        nsteps = np.zeros(n, dtype=np.int32)
        nt = nthreads or self.max_threads()
        self.lib.orc_fk_batch(C.byref(rb), _dp(states), n, cap_pts, _dp(p), _p(npts, C.c_int32),
                              _dp(L_i), _dp(tip), _p(flags, C.c_uint32), _p(iters, C.c_int32),
                              _p(nsteps, C.c_int32), nt)
        return dict(p=p, npts=npts, L_i=L_i, tip=tip, flags=flags, iters=iters, nsteps=nsteps)

    # ---- voxels --------------------------------------------------------------
    def octree(self, grid):
        return Octree(self, self.lib.orc_octree_new(C.byref(grid)), grid)

    def find_cell(self, grid, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        cell = np.zeros(3, dtype=np.int64)
        err = self.lib.orc_find_cell(C.byref(grid), _dp(p), _p(cell, C.c_int64))
        return (None if err else tuple(int(c) for c in cell))

    def morton_key(self, bx, by, bz, Nb):
        return int(self.lib.orc_morton_key(bx, by, bz, Nb))

    def valid_segment_count(self, rb, sp, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        return int(self.lib.orc_valid_segment_count(C.byref(rb), C.byref(sp), _dp(a), _dp(b)))

    def interpolate(self, rb, a, b, t):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        out = np.zeros_like(a)
        self.lib.orc_interpolate(C.byref(rb), _dp(a), _dp(b), C.c_double(t), _dp(out))
        return out

    def voxelize_shape(self, grid, p):
        tree = self.octree(grid)
        p = np.ascontiguousarray(p, dtype=np.float64)
        self.lib.orc_voxelize_shape(C.byref(grid), _dp(p), len(p), tree.h)
        return tree

    def voxelize_edge(self, rb, grid, sp, a, b, env=None):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        tree = self.octree(grid)
        info = OrcEdgeOut()
        self.lib.orc_voxelize_edge(C.byref(rb), C.byref(grid), C.byref(sp), _dp(a), _dp(b),
                                   env.h if env is not None else None, tree.h, C.byref(info))
        S = self.state_size(rb)
        return tree, dict(is_fully_valid=bool(info.is_fully_valid), nsamples=info.nsamples,
                          out_of_domain=bool(info.out_of_domain), t=info.t,
                          last_valid=np.array(info.last_valid[:S]))

    def setstore(self, grid, n):
        return SetStore(self, grid, n)

    def voxelize_vertices_batch(self, rb, grid, states, nthreads=0):
        states = np.ascontiguousarray(states, dtype=np.float64)
        n = states.shape[0]
        store = self.setstore(grid, n)
        flags = np.zeros(n, dtype=np.uint32)
        self.lib.orc_voxelize_vertices_batch(C.byref(rb), C.byref(grid), _dp(states), n, store.h,
                                             _p(flags, C.c_uint32), nthreads or self.max_threads())
        return store, flags

    def voxelize_edges_batch(self, rb, grid, sp, a, b, nthreads=0):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        n = a.shape[0]
        store = self.setstore(grid, n)
        flags = np.zeros(n, dtype=np.uint32)
        t_last = np.zeros(n)
        nsamples = np.zeros(n, dtype=np.int32)
        self.lib.orc_voxelize_edges_batch(C.byref(rb), C.byref(grid), C.byref(sp), _dp(a), _dp(b), n,
                                          store.h, _p(flags, C.c_uint32), _dp(t_last),
                                          _p(nsamples, C.c_int32), nthreads or self.max_threads())
        return store, dict(flags=flags, t_last=t_last, nsamples=nsamples)

    def check_sets_batch(self, store, env, begin=0, end=None, nthreads=0):
        end = store.size() if end is None else end
        verdict = np.zeros(end - begin, dtype=np.uint8)
        self.lib.orc_check_sets_batch(store.h, env.h, begin, end, _p(verdict, C.c_uint8),
                                      nthreads or self.max_threads())
        return verdict

    def max_threads(self):
        return int(self.lib.orc_max_threads())


class Octree:
    def __init__(self, orc, handle, grid, owned=True):
        self.orc, self.h, self.grid, self.owned = orc, C.c_void_p(handle), grid, owned

    def __del__(self):
        if getattr(self, "owned", False) and self.h:
            self.orc.lib.orc_octree_free(self.h)
            self.h = None

    def copy(self):
        return Octree(self.orc, self.orc.lib.orc_octree_copy(self.h), self.grid)

    def block(self, bx, by, bz):
        return int(self.orc.lib.orc_octree_block(self.h, bx, by, bz))

    def set_block(self, bx, by, bz, v):
        self.orc.lib.orc_octree_set_block(self.h, bx, by, bz, C.c_uint64(v))

    def union_block(self, bx, by, bz, v):
        return int(self.orc.lib.orc_octree_union_block(self.h, bx, by, bz, C.c_uint64(v)))

    def nblocks(self):
        return int(self.orc.lib.orc_octree_nblocks(self.h))

    def ncells(self):
        return int(self.orc.lib.orc_octree_ncells(self.h))

    def add_line(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        self.orc.lib.orc_octree_add_line(self.h, _dp(a), _dp(b))

    def add_piecewise_line(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        self.orc.lib.orc_octree_add_piecewise_line(self.h, _dp(pts), len(pts))

    def add_point(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        self.orc.lib.orc_octree_add_point(self.h, _dp(p))

    def add_sphere(self, c, r):
        c = np.ascontiguousarray(c, dtype=np.float64)
        self.orc.lib.orc_octree_add_sphere(self.h, _dp(c), r)

    def add_capsule(self, a, b, r):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        self.orc.lib.orc_octree_add_capsule(self.h, _dp(a), _dp(b), r)

    def add_voxels(self, other):
        self.orc.lib.orc_octree_add_voxels(self.h, other.h)

    def dilate(self, num=1, use_diagonal=False):
        (self.orc.lib.orc_octree_dilate_27neighbor if use_diagonal else self.orc.lib.orc_octree_dilate_6neighbor)(self.h, num)

    def dilate_sphere(self, r):
        self.orc.lib.orc_octree_dilate_sphere(self.h, r)

    def remove_interior(self, keep_diagonal=True):
        self.orc.lib.orc_octree_remove_interior(self.h, int(keep_diagonal))

    def collides(self, other):
        return int(self.orc.lib.orc_octree_collides(self.h, other.h))

    def export(self):
        """(bxyz uint8[n,3], bits uint64[n]) in visit_leaves order."""
        n = int(self.orc.lib.orc_octree_export(self.h, 0, None, None))
        bxyz = np.zeros((max(n, 1), 3), dtype=np.uint8)
        bits = np.zeros(max(n, 1), dtype=np.uint64)
        self.orc.lib.orc_octree_export(self.h, n, _p(bxyz, C.c_uint8), _p(bits, C.c_uint64))
        return bxyz[:n], bits[:n]

    def cells(self):
        """set of occupied (ix,iy,iz)."""
        bxyz, bits = self.export()
        out = set()
        for (bx, by, bz), b in zip(bxyz.tolist(), bits.tolist()):
            for x in range(4):
                for y in range(4):
                    for z in range(4):
                        if (b >> (x * 16 + y * 4 + z)) & 1:
                            out.add((4 * bx + x, 4 * by + y, 4 * bz + z))
        return out

    def dense_morton(self):
        """dense uint64[Nb^3] indexed by morton key (x-major octant order)."""
        Nb = self.grid.Ng // 4
        out = np.zeros(Nb ** 3, dtype=np.uint64)
        bxyz, bits = self.export()
        for (bx, by, bz), b in zip(bxyz.tolist(), bits.tolist()):
            out[self.orc.morton_key(bx, by, bz, Nb)] = b
        return out


class SetStore:
    def __init__(self, orc, grid, n):
        self.orc, self.grid = orc, grid
        self.h = C.c_void_p(orc.lib.orc_setstore_new(C.byref(grid), n))

    def __del__(self):
        if getattr(self, "h", None):
            self.orc.lib.orc_setstore_free(self.h)
            self.h = None

    def size(self):
        return int(self.orc.lib.orc_setstore_size(self.h))

    def get(self, i):
        return Octree(self.orc, self.orc.lib.orc_setstore_get(self.h, i), self.grid, owned=False)

    def export(self):
        """CSR: offsets uint64[n+1], keys uint32[nb] (morton), bits uint64[nb]."""
        n = self.size()
        nb = int(self.orc.lib.orc_setstore_total_blocks(self.h))
        offsets = np.zeros(n + 1, dtype=np.uint64)
        keys = np.zeros(max(nb, 1), dtype=np.uint32)
        bits = np.zeros(max(nb, 1), dtype=np.uint64)
        self.orc.lib.orc_setstore_export(self.h, _p(offsets, C.c_uint64), _p(keys, C.c_uint32),
                                         _p(bits, C.c_uint64))
        return offsets, keys[:nb], bits[:nb]
