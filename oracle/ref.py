"""ctypes binding of oracle/_ref/ — pieces of the REFERENCE ITSELF, compiled here.

TEST INFRASTRUCTURE ONLY (same rule as oracle/oracle.py).

* libtreenode_ref.so: the reference's collision/detail/TreeNode.h(.hxx) (std-only header), i.e. the
  real octree storage, set algebra, collides() and visit_leaves() order.
* libfk_ref.so: the reference's tendon/get_r_info.cpp, tendon/tendon_deriv.cpp,
  tendon/solve_initial_bending.cpp and collision/collision_primitives.cpp, unmodified, compiled against
  the Eigen stand-in in oracle/ref_shim/eigen_standin (Eigen is not installed; see that header for what
  this does and does not pin).

* libvoxeloctree_ref.so, libselfcol_ref.so, libsweptvol_ref.so: the hot-path core of
  collision/VoxelOctree.{h,cpp} (add_line, find_cell, dilate_*, remove_interior_* ...), collides_self
  (collision/collision.cpp) and VoxelEnvironment::voxelize_valid_backbone_motion, cut out of the
  reference's files by function-name anchors at build time and compiled unmodified.
* liblevmar_ref.so: the levmar-2.6 the reference vendors (without LAPACK).
* libtendonrobot_ref.so: TendonRobot::shape / home_shape / is_valid (tension_shape cut out by anchors).
* librmp_ref.so: the .rmp writer and reader (RmpStreamer, LazyRmpParser) of VoxelCachedLazyPRM.cpp.
* libik_ref.so: tip_control::inverse_kinematics (inverse_kinematics_impl, Bounds) over the vendored levmar;
  libtendonrobot_ref.so also carries tip_control::Jacobian.

All are built by `make -C oracle ref` only where /root/reference exists; they travel to the GPU box
as prebuilt files.  Nothing here is needed at run time by the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
REF_SRC = "/root/reference/cpp/src"


def build():
    """(Re)build oracle/_ref when the reference sources are present; no-op otherwise."""
    if os.path.isdir(REF_SRC):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"], env=dict(os.environ, CXX="g++"))


def available():
    return all(os.path.exists(os.path.join(REF_DIR, n))
               for n in ("libtreenode_ref.so", "libfk_ref.so"))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class RefTree:
    """collision::detail::TreeNode<Ng> of the reference behind a handle."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(REF_DIR, "libtreenode_ref.so"))
            u64, vp = C.c_uint64, C.c_void_p
            L.tnref_new.restype = vp
            L.tnref_new.argtypes = [u64]
            L.tnref_free.argtypes = [vp]
            L.tnref_clone.restype = vp
            L.tnref_clone.argtypes = [vp]
            L.tnref_nblocks.restype = u64
            L.tnref_nblocks.argtypes = [vp]
            L.tnref_is_empty.argtypes = [vp]
            L.tnref_block.restype = u64
            L.tnref_block.argtypes = [vp, u64, u64, u64]
            L.tnref_set_block.argtypes = [vp, u64, u64, u64, u64]
            for n in ("tnref_union_block", "tnref_intersect_block"):
                getattr(L, n).restype = u64
                getattr(L, n).argtypes = [vp, u64, u64, u64, u64]
            for n in ("tnref_union_tree", "tnref_intersect_tree", "tnref_remove_tree",
                      "tnref_collides", "tnref_equals"):
                getattr(L, n).restype = C.c_int
                getattr(L, n).argtypes = [vp, vp]
            L.tnref_leaves.restype = u64
            L.tnref_leaves.argtypes = [vp, C.POINTER(u64), u64]
            L.tnref_check_csr.restype = C.c_int
            L.tnref_check_csr.argtypes = [vp, u64, C.POINTER(u64), C.POINTER(C.c_uint8),
                                          C.POINTER(C.c_uint8), C.POINTER(C.c_uint8),
                                          C.POINTER(u64), C.POINTER(C.c_uint8)]
            L.tnref_sets_build.restype = vp
            L.tnref_sets_build.argtypes = [u64, u64, C.POINTER(u64), C.POINTER(C.c_uint8),
                                           C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.POINTER(u64)]
            L.tnref_sets_free.argtypes = [vp]
            L.tnref_sets_check.restype = C.c_int
            L.tnref_sets_check.argtypes = [vp, vp, C.POINTER(C.c_uint8), C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, Ng, _h=None):
        self.Ng = int(Ng)
        self.h = _h if _h is not None else self.lib().tnref_new(self.Ng)
        if not self.h:
            raise ValueError("unsupported Ng %r" % (Ng,))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib().tnref_free(self.h)
            self.h = None

    def copy(self):
        return RefTree(self.Ng, self.lib().tnref_clone(self.h))

    def nblocks(self):
        return int(self.lib().tnref_nblocks(self.h))

    def is_empty(self):
        return bool(self.lib().tnref_is_empty(self.h))

    def block(self, bx, by, bz):
        return int(self.lib().tnref_block(self.h, bx, by, bz))

    def set_block(self, bx, by, bz, v):
        self.lib().tnref_set_block(self.h, bx, by, bz, int(v))

    def union_block(self, bx, by, bz, v):
        """VoxelOctree::union_block (collision/VoxelOctree.cpp:224-233): the wrapper the callers use
        forwards to TreeNode::union_block only for a non-zero value."""
        if not v:
            return self.block(bx, by, bz)
        return int(self.lib().tnref_union_block(self.h, bx, by, bz, int(v)))

    def node_union_block(self, bx, by, bz, v):
        """raw TreeNode::union_block (creates an empty leaf for v == 0)"""
        return int(self.lib().tnref_union_block(self.h, bx, by, bz, int(v)))

    def intersect_block(self, bx, by, bz, v):
        return int(self.lib().tnref_intersect_block(self.h, bx, by, bz, int(v)))

    def union_tree(self, o):
        assert self.lib().tnref_union_tree(self.h, o.h) == 0

    def intersect_tree(self, o):
        assert self.lib().tnref_intersect_tree(self.h, o.h) == 0

    def remove_tree(self, o):
        assert self.lib().tnref_remove_tree(self.h, o.h) == 0

    def collides(self, o):
        r = self.lib().tnref_collides(self.h, o.h)
        assert r >= 0
        return bool(r)

    def equals(self, o):
        return bool(self.lib().tnref_equals(self.h, o.h))

    def leaves(self):
        """(bx, by, bz, bits) arrays in the reference's visit_leaves order."""
        n = self.nblocks()
        buf = np.zeros((max(n, 1), 4), dtype=np.uint64)
        m = int(self.lib().tnref_leaves(self.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), n))
        assert m == n
        buf = buf[:n]
        return (buf[:, 0].astype(np.uint8), buf[:, 1].astype(np.uint8), buf[:, 2].astype(np.uint8),
                buf[:, 3].copy())

    def check_csr(self, off, bx, by, bz, bits):
        """Verdicts of env.collides(set_i) for sets given as CSR block lists."""
        off = np.ascontiguousarray(off, dtype=np.uint64)
        bx, by, bz = (np.ascontiguousarray(a, dtype=np.uint8) for a in (bx, by, bz))
        bits = np.ascontiguousarray(bits, dtype=np.uint64)
        n = len(off) - 1
        out = np.zeros(n, dtype=np.uint8)
        u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
        rc = self.lib().tnref_check_csr(self.h, n, off.ctypes.data_as(u64p), bx.ctypes.data_as(u8p),
                                        by.ctypes.data_as(u8p), bz.ctypes.data_as(u8p),
                                        bits.ctypes.data_as(u64p), out.ctypes.data_as(u8p))
        assert rc == 0
        return out.astype(bool)


class RefSets:
    """Cached voxel sets kept as reference TreeNodes (the form the reference's planner stores them in);
    check(env) is the OpenMP loop of VoxelCachedLazyPRM.cpp:1584-1591."""

    def __init__(self, Ng, off, bx, by, bz, bits):
        off = np.ascontiguousarray(off, dtype=np.uint64)
        bx, by, bz = (np.ascontiguousarray(a, dtype=np.uint8) for a in (bx, by, bz))
        bits = np.ascontiguousarray(bits, dtype=np.uint64)
        self.n = len(off) - 1
        u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
        self.h = RefTree.lib().tnref_sets_build(int(Ng), self.n, off.ctypes.data_as(u64p),
                                                bx.ctypes.data_as(u8p), by.ctypes.data_as(u8p),
                                                bz.ctypes.data_as(u8p), bits.ctypes.data_as(u64p))
        if not self.h:
            raise ValueError("unsupported Ng %r" % (Ng,))

    def __del__(self):
        if getattr(self, "h", None):
            RefTree.lib().tnref_sets_free(self.h)
            self.h = None

    def check(self, env, nthreads=1):
        out = np.zeros(self.n, dtype=np.uint8)
        RefTree.lib().tnref_sets_check(env.h, self.h, out.ctypes.data_as(C.POINTER(C.c_uint8)),
                                       int(nthreads))
        return out.astype(bool)


class RefFK:
    """The reference's FK arithmetic (see module docstring) for one robot spec
    (same dict as oracle.Oracle.robot)."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(REF_DIR, "libfk_ref.so"))
            L.fkref_initial_bending.restype = C.c_int
            L.fkref_shape.restype = C.c_int
            L.fkref_segment_aabox_intersect.restype = C.c_int
            L.fkref_closest_t_segment.restype = C.c_double
            cls._lib = L
        return cls._lib

    def __init__(self, spec):
        self.spec = spec
        self.N = len(spec["C"])
        self.Nc = len(spec["C"][0])
        self.Nd = len(spec["D"][0])
        self.C = np.ascontiguousarray(spec["C"], dtype=np.float64).reshape(self.N, self.Nc)
        self.D = np.ascontiguousarray(spec["D"], dtype=np.float64).reshape(self.N, self.Nd)
        self.mat = [C.c_double(float(spec[k])) for k in ("ro", "ri", "E", "nu")]

    def _hdr(self):
        return [C.c_int(self.N), C.c_int(self.Nc), C.c_int(self.Nd), _dp(self.C), _dp(self.D)]

    def r_info(self, t):
        r, rd, rdd = (np.zeros((self.N, 3)) for _ in range(3))
        self.lib().fkref_r_info(*self._hdr(), C.c_double(t), _dp(r), _dp(rd), _dp(rdd))
        return r, rd, rdd

    def deriv(self, tau, x, t, unopt=False):
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert len(tau) == self.N and len(x) == 19 + self.N
        out = np.zeros_like(x)
        self.lib().fkref_deriv(*self._hdr(), _dp(tau), *self.mat, _dp(x), C.c_double(t), _dp(out),
                               C.c_int(int(unopt)))
        return out

    def initial_bending(self, tau, s_start):
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        v0, u0 = np.zeros(3), np.zeros(3)
        it = self.lib().fkref_initial_bending(*self._hdr(), _dp(tau), *self.mat,
                                              C.c_double(float(self.spec["residual_threshold"])),
                                              C.c_double(s_start), _dp(v0), _dp(u0))
        return v0, u0, int(it)

    def shape_states(self, tau, times):
        """RK4 walk over `times` (driver is NOT the reference's, see fk_ref.cpp) with the reference's
        derivative and initial condition.  Returns states [nt][19+N] and the step count."""
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        times = np.ascontiguousarray(times, dtype=np.float64)
        st = np.zeros((len(times), 19 + self.N))
        n = self.lib().fkref_shape(*self._hdr(), _dp(tau), *self.mat,
                                   C.c_double(float(self.spec["residual_threshold"])),
                                   C.c_double(float(self.spec["dL"])), _dp(times), C.c_int(len(times)),
                                   _dp(st))
        return st, int(n)

    def closest_st_segment(self, A, B, Cc, D):
        a, b, c, d = (np.ascontiguousarray(v, dtype=np.float64) for v in (A, B, Cc, D))
        s, t = C.c_double(), C.c_double()
        self.lib().fkref_closest_st_segment(_dp(a), _dp(b), _dp(c), _dp(d), C.byref(s), C.byref(t))
        return s.value, t.value

    def segment_aabox_intersect(self, A, B, Cc, D):
        a, b, c, d = (np.ascontiguousarray(v, dtype=np.float64) for v in (A, B, Cc, D))
        return bool(self.lib().fkref_segment_aabox_intersect(_dp(a), _dp(b), _dp(c), _dp(d)))


class RefLevmar:
    """levmar-2.6 as vendored by the reference (3rdparty/levmar-2.6, compiled WITHOUT LAPACK, see
    oracle/Makefile): the optimiser behind tip_control::inverse_kinematics (tip_control.cpp:34-153).
    Callbacks are Python callables f(p: ndarray[m]) -> ndarray[n] (and J(p) -> ndarray[n][m])."""
    _lib = None
    FUNC = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_void_p)
    INFO_SZ = 10

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(REF_DIR, "liblevmar_ref.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(REF_DIR, "liblevmar_ref.so"))
            dp = C.POINTER(C.c_double)
            L.dlevmar_fdif_forw_jac_approx.restype = None
            L.dlevmar_fdif_forw_jac_approx.argtypes = [cls.FUNC, dp, dp, dp, C.c_double, dp, C.c_int, C.c_int, C.c_void_p]
            L.dlevmar_fdif_cent_jac_approx.restype = None
            L.dlevmar_fdif_cent_jac_approx.argtypes = [cls.FUNC, dp, dp, dp, C.c_double, dp, C.c_int, C.c_int, C.c_void_p]
            L.dlevmar_bc_dif.restype = C.c_int
            L.dlevmar_bc_dif.argtypes = [cls.FUNC, dp, dp, C.c_int, C.c_int, dp, dp, dp, C.c_int, dp, dp, dp, dp, C.c_void_p]
            L.dlevmar_bc_der.restype = C.c_int
            L.dlevmar_bc_der.argtypes = [cls.FUNC, cls.FUNC, dp, dp, C.c_int, C.c_int, dp, dp, dp, C.c_int, dp, dp, dp, dp, C.c_void_p]
            cls._lib = L
        return cls._lib

    @classmethod
    def _wrap(cls, f, rows=None):
        def cb(p, out, m, n, _):
            v = np.asarray(f(np.array([p[i] for i in range(m)])), dtype=np.float64).reshape(-1)
            for i in range(len(v)):
                out[i] = v[i]
        return cls.FUNC(cb)

    @classmethod
    def fdif_jac(cls, f, p, n, delta, central):
        """levmar's own finite-difference Jacobian (misc_core.c:137-211), jac[i*m+j] layout -> [n][m]"""
        p = np.array(p, dtype=np.float64)
        m = len(p)
        J, w1, w2 = np.zeros((n, m)), np.zeros(n), np.zeros(n)
        cb = cls._wrap(f)
        if central:
            cls.lib().dlevmar_fdif_cent_jac_approx(cb, _dp(p), _dp(w1), _dp(w2), delta, _dp(J), m, n, None)
        else:
            hx = np.asarray(f(p.copy()), dtype=np.float64).copy()
            cls.lib().dlevmar_fdif_forw_jac_approx(cb, _dp(p), _dp(hx), _dp(w1), delta, _dp(J), m, n, None)
        return J

    @classmethod
    def bc_dif(cls, f, p0, x, lb, ub, itmax, opts5):
        """dlevmar_bc_dif as called at tip_control.cpp:124-137.  Returns (p, info[10], rc)."""
        p = np.array(p0, dtype=np.float64)
        x = np.array(x, dtype=np.float64)
        lb, ub = np.array(lb, dtype=np.float64), np.array(ub, dtype=np.float64)
        opts, info = np.array(opts5, dtype=np.float64), np.zeros(cls.INFO_SZ)
        rc = cls.lib().dlevmar_bc_dif(cls._wrap(f), _dp(p), _dp(x), len(p), len(x), _dp(lb), _dp(ub), None,
                                      itmax, _dp(opts), _dp(info), None, None, None)
        return p, info, rc

    @classmethod
    def bc_der(cls, f, jacf, p0, x, lb, ub, itmax, opts4):
        """dlevmar_bc_der: the same driver with a caller-supplied Jacobian (dlevmar_bc_dif is this
        function with levmar's own finite-difference wrappers, lmbc_core.c)."""
        p = np.array(p0, dtype=np.float64)
        x = np.array(x, dtype=np.float64)
        lb, ub = np.array(lb, dtype=np.float64), np.array(ub, dtype=np.float64)
        opts, info = np.array(opts4, dtype=np.float64), np.zeros(cls.INFO_SZ)
        rc = cls.lib().dlevmar_bc_der(cls._wrap(f), cls._wrap(jacf), _dp(p), _dp(x), len(p), len(x), _dp(lb),
                                      _dp(ub), None, itmax, _dp(opts), _dp(info), None, None, None)
        return p, info, rc


class RefVoxelOctree:
    """collision::VoxelOctree of the reference: its own class declaration and member definitions
    (constructor, limits, cells, find_cell, add_line, add_piecewise_line, add_voxels, dilate_*,
    remove_interior_*, collides, visit_leaves), cut out of VoxelOctree.{h,cpp} at build time and compiled
    unmodified over the real TreeNode.h; Point arithmetic through the Eigen stand-in
    (oracle/ref_shim/voxeloctree_ref.cpp)."""
    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(REF_DIR, "libvoxeloctree_ref.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(REF_DIR, "libvoxeloctree_ref.so"))
            u64, vp, dp = C.c_uint64, C.c_void_p, C.POINTER(C.c_double)
            L.voref_new.restype = vp
            L.voref_new.argtypes = [u64, dp]
            L.voref_free.argtypes = [vp]
            L.voref_copy.restype = vp
            L.voref_copy.argtypes = [vp]
            for n in ("voref_nblocks", "voref_ncells"):
                getattr(L, n).restype = u64
                getattr(L, n).argtypes = [vp]
            L.voref_block.restype = u64
            L.voref_block.argtypes = [vp, u64, u64, u64]
            L.voref_set_block.argtypes = [vp, u64, u64, u64, u64]
            L.voref_union_block.restype = u64
            L.voref_union_block.argtypes = [vp, u64, u64, u64, u64]
            L.voref_set_cell.argtypes = [vp, u64, u64, u64]
            L.voref_add_line.argtypes = [vp, dp, dp]
            L.voref_add_piecewise_line.argtypes = [vp, dp, C.c_int]
            L.voref_add_voxels.argtypes = [vp, vp]
            if hasattr(L, "voref_add_sphere"):
                L.voref_add_point.argtypes = [vp, dp]
                L.voref_add_sphere.argtypes = [vp, dp, C.c_double]
                L.voref_add_capsule.argtypes = [vp, dp, dp, C.c_double]
            L.voref_find_cell.restype = C.c_int
            L.voref_find_cell.argtypes = [vp, dp, C.POINTER(C.c_int64)]
            L.voref_nearest_cell.argtypes = [vp, dp, C.POINTER(C.c_int64)]
            L.voref_collides.restype = C.c_int
            L.voref_collides.argtypes = [vp, vp]
            L.voref_dilate.argtypes = [vp, C.c_int, C.c_int]
            L.voref_dilate_sphere.argtypes = [vp, C.c_double]
            L.voref_remove_interior.argtypes = [vp, C.c_int]
            L.voref_bitmask.restype = u64
            L.voref_bitmask.argtypes = [C.c_int, C.c_int, C.c_int]
            L.voref_leaves.restype = u64
            L.voref_leaves.argtypes = [vp, C.POINTER(u64), u64]
            cls._lib = L
        return cls._lib

    def __init__(self, Ng, lim, _h=None):
        self.Ng, self.lim = int(Ng), [float(v) for v in lim]
        if _h is None:
            l = np.array(self.lim, dtype=np.float64)
            _h = self.lib().voref_new(self.Ng, _dp(l))
        if not _h:
            raise ValueError("unsupported voxel dimension %r" % (Ng,))
        self.h = _h

    def __del__(self):
        if getattr(self, "h", None):
            self.lib().voref_free(self.h)
            self.h = None

    def copy(self):
        return RefVoxelOctree(self.Ng, self.lim, self.lib().voref_copy(self.h))

    def nblocks(self):
        return int(self.lib().voref_nblocks(self.h))

    def ncells(self):
        return int(self.lib().voref_ncells(self.h))

    def block(self, bx, by, bz):
        return int(self.lib().voref_block(self.h, bx, by, bz))

    def set_block(self, bx, by, bz, v):
        self.lib().voref_set_block(self.h, bx, by, bz, int(v))

    def union_block(self, bx, by, bz, v):
        return int(self.lib().voref_union_block(self.h, bx, by, bz, int(v)))

    def set_cell(self, ix, iy, iz):
        return bool(self.lib().voref_set_cell(self.h, ix, iy, iz))

    def add_line(self, a, b):
        a, b = (np.ascontiguousarray(v, dtype=np.float64) for v in (a, b))
        self.lib().voref_add_line(self.h, _dp(a), _dp(b))

    def add_piecewise_line(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        self.lib().voref_add_piecewise_line(self.h, _dp(pts), len(pts))

    def add_voxels(self, other):
        self.lib().voref_add_voxels(self.h, other.h)

    @classmethod
    def has_primitives(cls):
        return cls.available() and hasattr(cls.lib(), "voref_add_sphere")

    def add_point(self, p):
        """VoxelOctree::add(Point) -> add_point (VoxelOctree.cpp:319-323)"""
        p = np.ascontiguousarray(p, dtype=np.float64)
        self.lib().voref_add_point(self.h, _dp(p))

    def add_sphere(self, c, r):
        """VoxelOctree::add_sphere (VoxelOctree.cpp:434-469)"""
        c = np.ascontiguousarray(c, dtype=np.float64)
        self.lib().voref_add_sphere(self.h, _dp(c), float(r))

    def add_capsule(self, a, b, r):
        """VoxelOctree::add_capsule (VoxelOctree.cpp:471-515)"""
        a, b = (np.ascontiguousarray(v, dtype=np.float64) for v in (a, b))
        self.lib().voref_add_capsule(self.h, _dp(a), _dp(b), float(r))

    def find_cell(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        cell = np.zeros(3, dtype=np.int64)
        err = self.lib().voref_find_cell(self.h, _dp(p), cell.ctypes.data_as(C.POINTER(C.c_int64)))
        return None if err else tuple(int(c) for c in cell)

    def nearest_cell(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        cell = np.zeros(3, dtype=np.int64)
        self.lib().voref_nearest_cell(self.h, _dp(p), cell.ctypes.data_as(C.POINTER(C.c_int64)))
        return tuple(int(c) for c in cell)

    def collides(self, other):
        r = self.lib().voref_collides(self.h, other.h)
        if r < 0:
            raise ValueError("voxel dimension mismatch")
        return bool(r)

    def dilate(self, num=1, use_diagonal=False):
        self.lib().voref_dilate(self.h, int(num), int(use_diagonal))

    def dilate_sphere(self, r):
        self.lib().voref_dilate_sphere(self.h, float(r))

    def remove_interior(self, keep_diagonal=True):
        self.lib().voref_remove_interior(self.h, int(keep_diagonal))

    @classmethod
    def bitmask(cls, x, y, z):
        return int(cls.lib().voref_bitmask(x, y, z))

    def export(self):
        """(bxyz uint8[n,3], bits uint64[n]) in visit_leaves order (same form as oracle Octree.export)"""
        n = self.nblocks()
        buf = np.zeros((max(n, 1), 4), dtype=np.uint64)
        m = int(self.lib().voref_leaves(self.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), n))
        assert m == n
        return buf[:n, :3].astype(np.uint8), buf[:n, 3].copy()


class RefSelfCollision:
    """collision::collides_self(CapsuleSequence) of the reference (collision/collision.cpp:6-46) with the
    structs and inline capsule tests it needs, compiled from the reference's own text
    (oracle/ref_shim/selfcol_ref.cpp)."""
    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(REF_DIR, "libselfcol_ref.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(REF_DIR, "libselfcol_ref.so"))
            dp = C.POINTER(C.c_double)
            L.scref_collides_self.restype = C.c_int
            L.scref_collides_self.argtypes = [dp, C.c_int, C.c_double]
            L.scref_capsules_collide.restype = C.c_int
            L.scref_capsules_collide.argtypes = [dp, dp, C.c_double, dp, dp, C.c_double]
            cls._lib = L
        return cls._lib

    @classmethod
    def collides_self(cls, p, r):
        p = np.ascontiguousarray(p, dtype=np.float64)
        return bool(cls.lib().scref_collides_self(_dp(p), len(p), float(r)))


class RefSweptVolume:
    """VoxelEnvironment::voxelize_valid_backbone_motion of the reference (VoxelEnvironment.cpp:207-444)
    compiled from its own text (oracle/ref_shim/sweptvol_ref.cpp).  FK, validity and interpolation are
    callbacks, as in the reference: fk(state) -> points [n][3]; valid(state, points) -> bool;
    interp(a, b, t) -> state."""
    _lib = None
    INTERP = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_double,
                         C.POINTER(C.c_double))
    FK = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int)
    VALID = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int)

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(REF_DIR, "libsweptvol_ref.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(REF_DIR, "libsweptvol_ref.so"))
            dp, u64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
            L.veref_voxelize_valid_backbone_motion.restype = C.c_int
            L.veref_voxelize_valid_backbone_motion.argtypes = [
                C.c_uint64, dp, dp, dp, dp, C.c_int, C.c_double, C.c_int, cls.INTERP, cls.FK, cls.VALID,
                C.POINTER(C.c_int), dp, dp, C.POINTER(C.c_int), u64p, u64p]
            cls._lib = L
        return cls._lib

    @classmethod
    def voxelize(cls, Ng, lim, inv_rot, a, b, rel_threshold, cap_pts, fk, valid, interp):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        S = len(a)
        lim = np.ascontiguousarray(lim, dtype=np.float64)
        R = np.ascontiguousarray(np.eye(3) if inv_rot is None else inv_rot, dtype=np.float64).reshape(3, 3)

        def interp_cb(pa, pb, s, t, out):
            v = interp(np.array([pa[i] for i in range(s)]), np.array([pb[i] for i in range(s)]), t)
            for i in range(s):
                out[i] = v[i]

        def fk_cb(ps, s, pout, cap):
            p = np.asarray(fk(np.array([ps[i] for i in range(s)])), dtype=np.float64).reshape(-1, 3)
            assert len(p) <= cap
            for i, row in enumerate(p):
                pout[3 * i], pout[3 * i + 1], pout[3 * i + 2] = row
            return len(p)

        def valid_cb(ps, s, pp, n):
            pts = np.array([pp[i] for i in range(3 * n)]).reshape(n, 3)
            return int(bool(valid(np.array([ps[i] for i in range(s)]), pts)))

        ok, nfk, t_last = C.c_int(), C.c_int(), C.c_double()
        last = np.zeros(S)
        cap_leaves = 1 << 16
        leaves = np.zeros((cap_leaves, 4), dtype=np.uint64)
        nl = C.c_uint64(cap_leaves)
        rc = cls.lib().veref_voxelize_valid_backbone_motion(
            int(Ng), _dp(lim), _dp(R), _dp(a), _dp(b), S, float(rel_threshold), int(cap_pts),
            cls.INTERP(interp_cb), cls.FK(fk_cb), cls.VALID(valid_cb), C.byref(ok), C.byref(t_last), _dp(last),
            C.byref(nfk), leaves.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(nl))
        n = int(nl.value)
        assert n <= cap_leaves
        return dict(rc=rc, is_fully_valid=bool(ok.value), t=t_last.value, last_valid=last, nsamples=nfk.value,
                    bxyz=leaves[:n, :3].astype(np.uint8), bits=leaves[:n, 3].copy())


class RefTendonRobot:
    """tendon::TendonRobot of the reference: TendonRobot.h as is, tension_shape / home_shape /
    calc_point_forces / is_valid cut out of TendonRobot.cpp at build time, over the unmodified derivative /
    initial-condition / routing sources, against the Eigen and Boost.odeint stand-ins
    (oracle/ref_shim/tendonrobot_ref.cpp)."""
    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(REF_DIR, "libtendonrobot_ref.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(REF_DIR, "libtendonrobot_ref.so"))
            L.trref_shape.restype = C.c_int
            L.trref_flags.restype = C.c_uint
            cls._lib = L
        return cls._lib

    def __init__(self, spec):
        self.spec = spec
        self.N = len(spec["C"])
        self.Nc, self.Nd = len(spec["C"][0]), len(spec["D"][0])
        self.C = np.ascontiguousarray(spec["C"], dtype=np.float64).reshape(self.N, self.Nc)
        self.D = np.ascontiguousarray(spec["D"], dtype=np.float64).reshape(self.N, self.Nd)
        self.hdr = np.array([spec["r"], spec["L"], spec["dL"], spec["ro"], spec["ri"], spec["E"], spec["nu"],
                             spec["residual_threshold"], float(bool(spec.get("enable_rotation"))),
                             float(bool(spec.get("enable_retraction")))], dtype=np.float64)
        self.lim = np.ascontiguousarray(np.stack([spec["max_tension"], spec["min_length"], spec["max_length"]],
                                                 axis=1), dtype=np.float64)

    def _args(self):
        return [_dp(self.hdr), C.c_int(self.N), C.c_int(self.Nc), C.c_int(self.Nd), _dp(self.C), _dp(self.D),
                _dp(self.lim)]

    def shape(self, state, cap=1024):
        state = np.ascontiguousarray(state, dtype=np.float64)
        t, p, R = np.zeros(cap), np.zeros((cap, 3)), np.zeros((cap, 9))
        L_i, misc = np.zeros(self.N), np.zeros(14)
        n = self.lib().trref_shape(*self._args(), _dp(state), C.c_int(cap), _dp(t), _dp(p), _dp(R), _dp(L_i),
                                   _dp(misc))
        assert n >= 0
        return dict(t=t[:n].copy(), p=p[:n].copy(), R=R[:n].reshape(n, 3, 3).transpose(0, 2, 1).copy(),
                    L=misc[0], L_i=L_i, u_i=misc[1:4].copy(), u_f=misc[4:7].copy(), v_i=misc[7:10].copy(),
                    v_f=misc[10:13].copy(), converged=bool(misc[13]))

    _release = None

    @classmethod
    def release_available(cls):
        return os.path.exists(os.path.join(REF_DIR, "libtendonrobot_ref_release.so"))

    def shape_batch(self, states, nthreads=1, release=True):
        """TendonRobot::shape over a batch in the OpenMP loop of apps/estimate_length_discretization.cpp:62-71.
        release=True: the build with the reference's own Release flags (-Ofast ..., CMakeLists.txt:66), the
        TIMING baseline of bench.py; never a parity checker.  Returns (tips[n][3], npts[n])."""
        cls = type(self)
        if release:
            if cls._release is None:
                cls._release = C.CDLL(os.path.join(REF_DIR, "libtendonrobot_ref_release.so"))
            L = cls._release
        else:
            L = self.lib()
        states = np.ascontiguousarray(states, dtype=np.float64)
        n = states.shape[0]
        tips, npts = np.zeros((n, 3)), np.zeros(n, dtype=np.int32)
        L.trref_shape_batch(*self._args(), _dp(states), C.c_longlong(n), C.c_int(int(nthreads)), _dp(tips),
                            npts.ctypes.data_as(C.POINTER(C.c_int)))
        return tips, npts

    def tip_jacobian(self, state, ps, dist):
        """tip_control::Jacobian(robot, ps, dist, state) (tip-control/tip_control.cpp:243-265), its own text;
        `dist` is a C float there.  Returns J as [3][S]."""
        state = np.ascontiguousarray(state, dtype=np.float64)
        ps = np.ascontiguousarray(ps, dtype=np.float64)
        J = np.zeros((3, len(state)))
        self.lib().trref_tip_jacobian(*self._args(), _dp(state), _dp(ps), C.c_float(dist), _dp(J))
        return J

    _ik = None

    @classmethod
    def ik_available(cls):
        return os.path.exists(os.path.join(REF_DIR, "libik_ref.so"))

    def inverse_kinematics(self, initial_state, des, max_iters=100, mu_init=1e-3, stop_JT_err_inf=1e-9,
                           stop_Dp=1e-4, stop_err=1e-4, fd_delta=1e-6):
        """tip_control::inverse_kinematics over the robot's own forward_kinematics (tip_control.cpp:29-153,
        160-184, 347-368 without the printing), its own text, driving the reference's vendored levmar.
        Returns dict(state, tip, error, iters, num_fk_calls, info[10])."""
        cls = type(self)
        if cls._ik is None:
            cls._ik = C.CDLL(os.path.join(REF_DIR, "libik_ref.so"))
            cls._ik.trref_inverse_kinematics.restype = C.c_int
        s0 = np.ascontiguousarray(initial_state, dtype=np.float64)
        des = np.ascontiguousarray(des, dtype=np.float64)
        opts = np.array([max_iters, mu_init, stop_JT_err_inf, stop_Dp, stop_err, fd_delta], dtype=np.float64)
        state, tip, misc, info = np.zeros(len(s0)), np.zeros(3), np.zeros(3), np.zeros(10)
        rc = cls._ik.trref_inverse_kinematics(*self._args(), _dp(s0), _dp(des), _dp(opts), _dp(state), _dp(tip),
                                              _dp(misc), _dp(info))
        assert rc == 0, "the reference's inverse_kinematics threw"
        return dict(state=state, tip=tip, error=misc[0], iters=int(misc[1]), num_fk_calls=int(misc[2]), info=info)

    def home_lengths(self, state):
        state = np.ascontiguousarray(state, dtype=np.float64)
        out = np.zeros(self.N)
        self.lib().trref_home_lengths(*self._args(), _dp(state), _dp(out))
        return out

    def flags(self, state):
        """bit 0 !converged, bit 1 length limits violated, bit 2 self-collision (the ingredients of
        AbstractValidityChecker::is_valid_shape, evaluated separately)"""
        state = np.ascontiguousarray(state, dtype=np.float64)
        return int(self.lib().trref_flags(*self._args(), _dp(state)))


class RefRmp:
    """The reference's .rmp writer (RmpStreamer) and reader (LazyRmpParser), VoxelCachedLazyPRM.cpp:635-1114,
    compiled from their own text (oracle/ref_shim/rmp_ref.cpp).  Roadmaps are exchanged in the dict form of
    irt_b200.read_rmp / write_rmp, with block coordinates instead of Morton keys
    (v_bxyz / e_bxyz uint8[nb][3])."""
    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(REF_DIR, "librmp_ref.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(REF_DIR, "librmp_ref.so"))
            vp, u32, u64, dp = C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_double)
            u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
            L.rmpref_writer_open.restype = vp
            L.rmpref_writer_open.argtypes = [C.c_char_p, u32, u32]
            L.rmpref_write_reference.argtypes = [vp, u64, dp]
            L.rmpref_write_vertex.argtypes = [vp, u32, dp, C.c_int, C.c_int, dp, C.c_int, u64, dp, u64, u8p, u64p]
            L.rmpref_write_edge.argtypes = [vp, u32, u32, C.c_double, C.c_int, u64, dp, u64, u8p, u64p]
            L.rmpref_writer_close.argtypes = [vp]
            L.rmpref_reader_open.restype = vp
            L.rmpref_reader_open.argtypes = [C.c_char_p]
            L.rmpref_reader_close.argtypes = [vp]
            L.rmpref_next.argtypes = [vp, C.POINTER(u32), dp, dp, C.c_int, u64p, u64]
            cls._lib = L
        return cls._lib

    @classmethod
    def write(cls, path, d):
        L = cls.lib()
        u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
        h = L.rmpref_writer_open(os.fsencode(path), int(d["n_verts"]), int(d["n_edges"]))
        assert h
        lim = np.ascontiguousarray(d["lims"], dtype=np.float64)
        Ng = int(d["Ng"])
        if d["has_voxels"]:
            assert L.rmpref_write_reference(h, Ng, _dp(lim)) == 0
        # without reference voxels RmpStreamer::try_write_voxels throws on any voxel object
        # (VoxelCachedLazyPRM.cpp:1082-1086); such a roadmap is written with null voxel pointers
        hv = bool(d["has_voxels"])

        def blocks(off, bxyz, bits, i):
            lo, hi = int(off[i]), int(off[i + 1])
            xb = np.ascontiguousarray(bxyz[lo:hi], dtype=np.uint8).reshape(-1, 3)
            bb = np.ascontiguousarray(bits[lo:hi], dtype=np.uint64)
            return hi - lo, xb, bb

        for i in range(int(d["n_verts"])):
            st = np.ascontiguousarray(d["v_state"][i], dtype=np.float64)
            tip = np.ascontiguousarray(d["v_tip"][i], dtype=np.float64)
            nb, xb, bb = blocks(d["v_off"], d["v_bxyz"], d["v_bits"], i)
            assert L.rmpref_write_vertex(h, int(d["v_index"][i]), _dp(st), len(st), int(bool(d["v_has_tip"][i])),
                                         _dp(tip), int(hv and bool(d["v_has_vox"][i])), Ng, _dp(lim), nb,
                                         xb.ctypes.data_as(u8p), bb.ctypes.data_as(u64p)) == 0
        for i in range(int(d["n_edges"])):
            nb, xb, bb = blocks(d["e_off"], d["e_bxyz"], d["e_bits"], i)
            assert L.rmpref_write_edge(h, int(d["e_src"][i]), int(d["e_dst"][i]), float(d["e_weight"][i]),
                                       int(hv and bool(d["e_has_vox"][i])), Ng, _dp(lim), nb,
                                       xb.ctypes.data_as(u8p), bb.ctypes.data_as(u64p)) == 0
        L.rmpref_writer_close(h)

    @classmethod
    def read(cls, path, cap_state=64, cap_leaves=1 << 16):
        """list of ("vertex", index, state, tip or None, leaves or None) / ("edge", src, dst, weight, leaves)"""
        L = cls.lib()
        h = L.rmpref_reader_open(os.fsencode(path))
        assert h
        out = []
        hdr = (C.c_uint32 * 6)()
        vals, state = np.zeros(4), np.zeros(cap_state)
        leaves = np.zeros((cap_leaves, 4), dtype=np.uint64)
        while True:
            t = L.rmpref_next(h, hdr, _dp(vals), _dp(state), cap_state, leaves.ctypes.data_as(C.POINTER(C.c_uint64)),
                              cap_leaves)
            assert t >= 0, "reference parser failed"
            if t == 3:
                break
            lv = leaves[:hdr[5]].copy() if hdr[4] else None
            if t == 1:
                out.append(("vertex", int(hdr[0]), state[:hdr[2]].copy(), vals[1:4].copy() if hdr[3] else None, lv))
            else:
                out.append(("edge", int(hdr[0]), int(hdr[1]), float(vals[0]), lv))
        L.rmpref_reader_close(h)
        return out
