"""Batch side of motion_planning::VoxelCachedLazyPRM, B200-native.

Mirrors the batch entry points of the reference planner (motion-planning/VoxelCachedLazyPRM.h:
495-520) whose loop bodies are the hot path:

  createRoadmap(N, ...)            VoxelCachedLazyPRM.cpp:1380-1561 (sampling + voxel caches)
  precomputeVertexVoxelCache()     :1687-1734     precomputeEdgeVoxelCache()   :1736-1782
  precomputeVertexValidity()       :1563-1598     precomputeEdgeValidity()     :1600-1647
  clearValidity()                  :1656-1663     clear{Vertex,Edge}VoxelCache() :1814-1829

The graph search, nearest-neighbour structures, IK and file formats of the reference planner
stay on the host and are not re-implemented here; the roadmap is (states, edge index pairs).

Multi-GPU (one process per GPU): vertices and edges are sharded by contiguous index range
(boundaries aligned to 64 so verdict words never straddle shards), every rank voxelises and
checks only its shard, the environment grid is replicated, and ONLY the collision-verdict
bitmask is exchanged.  On GPUs the exchange is fused into K3: every warp stores its verdict word
straight into all peers' copies of the gathered array over NVLink (CUDA IPC peer memory) and a
per-rank epoch flag replaces the collective (VerdictExchange / irt_check_sets_allgather_dev);
torch.distributed (NCCL) only publishes the IPC handles once.  The plain all_gather of uint32 words
(NCCL, or gloo in the CPU tests of the host logic) remains as the fallback path (fused_gather=False).
"""
import numpy as np

from . import (INVALID_MASK, FLAG_PARTIAL, Env, SetStore, make_space, shard_range, unpack_verdicts)

VALIDITY_UNKNOWN = 0  # VoxelCachedLazyPRM.h:598-601
VALIDITY_TRUE = 1
# CreateRoadmapOption, VoxelCachedLazyPRM.h:468-479
LazyRoadmap, VoxelizeVertices, ValidateVertices, VoxelizeEdges, ValidateEdges = 0x0, 0x1, 0x2, 0x4, 0x8


def shard_words(n, world, align=64):
    """uint32 verdict words every shard contributes to the all_gather (equal on all ranks)."""
    per = -(-n // world)
    per = -(-per // align) * align
    return per // 32


def gather_verdict_words(local_words, dist=None, group=None):
    """all_gather of the per-shard verdict words.  `local_words` is a torch tensor (CUDA for
    NCCL, CPU for gloo) of identical length on every rank; returns the concatenation."""
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_words
    import torch
    world = dist.get_world_size(group)
    out = torch.empty(world * local_words.numel(), dtype=local_words.dtype, device=local_words.device)
    dist.all_gather_into_tensor(out, local_words.contiguous(), group=group)
    return out


def assemble_verdicts(all_words, n, world, align=64):
    """global bool[n] (True = collides) from the gathered words of `world` aligned shards."""
    w = shard_words(n, world, align)
    if n == 0:
        return np.zeros(0, dtype=bool)
    words = np.ascontiguousarray(all_words, dtype=np.uint32).reshape(world, w)
    if world == 1:              # one shard: the unpacked bits are the answer, no second copy
        return unpack_verdicts(words[0], n)
    out = np.zeros(n, dtype=bool)
    for r in range(world):
        lo, hi = shard_range(n, r, world, align)
        if hi > lo:
            out[lo:hi] = unpack_verdicts(words[r], hi - lo)
    return out


def _validity_table(collides, invalid):
    """uint8 table, VALIDITY_TRUE (1) where the item neither collides nor carries an invalidating flag, else
    VALIDITY_UNKNOWN (0).  `collides` is this sweep's own freshly unpacked table and is reused as the output
    (two in-place byte passes over it instead of three temporaries; 10 M edges per tick)."""
    c = collides.view(np.uint8) if collides.flags.writeable else collides.astype(np.uint8)
    np.bitwise_or(c, invalid.view(np.uint8), out=c)
    np.bitwise_xor(c, 1, out=c)
    return c


class LockstepFk:
    """Batching layer under k IK solvers that run side by side: every solver asks for the tip and
    the finite-difference tip Jacobian of ONE state at a time, like the reference's ikController_ does through
    fk_wrap (tip-control/tip_control.cpp:92-122); the requests of all solvers that are still running are answered
    by ONE K1 launch (irt_fk_tip_jacobian_batch over n (2S+1) states).  The solvers themselves stay what they
    are -- e.g. the reference's levmar driver, sequential host code -- and run in one host thread each."""

    def __init__(self, robot, n_workers, mode=None, delta=1e-6):
        import threading
        from . import JAC_LEVMAR_CENTRAL
        self.robot, self.mode, self.delta = robot, (JAC_LEVMAR_CENTRAL if mode is None else mode), delta
        self.cv = threading.Condition()
        self.active, self.pending, self.results, self.generation = n_workers, {}, {}, 0
        self.batches = self.evaluations = 0

    def _flush(self):       # called with the lock held by the thread that completed the round
        ids = sorted(self.pending)
        states = np.stack([self.pending[i] for i in ids])
        tips, J = self.robot.tip_jacobian_batch(states, mode=self.mode, delta=self.delta)
        self.results = {i: (tips[k].copy(), J[k].copy()) for k, i in enumerate(ids)}
        self.pending = {}
        self.batches += 1
        self.evaluations += len(ids)
        self.generation += 1
        self.cv.notify_all()

    def evaluate(self, worker, state):
        """tip [3] and Jacobian [3][S] at `state`; blocks until every running solver has asked"""
        with self.cv:
            self.pending[worker] = np.array(state, dtype=np.float64)
            gen = self.generation
            if len(self.pending) == self.active:
                self._flush()
            else:
                while self.generation == gen:
                    self.cv.wait()
            return self.results[worker]

    def finished(self, worker):
        with self.cv:
            self.active -= 1
            if self.active > 0 and len(self.pending) == self.active:
                self._flush()


def solve_ik_lockstep(robot, solver, starts, request, mode=None, delta=1e-6):
    """runs solver(start, request, fk) for every start state side by side, fk(state) -> (tip, J) answered in
    lockstep batches.  Returns (final states [k][S], LockstepFk with its batch statistics)."""
    import threading
    starts = np.asarray(starts, dtype=np.float64)
    k = len(starts)
    fk = LockstepFk(robot, k, mode, delta)
    out, errors = [None] * k, [None] * k

    def run(i):
        try:
            out[i] = np.asarray(solver(starts[i].copy(), np.asarray(request, dtype=np.float64),
                                       lambda st, _i=i: fk.evaluate(_i, st)), dtype=np.float64)
        except BaseException as e:   # noqa: a failing solver must not leave the others waiting for it
            errors[i] = e
        finally:
            fk.finished(i)

    threads = [threading.Thread(target=run, args=(i,)) for i in range(k)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    return np.stack(out), fk


class VoxelCachedLazyPRM:
    """Roadmap with device-resident voxel caches and batch validity sweeps."""

    def __init__(self, ctx, robot, grid, space=None, rank=0, world=1, dist=None):
        self.ctx, self.robot, self.grid = ctx, robot, grid
        self.space = space or make_space()
        self.rank, self.world, self.dist = rank, world, dist
        self.states = np.zeros((0, robot.state_size))
        self.edges = np.zeros((0, 2), dtype=np.int64)
        self.vertex_store = SetStore(ctx, grid)
        self.edge_store = SetStore(ctx, grid)
        self.env = Env(ctx, grid)
        self.vertex_flags = self.edge_flags = None
        self.tips = None
        self.vertex_validity = np.zeros(0, dtype=np.uint8)
        self.edge_validity = np.zeros(0, dtype=np.uint8)
        self._have_vcache = self._have_ecache = False
        self.seed = 20220801         # of the default sampler (the reference seeds from std::random_device)
        self.max_nearest_neighbors = 5   # magic::DEFAULT_NEAREST_NEIGHBORS_LAZY (VoxelCachedLazyPRM.cpp:125)
        self.range = None            # maxDistance_; None = configurePlannerRange's 20 % of the maximum extent
        self.fused_gather = True     # multi-GPU sweeps: fuse the verdict all-gather into K3 (peer memory)
        self._xchg = {}
        self._sampler_calls = 0      # createRoadmap calls so far: every call draws from its own Philox stream
        # lazy-path consumers (computeVertexValidity / computeEdgeValidity / constructSolution)
        self._vertex_swept = self._edge_swept = False   # a full sweep ran since the last clearValidity
        self.vertex_removed = np.zeros(0, dtype=bool)   # removeVertices / removeEdge of constructSolution
        self.edge_removed = np.zeros(0, dtype=bool)
        self._adj = None
        self._adj_extra, self._adj_extra_count = {}, 0   # edges appended since the CSR adjacency was built
        self._bufs = {}                                  # growth buffers of the arrays queries append to (_grow)
        self.lookups = {"vertex": 0, "edge": 0, "sweeps": 0}
        # vertices / edges that joined the roadmap after the last sweep with VALIDITY_UNKNOWN (addMilestone's lazy
        # connections): checked one at a time when a query first asks, like the reference's lazy checks
        self._vertex_unchecked, self._edge_unchecked = set(), set()
        self.max_uncached = 4096     # more items than this without a cached set: the next sweep rebuilds the cache
        self._flag_tables = {}

    def _exchange(self, store, slot_words):
        """one exchange buffer per (store, slot size); creating it is collective (IPC handle all-gather)"""
        from . import VerdictExchange
        key = (id(store), int(slot_words))
        if key not in self._xchg:
            self._xchg[key] = VerdictExchange(self.ctx, self.rank, self.world, max(int(slot_words), 1), self.dist)
        return self._xchg[key]

    # ---- roadmap content ----------------------------------------------------------------
    def set_roadmap(self, states, edges):
        self.states = np.ascontiguousarray(states, dtype=np.float64)
        self.edges = np.ascontiguousarray(edges, dtype=np.int64).reshape(-1, 2)
        self._have_vcache = self._have_ecache = False
        self._adj = None
        self.vertex_removed = np.zeros(len(self.states), dtype=bool)
        self.edge_removed = np.zeros(len(self.edges), dtype=bool)
        self.clearValidity()

    def distance(self, a, b):
        """distanceFunction / motionCost of the compound space of Problem::create_space_information
        (Problem.cpp:101-163): |d tau| + (ext / 4 pi) * SO2 distance + (2 ext / L) * |d retraction|,
        ext = |max_tension|.  a: [S] or [m][S], b: [n][S] -> [n] or [m][n]."""
        d = self.robot.spec
        N = len(d["C"])
        ext = float(np.linalg.norm(np.asarray(d["max_tension"], dtype=np.float64)[:N]))
        a = np.asarray(a, dtype=np.float64)[..., None, :]
        b = np.asarray(b, dtype=np.float64)
        out = np.sqrt(((a[..., :N] - b[..., :N]) ** 2).sum(-1))
        k = N
        if d.get("enable_rotation"):
            dr = np.abs(a[..., k] - b[..., k])
            out = out + (ext / (4 * np.pi)) * np.where(dr > np.pi, 2 * np.pi - dr, dr)
            k += 1
        if d.get("enable_retraction"):
            out = out + (2 * ext / d["L"]) * np.abs(a[..., k] - b[..., k])
        return out

    def maximum_extent(self):
        """CompoundStateSpace::getMaximumExtent = sum of weight_i * extent_i"""
        d = self.robot.spec
        ext = float(np.linalg.norm(np.asarray(d["max_tension"], dtype=np.float64)[:len(d["C"])]))
        return ext * (1.0 + (0.25 if d.get("enable_rotation") else 0.0) + (2.0 if d.get("enable_retraction") else 0.0))

    def k_bounded_neighbors(self, v, k=None, bound=None):
        """KBoundedStrategy(k, range) over nn_ (VoxelCachedLazyPRM.cpp:1329-1345): the k nearest milestones of
        v -- v itself is one of them, it is in nn_ by then -- no further away than the range.  Exact."""
        k = self.max_nearest_neighbors if k is None else k
        bound = (self.range or 0.2 * self.maximum_extent()) if bound is None else bound
        d = self.distance(self.states[v], self.states)
        order = np.lexsort((np.arange(len(d)), d))[:k]
        return order[d[order] <= bound]

    def random_states(self, count, rnd=0):
        """TendonRobot::random_state (tendon/TendonRobot.cpp:219-247) for a batch, seeded: tensions
        U[0, max_tension], rotation U[-pi, pi], retraction U[0, L], in state order"""
        d = self.robot.spec
        # one stream per (createRoadmap call, round): growing the roadmap never replays old candidates
        g = np.random.Generator(np.random.Philox(key=[self.seed, (self._sampler_calls << 32) + rnd]))
        N = len(d["C"])
        cols = [g.uniform(0.0, d["max_tension"][j], count) for j in range(N)]
        if d.get("enable_rotation"):
            cols.append(g.uniform(-np.pi, np.pi, count))
        if d.get("enable_retraction"):
            cols.append(g.uniform(0.0, d["L"], count))
        return np.ascontiguousarray(np.stack(cols, axis=1))

    def createRoadmap(self, n_vertices, sampler=None, connect=None, max_rounds=64, opt=VoxelizeVertices):
        """Brings the roadmap up to `n_vertices` (VoxelCachedLazyPRM.cpp:1380-1561), in the reference's order:
        rejection-sample the new vertices, add them all, connect every new vertex, then voxelise / validate
        the new edges as one batch and remove the invalid ones.  `opt` is the reference's
        CreateRoadmapOption bit set (VoxelCachedLazyPRM.h:468-479) and only pertains to what is added:
          LazyRoadmap       nothing is checked
          VoxelizeVertices  a sample is kept iff is_valid_shape                    (.cpp:1415-1443)
          ValidateVertices  ... and its voxels miss the environment
          VoxelizeEdges     new edges that are not fully valid are removed         (.cpp:1505-1551)
          ValidateEdges     ... and new edges whose swept volume hits the environment
        sampler(count, round) -> states (default: random_states = TendonRobot::random_state);
        connect(states) -> int64[n_edges][2] for the whole vertex list (default: k_bounded_neighbors of every
        new vertex, new vertices in index order, no duplicates -- connectionStrategy_ + getEdge, .cpp:1486-1500).
        Candidates are judged as whole batches (one FK / voxelise / check call per round), on every rank
        alike; the edge work is sharded like every other sweep."""
        validate_verts, validate_edges = bool(opt & ValidateVertices), bool(opt & ValidateEdges)
        voxelize_verts = validate_verts or bool(opt & VoxelizeVertices)
        voxelize_edges = validate_edges or bool(opt & VoxelizeEdges)
        nv0, ne0 = len(self.states), len(self.edges)
        if n_vertices <= nv0:
            return self      # "Graph is already at or bigger than N, skipping roadmap creation"
        sampler = sampler or self.random_states
        self._sampler_calls += 1
        kept, total = [], 0
        # addMilestone's duplicate rejection (the was_added loop of createRoadmap): a candidate equal to an
        # existing or already kept state is dropped and resampled
        seen = set(r.tobytes() for r in np.ascontiguousarray(self.states, dtype=np.float64))
        scratch = SetStore(self.ctx, self.grid) if validate_verts else None
        for rnd in range(max_rounds):
            need = n_vertices - nv0 - total
            if need <= 0:
                break
            cand = np.ascontiguousarray(sampler(int(need * 1.25) + 64, rnd), dtype=np.float64)
            ok = np.ones(len(cand), dtype=bool)
            if validate_verts:
                flags, _ = scratch.voxelize_vertices(self.robot, cand)
                ok = ((flags & INVALID_MASK) == 0) & ~scratch.check(self.env)
            elif voxelize_verts:
                ok = (self.robot.shape_batch(cand, want=("flags",))["flags"] & INVALID_MASK) == 0
            good = []
            for r in cand[ok]:
                key = r.tobytes()
                if key not in seen:
                    seen.add(key)
                    good.append(r)
                    if len(good) == need:
                        break
            good = np.asarray(good, dtype=np.float64).reshape(-1, cand.shape[1])
            kept.append(good)
            total += len(good)
        if total < n_vertices - nv0:
            raise RuntimeError("createRoadmap: rejection sampling kept %d of %d vertices in %d rounds"
                               % (total, n_vertices - nv0, max_rounds))
        old_vv, old_ev = self.vertex_validity, self.edge_validity
        had_vcache, had_ecache = self._have_vcache, self._have_ecache
        states = np.concatenate([self.states.reshape(-1, self.robot.state_size)] + kept, axis=0)[:n_vertices]
        self.states = np.ascontiguousarray(states)
        # ---- empty edges ----
        if connect is not None:
            new_edges = np.ascontiguousarray(connect(self.states), dtype=np.int64).reshape(-1, 2)
            if ne0:     # keep what was there, add what is new (getEdge: an edge can exist already)
                have = set(map(tuple, np.sort(self.edges, axis=1).tolist()))
                new_edges = np.array([e for e in new_edges.tolist() if tuple(sorted(e)) not in have],
                                     dtype=np.int64).reshape(-1, 2)
        else:
            have = set(map(tuple, np.sort(self.edges, axis=1).tolist()))
            out = []
            for v in range(nv0, n_vertices):
                for n in self.k_bounded_neighbors(v).tolist():
                    key = (min(v, n), max(v, n))
                    if n != v and key not in have:      # connectVertices: a == b makes no edge
                        have.add(key)
                        out.append((v, n))
            new_edges = np.array(out, dtype=np.int64).reshape(-1, 2)
        self.edges = np.concatenate([self.edges, new_edges], axis=0)
        self._have_vcache = self._have_ecache = False
        # ---- voxelise / validate the new edges, remove the invalid ones ----
        if voxelize_edges and len(new_edges):
            self.precomputeEdgeVoxelCache()
            n = len(self.edges)
            bad = self._gather_flags(self.edge_flags, n, FLAG_PARTIAL).copy()   # the table itself is cached
            if validate_edges:
                bad |= self._sweep(self.edge_store, n, self.edge_flags)
            bad[:ne0] = False
            self.edges = np.ascontiguousarray(self.edges[~bad])
            self._have_ecache = False
        # ---- caches and validity words of the grown roadmap ----
        if voxelize_verts or had_vcache:
            self.precomputeVertexVoxelCache()
        if voxelize_edges or had_ecache:
            self.precomputeEdgeVoxelCache()
        self.vertex_validity = np.zeros(len(self.states), dtype=np.uint8)
        self.vertex_validity[:nv0] = old_vv[:nv0]
        self.edge_validity = np.zeros(len(self.edges), dtype=np.uint8)
        self.edge_validity[:ne0] = old_ev[:ne0]
        if validate_verts:
            self.vertex_validity[nv0:] = VALIDITY_TRUE
        if validate_edges:
            self.edge_validity[ne0:] = VALIDITY_TRUE
        self._adj = None
        self._vertex_swept = self._edge_swept = False
        vr, er = self.vertex_removed, self.edge_removed
        self.vertex_removed = np.zeros(len(self.states), dtype=bool)
        self.vertex_removed[:min(nv0, len(vr))] = vr[:nv0]
        self.edge_removed = np.zeros(len(self.edges), dtype=bool)
        self.edge_removed[:min(ne0, len(er))] = er[:ne0]
        return self

    def shard(self, n):
        return shard_range(n, self.rank, self.world)

    # ---- voxel caches ---------------------------------------------------------------------
    def precomputeVertexVoxelCache(self):
        lo, hi = self.shard(len(self.states))
        self.vertex_flags, self.tips = self.vertex_store.voxelize_vertices(self.robot, self.states[lo:hi])
        self._have_vcache = True
        return self.vertex_flags

    def precomputeEdgeVoxelCache(self):
        lo, hi = self.shard(len(self.edges))
        e = self.edges[lo:hi]
        info = self.edge_store.voxelize_edges_indexed(self.robot, self.space, self.states, e)
        self.edge_flags = info["flags"]
        self.edge_info = info
        self._have_ecache = True
        return info

    def precomputeVoxelCache(self):
        self.precomputeVertexVoxelCache()
        self.precomputeEdgeVoxelCache()

    def clearVertexVoxelCache(self):
        """clearVertexVoxelCache (VoxelCachedLazyPRM.cpp:1814-1818): the cached sets go, validity words stay"""
        self.vertex_store = SetStore(self.ctx, self.grid)
        self.vertex_flags = self.tips = None
        self._have_vcache = False
        self._xchg = {k: v for k, v in self._xchg.items() if k[0] == id(self.edge_store)}

    def clearEdgeVoxelCache(self):
        """clearEdgeVoxelCache (VoxelCachedLazyPRM.cpp:1820-1824)"""
        self.edge_store = SetStore(self.ctx, self.grid)
        self.edge_flags = None
        self._have_ecache = False
        self._xchg = {k: v for k, v in self._xchg.items() if k[0] == id(self.vertex_store)}

    def clearVoxelCache(self):
        self.clearVertexVoxelCache()
        self.clearEdgeVoxelCache()

    # ---- environment ------------------------------------------------------------------------
    def setEnvironment(self, blocks):
        """Replace the obstacle grid (replicated on every rank) and forget old verdicts.  The
        reference has no live-update API: it rebuilds the validators (Problem.h:175-216) and
        calls clearValidity()."""
        self.env.update(blocks)
        self.clearValidity()

    # ---- validity sweeps ------------------------------------------------------------------------
    def _sweep(self, store, n_total, flags):
        import torch
        if n_total == 0:        # a roadmap without edges (or vertices): nothing collides
            return np.zeros(0, dtype=bool)
        lo, hi = self.shard(n_total)
        w = shard_words(n_total, self.world)
        # the verdict words live where the store's K3 writes them: on this rank's GPU (a store may name another
        # device through `words_device`; the product's SetStore does not)
        dev = torch.device(getattr(store, "words_device", None) or ("cuda:%d" % self.ctx.device))
        use_cuda = dev.type == "cuda"
        fused = (use_cuda and self.world > 1 and self.dist is not None and self.dist.is_initialized()
                 and self.dist.get_backend() == "nccl" and self.fused_gather)
        if fused:
            # K3 stores its verdict words straight into every peer's gathered array (NVLink P2P)
            x = self._exchange(store, w)
            stream = torch.cuda.current_stream(dev).cuda_stream
            allw = x.check(store, self.env, 0, hi - lo, stream=stream or None)
            if not stream:
                self.ctx.synchronize()
            collides = assemble_verdicts(allw.cpu().numpy().view(np.uint32), n_total, self.world)
            if x.status():
                raise RuntimeError("verdict exchange: peer %d never arrived" % (x.status() - 1))
            return collides
        words = torch.zeros(max(w, 1), dtype=torch.int32, device=dev)
        if hi > lo:
            stream = torch.cuda.current_stream(dev).cuda_stream if use_cuda else None
            store.check_dev(self.env, words, 0, hi - lo, stream=stream)
            if not stream:  # legacy default stream: the library ran on its own stream
                self.ctx.synchronize()
        allw = gather_verdict_words(words, self.dist)
        collides = assemble_verdicts(allw.cpu().numpy().view(np.uint32), n_total, self.world)
        return collides

    def precomputeVertexValidity(self):
        """vertexValidity = VALIDITY_TRUE iff the shape is valid and its voxels miss the
        environment (computeVertexValidity, VoxelCachedLazyPRM.cpp:2607-2618)."""
        if not self._have_vcache or self._tail(self.vertex_flags, len(self.states)) > self.max_uncached:
            self.precomputeVertexVoxelCache()
        n = len(self.states) - self._tail(self.vertex_flags, len(self.states))   # the sets the store holds
        collides = self._sweep(self.vertex_store, n, self.vertex_flags)
        invalid = self._gather_flags(self.vertex_flags, n, INVALID_MASK)
        self.vertex_validity = _validity_table(collides, invalid)
        if n < len(self.states):     # vertices that joined since (roadmapIk / addMilestone): one small batch
            self.vertex_validity = np.concatenate([self.vertex_validity, self._check_vertices_now(np.arange(n, len(self.states)))])
        self._vertex_unchecked.clear()
        self._vertex_swept = True
        self.lookups["sweeps"] += 1
        return self.vertex_validity

    def precomputeEdgeValidity(self):
        """computeEdgeValidity (VoxelCachedLazyPRM.cpp:2620-2631): is_fully_valid and no hit."""
        if not self._have_ecache or self._tail(self.edge_flags, len(self.edges)) > self.max_uncached:
            self.precomputeEdgeVoxelCache()
        n = len(self.edges) - self._tail(self.edge_flags, len(self.edges))
        collides = self._sweep(self.edge_store, n, self.edge_flags)
        invalid = self._gather_flags(self.edge_flags, n, FLAG_PARTIAL)
        self.edge_validity = _validity_table(collides, invalid)
        if n < len(self.edges):
            self.edge_validity = np.concatenate([self.edge_validity, self._check_edges_now(np.arange(n, len(self.edges)))])
        self._edge_unchecked.clear()
        self._edge_swept = True
        self.lookups["sweeps"] += 1
        return self.edge_validity

    def precomputeValidity(self):
        self.precomputeVertexValidity()
        self.precomputeEdgeValidity()

    def clearValidity(self):
        """clearValidity (VoxelCachedLazyPRM.cpp:1656-1663): every validity word back to VALIDITY_UNKNOWN"""
        self.vertex_validity = np.zeros(len(self.states), dtype=np.uint8)
        self.edge_validity = np.zeros(len(self.edges), dtype=np.uint8)
        self._vertex_swept = self._edge_swept = False
        self._vertex_unchecked.clear()
        self._edge_unchecked.clear()

    def _tail(self, flags, n_items):
        """items that joined the roadmap after its voxel cache was built (roadmapIk / addMilestone, one rank): they
        have no cached set; sweeps and look-ups answer them from small scratch batches until the cache is rebuilt"""
        if self.world != 1 or flags is None:
            return 0
        return max(0, n_items - len(flags))

    def _check_vertices_now(self, ids):
        """voxelizeVertex + collides for a few vertices (VoxelCachedLazyPRM.cpp:2607-2618, 2803-2837): uint8 validity"""
        ids = np.asarray(ids, dtype=np.int64)
        scratch = SetStore(self.ctx, self.grid)
        flags, _ = scratch.voxelize_vertices(self.robot, self.states[ids])
        ok = ((flags & INVALID_MASK) == 0) & ~scratch.check(self.env)
        return ok.astype(np.uint8) * np.uint8(VALIDITY_TRUE)

    def _check_edges_now(self, ids):
        """voxelizeEdge + collides for a few edges (VoxelCachedLazyPRM.cpp:2620-2631, 2879-2902): uint8 validity"""
        ids = np.asarray(ids, dtype=np.int64)
        scratch = SetStore(self.ctx, self.grid)
        info = scratch.voxelize_edges_indexed(self.robot, self.space, self.states, self.edges[ids])
        ok = ((info["flags"] & FLAG_PARTIAL) == 0) & ~scratch.check(self.env)
        return ok.astype(np.uint8) * np.uint8(VALIDITY_TRUE)

    # ---- lazy-path consumers: what the planner's query side calls --------------------
    def computeVertexValidity(self, v):
        """computeVertexValidity (VoxelCachedLazyPRM.cpp:2607-2618).  The reference checks ONE vertex here
        (voxelise if needed + collides); with the verdict words of a full sweep on the host this is a table
        look-up.  The first query after clearValidity() triggers the sweep (one K3 launch over every cached
        set + the gather), every later one only reads the table."""
        if not self._vertex_swept:
            self.precomputeVertexValidity()
        self.lookups["vertex"] += 1
        if v in self._vertex_unchecked:      # joined the roadmap after the sweep, never checked: the reference's lazy check
            self._vertex_unchecked.discard(v)
            self.vertex_validity[v] = self._check_vertices_now([v])[0]
            self.lookups["single"] = self.lookups.get("single", 0) + 1
        return bool(self.vertex_validity[v] & VALIDITY_TRUE)

    def computeEdgeValidity(self, e):
        """computeEdgeValidity (VoxelCachedLazyPRM.cpp:2620-2631): is_fully_valid (no IRT_FLAG_PARTIAL) and the
        swept volume misses the environment -- a look-up into the gathered words, as above."""
        if not self._edge_swept:
            self.precomputeEdgeValidity()
        self.lookups["edge"] += 1
        if e in self._edge_unchecked:
            self._edge_unchecked.discard(e)
            self.edge_validity[e] = self._check_edges_now([e])[0]
            self.lookups["single"] = self.lookups.get("single", 0) + 1
        return bool(self.edge_validity[e] & VALIDITY_TRUE)

    def _adjacency(self):
        """CSR adjacency (neighbour vertex, edge id) of the undirected roadmap.  Edges appended one query at a time
        (roadmapIk / addMilestone) go to per-vertex lists next to it (_adj_extra) instead of costing a rebuild each;
        it is rebuilt when the edge list was replaced or the appended share passes 1/64.  Use _neighbors(v)."""
        m = len(self.edges)
        base = -1 if self._adj is None else self._adj[3]
        if base < 0 or base > m or m - base != self._adj_extra_count or m - base > max(1024, base // 64):
            n = len(self.states)
            src = np.concatenate([self.edges[:, 0], self.edges[:, 1]])
            dst = np.concatenate([self.edges[:, 1], self.edges[:, 0]])
            eid = np.concatenate([np.arange(m), np.arange(m)])
            order = np.argsort(src, kind="stable")
            ptr = np.zeros(n + 1, dtype=np.int64)
            ptr[1:] = np.cumsum(np.bincount(src, minlength=n))
            self._adj = (ptr, dst[order], eid[order], m)
            self._adj_extra, self._adj_extra_count = {}, 0
            if len(self.vertex_removed) != n:
                self.vertex_removed = np.zeros(n, dtype=bool)
            if len(self.edge_removed) != m:
                self.edge_removed = np.zeros(m, dtype=bool)
        return self._adj[:3]

    def _neighbors(self, v):
        """(neighbour vertices, edge ids) of v: its CSR row plus the edges appended since the CSR was built"""
        ptr, nbr, eid = self._adjacency()
        if v + 1 < len(ptr):
            lo, hi = int(ptr[v]), int(ptr[v + 1])
            bn, be = nbr[lo:hi], eid[lo:hi]
        else:                                   # a vertex that joined after the CSR was built
            bn = be = np.zeros(0, dtype=np.int64)
        extra = self._adj_extra.get(int(v))
        if extra:
            ex = np.asarray(extra, dtype=np.int64)
            bn, be = np.concatenate([bn, ex[:, 0]]), np.concatenate([be, ex[:, 1]])
        return bn, be

    def astarSearch(self, start, goal):
        """astarSearch (VoxelCachedLazyPRM.cpp:2950-2976): A* over the current graph (removed vertices / edges
        left out) with the motion cost as edge weight and as heuristic.  Host code in the reference (Boost.Graph)
        and here; it never touches the device.  Returns the vertex list start..goal or None."""
        import heapq
        self._adjacency()
        if self.vertex_removed[start] or self.vertex_removed[goal]:
            return None
        gs = self.states[goal]
        dist = {start: 0.0}
        prev = {start: start}
        done = set()
        heap = [(float(self.distance(self.states[start], gs[None])[0]), start)]
        while heap:
            _, u = heapq.heappop(heap)
            if u in done:
                continue
            if u == goal:
                path = [goal]
                while path[-1] != start:
                    path.append(prev[path[-1]])
                return path[::-1]
            done.add(u)
            nb_u, eid_u = self._neighbors(u)
            ok = ~(self.edge_removed[eid_u] | self.vertex_removed[nb_u])
            vs = nb_u[ok]
            if not len(vs):
                continue
            w = self.distance(self.states[u], self.states[vs])
            h = self.distance(gs, self.states[vs])
            for v, wi, hi_ in zip(vs.tolist(), w.tolist(), h.tolist()):
                nd = dist[u] + wi
                if v not in done and nd < dist.get(v, np.inf):
                    dist[v] = nd
                    prev[v] = u
                    heapq.heappush(heap, (nd + hi_, v))
        return None

    def edge_index(self, a, b):
        nb_a, eid_a = self._neighbors(a)
        hit = np.nonzero(nb_a == b)[0]
        return int(eid_a[hit[0]]) if len(hit) else -1

    def constructSolution(self, start, goal):
        """constructSolution (VoxelCachedLazyPRM.cpp:2689-2771): A* path, then the lazy validity checks along it --
        every intermediate vertex first (ALL invalid ones are removed), then the edges from the goal side (the
        FIRST invalid one is removed).  Returns the vertex list of a fully validated path, or None when
        something was removed (the caller searches again) or no path exists."""
        if start == goal:
            return [start]
        path = self.astarSearch(start, goal)
        if path is None:
            return None
        bad = [v for v in path[-2:0:-1] if not self.computeVertexValidity(v)]   # from the goal side, ends excluded
        if bad:
            self.vertex_removed[bad] = True          # removeVertices(milestonesToRemove)
            return None
        for i in range(len(path) - 1, 0, -1):        # edge (path[i-1], path[i]), goal side first
            e = self.edge_index(path[i - 1], path[i])
            if not self.computeEdgeValidity(e):
                self.edge_removed[e] = True          # removeEdge(e): the first invalid edge only
                return None
        return path

    def solveWithRoadmap(self, start, goal, max_iterations=10000):
        """the remove-and-retry loop of solveWithRoadmap (VoxelCachedLazyPRM.cpp:1977-2096) around
        constructSolution: search again while something was removed and the two ends stay connected.
        Returns (path or None, number of constructSolution calls)."""
        for it in range(1, max_iterations + 1):
            nv, ne = int(self.vertex_removed.sum()), int(self.edge_removed.sum())
            path = self.constructSolution(start, goal)
            if path is not None:
                return path, it
            if nv == int(self.vertex_removed.sum()) and ne == int(self.edge_removed.sum()):
                return None, it                       # nothing removed: A* found no path
        return None, max_iterations

    def restoreRemoved(self):
        """forget the removals of constructSolution / roadmapIk (the reference reloads its roadmap instead)"""
        self.vertex_removed[:] = False
        self.edge_removed[:] = False

    # ---- roadmapIk as a batch (VoxelCachedLazyPRM.cpp:3095-3420) -----------------------------------------------
    def nearest_tips(self, request, k):
        """nnTip_->nearestK: the k vertices whose cached tip positions are nearest to `request` (exact)"""
        if self.tips is None or len(self.tips) != len(self.states):
            self.precomputeVertexVoxelCache()
        d = np.linalg.norm(self.tips - np.asarray(request, dtype=np.float64)[None], axis=1)
        d = np.where(self.vertex_removed[:len(d)] if len(self.vertex_removed) == len(d) else False, np.inf, d)
        order = np.lexsort((np.arange(len(d)), d))[:k]
        return order[np.isfinite(d[order])]

    def _enforce_bounds(self, x):
        """CompoundStateSpace::enforceBounds of the space of Problem.cpp:101-163: tensions clamped to
        [0, max_tension], rotation wrapped into [-pi, pi] (SO2StateSpace::enforceBounds: fmod + one wrap), retraction
        clamped to [0, L]"""
        d = self.robot.spec
        N = len(d["C"])
        x = np.array(x, dtype=np.float64)
        x[:N] = np.clip(x[:N], 0.0, np.asarray(d["max_tension"], dtype=np.float64)[:N])
        kk = N
        if d.get("enable_rotation"):
            v = np.fmod(x[kk], 2.0 * np.pi)
            x[kk] = v + 2.0 * np.pi if v < -np.pi else (v - 2.0 * np.pi if v >= np.pi else v)
            kk += 1
        if d.get("enable_retraction"):
            x[kk] = min(max(x[kk], 0.0), d["L"])
        return x

    def _find_state(self, x):
        """tryAddToGraph (VoxelCachedLazyPRM.cpp:2777-2801): the vertex whose state equals x (si_->equalStates:
        every component within 2 epsilon), else -1"""
        if not len(self.states):
            return -1
        same = np.all(np.abs(self.states - x[None]) <= 2.0 * np.finfo(np.float64).eps, axis=1)
        if len(self.vertex_removed) == len(same):
            same &= ~self.vertex_removed
        hit = np.nonzero(same)[0]
        return int(hit[0]) if len(hit) else -1

    def _out_degree(self, v):
        nb_v, eid_v = self._neighbors(v)
        return int((~(self.edge_removed[eid_v] | self.vertex_removed[nb_v])).sum())

    def _ik_nearest(self, x, own, ik_neighbor, accurate, removed):
        """the would-be edge sources of an IK result x (`nearest` of VoxelCachedLazyPRM.cpp:3238-3246, 3340-3349,
        3470-3478): its IK neighbour only, or with RMAP_IK_ACCURATE connectionStrategy_(vertex) -- the k nearest
        milestones within the range, the (temporarily added) vertex itself among them since it is in nn_ by then
        -- plus the IK neighbour when it is not one of them.  The result's own vertex is `own` when the state was
        already in the roadmap, else -1 (a vertex that only exists for the duration of the query)."""
        if not accurate:
            return [int(ik_neighbor)]
        bound = self.range or 0.2 * self.maximum_extent()
        d = self.distance(x, self.states)
        d = np.where(removed, np.inf, d)
        ids = np.arange(len(d))
        if own < 0:                       # the temporary vertex: distance 0, the highest index
            d, ids = np.append(d, 0.0), np.append(ids, -1)
        order = np.lexsort((np.arange(len(d)), d))[:self.max_nearest_neighbors]
        out = [int(ids[j]) for j in order if d[j] <= bound]
        if int(ik_neighbor) not in out:
            out.append(int(ik_neighbor))
        return out

    def _grow(self, name, rows):
        """append rows to the array attribute `name` in amortised O(rows): the attribute is a view of the leading rows
        of a buffer with spare capacity (a 10M-edge roadmap is 160 MB of index pairs; a query appends a handful)"""
        cur = getattr(self, name)
        rows = np.asarray(rows, dtype=cur.dtype).reshape((-1,) + cur.shape[1:])
        n, k = len(cur), len(rows)
        buf = self._bufs.get(name)
        if buf is None or cur.base is not buf or n + k > len(buf) or (n and cur.ctypes.data != buf.ctypes.data):
            buf = np.empty((max(2 * n, n + k + 1024),) + cur.shape[1:], dtype=cur.dtype)
            buf[:n] = cur
            self._bufs[name] = buf
        buf[n:n + k] = rows
        setattr(self, name, buf[:n + k])

    def _append_vertex(self, state, tip, validity=VALIDITY_TRUE):
        v = len(self.states)
        self._adjacency()
        self._grow("states", np.asarray(state, dtype=np.float64)[None])
        self._grow("vertex_validity", [validity])
        if not validity & VALIDITY_TRUE:
            self._vertex_unchecked.add(v)
        self._grow("vertex_removed", [False])
        self._grow("tips", np.asarray(tip, dtype=np.float64)[None])
        return v                          # the voxel cache lacks the new vertex: sweeps treat it as the tail (_tail)

    def _append_edges(self, pairs, validity):
        self._adjacency()
        pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
        e0 = len(self.edges)
        self._grow("edges", pairs)
        self._grow("edge_validity", np.full(len(pairs), validity, dtype=np.uint8))
        self._grow("edge_removed", np.zeros(len(pairs), dtype=bool))
        if not validity & VALIDITY_TRUE:
            self._edge_unchecked.update(range(e0, e0 + len(pairs)))
        for j, (a, b) in enumerate(pairs.tolist()):
            self._adj_extra.setdefault(a, []).append((b, e0 + j))
            self._adj_extra.setdefault(b, []).append((a, e0 + j))
        self._adj_extra_count += len(pairs)

    def addMilestone(self, state, connect=True):
        """addMilestone(state, connect) (VoxelCachedLazyPRM.cpp:1854-1885): the vertex of an equal state when the
        roadmap has one (tryAddToGraph), else a new vertex with VALIDITY_UNKNOWN, lazily connected -- edges of unknown
        validity -- to its connection-strategy neighbours (the k nearest milestones within the range; the vertex itself
        is not in nn_ yet).  Nothing is voxelised here: queries check the new items when they first ask for them.
        Returns (vertex, was_added)."""
        if self.world != 1:
            raise NotImplementedError("addMilestone runs on the full roadmap of one rank")
        x = np.ascontiguousarray(state, dtype=np.float64)
        self._adjacency()
        v = self._find_state(x)
        if v >= 0:
            return v, False
        conn = []
        if connect and len(self.states):
            d = np.where(self.vertex_removed, np.inf, self.distance(x, self.states))
            order = np.lexsort((np.arange(len(d)), d))[:self.max_nearest_neighbors]
            conn = [int(j) for j in order if d[j] <= (self.range or 0.2 * self.maximum_extent())]
        if self.tips is None or len(self.tips) != len(self.states):
            self.precomputeVertexVoxelCache()
        tip = self.robot.shape_batch(x[None], want=("tip",))["tip"][0]
        v = self._append_vertex(x, tip, VALIDITY_UNKNOWN)
        if conn:
            self._append_edges([[v, n] for n in conn], VALIDITY_UNKNOWN)
        return v, True

    def chainedPlan(self, start_state, requests, tolerance, k, solver, auto_add=True, accurate=False, lazy_add=False,
                    mode=None, delta=1e-6, common_start=False):
        """The milestone loop of apps/roadmap_chained_plan.cpp:535-679: for every requested tip position roadmapIk
        gives the goal configuration, start and goal join the roadmap as milestones (solvePrep,
        VoxelCachedLazyPRM.cpp:2978-3025), solveWithRoadmap plans between them, and the plan's last state is where the
        next milestone starts (unless common_start).  Everything the loop asks of the device is a batch or a look-up:
        the IK's lockstep FK launches, one validity call and one until-invalid call per request, table look-ups
        along the A* path, and single checks only for the few items that joined the roadmap since the last sweep.
        Returns one dict per request: ik (roadmapIk's result), status ('exact' / 'empty': no path, the plan stays
        put), path (vertex list or None), plan (states), searches, n_vertices / n_edges of the roadmap the search ran on,
        tip_error of the plan's last state."""
        current = np.ascontiguousarray(start_state, dtype=np.float64)
        first = current.copy()
        out = []
        for request in requests:
            request = np.asarray(request, dtype=np.float64)
            ik = self.roadmapIk(request, tolerance, k, solver, auto_add=auto_add, mode=mode, delta=delta,
                                accurate=accurate, lazy_add=lazy_add)
            if ik is None:
                raise RuntimeError("no IK results returned")
            start_v, _ = self.addMilestone(first if common_start else current)
            goal_v, _ = self.addMilestone(np.ascontiguousarray(ik["controls"], dtype=np.float64))
            n_v, n_e = len(self.states), len(self.edges)      # the roadmap the search runs on
            path, searches = self.solveWithRoadmap(start_v, goal_v)
            if path is None:     # "Could not reach goal, no solution": make plan to just stay put
                plan, status = self.states[[start_v]].copy(), "empty"
            else:
                plan, status = self.states[path].copy(), "exact"
            tip = self.robot.shape_batch(plan[-1:], want=("tip",))["tip"][0]
            out.append(dict(ik=ik, status=status, path=path, plan=plan, searches=searches, start_vertex=start_v,
                            goal_vertex=goal_v, n_vertices=n_v, n_edges=n_e,
                            tip_error=float(np.linalg.norm(tip - request))))
            if not common_start:
                current = plan[-1].copy()
        return out

    def roadmapIk(self, request, tolerance, k, solver, auto_add=False, mode=None, delta=1e-6, accurate=False,
                  lazy_add=False):
        """roadmapIk(request, tolerance, k, opt) of the reference (VoxelCachedLazyPRM.cpp:3095-3565) with its
        per-neighbour loop turned into batches; auto_add / accurate / lazy_add are RMAP_IK_AUTO_ADD / _ACCURATE /
        _LAZY_ADD (VoxelCachedLazyPRM.h:364-381):
          1. the k nearest VALID neighbours in tip space (invalid ones are removed and the query repeated,
             .cpp:3112-3137; validity = table look-ups),
          2. the k IK problems solved side by side, their FK / Jacobian requests answered in lockstep by single
             K1 launches (solve_ik_lockstep; `solver(start, request, fk)` is the reference's ikController_),
          3. ONE FK + is_valid_shape + voxelise + collides call over the k results (.cpp:3182-3192),
          4. ONE voxelize_until_invalid call over every would-be edge source -> result that the branches below can
             ask for (the reference computes them one at a time where it needs them; an edge it would not have
             reached costs device time here but changes nothing), sources per result as _ik_nearest says,
          5. the reference's order of acceptance, walked over those tables:
             without auto_add the FIRST neighbour (nearest first) whose result is valid and within tolerance; failing
             that the closest valid result; failing that the closest last-valid state of the edges source -> result
             ("stepping them backwards", .cpp:3298-3428);
             with auto_add the first result within tolerance that a fully valid edge connects to a valid source
             (the result joins the roadmap with that edge, .cpp:3212-3292); failing that the closest last-valid state
             over the edges of ALL results joins the roadmap, connected to its source and, lazily, to its
             connection-strategy neighbours, whose edges are validated unless lazy_add (.cpp:3430-3565).
        Sources found invalid on the way are removed like removeVertices does (vertex_removed).  Returns
        dict(controls, tip_position, neighbor, error, index, vertex, lockstep_batches, ...) or None (where the
        reference would index an empty list: no valid source anywhere)."""
        if self.world != 1:
            raise NotImplementedError("roadmapIk runs on the full roadmap of one rank")
        request = np.asarray(request, dtype=np.float64)
        while True:
            nb = self.nearest_tips(request, k)
            if not len(nb):
                raise RuntimeError("roadmapIk(): No neighbors were able to be found")
            bad = [int(v) for v in nb if not self.computeVertexValidity(int(v))]
            if not bad:
                break
            self._adjacency()
            self.vertex_removed[bad] = True
        starts = self.states[nb]
        finals, fk = solve_ik_lockstep(self.robot, solver, starts, request, mode, delta)
        scratch = SetStore(self.ctx, self.grid)
        flags, tips = scratch.voxelize_vertices(self.robot, finals)
        collides = scratch.check(self.env)
        valid = ((flags & INVALID_MASK) == 0) & ~collides
        err = np.linalg.norm(tips - request[None], axis=1)
        info = dict(lockstep_batches=fk.batches, fk_requests=fk.evaluations, neighbors=nb, errors=err, valid=valid)
        m = len(nb)

        def result(i, controls=None, tip=None, error=None, neighbor=None, **kw):
            return dict(controls=finals[i] if controls is None else controls,
                        tip_position=tips[i] if tip is None else tip, neighbor=starts[i] if neighbor is None else neighbor,
                        error=float(err[i] if error is None else error), index=int(i), vertex=int(nb[i]), **info, **kw)

        if not auto_add:
            for i in range(m):
                if valid[i] and err[i] < tolerance:
                    return result(i)
            ok = np.nonzero(valid)[0]
            if len(ok):          # "All IKs rejected, returning closest valid one"
                return result(int(ok[np.argmin(err[ok])]), accepted=False)

        # ---- the would-be edges: sources per result in the reference's order, validity look-ups and removals
        # replayed on a copy first so that ONE until-invalid batch can answer every edge the walk may ask for
        self._adjacency()
        bounded = np.stack([self._enforce_bounds(x) for x in finals])   # space->enforceBounds(state)
        own = [self._find_state(x) for x in bounded]                    # addMilestone: was_added == (own < 0)
        removed = self.vertex_removed.copy()
        first_pass = [i for i in range(m) if err[i] < tolerance] if auto_add else list(range(m))

        def plan_result(i):
            rows = []                  # (source, is_self, valid) in the reference's order
            for s in self._ik_nearest(bounded[i], own[i], nb[i], accurate, removed):
                is_self = s < 0 or s == own[i]
                ok_s = bool(valid[i]) if s < 0 else self.computeVertexValidity(s)
                if not ok_s and not is_self:
                    removed[s] = True
                rows.append((s, is_self, ok_s))
            return rows

        plan1, plan3 = {}, {}
        for i in first_pass:
            if not (auto_add and own[i] >= 0 and self._out_degree(own[i]) > 0):   # else returned before any edge
                plan1[i] = plan_result(i)
        jobs = [(i, s) for i, rows in plan1.items() for (s, is_self, ok_s) in rows if ok_s and not (auto_add and is_self)]
        if auto_add:                   # the fallback asks again for the results the loop above left without sources
            for i in range(m):
                if own[i] < 0 and not any(ok_s and not is_self for (_, is_self, ok_s) in plan1.get(i, [])):
                    plan3[i] = plan_result(i)
                    jobs += [(i, s) for (s, _, ok_s) in plan3[i] if ok_s and (i, s) not in jobs]
        pe_of = {}
        if jobs:
            src = np.stack([bounded[i] if s < 0 else self.states[s] for i, s in jobs])
            dst = np.stack([finals[i] for i, _ in jobs])
            est = SetStore(self.ctx, self.grid)
            pe = est.voxelize_edges_until_invalid(self.robot, self.space, src, dst, self.env)
            last_valid = self._interpolate(src, dst, pe["t_last"][:, None])
            lt = self.robot.shape_batch(last_valid, want=("tip",))["tip"]       # last_backbone.back()
            lerr = np.linalg.norm(lt - request[None], axis=1)
            for j, key in enumerate(jobs):
                pe_of[key] = dict(fully_valid=(pe["flags"][j] & FLAG_PARTIAL) == 0, last_valid=last_valid[j],
                                  tip=lt[j], error=float(lerr[j]))

        def walk(rows):
            """the valid sources of a result in order; applies the removals the reference makes on the way"""
            for s, is_self, ok_s in rows:
                if not ok_s:
                    if not is_self:
                        self.vertex_removed[s] = True
                    continue
                yield s, is_self

        if not auto_add:
            # "All IKs are in collision, stepping them backwards"
            best = None
            for i in range(m):
                for s, _ in walk(plan1[i]):
                    e = pe_of[(i, s)]
                    if best is None or e["error"] < best[0]["error"]:
                        best = (e, i, s)
            if best is None:
                return None
            e, i, s = best
            return result(i, controls=e["last_valid"], tip=e["tip"], error=e["error"], accepted=False,
                          stepped_back=True, source=int(s))

        nearest = {}                   # res.nearest: the sources whose partial edge was computed
        for i in first_pass:
            if own[i] >= 0 and self._out_degree(own[i]) > 0:
                return result(i, added_vertex=None, already_in_roadmap=True)
            nearest[i] = []
            for s, is_self in walk(plan1[i]):
                if is_self:            # "State already in the roadmap, returning it."
                    v = own[i]
                    if v < 0:          # the vertex addMilestone created stays, without an edge
                        v = self._append_vertex(bounded[i], tips[i])
                    else:
                        self.tips[v] = tips[i]
                        self.vertex_validity[v] = VALIDITY_TRUE
                    return result(i, added_vertex=v if own[i] < 0 else None, already_in_roadmap=own[i] >= 0)
                nearest[i].append(s)
                if pe_of[(i, s)]["fully_valid"]:          # connect source -- result vertex
                    v = own[i]
                    if v < 0:
                        v = self._append_vertex(bounded[i], tips[i])
                    else:
                        self.tips[v] = tips[i]
                        self.vertex_validity[v] = VALIDITY_TRUE
                    self._append_edges([[int(s), v]], VALIDITY_TRUE)
                    return result(i, added_vertex=v if own[i] < 0 else None, source=int(s))
        # "All IKs rejected, finding closest collision-free connection"
        best = None
        for i in range(m):
            if not nearest.get(i):
                if own[i] >= 0:    # "already part of the roadmap, no need to connect it"
                    continue
                nearest[i] = [s for s, _ in walk(plan3[i])]
            for s in nearest[i]:
                e = pe_of[(i, s)]
                if best is None or e["error"] < best[0]["error"]:
                    best = (e, i, s)
        if best is None:
            return None
        e, i, s = best
        src_state = bounded[i] if s < 0 else self.states[s].copy()
        v = self._find_state(e["last_valid"])
        added = v < 0
        if added:
            # addMilestone(state, true): connected lazily to its connection-strategy neighbours (itself not yet in nn_)
            d = np.where(self.vertex_removed, np.inf, self.distance(e["last_valid"], self.states))
            order = np.lexsort((np.arange(len(d)), d))[:self.max_nearest_neighbors]
            conn = [int(j) for j in order if d[j] <= (self.range or 0.2 * self.maximum_extent())]
            v = self._append_vertex(e["last_valid"], tips[i])      # the reference caches the IK RESULT's tip here
            if conn:
                self._append_edges([[v, n] for n in conn], VALIDITY_UNKNOWN)
        else:
            self.tips[v] = tips[i]
            self.vertex_validity[v] = VALIDITY_TRUE
        if s >= 0 and v != s:
            eid = self.edge_index(int(s), v)
            if eid < 0:
                self._append_edges([[int(s), v]], VALIDITY_TRUE)
            else:
                self.edge_validity[eid] = VALIDITY_TRUE
            if not lazy_add:           # every edge of the vertex is validated now, the invalid ones removed
                mine = [int(x) for x in self._neighbors(v)[1] if not self.edge_removed[x]]
                todo = [x for x in mine if not self.edge_validity[x] & VALIDITY_TRUE]
                if todo:
                    good = self._check_edges_now(todo) != 0
                    self._edge_unchecked.difference_update(todo)
                    self.edge_validity[np.asarray(todo)[good]] = VALIDITY_TRUE
                    self.edge_removed[np.asarray(todo)[~good]] = True
        return result(i, controls=e["last_valid"], tip=e["tip"], error=e["error"], neighbor=src_state, accepted=False,
                      stepped_back=True, source=int(s), added_vertex=v if added else None)

    def _interpolate(self, a, b, t):
        """OMPL compound interpolate of the space of Problem.cpp:101-163 (RealVector linear, SO2 shortest arc)"""
        d = self.robot.spec
        out = a + (b - a) * t
        if d.get("enable_rotation"):
            k = len(d["C"])
            diff = b[:, k] - a[:, k]
            wrap = np.abs(diff) > np.pi
            dd = np.where(diff > 0, 2 * np.pi - diff, -2 * np.pi - diff)
            v = a[:, k] - dd * t[:, 0]
            v = np.where(v > np.pi, v - 2 * np.pi, np.where(v < -np.pi, v + 2 * np.pi, v))
            out[:, k] = np.where(wrap, v, out[:, k])
        return out

    def _gather_flags(self, local_flags, n_total, mask):
        """validity flags of all shards as a global bool array (1 bit per item on the wire)."""
        import torch
        if n_total == 0:
            return np.zeros(0, dtype=bool)
        # the flags only change when a voxel cache is rebuilt: every tick between two rebuilds reuses the table
        cached = self._flag_tables.get(mask)
        if cached is not None and cached[0] is local_flags:
            return cached[1]
        out = self._gather_flags_uncached(local_flags, n_total, mask)
        self._flag_tables[mask] = (local_flags, out)
        return out

    def _gather_flags_uncached(self, local_flags, n_total, mask):
        import torch
        if self.dist is None or self.world == 1:     # one rank holds everything: nothing to pack or exchange
            return (np.asarray(local_flags) & mask) != 0
        lo, hi = self.shard(n_total)
        w = shard_words(n_total, self.world)
        bits = np.zeros(w * 32, dtype=np.uint8)
        bits[:hi - lo] = (np.asarray(local_flags) & mask) != 0
        words = np.packbits(bits, bitorder="little").view(np.uint32)
        if self.dist is None or self.world == 1:
            allw = words
        else:
            # the flag bits travel over the process group's own transport (NCCL: device tensors)
            dev = torch.device("cuda", self.ctx.device) if self.dist.get_backend() == "nccl" else torch.device("cpu")
            t = torch.from_numpy(words.view(np.int32).copy()).to(dev)
            allw = gather_verdict_words(t, self.dist).cpu().numpy().view(np.uint32)
        return assemble_verdicts(allw, n_total, self.world)
