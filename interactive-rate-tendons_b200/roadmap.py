"""Batch side of motion_planning::VoxelCachedLazyPRM, B200-native.

Mirrors the batch entry points of the reference planner (motion-planning/VoxelCachedLazyPRM.h:
495-520) whose loop bodies are the hot path:

  createRoadmap(N, ...)            VoxelCachedLazyPRM.cpp:1380-1561 (sampling + voxel caches)
  precomputeVertexVoxelCache()     :1687-1734     precomputeEdgeVoxelCache()   :1736-1782
  precomputeVertexValidity()       :1563-1598     precomputeEdgeValidity()     :1600-1647
  clearValidity()                  :1656-1663

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
    words = np.ascontiguousarray(all_words, dtype=np.uint32).reshape(world, w)
    out = np.zeros(n, dtype=bool)
    for r in range(world):
        lo, hi = shard_range(n, r, world, align)
        if hi > lo:
            out[lo:hi] = unpack_verdicts(words[r], hi - lo)
    return out


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
        self.fused_gather = True     # multi-GPU sweeps: fuse the verdict all-gather into K3 (peer memory)
        self._xchg = {}

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
        self.clearValidity()

    def createRoadmap(self, n_vertices, sampler, connect, max_rounds=64):
        """Rejection-sample `n_vertices` valid configurations (VoxelCachedLazyPRM.cpp:1415-1455:
        a sample is kept iff is_valid_shape) and connect them with `connect(states) -> edges`
        (the reference's connectionStrategy_, host side).  `sampler(count, round) -> states`."""
        kept = []
        total = 0
        for rnd in range(max_rounds):
            need = n_vertices - total
            if need <= 0:
                break
            cand = sampler(int(need * 1.25) + 64, rnd)
            out = self.robot.shape_batch(cand, want=("flags",))
            ok = (out["flags"] & INVALID_MASK) == 0
            good = cand[ok][:need]
            kept.append(good)
            total += len(good)
        states = np.concatenate(kept, axis=0)[:n_vertices]
        self.set_roadmap(states, connect(states))
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
        lo, hi = self.shard(n_total)
        w = shard_words(n_total, self.world)
        use_cuda = torch.cuda.is_available()
        dev = torch.device("cuda", self.ctx.device) if use_cuda else torch.device("cpu")
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
        if not self._have_vcache:
            self.precomputeVertexVoxelCache()
        n = len(self.states)
        collides = self._sweep(self.vertex_store, n, self.vertex_flags)
        invalid = self._gather_flags(self.vertex_flags, n, INVALID_MASK)
        self.vertex_validity = np.where(~collides & ~invalid, VALIDITY_TRUE, VALIDITY_UNKNOWN).astype(np.uint8)
        return self.vertex_validity

    def precomputeEdgeValidity(self):
        """computeEdgeValidity (VoxelCachedLazyPRM.cpp:2620-2631): is_fully_valid and no hit."""
        if not self._have_ecache:
            self.precomputeEdgeVoxelCache()
        n = len(self.edges)
        collides = self._sweep(self.edge_store, n, self.edge_flags)
        invalid = self._gather_flags(self.edge_flags, n, FLAG_PARTIAL)
        self.edge_validity = np.where(~collides & ~invalid, VALIDITY_TRUE, VALIDITY_UNKNOWN).astype(np.uint8)
        return self.edge_validity

    def precomputeValidity(self):
        self.precomputeVertexValidity()
        self.precomputeEdgeValidity()

    def clearValidity(self):
        self.vertex_validity = np.zeros(len(self.states), dtype=np.uint8)
        self.edge_validity = np.zeros(len(self.edges), dtype=np.uint8)

    def _gather_flags(self, local_flags, n_total, mask):
        """validity flags of all shards as a global bool array (1 bit per item on the wire)."""
        import torch
        lo, hi = self.shard(n_total)
        w = shard_words(n_total, self.world)
        bits = np.zeros(w * 32, dtype=np.uint8)
        bits[:hi - lo] = (np.asarray(local_flags) & mask) != 0
        words = np.packbits(bits, bitorder="little").view(np.uint32)
        if self.dist is None or self.world == 1:
            allw = words
        else:
            use_cuda = torch.cuda.is_available() and self.dist.get_backend() == "nccl"
            dev = torch.device("cuda", self.ctx.device) if use_cuda else torch.device("cpu")
            t = torch.from_numpy(words.view(np.int32).copy()).to(dev)
            allw = gather_verdict_words(t, self.dist).cpu().numpy().view(np.uint32)
        return assemble_verdicts(allw, n_total, self.world)
