"""Multi-GPU check of the verdict exchange fused into K3 (run under torchrun, one rank per GPU):
fused peer-memory gather == NCCL all_gather == unsharded single-GPU result, over several replanning
ticks (epoch / double-buffer logic), plus device-timed sweeps of both paths."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import irt_b200  # noqa: E402
import irt_b200.workloads as wl  # noqa: E402
from irt_b200.roadmap import VoxelCachedLazyPRM, shard_words, gather_verdict_words  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dist.all_reduce(torch.zeros(1, device=dev))
    torch.cuda.synchronize()
    os.dup2(saved, 1)
    ctx = irt_b200.Context(local)
    spec = wl.robot_b(0.003)
    rb = irt_b200.Robot(ctx, spec)
    g = wl.workspace_grid(spec)
    grid = irt_b200.make_grid(g["Ng"], g["lim"], g["inv_rot"])
    nv = int(os.environ.get("MGPU_VERTICES", "20000"))
    states = wl.sample_states(spec, nv, stream=300)
    sys.path.insert(0, ROOT)
    from bench import knn_edges_gpu
    edges = knn_edges_gpu(torch, states, spec, 8, dev)
    env0 = wl.dense_to_morton_blocks(wl.lung_like_env_dense(spec, g))

    prm = VoxelCachedLazyPRM(ctx, rb, grid, rank=rank, world=world, dist=dist)
    prm.set_roadmap(states, edges)
    prm.precomputeVoxelCache()
    solo = VoxelCachedLazyPRM(ctx, rb, grid, rank=0, world=1, dist=None)   # unsharded reference on this GPU
    solo.set_roadmap(states, edges)
    solo.precomputeVoxelCache()
    rng = np.random.default_rng(11)
    env = env0
    for tick in range(6):
        if tick:
            env = wl.toggle_blob(env, g, rng.uniform(-0.1, 0.1, 3) + np.array([0, 0, 0.1]), 0.012)
        for p in (prm, solo):
            p.setEnvironment(env)
        prm.fused_gather = True
        ev_f, vv_f = prm.precomputeEdgeValidity().copy(), prm.precomputeVertexValidity().copy()
        prm.clearValidity()
        prm.fused_gather = False
        ev_n, vv_n = prm.precomputeEdgeValidity().copy(), prm.precomputeVertexValidity().copy()
        ev_s, vv_s = solo.precomputeEdgeValidity(), solo.precomputeVertexValidity()
        assert np.array_equal(ev_f, ev_n) and np.array_equal(vv_f, vv_n), "fused != NCCL at tick %d" % tick
        assert np.array_equal(ev_f, ev_s) and np.array_equal(vv_f, vv_s), "sharded != unsharded at tick %d" % tick
        assert 0 < ev_f.sum() < len(ev_f)
    # timing of one edge sweep, both paths, device events on the launching stream
    ne = len(edges)
    lo, hi = prm.shard(ne)
    w = shard_words(ne, world)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    x = prm._exchange(prm.edge_store, w)
    d_words = torch.zeros(max(w, 1), dtype=torch.int32, device=dev)

    def fused():
        return x.check(prm.edge_store, prm.env, 0, hi - lo, stream=stream.cuda_stream)

    def nccl():
        prm.edge_store.check_dev(prm.env, d_words, 0, hi - lo, stream=stream.cuda_stream)
        return gather_verdict_words(d_words, dist)

    res = {}
    for name, fn in (("fused", fused), ("nccl", nccl)):
        for _ in range(5):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        a.record(stream)
        for _ in range(reps):
            out = fn()
        b.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = float(t.item())
        res[name + "_sum"] = int(out.to(torch.int64).abs().sum().item())
    assert x.status() == 0
    assert torch.equal(fused().cpu(), nccl().cpu())
    # stress of the flag protocol: back-to-back sweeps against three rotating environments (so the slot of this
    # parity holds a DIFFERENT environment's words from two sweeps ago), the gathered table read on the device
    # right behind every sweep, no host synchronisation in between; a peer store that became visible after its
    # flag would leave a stale word
    envs, expect = [], []
    for i in range(3):
        e = irt_b200.Env(ctx, grid)
        e.update(wl.toggle_blob(env0, g, np.array([0.03 * (i - 1), 0.0, 0.1]), 0.015 + 0.004 * i))
        envs.append(e)
        prm.edge_store.check_dev(e, d_words, 0, hi - lo, stream=stream.cuda_stream)
        expect.append(gather_verdict_words(d_words, dist).clone())
    assert not torch.equal(expect[0], expect[1]) and not torch.equal(expect[1], expect[2])
    bad = torch.zeros((), dtype=torch.int64, device=dev)
    sweeps = int(os.environ.get("MGPU_STRESS_SWEEPS", "600"))
    for k in range(sweeps):
        out = x.check(prm.edge_store, envs[k % 3], 0, hi - lo, stream=stream.cuda_stream)
        bad += (out != expect[k % 3]).sum()
    torch.cuda.synchronize()
    assert x.status() == 0
    assert int(bad.item()) == 0, "stale verdict words in %d places over %d sweeps" % (int(bad.item()), sweeps)
    if rank == 0:
        print("MGPU_OK world=%d vertices=%d edges=%d fused %.4f ms nccl %.4f ms per edge sweep"
              % (world, nv, ne, res["fused"], res["nccl"]), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
