import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import irt_b200, irt_b200.workloads as wl
from bench import knn_edges_gpu
ctx = irt_b200.Context(0)
spec = wl.robot_b(0.003)
rb = irt_b200.Robot(ctx, spec)
g = wl.workspace_grid(spec)
grid = irt_b200.make_grid(g["Ng"], g["lim"])
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
st = wl.sample_states(spec, nv, stream=200)
k = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 10
pairs = knn_edges_gpu(torch, st, spec, k, torch.device("cuda"))
store = irt_b200.SetStore(ctx, grid)
for rep in range(2):
    t0 = time.perf_counter()
    info = store.voxelize_edges_indexed(rb, irt_b200.make_space(), st, pairs)
    dt = time.perf_counter() - t0
    print("indexed edges %d: %.3f s, %.2f Medges/s, samples/edge %.2f, blocks/edge %.1f" % (len(pairs), dt, len(pairs) / dt / 1e6, info["nsamples"].mean(), store.num_blocks / len(pairs)), flush=True)
if "--pairwise" in sys.argv:
    a, b = st[pairs[:, 0]].copy(), st[pairs[:, 1]].copy()
    t0 = time.perf_counter(); store.voxelize_edges(rb, irt_b200.make_space(), a, b); dt = time.perf_counter() - t0
    print("pairwise edges %d: %.3f s" % (len(pairs), dt))
vs = irt_b200.SetStore(ctx, grid)
for rep in range(2):
    t0 = time.perf_counter(); vs.voxelize_vertices(rb, st); dt = time.perf_counter() - t0
    print("vertices %d: %.3f s" % (nv, dt))
