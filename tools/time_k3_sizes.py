"""K3 over sub-ranges of a real 10M-edge store: fixed cost vs streaming cost of a sweep (plain and fused self-exchange)"""
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
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
st = wl.sample_states(spec, nv, stream=200)
pairs = knn_edges_gpu(torch, st, spec, 17, torch.device("cuda"))
store = irt_b200.SetStore(ctx, grid)
store.voxelize_edges_indexed(rb, irt_b200.make_space(), st, pairs)
env = irt_b200.Env(ctx, grid)
env.update(wl.dense_to_morton_blocks(wl.lung_like_env_dense(spec, g)))
n = store.num_sets
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
words = torch.zeros((n + 31) // 32 + 64, dtype=torch.int32, device="cuda")
flush = torch.zeros(512 << 17, dtype=torch.int64, device="cuda")
x = irt_b200.VerdictExchange(ctx, 0, 1, (n + 63) // 64 * 2)
def timeit(fn, reps=20):
    for _ in range(3):
        flush.sum(); fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.sum(); a.record(s); fn(); b.record(s)
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3
for m in (0, 1024, 157696, 315392, 630784, 1261568, 2523136, 5046272, n):
    m = min(m, n)
    t_plain = timeit(lambda: store.check_dev(env, words, 0, m, stream=s.cuda_stream)) if m else 0.0
    t_fused = timeit(lambda: x.check(store, env, 0, m, stream=s.cuda_stream))
    by = store.algorithmic_bytes(0, m) if m else 0
    print("sets %9d  bytes %7.1f MB  plain %7.1f us (%5.0f GB/s)  fused(self) %7.1f us" % (m, by / 1e6, t_plain, by / max(t_plain, 1e-9) / 1e3, t_fused), flush=True)
