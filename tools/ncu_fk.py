"""small single-purpose drivers for ncu captures (one kernel each)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import irt_b200, irt_b200.workloads as wl

which = sys.argv[1] if len(sys.argv) > 1 else "fk"
ctx = irt_b200.Context(0)
if which == "fk":
    spec = wl.robot_b(0.005)
    rb = irt_b200.Robot(ctx, spec)
    n = 1_000_000
    st = torch.from_numpy(wl.sample_states(spec, n, stream=100)).cuda()
    outs = dict(p=torch.zeros(n, rb.max_points, 3, dtype=torch.float64, device="cuda"),
                npts=torch.zeros(n, dtype=torch.int32, device="cuda"),
                L=torch.zeros(n, dtype=torch.float64, device="cuda"),
                L_i=torch.zeros(n, 6, dtype=torch.float64, device="cuda"))
    for _ in range(3):
        rb.shape_batch_dev(st, n, outs)
        ctx.synchronize()
elif which == "fkv":      # the headline step: FK + validity epilogue (flags)
    spec = wl.robot_b(0.005)
    rb = irt_b200.Robot(ctx, spec)
    n = 1_000_000
    st = torch.from_numpy(wl.sample_states(spec, n, stream=100)).cuda()
    outs = dict(p=torch.zeros(n, rb.max_points, 3, dtype=torch.float64, device="cuda"),
                npts=torch.zeros(n, dtype=torch.int32, device="cuda"),
                L=torch.zeros(n, dtype=torch.float64, device="cuda"),
                L_i=torch.zeros(n, 6, dtype=torch.float64, device="cuda"),
                flags=torch.zeros(n, dtype=torch.int32, device="cuda"))
    for _ in range(3):
        rb.shape_batch_dev(st, n, outs)
        ctx.synchronize()
elif which == "peak":     # the FP64 roofline denominator: the DFMA-chain probe
    print("fp64 peak", ctx.fp64_peak() / 1e12, "TFLOP/s")
elif which == "k2":       # K2 over a small k-NN roadmap (numpy k-NN: no torch kernels in the launch list)
    spec = wl.robot_b(0.003)
    rb = irt_b200.Robot(ctx, spec)
    g = wl.workspace_grid(spec)
    grid = irt_b200.make_grid(g["Ng"], g["lim"])
    nv = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
    st = wl.sample_states(spec, nv, stream=200)
    pairs = wl.knn_edges(spec, st, k=10)
    store = irt_b200.SetStore(ctx, grid)
    import time
    for rep in range(2):
        t0 = time.perf_counter()
        info = store.voxelize_edges_indexed(rb, irt_b200.make_space(), st, pairs)
        print("k2 edges", len(pairs), "%.1f ms" % ((time.perf_counter() - t0) * 1e3), "samples/edge",
              info["nsamples"].mean(), "blocks/edge", store.num_blocks / len(pairs))
elif which == "k3real":
    # K3 over a roadmap store built by the real pipeline (bench.py's edge_check at 300k vertices)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import knn_edges_gpu
    spec = wl.robot_b(0.003)
    rb = irt_b200.Robot(ctx, spec)
    g = wl.workspace_grid(spec)
    grid = irt_b200.make_grid(g["Ng"], g["lim"])
    st = wl.sample_states(spec, 300000, stream=200)
    pairs = knn_edges_gpu(torch, st, spec, 10, torch.device("cuda"))
    store = irt_b200.SetStore(ctx, grid)
    store.voxelize_edges_indexed(rb, irt_b200.make_space(), st, pairs)
    env = irt_b200.Env(ctx, grid)
    env.update(wl.dense_to_morton_blocks(wl.lung_like_env_dense(spec, g)))
    words = torch.zeros((len(pairs) + 31) // 32, dtype=torch.int32, device="cuda")
    for _ in range(3):
        store.check_dev(env, words); ctx.synchronize()
    print("k3real sets", store.num_sets, "leaves", store.num_blocks, "alg bytes", store.algorithmic_bytes(),
          "collide frac", irt_b200.unpack_verdicts(words.cpu().numpy().view(np.uint32), len(pairs)).mean())
else:
    rng = np.random.default_rng(8)
    Ng, Nb = 128, 32
    grid = irt_b200.make_grid(Ng, [-0.21, 0.21] * 3)
    n_sets = 4_000_000
    sizes = rng.integers(20, 60, size=n_sets)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    nb = int(off[-1])
    # clustered keys like a real swept volume: consecutive morton keys around a random start
    starts = rng.integers(0, Nb ** 3 - 64, size=n_sets)
    keys = (np.repeat(starts, sizes) + (np.arange(nb) - np.repeat(off[:-1].astype(np.int64), sizes))).astype(np.uint32)
    bits = rng.integers(1, 2 ** 63, size=nb, dtype=np.uint64)
    store = irt_b200.SetStore(ctx, grid); store.import_csr(off, keys, bits)
    env = irt_b200.Env(ctx, grid)
    e1 = np.zeros(Nb ** 3, dtype=np.uint64)
    occ = rng.choice(Nb ** 3, size=Nb ** 3 // 25, replace=False)
    e1[occ] = rng.integers(1, 2 ** 62, size=len(occ), dtype=np.uint64)
    env.update(e1)
    words = torch.zeros((n_sets + 31) // 32, dtype=torch.int32, device="cuda")
    for _ in range(3):
        store.check_dev(env, words); ctx.synchronize()
    print("alg bytes", store.algorithmic_bytes(), "collide frac", irt_b200.unpack_verdicts(words.cpu().numpy().view(np.uint32), n_sets).mean())
