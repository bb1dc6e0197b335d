"""time K3 over a large synthetic store (clustered keys like swept volumes)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import irt_b200
ctx = irt_b200.Context(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
rng = np.random.default_rng(8)
Ng, Nb = 128, 32
grid = irt_b200.make_grid(Ng, [-0.21, 0.21] * 3)
for (n_sets, lo, hi, envfrac) in ((4_000_000, 20, 60, 25), (10_000_000, 15, 40, 25), (10_000_000, 15, 40, 400)):
    sizes = rng.integers(lo, hi, size=n_sets)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    nb = int(off[-1])
    starts = rng.integers(0, Nb ** 3 - 64, size=n_sets)
    keys = (np.repeat(starts, sizes) + (np.arange(nb) - np.repeat(off[:-1].astype(np.int64), sizes))).astype(np.uint32)
    bits = rng.integers(1, 2 ** 63, size=nb, dtype=np.uint64)
    store = irt_b200.SetStore(ctx, grid); store.import_csr(off, keys, bits)
    env = irt_b200.Env(ctx, grid)
    e1 = np.zeros(Nb ** 3, dtype=np.uint64)
    occ = rng.choice(Nb ** 3, size=Nb ** 3 // envfrac, replace=False)
    e1[occ] = rng.integers(1, 2 ** 62, size=len(occ), dtype=np.uint64)
    env.update(e1)
    words = torch.zeros((n_sets + 31) // 32, dtype=torch.int32, device="cuda")
    for _ in range(3):
        store.check_dev(env, words, stream=s.cuda_stream)
    torch.cuda.synchronize()
    e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(10):
        store.check_dev(env, words, stream=s.cuda_stream)
    e1_.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1_) / 10
    ab = store.algorithmic_bytes()
    v = irt_b200.unpack_verdicts(words.cpu().numpy().view(np.uint32), n_sets)
    x = bits & e1[keys]
    want = np.add.reduceat((x != 0).astype(np.int64), off[:-1].astype(np.int64)) > 0
    print("sets %d leaves %d: %.3f ms, %.0f GB/s alg (%.1f%% of 6552), %.2f Gsets/s, collide %.3f, verdicts ok %s"
          % (n_sets, nb, ms, ab / ms / 1e6, 100 * ab / ms / 1e6 / 6552, n_sets / ms / 1e6, v.mean(), np.array_equal(v, want)))
    del store, env
