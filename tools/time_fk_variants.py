"""Time K1 for several builds of the library in ONE process (launch-shape / register-cap experiments).
usage: python tools/time_fk_variants.py build/variants/libirt_*.so   -- the first one is the parity reference."""
import sys, os, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import irt_b200, irt_b200.workloads as wl

paths = sorted(p for a in sys.argv[1:] for p in glob.glob(a))
n = 1_000_000
robots = (("B.005", wl.robot_b(0.005)), ("B.003", wl.robot_b(0.003)), ("A.005", wl.robot_a(0.005)))
states = {name: torch.from_numpy(wl.sample_states(spec, n, stream=100)).cuda() for name, spec in robots}
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
ref = {}
for path in paths:
    irt_b200._lib = None
    irt_b200.LIB_PATH = os.path.abspath(path)
    ctx = irt_b200.Context(0)
    res = []
    for name, spec in robots:
        rb = irt_b200.Robot(ctx, spec)
        st = states[name]
        outs = dict(p=torch.zeros(n, rb.max_points, 3, dtype=torch.float64, device="cuda"),
                    npts=torch.zeros(n, dtype=torch.int32, device="cuda"),
                    L_i=torch.zeros(n, rb.n_tendons, dtype=torch.float64, device="cuda"),
                    flags=torch.zeros(n, dtype=torch.int32, device="cuda"),
                    nsteps=torch.zeros(n, dtype=torch.int32, device="cuda"),
                    iters=torch.zeros(n, dtype=torch.int32, device="cuda"))
        for _ in range(2):
            rb.shape_batch_dev(st, n, outs, stream=s.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            rb.shape_batch_dev(st, n, outs, stream=s.cuda_stream)
        e1.record(s); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        N = rb.n_tendons
        steps = outs["nsteps"].double().mean().item(); iters = outs["iters"].double().mean().item()
        flop = steps * (4 * (346 + 162 * N) + 13 * (19 + N)) + iters * (30 + 46 * N)
        if name not in ref:
            ref[name] = {k: v.clone() for k, v in outs.items()}
            d = "ref"
        else:
            r = ref[name]
            same = all(torch.equal(outs[k], r[k]) for k in ("npts", "flags", "nsteps", "iters"))
            d = "dp %.1e %s" % ((outs["p"] - r["p"]).abs().max().item() / spec["L"], "ints==" if same else "INTS DIFFER")
        res.append("%s %.3f ms %.2f TF (%s)" % (name, ms, n * flop / ms / 1e9, d))
        del outs, rb
    print(os.path.basename(path), " | ".join(res), flush=True)
    del ctx
