"""first-contact GPU diagnostic (not a test): parity + rough timings"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import irt_b200, irt_b200.workloads as wl
from oracle.oracle import Oracle

ctx = irt_b200.Context(0)
print("fp64 peak TFLOP/s:", ctx.fp64_peak() / 1e12)
orc = Oracle("canonical")
for name, spec in [("A.005", wl.robot_a(0.005)), ("A.003", wl.robot_a(0.003)), ("B.005", wl.robot_b(0.005)), ("B.003", wl.robot_b(0.003)), ("Brot", wl.robot_b(0.005, rotation=True))]:
    rb = irt_b200.Robot(ctx, spec)
    st = wl.sample_states(spec, 2000, stream=1)
    out = rb.shape_batch(st, want=("p", "R", "t", "npts", "L", "L_i", "tip", "uv", "flags", "iters", "nsteps"))
    ref = orc.fk_batch(orc.robot(spec), st, rb.max_points)
    print(name, "npts eq", np.array_equal(out["npts"], ref["npts"]), "p err", np.abs(out["p"] - ref["p"]).max(),
          "Li err", np.abs(out["L_i"] - ref["L_i"]).max(), "flags eq", np.array_equal(out["flags"], ref["flags"]),
          "iters eq", np.array_equal(out["iters"], ref["iters"]), "nsteps eq", np.array_equal(out["nsteps"], ref["nsteps"]),
          "nflag", int((out["flags"] != 0).sum()))
import torch
for name, spec, n in [("A.005", wl.robot_a(0.005), 1 << 20), ("B.005", wl.robot_b(0.005), 1 << 20), ("B.003", wl.robot_b(0.003), 1 << 20)]:
    rb = irt_b200.Robot(ctx, spec)
    st = torch.from_numpy(wl.sample_states(spec, n, stream=2)).cuda()
    outs = dict(p=torch.empty(n, rb.max_points, 3, dtype=torch.float64, device="cuda"),
                npts=torch.empty(n, dtype=torch.int32, device="cuda"),
                L_i=torch.empty(n, rb.n_tendons, dtype=torch.float64, device="cuda"),
                nsteps=torch.empty(n, dtype=torch.int32, device="cuda"), iters=torch.empty(n, dtype=torch.int32, device="cuda"))
    for withflags in (False, True):
        o = dict(outs)
        if withflags:
            o["flags"] = torch.empty(n, dtype=torch.int32, device="cuda")
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.time()
            rb.shape_batch_dev(st, n, o); ctx.synchronize()
            dt = time.time() - t0
        N = rb.n_tendons
        steps = outs["nsteps"].double().mean().item(); iters = outs["iters"].double().mean().item()
        flop = steps * (4 * (346 + 162 * N) + 13 * (19 + N)) + iters * (30 + 46 * N)
        print(name, "flags" if withflags else "noflags", "n", n, "ms", dt * 1e3, "Mshapes/s", n / dt / 1e6, "mean steps", steps, "iters", iters,
              "TFLOP/s(alg)", n * flop / dt / 1e12)
