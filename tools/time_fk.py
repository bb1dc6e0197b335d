"""time the fk kernel (device-resident) for a few robots; used to compare library variants"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import irt_b200, irt_b200.workloads as wl
ctx = irt_b200.Context(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
res = []
for name, spec in (("B.005", wl.robot_b(0.005)), ("A.005", wl.robot_a(0.005)), ("B.003", wl.robot_b(0.003))):
    rb = irt_b200.Robot(ctx, spec)
    n = 1_000_000
    st = torch.from_numpy(wl.sample_states(spec, n, stream=100)).cuda()
    outs = dict(p=torch.zeros(n, rb.max_points, 3, dtype=torch.float64, device="cuda"),
                npts=torch.zeros(n, dtype=torch.int32, device="cuda"),
                L_i=torch.zeros(n, rb.n_tendons, dtype=torch.float64, device="cuda"),
                nsteps=torch.zeros(n, dtype=torch.int32, device="cuda"), iters=torch.zeros(n, dtype=torch.int32, device="cuda"))
    for _ in range(3):
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
    res.append("%s %.2f ms %.1f TF" % (name, ms, n * flop / ms / 1e9))
    chk = float(outs["p"].sum().item())
print(os.environ.get("IRT_B200_LIB", "default").split("/")[-1], " | ".join(res), "chk %.9f" % chk)
