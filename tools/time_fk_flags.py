"""FK with and without the validity epilogue (flags) on 1M device-resident shapes, robots B dL=0.005 / dL=0.003:
what the epilogue costs on top of K1 (written for the K1 turning-hint experiment)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import irt_b200, irt_b200.workloads as wl
ctx = irt_b200.Context(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
res = []
for name, spec in (("B.005", wl.robot_b(0.005)), ("B.003", wl.robot_b(0.003))):
    rb = irt_b200.Robot(ctx, spec)
    n = 1_000_000
    st = torch.from_numpy(wl.sample_states(spec, n, stream=100)).cuda()
    base = dict(p=torch.zeros(n, rb.max_points, 3, dtype=torch.float64, device="cuda"),
                npts=torch.zeros(n, dtype=torch.int32, device="cuda"),
                L_i=torch.zeros(n, rb.n_tendons, dtype=torch.float64, device="cuda"))
    for flags in (False, True):
        outs = dict(base)
        if flags:
            outs["flags"] = torch.zeros(n, dtype=torch.int32, device="cuda")
        for _ in range(3):
            rb.shape_batch_dev(st, n, outs, stream=s.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            rb.shape_batch_dev(st, n, outs, stream=s.cuda_stream)
        e1.record(s); torch.cuda.synchronize()
        res.append("%s %s %.3f ms" % (name, "with flags" if flags else "no flags  ", e0.elapsed_time(e1) / 5))
        if flags:
            f = outs["flags"].cpu().numpy().view(np.uint32)
            res.append("flag bits seen 0x%x, self-colliding %.4f" % (int(np.bitwise_or.reduce(f)), float(((f & 4) != 0).mean())))
print(" | ".join(res))
