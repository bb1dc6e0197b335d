import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import irt_b200, irt_b200.workloads as wl
ctx = irt_b200.Context(0)
spec = wl.robot_b(0.005)
rb = irt_b200.Robot(ctx, spec)
n = 1_000_000; cap = rb.max_points
states = wl.sample_states(spec, n, stream=100)
h_states = torch.from_numpy(states).pin_memory()
h_out = dict(p=torch.zeros(n, cap, 3, dtype=torch.float64).pin_memory(), npts=torch.zeros(n, dtype=torch.int32).pin_memory(),
             L=torch.zeros(n, dtype=torch.float64).pin_memory(), L_i=torch.zeros(n, 6, dtype=torch.float64).pin_memory())
o = irt_b200.FkOutputs()
for k, t in h_out.items(): setattr(o, k, t.data_ptr())
ts = []
for i in range(12):
    t0 = time.perf_counter()
    ctx.check(ctx.L.irt_fk_batch(ctx.h, rb.h, C.c_void_p(h_states.data_ptr()), rb.state_size, n, cap, C.byref(o)))
    ts.append((time.perf_counter() - t0) * 1e3)
print("e2e ms:", " ".join("%.1f" % t for t in ts))
# raw pinned D2H bandwidth with torch for reference
d = torch.zeros(n, cap, 3, dtype=torch.float64, device="cuda")
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); h_out["p"].copy_(d, non_blocking=True); torch.cuda.synchronize()
    print("torch D2H 984MB: %.1f ms -> %.1f GB/s" % ((time.perf_counter() - t0) * 1e3, 0.984 / (time.perf_counter() - t0)))
