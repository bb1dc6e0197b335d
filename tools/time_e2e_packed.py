"""e2e FK through host pointers: dense irt_fk_batch vs packed irt_fk_batch_packed (pinned buffers)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import irt_b200, irt_b200.workloads as wl
ctx = irt_b200.Context(0)
spec = wl.robot_b(0.005)
rb = irt_b200.Robot(ctx, spec)
n = 1_000_000
st = torch.from_numpy(wl.sample_states(spec, n, stream=100)).pin_memory()
cap = rb.max_points
pin = lambda *shape, dt=torch.float64: torch.empty(*shape, dtype=dt).pin_memory()
outs = dict(p=pin(n * cap, 3), npts=pin(n, dt=torch.int32), L=pin(n), L_i=pin(n, 6), row_offsets=pin(n + 1, dt=torch.int64))
dense = irt_b200.FkOutputs()
import ctypes as C
for k in ("p", "npts", "L", "L_i"):
    setattr(dense, k, outs[k].data_ptr())
def run_dense():
    ctx.check(ctx.L.irt_fk_batch(ctx.h, rb.h, C.c_void_p(st.data_ptr()), 7, n, cap, C.byref(dense)))
def run_packed():
    ctx.check(ctx.L.irt_fk_batch_packed(ctx.h, rb.h, C.c_void_p(st.data_ptr()), 7, n, C.byref(dense), n * cap, C.c_void_p(outs["row_offsets"].data_ptr())))
for name, f in (("dense", run_dense), ("packed", run_packed), ("dense", run_dense), ("packed", run_packed)):
    for _ in range(2): f()
    t0 = time.perf_counter()
    for _ in range(5): f()
    dt = (time.perf_counter() - t0) / 5
    print("%s: %.2f ms per 1M shapes, %.1f M shapes/s (rows %d)" % (name, dt * 1e3, n / dt / 1e6, int(outs["row_offsets"][-1])))
