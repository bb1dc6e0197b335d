"""Latency / throughput of the batched finite-difference tip Jacobian.  (The CPU
figure quoted beside it in INTEGRATION.md -- 0.72 ms for the 15 sequential FKs of one Jacobian on one
host thread -- was taken once with the oracle port; tools/ does not load oracle/.)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import irt_b200, irt_b200.workloads as wl
ctx = irt_b200.Context(0)
spec = wl.robot_b(0.005)
rb = irt_b200.Robot(ctx, spec)
st = wl.sample_states(spec, 100000, stream=5)
for n in (1, 10, 100, 1000, 10000, 100000):
    s = st[:n]
    rb.tip_jacobian_batch(s, mode=2, delta=1e-6)
    reps = max(3, min(200, 20000 // n))
    t0 = time.perf_counter()
    for _ in range(reps):
        rb.tip_jacobian_batch(s, mode=2, delta=1e-6)
    dt = (time.perf_counter() - t0) / reps
    print("GPU n=%6d seeds: %.3f ms per batch (host buffers), %.0f Jacobians/s, %.2f M FK/s" % (n, dt * 1e3, n / dt, n * 15 / dt / 1e6), flush=True)
