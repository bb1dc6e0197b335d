import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import irt_b200, irt_b200.workloads as wl
from oracle.oracle import Oracle
ctx = irt_b200.Context(0); orc = Oracle()
spec = wl.robot_b(0.005, rotation=True)
rb = irt_b200.Robot(ctx, spec)
st = wl.sample_states(spec, 2000, stream=1)
ref = orc.fk_batch(orc.robot(spec), st, rb.max_points)
for want in [("p","npts"), ("p","npts","R"), ("p","npts","t"), ("p","npts","uv"), ("p","npts","L","L_i"), ("p","npts","flags"), ("p","npts","iters","nsteps"),
             ("p", "R", "t", "npts", "L", "L_i", "tip", "uv", "flags", "iters", "nsteps")]:
    out = rb.shape_batch(st, want=want)
    err = np.abs(out["p"]-ref["p"]).max(axis=(1,2))
    bad = np.nonzero(err > 1e-12)[0]
    print(want, len(bad), bad[:10], err.max())
    if len(bad):
        i = bad[0]; n = out["npts"][i]
        d = np.abs(out["p"][i]-ref["p"][i]).max(axis=1)
        print("   row", i, "npts", n, ref["npts"][i], "first bad point", np.nonzero(d > 1e-12)[0][:5], st[i])
