"""Times the device environment-preparation steps on the C4 grid (128^3 cells = 32^3... leaf blocks)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import irt_b200 as irt
from irt_b200 import workloads as wl

spec = wl.robot_b(0.005)
g = wl.workspace_grid(spec)
ctx = irt.Context(0)
grid = irt.make_grid(g["Ng"], g["lim"])
blocks = wl.dense_to_morton_blocks(wl.lung_like_env_dense(spec, g))
env = irt.Env(ctx, grid)
for name, args in [("dilate", (1,)), ("dilate", (4,)), ("dilate", (4, True)), ("dilate_sphere", (spec["r"],)),
                   ("remove_interior", (False,)), ("remove_interior", (True,))]:
    ts = []
    for _ in range(5):
        env.update(blocks)
        t0 = time.perf_counter(); getattr(env, name)(*args); ts.append(time.perf_counter() - t0)
    print(f"{name}{args}: {min(ts)*1e3:.3f} ms (host wall, synchronous call)  nblocks={env.nblocks()}")
for Ng in (256, 512):
    grid = irt.make_grid(Ng, g["lim"])
    env = irt.Env(ctx, grid)
    rng = np.random.default_rng(1)
    b = np.zeros((Ng // 4) ** 3, np.uint64)
    idx = rng.integers(0, b.size, b.size // 20)
    b[idx] = rng.integers(1, 2 ** 63, len(idx), dtype=np.uint64)
    for name, args in [("dilate", (4,)), ("remove_interior", (True,))]:
        ts = []
        for _ in range(3):
            env.update(b)
            t0 = time.perf_counter(); getattr(env, name)(*args); ts.append(time.perf_counter() - t0)
        print(f"Ng={Ng} {name}{args}: {min(ts)*1e3:.3f} ms")
