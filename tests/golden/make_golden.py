"""Generates the committed golden fixtures from the CPU oracle.

The reference ships no golden vectors and cannot be built or imported in this image,
so these vectors pin the ORACLE's current behaviour (itself pinned by tests/test_oracle_kat.py):
the CPU suite checks the oracle still reproduces them, the GPU suite checks the CUDA path does.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import irt_b200.workloads as wl  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    orc = Oracle("canonical")
    out = {}
    for name, spec in (("a", wl.robot_a(0.003)), ("b", wl.robot_b(0.003)), ("brot", wl.robot_b(0.003, rotation=True))):
        rb = orc.robot(spec)
        st = wl.sample_states(spec, 12, stream=301)
        cap = len(orc.t_range(0.0, spec["L"], spec["dL"]))
        ref = orc.fk_batch(rb, st, cap)
        out["fk_%s_states" % name] = st
        for k in ("p", "npts", "L_i", "tip", "flags", "iters", "nsteps"):
            out["fk_%s_%s" % (name, k)] = ref[k]
        g = wl.workspace_grid(spec)
        grid = orc.grid(g["Ng"], g["lim"])
        store, flags = orc.voxelize_vertices_batch(rb, grid, st)
        off, keys, bits = store.export()
        out["vv_%s_off" % name], out["vv_%s_keys" % name], out["vv_%s_bits" % name] = off, keys, bits
        a = st[:6]
        b = st[:6] + 0.2 * (st[6:12] - st[:6])
        es, info = orc.voxelize_edges_batch(rb, grid, orc.space(), a, b)
        off, keys, bits = es.export()
        out["ve_%s_a" % name], out["ve_%s_b" % name] = a, b
        out["ve_%s_off" % name], out["ve_%s_keys" % name], out["ve_%s_bits" % name] = off, keys, bits
        for k in ("flags", "t_last", "nsamples"):
            out["ve_%s_%s" % (name, k)] = info[k]
        # environment: two spheres; verdicts of the vertex and edge sets
        env = orc.octree(grid)
        env.add_sphere([0.03, 0.0, 0.12], 0.025)
        env.add_sphere([-0.04, 0.03, 0.08], 0.02)
        bxyz, ebits = env.export()
        out["env_%s_bxyz" % name], out["env_%s_bits" % name] = bxyz, ebits
        out["vv_%s_verdict" % name] = orc.check_sets_batch(store, env)
        out["ve_%s_verdict" % name] = orc.check_sets_batch(es, env)
    np.savez_compressed(os.path.join(HERE, "golden_r1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_r1.npz"), "with", len(out), "arrays")


if __name__ == "__main__":
    main()
