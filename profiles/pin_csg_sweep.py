#!/usr/bin/env python
"""The checker's CSG path against the reference's own kernels over many random scripts (CPU only;
needs oracle/_ref): the long form of tests/test_ref_pin.py::test_live_csg_random_scripts.  Each trial
takes one of several surface chunks, applies 1-6 random operations (sphere / cube, add with
materials 1-4 / subtract, half-dimensions 1-31 voxels, rotations on two trials of three) through
apply_csg_operation.cl compiled for the host and through oracle/lvn_oracle.c, and compares the
material field, the edge set with (normal, t), and the octree / mesh / seam nodes built from them.
    python profiles/pin_csg_sweep.py [TRIALS]  ->  one JSON line"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O
from oracle import ref as R
import test_ref_pin as T

TRIALS = int(sys.argv[1]) if len(sys.argv) > 1 else 40
O.build(); R.build()
image = O.noise_image(T.SEED)
rw = R.RefWorld(image)
probe = O.World(image=image, default_material=0, voxels_per_chunk=64)
cy = int(900 * probe.terrain(0.0, 0.0) // 64)
probe.close()
rng = np.random.default_rng(4242)
bad, t0 = [], time.perf_counter()
for trial in range(TRIALS):
    mn = [int(rng.integers(-2, 3)) * 256, (cy + int(rng.integers(-1, 1))) * 256, int(rng.integers(-2, 3)) * 256]
    w = O.World(image=image, default_material=0, voxels_per_chunk=64)
    try:
        base = w.generate_chunk_mesh(mn, 256)
        w.free_chunk_octree(mn, 256)
        n = int(rng.integers(1, 7))
        ops = np.zeros(n, R.CSG_DTYPE)
        for i in range(n):
            add = bool(rng.integers(0, 2))
            org = [mn[0] / 4 + float(rng.integers(-2, 67)) + 0.5, mn[1] / 4 + float(rng.integers(-2, 67)) + 0.5, mn[2] / 4 + float(rng.integers(-2, 67)) + 0.5]
            dim = [float(rng.integers(1, 32)) for _ in range(3)]
            ops[i] = (0 if add else 1, int(rng.integers(0, 2)), int(rng.integers(1, 5)) if add else 201,
                      float(rng.uniform(-3.0, 3.0)) if trial % 3 else 0.0, org + [0.0], dim + [0.0])
        m, keys, info = rw.apply_csg(mn, 256, ops, base["materials"], base["edgeKeys"], base["edgeInfo"])
        w.apply_csg_operations(T.oracle_csg_ops(O, ops), mn, 256)
        o = w.generate_chunk_mesh(mn, 256)
        what = None
        idx = keys >> 2
        ingrid = ((idx & 127) < 65) & (((idx >> 7) & 127) < 65) & (((idx >> 14) & 127) < 65)     # DESIGN.md deviation 4
        ro, oo = np.argsort(keys[ingrid], kind="stable"), np.argsort(o["edgeKeys"], kind="stable")
        if not np.array_equal(o["materials"], m): what = "materials"
        elif not np.array_equal(keys[ingrid][ro], o["edgeKeys"][oo]): what = "edge set"
        elif not T.beq(info[ingrid][ro], o["edgeInfo"][oo]): what = "edge info"
        else:
            oc = rw.construct_octree(mn, 256, m, keys, info)
            if oc is None:
                if o["numNodes"] != 0: what = "empty octree"
            else:
                verts, tris = rw.generate_mesh(256, oc)
                if not (T.beq(verts, o["vertices"]) and T.beq(tris, o["indices"]) and T.beq(rw.gather_seam_nodes(oc), o["seams"])): what = "mesh / seam nodes"
        if what: bad.append({"trial": trial, "min": mn, "ops": n, "what": what})
    finally:
        w.close()
print(json.dumps({"trials": TRIALS, "mismatches": len(bad), "first": bad[:3], "seconds": time.perf_counter() - t0,
                  "what": "oracle/lvn_oracle.c CSG path vs the reference's apply_csg_operation.cl + octree.cl compiled for the host, bit for bit"}))
