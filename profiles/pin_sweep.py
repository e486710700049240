#!/usr/bin/env python
"""The checker against the reference's own kernels over many chunks (CPU only; needs oracle/_ref,
i.e. the container that holds /root/reference): N surface chunks of the default terrain at LOD0/1/2,
every stage array of oracle/lvn_oracle.c compared bit for bit with what leven/cl/*.cl compiled for
the host produce through the reference's host sequence (oracle/ref.py).  The test suite does this
for six random chunks (tests/test_ref_pin.py::test_live_random_chunks); this is the long form.
    python profiles/pin_sweep.py [N]  ->  one JSON line"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O
from oracle import ref as R
import test_ref_pin as T

N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
O.build(); R.build()
world = O.World(seed=T.SEED)
rw = R.RefWorld(world.image)
rng = np.random.default_rng(777)
done, bad, seen, t0 = 0, [], set(), time.perf_counter()
per_lod = {256: 0, 512: 0, 1024: 0}
while done < N:
    size = int(rng.choice([256, 256, 512, 1024]))
    cx, cz = (int(v) for v in rng.integers(-8, 8, 2))
    h = -rw.density(np.float32(cx * 64.0 + 32), np.float32(0.0), np.float32(cz * 64.0 + 32))
    cy = int(h * 4 // size) + int(rng.integers(-1, 2)) * int(rng.integers(0, 2))
    mn = (cx * 256 // size * size, cy * size, cz * 256 // size * size)
    if (mn, size) in seen:
        continue
    seen.add((mn, size))
    r = rw.generate_chunk_mesh(list(mn), size)
    o = world.generate_chunk_mesh(list(mn), size)
    world.free_chunk_octree(list(mn), size)
    ok = all(o[k] == r[k] for k in T.COUNTS)
    if ok and r["numNodes"] == 0:
        continue
    if ok:
        for k in T.STAGES:
            a, b = (T.zero_pad(o[k]), T.zero_pad(r[k])) if k == "qefs" else (o[k], r[k])
            if not T.beq(a, b):
                ok = False
                bad.append({"min": mn, "size": size, "stage": k})
                break
    else:
        bad.append({"min": mn, "size": size, "stage": "counts"})
    done += 1
    per_lod[size] += 1
print(json.dumps({"surface_chunks_compared": done, "per_size": per_lod, "mismatches": len(bad), "first": bad[:3],
                  "stages": list(T.STAGES), "seconds": time.perf_counter() - t0,
                  "what": "oracle/lvn_oracle.c vs the reference's leven/cl kernels compiled for the host (oracle/_ref), bit for bit"}))
