"""The five BASELINE.json configurations as concrete chunk lists / edit scripts.

Shared by bench.py, the GPU parity tests and the scripts under profiles/, so that "config 3" is one
script everywhere.  Host-side only: numpy, no device.

All configs: V = 64, noise seed 93923590 (leven/default.cfg:8), defaultMaterial 0, LOD0
(size 256 = 64 voxels * LEAF_SIZE_SCALE, volume_constants.h:7-15), chunk (cx, cy, cz) -> min = 256 * (cx, cy, cz).
"""
import numpy as np

SEED = 93923590          # leven/default.cfg:8
V = 64
SIZE = 256               # LOD0 chunk: 64 voxels * LEAF_SIZE_SCALE
CY0 = 9                  # floor(h(0,0) / 64): h(0,0) = 598.49 voxels for this seed (bench.py checks it)
RING = 8                 # config 2: 8 x 8 x 8 chunks
SWEEP = 16               # config 5: 16 x 16 x 16 chunks
STRESS_THRESHOLD = 0.735  # config 4: tuned once to 30 % active voxels
CSG_OPS = 32             # config 3


def single_chunk():
    """configs[0]: the surface chunk above the world origin"""
    return np.array([[0, CY0 * SIZE, 0, SIZE]], np.int32)


def ring_chunks(x_shift=0):
    """configs[1]: clipmap LOD0 ring, cx, cz in [-4, 4), cy in [CY0 - 4, CY0 + 4): 512 chunks"""
    h = RING // 2
    return np.array([[(cx + x_shift) * SIZE, (CY0 + dy) * SIZE, cz * SIZE, SIZE]
                     for dy in range(-h, h) for cz in range(-h, h) for cx in range(-h, h)], np.int32)


def sweep_chunks():
    """configs[4]: the default world's whole XZ extent (viewer.cpp:64-69), cy in [0, 16): 4096 chunks,
    linear index i = cx' + 16 * (cz' + 16 * cy) -- the index the round-robin sharding cuts (i mod G)"""
    n, h = SWEEP, SWEEP // 2
    return np.array([[(cx - h) * SIZE, cy * SIZE, (cz - h) * SIZE, SIZE]
                     for cy in range(n) for cz in range(n) for cx in range(n)], np.int32)


def stress_chunks():
    """configs[3]: 4 x 4 x 4 chunks of the dense 3-D field (density kind 1, STRESS_THRESHOLD)"""
    return np.array([[cx * SIZE, cy * SIZE, cz * SIZE, SIZE] for cy in range(4) for cz in range(4) for cx in range(4)], np.int32)


def csg_script(num_ops=CSG_OPS, seed=12345):
    """configs[2]: the fixed edit script.  One op per step, alternating cube / sphere, two adds
    (materials 1-3) then two subtracts, centres from MT19937(seed) within +-2 chunks of the surface
    point above the origin, half-dimensions 1-31 voxels (the viewer's brush range 8-248 world units,
    viewer.h:51-52).  Returns tuples (type, shape, material, origin[3], dimensions[3], rotateY) --
    the argument order of CSGOperationInfo.make.  Another seed gives another script of the same kind (bench.py warms
    the context with one before it times the fixed script)."""
    rng = np.random.RandomState(seed)
    sy = CY0 * 64
    ops = []
    for step in range(num_ops):
        shape = step % 2
        add = (step // 2) % 2 == 0
        origin = [float(rng.randint(-128, 128)) + 0.5, sy + float(rng.randint(-40, 60)) + 0.5, float(rng.randint(-128, 128)) + 0.5]
        half = float(rng.randint(1, 32))
        dims = [half, half, half] if shape == 1 else [float(rng.randint(1, 32)) for _ in range(3)]
        material = int(rng.randint(1, 4)) if add else 201
        ops.append((0 if add else 1, shape, material, origin, dims, 0.0))
    return ops


def touched_chunks(chunks, lo, hi):
    """the chunks of a list whose AABB overlaps an operation's bounds (clipmap.cpp:1647-1744 re-meshes
    exactly these; AABB::overlaps, aabb.h:24-33)"""
    c = np.asarray(chunks, np.int32).reshape(-1, 4)
    s = c[:, 3]
    keep = ~((c[:, 0] + s < lo[0]) | (c[:, 1] + s < lo[1]) | (c[:, 2] + s < lo[2]) |
             (c[:, 0] > hi[0]) | (c[:, 1] > hi[1]) | (c[:, 2] > hi[2]))
    return np.ascontiguousarray(c[keep])
