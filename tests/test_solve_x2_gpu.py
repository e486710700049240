"""k_solve's arithmetic on adversarial QEFs (lvn_debug_solve_qefs): the production form -- two nodes per
thread, packed FP32, divisions / reciprocals / square roots written out as FMA sequences with their operand
ranges established once per Jacobi rotation -- must equal the oracle's qef_solve (qef.cl:239-256 restated)
bit for bit, including the regimes the range analysis separates: off-diagonals so small that tau exceeds
2^60 / 2^64 or overflows, denormal off-diagonals, exactly-zero off-diagonals in one node of a pair only,
huge and tiny scales, singular matrices (svd_invdet's cut-offs), and ordinary terrain-like QEFs."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_qefs(rng, n):
    q = np.zeros((n, 16), np.float32)
    kind = rng.randint(0, 10, size=n)
    for i in range(n):
        k = kind[i]
        m = rng.randint(1, 7)
        nrm = rng.normal(size=(m, 3)).astype(np.float32)
        nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-6).astype(np.float32)
        pts = rng.uniform(0, 64, size=(m, 3)).astype(np.float32)
        ata = (nrm.T @ nrm).astype(np.float32)
        atb = (nrm.T @ np.sum(nrm * pts, axis=1)).astype(np.float32)
        if k == 1:      # nearly axis-aligned normals: tiny off-diagonals
            ata[0, 1] *= np.float32(10.0 ** -rng.randint(5, 30)); ata[0, 2] *= np.float32(10.0 ** -rng.randint(5, 38)); ata[1, 2] *= np.float32(1e-12)
        elif k == 2:    # denormal / vanishing off-diagonals
            ata[0, 1] = np.float32(rng.choice([1e-39, -3e-42, 1e-45, 0.0])); ata[1, 2] = np.float32(rng.choice([2e-40, 0.0, -1e-44]))
        elif k == 3:    # one zero off-diagonal
            ata[0, 2] = 0
        elif k == 4:    # diagonal already
            ata[0, 1] = ata[0, 2] = ata[1, 2] = 0
        elif k == 5:    # large scale
            s = np.float32(2.0 ** rng.randint(10, 40)); ata *= s; atb *= s
        elif k == 6:    # small scale
            s = np.float32(2.0 ** -rng.randint(10, 45)); ata *= s; atb *= s
        elif k == 7:    # equal diagonal entries: tau = 0
            ata[1, 1] = ata[0, 0]
        elif k == 8:    # rank one
            ata = np.outer(nrm[0], nrm[0]).astype(np.float32); atb = (nrm[0] * np.dot(nrm[0], pts[0])).astype(np.float32)
        q[i, 0:6] = [ata[0, 0], ata[0, 1], ata[0, 2], ata[1, 1], ata[1, 2], ata[2, 2]]
        q[i, 8:11] = atb
        q[i, 12:15] = pts.mean(axis=0)
        q[i, 15] = 1.0
    return q


def test_packed_solve_equals_oracle(lc, oracle_mod):
    rng = np.random.RandomState(20261017)
    q = make_qefs(rng, 40000)
    # chunk QEFs as the path produces them, too
    world = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=64)
    cy = int(900 * world.terrain(0.0, 0.0) // 64)
    ref = world.generate_chunk_mesh([0, cy * 256, 0], 256)
    world.close()
    rq = ref["qefs"]
    q2 = np.concatenate([rq["ATA"], rq["pad"], rq["ATb"], rq["masspoint"]], axis=1).astype(np.float32)
    q = np.concatenate([q, q2])
    want = np.zeros((len(q), 4), np.float32)
    mn = (C.c_int * 3)(0, 0, 0)
    oracle_mod.lib().lvo_solve_qefs(mn, q.ctypes.data_as(C.c_void_p), len(q), want.ctypes.data_as(C.c_void_p))
    for packed in (False, True):
        got = lc.DebugSolveQEFs(q, packed=packed)
        same = (got.view(np.uint32) == want.view(np.uint32)).all(axis=1) | (np.isnan(got).any(axis=1) & np.isnan(want).any(axis=1))
        bad = np.nonzero(~same)[0]
        assert len(bad) == 0, f"packed={packed}: {len(bad)} of {len(q)} differ, first {bad[:5]}: {got[bad[:2]]} vs {want[bad[:2]]}"
    # an odd count exercises the last thread's single node
    got = lc.DebugSolveQEFs(q[:1001], packed=True)
    assert np.array_equal(got.view(np.uint32)[~np.isnan(want[:1001]).any(axis=1)], want[:1001].view(np.uint32)[~np.isnan(want[:1001]).any(axis=1)])
