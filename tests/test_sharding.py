"""N > 1 host logic on CPU: world_size-2 gloo run of the round-robin sharding and the count
gather.  The per-chunk work is done by the oracle here (no GPU in this suite);
the GPU suite checks that the CUDA path gives the same per-chunk counts."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, SEED


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, chunks, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from leven_b200 import sharding
    from oracle import oracle as O
    mine = sharding.shard_round_robin(len(chunks), rank, world_size)
    world = O.World(seed=SEED)
    counts, _ = world.batch_counts(chunks[mine])
    local = np.stack([np.where(counts[:, 2] > 0, counts[:, 1], 0), counts[:, 2], counts[:, 3]], axis=1)
    c, off, tot = sharding.gather_global_offsets(local, mine, len(chunks))
    # the repeating form bench.py's sweep uses (pre-allocated buffers, one all_gather of 12 B per chunk):
    # an odd chunk count exercises the padded last row of the shorter rank
    g = sharding.CountGather(len(chunks), rank, world_size)
    for _ in range(2):
        c2, off2, tot2 = g.gather(local[:, 0], local[:, 1], local[:, 2])
    assert np.array_equal(c, c2) and np.array_equal(off, off2) and np.array_equal(tot, tot2)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), counts=c, offsets=off, totals=tot, mine=mine)
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_partition():
    from leven_b200 import sharding
    for world in (1, 2, 4, 8):
        parts = [sharding.shard_round_robin(4096, r, world) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(4096))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    sw = sharding.sweep_chunks()
    assert len(sw) == 4096 and sw[:, 0].min() == -8 * 256 and sw[:, 0].max() == 7 * 256 and sw[:, 1].max() == 15 * 256
    owners = [sharding.stable_owner(c[:3], 256, 8) for c in sw[:512]]
    assert set(owners) == set(range(8))
    assert owners == [sharding.stable_owner(c[:3], 256, 8) for c in sw[:512]]


def test_world_size_2_gloo_count_gather(tmp_path, built, surface_cy):
    from leven_b200 import sharding
    from oracle import oracle as O
    chunks = np.array([[cx * 256, (surface_cy + dy) * 256, 0, 256] for dy in (-1, 0) for cx in range(-2, 2)], np.int32)[:7]
    port = _free_port()
    mp.spawn(_worker, args=(2, port, chunks, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["counts"], r1["counts"]) and np.array_equal(r0["offsets"], r1["offsets"])
    assert set(r0["mine"]).isdisjoint(set(r1["mine"])) and len(r0["mine"]) + len(r1["mine"]) == len(chunks)
    # single-process answer
    world = O.World(seed=SEED)
    counts, _ = world.batch_counts(chunks)
    ref = np.stack([np.where(counts[:, 2] > 0, counts[:, 1], 0), counts[:, 2], counts[:, 3]], axis=1)
    assert np.array_equal(r0["counts"], ref)
    assert np.array_equal(r0["offsets"], np.cumsum(ref, axis=0) - ref)
    assert np.array_equal(r0["totals"], ref.sum(axis=0)) and ref.sum() > 0


def _exchange_worker(rank, world_size, port, nodes, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from leven_b200 import sharding
    from oracle import oracle as O
    mine = sharding.shard_round_robin(len(nodes), rank, world_size)
    world = O.World(seed=SEED)
    tris, counts, offsets, parts, used = [], [], [], [], 0
    for i in mine:          # this rank's pass 1 (the oracle stands in for the CUDA path in this suite)
        r = world.generate_chunk_mesh(list(nodes[i][:3]), int(nodes[i][3]))
        world.free_chunk_octree(list(nodes[i][:3]), int(nodes[i][3]))
        tris.append(r["numTriangles"]); counts.append(len(r["seams"])); offsets.append(used)
        parts.append(r["seams"]); used += len(r["seams"])
    local = np.concatenate(parts) if used else np.zeros(0, parts[0].dtype if parts else np.uint8)
    table, arena, per_rank = sharding.exchange_seam_nodes(len(nodes), mine, tris, counts, offsets, local)
    np.savez(os.path.join(out_dir, f"x{rank}.npz"), table=table, arena=np.asarray(arena), per_rank=per_rank, mine=mine)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_seam_node_exchange(tmp_path, built, surface_cy):
    """the one exchange step of the batched update over G GPUs (leven_b200/sharding.py): after it every
    rank holds every node's triangle count and seam nodes, addressed through the same table"""
    from oracle import oracle as O
    nodes = np.array([[cx * 256, (surface_cy + dy) * 256, cz * 256, 256] for dy in (-1, 0) for cx in (-1, 0) for cz in (-1, 0, 1)]
                     + [[0, 15 * 256, 0, 256]], np.int32)          # 12 surface-side nodes + one of air
    port = _free_port()
    mp.spawn(_exchange_worker, args=(2, port, nodes, str(tmp_path)), nprocs=2, join=True)
    x0, x1 = np.load(tmp_path / "x0.npz"), np.load(tmp_path / "x1.npz")
    assert np.array_equal(x0["table"], x1["table"]) and x0["arena"].tobytes() == x1["arena"].tobytes()
    assert x0["per_rank"] == x1["per_rank"] and len(x0["arena"]) == 2 * int(x0["per_rank"]) * 48
    world = O.World(seed=SEED)
    seen = 0
    for i, n in enumerate(nodes):
        r = world.generate_chunk_mesh(list(n[:3]), int(n[3]))
        world.free_chunk_octree(list(n[:3]), int(n[3]))
        tri, cnt, first = (int(v) for v in x0["table"][i])
        assert tri == r["numTriangles"] and cnt == len(r["seams"])
        got = x0["arena"][first * 48:(first + cnt) * 48]
        assert got.tobytes() == np.ascontiguousarray(r["seams"]).tobytes(), i
        seen += cnt
    assert seen > 1000 and x0["table"][-1].tolist() == [0, 0, 0]
