#!/usr/bin/env python
"""The batched Clipmap::update over G GPUs (leven_b200/sharding.py: sharded_clipmap_update), one
process per GPU under torchrun, NCCL:  pass 1 on this rank's nodes, the all-gather of the seam
nodes over NVLink (the one exchange step of the widened path), pass 2 on this rank's share of the
seam-update set.  Strong scaling: the workload is fixed (a 16 x 3 x 16 LOD0 block, 768 nodes), timed
with barriers on both sides, max over ranks; rank 0 checks that the ranks' seam meshes together
are the seam meshes of the one-GPU call (digests).
    torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 profiles/bench_update_multi.py"""
import hashlib, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import leven_b200.compute as lc
from leven_b200 import sharding
import bench as B

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lc.lib().lvn_compute_set_device(local)
assert lc.Compute_Initialise(B.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(64)
cy = B.CY0
ms = np.array([[cx * 256, (cy + dy) * 256, cz * 256, 256] for cx in range(-8, 8) for dy in (-1, 0, 1) for cz in range(-8, 8)], np.int32)
keep = []
pin = lambda n, dt: (keep.append(lc.PinnedArray(n, dt)), keep[-1].array)[1]
V, T, Sn = pin(3000000, lc.MeshVertex), pin(6000000, lc.MeshTriangle), pin(1000000, lc.SeamNodeInfo)


def digest(v, t):
    return hashlib.sha256(v.tobytes() + np.sort(np.ascontiguousarray(t["indices_"]).view([("", np.int32)] * 3), axis=0).tobytes()).hexdigest()


times = []
for it in range(8):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = sharding.sharded_clipmap_update(lc, ctx, ms, V, T, Sn)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times.append(time.perf_counter() - t0)
dt = torch.tensor([float(np.median(times[2:]))], device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
mine = {int(k): digest(V[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]], T[r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]])
        for k, r in zip(out["seam_update_nodes"], out["seam_results"])}
gathered = [None] * world
if world > 1:
    dist.all_gather_object(gathered, mine)
else:
    gathered = [mine]
if rank == 0:
    allseams = {}
    for g in gathered:
        allseams.update(g)
    # the one-GPU answer, on this rank
    nodes = np.zeros(len(ms), lc.ClipmapNode)
    nodes["min"] = ms[:, :3]; nodes["size"] = ms[:, 3]
    rc, cres, upd, sres, tot = lc.ClipmapUpdateBatch(ctx, nodes, 0, Sn, 0, V, T)
    assert rc == 0
    one = {int(k): digest(V[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]], T[r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]])
           for k, r in zip(upd, sres)}
    print(json.dumps({"n_gpus": world, "nodes": len(ms), "seam_updates": int(out["num_seam_updates_all"]), "update_ms": float(dt.item()) * 1e3,
                      "nodes_per_s": len(ms) / float(dt.item()), "scaling": "strong", "seams_match_one_gpu": allseams == one,
                      "exchange": "all_reduce of 4 int64 per node + all_gather_into_tensor of the ranks' SeamNodeInfo records (NCCL, device memory)",
                      "arena_records_per_rank": int(len(out["arena"]) // 48 // world) if hasattr(out["arena"], "__len__") else None}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
