"""Multi-GPU sharding of the chunk-meshing path.

Chunks are independent: no kernel of one chunk reads another chunk's data, seams are stitched by
the caller from per-chunk SeamNodeInfo (clipmap.cpp:573-611).  So the path shards with NO
collective on the data path; the only exchange is the per-chunk (numVertices, numTriangles,
numSeamNodes) count gather needed when one global mesh is assembled.
"""
import numpy as np


def shard_round_robin(num_chunks, rank, world_size):
    """BASELINE config 5: chunk linear index i goes to GPU i mod G."""
    return np.arange(rank, num_chunks, world_size, dtype=np.int64)


def stable_owner(chunk_min, size, world_size):
    """Owner of a chunk's cached state (CSG-edited field, octree): a hash of the cache key
    ivec4(min, size) (compute_local.h:40,53), so edits and re-meshes hit the GPU holding the field."""
    h = 1469598103934665603
    for v in (int(chunk_min[0]), int(chunk_min[1]), int(chunk_min[2]), int(size)):
        h ^= v & 0xffffffff
        h = (h * 1099511628211) & 0xffffffffffffffff
    h ^= h >> 33                                   # avalanche: FNV's low bits alone are weak
    h = (h * 0xff51afd7ed558ccd) & 0xffffffffffffffff
    h ^= h >> 33
    return int(h % world_size)


def sweep_chunks(nx=16, ny=16, nz=16, size=256):
    """BASELINE config 5: cx, cz in [-8, 8), cy in [0, 16); linear index cx' + 16 (cz' + 16 cy)."""
    out = np.zeros((nx * ny * nz, 4), np.int32)
    i = 0
    for cy in range(ny):
        for cz in range(nz):
            for cx in range(nx):
                out[i] = ((cx - nx // 2) * size, cy * size, (cz - nz // 2) * size, size)
                i += 1
    return out


def gather_global_offsets(local_counts, local_indices, num_chunks, group=None):
    """All-gather the per-chunk counts of every rank and return, for ALL chunks in linear index
    order, (counts[num_chunks, 3], offsets[num_chunks, 3], totals[3]): the base offsets of each
    chunk's vertices / triangles / seam nodes in one global mesh.  12 bytes per chunk cross the
    wire; works with the gloo backend on CPU tensors and with NCCL on CUDA tensors."""
    import torch
    import torch.distributed as dist

    counts = torch.zeros((num_chunks, 3), dtype=torch.int64)
    counts[torch.as_tensor(np.asarray(local_indices), dtype=torch.long)] = torch.as_tensor(
        np.asarray(local_counts, dtype=np.int64).reshape(-1, 3))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        backend = dist.get_backend(group)
        buf = counts.cuda() if backend == "nccl" else counts
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)   # shards are disjoint: sum == gather
        counts = buf.cpu()
    offsets = torch.cumsum(counts, dim=0) - counts
    return counts.numpy(), offsets.numpy(), counts.sum(dim=0).numpy()
