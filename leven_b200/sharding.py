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


class CountGather:
    """gather_global_offsets for a step that repeats: the buffers (pinned host staging + one device
    tensor per direction with NCCL; plain CPU tensors with gloo / one rank) are allocated once and a
    gather is one all_gather of exactly 12 B per chunk.  Round-robin sharding (chunk i on rank
    i mod G) makes the gathered array [G][ceil(n / G)][3] the transpose of the linear order.

    gather(numVertices, numTriangles, numSeamNodes) -> (counts[n, 3], offsets[n, 3], totals[3]),
    identical on every rank."""

    def __init__(self, num_chunks, rank, world_size, device=None, group=None):
        import torch
        self.torch, self.n, self.rank, self.world, self.group = torch, int(num_chunks), int(rank), int(world_size), group
        self.per = (self.n + self.world - 1) // self.world
        self.mine = len(shard_round_robin(self.n, self.rank, self.world))
        self.device = device
        on_gpu = device is not None
        # structure of arrays, [3][per]: every step below then runs over contiguous rows
        self.h_in = torch.zeros((3, self.per), dtype=torch.int32, pin_memory=on_gpu)
        self.h_out = torch.zeros((self.world, 3, self.per), dtype=torch.int32, pin_memory=on_gpu)
        self.a_in, self.a_out = self.h_in.numpy(), self.h_out.numpy()
        self.counts = np.zeros((self.n, 3), np.int64)
        self.offsets = np.zeros((self.n, 3), np.int64)
        self.totals = np.zeros(3, np.int64)
        if on_gpu:
            self.d_in = torch.zeros((3, self.per), dtype=torch.int32, device=device)
            self.d_out = torch.zeros((self.world, 3, self.per), dtype=torch.int32, device=device)
            # the gather has no input on the device: on a stream of its own it does not queue behind the kernels of
            # the batch whose counts it carries (lvn_meshgen_generate_batch_device_async returns while they run)
            self.stream = torch.cuda.Stream(device=device)

    def gather(self, num_vertices, num_triangles, num_seam_nodes):
        torch = self.torch
        import torch.distributed as dist
        a = self.a_in
        a[0, :self.mine] = num_vertices; a[1, :self.mine] = num_triangles; a[2, :self.mine] = num_seam_nodes
        multi = self.world > 1 and dist.is_available() and dist.is_initialized()
        if not multi:
            self.a_out[0] = a
        elif self.device is not None:
            with torch.cuda.stream(self.stream):
                self.d_in.copy_(self.h_in, non_blocking=True)
                dist.all_gather_into_tensor(self.d_out.view(-1), self.d_in.view(-1), group=self.group)
                self.h_out.copy_(self.d_out, non_blocking=True)
            self.stream.synchronize()
        else:
            dist.all_gather_into_tensor(self.h_out.view(-1), self.h_in.view(-1), group=self.group)
        # chunk i = j * world + r sits at [r][k][j]; the transposition and the exclusive scan are one host loop
        # (lvn_global_mesh_offsets: numpy's cumsum alone costs 40 us for these 12 k integers)
        from . import compute as lc
        rc = lc.GlobalMeshOffsets(self.a_out, self.world, self.per, self.n, self.counts, self.offsets, self.totals)
        assert rc == 0
        return self.counts, self.offsets, self.totals


# ---------------------------------------------------------------------------------------------
# The batched Clipmap::update over G GPUs.  Pass 1 -- constructing nodes -- shards
# like the chunk path: node i of the list goes to GPU i mod G, no exchange.  Pass 2 -- the seam
# meshes -- reads the seam nodes of a host's up to 8 neighbours, which other GPUs may have built:
# the one real exchange step of the widened path, an all-gather of every rank's SeamNodeInfo
# records (48 B each, a few hundred per node) and of three integers per node.  After it every rank
# knows the whole node table, derives the same seam-update set and contours its share of it.
# ---------------------------------------------------------------------------------------------
SEAM_NODE_BYTES = 48


def exchange_seam_nodes(num_nodes, mine, local_triangles, local_seam_counts, local_seam_offsets, local_seam_nodes, group=None):
    """mine: global indices of this rank's nodes; local_*: per such node its triangle count, its
    SeamNodeInfo count and where those records start in local_seam_nodes (a numpy structured
    array, or a torch uint8 tensor of SEAM_NODE_BYTES-byte records on the rank's GPU).
    Returns (table int64[num_nodes, 3] = numTriangles, numSeamNodes, firstSeamNode in the gathered
    arena; arena; records_per_rank).  The gathered arena is rank-major and padded: rank r's
    records keep their local order at [r * records_per_rank, ...), so no record is moved twice.
    gloo: CPU tensors, the arena comes back as a numpy uint8 array; nccl: CUDA tensors, the arena
    is a CUDA uint8 tensor (pass arena.data_ptr() on: the seam batch reads device memory)."""
    import torch
    import torch.distributed as dist

    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    nccl = distributed and dist.get_backend(group) == "nccl"
    is_tensor = isinstance(local_seam_nodes, torch.Tensor)
    local_records = (local_seam_nodes.numel() // SEAM_NODE_BYTES) if is_tensor else len(local_seam_nodes)

    table = torch.zeros((num_nodes, 4), dtype=torch.int64)          # triangles, seam nodes, local offset, owner rank
    idx = torch.as_tensor(np.asarray(mine), dtype=torch.long)
    table[idx, 0] = torch.as_tensor(np.asarray(local_triangles, np.int64))
    table[idx, 1] = torch.as_tensor(np.asarray(local_seam_counts, np.int64))
    table[idx, 2] = torch.as_tensor(np.asarray(local_seam_offsets, np.int64))
    table[idx, 3] = rank
    sizes = torch.zeros(world, dtype=torch.int64)
    sizes[rank] = local_records
    if distributed:
        t, s = (table.cuda(), sizes.cuda()) if nccl else (table, sizes)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)       # shards are disjoint: sum == gather
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
        table, sizes = t.cpu(), s.cpu()
    per_rank = max(int(sizes.max()), 1)

    if is_tensor:
        send = local_seam_nodes.reshape(-1)
    else:
        send = torch.from_numpy(np.ascontiguousarray(local_seam_nodes).view(np.uint8).reshape(-1))
    if not distributed:       # one rank: the local records are the arena
        out = torch.stack([table[:, 0], table[:, 1], table[:, 2]], dim=1).numpy()
        out[out[:, 1] == 0, 2] = 0
        return out, (send if send.is_cuda else send.numpy()), max(local_records, 1)
    if nccl and not send.is_cuda:
        send = send.cuda()
    padded = torch.zeros(per_rank * SEAM_NODE_BYTES, dtype=torch.uint8, device=send.device)
    padded[:local_records * SEAM_NODE_BYTES] = send[:local_records * SEAM_NODE_BYTES]
    if distributed:
        arena = torch.empty(world * per_rank * SEAM_NODE_BYTES, dtype=torch.uint8, device=send.device)
        dist.all_gather_into_tensor(arena, padded, group=group)
    else:
        arena = padded
    out = torch.stack([table[:, 0], table[:, 1], table[:, 3] * per_rank + table[:, 2]], dim=1).numpy()
    out[out[:, 1] == 0, 2] = 0
    return out, (arena if arena.is_cuda else arena.numpy()), per_rank


def sharded_clipmap_update(lc, ctx, node_min_size, vertices, triangles, seam_nodes, unit_options=None, group=None):
    """One update that loads `node_min_size` (int32[n][4], no node active before) on G GPUs:
    pass 1 on this rank's nodes (i mod G), the seam-node exchange, pass 2 on this rank's share of
    the seam-update set.  Host arenas of this rank: node meshes first, its seam meshes after them.
    Returns dict(mine, results, table, seam_update_nodes, seam_results, num_seam_updates_all, node_totals)."""
    import torch
    import torch.distributed as dist

    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    ms = np.ascontiguousarray(node_min_size, np.int32).reshape(-1, 4)
    mine = shard_round_robin(len(ms), rank, world)
    rc, res, simp = ctx.generateSimplifiedBatch(ms[mine], vertices, triangles, seam_nodes, unit_options)
    if rc < 0:
        raise RuntimeError(f"generateSimplifiedBatch: {lc.GetCLErrorString(rc)}")
    used = int(res["numSeamNodes"].sum())
    nv, nt = int(res["numVertices"].sum()), int(res["numTriangles"].sum())
    local = seam_nodes[:max(used, 1)]
    if distributed and dist.get_backend(group) == "nccl":
        local = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).reshape(-1)).cuda()
    table, arena, per_rank = exchange_seam_nodes(len(ms), mine, res["numTriangles"], res["numSeamNodes"], res["seamOffset"], local, group)
    nodes = np.zeros(len(ms), lc.ClipmapNode)
    nodes["min"] = ms[:, :3]; nodes["size"] = ms[:, 3]
    nodes["firstSeamNode"] = table[:, 2]; nodes["numSeamNodes"] = table[:, 1]
    active = np.nonzero((table[:, 0] > 0) | (table[:, 1] > 0))[0].astype(np.int32)
    arena_arg = int(arena.data_ptr()) if isinstance(arena, torch.Tensor) else arena.view(lc.SeamNodeInfo)
    rc, upd, sres, n_all = lc.ClipmapSeamUpdateBatch(ctx.voxelsPerChunk(), nodes, active, active, arena_arg, world * per_rank,
                                                     vertices[nv:], triangles[nt:], rank, world)
    if rc < 0:
        raise RuntimeError(f"ClipmapSeamUpdateBatch: {lc.GetCLErrorString(rc)} {lc.lib().lvn_seam_last_error()}")
    sres = sres.copy()
    sres["vertexOffset"] += nv; sres["triangleOffset"] += nt
    return dict(mine=mine, results=res, table=table, seam_update_nodes=upd, seam_results=sres, num_seam_updates_all=n_all,
                node_totals=(nv, nt), arena=arena)
