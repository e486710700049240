"""Clipmap covers for the seam-mesh tests: sets of active nodes (min, size) around the terrain
surface, and the neighbour listing GenerateClipmapSeamMesh (clipmap.cpp:573-611) makes for a host
node over such a cover (findNode + findActiveNodes, clipmap.cpp:1011-1014,1449-1481)."""
CHILD_MIN_OFFSETS = [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)]


def neighbours_for(host_min, host_size, active, seams_of):
    """[(slot index, node min, node size, SeamNodeInfo array)]: for each of the host's 8 slots, the
    active nodes that contain the slot's min or whose min lies inside the slot (FindActiveNodes);
    a coarser node therefore appears once per slot it covers, as in the reference."""
    out = []
    for i, off in enumerate(CHILD_MIN_OFFSETS):
        cmin = [host_min[k] + off[k] * host_size for k in range(3)]
        for (amin, asize) in active:
            a_contains_c = all(amin[k] <= cmin[k] < amin[k] + asize for k in range(3))
            c_contains_a = all(cmin[k] <= amin[k] < cmin[k] + host_size for k in range(3))
            if a_contains_c or c_contains_a:
                sn = seams_of(amin, asize)
                if len(sn):
                    out.append((i, list(amin), asize, sn))
    return out


def uniform_lod0(cy):
    return [((cx * 256, (cy + dy) * 256, cz * 256), 256) for cx in range(-2, 2) for dy in (-1, 0, 1) for cz in range(-2, 2)]


def mixed_lod01(cy):
    """a 3x3x3 block of LOD1 nodes whose centre node is split into its 8 LOD0 children"""
    y1 = (cy * 256 // 512) * 512
    out = []
    for px in (-512, 0, 512):
        for py in (y1 - 512, y1, y1 + 512):
            for pz in (-512, 0, 512):
                if (px, py, pz) == (0, y1, 0):
                    out += [((px + ox * 256, py + oy * 256, pz + oz * 256), 256) for ox in (0, 1) for oy in (0, 1) for oz in (0, 1)]
                else:
                    out.append(((px, py, pz), 512))
    return out


def mixed_lod012(cy):
    """LOD2 nodes, one split into LOD1 nodes, one of those split into LOD0 nodes"""
    y2 = (cy * 256 // 1024) * 1024
    out = []
    for px in (-1024, 0):
        for pz in (-1024, 0):
            if (px, pz) != (0, 0):
                out.append(((px, y2, pz), 1024))
                continue
            for ox in (0, 1):
                for oy in (0, 1):
                    for oz in (0, 1):
                        q = (px + ox * 512, y2 + oy * 512, pz + oz * 512)
                        if (ox, oz) == (0, 0):
                            out += [((q[0] + a * 256, q[1] + b * 256, q[2] + c * 256), 256) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
                        else:
                            out.append((q, 512))
    return out


def collision_nodes(cy):
    """the collision feed: Clipmap::loadCollisionNodes / GenerateCollisionSeamMesh
    (clipmap.cpp:474-503,613-640) mesh nodes of COLLISION_NODE_SIZE = 512 world units with a 64-voxel
    context (sampleScale 2) and stitch them with the same seam code, all neighbours at that one size"""
    y1 = (cy * 256 // 512) * 512
    return [((px, py, pz), 512) for px in (-1024, -512, 0, 512) for py in (y1 - 512, y1) for pz in (-1024, -512, 0, 512)]


SCENARIOS = {"uniform_lod0": uniform_lod0, "mixed_lod01": mixed_lod01, "mixed_lod012": mixed_lod012,
             "collision_nodes": collision_nodes}


def build_jobs(active, seams_of):
    return [(list(mn), size, neighbours_for(mn, size, active, seams_of)) for (mn, size) in active]
