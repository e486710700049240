// CSG brush edits on a cached density field (a16).
//
// Reference functions restated (paths relative to the reference tree):
//   BrushDensity / Density_Cuboid / Density_Sphere   leven/cl/apply_csg_operation.cl:42-82
//   pR, fBox, vmax3                                  leven/cl/hg_sdf.glsl:164-166,218-221,460-463
//   BrushMaterial / CSG_HermiteIndices / UpdateFieldMaterials   apply_csg_operation.cl:114-241
//   FindUpdatedEdges ... PruneFieldEdges / CompactFieldEdges    apply_csg_operation.cl:253-437
//   BrushZeroCrossing / BrushNormal / FindEdgeIntersectionInfo  apply_csg_operation.cl:86-175,443-477
//   host sequence ApplyCSGOperations                 leven/src/compute_csg.cpp:11-220
//
// B200 shape of the same result: the reference builds the set of edges touching a changed
// sample with 6-way expansion, compaction and a multi-round hash dedupe, then tests every old
// edge against that list linearly (O(E*K)).  Here the set is a bitmap over the 3*H^3 Hermite
// edges, so "dedupe" is an atomicOr and "prune" is one bit test per old edge.
#include <float.h>

#include "density.cuh"

namespace lvn {

__device__ __forceinline__ float length3(float x, float y, float z) { return sqrtf((x * x + y * y) + z * z); }

__device__ __forceinline__ float brush_density(float x, float y, float z, const CsgOpDev &op)
{
    const float lx = x - op.ox, ly = y - op.oy, lz = z - op.oz;
    if (op.shape == 0) {
        const float rx = op.c * lx + op.s * lz;      // pR on (x, z); cos/sin computed on the host
        const float rz = op.c * lz + op.s * (-lx);
        const float dx = fabsf(rx) - op.dx, dy = fabsf(ly) - op.dy, dz = fabsf(rz) - op.dz;
        const float outside = length3(fmaxf(dx, 0.f), fmaxf(dy, 0.f), fmaxf(dz, 0.f));
        const float inside = fmaxf(fmaxf(fminf(dx, 0.f), fminf(dy, 0.f)), fminf(dz, 0.f));
        return outside + inside;
    }
    return length3(lx, ly, lz) - op.dx;
}

__device__ __forceinline__ int edge_bit_index(int x, int y, int z, int axis, int H)
{
    return (x + H * (y + H * z)) * 3 + axis;
}

// All three kernels are batched: blockIdx.y selects the chunk (CsgChunk), so that an edit that
// overlaps 27 or 64 chunks is three launches and two host waits, not that many per chunk.

// CSG_HermiteIndices + UpdateFieldMaterials + FindUpdatedEdges in one pass over the field
__global__ void k_csg_materials(Dims d, const CsgChunk *__restrict__ chunks, const CsgOpDev *__restrict__ allOps)
{
    const CsgChunk cc = chunks[blockIdx.y];
    const CsgOpDev *ops = allOps + cc.opFirst;
    uint8_t *field = cc.field;
    unsigned int *touched = cc.touched;
    const int F = d.F, H = d.H, F3 = F * F * F;
    unsigned int changed = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F3; i += gridDim.x * blockDim.x) {
        const int x = i % F, y = (i / F) % F, z = i / (F * F);
        const int oldMaterial = field[i];
        const float wx = (float)(cc.ox + cc.scale * x), wy = (float)(cc.oy + cc.scale * y), wz = (float)(cc.oz + cc.scale * z);
        int m = oldMaterial;
        for (int k = 0; k < cc.numOps; k++) {   // BrushMaterial: the last op with d <= 0 wins
            const CsgOpDev op = ops[k];
            if (brush_density(wx, wy, wz, op) <= 0.f) m = op.type == 0 ? op.material : LVN_MATERIAL_AIR;
        }
        if (m == oldMaterial) continue;
        field[i] = (uint8_t)m;
        changed++;
        const int p[3] = {x, y, z};
        if (x < H && y < H && z < H)
            for (int k = 0; k < 3; k++) {
                const int e = edge_bit_index(x, y, z, k, H);
                atomicOr(&touched[e >> 5], 1u << (e & 31));
            }
        for (int k = 0; k < 3; k++) {
            if (p[k] <= 0) continue;
            const int qx = x - (k == 0), qy = y - (k == 1), qz = z - (k == 2);
            if (qx < H && qy < H && qz < H) {
                const int e = edge_bit_index(qx, qy, qz, k, H);
                atomicOr(&touched[e >> 5], 1u << (e & 31));
            }
        }
    }
    changed = __reduce_add_sync(0xffffffffu, changed);   // one atomic per warp: the counters of a chunk are three addresses
    if (changed && (threadIdx.x & 31) == 0) atomicAdd(&cc.counts[2], changed);
}

__device__ __forceinline__ bool edge_sign_change(const uint8_t *__restrict__ field, int F, int x, int y, int z, int axis)
{
    const int m0 = field[x + F * (y + F * z)];
    const int m1 = field[(x + (axis == 0)) + F * ((y + (axis == 1)) + F * (z + (axis == 2)))];
    return (m0 == LVN_MATERIAL_AIR) != (m1 == LVN_MATERIAL_AIR);
}

__device__ __forceinline__ int key_to_bit(int key, const Dims &d)
{
    const int axis = key & 3, idx = key >> 2;
    return edge_bit_index(idx & d.mask, (idx >> d.shift) & d.mask, (idx >> (d.shift * 2)) & d.mask, axis, d.H);
}

// counts[0] = old edges that survive the prune; counts[1] = touched edges that now change sign
__global__ void k_csg_count(Dims d, const CsgChunk *__restrict__ chunks, unsigned int *ticket, unsigned int *hostCounts)
{
    const CsgChunk cc = chunks[blockIdx.y];
    const uint8_t *field = cc.field;
    const unsigned int *touched = cc.touched;
    const int H = d.H, numBits = 3 * H * H * H, numWords = (numBits + 31) / 32;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    unsigned int kept = 0, created = 0;
    for (int i = gid; i < cc.numOld; i += stride) {
        const int e = key_to_bit(cc.oldKeys[i], d);
        if (!((touched[e >> 5] >> (e & 31)) & 1u)) kept++;
    }
    // the bitmap: every non-empty word is looked at by a whole warp, lane b taking bit b (the touched edges of a
    // brush sit in a few hundred neighbouring words: a thread that walked the set bits of its own word did 32 edges
    // one after the other while most of the grid had none).  Neighbouring words go to different warps -- warp v
    // takes words v, v + W, v + 2W, ... (W warps), fetched 32 at a time, one per lane
    const int lane32 = threadIdx.x & 31, warpId = gid >> 5, numWarps = stride >> 5;
    for (int base = 0; base < numWords; base += numWarps * 32) {
        const int myWord = base + warpId + lane32 * numWarps;
        const unsigned int mine = myWord < numWords ? touched[myWord] : 0u;
        unsigned int nonEmpty = __ballot_sync(0xffffffffu, mine != 0u);
        while (nonEmpty) {
            const int src = __ffs((int)nonEmpty) - 1;
            nonEmpty &= nonEmpty - 1;
            const unsigned int bits = __shfl_sync(0xffffffffu, mine, src);
            if (!((bits >> lane32) & 1u)) continue;
            const int e = (base + warpId + src * numWarps) * 32 + lane32, axis = e % 3, cell = e / 3;
            const int x = cell % H, y = (cell / H) % H, z = cell / (H * H);
            if (edge_sign_change(field, d.F, x, y, z, axis)) created++;
        }
    }
    kept = __reduce_add_sync(0xffffffffu, kept);
    created = __reduce_add_sync(0xffffffffu, created);
    if ((threadIdx.x & 31) == 0) {
        if (kept) atomicAdd(&cc.counts[0], kept);
        if (created) atomicAdd(&cc.counts[1], created);
    }
    // the launch's last block hands every chunk's counters to the host (mapped pinned mirror)
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const unsigned int *all = chunks[0].counts;   // chunk i's counters are at all + 8 i
    for (int i = threadIdx.x; i < 8 * (int)gridDim.y; i += blockDim.x) hostCounts[i] = __ldcg(&all[i]);
    __threadfence_system();
}

// PruneFieldEdges + CompactFieldEdges, then the CSG FindEdgeIntersectionInfo for created edges.
// counts[4] counts kept edges, counts[5] created edges (appended after the kept ones: the final
// kept count is known from k_csg_count).
__global__ void k_csg_emit(Dims d, const CsgChunk *__restrict__ chunks, const CsgOpDev *__restrict__ allOps)
{
    const CsgChunk cc = chunks[blockIdx.y];
    if (cc.skip) return;
    const CsgOpDev *ops = allOps + cc.opFirst;
    const int numOps = cc.numOps;
    const uint8_t *field = cc.field;
    const unsigned int *touched = cc.touched;
    const int H = d.H, numBits = 3 * H * H * H, numWords = (numBits + 31) / 32;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x, lane32 = threadIdx.x & 31;
    // kept edges: one atomic per warp (the order of the new list is arbitrary, as in the reference)
    for (int i0 = gid - lane32; i0 < cc.numOld; i0 += stride) {
        const int i = i0 + lane32;
        int key = 0;
        bool keep = false;
        if (i < cc.numOld) {
            key = cc.oldKeys[i];
            const int e = key_to_bit(key, d);
            keep = !((touched[e >> 5] >> (e & 31)) & 1u);
        }
        const unsigned int bal = __ballot_sync(0xffffffffu, keep);
        if (!bal) continue;
        unsigned int base = 0;
        if (lane32 == 0) base = atomicAdd(&cc.counts[4], (unsigned int)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
            const unsigned int o = base + __popc(bal & ((1u << lane32) - 1u));
            cc.newKeys[o] = key;
            cc.newInfo[o] = cc.oldInfo[i];
        }
    }
    // created edges: one lane per bit of every non-empty word, neighbouring words on different warps (see
    // k_csg_count), one atomic per warp and word
    const int warpId = gid >> 5, numWarps = stride >> 5;
    for (int wbase = 0; wbase < numWords; wbase += numWarps * 32) {
      const int myWord = wbase + warpId + lane32 * numWarps;
      const unsigned int mine = myWord < numWords ? touched[myWord] : 0u;
      unsigned int nonEmpty = __ballot_sync(0xffffffffu, mine != 0u);
      while (nonEmpty) {
        const int src = __ffs((int)nonEmpty) - 1;
        nonEmpty &= nonEmpty - 1;
        const unsigned int bits = __shfl_sync(0xffffffffu, mine, src);
        const int e = (wbase + warpId + src * numWarps) * 32 + lane32, axis = e % 3, cell = e / 3;
        const int x = cell % H, y = (cell / H) % H, z = cell / (H * H);
        const bool create = ((bits >> lane32) & 1u) && edge_sign_change(field, d.F, x, y, z, axis);
        const unsigned int bal = __ballot_sync(0xffffffffu, create);
        if (!bal) continue;
        unsigned int base = 0;
        if (lane32 == 0) base = atomicAdd(&cc.counts[5], (unsigned int)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (!create) continue;
        // FindEdgeIntersectionInfo, apply_csg_operation.cl:443-477
        const int wx = (cc.scale * x) + cc.ox, wy = (cc.scale * y) + cc.oy, wz = (cc.scale * z) + cc.oz;
        const float p0x = (float)wx, p0y = (float)wy, p0z = (float)wz;
        const float p1x = (float)(wx + (axis == 0 ? cc.scale : 0)), p1y = (float)(wy + (axis == 1 ? cc.scale : 0)),
                    p1z = (float)(wz + (axis == 2 ? cc.scale : 0));
        // BrushZeroCrossing: first minimum of |density| over 17 steps x ops
        float minDensity = FLT_MAX, crossing = 0.f;
        for (float t = 0.f; t <= 1.f; t += (1.f / 16.f)) {
            const float px = mixf(p0x, p1x, t), py = mixf(p0y, p1y, t), pz = mixf(p0z, p1z, t);
            for (int k = 0; k < numOps; k++) {
                const float dd = fabsf(brush_density(px, py, pz, ops[k]));
                if (dd < minDensity) { crossing = t; minDensity = dd; }
            }
        }
        const float px = mixf(p0x, p1x, crossing), py = mixf(p0y, p1y, crossing), pz = mixf(p0z, p1z, crossing);
        // BrushNormal: the last op whose density at p is <= 0, flipped for subtract
        float nx = 0.f, ny = 0.f, nz = 0.f;
        const float h = 0.001f;
        for (int k = 0; k < numOps; k++) {
            const CsgOpDev op = ops[k];
            if (brush_density(px, py, pz, op) > 0.f) continue;
            float gx = brush_density(px + h, py, pz, op) - brush_density(px - h, py, pz, op);
            float gy = brush_density(px, py + h, pz, op) - brush_density(px, py - h, pz, op);
            float gz = brush_density(px, py, pz + h, op) - brush_density(px, py, pz - h, op);
            const float flip = op.type == 0 ? 1.f : -1.f;
            normalize3(gx, gy, gz);
            nx = flip * gx; ny = flip * gy; nz = flip * gz;
        }
        const unsigned int o = (unsigned int)cc.numKept + base + __popc(bal & ((1u << lane32) - 1u));
        cc.newKeys[o] = ((x | (y << d.shift) | (z << (d.shift * 2))) << 2) | axis;
        cc.newInfo[o] = make_float4(nx, ny, nz, crossing);
      }
    }
}

void launch_csg_materials_count(const Dims &d, const CsgChunk *chunks, int n, const CsgOpDev *ops, unsigned int *ticket,
                                unsigned int *hostCounts, cudaStream_t s)
{
    if (n <= 0) return;
    k_csg_materials<<<dim3(296, n), 256, 0, s>>>(d, chunks, ops);
    k_csg_count<<<dim3(148, n), 256, 0, s>>>(d, chunks, ticket, hostCounts);
}

void launch_csg_emit(const Dims &d, const CsgChunk *chunks, int n, const CsgOpDev *ops, cudaStream_t s)
{
    if (n <= 0) return;
    k_csg_emit<<<dim3(296, n), 256, 0, s>>>(d, chunks, ops);
}

}  // namespace lvn
