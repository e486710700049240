// simplify.cu -- the mesh simplifier that runs on every chunk mesh right after export
// (ngMeshSimplifier, leven/src/ng_mesh_simplify.cpp:441-540; called at clipmap.cpp:449-465,495-501),
// on the GPU: SURVEY.md 8f-2.
//
// The reference's algorithm is already iteration-parallel in spirit (random candidate edges, each
// vertex keeps its cheapest collapse, an edge collapses when both ends chose it); its loops are
// sequential only in how they compact.  Here one thread block owns one mesh (a batch of meshes is
// one launch) and every sequential loop becomes a block-wide pass with the same result:
//   BuildCandidateEdges   std::sort by (max, min) -> counting sort by max vertex + a tiny sort of each
//                         vertex's bucket by min; duplicate runs, boundary marks and the stable filter
//                         by scans.  (The reference never flushes the last run of its scan: the
//                         greatest edge is neither a candidate nor a boundary mark.  Kept.)
//   FindValidCollapses    the candidate sample is std::mt19937(42) through libstdc++'s
//                         std::uniform_int_distribution (Lemire multiply-shift with rejection),
//                         reproduced from the precomputed raw stream; "first cheapest edge wins" in
//                         ascending edge order = atomicMin on (error bits, edge index)
//   CollapseEdges         one thread per vertex; the winner re-solves its 2-point QEF
//   RemoveTriangles / RemoveEdges / CompactVertices   remap + stable compaction (scan + scatter)
// The 4-D QEF solve restates qef_simd.h lane by lane, including what look like slips in
// rotateq_xy (vtav[0][0] written twice, `cc + v`): the reference's results are the bar.
// _mm_rsqrt_ps := 1 / sqrt(x) (arithmetic spec; the x86 estimate differs between CPU vendors).
#include <algorithm>
#include <cstring>
#include <random>
#include <vector>

#include "lvn_internal.h"

namespace lvn {

constexpr int SIMP_BLOCK = 512;
constexpr int SIMP_RAW = 1 << 17;        // raw mt19937(42) outputs kept on the device
constexpr int SIMP_MAX_DEGREE = 16;      // COLLAPSE_MAX_DEGREE

struct SimpOptionsDev { float edgeFraction; int maxIterations; float targetPercentage, maxError, maxEdgeSize, minAngleCosine; };

struct SimpJobDev {
    int vertexOffset, numVertices;       // into the packed input arrays
    int triangleOffset, numTriangles;
    float offset[4];                     // worldSpaceOffset
    SimpOptionsDev opt;                  // this mesh's options (they scale with the node's leaf size, clipmap.cpp:455-462)
    // scratch slices (element offsets)
    long long edgeOff;                   // capacity 3 * numTriangles (x2 buffers, bucket, flags)
    int vtxOff;                          // capacity numVertices
};

struct SimpScratch {
    float4 *vx, *vn, *vc;                // working vertices
    int *tri[2];                         // triangle ping-pong, 3 ints each
    uint2 *edge[2];                      // (min, max) ping-pong
    unsigned int *bucket;                // per raw edge: min, grouped by max
    int *eflag, *escan;                  // per edge / triangle flags and scans
    int *vcount, *vstart, *vfill;        // per vertex: bucket count / start / cursor, later triangle counts
    int *boundary, *target, *vflag, *vscan;
    unsigned long long *best;            // per vertex (error bits << 32 | edge)
    const unsigned int *raw;             // mt19937(42) outputs
};

// ---- block-wide exclusive scan of n ints in global memory (in -> out), returns the total ----
__device__ int block_scan(const int *in, int *out, int n, int *s_warp, int *s_run)
{
    const int tid = threadIdx.x;
    if (tid == 0) *s_run = 0;
    __syncthreads();
    for (int base = 0; base < n; base += SIMP_BLOCK) {
        const int i = base + tid;
        const int v = i < n ? in[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t; }
        if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < (tid >> 5); w++) woff += s_warp[w];
        if (i < n) out[i] = *s_run + woff + incl - v;
        __syncthreads();
        if (tid == SIMP_BLOCK - 1) *s_run += woff + incl;
        __syncthreads();
    }
    return *s_run;
}

// ---- qef_simd.h, lane by lane -------------------------------------------------------------
struct V4 { float f[4]; };
struct M4 { float m[4][4]; };
__device__ __forceinline__ float v4_dot(const V4 &a, const V4 &b)
{   // vec4_dot: mul, pairwise shuffle-add, reversed shuffle-add -> lane 0
    const float m0 = a.f[0] * b.f[0], m1 = a.f[1] * b.f[1], m2 = a.f[2] * b.f[2], m3 = a.f[3] * b.f[3];
    return (m0 + m1) + (m3 + m2);
}
__device__ __forceinline__ V4 v4_mul_m4(const V4 &a, const M4 &B)
{   // vec4_mul_m4x4
    V4 r;
#pragma unroll
    for (int l = 0; l < 4; l++) {
        float t = a.f[0] * B.m[0][l];
        t = t + a.f[1] * B.m[1][l];
        t = t + a.f[2] * B.m[2][l];
        t = t + a.f[3] * B.m[3][l];
        r.f[l] = t;
    }
    return r;
}
__device__ __forceinline__ void qef4_givens(float c[3], float s[3], const M4 &vtav)
{   // givens_coeffs_sym: three (p, q) pairs at once, all from the same vtav
    const float pp[3] = {vtav.m[0][0], vtav.m[0][0], vtav.m[1][1]};
    const float pq[3] = {vtav.m[0][1], vtav.m[0][2], vtav.m[1][2]};
    const float qq[3] = {vtav.m[1][1], vtav.m[2][2], vtav.m[2][2]};
#pragma unroll
    for (int l = 0; l < 3; l++) {
        if (pq[l] == 0.f) { c[l] = 1.f; s[l] = 0.f; continue; }
        const float pq2 = pq[l] * 2.f;
        const float tau = (qq[l] - pp[l]) / pq2;
        const float stt = sqrtf(tau * tau + 1.f);
        const float tan_inv = (tau >= 0.f) ? (tau + stt) : (tau - stt);
        const float tan_ = 1.f / tan_inv;
        const float cc = 1.f / sqrtf(1.f + tan_ * tan_);     // _mm_rsqrt_ps := 1 / sqrt
        c[l] = cc;
        s[l] = tan_ * cc;
    }
}
__device__ __forceinline__ void qef4_rotateq(M4 &vtav, const float c[3], const float s[3])
{   // rotateq_xy, as written (including `y1 = x0 + mx`, `y2 = cc + v` and the double write of [0][0])
    const float u[3] = {vtav.m[0][0], vtav.m[0][0], vtav.m[1][1]};
    const float v[3] = {vtav.m[1][1], vtav.m[2][2], vtav.m[2][2]};
    const float a[3] = {vtav.m[0][1], vtav.m[0][2], vtav.m[1][2]};
    float x[3], y[3];
#pragma unroll
    for (int l = 0; l < 3; l++) {
        const float cc = c[l] * c[l], ss = s[l] * s[l];
        const float c2 = 2.f * c[l];
        const float c2s = c2 * s[l];
        const float mx = c2s * a[l];
        const float x0 = cc * u[l];
        const float x1 = x0 - mx;
        const float x2 = ss * v[l];
        x[l] = x1 + x2;
        const float y1 = x0 + mx;
        const float y2 = cc + v[l];
        y[l] = y1 + y2;
    }
    vtav.m[0][0] = x[0];
    vtav.m[0][0] = x[1];
    vtav.m[1][1] = x[2];
    vtav.m[0][1] = y[0];
    vtav.m[0][2] = y[1];
    vtav.m[1][2] = y[2];
}
template <int A, int B>
__device__ __forceinline__ void qef4_rotate(M4 &vtav, M4 &v, float c, float s)
{   // svd_rotate
    if (vtav.m[A][B] == 0.f) return;
    const float u[4] = {v.m[0][A], v.m[1][A], v.m[2][A], vtav.m[0][3 - B]};
    const float w[4] = {v.m[0][B], v.m[1][B], v.m[2][B], vtav.m[1 - A][2]};
    float x[4], y[4];
#pragma unroll
    for (int l = 0; l < 4; l++) {
        x[l] = c * u[l] - s * w[l];
        y[l] = s * u[l] + c * w[l];
    }
    v.m[0][A] = x[0]; v.m[1][A] = x[1]; v.m[2][A] = x[2];
    vtav.m[0][3 - B] = x[3];
    v.m[0][B] = y[0]; v.m[1][B] = y[1]; v.m[2][B] = y[2];
    vtav.m[1 - A][2] = y[3];
    vtav.m[A][B] = 0.f;
}
// qef_solve_from_points with two points: returns the error, the solved position in `out`
__device__ float qef4_solve2(const float4 p0, const float4 n0, const float4 p1, const float4 n1, float out[4])
{
    M4 ATA;
    V4 ATb, acc;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int l = 0; l < 4; l++) ATA.m[r][l] = 0.f;
#pragma unroll
    for (int l = 0; l < 4; l++) { ATb.f[l] = 0.f; acc.f[l] = 0.f; }
#pragma unroll
    for (int k = 0; k < 2; k++) {   // qef_simd_add
        const float4 pp = k ? p1 : p0, nn = k ? n1 : n0;
        const V4 p = {{pp.x, pp.y, pp.z, pp.w}}, n = {{nn.x, nn.y, nn.z, nn.w}};
#pragma unroll
        for (int l = 0; l < 4; l++) {
            ATA.m[0][l] = ATA.m[0][l] + n.f[0] * n.f[l];
            ATA.m[1][l] = ATA.m[1][l] + n.f[1] * n.f[l];
            ATA.m[2][l] = ATA.m[2][l] + n.f[2] * n.f[l];
        }
        const float d = v4_dot(p, n);
        const float xd[4] = {d, d, d, 0.f};
#pragma unroll
        for (int l = 0; l < 4; l++) { ATb.f[l] = ATb.f[l] + xd[l] * n.f[l]; acc.f[l] = acc.f[l] + p.f[l]; }
    }
    // qef_simd_solve
    V4 mp;
#pragma unroll
    for (int l = 0; l < 4; l++) mp.f[l] = acc.f[l] / acc.f[3];
    V4 p = v4_mul_m4(mp, ATA);
#pragma unroll
    for (int l = 0; l < 4; l++) p.f[l] = ATb.f[l] - p.f[l];
    // svd_solve_ATA_ATb
    M4 V;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int l = 0; l < 4; l++) V.m[r][l] = (r == l && r < 3) ? 1.f : 0.f;
    M4 vtav = ATA;
    for (int i = 0; i < 5; ++i) {   // SVD_NUM_SWEEPS
        float c[3], s[3];
        qef4_givens(c, s, vtav);
        qef4_rotateq(vtav, c, s);
        qef4_rotate<0, 1>(vtav, V, c[0], s[0]);
        qef4_rotate<0, 2>(vtav, V, c[1], s[1]);
        qef4_rotate<1, 2>(vtav, V, c[2], s[2]);
    }
    const float sigma[4] = {vtav.m[0][0], vtav.m[1][1], vtav.m[2][2], 0.f};
    float invdet[4];
#pragma unroll
    for (int l = 0; l < 4; l++) {   // svd_invdet
        const float ax = fabsf(sigma[l]);
        const float inv = 1.f / sigma[l];
        const float ai = fabsf(inv);
        const float mn = ax < ai ? ax : ai;            // _mm_min_ps
        invdet[l] = (mn >= 0.1f) ? inv : 0.f;
    }
    M4 o;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int l = 0; l < 4; l++) o.m[r][l] = V.m[r][l] * invdet[l];
#pragma unroll
    for (int l = 0; l < 4; l++) o.m[3][l] = 0.f;
    M4 Vinv;
#pragma unroll
    for (int r = 0; r < 4; r++) {   // m4x4_mul_m4x4(o, o, v), row by row
        const V4 a = {{o.m[r][0], o.m[r][1], o.m[r][2], o.m[r][3]}};
        const V4 t = v4_mul_m4(a, V);
#pragma unroll
        for (int l = 0; l < 4; l++) Vinv.m[r][l] = t.f[l];
    }
    V4 x = v4_mul_m4(p, Vinv);
    // qef_simd_calc_error
    V4 tmp = v4_mul_m4(x, ATA);
#pragma unroll
    for (int l = 0; l < 4; l++) tmp.f[l] = ATb.f[l] - tmp.f[l];
    const float error = v4_dot(tmp, tmp);
#pragma unroll
    for (int l = 0; l < 4; l++) out[l] = x.f[l] + mp.f[l];
    return error;
}

__device__ __forceinline__ float dot4_lr(const float4 a, const float4 b)   // GLM 0.9.3 dot(vec4, vec4): left to right
{
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

__global__ void __launch_bounds__(SIMP_BLOCK)
k_simplify(const SimpJobDev *__restrict__ jobs, SimpScratch ws,
           const lvn_mesh_vertex *__restrict__ inV, const int *__restrict__ inT,
           lvn_mesh_vertex *__restrict__ outV, int *__restrict__ outT, int4 *__restrict__ results)
{
    __shared__ int s_warp[SIMP_BLOCK / 32], s_run, s_cnt, s_last[2], s_bad;
    const SimpJobDev job = jobs[blockIdx.x];
    const SimpOptionsDev opt = job.opt;
    const int tid = threadIdx.x, NV = job.numVertices, NT0 = job.numTriangles;
    const lvn_mesh_vertex *srcV = inV + job.vertexOffset;
    const int *srcT = inT + (size_t)job.triangleOffset * 3;
    lvn_mesh_vertex *dstV = outV + job.vertexOffset;
    int *dstT = outT + (size_t)job.triangleOffset * 3;

    // an index outside the mesh's vertices would be a wild write below (the reference would crash):
    // such a mesh is passed through untouched and reported with iterations = -1
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int i = tid; i < NT0 * 3; i += SIMP_BLOCK) if ((unsigned int)srcT[i] >= (unsigned int)NV) s_bad = 1;
    __syncthreads();
    const bool bad = s_bad != 0;
    if (bad || NT0 < 100 || NV < 100) {   // ng_mesh_simplify.cpp:446-449: too small, returned untouched
        for (int i = tid; i < NV * 3; i += SIMP_BLOCK) reinterpret_cast<float4 *>(dstV)[i] = reinterpret_cast<const float4 *>(srcV)[i];
        for (int i = tid; i < NT0 * 3; i += SIMP_BLOCK) dstT[i] = srcT[i];
        if (tid == 0) results[blockIdx.x] = make_int4(NV, NT0, bad ? -1 : 0, 0);
        return;
    }
    float4 *vx = ws.vx + job.vtxOff, *vn = ws.vn + job.vtxOff, *vc = ws.vc + job.vtxOff;
    int *tri[2] = {ws.tri[0] + job.edgeOff, ws.tri[1] + job.edgeOff};   // 3 ints per triangle: same capacity as edges
    uint2 *edge[2] = {ws.edge[0] + job.edgeOff, ws.edge[1] + job.edgeOff};
    unsigned int *bucket = ws.bucket + job.edgeOff;
    int *eflag = ws.eflag + job.edgeOff, *escan = ws.escan + job.edgeOff;
    int *vcount = ws.vcount + job.vtxOff, *vstart = ws.vstart + job.vtxOff, *vfill = ws.vfill + job.vtxOff;
    int *boundary = ws.boundary + job.vtxOff, *target = ws.target + job.vtxOff, *vflag = ws.vflag + job.vtxOff, *vscan = ws.vscan + job.vtxOff;
    unsigned long long *best = ws.best + job.vtxOff;
    const float4 off = make_float4(job.offset[0], job.offset[1], job.offset[2], job.offset[3]);

    // ---- copy in; v.xyz -= worldSpaceOffset; per-vertex triangle counts ----
    for (int i = tid; i < NV; i += SIMP_BLOCK) {
        const float4 *p = reinterpret_cast<const float4 *>(&srcV[i]);
        const float4 x = p[0];
        vx[i] = make_float4(x.x - off.x, x.y - off.y, x.z - off.z, x.w - off.w);
        vn[i] = p[1];
        vc[i] = p[2];
        vcount[i] = 0; vfill[i] = 0; boundary[i] = 0; vflag[i] = 0;
    }
    for (int i = tid; i < NT0 * 3; i += SIMP_BLOCK) tri[0][i] = srcT[i];
    __syncthreads();

    // ---- BuildCandidateEdges (ng_mesh_simplify.cpp:122-177) ----
    // raw edge j of triangle t: (0,1), (1,2), (0,2); bucket by max vertex
    const int NE0 = NT0 * 3;
    for (int j = tid; j < NE0; j += SIMP_BLOCK) {
        const int t = j / 3, k = j - t * 3;
        const int a = tri[0][t * 3 + (k == 2 ? 0 : k)], b = tri[0][t * 3 + (k == 0 ? 1 : 2)];
        atomicAdd(&vcount[max(a, b)], 1);
    }
    __syncthreads();
    block_scan(vcount, vstart, NV, s_warp, &s_run);
    for (int j = tid; j < NE0; j += SIMP_BLOCK) {
        const int t = j / 3, k = j - t * 3;
        const int a = tri[0][t * 3 + (k == 2 ? 0 : k)], b = tri[0][t * 3 + (k == 0 ? 1 : 2)];
        const int mx = max(a, b), mn = min(a, b);
        bucket[vstart[mx] + atomicAdd(&vfill[mx], 1)] = (unsigned int)mn;
    }
    if (tid == 0) { s_last[0] = -1; s_last[1] = -1; }
    __syncthreads();
    // the greatest edge of the sorted list: greatest max vertex with a non-empty bucket, its greatest min
    for (int v = tid; v < NV; v += SIMP_BLOCK) if (vcount[v] > 0) atomicMax(&s_last[0], v);
    __syncthreads();
    // each vertex sorts its bucket by min, counts runs: run of 1 -> boundary marks, longer -> one filtered edge
    for (int v = tid; v < NV; v += SIMP_BLOCK) {
        const int n = vcount[v], st = vstart[v];
        for (int i = 1; i < n; i++) {   // insertion sort (a vertex has a dozen raw edges)
            const unsigned int key = bucket[st + i];
            int j = i - 1;
            while (j >= 0 && bucket[st + j] > key) { bucket[st + j + 1] = bucket[st + j]; j--; }
            bucket[st + j + 1] = key;
        }
        int kept = 0;
        for (int i = 0; i < n;) {
            int j = i + 1;
            while (j < n && bucket[st + j] == bucket[st + i]) j++;
            const bool lastRun = (v == s_last[0]) && (j == n);   // never flushed by the reference's scan
            if (!lastRun) {
                if (j - i == 1) { boundary[bucket[st + i]] = 1; boundary[v] = 1; }
                else bucket[st + kept++] = bucket[st + i];
            }
            i = j;
        }
        vfill[v] = kept;   // filtered edges of this bucket, ascending min, at bucket[st .. st + kept)
    }
    __syncthreads();
    block_scan(vfill, vscan, NV, s_warp, &s_run);
    __syncthreads();
    const int numFiltered = s_run;
    for (int v = tid; v < NV; v += SIMP_BLOCK) {
        const int st = vstart[v], o = vscan[v];
        for (int i = 0; i < vfill[v]; i++) edge[1][o + i] = make_uint2(bucket[st + i], (unsigned int)v);
    }
    __syncthreads();
    for (int i = tid; i < numFiltered; i += SIMP_BLOCK) { const uint2 e = edge[1][i]; eflag[i] = !boundary[e.x] && !boundary[e.y]; }
    __syncthreads();
    block_scan(eflag, escan, numFiltered, s_warp, &s_run);
    __syncthreads();
    int NE = s_run;
    for (int i = tid; i < numFiltered; i += SIMP_BLOCK) if (eflag[i]) edge[0][escan[i]] = edge[1][i];
    // vertexTriangleCounts (ng_mesh_simplify.cpp:478-489)
    for (int i = tid; i < NV; i += SIMP_BLOCK) vcount[i] = 0;
    __syncthreads();
    for (int i = tid; i < NT0 * 3; i += SIMP_BLOCK) atomicAdd(&vcount[tri[0][i]], 1);
    __syncthreads();

    int NT = NT0, curT = 0, curE = 0;
    const int targetTriangleCount = (int)((float)NT0 * opt.targetPercentage);
    const float maxEdge2 = opt.maxEdgeSize * opt.maxEdgeSize;
    int iterations = 0;
    while (NT > targetTriangleCount && iterations++ < opt.maxIterations) {
        for (int i = tid; i < NV; i += SIMP_BLOCK) { best[i] = ~0ull; target[i] = -1; }
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        // ---- FindValidCollapses (ng_mesh_simplify.cpp:181-283) ----
        const int numRandom = (int)((float)NE * opt.edgeFraction);
        if (numRandom > 0) {
            // std::uniform_int_distribution<int>(0, NE - 1) over std::mt19937(42), libstdc++:
            // product = raw * range; reject while (uint32)product < (2^32 - range) % range
            const unsigned int range = (unsigned int)NE, threshold = (0u - range) % range;
            const int K = min(numRandom + 256, SIMP_RAW);
            for (int base = 0; base < K; base += SIMP_BLOCK) {
                const int k = base + tid;
                if (k < K) {
                    const unsigned long long prod = (unsigned long long)ws.raw[k] * range;
                    eflag[k] = ((unsigned int)prod >= threshold) ? 1 : 0;   // accepted draw
                }
            }
            __syncthreads();
            block_scan(eflag, escan, K, s_warp, &s_run);
            __syncthreads();
            for (int k = tid; k < K; k += SIMP_BLOCK) {
                if (!eflag[k] || escan[k] >= numRandom) continue;
                const int i = (int)(((unsigned long long)ws.raw[k] * range) >> 32);
                const uint2 e = edge[curE][i];
                const float4 nMin = vn[e.x], nMax = vn[e.y];
                if (dot4_lr(nMin, nMax) < opt.minAngleCosine) continue;
                const float4 pMin = vx[e.x], pMax = vx[e.y];
                const float4 d = make_float4(pMax.x - pMin.x, pMax.y - pMin.y, pMax.z - pMin.z, pMax.w - pMin.w);
                if (dot4_lr(d, d) > maxEdge2) continue;
                if ((double)fabsf(vc[e.x].w - vc[e.y].w) > 1e-3) continue;
                const int degree = vcount[e.x] + vcount[e.y];
                if (degree > SIMP_MAX_DEGREE) continue;
                float pos[4];
                float error = qef4_solve2(pMin, nMin, pMax, nMax, pos);
                if (error > 0.f) error = 1.f / error;
                const int penalty = max(0, degree - 10);
                error += (float)penalty * (opt.maxError * 0.1f);
                if (error > opt.maxError) continue;
                const unsigned long long pack = ((unsigned long long)__float_as_uint(error) << 32) | (unsigned int)i;
                atomicMin(&best[e.x], pack);
                atomicMin(&best[e.y], pack);
                atomicAdd(&s_cnt, 1);
            }
        }
        __syncthreads();
        if (s_cnt == 0) break;
        // ---- CollapseEdges (ng_mesh_simplify.cpp:287-311): one thread per min vertex ----
        for (int v = tid; v < NV; v += SIMP_BLOCK) {
            const unsigned long long b = best[v];
            if (b == ~0ull) continue;
            const int i = (int)(unsigned int)b;
            const uint2 e = edge[curE][i];
            if ((int)e.x != v || e.x == e.y) continue;
            const unsigned long long bo = best[e.y];
            if (bo == ~0ull || (int)(unsigned int)bo != i) continue;
            float pos[4];
            const float4 nMin = vn[e.x], nMax = vn[e.y];
            qef4_solve2(vx[e.x], nMin, vx[e.y], nMax, pos);
            target[e.y] = (int)e.x;
            vx[e.x] = make_float4(pos[0], pos[1], pos[2], 1.f);
            vn[e.x] = make_float4((nMin.x - nMax.x) * 0.5f, (nMin.y - nMax.y) * 0.5f, (nMin.z - nMax.z) * 0.5f, (nMin.w - nMax.w) * 0.5f);
        }
        __syncthreads();
        // ---- RemoveTriangles (ng_mesh_simplify.cpp:315-360) ----
        for (int i = tid; i < NV; i += SIMP_BLOCK) vcount[i] = 0;
        for (int t = tid; t < NT; t += SIMP_BLOCK) {
            int a = tri[curT][t * 3], b = tri[curT][t * 3 + 1], c = tri[curT][t * 3 + 2];
            const int ta = target[a], tb = target[b], tc = target[c];
            if (ta != -1) a = ta;
            if (tb != -1) b = tb;
            if (tc != -1) c = tc;
            tri[curT][t * 3] = a; tri[curT][t * 3 + 1] = b; tri[curT][t * 3 + 2] = c;
            eflag[t] = !(a == b || a == c || b == c);
        }
        __syncthreads();
        block_scan(eflag, escan, NT, s_warp, &s_run);
        __syncthreads();
        const int newNT = s_run;
        for (int t = tid; t < NT; t += SIMP_BLOCK) {
            if (!eflag[t]) continue;
            const int o = escan[t] * 3;
#pragma unroll
            for (int k = 0; k < 3; k++) { const int idx = tri[curT][t * 3 + k]; tri[curT ^ 1][o + k] = idx; atomicAdd(&vcount[idx], 1); }
        }
        __syncthreads();
        NT = newNT; curT ^= 1;
        // ---- RemoveEdges (ng_mesh_simplify.cpp:364-391) ----
        for (int i = tid; i < NE; i += SIMP_BLOCK) {
            uint2 e = edge[curE][i];
            const int t0 = target[e.x], t1 = target[e.y];
            if (t0 != -1) e.x = (unsigned int)t0;
            if (t1 != -1) e.y = (unsigned int)t1;
            edge[curE][i] = e;
            eflag[i] = e.x != e.y;
        }
        __syncthreads();
        block_scan(eflag, escan, NE, s_warp, &s_run);
        __syncthreads();
        const int newNE = s_run;
        for (int i = tid; i < NE; i += SIMP_BLOCK) if (eflag[i]) edge[curE ^ 1][escan[i]] = edge[curE][i];
        __syncthreads();
        NE = newNE; curE ^= 1;
    }
    __syncthreads();

    // ---- CompactVertices + write back (ng_mesh_simplify.cpp:395-437,520-539) ----
    for (int i = tid; i < NV; i += SIMP_BLOCK) vflag[i] = 0;
    __syncthreads();
    for (int i = tid; i < NT * 3; i += SIMP_BLOCK) vflag[tri[curT][i]] = 1;
    __syncthreads();
    block_scan(vflag, vscan, NV, s_warp, &s_run);
    __syncthreads();
    const int newNV = s_run;
    for (int i = tid; i < NV; i += SIMP_BLOCK) {
        if (!vflag[i]) continue;
        float4 *p = reinterpret_cast<float4 *>(&dstV[vscan[i]]);
        const float4 x = vx[i];
        p[0] = make_float4(x.x + off.x, x.y + off.y, x.z + off.z, x.w + off.w);
        p[1] = vn[i];
        p[2] = vc[i];
    }
    for (int i = tid; i < NT * 3; i += SIMP_BLOCK) dstT[i] = vscan[tri[curT][i]];
    if (tid == 0) results[blockIdx.x] = make_int4(newNV, NT, iterations, NE);
}

// ---------------------------------------------------------------------------------------------
struct SimpState {
    cudaStream_t stream = nullptr;
    void *d_blob = nullptr; size_t blobCap = 0;
    unsigned int *d_raw = nullptr;
};
static SimpState g_simp;
static const char *g_simpError = "";
static int simp_fail(cudaError_t e) { g_simpError = cudaGetErrorString(e); cudaGetLastError(); return LVN_ERR_CUDA; }
#define MCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return simp_fail(e_); } while (0)
static size_t simp_align(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace lvn

using namespace lvn;

extern "C" const char *lvn_mesh_simplify_last_error(void) { return g_simpError; }

extern "C" int lvn_mesh_simplify_batch(int numMeshes, const lvn_simplify_job *jobs, const lvn_simplify_options *options, int numOptions,
                                       lvn_mesh_vertex *vertices, int64_t numVerticesTotal,
                                       lvn_mesh_triangle *triangles, int64_t numTrianglesTotal,
                                       lvn_simplify_result *results)
{
    if (numMeshes < 0 || !options || (numOptions != 1 && numOptions != numMeshes) || (numMeshes > 0 && (!jobs || !results || !vertices || !triangles))) return LVN_ERR_INVALID_VALUE;
    if (numMeshes == 0) return LVN_SUCCESS;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return LVN_ERR_NO_DEVICE; }
    if (!g_simp.stream) MCU(cudaStreamCreateWithFlags(&g_simp.stream, cudaStreamNonBlocking));
    cudaStream_t st = g_simp.stream;
    if (!g_simp.d_raw) {
        // the raw stream of std::mt19937 seeded with 42 (ng_mesh_simplify.cpp:195-196); the engine is standardised
        std::vector<unsigned int> raw(SIMP_RAW);
        std::mt19937 prng;
        prng.seed(42);
        for (int i = 0; i < SIMP_RAW; i++) raw[i] = (unsigned int)prng();
        MCU(cudaMalloc(&g_simp.d_raw, sizeof(unsigned int) * SIMP_RAW));
        MCU(cudaMemcpy(g_simp.d_raw, raw.data(), sizeof(unsigned int) * SIMP_RAW, cudaMemcpyHostToDevice));
    }
    std::vector<SimpJobDev> jd(numMeshes);
    long long edgeTotal = 0;
    long long vtxTotal = 0;
    for (int m = 0; m < numMeshes; m++) {
        const lvn_simplify_job &j = jobs[m];
        if (j.numVertices < 0 || j.numTriangles < 0 || j.vertexOffset < 0 || j.triangleOffset < 0 ||
            (int64_t)j.vertexOffset + j.numVertices > numVerticesTotal || (int64_t)j.triangleOffset + j.numTriangles > numTrianglesTotal)
            return LVN_ERR_INVALID_VALUE;
        // the candidate sample of one iteration must fit the precomputed raw stream
        const lvn_simplify_options &o = options[numOptions == 1 ? 0 : m];
        if (!(o.edgeFraction >= 0.f)) return LVN_ERR_INVALID_VALUE;
        if ((int64_t)j.numTriangles * 3 * (double)o.edgeFraction + 256 > SIMP_RAW) return LVN_ERR_CAPACITY;
        SimpJobDev &d = jd[m];
        d.opt = SimpOptionsDev{o.edgeFraction, o.maxIterations, o.targetPercentage, o.maxError, o.maxEdgeSize, o.minAngleCosine};
        d.vertexOffset = j.vertexOffset; d.numVertices = j.numVertices;
        d.triangleOffset = j.triangleOffset; d.numTriangles = j.numTriangles;
        memcpy(d.offset, j.worldSpaceOffset, sizeof(d.offset));
        d.edgeOff = edgeTotal; d.vtxOff = (int)vtxTotal;
        // eflag / escan also hold the sampler's flags: numRandom + 256 draws of one iteration
        edgeTotal += (long long)((double)j.numTriangles * 3 * std::max(1.0, (double)o.edgeFraction)) + 256;
        vtxTotal += j.numVertices;
    }
    const size_t E = (size_t)std::max<long long>(edgeTotal, 1), V = (size_t)std::max<long long>(vtxTotal, 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += simp_align(bytes); return o; };
    const size_t oJobs = take(sizeof(SimpJobDev) * numMeshes), oInV = take(sizeof(lvn_mesh_vertex) * (size_t)numVerticesTotal),
                 oInT = take(12 * (size_t)numTrianglesTotal), oOutV = take(sizeof(lvn_mesh_vertex) * (size_t)numVerticesTotal),
                 oOutT = take(12 * (size_t)numTrianglesTotal), oRes = take(sizeof(int4) * numMeshes),
                 oVx = take(16 * V), oVn = take(16 * V), oVc = take(16 * V), oT0 = take(4 * E), oT1 = take(4 * E),
                 oE0 = take(8 * E), oE1 = take(8 * E), oBk = take(4 * E), oEf = take(4 * E), oEs = take(4 * E),
                 oVcnt = take(4 * V), oVst = take(4 * V), oVfl = take(4 * V), oBd = take(4 * V), oTg = take(4 * V), oVf = take(4 * V),
                 oVs = take(4 * V), oBest = take(8 * V);
    if (off > g_simp.blobCap) {
        if (g_simp.d_blob) MCU(cudaFree(g_simp.d_blob));
        g_simp.d_blob = nullptr; g_simp.blobCap = 0;
        MCU(cudaMalloc(&g_simp.d_blob, off + off / 4));
        g_simp.blobCap = off + off / 4;
    }
    char *B = (char *)g_simp.d_blob;
    MCU(cudaMemcpyAsync(B + oJobs, jd.data(), sizeof(SimpJobDev) * numMeshes, cudaMemcpyHostToDevice, st));
    MCU(cudaMemcpyAsync(B + oInV, vertices, sizeof(lvn_mesh_vertex) * (size_t)numVerticesTotal, cudaMemcpyHostToDevice, st));
    MCU(cudaMemcpyAsync(B + oInT, triangles, 12 * (size_t)numTrianglesTotal, cudaMemcpyHostToDevice, st));
    SimpScratch ws;
    ws.vx = (float4 *)(B + oVx); ws.vn = (float4 *)(B + oVn); ws.vc = (float4 *)(B + oVc);
    ws.tri[0] = (int *)(B + oT0); ws.tri[1] = (int *)(B + oT1);
    ws.edge[0] = (uint2 *)(B + oE0); ws.edge[1] = (uint2 *)(B + oE1);
    ws.bucket = (unsigned int *)(B + oBk); ws.eflag = (int *)(B + oEf); ws.escan = (int *)(B + oEs);
    ws.vcount = (int *)(B + oVcnt); ws.vstart = (int *)(B + oVst); ws.vfill = (int *)(B + oVfl); ws.boundary = (int *)(B + oBd);
    ws.target = (int *)(B + oTg); ws.vflag = (int *)(B + oVf); ws.vscan = (int *)(B + oVs); ws.best = (unsigned long long *)(B + oBest);
    ws.raw = g_simp.d_raw;
    k_simplify<<<numMeshes, SIMP_BLOCK, 0, st>>>((const SimpJobDev *)(B + oJobs), ws, (const lvn_mesh_vertex *)(B + oInV),
                                                 (const int *)(B + oInT), (lvn_mesh_vertex *)(B + oOutV), (int *)(B + oOutT), (int4 *)(B + oRes));
    MCU(cudaGetLastError());
    std::vector<int4> res(numMeshes);
    MCU(cudaMemcpyAsync(res.data(), B + oRes, sizeof(int4) * numMeshes, cudaMemcpyDeviceToHost, st));
    // the simplified meshes stay in their input slots (a mesh never grows): two copies back
    MCU(cudaMemcpyAsync(vertices, B + oOutV, sizeof(lvn_mesh_vertex) * (size_t)numVerticesTotal, cudaMemcpyDeviceToHost, st));
    MCU(cudaMemcpyAsync(triangles, B + oOutT, 12 * (size_t)numTrianglesTotal, cudaMemcpyDeviceToHost, st));
    MCU(cudaStreamSynchronize(st));
    int rc = LVN_SUCCESS;
    for (int m = 0; m < numMeshes; m++) {
        results[m].numVertices = res[m].x; results[m].numTriangles = res[m].y;
        results[m].iterations = res[m].z; results[m].numEdges = res[m].w;
        if (res[m].z < 0) { rc = LVN_ERR_INVALID_VALUE; g_simpError = "a triangle index lies outside its mesh's vertices"; }
    }
    return rc;
}
