// simplify.cu -- the mesh simplifier that runs on every chunk mesh right after export
// (ngMeshSimplifier, leven/src/ng_mesh_simplify.cpp:441-540; called at clipmap.cpp:449-465,495-501),
// on the GPU: SURVEY.md 8f-2.
//
// The reference's algorithm is already iteration-parallel in spirit (random candidate edges, each
// vertex keeps its cheapest collapse, an edge collapses when both ends chose it); its loops are
// sequential only in how they compact.  Here one thread block owns one mesh (a batch of meshes is
// one launch), the mesh is simplified IN PLACE in its slices of the device arrays, and every
// sequential loop becomes a block-wide pass with the same result:
//   BuildCandidateEdges   std::sort by (max, min) -> counting sort by max vertex, then a rank sort
//                         inside each vertex's bucket (one thread per raw edge, independent loads);
//                         duplicate runs, boundary marks and the stable filter from neighbours in
//                         the sorted list.  (The reference never flushes the last run of its scan:
//                         the greatest edge is neither a candidate nor a boundary mark.  Kept.)
//   FindValidCollapses    the candidate sample is std::mt19937(42) through libstdc++'s
//                         std::uniform_int_distribution (Lemire multiply-shift with rejection),
//                         reproduced from the precomputed raw stream; "first cheapest edge wins" in
//                         ascending edge order = atomicMin on (error bits, edge index)
//   CollapseEdges         one thread per vertex; the winner re-solves its 2-point QEF
//   RemoveTriangles / RemoveEdges / CompactVertices   stable compaction
// Stable compaction = warp ballots into shared-memory mask words, one scan of the words' popcounts,
// destination = word prefix + popc(mask below my lane): two passes over the data and a handful of
// block barriers per compaction, nothing written but the survivors.
// The 4-D QEF solve restates qef_simd.h lane by lane, including what look like slips in
// rotateq_xy (vtav[0][0] written twice, `cc + v`): the reference's results are the bar.
// _mm_rsqrt_ps := 1 / sqrt(x) (arithmetic spec; the x86 estimate differs between CPU vendors).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#include "lvn_internal.h"

namespace lvn {

constexpr int SIMP_BLOCK = 1024;          // one block per mesh; 64 registers per thread
constexpr int SIMP_WARPS = SIMP_BLOCK / 32;
constexpr int SIMP_MAX_DEGREE = 16;      // COLLAPSE_MAX_DEGREE
constexpr int SIMP_SLACK = 256;          // raw draws beyond numRandom that one iteration may consume on rejections
constexpr size_t SIMP_MAX_SMEM = 200 * 1024;

struct SimpJobDev {
    int vertexOffset, numVertices;       // the mesh's slices of the vertex / triangle arrays
    int triangleOffset, numTriangles;
    float offset[4];                     // worldSpaceOffset
    lvn_simplify_options opt;            // this mesh's options (they scale with the node's leaf size, clipmap.cpp:455-462)
    long long edgeOff;                   // scratch slices (element offsets): capacity 3 * numTriangles
    int vtxOff;                          // capacity numVertices
    int result;                          // where this mesh's result goes (jobs are launched largest first)
    int skip;                            // too large for this launch's shared-memory masks: reported, left untouched
};

struct SimpScratch {
    float4 *vx, *vn, *vc;                // working vertices
    int *tri1;                           // staging of the compacted triangle list, 3 ints each
    uint2 *edge[2];                      // (min, max): sort buffers; then [0] the candidate edges, [1] a round's candidate list
    int *boundary;                       // per vertex: boundary mark, later per-draw / per-vertex gather lists
    unsigned long long *best;            // per vertex (error bits << 32 | edge)
    int *rep, *vcount;                   // per vertex: representative, triangle count -- global fallbacks (vtxSmem == 0)
    const unsigned int *raw;             // mt19937(42) outputs
    int wordsE, wordsS;                  // capacities of the shared-memory mask / prefix word arrays
    int vtxSmem;                         // capacity (vertices) of the two shared-memory per-vertex arrays, or 0
    long long *timing;                   // LVN_SIMP_TIMING builds only
};

extern __shared__ unsigned int s_dyn[];

// from the per-warp sums of a round: the sum of the warps before mine, and of all (SIMP_WARPS == 32:
// every warp scans the 32 sums with shuffles instead of reading them all)
static_assert(SIMP_WARPS == 32, "warp_totals scans one warp-sum per lane");
__device__ __forceinline__ void warp_totals(const int *s_warp, int lane, int warp, int &before, int &total)
{
    const int v = s_warp[lane];
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    before = __shfl_sync(0xffffffffu, incl - v, warp);
    total = __shfl_sync(0xffffffffu, incl, 31);
}

// exclusive prefix of the popcounts of s_mask[0 .. nwords) into s_pref; returns the total (same in every thread)
__device__ __forceinline__ int words_prefix(const unsigned int *s_mask, int *s_pref, int nwords, int *s_warp)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = 0;
    for (int base = 0; base < nwords; base += SIMP_BLOCK) {
        const int w = base + tid;
        const int c = w < nwords ? __popc(s_mask[w]) : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int woff, tot;
        warp_totals(s_warp, lane, warp, woff, tot);
        if (w < nwords) s_pref[w] = carry + woff + incl - c;
        carry += tot;
        __syncthreads();
    }
    return carry;
}

// block-wide exclusive scan of n ints in global memory (in -> out, may alias), four per thread per round;
// returns the total (same in every thread)
__device__ __forceinline__ int block_scan(const int *in, int *out, int n, int *s_warp)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += SIMP_BLOCK * 4) {
        const int i = base + tid * 4;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = i + k < n ? in[i + k] : 0;
        const int sum = v[0] + v[1] + v[2] + v[3];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int woff, tot;
        warp_totals(s_warp, lane, warp, woff, tot);
        int run = carry + woff + incl - sum;
#pragma unroll
        for (int k = 0; k < 4; k++) { if (i + k < n) out[i + k] = run; run += v[k]; }
        carry += tot;
        __syncthreads();
    }
    return carry;
}

// pass 1 of a stable compaction: s_mask[i / 32] bit (i % 32) = pred(i), i < n.
// The mesh lives in L2, not in shared memory, so a pass is bound by load latency, not bandwidth:
// every thread takes U elements per round and issues all their loads (`load`, index clamped
// instead of branched) and then all the dependent gathers (`gather`) before the first ballot.
template <int U, class Load, class Gather, class Pred>
__device__ __forceinline__ void ballot_pass(int n, unsigned int *s_mask, Load load, Gather gather, Pred pred)
{
    const int tid = threadIdx.x;
    const int nr = (n + 31) & ~31;
    for (int base = tid; base < nr; base += SIMP_BLOCK * U) {
        decltype(load(0)) a[U];
        decltype(gather(0, a[0])) b[U];
#pragma unroll
        for (int u = 0; u < U; u++) a[u] = load(min(base + u * SIMP_BLOCK, n - 1));
#pragma unroll
        for (int u = 0; u < U; u++) b[u] = gather(min(base + u * SIMP_BLOCK, n - 1), a[u]);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = base + u * SIMP_BLOCK;
            if (i < nr) {   // warp-uniform
                const unsigned int m = __ballot_sync(0xffffffffu, i < n && pred(i, a[u], b[u]));
                if ((tid & 31) == 0) s_mask[i >> 5] = m;
            }
        }
    }
}
// pass 2: where element i goes, or -1
__device__ __forceinline__ int compact_slot(const unsigned int *s_mask, const int *s_pref, int i)
{
    const unsigned int m = s_mask[i >> 5], bit = 1u << (i & 31);
    return (m & bit) ? s_pref[i >> 5] + __popc(m & (bit - 1u)) : -1;
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

// the i-th set bit over the mask words (i < total): the word by binary search over the exclusive
// prefix, the bit by __fns
__device__ __forceinline__ int select_bit(const unsigned int *s_mask, const int *s_pref, int nwords, int i)
{
    int lo = 0, hi = nwords - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_pref[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return (lo << 5) + (int)__fns(s_mask[lo], 0, i - s_pref[lo] + 1);
}

// LVN_SIMP_TIMING: thread 0 of every block accumulates clock64 per phase into ws.timing[block][16]
#ifdef LVN_SIMP_TIMING
#define PHASE(k) do { __syncthreads(); if (threadIdx.x == 0) { const long long now_ = clock64(); ws.timing[blockIdx.x * 16 + ph_] += now_ - t_; t_ = now_; ph_ = (k); } } while (0)
#else
#define PHASE(k) do { } while (0)
#endif

// One block simplifies one mesh.  The reference rewrites and compacts its triangle and edge lists
// after every round of collapses; here both lists stay as built and a per-vertex representative
// (rep[v]: the vertex v has been merged into) stands for the rewriting: element k of a list is
// alive while its ends' representatives differ, and "the i-th edge of the compacted list" -- what
// the random candidate sample indexes -- is the i-th alive edge, found by a select over the
// alive mask.  The lists are compacted once, at the end.
__global__ void __launch_bounds__(SIMP_BLOCK)
k_simplify(const SimpJobDev *__restrict__ jobs, SimpScratch ws, lvn_mesh_vertex *V, int *T, int4 *__restrict__ results)
{
    __shared__ int s_warp[SIMP_WARPS], s_any, s_last, s_bad, s_count;
    // dynamic shared memory: [wordsE] alive-edge mask, [wordsE] its prefix, [wordsS] scratch mask, [wordsS] its
    // prefix, [wordsS] + [wordsS] the same for the accepted draws, then the two per-vertex arrays when they fit
    unsigned int *s_maskE = s_dyn;
    int *s_prefE = reinterpret_cast<int *>(s_dyn + ws.wordsE);
    unsigned int *s_mask = s_dyn + 2 * ws.wordsE;
    int *s_pref = reinterpret_cast<int *>(s_mask + ws.wordsS);
    unsigned int *s_maskA = s_mask + 2 * ws.wordsS;
    int *s_prefA = reinterpret_cast<int *>(s_maskA + ws.wordsS);
    const SimpJobDev job = jobs[blockIdx.x];
    const lvn_simplify_options opt = job.opt;
    const int tid = threadIdx.x, NV = job.numVertices, NT0 = job.numTriangles;
    lvn_mesh_vertex *meshV = V + job.vertexOffset;
    int *tri = T + (size_t)job.triangleOffset * 3;

    if (job.skip || NT0 < 100 || NV < 100) {   // ng_mesh_simplify.cpp:446-449: too small, returned untouched
        if (tid == 0) results[job.result] = make_int4(NV, NT0, job.skip ? -2 : 0, 0);
        return;
    }
    // The per-vertex arrays every pass gathers from live in shared memory when the launch's largest
    // mesh allows (ws.vtxSmem): a pass over triangles or edges then has one global load per element
    // and the count updates are shared-memory atomics.  Otherwise the same code runs on global fallbacks.
    //   rep     the vertex this one has been merged into (itself while it lives)
    //   vcount  triangles per vertex (FindValidCollapses' degree); between the sample and the next
    //           triangle pass the same array holds the round's collapse targets
    int *rep = ws.vtxSmem ? s_prefA + ws.wordsS : ws.rep + job.vtxOff;
    int *vcount = ws.vtxSmem ? rep + ws.vtxSmem : ws.vcount + job.vtxOff;
    int *target = vcount;
    float4 *vx = ws.vx + job.vtxOff, *vn = ws.vn + job.vtxOff, *vc = ws.vc + job.vtxOff;
    uint2 *edgeA = ws.edge[0] + job.edgeOff, *edgeB = ws.edge[1] + job.edgeOff;
    int *boundary = ws.boundary + job.vtxOff;        // BuildCandidateEdges; afterwards the list of collapsing vertices
    unsigned long long *best = ws.best + job.vtxOff;
    const float4 off = make_float4(job.offset[0], job.offset[1], job.offset[2], job.offset[3]);
    const int NE0 = NT0 * 3;
#ifdef LVN_SIMP_TIMING
    long long t_ = clock64();
    int ph_ = 0;
#endif

    // ---- copy in; v.xyz -= worldSpaceOffset (ng_mesh_simplify.cpp:451-463) ----
    if (tid == 0) { s_bad = 0; s_last = -1; }
    for (int i = tid; i < NV; i += SIMP_BLOCK) {
        const float4 *p = reinterpret_cast<const float4 *>(&meshV[i]);
        const float4 x = p[0];
        vx[i] = make_float4(x.x - off.x, x.y - off.y, x.z - off.z, x.w - off.w);
        vn[i] = p[1];
        vc[i] = p[2];
        vcount[i] = 0; boundary[i] = 0;
    }
    __syncthreads();

    // ---- BuildCandidateEdges (ng_mesh_simplify.cpp:122-177) ----
    // counting sort of the raw edges by max vertex ...
    // An index outside the mesh's vertices would be a wild write (the reference would crash): such a
    // mesh is left untouched and reported with iterations = -1.
    int lmax = -1;
#pragma unroll 4
    for (int t = tid; t < NT0; t += SIMP_BLOCK) {      // a triangle's raw edges: (0,1), (1,2), (0,2) (ng_mesh_simplify.cpp:127-133)
        const int a = tri[t * 3], b = tri[t * 3 + 1], c = tri[t * 3 + 2];
        if ((unsigned int)a >= (unsigned int)NV || (unsigned int)b >= (unsigned int)NV || (unsigned int)c >= (unsigned int)NV) { s_bad = 1; continue; }
        atomicAdd(&vcount[max(a, b)], 1); atomicAdd(&vcount[max(b, c)], 1); atomicAdd(&vcount[max(a, c)], 1);
        lmax = max(lmax, max(a, max(b, c)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if ((tid & 31) == 0) atomicMax(&s_last, lmax);     // the greatest max vertex: owner of the list's last run
    __syncthreads();
    if (s_bad) {
        if (tid == 0) results[job.result] = make_int4(NV, NT0, -1, 0);
        return;
    }
    int *vstart = rep;                                 // (not in use yet)
    block_scan(vcount, vstart, NV, s_warp);
#pragma unroll 4
    for (int t = tid; t < NT0; t += SIMP_BLOCK) {      // vstart[v] ends as the bucket's end
        const int a = tri[t * 3], b = tri[t * 3 + 1], c = tri[t * 3 + 2];
        edgeB[atomicAdd(&vstart[max(a, b)], 1)] = make_uint2((unsigned int)min(a, b), (unsigned int)max(a, b));
        edgeB[atomicAdd(&vstart[max(b, c)], 1)] = make_uint2((unsigned int)min(b, c), (unsigned int)max(b, c));
        edgeB[atomicAdd(&vstart[max(a, c)], 1)] = make_uint2((unsigned int)min(a, c), (unsigned int)max(a, c));
    }
    __syncthreads();
    PHASE(1);
    // ... then a rank sort by min inside each bucket (equal keys are interchangeable): edgeA = std::sort's order
#pragma unroll 2
    for (int p = tid; p < NE0; p += SIMP_BLOCK) {
        const uint2 e = edgeB[p];
        const int n = vcount[e.y], st = vstart[e.y] - n;
        int rank = 0;
        for (int q = 0; q < n; q++) {
            const unsigned int k = edgeB[st + q].x;
            rank += (k < e.x) || (k == e.x && st + q < p);
        }
        edgeA[st + rank] = e;
    }
    __syncthreads();
    PHASE(2);
    // runs of equal edges: a run of one marks both ends as boundary, a longer run is one filtered edge;
    // the last run of the list is never flushed by the reference's scan
    const int lastV = s_last;
    const unsigned int lastKey = edgeA[NE0 - 1].x;
    ballot_pass<4>(NE0, s_maskE,
        [&](int p) { return edgeA[p]; },
        [&](int p, uint2) { return make_uint4(edgeA[max(p - 1, 0)].x, edgeA[max(p - 1, 0)].y, edgeA[min(p + 1, NE0 - 1)].x, edgeA[min(p + 1, NE0 - 1)].y); },
        [&](int p, uint2 e, uint4 nb) {
            const bool start = p == 0 || nb.x != e.x || nb.y != e.y;
            const bool more = p + 1 < NE0 && nb.z == e.x && nb.w == e.y;
            const bool lastRun = (int)e.y == lastV && e.x == lastKey;
            if (start && !more && !lastRun) { boundary[e.x] = 1; boundary[e.y] = 1; }
            return start && more && !lastRun;
        });
    __syncthreads();
    const int numFiltered = words_prefix(s_maskE, s_prefE, (NE0 + 31) >> 5, s_warp);
#pragma unroll 4
    for (int p = tid; p < NE0; p += SIMP_BLOCK) {
        const uint2 e = edgeA[p];
        const int o = compact_slot(s_maskE, s_prefE, p);
        if (o >= 0) edgeB[o] = e;
    }
    __syncthreads();
    ballot_pass<4>(numFiltered, s_maskE,
        [&](int i) { return edgeB[i]; },
        [&](int, uint2 e) { return make_int2(boundary[e.x], boundary[e.y]); },
        [&](int, uint2, int2 b) { return !b.x && !b.y; });
    __syncthreads();
    const int NEinit = words_prefix(s_maskE, s_prefE, (numFiltered + 31) >> 5, s_warp);
#pragma unroll 4
    for (int i = tid; i < numFiltered; i += SIMP_BLOCK) {
        const uint2 e = edgeB[i];
        const int o = compact_slot(s_maskE, s_prefE, i);
        if (o >= 0) edgeA[o] = e;
    }
    const uint2 *edges = edgeA;                        // the candidate edges, as built; never rewritten
    const int wordsE = (NEinit + 31) >> 5;
    __syncthreads();
    PHASE(3);

    // alive edges: mask + prefix; returns their number (RemoveEdges, ng_mesh_simplify.cpp:364-391)
    auto alive_edges = [&]() {
        ballot_pass<4>(NEinit, s_maskE,
            [&](int i) { return edges[i]; },
            [&](int, uint2 e) { return make_int2(rep[e.x], rep[e.y]); },
            [&](int, uint2, int2 r) { return r.x != r.y; });
        __syncthreads();
        return words_prefix(s_maskE, s_prefE, wordsE, s_warp);
    };
    // alive triangles: their number, and the triangle count of every vertex
    // (RemoveTriangles, ng_mesh_simplify.cpp:315-360; vertexTriangleCounts :478-489)
    auto alive_triangles = [&]() {
        for (int i = tid; i < NV; i += SIMP_BLOCK) vcount[i] = 0;
        if (tid == 0) s_count = 0;
        __syncthreads();
        int cnt = 0;
        for (int base = tid; base < NT0; base += SIMP_BLOCK * 4) {
            int3 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const int t = min(base + u * SIMP_BLOCK, NT0 - 1); v[u] = make_int3(tri[t * 3], tri[t * 3 + 1], tri[t * 3 + 2]); }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int a = rep[v[u].x], b = rep[v[u].y], c = rep[v[u].z];
                if (base + u * SIMP_BLOCK < NT0 && !(a == b || a == c || b == c)) {
                    cnt++;
                    atomicAdd(&vcount[a], 1); atomicAdd(&vcount[b], 1); atomicAdd(&vcount[c], 1);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((tid & 31) == 0 && cnt) atomicAdd(&s_count, cnt);
        __syncthreads();
        return s_count;
    };

    for (int i = tid; i < NV; i += SIMP_BLOCK) rep[i] = i;
    __syncthreads();
    int NT = alive_triangles();
    int NE = alive_edges();
    const int targetTriangleCount = (int)((float)NT0 * opt.targetPercentage);
    const float maxEdge2 = opt.maxEdgeSize * opt.maxEdgeSize;
    int *list = boundary;
    int *cand = reinterpret_cast<int *>(edgeB);
    int iterations = 0;
    while (NT > targetTriangleCount && iterations++ < opt.maxIterations) {
        PHASE(4);
        for (int i = tid; i < NV; i += SIMP_BLOCK) best[i] = ~0ull;
        if (tid == 0) s_any = 0;
        // ---- FindValidCollapses (ng_mesh_simplify.cpp:181-283) ----
        const int numRandom = (int)((float)NE * opt.edgeFraction);
        if (numRandom > 0) {
            // std::uniform_int_distribution<int>(0, NE - 1) over std::mt19937(42), libstdc++:
            // product = raw * range; reject while (uint32)product < (2^32 - range) % range
            const unsigned int range = (unsigned int)NE, threshold = (0u - range) % range;
            const int K = numRandom + SIMP_SLACK;
            int *park = cand + K;
            ballot_pass<2>(K, s_maskA,
                [&](int k) { return ws.raw[k]; },
                [&](int, unsigned int) { return 0; },
                [&](int, unsigned int r, int) { return (unsigned int)((unsigned long long)r * range) >= threshold; });
            __syncthreads();
            words_prefix(s_maskA, s_prefA, (K + 31) >> 5, s_warp);
            // the cheap tests first, on every draw (its place among the accepted ones from the first pair of
            // word arrays); the survivors are gathered so that the QEF solves below run on full warps
            ballot_pass<2>(K, s_mask,
                [&](int k) {
                    const int rank = compact_slot(s_maskA, s_prefA, k);
                    const int i = (int)(((unsigned long long)ws.raw[k] * range) >> 32);
                    const int j = select_bit(s_maskE, s_prefE, wordsE, i);        // the i-th alive edge
                    const uint2 e = edges[j];
                    return make_int4(rep[e.x], rep[e.y], j, rank >= 0 && rank < numRandom);
                },
                [&](int, int4 a) {
                    const float4 nMin = vn[a.x], nMax = vn[a.y];
                    const float4 pMin = vx[a.x], pMax = vx[a.y];
                    const float mMin = vc[a.x].w, mMax = vc[a.y].w;
                    const int degree = vcount[a.x] + vcount[a.y];
                    const float4 d = make_float4(pMax.x - pMin.x, pMax.y - pMin.y, pMax.z - pMin.z, pMax.w - pMin.w);
                    return (!(dot4_lr(nMin, nMax) < opt.minAngleCosine) && !(dot4_lr(d, d) > maxEdge2) &&
                            !((double)fabsf(mMin - mMax) > 1e-3) && !(degree > SIMP_MAX_DEGREE)) ? a.z : -1;
                },
                [&](int k, int4 a, int j) {
                    const bool ok = a.w && j >= 0;
                    if (ok) park[k] = j;          // (parked per draw; gathered below)
                    return ok;
                });
            __syncthreads();
            const int numCand = words_prefix(s_mask, s_pref, (K + 31) >> 5, s_warp);
#pragma unroll 2
            for (int k = tid; k < K; k += SIMP_BLOCK) {
                const int o = compact_slot(s_mask, s_pref, k);
                if (o >= 0) cand[o] = park[k];
            }
            __syncthreads();
            for (int c = tid; c < numCand; c += SIMP_BLOCK) {
                const int j = cand[c];
                const uint2 e0 = edges[j];
                const int x = rep[e0.x], y = rep[e0.y];
                const float4 nMin = vn[x], nMax = vn[y];
                const float4 pMin = vx[x], pMax = vx[y];
                const int degree = vcount[x] + vcount[y];
                float pos[4];
                float error = qef4_solve2(pMin, nMin, pMax, nMax, pos);
                if (error > 0.f) error = 1.f / error;
                const int penalty = max(0, degree - 10);
                error += (float)penalty * (opt.maxError * 0.1f);
                if (error > opt.maxError) continue;
                // "the first cheapest edge wins", edges visited in ascending order: the alive edges keep their order
                const unsigned long long pack = ((unsigned long long)__float_as_uint(error) << 32) | (unsigned int)j;
                atomicMin(&best[x], pack);
                atomicMin(&best[y], pack);
                s_any = 1;
            }
        }
        __syncthreads();
        if (s_any == 0) break;
        PHASE(5);
        // ---- CollapseEdges (ng_mesh_simplify.cpp:287-311): an edge collapses when both its ends chose it;
        //      the collapsing (min) vertices are gathered first, then each re-solves its 2-point QEF ----
        ballot_pass<2>(NV, s_mask,
            [&](int v) { return best[v]; },
            [&](int, unsigned long long b) {
                const uint2 e0 = edges[b == ~0ull ? 0 : (int)(unsigned int)b];
                const int x = rep[e0.x], y = rep[e0.y];
                const unsigned long long bo = best[y];
                return make_uint4((unsigned int)x, (unsigned int)y, (unsigned int)bo, (unsigned int)(bo >> 32));
            },
            [&](int v, unsigned long long b, uint4 g) {
                const unsigned long long bo = ((unsigned long long)g.w << 32) | g.z;
                return b != ~0ull && (int)g.x == v && g.x != g.y && bo != ~0ull && (unsigned int)bo == (unsigned int)b;
            });
        for (int i = tid; i < NV; i += SIMP_BLOCK) target[i] = -1;       // (the degrees are no longer needed)
        __syncthreads();
        const int numWin = words_prefix(s_mask, s_pref, (NV + 31) >> 5, s_warp);
        for (int v = tid; v < NV; v += SIMP_BLOCK) {
            const int o = compact_slot(s_mask, s_pref, v);
            if (o >= 0) list[o] = v;
        }
        __syncthreads();
        for (int w = tid; w < numWin; w += SIMP_BLOCK) {
            const int v = list[w];
            const uint2 e0 = edges[(int)(unsigned int)best[v]];
            const int x = rep[e0.x], y = rep[e0.y];
            float pos[4];
            const float4 nMin = vn[x], nMax = vn[y];
            qef4_solve2(vx[x], nMin, vx[y], nMax, pos);
            target[y] = x;
            vx[x] = make_float4(pos[0], pos[1], pos[2], 1.f);
            vn[x] = make_float4((nMin.x - nMax.x) * 0.5f, (nMin.y - nMax.y) * 0.5f, (nMin.z - nMax.z) * 0.5f, (nMin.w - nMax.w) * 0.5f);
        }
        __syncthreads();
        for (int i = tid; i < NV; i += SIMP_BLOCK) { const int t = target[rep[i]]; if (t != -1) rep[i] = t; }
        __syncthreads();
        PHASE(6);
        NT = alive_triangles();       // (zeroes the shared target / count array first)
        PHASE(7);
        NE = alive_edges();
    }
    __syncthreads();
    PHASE(8);

    // ---- the one compaction: alive triangles in order, the vertices they use in order
    //      (RemoveTriangles' list, CompactVertices + write back, ng_mesh_simplify.cpp:395-437,520-539) ----
    int *used = vcount;
    for (int i = tid; i < NV; i += SIMP_BLOCK) used[i] = 0;
    __syncthreads();
    ballot_pass<4>(NT0, s_maskE,
        [&](int t) { return make_int3(tri[t * 3], tri[t * 3 + 1], tri[t * 3 + 2]); },
        [&](int, int3 v) { return make_int3(rep[v.x], rep[v.y], rep[v.z]); },
        [&](int t, int3, int3 r) {
            const bool alive = !(r.x == r.y || r.x == r.z || r.y == r.z);
            if (alive) { used[r.x] = 1; used[r.y] = 1; used[r.z] = 1; }
            return alive;
        });
    __syncthreads();
    words_prefix(s_maskE, s_prefE, (NT0 + 31) >> 5, s_warp);
    ballot_pass<4>(NV, s_mask, [&](int i) { return used[i]; }, [&](int, int) { return 0; }, [&](int, int u, int) { return u != 0; });
    __syncthreads();
    const int newNV = words_prefix(s_mask, s_pref, (NV + 31) >> 5, s_warp);
    for (int i = tid; i < NV; i += SIMP_BLOCK) {
        const int o = compact_slot(s_mask, s_pref, i);
        if (o < 0) continue;
        float4 *p = reinterpret_cast<float4 *>(&meshV[o]);
        const float4 x = vx[i];
        p[0] = make_float4(x.x + off.x, x.y + off.y, x.z + off.z, x.w + off.w);
        p[1] = vn[i];
        p[2] = vc[i];
    }
    // a triangle moves towards the front of its own list: staged through scratch
    int *stage = ws.tri1 + job.edgeOff;
#pragma unroll 2
    for (int t = tid; t < NT0; t += SIMP_BLOCK) {
        const int o = compact_slot(s_maskE, s_prefE, t);
        if (o < 0) continue;
#pragma unroll
        for (int k = 0; k < 3; k++) stage[o * 3 + k] = compact_slot(s_mask, s_pref, rep[tri[t * 3 + k]]);
    }
    __syncthreads();
#pragma unroll 4
    for (int i = tid; i < NT * 3; i += SIMP_BLOCK) tri[i] = stage[i];
    PHASE(9);
    if (tid == 0) results[job.result] = make_int4(newNV, NT, iterations, NE);
}

// The simplified meshes, left at the start of their slices, gathered into dense arrays in the order
// of the list: one block per listed mesh; a mesh's place = *baseFrom (the end of an earlier group's
// region; null = 0) + the counts of the listed meshes before it.  results / packedOffsets are
// indexed by the mesh's own number (job.result); *totals = where this group's region ends.
__global__ void __launch_bounds__(256)
k_pack_meshes(const SimpJobDev *__restrict__ jobsByMesh, const int4 *__restrict__ results, int numMeshes,
              const lvn_mesh_vertex *__restrict__ V, const int *__restrict__ T,
              lvn_mesh_vertex *__restrict__ outV, int *__restrict__ outT, int2 *__restrict__ packedOffsets, int2 *__restrict__ totals,
              float4 *__restrict__ outP, float physicsScale, const int2 *__restrict__ baseFrom)
{
    __shared__ int s_v[8], s_t[8];
    const int m = blockIdx.x, tid = threadIdx.x;
    int sv = 0, st = 0;
    for (int i = tid; i < m; i += 256) { const int4 r = results[jobsByMesh[i].result]; sv += r.x; st += r.y; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sv += __shfl_xor_sync(0xffffffffu, sv, o); st += __shfl_xor_sync(0xffffffffu, st, o); }
    if ((tid & 31) == 0) { s_v[tid >> 5] = sv; s_t[tid >> 5] = st; }
    __syncthreads();
    int baseV = baseFrom ? baseFrom->x : 0, baseT = baseFrom ? baseFrom->y : 0;
#pragma unroll
    for (int q = 0; q < 8; q++) { baseV += s_v[q]; baseT += s_t[q]; }
    const SimpJobDev job = jobsByMesh[m];
    const int4 r = results[job.result];
    const float4 *src = reinterpret_cast<const float4 *>(V + job.vertexOffset);
    float4 *dst = reinterpret_cast<float4 *>(outV + baseV);
    if (outV) for (int i = tid; i < r.x * 3; i += 256) dst[i] = src[i];
    if (outP) {   // AddMeshToWorldImpl, physics.cpp:562-566: Scale_WorldToPhysics(vertex.xyz - vec4(origin, 0)), one vec4 per vertex
        const float4 o = make_float4(job.offset[0], job.offset[1], job.offset[2], 0.f);
        for (int i = tid; i < r.x; i += 256) {
            const float4 x = src[i * 3];
            outP[baseV + i] = make_float4((x.x - o.x) * physicsScale, (x.y - o.y) * physicsScale, (x.z - o.z) * physicsScale, (x.w - o.w) * physicsScale);
        }
    }
    const int *ts = T + (size_t)job.triangleOffset * 3;
    int *td = outT + (size_t)baseT * 3;
    for (int i = tid; i < r.y * 3; i += 256) td[i] = ts[i];
    if (tid == 0) {
        packedOffsets[job.result] = make_int2(baseV, baseT);
        if (m == numMeshes - 1) *totals = make_int2(baseV + r.x, baseT + r.y);
    }
}

// ---------------------------------------------------------------------------------------------
struct SimpState {
    cudaStream_t stream = nullptr;
    void *d_blob = nullptr; size_t blobCap = 0;       // scratch of a launch
    void *d_io = nullptr; size_t ioCap = 0;           // host path: the caller's arrays on the device
    unsigned int *d_raw = nullptr; int rawCount = 0;
    size_t smemSet = 0;
};
static SimpState g_simp;
static const char *g_simpError = "";
static int simp_fail(cudaError_t e) { g_simpError = cudaGetErrorString(e); cudaGetLastError(); return LVN_ERR_CUDA; }
#define LV(call) do { int r_ = (call); if (r_ < 0) return r_; } while (0)
#define MCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return simp_fail(e_); } while (0)
static size_t simp_align(size_t v) { return (v + 255) & ~(size_t)255; }

const char *simplify_last_error() { return g_simpError; }

// the raw stream of std::mt19937 seeded with 42 (ng_mesh_simplify.cpp:195-196); the engine is standardised
static int ensure_raw(int count)
{
    if (count <= g_simp.rawCount) return LVN_SUCCESS;
    int cap = 1 << 16;
    while (cap < count) cap <<= 1;
    std::vector<unsigned int> raw(cap);
    std::mt19937 prng;
    prng.seed(42);
    for (int i = 0; i < cap; i++) raw[i] = (unsigned int)prng();
    if (g_simp.d_raw) { MCU(cudaDeviceSynchronize()); MCU(cudaFree(g_simp.d_raw)); g_simp.d_raw = nullptr; g_simp.rawCount = 0; }
    MCU(cudaMalloc(&g_simp.d_raw, sizeof(unsigned int) * cap));
    MCU(cudaMemcpy(g_simp.d_raw, raw.data(), sizeof(unsigned int) * cap, cudaMemcpyHostToDevice));
    g_simp.rawCount = cap;
    return LVN_SUCCESS;
}

static int ensure_buffer(void **p, size_t *cap, size_t bytes)
{
    if (bytes <= *cap) return LVN_SUCCESS;
    if (*p) { MCU(cudaDeviceSynchronize()); MCU(cudaFree(*p)); *p = nullptr; *cap = 0; }
    MCU(cudaMalloc(p, bytes + bytes / 4));
    *cap = bytes + bytes / 4;
    return LVN_SUCCESS;
}

// Simplify n meshes in place in their slices of d_V / d_T, asynchronously on `st`.
// d_results[m] (device, mesh order) = (vertices, triangles, iterations, candidate edges left).
// With d_packT the simplified meshes are also gathered densely, in mesh order (vertices as MeshVertex
// into d_packV and / or in the physics engine's format into d_packP, either may be null):
// d_packOffsets[m] = (first vertex, first triangle), *d_packTotals = the totals.
//
// With `split` (and packing) the launch is cut in two.  A launch lasts as long as its largest mesh --
// one block per mesh -- while most blocks are done in half that time.  The meshes with more than
// half the triangles of the largest stay on `st` (the late group); the others run beside them on
// split->streamB, are packed there first, at the front of the packed arrays, and
// *split->d_earlyTotals says where their region ends: the caller can ship it while the late group is
// still being simplified.  The late group is packed behind it on `st`.  Packed order is then: early
// meshes in mesh order, late meshes in mesh order; d_packOffsets[m] addresses mesh m either way.
int simplify_device(int n, const SimplifyMesh *meshes, lvn_mesh_vertex *d_V, int *d_T, int4 *d_results,
                    lvn_mesh_vertex *d_packV, int *d_packT, int2 *d_packOffsets, int2 *d_packTotals, cudaStream_t st,
                    float4 *d_packP, float physicsScale, SimplifySplit *split)
{
    if (split) split->numEarly = 0;
    if (n <= 0) return LVN_SUCCESS;
    std::vector<SimpJobDev> jd(n);
    long long edgeTotal = 0, vtxTotal = 0;
    int maxEdges = 0, maxOther = 0, maxDraws = 0, maxVerts = 0;
    for (int m = 0; m < n; m++) {
        const SimplifyMesh &j = meshes[m];
        SimpJobDev &d = jd[m];
        if (!(j.opt.edgeFraction >= 0.f)) return LVN_ERR_INVALID_VALUE;
        d.vertexOffset = j.vertexOffset; d.numVertices = j.numVertices;
        d.triangleOffset = j.triangleOffset; d.numTriangles = j.numTriangles;
        memcpy(d.offset, j.offset, sizeof(d.offset));
        d.opt = j.opt;
        d.edgeOff = edgeTotal; d.vtxOff = (int)vtxTotal; d.result = m; d.skip = 0;
        if (j.numTriangles < 100 || j.numVertices < 100) continue;    // passes through: no scratch
        // one iteration consumes at most numRandom + SIMP_SLACK raw draws
        const double draws = (double)j.numTriangles * 3.0 * (double)j.opt.edgeFraction + SIMP_SLACK;
        const double other = std::max(draws, (double)j.numVertices);
        // shared memory: mask + prefix words over the raw edges, and over the draws / the vertices
        if (((double)j.numTriangles * 3.0 / 32.0 + 2.0 * other / 32.0 + 192.0) * 8.0 > (double)SIMP_MAX_SMEM) { d.skip = 1; continue; }   // iterations = -2
        maxEdges = std::max(maxEdges, j.numTriangles * 3);
        maxOther = std::max(maxOther, (int)other);
        maxDraws = std::max(maxDraws, (int)draws);
        maxVerts = std::max(maxVerts, j.numVertices);
        // the edge buffers hold 3 * numTriangles (min, max) pairs; the second one later a round's draws (two ints each)
        edgeTotal += (std::max((long long)j.numTriangles * 3, (long long)draws + 1) + 3) & ~3ll;
        vtxTotal += (j.numVertices + 3) & ~3;
        if (vtxTotal > 0x7fffffffll) return LVN_ERR_CAPACITY;
    }
    LV(ensure_raw(maxDraws));
    const int wordsE = ((maxEdges + 31) / 32 + 31) & ~31, wordsS = ((maxOther + 31) / 32 + 31) & ~31;
    int vtxSmem = (maxVerts + 31) & ~31;
    if (((size_t)wordsE + 2 * (size_t)wordsS + vtxSmem) * 8 > SIMP_MAX_SMEM) vtxSmem = 0;     // per-vertex arrays fall back to global memory
    if (getenv("LVN_SIMP_FORCE_GLOBAL")) vtxSmem = 0;      // test switch: run the fallback on meshes that would fit
    const size_t smem = ((size_t)wordsE + 2 * (size_t)wordsS + vtxSmem) * 8;
    if (smem > 48 * 1024 && smem > g_simp.smemSet) {
        MCU(cudaFuncSetAttribute(k_simplify, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SIMP_MAX_SMEM));
        g_simp.smemSet = SIMP_MAX_SMEM;
    }
    // launch order: largest meshes first (a block per mesh; the tail of the launch is small meshes)
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return jd[a].numTriangles > jd[b].numTriangles; });
    std::vector<SimpJobDev> launch(n);
    for (int i = 0; i < n; i++) launch[i] = jd[order[i]];

    const size_t E = (size_t)std::max<long long>(edgeTotal, 4), NVt = (size_t)std::max<long long>(vtxTotal, 4);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += simp_align(bytes); return o; };
    const size_t oLaunch = take(sizeof(SimpJobDev) * n), oJobs = take(sizeof(SimpJobDev) * n),
                 oVx = take(16 * NVt), oVn = take(16 * NVt), oVc = take(16 * NVt), oT1 = take(4 * E),
                 oE0 = take(8 * E), oE1 = take(8 * E), oVcnt = take(4 * NVt), oBd = take(4 * NVt), oTg = take(4 * NVt),
                 oBest = take(8 * NVt);
    LV(ensure_buffer(&g_simp.d_blob, &g_simp.blobCap, off));
    char *B = (char *)g_simp.d_blob;
    MCU(cudaMemcpyAsync(B + oLaunch, launch.data(), sizeof(SimpJobDev) * n, cudaMemcpyHostToDevice, st));
    SimpScratch ws;
    ws.vx = (float4 *)(B + oVx); ws.vn = (float4 *)(B + oVn); ws.vc = (float4 *)(B + oVc);
    ws.tri1 = (int *)(B + oT1);
    ws.edge[0] = (uint2 *)(B + oE0); ws.edge[1] = (uint2 *)(B + oE1);
    ws.vcount = (int *)(B + oVcnt); ws.boundary = (int *)(B + oBd); ws.rep = (int *)(B + oTg);
    ws.best = (unsigned long long *)(B + oBest);
    ws.raw = g_simp.d_raw;
    ws.wordsE = wordsE; ws.wordsS = wordsS;
    ws.vtxSmem = vtxSmem;
    ws.timing = nullptr;
#ifdef LVN_SIMP_TIMING
    static long long *d_timing = nullptr;
    if (!d_timing) MCU(cudaMalloc(&d_timing, sizeof(long long) * 16 * 65536));
    MCU(cudaMemsetAsync(d_timing, 0, sizeof(long long) * 16 * n, st));
    ws.timing = d_timing;
#endif
    // the late group = the front of the launch list (it is sorted by size)
    int nLate = n;
    if (split && d_packT && n >= 16 && !getenv("LVN_SIMP_NO_SPLIT")) {
        nLate = 0;
        static const int pct = getenv("LVN_SIMP_SPLIT_PCT") ? atoi(getenv("LVN_SIMP_SPLIT_PCT")) : 50;   // experiment switch
        while (nLate < n && 100 * (long long)launch[nLate].numTriangles > (long long)pct * launch[0].numTriangles) nLate++;
        if (nLate == 0 || nLate == n) nLate = n;
    }
    if (nLate < n) {
        // pack lists in mesh order: the early group, then the late group
        std::vector<char> late(n, 0);
        for (int i = 0; i < nLate; i++) late[launch[i].result] = 1;
        std::vector<SimpJobDev> lists;
        lists.reserve(n);
        for (int m = 0; m < n; m++) if (!late[m]) lists.push_back(jd[m]);
        for (int m = 0; m < n; m++) if (late[m]) lists.push_back(jd[m]);
        const int nEarly = n - nLate;
        MCU(cudaMemcpyAsync(B + oJobs, lists.data(), sizeof(SimpJobDev) * n, cudaMemcpyHostToDevice, st));
        MCU(cudaEventRecord(split->evFork, st));
        MCU(cudaStreamWaitEvent(split->streamB, split->evFork, 0));
        const SimpJobDev *dl = (const SimpJobDev *)(B + oLaunch), *dj = (const SimpJobDev *)(B + oJobs);
        k_simplify<<<nLate, SIMP_BLOCK, smem, st>>>(dl, ws, d_V, d_T, d_results);
        k_simplify<<<nEarly, SIMP_BLOCK, smem, split->streamB>>>(dl + nLate, ws, d_V, d_T, d_results);
        k_pack_meshes<<<nEarly, 256, 0, split->streamB>>>(dj, d_results, nEarly, d_V, d_T, d_packV, d_packT, d_packOffsets, split->d_earlyTotals,
                                                          d_packP, physicsScale, nullptr);
        MCU(cudaEventRecord(split->evEarly, split->streamB));
        MCU(cudaStreamWaitEvent(st, split->evEarly, 0));
        k_pack_meshes<<<nLate, 256, 0, st>>>(dj + nEarly, d_results, nLate, d_V, d_T, d_packV, d_packT, d_packOffsets, d_packTotals,
                                             d_packP, physicsScale, split->d_earlyTotals);
        MCU(cudaGetLastError());
        split->numEarly = nEarly;
        return LVN_SUCCESS;
    }
    k_simplify<<<n, SIMP_BLOCK, smem, st>>>((const SimpJobDev *)(B + oLaunch), ws, d_V, d_T, d_results);
    MCU(cudaGetLastError());
#ifdef LVN_SIMP_TIMING
    {
        std::vector<long long> t(16 * (size_t)n);
        MCU(cudaMemcpyAsync(t.data(), d_timing, sizeof(long long) * 16 * n, cudaMemcpyDeviceToHost, st));
        MCU(cudaStreamSynchronize(st));
        static const char *names[10] = {"copy-in + bucket fill", "rank sort", "runs + filters", "triangle counts", "sample + QEF", "collapse", "triangles", "edges", "write back", ""};   // phase k ends at PHASE(k + 1)
        long long sum[16] = {0};
        for (int b = 0; b < n; b++) for (int k = 0; k < 16; k++) sum[k] += t[16 * (size_t)b + k];
        fprintf(stderr, "[simplify timing] %d meshes; block 0 (largest: %d triangles) / mean over blocks, kilocycles\n", n, launch[0].numTriangles);
        for (int k = 0; k < 9; k++) fprintf(stderr, "[simplify timing]   %-24s %9.1f %9.1f\n", names[k], t[k] / 1e3, sum[k] / 1e3 / n);
    }
#endif
    if (d_packT) {
        MCU(cudaMemcpyAsync(B + oJobs, jd.data(), sizeof(SimpJobDev) * n, cudaMemcpyHostToDevice, st));
        k_pack_meshes<<<n, 256, 0, st>>>((const SimpJobDev *)(B + oJobs), d_results, n, d_V, d_T, d_packV, d_packT, d_packOffsets, d_packTotals,
                                         d_packP, physicsScale, nullptr);
        MCU(cudaGetLastError());
    }
    return LVN_SUCCESS;
}

}  // namespace lvn

using namespace lvn;

extern "C" const char *lvn_mesh_simplify_last_error(void) { return g_simpError; }

extern "C" int lvn_mesh_simplify_batch(int numMeshes, const lvn_simplify_job *jobs, const lvn_simplify_options *options, int numOptions,
                                       lvn_mesh_vertex *vertices, int64_t numVerticesTotal,
                                       lvn_mesh_triangle *triangles, int64_t numTrianglesTotal,
                                       lvn_simplify_result *results)
{
    if (numMeshes < 0 || !options || (numOptions != 1 && numOptions != numMeshes) ||
        (numMeshes > 0 && (!jobs || !results || !vertices || !triangles))) return LVN_ERR_INVALID_VALUE;
    if (numMeshes == 0) return LVN_SUCCESS;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return LVN_ERR_NO_DEVICE; }
    if (!g_simp.stream) MCU(cudaStreamCreateWithFlags(&g_simp.stream, cudaStreamNonBlocking));
    cudaStream_t st = g_simp.stream;
    std::vector<SimplifyMesh> meshes(numMeshes);
    for (int m = 0; m < numMeshes; m++) {
        const lvn_simplify_job &j = jobs[m];
        if (j.numVertices < 0 || j.numTriangles < 0 || j.vertexOffset < 0 || j.triangleOffset < 0 ||
            (int64_t)j.vertexOffset + j.numVertices > numVerticesTotal || (int64_t)j.triangleOffset + j.numTriangles > numTrianglesTotal)
            return LVN_ERR_INVALID_VALUE;
        SimplifyMesh &d = meshes[m];
        d.vertexOffset = j.vertexOffset; d.numVertices = j.numVertices;
        d.triangleOffset = j.triangleOffset; d.numTriangles = j.numTriangles;
        memcpy(d.offset, j.worldSpaceOffset, sizeof(d.offset));
        d.opt = options[numOptions == 1 ? 0 : m];
    }
    const size_t bV = simp_align(sizeof(lvn_mesh_vertex) * (size_t)numVerticesTotal), bT = simp_align(12 * (size_t)numTrianglesTotal);
    LV(ensure_buffer(&g_simp.d_io, &g_simp.ioCap, bV + bT + simp_align(sizeof(int4) * numMeshes)));
    char *IO = (char *)g_simp.d_io;
    lvn_mesh_vertex *dV = (lvn_mesh_vertex *)IO;
    int *dT = (int *)(IO + bV);
    int4 *dRes = (int4 *)(IO + bV + bT);
    MCU(cudaMemcpyAsync(dV, vertices, sizeof(lvn_mesh_vertex) * (size_t)numVerticesTotal, cudaMemcpyHostToDevice, st));
    MCU(cudaMemcpyAsync(dT, triangles, 12 * (size_t)numTrianglesTotal, cudaMemcpyHostToDevice, st));
    LV(simplify_device(numMeshes, meshes.data(), dV, dT, dRes, nullptr, nullptr, nullptr, nullptr, st, nullptr, 0.f, nullptr));
    std::vector<int4> res(numMeshes);
    MCU(cudaMemcpyAsync(res.data(), dRes, sizeof(int4) * numMeshes, cudaMemcpyDeviceToHost, st));
    // the simplified meshes stay in their input slots (a mesh never grows): two copies back
    MCU(cudaMemcpyAsync(vertices, dV, sizeof(lvn_mesh_vertex) * (size_t)numVerticesTotal, cudaMemcpyDeviceToHost, st));
    MCU(cudaMemcpyAsync(triangles, dT, 12 * (size_t)numTrianglesTotal, cudaMemcpyDeviceToHost, st));
    MCU(cudaStreamSynchronize(st));
    int rc = LVN_SUCCESS;
    for (int m = 0; m < numMeshes; m++) {
        results[m].numVertices = res[m].x; results[m].numTriangles = res[m].y;
        results[m].iterations = res[m].z; results[m].numEdges = res[m].w;
        if (res[m].z == -1) { rc = LVN_ERR_INVALID_VALUE; g_simpError = "a triangle index lies outside its mesh's vertices"; }
        else if (res[m].z == -2 && rc == LVN_SUCCESS) { rc = LVN_ERR_CAPACITY; g_simpError = "a mesh is too large for the simplifier's shared-memory masks"; }
    }
    return rc;
}
