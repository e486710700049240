// Host side of libleven_b200.so: the C ABI of include/leven_compute.h.
//
// Replaces the reference's OpenCL host layer (paths relative to the reference tree):
//   leven/src/compute.cpp               context, utilities, Compute_MeshGenContext
//   leven/src/compute_density_field.cpp noise table, field cache, CSG op store / replay
//   leven/src/compute_octree.cpp        octree cache, generateChunkMesh orchestration
//   leven/src/compute_csg.cpp           ApplyCSGOperations
//   leven/src/compute_cuckoo.cpp        table sizing, rehash loop
//
// One batch = one column kernel plus three kernels per lane (k_rows -> k_hermite -> k_leaves) and
// one host wait per lane (the reference: ~65 launches and ~10 blocking reads per chunk,
// SURVEY.md 3.2).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "lvn_internal.h"

using namespace lvn;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static std::string g_lastCudaError;

#define CU(call)                                                                             \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            char _b[512];                                                                    \
            snprintf(_b, sizeof(_b), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            g_lastCudaError = _b;                                                            \
            return _e == cudaErrorMemoryAllocation ? LVN_ERR_OUT_OF_MEMORY : LVN_ERR_CUDA;   \
        }                                                                                    \
    } while (0)
#define LV(call) do { int _r = (call); if (_r < 0) return _r; } while (0)

extern "C" const char *lvn_last_cuda_error(void) { return g_lastCudaError.c_str(); }

extern "C" const char *lvn_error_string(int error)
{
    switch (error) {
    case LVN_SUCCESS: return "LVN_SUCCESS";
    case LVN_CL_ERROR: return "LVN_CL_ERROR";
    case LVN_ERR_NO_DEVICE: return "LVN_ERR_NO_DEVICE";
    case LVN_ERR_OUT_OF_MEMORY: return "LVN_ERR_OUT_OF_MEMORY";
    case LVN_ERR_INVALID_VALUE: return "LVN_ERR_INVALID_VALUE";
    case LVN_ERR_NOT_INITIALISED: return "LVN_ERR_NOT_INITIALISED";
    case LVN_ERR_CAPACITY: return "LVN_ERR_CAPACITY";
    case LVN_ERR_CUDA: return "LVN_ERR_CUDA";
    default: return "Unknown leven_b200 error code!";
    }
}

// ---------------------------------------------------------------------------
// device buffers
// ---------------------------------------------------------------------------
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;   // elements
    int reserve(size_t n, bool keep = false)
    {
        if (n <= cap) return 0;
        size_t want = std::max(n, cap + cap / 2);
        T *np = nullptr;
        CU(cudaMalloc((void **)&np, want * sizeof(T)));
        if (keep && p && cap) CU(cudaMemcpy(np, p, cap * sizeof(T), cudaMemcpyDeviceToDevice));
        if (p) cudaFree(p);
        p = np;
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <typename T>
struct PinBuf {
    T *p = nullptr;
    size_t cap = 0;
    int reserve(size_t n)
    {
        if (n <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        // mapped: kernels may write results straight into it (zero-copy), no copy-engine transfer
        CU(cudaHostAlloc((void **)&p, n * sizeof(T), cudaHostAllocMapped));
        cap = n;
        return 0;
    }
    T *dev() const   // the device-side address of the mapping
    {
        T *d = nullptr;
        return (p && cudaHostGetDevicePointer((void **)&d, p, 0) == cudaSuccess) ? d : nullptr;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// ---------------------------------------------------------------------------
// process-wide state (ComputeContext, compute_local.h:17-24; g_storedOps, compute_density_field.cpp:23-24)
// ---------------------------------------------------------------------------
struct Global {
    bool initialised = false;
    int device = 0;
    bool deviceChosen = false;
    std::vector<uint8_t> image;
    float2 *d_grad2 = nullptr;
    float *d_grad2x = nullptr;     // x plane, then y plane
    float4 *d_grad3 = nullptr;
    int defaultMaterial = 0;
    int densityKind = 0;
    float densityParam = 0.5f;
    std::vector<lvn_csg_operation_info> storedOps;
    std::vector<lvn_aabb> storedAABBs;
    std::mt19937 cuckooRng;   // compute_cuckoo.cpp:46
};
static Global g;

static DensityParams density_params()
{
    DensityParams dp;
    dp.grad2 = g.d_grad2;
    dp.grad2x = g.d_grad2x;
    dp.grad3 = g.d_grad3;
    dp.kind = g.densityKind;
    dp.param = g.densityParam;
    dp.defaultMaterial = g.defaultMaterial;
    dp.negZero = -0.f;
    return dp;
}

extern "C" int lvn_compute_set_device(int cudaDevice)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return LVN_ERR_NO_DEVICE;
    if (cudaDevice < 0 || cudaDevice >= count) return LVN_ERR_INVALID_VALUE;
    g.device = cudaDevice;
    g.deviceChosen = true;
    CU(cudaSetDevice(cudaDevice));
    return LVN_SUCCESS;
}

// One process per GPU on a two-socket box: the pinned arenas a process allocates afterwards should
// sit on the socket its GPU hangs off, or every download crosses the socket link and the ranks
// share one memory controller.  Reads the device's PCI address, its NUMA node from sysfs, restricts
// the calling thread (threads made later inherit it) to that node's CPUs -- those the process is
// allowed to use -- and prefers the node for its page allocations.  Returns the node, or a negative
// value when the platform gives no answer (nothing is changed then); cpusBound = CPUs in the new
// mask (0: affinity left alone), memoryBound = 1 when set_mempolicy took.
extern "C" int lvn_compute_bind_host_numa(int cudaDevice, int *cpusBound, int *memoryBound)
{
    if (cpusBound) *cpusBound = 0;
    if (memoryBound) *memoryBound = 0;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), cudaDevice) != cudaSuccess) { cudaGetLastError(); return LVN_ERR_INVALID_VALUE; }
    for (char *c = bus; *c; c++) if (*c >= 'A' && *c <= 'F') *c = (char)(*c - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    int node = -1;
    if (FILE *f = fopen(path, "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    if (node < 0 || node >= 1024) return -1;
    // the node's CPUs, "a-b,c,d-e"
    cpu_set_t want, have, both;
    CPU_ZERO(&want);
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    if (FILE *f = fopen(path, "r")) {
        int a, b;
        while (fscanf(f, "%d", &a) == 1) {
            b = a;
            int ch = fgetc(f);
            if (ch == '-') { if (fscanf(f, "%d", &b) != 1) b = a; ch = fgetc(f); }
            for (int c = a; c <= b && c < CPU_SETSIZE; c++) CPU_SET(c, &want);
            if (ch != ',') break;
        }
        fclose(f);
    }
    if (sched_getaffinity(0, sizeof(have), &have) == 0) {
        CPU_AND(&both, &want, &have);
        const int nb = CPU_COUNT(&both);
        if (nb > 0 && sched_setaffinity(0, sizeof(both), &both) == 0 && cpusBound) *cpusBound = nb;
    }
    unsigned long mask[1024 / (8 * sizeof(unsigned long))] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
#ifdef SYS_set_mempolicy
    if (syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, 1024ul + 1) == 0 && memoryBound) *memoryBound = 1;
#endif
    return node;
}

// ---- noise table (compute_density_field.cpp:28-125) -------------------------
static const int kPerm256[256] = {151,160,137,91,90,15,
  131,13,201,95,96,53,194,233,7,225,140,36,103,30,69,142,8,99,37,240,21,10,23,
  190, 6,148,247,120,234,75,0,26,197,62,94,252,219,203,117,35,11,32,57,177,33,
  88,237,149,56,87,174,20,125,136,171,168, 68,175,74,165,71,134,139,48,27,166,
  77,146,158,231,83,111,229,122,60,211,133,230,220,105,92,41,55,46,245,40,244,
  102,143,54, 65,25,63,161, 1,216,80,73,209,76,132,187,208, 89,18,169,200,196,
  135,130,116,188,159,86,164,100,109,198,173,186, 3,64,52,217,226,250,124,123,
  5,202,38,147,118,126,255,82,85,212,207,206,59,227,47,16,58,17,182,189,28,42,
  223,183,170,213,119,248,152, 2,44,154,163, 70,221,153,101,155,167, 43,172,9,
  129,22,39,253, 19,98,108,110,79,113,224,232,178,185, 112,104,218,246,97,228,
  251,34,242,193,238,210,144,12,191,179,162,241, 81,51,145,235,249,14,239,107,
  49,192,214, 31,181,199,106,157,184, 84,204,176,115,121,50,45,127, 4,150,254,
  138,236,205,93,222,114,67,29,24,72,243,141,128,195,78,66,215,61,156,180};
static const int kGrad3[16][3] = {{0,1,1},{0,1,-1},{0,-1,1},{0,-1,-1},{1,0,1},{1,0,-1},{-1,0,1},{-1,0,-1},
    {1,1,0},{1,-1,0},{-1,1,0},{-1,-1,0},{1,0,-1},{-1,0,-1},{0,-1,1},{0,1,1}};

static unsigned int noise_hash(int x, int y, int seed)   // NoiseHash, compute_density_field.cpp:69-88
{
    const unsigned int key = (((unsigned int)x << 24) | ((unsigned int)y << 16)) ^ (unsigned int)seed;
    unsigned int hash = 0;
    for (int i = 0; i < 4; i++) {
        hash += (key >> (8 * i)) & 0xffu;
        hash += (hash << 10);
        hash ^= (hash >> 6);
    }
    hash += (hash << 3);
    hash ^= (hash >> 11);
    hash += (hash << 15);
    return hash;
}

static void make_noise_image(int seed, std::vector<uint8_t> &rgba)
{
    // the reference initialiser holds 506 values: the table, then the table again from entry 6
    // on; the last 6 of the 512 slots are zero (compute_density_field.cpp:28-53)
    int perm[512];
    for (int i = 0; i < 256; i++) perm[i] = kPerm256[i];
    for (int i = 0; i < 250; i++) perm[256 + i] = kPerm256[6 + i];
    for (int i = 506; i < 512; i++) perm[i] = 0;
    // documented stand-in for std::shuffle(std::default_random_engine(seed)): Fisher-Yates from
    // the top with std::mt19937(seed), j = next() % (i + 1)
    std::mt19937 mt((unsigned int)seed);
    for (int i = 511; i > 0; i--) std::swap(perm[i], perm[mt() % (unsigned int)(i + 1)]);
    rgba.resize(256 * 256 * 4);
    for (int i = 0; i < 256; i++)
        for (int j = 0; j < 256; j++) {
            const int offset = ((i * 256) + j) * 4;
            const unsigned char value = (unsigned char)perm[noise_hash(i, j, seed) & 0x1ff];
            rgba[offset + 0] = (uint8_t)(kGrad3[value & 0x0f][0] * 64 + 64);
            rgba[offset + 1] = (uint8_t)(kGrad3[value & 0x0f][1] * 64 + 64);
            rgba[offset + 2] = (uint8_t)(kGrad3[value & 0x0f][2] * 64 + 64);
            rgba[offset + 3] = value;
        }
}

static float int_as_float_host(int v) { float f; memcpy(&f, &v, 4); return f; }

// read_imagef(...).xyz * 4.f - 1.f for a UNORM8 texel (simplex.cl:124): byte/255 correctly rounded
static float unorm_grad(uint8_t b)
{
    volatile float q = (float)b / 255.0f;
    volatile float m = q * 4.f;
    return m - 1.f;
}

static int upload_noise_image()
{
    std::vector<float2> g2(257 * 257);   // + one wrapped column and row (density.cuh: LVN_G2PITCH)
    std::vector<float4> g3(65536);
    std::vector<uint8_t> perm3(65536);   // the first lookup of a snoise3 corner wants the column alone: a byte table behind the gradients
    for (int r = 0; r < 257; r++)
        for (int c = 0; c < 257; c++) {
            const uint8_t *px = &g.image[(size_t)(((r & 255) << 8) | (c & 255)) * 4];
            g2[(size_t)r * 257 + c] = make_float2(unorm_grad(px[0]), unorm_grad(px[1]));
        }
    for (int t = 0; t < 65536; t++) {
        const uint8_t *px = &g.image[(size_t)t * 4];
        // snoise3's second lookup uses the UNORM alpha v/255 as an un-centred x coordinate:
        // NEAREST + REPEAT lands on column v, and on column 0 for v = 255 (simplex.cl:184-185)
        const int col = px[3] == 255 ? 0 : (int)px[3];
        g3[t] = make_float4(unorm_grad(px[0]), unorm_grad(px[1]), unorm_grad(px[2]), int_as_float_host(col));
        perm3[t] = (uint8_t)col;
    }
    if (!g.d_grad2) CU(cudaMalloc((void **)&g.d_grad2, g2.size() * sizeof(float2)));
    if (!g.d_grad3) CU(cudaMalloc((void **)&g.d_grad3, 65536 * sizeof(float4) + 65536));
    std::vector<float> planes(2 * g2.size());
    for (size_t i = 0; i < g2.size(); i++) { planes[i] = g2[i].x; planes[g2.size() + i] = g2[i].y; }
    if (!g.d_grad2x) CU(cudaMalloc((void **)&g.d_grad2x, planes.size() * sizeof(float)));
    CU(cudaMemcpy(g.d_grad2x, planes.data(), planes.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(g.d_grad2, g2.data(), g2.size() * sizeof(float2), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(g.d_grad3, g3.data(), 65536 * sizeof(float4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(g.d_grad3 + 65536, perm3.data(), 65536, cudaMemcpyHostToDevice));
    return LVN_SUCCESS;
}

extern "C" int lvn_compute_set_noise_image(const uint8_t *rgba)
{
    if (!g.initialised) return LVN_ERR_NOT_INITIALISED;
    if (!rgba) return LVN_ERR_INVALID_VALUE;
    g.image.assign(rgba, rgba + 256 * 256 * 4);
    return upload_noise_image();
}

extern "C" int lvn_compute_get_noise_image(uint8_t *rgba)
{
    if (!g.initialised) return LVN_ERR_NOT_INITIALISED;
    memcpy(rgba, g.image.data(), 256 * 256 * 4);
    return LVN_SUCCESS;
}

extern "C" int lvn_compute_set_noise_seed(int noiseSeed)
{
    if (!g.initialised) return LVN_ERR_NOT_INITIALISED;
    make_noise_image(noiseSeed, g.image);
    return upload_noise_image();
}

extern "C" int lvn_compute_initialise(int noiseSeed, unsigned int defaultMaterial, int numCSGBrushes)
{
    (void)numCSGBrushes;   // NUM_CSG_BRUSHES: the two brush shapes are compiled in (compute.cpp:222)
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        g_lastCudaError = "no CUDA device";
        return LVN_ERR_NO_DEVICE;
    }
    if (defaultMaterial > 255u) return LVN_ERR_INVALID_VALUE;   // materials are stored as u8
    if (g.deviceChosen) CU(cudaSetDevice(g.device));
    else CU(cudaGetDevice(&g.device));
    {   // keep freed octree-cache blocks in the device's pool instead of returning them to the OS
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, g.device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    g.initialised = true;
    g.defaultMaterial = (int)defaultMaterial;
    g.cuckooRng = std::mt19937();
    return lvn_compute_set_noise_seed(noiseSeed);
}

extern "C" int lvn_compute_shutdown(void)
{
    if (g.d_grad2) cudaFree(g.d_grad2);
    if (g.d_grad3) cudaFree(g.d_grad3);
    if (g.d_grad2x) cudaFree(g.d_grad2x);
    g.d_grad2x = nullptr;
    g.d_grad2 = nullptr;
    g.d_grad3 = nullptr;
    g.initialised = false;
    g.storedOps.clear();
    g.storedAABBs.clear();
    return LVN_SUCCESS;
}

extern "C" int lvn_compute_set_density_function(int kind, float param)
{
    if (kind != 0 && kind != 1) return LVN_ERR_INVALID_VALUE;
    g.densityKind = kind;
    g.densityParam = param;
    return LVN_SUCCESS;
}

extern "C" int lvn_compute_store_csg_operation(const lvn_csg_operation_info *op, const lvn_aabb *aabb)
{
    if (!op || !aabb) return LVN_ERR_INVALID_VALUE;
    g.storedOps.push_back(*op);
    g.storedAABBs.push_back(*aabb);
    return LVN_SUCCESS;
}

extern "C" int lvn_compute_clear_csg_operations(void)
{
    g.storedOps.clear();
    g.storedAABBs.clear();
    return LVN_SUCCESS;
}

// ---------------------------------------------------------------------------
// mesh generation context (MeshGenerationContext, compute_local.h:57-72)
// ---------------------------------------------------------------------------
struct Key {
    int x, y, z, s;
    bool operator==(const Key &o) const { return x == o.x && y == o.y && z == o.z && s == o.s; }
};
struct KeyHash {
    size_t operator()(const Key &k) const
    {
        size_t h = 1469598103934665603ull;
        for (int v : {k.x, k.y, k.z, k.s}) { h ^= (size_t)(unsigned int)v; h *= 1099511628211ull; }
        return h;
    }
};

// GPUDensityField (compute_local.h:28-38) for fields that must persist: CSG-edited ones and
// those created by isChunkEmpty.  The cuckoo table maps edge key -> slot (compute_octree.cpp:107-109).
struct FieldEntry {
    int lastCSGOperation = 0;
    int numEdges = 0;
    uint8_t *d_field = nullptr;
    int *d_keys = nullptr;
    float4 *d_info = nullptr;
    unsigned long long *d_table = nullptr;
    unsigned int prime = 0;
    unsigned int params[8] = {0};
    int cuckooRetries = 0;
    void release(cudaStream_t st)   // stream-ordered pool, like the octree entries
    {
        if (d_field) cudaFreeAsync(d_field, st);
        if (d_keys) cudaFreeAsync(d_keys, st);
        if (d_info) cudaFreeAsync(d_info, st);
        if (d_table) cudaFreeAsync(d_table, st);
        d_field = nullptr; d_keys = nullptr; d_info = nullptr; d_table = nullptr;
    }
};

// GPUOctree (compute_local.h:44-50) reduced to what a later generateChunkMesh returns.  One
// stream-ordered allocation from the device's memory pool per entry: cudaMalloc / cudaFree
// would synchronise the device on every chunk request and eviction.
struct OctreeEntry {
    int numNodes = 0, numQuads = 0, numSeams = 0;
    void *block = nullptr;
    lvn_mesh_vertex *d_v = nullptr;
    int *d_t = nullptr;
    lvn_seam_node_info *d_s = nullptr;
    int alloc(cudaStream_t st)
    {
        const size_t bv = (size_t)numNodes * sizeof(lvn_mesh_vertex), bs = (size_t)numSeams * sizeof(lvn_seam_node_info),
                     bt = (size_t)numQuads * 6 * sizeof(int);
        if (bv + bs + bt == 0) return LVN_SUCCESS;
        CU(cudaMallocAsync(&block, bv + bs + bt, st));
        d_v = (lvn_mesh_vertex *)block;                               // 48 B records first: 16 B alignment holds
        d_s = (lvn_seam_node_info *)((char *)block + bv);
        d_t = (int *)((char *)block + bv + bs);
        return LVN_SUCCESS;
    }
    void release(cudaStream_t st)
    {
        if (block) cudaFreeAsync(block, st);
        block = nullptr; d_v = nullptr; d_t = nullptr; d_s = nullptr;
    }
};

constexpr int LVN_MAX_LANES = 32;
constexpr int LVN_MAX_STREAMS = 4;

struct lvn_meshgen {
    Dims dims;
    cudaStream_t stream = nullptr;      // where the kernels go (own stream or the caller's)
    cudaStream_t ownStream = nullptr;
    std::unordered_map<Key, FieldEntry, KeyHash> fields;
    std::unordered_map<Key, OctreeEntry, KeyHash> octrees;

    // batch workspace
    DevBuf<ChunkDesc> d_descs;
    DevBuf<ChunkHdr> d_hdrs;
    DevBuf<float> d_heights;
    DevBuf<unsigned long long> d_bitsLo;
    DevBuf<unsigned int> d_bitsHi, d_rowE, d_rowN, d_rowQ, d_rowS;
    DevBuf<int> d_edgeKeys;
    DevBuf<float4> d_edgeInfo;
    DevBuf<int2> d_xzList;         // (chunk, edge slot) of every x/z edge of a lane: the Hermite search list
    DevBuf<lvn_mesh_vertex> d_vertices;
    DevBuf<float4> d_qef;          // 4 x float4 per slot of the vertex arena: the QEF records between k_leaves and k_solve
    DevBuf<int> d_tris;
    DevBuf<lvn_seam_node_info> d_seams;
    DevBuf<uint4> d_slab;
    DevBuf<unsigned int> d_slabEy, d_ticket;
    DevBuf<int> d_candidates;      // per lane slice: the chunks that can contain surface (k_candidates -> k_rows)
    DevBuf<TileRef> d_edgeTiles, d_nodeTiles;  // tile directories, one slice per lane
    DevBuf<uint8_t> d_tmpFields;
    DevBuf<float> d_tmpDensity;   // the density values behind d_tmpFields (ChunkDesc::latticeDensity)
    DevBuf<uint8_t *> d_fieldPtrs;
    // debug stage outputs
    DevBuf<unsigned int> d_dbgCodes;
    DevBuf<int> d_dbgMasks, d_dbgMats;
    DevBuf<float> d_dbgQefs;
    DevBuf<float4> d_dbgPos, d_dbgNrm;
    // csg scratch
    DevBuf<unsigned int> d_touched, d_csgCounts;   // d_touched: an edit's counters, ticket and touched-edge bitmaps
    DevBuf<unsigned char> d_csgStage;              // an edit's [chunks | ops]
    PinBuf<unsigned char> h_csgStage;
    // hash tables built without waiting for their insert flags (apply_csg_items): the flags land in
    // h_tableFlags behind the stream's next synchronisation, run_batch looks at them before it trusts
    // a result that read those tables (validate_pending_tables)
    std::vector<FieldEntry *> pendingTables;
    PinBuf<unsigned int> h_tableFlags;
    int lookahead = 2;                     // host path: lanes queued ahead of the one whose copies are being queued (LVN_LOOKAHEAD)
    bool forceTableRetry = false;          // LVN_TEST_CUCKOO_RETRY=1: treat every first insertion as failed (tests)
    bool noLatticeDensity = false;         // LVN_TEST_NO_LATTICE_DENSITY=1: k_hermite evaluates all 17 steps itself (tests)
    // simplified batch (lvn_meshgen_generate_simplified_batch)
    DevBuf<int4> d_simpRes;
    DevBuf<int2> d_packOff;
    cudaStream_t simpStreamB = nullptr;    // the early group of a split simplifier launch (simplify.cu)
    cudaEvent_t evSimpFork = nullptr, evSimpEarly = nullptr, evSimpJoin = nullptr;
    DevBuf<lvn_mesh_vertex> d_packV;
    DevBuf<float4> d_packP;
    DevBuf<int> d_packT;
    PinBuf<int4> h_simpRes;
    PinBuf<int2> h_packOff;

    PinBuf<ChunkDesc> h_descs;
    PinBuf<ChunkHdr> h_hdrs;
    PinBuf<int4> h_colOrigins;
    PinBuf<unsigned int> h_small;

    ArenaCounters lastCounters = {};   // totals over the lanes of the last batch
    int hintChunks = 0;                // what the last batch of this many chunks produced: a batch list repeats
    unsigned int hintNodes = 0, hintNonEmpty = 0;   // in steady state, and a sparse one wants fewer lanes
    int lastN = 0;

    // lanes of a batch (run_batch): independent slices of the chunk list, each a chain
    // classify -> hermite -> leaves on one of `numStreams` streams, so that the tail of one
    // lane's kernel overlaps the next lane's work and (host path) the copy engine drains lane k
    // while lane k+1 computes
    int cfgLanes = 0;                  // 0 = chosen per call (choose_pipeline)
    int cfgStreams = 2;
    int numStreams = 1;                // of the last batch
    cudaStream_t laneStreams[LVN_MAX_STREAMS] = {nullptr};   // [0] unused: the context's stream
    cudaStream_t copyStream = nullptr;
    cudaStream_t pubStream = nullptr;      // k_publish of the host path (kernels_chunk.cu)
    cudaEvent_t evFork = nullptr, evCopy = nullptr, evPubJoin = nullptr;
    cudaEvent_t evLane[LVN_MAX_LANES] = {nullptr};
    cudaEvent_t evPub[LVN_MAX_LANES] = {nullptr};
    cudaEvent_t evRows[LVN_MAX_LANES] = {nullptr};
    cudaEvent_t evJoin[LVN_MAX_STREAMS] = {nullptr};
    int numLanes = 0;
    int laneFirst[LVN_MAX_LANES + 1] = {0};      // lane k owns internal chunks [laneFirst[k], laneFirst[k+1])
    ArenaCaps laneBase[LVN_MAX_LANES] = {};      // where lane k's slice of every arena starts
    ArenaCounters laneCounters[LVN_MAX_LANES] = {};
    int64_t hostBase[LVN_MAX_LANES][3] = {};     // vertices, triangles, seam nodes: lane k's place in the host arenas
    bool hostLayout = false;                     // results address the caller's host arenas
    std::vector<int> perm;                       // internal (lane-major) chunk order -> caller's index
    bool trace = false;                          // LVN_TRACE=1: per-lane kernel timeline on stderr
    std::vector<cudaEvent_t> traceEv, traceCopyEv;

    bool profiling = false;
    cudaEvent_t ev[2 * LVN_NUM_STAGES] = {nullptr};
    bool evUsed[LVN_NUM_STAGES] = {false};
    lvn_stage_stats stats = {};
};

static Key make_key(const int32_t min[3], int size) { return Key{min[0], min[1], min[2], size}; }

extern "C" lvn_meshgen *lvn_meshgen_create(int voxelsPerChunk)
{
    if (!g.initialised) return nullptr;
    // power of two, 8..64: one sign row must fit 96 bits
    if (voxelsPerChunk < 8 || voxelsPerChunk > 64 || (voxelsPerChunk & (voxelsPerChunk - 1))) return nullptr;
    lvn_meshgen *ctx = new lvn_meshgen;
    int l = 0;
    while ((1 << (l + 1)) <= voxelsPerChunk) l++;
    ctx->dims.V = voxelsPerChunk;
    ctx->dims.H = voxelsPerChunk + 1;
    ctx->dims.F = voxelsPerChunk + 2;
    ctx->dims.depth = l;            // MAX_OCTREE_DEPTH, compute.cpp:271
    ctx->dims.shift = l + 1;        // compute.cpp:251
    ctx->dims.mask = (1 << (l + 1)) - 1;
    if (cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return nullptr; }
    ctx->stream = ctx->ownStream;
    for (int i = 0; i < 2 * LVN_NUM_STAGES; i++) cudaEventCreate(&ctx->ev[i]);
    for (int i = 1; i < LVN_MAX_STREAMS; i++) cudaStreamCreateWithFlags(&ctx->laneStreams[i], cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking);
    {   // k_publish must slip in between the blocks of a running Hermite kernel: highest priority
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamCreateWithPriority(&ctx->pubStream, cudaStreamNonBlocking, hi);
    }
    cudaEventCreateWithFlags(&ctx->evPubJoin, cudaEventDisableTiming);
    for (int i = 0; i < LVN_MAX_LANES; i++) cudaEventCreateWithFlags(&ctx->evPub[i], cudaEventDisableTiming);
    for (int i = 0; i < LVN_MAX_LANES; i++) cudaEventCreateWithFlags(&ctx->evRows[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->evCopy, cudaEventDisableTiming);
    for (int i = 0; i < LVN_MAX_LANES; i++) cudaEventCreateWithFlags(&ctx->evLane[i], cudaEventDisableTiming);
    for (int i = 0; i < LVN_MAX_STREAMS; i++) cudaEventCreateWithFlags(&ctx->evJoin[i], cudaEventDisableTiming);
    if (const char *e = getenv("LVN_LANES")) ctx->cfgLanes = atoi(e);
    if (const char *e = getenv("LVN_STREAMS")) ctx->cfgStreams = atoi(e);
    if (const char *e = getenv("LVN_TRACE")) ctx->trace = atoi(e) != 0;
    if (const char *e = getenv("LVN_TEST_CUCKOO_RETRY")) ctx->forceTableRetry = atoi(e) != 0;
    if (const char *e = getenv("LVN_TEST_NO_LATTICE_DENSITY")) ctx->noLatticeDensity = atoi(e) != 0;
    if (const char *e = getenv("LVN_LOOKAHEAD")) ctx->lookahead = std::max(1, atoi(e));
    return ctx;
}

extern "C" void lvn_meshgen_destroy(lvn_meshgen *ctx)
{
    if (!ctx) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->fields) kv.second.release(ctx->stream);
    for (auto &kv : ctx->octrees) kv.second.release(ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    ctx->d_descs.release(); ctx->d_hdrs.release(); ctx->d_heights.release();
    ctx->d_bitsLo.release(); ctx->d_bitsHi.release(); ctx->d_rowE.release(); ctx->d_rowN.release();
    ctx->d_rowQ.release(); ctx->d_rowS.release(); ctx->d_edgeKeys.release(); ctx->d_edgeInfo.release(); ctx->d_xzList.release();
    ctx->d_vertices.release(); ctx->d_qef.release(); ctx->d_tris.release(); ctx->d_seams.release();
    ctx->d_slab.release(); ctx->d_slabEy.release(); ctx->d_ticket.release(); ctx->d_candidates.release();
    ctx->d_tmpFields.release(); ctx->d_tmpDensity.release(); ctx->d_fieldPtrs.release();
    ctx->d_edgeTiles.release(); ctx->d_nodeTiles.release();
    ctx->d_dbgCodes.release(); ctx->d_dbgMasks.release(); ctx->d_dbgMats.release(); ctx->d_dbgQefs.release();
    ctx->d_dbgPos.release(); ctx->d_dbgNrm.release(); ctx->d_touched.release(); ctx->d_csgCounts.release();
    ctx->d_csgStage.release(); ctx->h_csgStage.release(); ctx->h_tableFlags.release();
    ctx->d_simpRes.release(); ctx->d_packOff.release(); ctx->d_packV.release(); ctx->d_packP.release(); ctx->d_packT.release();
    ctx->h_simpRes.release(); ctx->h_packOff.release();
    if (ctx->simpStreamB) cudaStreamDestroy(ctx->simpStreamB);
    if (ctx->evSimpFork) cudaEventDestroy(ctx->evSimpFork);
    if (ctx->evSimpEarly) cudaEventDestroy(ctx->evSimpEarly);
    if (ctx->evSimpJoin) cudaEventDestroy(ctx->evSimpJoin);
    ctx->h_descs.release(); ctx->h_hdrs.release(); ctx->h_colOrigins.release();
    ctx->h_small.release();
    for (int i = 0; i < 2 * LVN_NUM_STAGES; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 1; i < LVN_MAX_STREAMS; i++) if (ctx->laneStreams[i]) cudaStreamDestroy(ctx->laneStreams[i]);
    if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
    if (ctx->pubStream) cudaStreamDestroy(ctx->pubStream);
    if (ctx->evPubJoin) cudaEventDestroy(ctx->evPubJoin);
    for (int i = 0; i < LVN_MAX_LANES; i++) if (ctx->evPub[i]) cudaEventDestroy(ctx->evPub[i]);
    for (int i = 0; i < LVN_MAX_LANES; i++) if (ctx->evRows[i]) cudaEventDestroy(ctx->evRows[i]);
    if (ctx->evFork) cudaEventDestroy(ctx->evFork);
    if (ctx->evCopy) cudaEventDestroy(ctx->evCopy);
    for (int i = 0; i < LVN_MAX_LANES; i++) if (ctx->evLane[i]) cudaEventDestroy(ctx->evLane[i]);
    for (int i = 0; i < LVN_MAX_STREAMS; i++) if (ctx->evJoin[i]) cudaEventDestroy(ctx->evJoin[i]);
    cudaStreamDestroy(ctx->ownStream);
    delete ctx;
}

extern "C" int lvn_meshgen_voxels_per_chunk(const lvn_meshgen *ctx) { return ctx ? ctx->dims.V : 0; }

extern "C" int lvn_meshgen_set_profiling(lvn_meshgen *ctx, int enabled)
{
    if (!ctx) return LVN_ERR_INVALID_VALUE;
    ctx->profiling = enabled != 0;
    return LVN_SUCCESS;
}

extern "C" int lvn_meshgen_set_stream(lvn_meshgen *ctx, void *cudaStream)
{
    if (!ctx) return LVN_ERR_INVALID_VALUE;
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = cudaStream ? (cudaStream_t)cudaStream : ctx->ownStream;
    return LVN_SUCCESS;
}

extern "C" int lvn_meshgen_set_pipeline(lvn_meshgen *ctx, int lanes, int streams)
{
    if (!ctx || lanes < 0 || lanes > LVN_MAX_LANES || streams < 1 || streams > LVN_MAX_STREAMS) return LVN_ERR_INVALID_VALUE;
    ctx->cfgLanes = lanes;
    ctx->cfgStreams = streams;
    return LVN_SUCCESS;
}

extern "C" int lvn_meshgen_get_pipeline(const lvn_meshgen *ctx, int *lanesOfLastBatch, int *streams)
{
    if (!ctx) return LVN_ERR_INVALID_VALUE;
    if (lanesOfLastBatch) *lanesOfLastBatch = ctx->numLanes;
    if (streams) *streams = ctx->numStreams;
    return LVN_SUCCESS;
}

extern "C" void *lvn_alloc_pinned(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

extern "C" void lvn_free_pinned(void *p) { if (p) cudaFreeHost(p); }

extern "C" int lvn_measure_fp32_peak(double *tflops)
{
    if (!g.initialised) return LVN_ERR_NOT_INITIALISED;
    if (!tflops) return LVN_ERR_INVALID_VALUE;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, g.device));
    float *sink = nullptr;
    CU(cudaMalloc((void **)&sink, 4));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CU(cudaEventRecord(e0, 0));
        launch_fma_peak(sink, iters, blocks, 0);
        CU(cudaEventRecord(e1, 0));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
        if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops = best;
    return LVN_SUCCESS;
}

extern "C" int lvn_meshgen_get_stats(lvn_meshgen *ctx, lvn_stage_stats *out, int reset)
{
    if (!ctx || !out) return LVN_ERR_INVALID_VALUE;
    *out = ctx->stats;
    if (reset) memset(&ctx->stats, 0, sizeof(ctx->stats));
    return LVN_SUCCESS;
}

struct StageTimer {   // CUDA events around one stage on the context's stream
    lvn_meshgen *ctx;
    int stage;
    StageTimer(lvn_meshgen *c, int s, int launches) : ctx(c), stage(s)
    {
        ctx->stats.launches[stage] += launches;
        if (ctx->profiling) cudaEventRecord(ctx->ev[2 * stage], ctx->stream);
    }
    ~StageTimer()
    {
        if (ctx->profiling) { cudaEventRecord(ctx->ev[2 * stage + 1], ctx->stream); ctx->evUsed[stage] = true; }
    }
};

static void collect_stage_times(lvn_meshgen *ctx)   // call after the stream is synchronised
{
    if (!ctx->profiling) return;
    for (int s = 0; s < LVN_NUM_STAGES; s++) {
        if (!ctx->evUsed[s]) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev[2 * s], ctx->ev[2 * s + 1]) == cudaSuccess) ctx->stats.ms[s] += ms;
        ctx->evUsed[s] = false;
    }
}

// ColourForMinLeafSize(clipmapNodeSize / CLIPMAP_LEAF_SIZE), clipmap.cpp:329-352, compute_octree.cpp:252
static void colour_for_size(int size, float rgb[3])
{
    switch (size / (LVN_LEAF_SIZE_SCALE * 64)) {
    case 1:  rgb[0] = 0.3f; rgb[1] = 0.1f; rgb[2] = 0.f;  break;
    case 2:  rgb[0] = 0.f;  rgb[1] = 0.f;  rgb[2] = 0.5f; break;
    case 4:  rgb[0] = 0.f;  rgb[1] = 0.5f; rgb[2] = 0.5f; break;
    case 8:  rgb[0] = 0.5f; rgb[1] = 0.f;  rgb[2] = 0.5f; break;
    case 16: rgb[0] = 0.f;  rgb[1] = 0.5f; rgb[2] = 0.f;  break;
    default: rgb[0] = 0.5f; rgb[1] = 0.f;  rgb[2] = 0.f;  break;
    }
}

// ---------------------------------------------------------------------------
// one pass of the path over n chunks; results stay in the context's arenas
// ---------------------------------------------------------------------------
struct BatchOpts {
    bool debug = false;          // fill the per-node stage dumps
    bool ignoreFieldCache = false;
    bool singleLane = false;     // keep descriptors / headers in the caller's order
    bool countsFirst = false;    // return once every lane's counts are published (they are final after k_rows) and, on
                                 // the host path, its copies are queued; the rest completes in stream order
};

struct HostOut {   // caller-owned destination of lvn_meshgen_generate_batch
    lvn_mesh_vertex *vertices;
    int64_t vertexCapacity;
    lvn_mesh_triangle *triangles;
    int64_t triangleCapacity;
    lvn_seam_node_info *seams;
    int64_t seamCapacity;
};

static int fill_desc(lvn_meshgen *ctx, const int32_t *ms, ChunkDesc &cd)
{
    const Dims &d = ctx->dims;
    const int size = ms[3];
    if (size <= 0 || size % (d.V * LVN_LEAF_SIZE_SCALE) != 0) return LVN_ERR_INVALID_VALUE;
    memset(&cd, 0, sizeof(cd));
    cd.scale = size / (d.V * LVN_LEAF_SIZE_SCALE);
    cd.ox = ms[0] / LVN_LEAF_SIZE_SCALE;   // LeafScaleVec: C++ integer division
    cd.oy = ms[1] / LVN_LEAF_SIZE_SCALE;
    cd.oz = ms[2] / LVN_LEAF_SIZE_SCALE;
    cd.minx = ms[0]; cd.miny = ms[1]; cd.minz = ms[2];
    cd.size = size;
    colour_for_size(size, cd.colour);
    return LVN_SUCCESS;
}

// Lanes x streams of a batch.  Measured on B200 (profiles/r01_pipeline.md): every kernel
// boundary costs ~30 us extra while a bulk device-to-host copy is in flight, and a lane has
// ~40 us of launch gaps and kernel tails, so few lanes win: the device-resident pass runs two
// lanes on two streams (the tail of one kernel overlaps the other lane), the host pass four
// lanes on one stream (lane k is copied out while lanes k+1.. compute).
static void choose_pipeline(const lvn_meshgen *ctx, int n, const BatchOpts &opts, bool hostPath, int &lanes, int &streams)
{
    lanes = ctx->cfgLanes;
    streams = ctx->cfgStreams;
    if (lanes <= 0) {
        // device-resident: 2 lanes x 2 streams; host path: 8 x 2 (the first copy starts after 1/8 of the
        // batch; since the lane headers are published early a lane boundary costs ~1 us under D2H
        // load, profiles/r01h_notes.md)
        lanes = n >= 128 ? (hostPath ? 8 : 2) : 1;
        streams = 2;
        // a lane costs ~40 us of launches and kernel tails whatever it holds: a batch that was sparse the last
        // time (one rank's share of the sweep at 8 GPUs: 512 chunks, 48 with surface) runs in fewer lanes
        if (lanes > 1 && ctx->hintChunks == n) {
            if (hostPath) lanes = (int)std::max(1u, std::min(8u, ctx->hintNodes / 50000u));
            else if (ctx->hintNonEmpty < 80u) lanes = 1;
        }
    }
    // per-stage event timing and the stage dumps want one kernel at a time on one stream
    if (ctx->profiling || opts.debug || opts.singleLane) lanes = 1;
    lanes = std::max(1, std::min(std::min(lanes, LVN_MAX_LANES), n));
    streams = lanes == 1 ? 1 : std::max(1, std::min(std::min(streams, LVN_MAX_STREAMS), lanes));
}

static double host_now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// every stream a lane ran on, and the copy stream, rejoin the context's stream
static int join_lanes(lvn_meshgen *ctx, int numStreams, bool copies, bool hostSync = true)
{
    for (int r = 1; r < numStreams; r++) {
        CU(cudaEventRecord(ctx->evJoin[r], ctx->laneStreams[r]));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->evJoin[r], 0));
    }
    if (copies) {
        CU(cudaEventRecord(ctx->evCopy, ctx->copyStream));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->evCopy, 0));
        CU(cudaEventRecord(ctx->evPubJoin, ctx->pubStream));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->evPubJoin, 0));
    }
    if (hostSync) CU(cudaStreamSynchronize(ctx->stream));
    return LVN_SUCCESS;
}

// One pass of the path over n chunks.  Results stay in the context's arenas; with `out` the
// mesh / seam arenas of every lane are also copied into the caller's host arenas while the
// following lanes are still computing.
static int validate_pending_tables(lvn_meshgen *ctx, bool *rebuilt);
constexpr int LVN_RETRY_TABLES = 1;   // internal: a hash table this batch read had to be rehashed, run the batch again

static int run_batch_once(lvn_meshgen *ctx, int n, const int32_t *chunkMinSize, const BatchOpts &opts, const HostOut *out);

static int run_batch(lvn_meshgen *ctx, int n, const int32_t *chunkMinSize, const BatchOpts &opts,
                     const HostOut *out = nullptr)
{
    int rc = run_batch_once(ctx, n, chunkMinSize, opts, out);
    if (rc == LVN_RETRY_TABLES) rc = run_batch_once(ctx, n, chunkMinSize, opts, out);
    return rc == LVN_RETRY_TABLES ? LVN_CL_ERROR : rc;
}

static int run_batch_once(lvn_meshgen *ctx, int n, const int32_t *chunkMinSize, const BatchOpts &opts, const HostOut *out)
{
    if (!g.initialised) return LVN_ERR_NOT_INITIALISED;
    if (!ctx || n < 0 || (n > 0 && !chunkMinSize)) return LVN_ERR_INVALID_VALUE;
    const Dims &d = ctx->dims;
    const size_t FF = (size_t)d.F * d.F, HH = (size_t)d.H * d.H, VV = (size_t)d.V * d.V, F3 = FF * d.F;
    cudaStream_t st = ctx->stream;
    const double hostStart = ctx->trace ? host_now_us() : 0.0;   // LVN_TRACE: where the host thread's time goes
    double hostWaited = 0.0;
    ctx->lastN = n;
    ctx->numLanes = 0;
    ctx->hostLayout = false;
    memset(&ctx->lastCounters, 0, sizeof(ctx->lastCounters));
    if (n == 0) return LVN_SUCCESS;

    // ---- lanes: lane k takes the caller's chunks k, k + S, k + 2S, ... (neighbouring chunks of a
    //      clipmap ring or a sweep do similar work, so strided lanes are balanced) ----
    int S = 1, R = 1;
    choose_pipeline(ctx, n, opts, out != nullptr, S, R);
    ctx->numLanes = S;
    ctx->numStreams = R;
    ctx->perm.resize(n);
    {
        int p = 0;
        for (int k = 0; k < S; k++) {
            ctx->laneFirst[k] = p;
            for (int i = k; i < n; i += S) ctx->perm[p++] = i;
        }
        ctx->laneFirst[S] = p;
    }

    // ---- descriptors (internal, lane-major order), column-set dedupe, cached fields ----
    // the batch head = the chunk descriptors, then (16-byte aligned) the column sets' origins and the
    // initial keys of their height ranges: one pinned staging block, one upload
    const size_t headColOffset = ((size_t)n * sizeof(ChunkDesc) + 15) & ~(size_t)15;
    const size_t headDescs = (headColOffset + (size_t)n * (sizeof(int4) + 2 * sizeof(int)) + sizeof(ChunkDesc) - 1) / sizeof(ChunkDesc);
    LV(ctx->h_descs.reserve(headDescs));
    LV(ctx->h_colOrigins.reserve(n));
    LV(ctx->h_hdrs.reserve(n + S));
    std::map<std::tuple<int, int, int>, int> colSets;
    int numColSets = 0, numTmpFields = 0;
    std::vector<int> tmpFieldChunk;
    for (int p = 0; p < n; p++) {
        const int32_t *ms = &chunkMinSize[4 * ctx->perm[p]];
        ChunkDesc &cd = ctx->h_descs.p[p];
        LV(fill_desc(ctx, ms, cd));
        auto it = opts.ignoreFieldCache ? ctx->fields.end() : ctx->fields.find(make_key(ms, cd.size));
        if (it != ctx->fields.end()) {
            const FieldEntry &fe = it->second;
            cd.source = SRC_FIELD;
            cd.edgeMode = EDGES_CACHED;
            cd.field = fe.d_field;
            cd.cachedNumEdges = fe.numEdges;
            cd.cachedKeys = fe.d_keys;
            cd.cachedInfo = fe.d_info;
            cd.cuckooTable = fe.d_table;
            cd.cuckooPrime = fe.prime;
            memcpy(cd.cuckooParams, fe.params, sizeof(fe.params));
        } else if (g.densityKind == 0) {
            cd.source = SRC_HEIGHTS;
            cd.edgeMode = EDGES_FRESH;
            auto key = std::make_tuple(cd.ox, cd.oz, cd.scale);
            auto cs = colSets.find(key);
            if (cs == colSets.end()) {
                ctx->h_colOrigins.p[numColSets] = make_int4(cd.ox, cd.oz, cd.scale, 0);
                cs = colSets.emplace(key, numColSets++).first;
            }
            cd.colSet = cs->second;
        } else {
            cd.source = SRC_FIELD;
            cd.edgeMode = EDGES_FRESH;
            tmpFieldChunk.push_back(p);
            numTmpFields++;
        }
    }

    // ---- workspace ----
    LV(ctx->d_descs.reserve(headDescs));
    LV(ctx->d_hdrs.reserve(n + S));
    LV(ctx->d_heights.reserve(std::max<size_t>((size_t)numColSets * FF, 1)));
    LV(ctx->d_bitsLo.reserve(n * FF));
    LV(ctx->d_bitsHi.reserve(n * FF));
    LV(ctx->d_rowE.reserve(n * HH));
    LV(ctx->d_rowN.reserve(n * VV));
    LV(ctx->d_rowQ.reserve(n * VV));
    LV(ctx->d_rowS.reserve(n * VV));
    LV(ctx->d_slab.reserve((size_t)n * LVN_MAX_LAYERS));
    LV(ctx->d_slabEy.reserve((size_t)n * LVN_MAX_LAYERS));
    LV(ctx->d_candidates.reserve(n));
    if ((size_t)n > ctx->d_ticket.cap) {
        LV(ctx->d_ticket.reserve(n));
        CU(cudaMemsetAsync(ctx->d_ticket.p, 0, ctx->d_ticket.cap * sizeof(unsigned int), st));   // k_rows leaves it zero
    }
    if (numTmpFields) {
        LV(ctx->d_tmpFields.reserve((size_t)numTmpFields * F3));
        // the density values behind the materials, for k_hermite (4 F^3 bytes per chunk: up to 8 GB of them)
        const bool keepDensity = !ctx->noLatticeDensity && (size_t)numTmpFields * F3 * sizeof(float) <= ((size_t)8 << 30);
        if (keepDensity) LV(ctx->d_tmpDensity.reserve((size_t)numTmpFields * F3));
        for (int k = 0; k < numTmpFields; k++) {
            ctx->h_descs.p[tmpFieldChunk[k]].field = ctx->d_tmpFields.p + (size_t)k * F3;
            ctx->h_descs.p[tmpFieldChunk[k]].latticeDensity = keepDensity ? ctx->d_tmpDensity.p + (size_t)k * F3 : nullptr;
        }
    }
    // first-guess arena sizes; a batch that needs more reports it in the counters and is re-run
    LV(ctx->d_edgeKeys.reserve(std::max<size_t>((size_t)n * 2048, 1u << 16)));
    LV(ctx->d_edgeInfo.reserve(ctx->d_edgeKeys.cap));
    LV(ctx->d_xzList.reserve(ctx->d_edgeKeys.cap));
    LV(ctx->d_vertices.reserve(std::max<size_t>((size_t)n * 2048, 1u << 16)));
    LV(ctx->d_qef.reserve(ctx->d_vertices.cap * 4));
    LV(ctx->d_tris.reserve(ctx->d_vertices.cap * 6 * 2));
    LV(ctx->d_seams.reserve(std::max<size_t>((size_t)n * 512, 1u << 14)));

    {
        char *hb = (char *)ctx->h_descs.p + headColOffset;
        memcpy(hb, ctx->h_colOrigins.p, (size_t)numColSets * sizeof(int4));
        int *hMin = (int *)(hb + (size_t)numColSets * sizeof(int4)), *hMax = hMin + numColSets;
        for (int k = 0; k < numColSets; k++) { hMin[k] = 0x7f7f7f7f; hMax[k] = (int)0x80808080u; }   // ordered keys of +3.4e38 / -3.4e38
    }
    const double hostPrepDone = ctx->trace ? host_now_us() : 0.0;
    {   // one upload instead of two and two memsets (measured against a kernel that reads the mapped staging
        // block itself, profiles/r01s_notes.md: the copy is as fast or faster)
        const size_t headBytes = headColOffset + (size_t)numColSets * (sizeof(int4) + 2 * sizeof(int));
        CU(cudaMemcpyAsync(ctx->d_descs.p, ctx->h_descs.p, headBytes, cudaMemcpyHostToDevice, st));
    }
    const int4 *d_colOrigins = (const int4 *)((char *)ctx->d_descs.p + headColOffset);
    int *d_colMin = (int *)(d_colOrigins + numColSets), *d_colMax = d_colMin + numColSets;

    const DensityParams dp = density_params();
    ChunkScratch ws;
    ws.bitsLo = ctx->d_bitsLo.p; ws.bitsHi = ctx->d_bitsHi.p;
    ws.rowE = ctx->d_rowE.p; ws.rowN = ctx->d_rowN.p; ws.rowQ = ctx->d_rowQ.p; ws.rowS = ctx->d_rowS.p;
    ws.layer = ctx->d_slab.p; ws.layerEy = ctx->d_slabEy.p; ws.ticket = ctx->d_ticket.p;

    // ---- S1 for the whole batch: the column sets are shared between the lanes ----
    if (numColSets) {
        // stage timing: the interval's first event would be stamped by the copy engine that delivered the batch
        // head, and the hand-over to the compute engine (~30 us) would be booked on k_columns (14 us under
        // ncu); a one-thread kernel in front makes the compute engine stamp it
        if (ctx->profiling) launch_publish(ctx->d_hdrs.p, ctx->d_hdrs.p, 0, st, true);
        StageTimer t(ctx, LVN_STAGE_COLUMNS, 1);
        launch_columns(dp, d, d_colOrigins, numColSets, ctx->d_heights.p, d_colMin, d_colMax, st);
        ctx->stats.terrainEvals += (int64_t)numColSets * (int64_t)FF;
    }
    if (numTmpFields) {
        // 3-D density: materialise the u8 field (descs of the other chunks are skipped by pointer)
        StageTimer t(ctx, LVN_STAGE_FIELD, 1);
        std::vector<ChunkDesc> sub(numTmpFields);
        std::vector<uint8_t *> subPtrs(numTmpFields);
        for (int k = 0; k < numTmpFields; k++) { sub[k] = ctx->h_descs.p[tmpFieldChunk[k]]; subPtrs[k] = (uint8_t *)sub[k].field; }
        DevBuf<ChunkDesc> dsub; DevBuf<uint8_t *> dptr;
        LV(dsub.reserve(numTmpFields)); LV(dptr.reserve(numTmpFields));
        CU(cudaMemcpyAsync(dsub.p, sub.data(), numTmpFields * sizeof(ChunkDesc), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(dptr.p, subPtrs.data(), numTmpFields * sizeof(uint8_t *), cudaMemcpyHostToDevice, st));
        launch_field_density(dp, d, dsub.p, numTmpFields, dptr.p, st);
        CU(cudaStreamSynchronize(st));
        dsub.release(); dptr.release();
    }

    // Header array: lane k's chunk headers, then one slot holding the lane's ArenaCounters.
    // Kernels of lane k index it with the chunk's internal index through a pointer advanced by k
    // slots.  The host's copy is a mapped pinned mirror that the kernels write directly (k_rows
    // the headers, k_leaves the final counters): a small copy-engine transfer at the end of a
    // lane would queue behind the bulk mesh copies of the lanes before it.
    static_assert(sizeof(ArenaCounters) <= sizeof(ChunkHdr), "the counters live in a header slot");
    ChunkHdr *h_hdrs_dev = ctx->h_hdrs.dev();
    if (!h_hdrs_dev) { g_lastCudaError = "pinned header mirror is not mapped"; return LVN_ERR_CUDA; }
    auto lane_counters_dev = [&](int k) { return (ArenaCounters *)(ctx->d_hdrs.p + ctx->laneFirst[k + 1] + k); };
    auto lane_counters_host = [&](int k) { return (const ArenaCounters *)(ctx->h_hdrs.p + ctx->laneFirst[k + 1] + k); };

    // counts first: nothing that needs the whole batch on the host may be pending
    const bool countsFirst = opts.countsFirst && !opts.debug && !ctx->profiling && !ctx->trace && ctx->pendingTables.empty();
    for (int attempt = 0; attempt < 3; attempt++) {
        if (opts.debug) {
            LV(ctx->d_dbgCodes.reserve(ctx->d_vertices.cap)); LV(ctx->d_dbgMasks.reserve(ctx->d_vertices.cap));
            LV(ctx->d_dbgMats.reserve(ctx->d_vertices.cap)); LV(ctx->d_dbgQefs.reserve(ctx->d_vertices.cap * 16));
            LV(ctx->d_dbgPos.reserve(ctx->d_vertices.cap)); LV(ctx->d_dbgNrm.reserve(ctx->d_vertices.cap));
        }
        // every lane owns an equal slice of every arena and of the tile directories
        ArenaCaps caps;
        caps.edges = (unsigned int)std::min<size_t>(ctx->d_edgeKeys.cap / S, 0x7fffffffu);
        caps.nodes = (unsigned int)std::min<size_t>(ctx->d_vertices.cap / S, 0x7fffffffu);
        caps.quads = (unsigned int)std::min<size_t>(ctx->d_tris.cap / 6 / S, 0x7fffffffu);
        caps.seams = (unsigned int)std::min<size_t>(ctx->d_seams.cap / S, 0x7fffffffu);
        const int maxLaneChunks = ctx->laneFirst[1] - ctx->laneFirst[0];
        const unsigned int tileCap = std::max(caps.edges, caps.nodes) / LVN_TILE + (unsigned int)maxLaneChunks + 1u;
        LV(ctx->d_edgeTiles.reserve((size_t)tileCap * S));
        LV(ctx->d_nodeTiles.reserve((size_t)tileCap * S));
        CU(cudaMemsetAsync(ctx->d_hdrs.p, 0, (size_t)(n + S) * sizeof(ChunkHdr), st));   // zero counters
        if (R > 1) {
            CU(cudaEventRecord(ctx->evFork, st));
            for (int r = 1; r < R; r++) CU(cudaStreamWaitEvent(ctx->laneStreams[r], ctx->evFork, 0));
        }
        if (ctx->trace) {
            while (ctx->traceEv.size() < (size_t)(1 + 5 * S)) { cudaEvent_t e; cudaEventCreate(&e); ctx->traceEv.push_back(e); }
            cudaEventRecord(ctx->traceEv[0], st);
        }
#define LVN_TRACE_EV(i) do { if (ctx->trace) cudaEventRecord(ctx->traceEv[1 + 5 * k + (i)], ls); } while (0)
        auto enqueue_lane = [&](int k) -> int {
            cudaStream_t ls = (k % R) == 0 ? st : ctx->laneStreams[k % R];
            const int first = ctx->laneFirst[k], cnt = ctx->laneFirst[k + 1] - first;
            ChunkHdr *hdrs = ctx->d_hdrs.p + k;
            LVN_TRACE_EV(0);
            LaneArenas lane;
            lane.caps = caps;
            lane.base.edges = caps.edges * k; lane.base.nodes = caps.nodes * k;
            lane.base.quads = caps.quads * k; lane.base.seams = caps.seams * k;
            lane.ctr = lane_counters_dev(k);
            lane.edgeTiles = ctx->d_edgeTiles.p + (size_t)tileCap * k;
            lane.nodeTiles = ctx->d_nodeTiles.p + (size_t)tileCap * k;
            lane.tileCap = tileCap;
            ctx->laneBase[k] = lane.base;
            {
                StageTimer t(ctx, LVN_STAGE_CLASSIFY, 2);
                launch_rows(d, ctx->d_descs.p, first, cnt, ctx->d_heights.p, d_colMin, d_colMax, hdrs, nullptr,
                            ws, lane, ctx->d_candidates.p, ls);
            }
            // The lane's headers and counters are final once k_rows has run.  On the host path they
            // are published right away from a side stream, so that the host can size and queue the
            // lane's copies (behind evLane) long before the lane's last kernel ends: the copy engine
            // then starts the moment the lane is done, with no host round trip in between.
            const bool earlyPublish = out != nullptr || countsFirst;   // (the host path always publishes early)
            if (earlyPublish) {
                CU(cudaEventRecord(ctx->evRows[k], ls));
                CU(cudaStreamWaitEvent(ctx->pubStream, ctx->evRows[k], 0));
                launch_publish(ctx->d_hdrs.p + first + k, h_hdrs_dev + first + k, cnt + 1, ctx->pubStream);
                CU(cudaEventRecord(ctx->evPub[k], ctx->pubStream));
                ctx->stats.launches[LVN_STAGE_CLASSIFY] += 1;   // k_publish
            }
            LVN_TRACE_EV(1);
            bool anyFresh = false;   // a lane of cached edge lists only (re-meshing edited chunks) has nothing for S3
            for (int p = first; p < first + cnt && !anyFresh; p++) anyFresh = ctx->h_descs.p[p].edgeMode == EDGES_FRESH;
            if (anyFresh) {
                StageTimer t(ctx, LVN_STAGE_HERMITE, 1);
                launch_hermite(dp, d, ctx->d_descs.p, hdrs, ws, lane, ctx->d_heights.p, ctx->d_edgeKeys.p,
                               ctx->d_edgeInfo.p, ctx->d_xzList.p, ls);
            }
            LVN_TRACE_EV(2);
            {
                NodeDebug dbg = {};
                if (opts.debug) {
                    dbg.codes = ctx->d_dbgCodes.p; dbg.edgeMasks = ctx->d_dbgMasks.p; dbg.matWords = ctx->d_dbgMats.p;
                    dbg.qefs = ctx->d_dbgQefs.p; dbg.positions = ctx->d_dbgPos.p; dbg.normals = ctx->d_dbgNrm.p;
                }
                {
                    StageTimer t(ctx, LVN_STAGE_LEAVES, 1);
                    launch_leaves(dp, d, ctx->d_descs.p, hdrs, ws, lane, nullptr, ctx->d_edgeInfo.p, ctx->d_qef.p,
                                  ctx->d_vertices.p, ctx->d_tris.p, ctx->d_seams.p, dbg, ls);
                }
                {
                    StageTimer t(ctx, LVN_STAGE_SOLVE, 1);
                    launch_solve(ctx->d_descs.p, lane, ctx->d_qef.p, ctx->d_vertices.p, ctx->d_seams.p, dbg.positions, ls);
                }
            }
            LVN_TRACE_EV(3);
            if (earlyPublish) CU(cudaEventRecord(ctx->evLane[k], ls));
            else {
                launch_publish(ctx->d_hdrs.p + first + k, h_hdrs_dev + first + k, cnt + 1, ls);
                ctx->stats.launches[LVN_STAGE_CLASSIFY] += 1;
            }
            LVN_TRACE_EV(4);
            return LVN_SUCCESS;
        };

        bool overflow = false, hostFull = false;
        bool joinInStreamOrder = false;
        if (!out) {
            for (int k = 0; k < S; k++) LV(enqueue_lane(k));
            if (countsFirst) {
                // every lane's headers and counters are in the mirror once its k_publish has run; without an
                // overflow nothing else of the batch is needed on the host
                joinInStreamOrder = true;
                for (int k = 0; k < S; k++) {
                    CU(cudaEventSynchronize(ctx->evPub[k]));
                    if (lane_counters_host(k)->overflow) joinInStreamOrder = false;
                }
            }
        } else {
            // host path: keep two lanes queued ahead of the one being drained over the copy
            // engine, so that neither the SMs nor the PCIe link wait for the host
            int64_t hv = 0, ht = 0, hs = 0;
            int issued = 0;
            for (int k = 0; k < S; k++) {
                while (issued < S && issued < k + ctx->lookahead) LV(enqueue_lane(issued++));
                if (overflow || hostFull) continue;   // keep issuing: the retry needs every lane's counts
                const double w0 = ctx->trace ? host_now_us() : 0.0;
                CU(cudaEventSynchronize(ctx->evPub[k]));
                if (ctx->trace) hostWaited += host_now_us() - w0;
                const ArenaCounters c = *lane_counters_host(k);
                if (c.overflow) { overflow = true; continue; }
                ctx->hostBase[k][0] = hv; ctx->hostBase[k][1] = ht; ctx->hostBase[k][2] = hs;
                if (hv + c.nodes > out->vertexCapacity || ht + 2 * (int64_t)c.quads > out->triangleCapacity ||
                    hs + c.seams > out->seamCapacity) { hostFull = true; continue; }
                const ArenaCaps &b = ctx->laneBase[k];
                // the lane's three arenas in one batched copy: 12 separate memcpys of a 4-lane batch
                // cost ~55 us more than the same bytes in one (profiles/micro/copy_granularity.cu)
                cudaStream_t cs = ctx->copyStream;
                CU(cudaStreamWaitEvent(cs, ctx->evLane[k], 0));   // the lane's last kernel
                void *dst[3], *src[3];
                size_t len[3];
                size_t m = 0;
                if (c.nodes) { dst[m] = out->vertices + hv; src[m] = ctx->d_vertices.p + b.nodes; len[m++] = (size_t)c.nodes * sizeof(lvn_mesh_vertex); }
                if (c.quads) { dst[m] = out->triangles + ht; src[m] = ctx->d_tris.p + (size_t)b.quads * 6; len[m++] = (size_t)c.quads * 6 * sizeof(int); }
                if (c.seams) { dst[m] = out->seams + hs; src[m] = ctx->d_seams.p + b.seams; len[m++] = (size_t)c.seams * sizeof(lvn_seam_node_info); }
#if CUDART_VERSION >= 12080 && CUDART_VERSION < 13000   // the 12.8 / 12.9 signature (size_t *failIdx); plain copies otherwise
                // (small single-lane batches usually land in pageable memory, where the batched call
                // costs ~0.6 ms more than three plain copies: measured with config 3's re-meshes)
                if (m > 1 && S > 1) {
                    cudaMemcpyAttributes attr = {};
                    attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
                    size_t attrIdx = 0, failIdx = 0;
                    if (cudaMemcpyBatchAsync(dst, src, len, m, &attr, &attrIdx, 1, &failIdx, cs) == cudaSuccess) m = 0;
                    else cudaGetLastError();   // older driver: plain copies below
                }
#endif
                for (size_t i = 0; i < m; i++) CU(cudaMemcpyAsync(dst[i], src[i], len[i], cudaMemcpyDeviceToHost, cs));
                hv += c.nodes; ht += 2 * (int64_t)c.quads; hs += c.seams;
                if (ctx->trace) {
                    while (ctx->traceCopyEv.size() < (size_t)S) { cudaEvent_t e; cudaEventCreate(&e); ctx->traceCopyEv.push_back(e); }
                    cudaEventRecord(ctx->traceCopyEv[k], cs);
                }
            }
            // every lane's counts are in and its copies are queued: a counts-first caller gets them now
            joinInStreamOrder = countsFirst && !overflow && !hostFull;
        }
#undef LVN_TRACE_EV
        const double j0 = ctx->trace ? host_now_us() : 0.0;
        LV(join_lanes(ctx, R, out != nullptr || countsFirst, !joinInStreamOrder));
        if (ctx->trace) {
            hostWaited += host_now_us() - j0;
            fprintf(stderr, "[lvn trace] host thread: %.0f us in the call, %.0f us of them waiting for the GPU (the rest: building the batch, driver calls); "
                            "%.0f us before the first enqueue, everything enqueued after %.0f us\n",
                    host_now_us() - hostStart, hostWaited, hostPrepDone - hostStart, j0 - hostStart);
        }
        CU(cudaGetLastError());
        collect_stage_times(ctx);
        {   // hash tables queued by the edit in front of this batch (apply_csg_items): their insert flags are in now
            bool rebuilt = false;
            LV(validate_pending_tables(ctx, &rebuilt));
            if (rebuilt) return LVN_RETRY_TABLES;
        }
        if (ctx->trace) {
            fprintf(stderr, "[lvn trace] n=%d lanes=%d streams=%d  (us since fork: start rows|hermite|leaves|d2h|end)\n", n, S, R);
            for (int k = 0; k < S; k++) {
                float t[5];
                for (int i = 0; i < 5; i++) cudaEventElapsedTime(&t[i], ctx->traceEv[0], ctx->traceEv[1 + 5 * k + i]);
                float tc = 0.f;
                if (out && !overflow && !hostFull && ctx->traceCopyEv.size() >= (size_t)S)
                    cudaEventElapsedTime(&tc, ctx->traceEv[0], ctx->traceCopyEv[k]);
                fprintf(stderr, "[lvn trace]  lane %2d stream %d: %7.1f %7.1f %7.1f %7.1f %7.1f  copied %7.1f\n", k, k % R,
                        t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, t[3] * 1e3, t[4] * 1e3, tc * 1e3);
            }
        }

        ArenaCounters tot = {}, mx = {};
        for (int k = 0; k < S; k++) {
            const ArenaCounters c = *lane_counters_host(k);
            ctx->laneCounters[k] = c;
            tot.edges += c.edges; tot.nodes += c.nodes; tot.quads += c.quads; tot.seams += c.seams;
            tot.nonEmpty += c.nonEmpty; tot.overflow |= c.overflow;
            mx.edges = std::max(mx.edges, c.edges); mx.nodes = std::max(mx.nodes, c.nodes);
            mx.quads = std::max(mx.quads, c.quads); mx.seams = std::max(mx.seams, c.seams);
        }
        if (!tot.overflow) {
            ctx->lastCounters = tot;
            ctx->hintChunks = n; ctx->hintNodes = tot.nodes; ctx->hintNonEmpty = tot.nonEmpty;
            int64_t ey = 0;
            for (int k = 0; k < S; k++)
                for (int p = ctx->laneFirst[k]; p < ctx->laneFirst[k + 1]; p++) ey += ctx->h_hdrs.p[p + k].Ey;
            ctx->stats.edges += tot.edges; ctx->stats.edgesY += ey; ctx->stats.nodes += tot.nodes;
            ctx->stats.triangles += 2 * (int64_t)tot.quads; ctx->stats.seamNodes += tot.seams;
            ctx->stats.chunks += n; ctx->stats.nonEmptyChunks += tot.nonEmpty;
            if (g.densityKind == 0) ctx->stats.terrainEvals += 4 * ey + 19 * ((int64_t)tot.edges - ey);
            if (out) {
                if (hostFull) return LVN_ERR_CAPACITY;
                ctx->hostLayout = true;
            }
            return LVN_SUCCESS;
        }
        // grow every lane's slice to what the fullest lane asked for (+12%) and run again: the
        // counters keep counting past the capacity, so one retry suffices
        LV(ctx->d_edgeKeys.reserve((size_t)S * ((size_t)mx.edges + mx.edges / 8 + 1024)));
        LV(ctx->d_edgeInfo.reserve(ctx->d_edgeKeys.cap));
        LV(ctx->d_xzList.reserve(ctx->d_edgeKeys.cap));
        LV(ctx->d_vertices.reserve((size_t)S * ((size_t)mx.nodes + mx.nodes / 8 + 1024)));
        LV(ctx->d_qef.reserve(ctx->d_vertices.cap * 4));
        LV(ctx->d_tris.reserve((size_t)S * ((size_t)mx.quads + mx.quads / 8 + 1024) * 6));
        LV(ctx->d_seams.reserve((size_t)S * ((size_t)mx.seams + mx.seams / 8 + 1024)));
    }
    return LVN_ERR_CAPACITY;
}

// per-chunk results in the caller's chunk order; offsets address the device arenas, or the host
// arenas after a host-path batch
static void fill_results(lvn_meshgen *ctx, int n, lvn_chunk_result *results)
{
    for (int k = 0; k < ctx->numLanes; k++) {
        const ArenaCaps &b = ctx->laneBase[k];
        for (int p = ctx->laneFirst[k]; p < ctx->laneFirst[k + 1]; p++) {
            const ChunkHdr &h = ctx->h_hdrs.p[p + k];   // lane k's headers sit k counter slots further
            lvn_chunk_result &r = results[ctx->perm[p]];
            r.numEdges = h.E;
            // a chunk whose octree yields no quad exports an empty mesh buffer, but still its seam
            // nodes (GenerateMeshFromOctree returns before filling the buffer, compute_octree.cpp:227-232)
            r.numVertices = h.Q > 0 ? h.N : 0;
            r.numTriangles = 2 * h.Q;
            r.numSeamNodes = h.S;
            if (ctx->hostLayout) {
                r.vertexOffset = (int32_t)(ctx->hostBase[k][0] + (h.N ? h.nodeBase - (int)b.nodes : 0));
                r.triangleOffset = (int32_t)(ctx->hostBase[k][1] + 2 * (h.Q ? h.quadBase - (int)b.quads : 0));
                r.seamOffset = (int32_t)(ctx->hostBase[k][2] + (h.S ? h.seamBase - (int)b.seams : 0));
            } else {
                r.vertexOffset = h.nodeBase;
                r.triangleOffset = 2 * h.quadBase;
                r.seamOffset = h.seamBase;
            }
            r.status = h.status;
        }
    }
    (void)n;
}

static int load_density_fields(lvn_meshgen *ctx, int n, const int32_t *minSize, bool forceEntry, bool withTables,
                               std::vector<FieldEntry *> &out);

// LoadDensityField's replay (compute_density_field.cpp:235-274) for the batch entry points: every
// generateChunkMesh of the reference passes through it, so a chunk that stored operations
// overlap is meshed from the edited field, whichever entry point asks.  Without stored
// operations this is one test.
static int replay_stored_ops(lvn_meshgen *ctx, int n, const int32_t *chunkMinSize)
{
    if (!g.initialised || !ctx || n <= 0 || !chunkMinSize || g.storedOps.empty()) return LVN_SUCCESS;
    std::vector<FieldEntry *> entries;
    return load_density_fields(ctx, n, chunkMinSize, false, true, entries);
}

extern "C" int lvn_meshgen_generate_batch_device(lvn_meshgen *ctx, int nChunks, const int32_t *chunkMinSize,
                                                 lvn_chunk_result *results, lvn_batch_device_view *view)
{
    BatchOpts opts;
    LV(replay_stored_ops(ctx, nChunks, chunkMinSize));
    LV(run_batch(ctx, nChunks, chunkMinSize, opts));
    if (results) fill_results(ctx, nChunks, results);
    if (view) {
        view->vertices = ctx->d_vertices.p;
        view->triangles = (const lvn_mesh_triangle *)ctx->d_tris.p;
        view->seamNodes = ctx->d_seams.p;
        view->totalVertices = ctx->lastCounters.nodes;
        view->totalTriangles = 2 * (int64_t)ctx->lastCounters.quads;
        view->totalSeamNodes = ctx->lastCounters.seams;
        view->totalEdges = ctx->lastCounters.edges;
        view->nonEmptyChunks = (int32_t)ctx->lastCounters.nonEmpty;
    }
    return LVN_SUCCESS;
}

// The device batch for a caller that has something to do with the counts while the meshes are still being made
// (the count gather of a sharded sweep, sizing a consumer's buffers): returns once every chunk's counts and
// arena offsets are final -- they are after the classify kernels -- with the rest of the batch queued on the
// context's stream.  The arenas are complete in stream order on that stream, or after lvn_meshgen_wait.
extern "C" int lvn_meshgen_generate_batch_device_async(lvn_meshgen *ctx, int nChunks, const int32_t *chunkMinSize,
                                                       lvn_chunk_result *results, lvn_batch_device_view *view)
{
    BatchOpts opts;
    opts.countsFirst = true;
    LV(replay_stored_ops(ctx, nChunks, chunkMinSize));
    LV(run_batch(ctx, nChunks, chunkMinSize, opts));
    if (results) fill_results(ctx, nChunks, results);
    if (view) {
        view->vertices = ctx->d_vertices.p;
        view->triangles = (const lvn_mesh_triangle *)ctx->d_tris.p;
        view->seamNodes = ctx->d_seams.p;
        view->totalVertices = ctx->lastCounters.nodes;
        view->totalTriangles = 2 * (int64_t)ctx->lastCounters.quads;
        view->totalSeamNodes = ctx->lastCounters.seams;
        view->totalEdges = ctx->lastCounters.edges;
        view->nonEmptyChunks = (int32_t)ctx->lastCounters.nonEmpty;
    }
    return LVN_SUCCESS;
}

extern "C" int lvn_meshgen_wait(lvn_meshgen *ctx) { return lvn::meshgen_wait(ctx); }

// lvn_meshgen_generate_batch for the same kind of caller: returns once every lane's counts are published and its
// copies are queued; `results` are final, the host arenas are complete after lvn_meshgen_wait.
extern "C" int lvn_meshgen_generate_batch_async(lvn_meshgen *ctx, int nChunks, const int32_t *chunkMinSize,
                                                lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                                                lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                                                lvn_seam_node_info *seamNodes, int64_t seamCapacity,
                                                lvn_chunk_result *results)
{
    if (!results) return LVN_ERR_INVALID_VALUE;
    BatchOpts opts;
    opts.countsFirst = true;
    HostOut out = {vertices, vertexCapacity, triangles, triangleCapacity, seamNodes, seamCapacity};
    LV(replay_stored_ops(ctx, nChunks, chunkMinSize));
    const int rc = run_batch(ctx, nChunks, chunkMinSize, opts, &out);
    if (rc == LVN_SUCCESS || rc == LVN_ERR_CAPACITY) {
        if (ctx && ctx->numLanes) fill_results(ctx, nChunks, results);
    }
    return rc;
}

extern "C" int lvn_meshgen_generate_batch(lvn_meshgen *ctx, int nChunks, const int32_t *chunkMinSize,
                                          lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                                          lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                                          lvn_seam_node_info *seamNodes, int64_t seamCapacity,
                                          lvn_chunk_result *results)
{
    if (!results) return LVN_ERR_INVALID_VALUE;
    BatchOpts opts;
    HostOut out = {vertices, vertexCapacity, triangles, triangleCapacity, seamNodes, seamCapacity};
    LV(replay_stored_ops(ctx, nChunks, chunkMinSize));
    const int rc = run_batch(ctx, nChunks, chunkMinSize, opts, &out);
    if (rc == LVN_SUCCESS || rc == LVN_ERR_CAPACITY) {
        // on LVN_ERR_CAPACITY the counts say what the caller must provide
        if (ctx && ctx->numLanes) fill_results(ctx, nChunks, results);
    }
    return rc;
}

// ConstructClipmapNodeData / ConstructCollisionNodeData (clipmap.cpp:432-504) for many nodes:
// generateChunkMesh, then ngMeshSimplifier with the options the clipmap derives from the node size.
// The meshes stay in HBM between the two; only the simplified meshes cross PCIe -- as MeshVertex
// (`vertices`) and / or in the physics engine's format (`physicsVertices`, AddMeshToWorldImpl).
// deferMeshCopies: return once the two mesh copies are queued (results, counts and seam nodes are
// final by then); the caller overlaps other work and calls lvn::meshgen_wait before reading the meshes.
int lvn::generate_simplified(lvn_meshgen *ctx, int nChunks, const int32_t *chunkMinSize,
                             const lvn_simplify_options *unitOptions,
                             lvn_mesh_vertex *vertices, float *physicsVertices, float physicsScale, int64_t vertexCapacity,
                             lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                             lvn_seam_node_info *seamNodes, int64_t seamCapacity,
                             lvn_chunk_result *results, lvn_simplify_result *simplified, bool deferMeshCopies,
                             uint8_t *hadMesh)
{
    if (!results || !unitOptions) return LVN_ERR_INVALID_VALUE;
    BatchOpts opts;
    LV(replay_stored_ops(ctx, nChunks, chunkMinSize));
    LV(run_batch(ctx, nChunks, chunkMinSize, opts));
    if (nChunks == 0) return LVN_SUCCESS;
    fill_results(ctx, nChunks, results);      // offsets into the device arenas
    cudaStream_t st = ctx->stream;

    std::vector<SimplifyMesh> meshes;
    std::vector<int> chunkOf;
    meshes.reserve(ctx->lastCounters.nonEmpty);
    for (int i = 0; i < nChunks; i++) {
        const lvn_chunk_result &r = results[i];
        if (simplified) simplified[i] = lvn_simplify_result{0, 0, 0, 0};
        if (hadMesh) hadMesh[i] = r.status >= 0 && r.numTriangles > 0;   // renderMesh != null before ngMeshSimplifier (clipmap.cpp:466)
        if (r.status < 0 || r.numTriangles <= 0) continue;
        const int32_t *ms = chunkMinSize + 4 * (size_t)i;
        SimplifyMesh m;
        m.vertexOffset = r.vertexOffset; m.numVertices = r.numVertices;
        m.triangleOffset = r.triangleOffset; m.numTriangles = r.numTriangles;
        // centrePos = vec4(vec3(min) + vec3(size / 2.f), 0); leafSize = LEAF_SIZE_SCALE * (size / CLIPMAP_LEAF_SIZE)
        for (int a = 0; a < 3; a++) m.offset[a] = (float)ms[a] + (float)ms[3] / 2.f;
        m.offset[3] = 0.f;
        const float leafSize = (float)(LVN_LEAF_SIZE_SCALE * (ms[3] / (LVN_LEAF_SIZE_SCALE * 64)));
        m.opt = *unitOptions;
        m.opt.maxError = unitOptions->maxError * leafSize;
        m.opt.maxEdgeSize = unitOptions->maxEdgeSize * leafSize;
        meshes.push_back(m);
        chunkOf.push_back(i);
    }
    const int M = (int)meshes.size();
    const int64_t totalSeams = ctx->lastCounters.seams;
    int2 totals = make_int2(0, 0), early = make_int2(0, 0);
    SimplifySplit split = {};
    if (M > 0) {
        if (!ctx->simpStreamB) {
            CU(cudaStreamCreateWithFlags(&ctx->simpStreamB, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&ctx->evSimpFork, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ctx->evSimpEarly, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ctx->evSimpJoin, cudaEventDisableTiming));
        }
        LV(ctx->d_simpRes.reserve(M));
        LV(ctx->d_packOff.reserve((size_t)M + 2));      // per mesh, [M] the totals, [M + 1] the end of the early region
        if (vertices) LV(ctx->d_packV.reserve(ctx->lastCounters.nodes));
        if (physicsVertices) LV(ctx->d_packP.reserve(ctx->lastCounters.nodes));
        LV(ctx->d_packT.reserve(6 * (size_t)ctx->lastCounters.quads));
        LV(ctx->h_simpRes.reserve(M));
        LV(ctx->h_packOff.reserve((size_t)M + 2));
        split.streamB = ctx->simpStreamB; split.evFork = ctx->evSimpFork; split.evEarly = ctx->evSimpEarly;
        split.d_earlyTotals = ctx->d_packOff.p + M + 1;
        const int rc = simplify_device(M, meshes.data(), ctx->d_vertices.p, ctx->d_tris.p, ctx->d_simpRes.p,
                                       vertices ? ctx->d_packV.p : nullptr, ctx->d_packT.p, ctx->d_packOff.p, ctx->d_packOff.p + M, st,
                                       physicsVertices ? ctx->d_packP.p : nullptr, physicsScale,
                                       (deferMeshCopies && !getenv("LVN_SIMP_SPLIT_DEFER")) ? nullptr : &split);   // (a caller that defers overlaps the download itself)
        if (rc < 0) { g_lastCudaError = simplify_last_error(); return rc; }
        CU(cudaMemcpyAsync(ctx->h_simpRes.p, ctx->d_simpRes.p, sizeof(int4) * M, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ctx->h_packOff.p, ctx->d_packOff.p, sizeof(int2) * ((size_t)M + 1), cudaMemcpyDeviceToHost, st));
    }
    // The seam nodes do not wait for the simplifier: lane by lane (each lane's slice is dense).  With a
    // split launch they ride on the early group's stream, which ends long before the late group.
    const bool splitRun = split.numEarly > 0;
    cudaStream_t sideStream = splitRun ? ctx->simpStreamB : st;
    if (splitRun) CU(cudaMemcpyAsync(ctx->h_packOff.p + M + 1, ctx->d_packOff.p + M + 1, sizeof(int2), cudaMemcpyDeviceToHost, sideStream));
    const bool seamsFit = totalSeams <= seamCapacity && (seamNodes || totalSeams == 0);
    int64_t hs = 0;
    for (int k = 0; k < ctx->numLanes; k++) {
        const unsigned int cnt = ctx->laneCounters[k].seams;
        ctx->hostBase[k][2] = hs;
        if (seamsFit && cnt)
            CU(cudaMemcpyAsync(seamNodes + hs, ctx->d_seams.p + ctx->laneBase[k].seams, sizeof(lvn_seam_node_info) * cnt, cudaMemcpyDeviceToHost, sideStream));
        hs += cnt;
    }
    if (splitRun) {
        // the early region goes out while the largest meshes are still being simplified
        CU(cudaStreamSynchronize(sideStream));
        early = ctx->h_packOff.p[M + 1];
        const bool fits = early.x <= vertexCapacity && early.y <= triangleCapacity && (early.x == 0 || vertices || physicsVertices) &&
                          (early.y == 0 || triangles);
        if (!fits) early = make_int2(0, 0);      // everything is decided (and refused) below, once the totals are known
        if (early.x && vertices) CU(cudaMemcpyAsync(vertices, ctx->d_packV.p, sizeof(lvn_mesh_vertex) * (size_t)early.x, cudaMemcpyDeviceToHost, sideStream));
        if (early.x && physicsVertices) CU(cudaMemcpyAsync(physicsVertices, ctx->d_packP.p, sizeof(float4) * (size_t)early.x, cudaMemcpyDeviceToHost, sideStream));
        if (early.y) CU(cudaMemcpyAsync(triangles, ctx->d_packT.p, 12 * (size_t)early.y, cudaMemcpyDeviceToHost, sideStream));
        CU(cudaEventRecord(ctx->evSimpJoin, sideStream));
        CU(cudaStreamWaitEvent(st, ctx->evSimpJoin, 0));     // whoever waits for the context's stream waits for these copies too
    }
    CU(cudaStreamSynchronize(st));
    if (M > 0) totals = ctx->h_packOff.p[M];
    for (int k = 0; k < ctx->numLanes; k++)
        for (int p = ctx->laneFirst[k]; p < ctx->laneFirst[k + 1]; p++) {
            const ChunkHdr &h = ctx->h_hdrs.p[p + k];
            results[ctx->perm[p]].seamOffset = (int32_t)(ctx->hostBase[k][2] + (h.S ? h.seamBase - (int)ctx->laneBase[k].seams : 0));
        }
    for (int i = 0; i < nChunks; i++) { results[i].vertexOffset = 0; results[i].triangleOffset = 0; }
    for (int m = 0; m < M; m++) {
        lvn_chunk_result &r = results[chunkOf[m]];
        const int4 sr = ctx->h_simpRes.p[m];
        r.numVertices = sr.x; r.numTriangles = sr.y;
        r.vertexOffset = ctx->h_packOff.p[m].x; r.triangleOffset = ctx->h_packOff.p[m].y;
        if (simplified) simplified[chunkOf[m]] = lvn_simplify_result{sr.x, sr.y, sr.z, sr.w};
    }
    // on LVN_ERR_CAPACITY the counts say what the caller must provide
    if (!seamsFit || totals.x > vertexCapacity || totals.y > triangleCapacity ||
        (totals.x > 0 && !vertices && !physicsVertices) || (totals.y > 0 && !triangles)) return LVN_ERR_CAPACITY;
    // what has not left yet: everything, or the late group's region behind the early one
    const size_t v0 = (size_t)early.x, t0 = (size_t)early.y, nv = (size_t)totals.x - v0, nt = (size_t)totals.y - t0;
    if (nv && vertices) CU(cudaMemcpyAsync(vertices + v0, ctx->d_packV.p + v0, sizeof(lvn_mesh_vertex) * nv, cudaMemcpyDeviceToHost, st));
    if (nv && physicsVertices) CU(cudaMemcpyAsync(physicsVertices + 4 * v0, ctx->d_packP.p + v0, sizeof(float4) * nv, cudaMemcpyDeviceToHost, st));
    if (nt) CU(cudaMemcpyAsync(triangles + t0, ctx->d_packT.p + 3 * t0, 12 * nt, cudaMemcpyDeviceToHost, st));
    if (!deferMeshCopies) CU(cudaStreamSynchronize(st));
    return LVN_SUCCESS;
}

int lvn::meshgen_wait(lvn_meshgen *ctx)
{
    if (!ctx) return LVN_ERR_INVALID_VALUE;
    CU(cudaStreamSynchronize(ctx->stream));
    return LVN_SUCCESS;
}

extern "C" int lvn_meshgen_generate_simplified_batch(lvn_meshgen *ctx, int nChunks, const int32_t *chunkMinSize,
                                                     const lvn_simplify_options *unitOptions,
                                                     lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                                                     lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                                                     lvn_seam_node_info *seamNodes, int64_t seamCapacity,
                                                     lvn_chunk_result *results, lvn_simplify_result *simplified)
{
    return generate_simplified(ctx, nChunks, chunkMinSize, unitOptions, vertices, nullptr, 0.f, vertexCapacity, triangles, triangleCapacity,
                               seamNodes, seamCapacity, results, simplified, false, nullptr);
}

// Clipmap::loadCollisionNodes' per-node work (clipmap.cpp:1346-1385): ConstructCollisionNodeData, then the
// conversion Physics_UpdateWorldNodeMainMesh -> AddMeshToWorldImpl makes for Bullet (physics.cpp:549-573)
extern "C" int lvn_meshgen_generate_collision_batch(lvn_meshgen *ctx, int nNodes, const int32_t *nodeMinSize,
                                                    const lvn_simplify_options *unitOptions, float physicsScale,
                                                    float *physicsVertices, lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                                                    int32_t *triangles, int64_t triangleCapacity,
                                                    lvn_seam_node_info *seamNodes, int64_t seamCapacity,
                                                    lvn_chunk_result *results, lvn_simplify_result *simplified)
{
    if (!physicsVertices && !vertices && vertexCapacity > 0) return LVN_ERR_INVALID_VALUE;
    return generate_simplified(ctx, nNodes, nodeMinSize, unitOptions, vertices, physicsVertices, physicsScale, vertexCapacity,
                               (lvn_mesh_triangle *)triangles, triangleCapacity, seamNodes, seamCapacity, results, simplified, false, nullptr);
}

// ---------------------------------------------------------------------------
// cuckoo tables of field entries (compute_cuckoo.cpp:49-138), many at a time
// ---------------------------------------------------------------------------
static void draw_cuckoo_params(unsigned int *params8)
{
    std::uniform_int_distribution<uint32_t> distribution(1u << 15, 1u << 30);   // compute_cuckoo.cpp:47
    for (int i = 0; i < 8; i++) params8[i] = distribution(g.cuckooRng);
}

// Cuckoo_InitialiseTable + Cuckoo_InsertKeys for every entry of the list: all fills and inserts
// are queued, one host wait reads every failure flag, entries whose insertion failed are
// rehashed with fresh parameters (the reference's loop, compute_cuckoo.cpp:89-132).
static int validate_pending_tables(lvn_meshgen *ctx, bool *rebuilt);

// queue Cuckoo_InitialiseTable + Cuckoo_InsertKeys with fresh parameters for every entry of the list; flag i
// (device) becomes non-zero when a key of entry i could not be placed
static int enqueue_table_builds(lvn_meshgen *ctx, const std::vector<FieldEntry *> &todo, unsigned int *d_flags, int attempt,
                                bool flagsClearedByHost = false)
{
    cudaStream_t st = ctx->stream;
    if (!flagsClearedByHost) CU(cudaMemsetAsync(d_flags, 0, todo.size() * sizeof(unsigned int), st));
    for (size_t first = 0; first < todo.size(); first += LVN_TABLE_JOBS) {
        const int m = (int)std::min<size_t>(LVN_TABLE_JOBS, todo.size() - first);
        TableJobs jobs = {};
        for (int i = 0; i < m; i++) {
            FieldEntry *fe = todo[first + i];
            draw_cuckoo_params(fe->params);
            TableJob &j = jobs.job[i];
            j.keys = (const unsigned int *)fe->d_keys;
            j.table = fe->d_table;
            j.failed = d_flags + first + i;
            j.count = (unsigned int)fe->numEdges;
            j.prime = fe->prime;
            memcpy(j.p, fe->params, sizeof(j.p));
            fe->cuckooRetries = attempt;
        }
        launch_table_builds(jobs, m, st);
        ctx->stats.launches[LVN_STAGE_CUCKOO] += 2;
    }
    return LVN_SUCCESS;
}

// the reference's loop (compute_cuckoo.cpp:89-132): entries whose insertion failed are rehashed with fresh parameters
static int retry_table_builds(lvn_meshgen *ctx, std::vector<FieldEntry *> todo, int firstAttempt)
{
    cudaStream_t st = ctx->stream;
    LV(ctx->d_csgCounts.reserve(todo.size()));
    LV(ctx->h_small.reserve(todo.size()));
    for (int attempt = firstAttempt; attempt < 64 && !todo.empty(); attempt++) {
        LV(enqueue_table_builds(ctx, todo, ctx->d_csgCounts.p, attempt));
        CU(cudaMemcpyAsync(ctx->h_small.p, ctx->d_csgCounts.p, todo.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        std::vector<FieldEntry *> failed;
        for (size_t i = 0; i < todo.size(); i++)
            if (ctx->h_small.p[i] != 0) failed.push_back(todo[i]);
        todo.swap(failed);
    }
    return todo.empty() ? LVN_SUCCESS : LVN_CL_ERROR;
}

// Cuckoo_InitialiseTable + Cuckoo_InsertKeys for every entry of the list: all fills and inserts
// are queued, one host wait reads every failure flag, entries whose insertion failed are
// rehashed with fresh parameters.  defer: do not wait -- the first attempt almost always succeeds
// (four hash functions at load 0.5); the flags are looked at behind the stream's next
// synchronisation (validate_pending_tables), before any result that read the tables is trusted.
static int build_cuckoo_tables(lvn_meshgen *ctx, const std::vector<FieldEntry *> &entries, bool defer = false)
{
    cudaStream_t st = ctx->stream;
    if (!ctx->pendingTables.empty()) {      // an earlier deferred build nobody has looked at yet
        CU(cudaStreamSynchronize(st));
        LV(validate_pending_tables(ctx, nullptr));
    }
    std::vector<FieldEntry *> todo;
    for (FieldEntry *fe : entries) {
        if (fe->d_table) { cudaFreeAsync(fe->d_table, st); fe->d_table = nullptr; }
        fe->prime = 0;
        fe->cuckooRetries = 0;
        if (fe->numEdges <= 0) continue;
        fe->prime = (unsigned int)host_find_next_prime((int)std::max(2048u, (unsigned int)fe->numEdges * 2u));   // MIN_TABLE_SIZE
        CU(cudaMallocAsync((void **)&fe->d_table, (size_t)fe->prime * sizeof(unsigned long long), st));
        todo.push_back(fe);
    }
    if (todo.empty()) return LVN_SUCCESS;
    StageTimer t(ctx, LVN_STAGE_CUCKOO, 0);
    if (!defer) return retry_table_builds(ctx, todo, 0);
    // the flags of a deferred build live in the mapped pinned mirror itself: cleared here (nothing is pending, see
    // above), set by the insert kernel in the rare failure: no memset and no copy on the stream
    LV(ctx->h_tableFlags.reserve(todo.size()));
    unsigned int *flagsDev = ctx->h_tableFlags.dev();
    if (!flagsDev) { g_lastCudaError = "pinned flag mirror is not mapped"; return LVN_ERR_CUDA; }
    memset(ctx->h_tableFlags.p, 0, todo.size() * sizeof(unsigned int));
    LV(enqueue_table_builds(ctx, todo, flagsDev, 0, true));
    ctx->pendingTables = todo;
    return LVN_SUCCESS;
}

// Call with the context's stream synchronised.  *rebuilt = a table had to be rehashed: whatever read it
// before (a batch that looked edges up through it) has to run again.
static int validate_pending_tables(lvn_meshgen *ctx, bool *rebuilt)
{
    if (rebuilt) *rebuilt = false;
    if (ctx->pendingTables.empty()) return LVN_SUCCESS;
    std::vector<FieldEntry *> failed;
    for (size_t i = 0; i < ctx->pendingTables.size(); i++)
        if (ctx->h_tableFlags.p[i] != 0 || ctx->forceTableRetry) failed.push_back(ctx->pendingTables[i]);
    ctx->pendingTables.clear();
    if (failed.empty()) return LVN_SUCCESS;
    if (rebuilt) *rebuilt = true;
    return retry_table_builds(ctx, failed, 1);
}

// ---------------------------------------------------------------------------
// density-field cache (LoadDensityField / StoreDensityField, compute_density_field.cpp:235-308)
// ---------------------------------------------------------------------------

// GenerateDefaultDensityField + FindDefaultEdges into persistent entries, n chunks in one pass of
// the path.  The entries come without cuckoo tables (an edit that follows rebuilds them anyway).
static int materialise_default_fields(lvn_meshgen *ctx, int n, const int32_t *minSize, std::vector<FieldEntry> &out)
{
    out.assign(n, FieldEntry());
    if (n <= 0) return LVN_SUCCESS;
    const Dims &d = ctx->dims;
    const size_t F3 = (size_t)d.F * d.F * d.F;
    BatchOpts opts;
    opts.ignoreFieldCache = true;
    opts.singleLane = true;     // descriptors and headers stay in the caller's order
    LV(run_batch(ctx, n, minSize, opts));
    cudaStream_t st = ctx->stream;
    std::vector<uint8_t *> ptrs(n);
    for (int i = 0; i < n; i++) {
        const ChunkHdr &h = ctx->h_hdrs.p[i];
        FieldEntry &fe = out[i];
        CU(cudaMallocAsync((void **)&fe.d_field, F3, st));
        ptrs[i] = fe.d_field;
        fe.numEdges = h.E;
        fe.lastCSGOperation = 0;
        if (h.E > 0) {
            CU(cudaMallocAsync((void **)&fe.d_keys, (size_t)h.E * sizeof(int), st));
            CU(cudaMallocAsync((void **)&fe.d_info, (size_t)h.E * sizeof(float4), st));
            CU(cudaMemcpyAsync(fe.d_keys, ctx->d_edgeKeys.p + h.edgeBase, (size_t)h.E * sizeof(int), cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(fe.d_info, ctx->d_edgeInfo.p + h.edgeBase, (size_t)h.E * sizeof(float4), cudaMemcpyDeviceToDevice, st));
        }
        if (g.densityKind != 0)
            CU(cudaMemcpyAsync(fe.d_field, ctx->d_tmpFields.p + (size_t)i * F3, F3, cudaMemcpyDeviceToDevice, st));
    }
    if (g.densityKind == 0) {
        StageTimer t(ctx, LVN_STAGE_FIELD, 1);
        LV(ctx->d_fieldPtrs.reserve(n));
        CU(cudaMemcpyAsync(ctx->d_fieldPtrs.p, ptrs.data(), n * sizeof(uint8_t *), cudaMemcpyHostToDevice, st));
        launch_field_from_heights(d, ctx->d_descs.p, n, ctx->d_heights.p, g.defaultMaterial, ctx->d_fieldPtrs.p, st);
        CU(cudaStreamSynchronize(st));   // ptrs must outlive the upload
    }
    return LVN_SUCCESS;
}

// One chunk of a batched edit: the field it applies to and its own op list (replayed stored ops
// differ from chunk to chunk).
struct CsgItem {
    FieldEntry *fe;
    int32_t ms[4];
    const lvn_csg_operation_info *ops;
    int numOps;
};

// ApplyCSGOperations, compute_csg.cpp:11-220, for many chunks: three batched launches, two host
// waits (the counts that size the new edge lists; the hash-table insert flags).
static int apply_csg_items(lvn_meshgen *ctx, const std::vector<CsgItem> &items)
{
    const int n = (int)items.size();
    if (n == 0) return LVN_SUCCESS;
    const Dims &d = ctx->dims;
    cudaStream_t st = ctx->stream;
    const double t0 = ctx->trace ? host_now_us() : 0.0;   // LVN_TRACE: where an edit's host time goes
    double t1 = 0.0, t2 = 0.0, t3 = 0.0;
    std::vector<CsgOpDev> hops;
    std::vector<CsgChunk> hc(n);
    const size_t numWords = ((size_t)3 * d.H * d.H * d.H + 31) / 32;
    for (int i = 0; i < n; i++) {
        const CsgItem &it = items[i];
        ChunkDesc cd;
        LV(fill_desc(ctx, it.ms, cd));
        CsgChunk &c = hc[i];
        memset(&c, 0, sizeof(c));
        c.ox = cd.ox; c.oy = cd.oy; c.oz = cd.oz; c.scale = cd.scale;
        c.opFirst = (int)hops.size();
        c.numOps = it.numOps;
        for (int k = 0; k < it.numOps; k++) {
            const lvn_csg_operation_info &o = it.ops[k];
            if (o.material < 0 || o.material > 255 || (o.brushShape != 0 && o.brushShape != 1) || (o.type != 0 && o.type != 1))
                return LVN_ERR_INVALID_VALUE;
            CsgOpDev v;
            v.type = o.type; v.shape = o.brushShape; v.material = o.material; v.pad = 0;
            v.ox = o.origin[0]; v.oy = o.origin[1]; v.oz = o.origin[2];
            v.dx = o.dimensions[0]; v.dy = o.dimensions[1]; v.dz = o.dimensions[2];
            v.c = cosf(o.rotateY);   // pR(), hg_sdf.glsl:460-463; host libm so that every
            v.s = sinf(o.rotateY);   // implementation fed the same op sees the same rotation
            hops.push_back(v);
        }
    }
    // one pinned staging block [chunks | ops] and one upload; the counters and the touched-edge bitmaps are one
    // block [8 counters per chunk | ticket | bitmaps] and one memset; the counters come back through the mapped
    // mirror, written by k_csg_count's last block (no copy-engine hop in front of the host wait)
    const size_t chunkBytes = ((size_t)n * sizeof(CsgChunk) + 15) & ~(size_t)15;
    const size_t stageBytes = chunkBytes + std::max<size_t>(hops.size(), 1) * sizeof(CsgOpDev);
    const size_t countWords = (((size_t)8 * n + 1) + 3) & ~(size_t)3;
    LV(ctx->h_csgStage.reserve(stageBytes + chunkBytes));   // second part: the chunks as re-sent for the emit pass
    LV(ctx->d_csgStage.reserve(stageBytes));
    LV(ctx->d_touched.reserve(countWords + numWords * n));
    LV(ctx->h_small.reserve((size_t)8 * n));
    unsigned int *h_counts_dev = ctx->h_small.dev();
    if (!h_counts_dev) { g_lastCudaError = "pinned counter mirror is not mapped"; return LVN_ERR_CUDA; }
    CsgChunk *d_chunks = (CsgChunk *)ctx->d_csgStage.p;
    const CsgOpDev *d_ops = (const CsgOpDev *)(ctx->d_csgStage.p + chunkBytes);
    unsigned int *d_counts = ctx->d_touched.p, *d_ticket = d_counts + (size_t)8 * n, *d_bitmaps = d_counts + countWords;
    for (int i = 0; i < n; i++) {
        FieldEntry &fe = *items[i].fe;
        CsgChunk &c = hc[i];
        c.field = fe.d_field;
        c.touched = d_bitmaps + numWords * i;
        c.oldKeys = fe.d_keys; c.oldInfo = fe.d_info; c.numOld = fe.numEdges;
        c.counts = d_counts + 8 * i;
    }
    memcpy(ctx->h_csgStage.p, hc.data(), (size_t)n * sizeof(CsgChunk));
    memcpy(ctx->h_csgStage.p + chunkBytes, hops.data(), hops.size() * sizeof(CsgOpDev));
    CU(cudaMemcpyAsync(ctx->d_csgStage.p, ctx->h_csgStage.p, stageBytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_counts, 0, (countWords + numWords * n) * sizeof(unsigned int), st));
    {
        StageTimer t(ctx, LVN_STAGE_CSG, 2);
        launch_csg_materials_count(d, d_chunks, n, d_ops, d_ticket, h_counts_dev, st);
    }
    if (ctx->trace) t1 = host_now_us();
    CU(cudaStreamSynchronize(st));
    if (ctx->trace) t2 = host_now_us();
    collect_stage_times(ctx);

    // When every old edge is invalidated the reference keeps the stale list (numPrunedEdges == 0
    // skips the swap, compute_csg.cpp:160) and ends up with duplicate keys whose winner depends
    // on hash order; here the evident intent: the surviving list is empty.  DESIGN.md "deviations".
    bool anyEmit = false;
    for (int i = 0; i < n; i++) {
        const unsigned int kept = ctx->h_small.p[8 * i + 0], created = ctx->h_small.p[8 * i + 1], changed = ctx->h_small.p[8 * i + 2];
        CsgChunk &c = hc[i];
        c.skip = changed == 0;   // numUpdatedPoints <= 0, compute_csg.cpp:62-66
        c.numKept = (int)kept;
        if (c.skip) continue;
        const unsigned int newE = kept + created;
        if (newE) {
            CU(cudaMallocAsync((void **)&c.newKeys, (size_t)newE * sizeof(int), st));
            CU(cudaMallocAsync((void **)&c.newInfo, (size_t)newE * sizeof(float4), st));
            anyEmit = true;
        }
    }
    if (anyEmit) {
        // the staging block's second part: rewritten by the next edit only behind that edit's own host wait
        memcpy(ctx->h_csgStage.p + stageBytes, hc.data(), (size_t)n * sizeof(CsgChunk));
        CU(cudaMemcpyAsync(d_chunks, ctx->h_csgStage.p + stageBytes, (size_t)n * sizeof(CsgChunk), cudaMemcpyHostToDevice, st));
        StageTimer t(ctx, LVN_STAGE_CSG, 1);
        launch_csg_emit(d, d_chunks, n, d_ops, st);
    }
    std::vector<FieldEntry *> rebuild;
    for (int i = 0; i < n; i++) {
        FieldEntry &fe = *items[i].fe;
        const CsgChunk &c = hc[i];
        if (c.skip) {
            if (fe.numEdges > 0 && !fe.d_table) rebuild.push_back(&fe);   // a fresh entry the edit did not touch
            continue;
        }
        if (fe.d_keys) cudaFreeAsync(fe.d_keys, st);
        if (fe.d_info) cudaFreeAsync(fe.d_info, st);
        fe.d_keys = c.newKeys;
        fe.d_info = c.newInfo;
        fe.numEdges = (int)(ctx->h_small.p[8 * i + 0] + ctx->h_small.p[8 * i + 1]);
        rebuild.push_back(&fe);
    }
    // no host wait here: the uploads above came from pageable memory (staged before the call returned), the
    // tables' insert flags are read behind the next synchronisation of the stream (validate_pending_tables)
    if (ctx->trace) t3 = host_now_us();
    LV(build_cuckoo_tables(ctx, rebuild, !ctx->profiling));
    if (ctx->profiling) { CU(cudaStreamSynchronize(st)); collect_stage_times(ctx); }
    CU(cudaGetLastError());
    if (ctx->trace)
        fprintf(stderr, "[lvn trace] edit of %d chunks: host %.0f us to enqueue materials + count, %.0f us waiting for the counts, "
                        "%.0f us to size and enqueue emit, %.0f us to enqueue %d hash tables\n",
                n, t1 - t0, t2 - t1, t3 - t2, host_now_us() - t3, (int)rebuild.size());
    return LVN_SUCCESS;
}

static bool aabb_overlaps(const lvn_aabb &a, const lvn_aabb &b)   // AABB::overlaps, aabb.h:24-33
{
    return !(a.max[0] < b.min[0] || a.max[1] < b.min[1] || a.max[2] < b.min[2] ||
             a.min[0] > b.max[0] || a.min[1] > b.max[1] || a.min[2] > b.max[2]);
}

// LoadDensityField, compute_density_field.cpp:235-274, for n chunks.  out[i] = the cache entry,
// or nullptr when chunk i is a plain default field that never needs to persist.  Missing entries
// are materialised in one pass of the path; stored operations overlapping a chunk since its
// lastCSGOperation are replayed (one ApplyCSGOperations call per chunk, batched over chunks).
// withTables: entries that were created here and not edited get their hash table too.
static int load_density_fields(lvn_meshgen *ctx, int n, const int32_t *minSize, bool forceEntry, bool withTables,
                               std::vector<FieldEntry *> &out)
{
    out.assign(n, nullptr);
    std::vector<std::vector<lvn_csg_operation_info>> replay(n);
    std::vector<int> missing;
    std::vector<int32_t> missingMs;
    for (int i = 0; i < n; i++) {
        const int32_t *ms = &minSize[4 * i];
        auto it = ctx->fields.find(make_key(ms, ms[3]));
        const int startOp = it != ctx->fields.end() ? it->second.lastCSGOperation : 0;
        lvn_aabb bb;
        for (int k = 0; k < 3; k++) { bb.min[k] = ms[k]; bb.max[k] = ms[k] + ms[3]; }
        for (size_t o = (size_t)startOp; o < g.storedOps.size(); o++)
            if (aabb_overlaps(bb, g.storedAABBs[o])) replay[i].push_back(g.storedOps[o]);
        if (it != ctx->fields.end()) { out[i] = &it->second; continue; }
        if (replay[i].empty() && !forceEntry) continue;
        bool dup = false;   // the same chunk twice in one list: one entry
        for (int j : missing) dup = dup || (memcmp(&minSize[4 * j], ms, 4 * sizeof(int32_t)) == 0);
        if (!dup) { missing.push_back(i); missingMs.insert(missingMs.end(), ms, ms + 4); }
    }
    if (!missing.empty()) {
        std::vector<FieldEntry> fresh;
        const int rc = materialise_default_fields(ctx, (int)missing.size(), missingMs.data(), fresh);
        if (rc < 0) { for (FieldEntry &fe : fresh) fe.release(ctx->stream); return rc; }
        for (size_t k = 0; k < missing.size(); k++) {
            const int32_t *ms = &minSize[4 * missing[k]];
            ctx->fields.emplace(make_key(ms, ms[3]), fresh[k]);
        }
        for (int i = 0; i < n; i++)
            if (!out[i] && (!replay[i].empty() || forceEntry)) out[i] = &ctx->fields.find(make_key(&minSize[4 * i], minSize[4 * i + 3]))->second;
    }
    std::vector<CsgItem> items;
    std::vector<FieldEntry *> untouched;
    for (int i = 0; i < n; i++) {
        if (!out[i]) continue;
        FieldEntry &fe = *out[i];
        fe.lastCSGOperation = (int)g.storedOps.size();
        if (!replay[i].empty()) {
            CsgItem it;
            it.fe = &fe; memcpy(it.ms, &minSize[4 * i], sizeof(it.ms));
            it.ops = replay[i].data(); it.numOps = (int)replay[i].size();
            items.push_back(it);
        } else if (withTables && fe.numEdges > 0 && !fe.d_table) {
            untouched.push_back(&fe);
        }
    }
    // a chunk listed twice replays once (its replay list was taken before lastCSGOperation moved)
    {
        std::vector<CsgItem> uniq;
        for (const CsgItem &it : items) {
            bool seen = false;
            for (const CsgItem &u : uniq) seen = seen || u.fe == it.fe;
            if (!seen) uniq.push_back(it);
        }
        LV(apply_csg_items(ctx, uniq));
    }
    if (!untouched.empty()) {
        std::sort(untouched.begin(), untouched.end());
        untouched.erase(std::unique(untouched.begin(), untouched.end()), untouched.end());
        LV(build_cuckoo_tables(ctx, untouched));
    }
    return LVN_SUCCESS;
}

static int load_density_field(lvn_meshgen *ctx, const int32_t min[3], int size, bool forceEntry, FieldEntry **out)
{
    const int32_t ms[4] = {min[0], min[1], min[2], size};
    std::vector<FieldEntry *> v;
    LV(load_density_fields(ctx, 1, ms, forceEntry, true, v));
    *out = v[0];
    return LVN_SUCCESS;
}

extern "C" int lvn_meshgen_apply_csg_operations_batch(lvn_meshgen *ctx, const lvn_csg_operation_info *ops, int numOps,
                                                      int nChunks, const int32_t *chunkMinSize)
{
    if (!g.initialised) return LVN_ERR_NOT_INITIALISED;
    if (!ctx || numOps < 0 || (numOps > 0 && !ops) || nChunks < 0 || (nChunks > 0 && !chunkMinSize)) return LVN_ERR_INVALID_VALUE;
    if (nChunks == 0) return LVN_SUCCESS;
    // Compute_ApplyCSGOperations, compute_csg.cpp:224-242, per chunk: LoadDensityField (create +
    // replay), ApplyCSGOperations(new ops), lastCSGOperation += numOps
    std::vector<FieldEntry *> entries;
    LV(load_density_fields(ctx, nChunks, chunkMinSize, true, numOps == 0, entries));
    std::vector<CsgItem> items;
    for (int i = 0; i < nChunks; i++) {
        bool seen = false;   // the reference would apply the ops twice to a chunk listed twice; so do we, in order
        for (const CsgItem &u : items) seen = seen || u.fe == entries[i];
        if (seen) { LV(apply_csg_items(ctx, items)); items.clear(); }
        CsgItem it;
        it.fe = entries[i]; memcpy(it.ms, &chunkMinSize[4 * i], sizeof(it.ms));
        it.ops = ops; it.numOps = numOps;
        if (numOps > 0) items.push_back(it);
        entries[i]->lastCSGOperation += numOps;   // compute_csg.cpp:237
    }
    LV(apply_csg_items(ctx, items));
    return LVN_SUCCESS;
}

extern "C" int lvn_meshgen_apply_csg_operations(lvn_meshgen *ctx, const lvn_csg_operation_info *ops, int numOps,
                                                const int32_t clipmapNodeMin[3], int clipmapNodeSize)
{
    if (!clipmapNodeMin) return LVN_ERR_INVALID_VALUE;
    const int32_t ms[4] = {clipmapNodeMin[0], clipmapNodeMin[1], clipmapNodeMin[2], clipmapNodeSize};
    return lvn_meshgen_apply_csg_operations_batch(ctx, ops, numOps, 1, ms);
}

extern "C" int lvn_meshgen_free_chunk_octree(lvn_meshgen *ctx, const int32_t min[3], int size)
{
    if (!ctx) return LVN_ERR_INVALID_VALUE;
    auto it = ctx->octrees.find(make_key(min, size));
    if (it != ctx->octrees.end()) {
        it->second.release(ctx->stream);
        ctx->octrees.erase(it);
    }
    return LVN_SUCCESS;
}

extern "C" int lvn_meshgen_is_chunk_empty(lvn_meshgen *ctx, const int32_t min[3], int size, int *isEmpty)
{
    if (!g.initialised) return LVN_ERR_NOT_INITIALISED;
    if (!ctx || !isEmpty) return LVN_ERR_INVALID_VALUE;
    // Compute_ChunkIsEmpty (compute_density_field.cpp:278-299): cached field, else generate the
    // default field and store it -- stored CSG operations are NOT replayed on this path
    auto it = ctx->fields.find(make_key(min, size));
    if (it == ctx->fields.end()) {
        const int32_t ms[4] = {min[0], min[1], min[2], size};
        std::vector<FieldEntry> fresh;
        const int rc = materialise_default_fields(ctx, 1, ms, fresh);
        if (rc < 0) { for (FieldEntry &fe : fresh) fe.release(ctx->stream); return rc; }
        it = ctx->fields.emplace(make_key(min, size), fresh[0]).first;
        LV(build_cuckoo_tables(ctx, std::vector<FieldEntry *>(1, &it->second)));
    }
    *isEmpty = it->second.numEdges == 0;
    return LVN_SUCCESS;
}

// ---------------------------------------------------------------------------
// generateChunkMesh (compute_octree.cpp:351-375) with LoadOctree's cache (:154-181)
// ---------------------------------------------------------------------------
extern "C" int lvn_meshgen_generate_chunk_mesh(lvn_meshgen *ctx, const int32_t min[3], int clipmapNodeSize,
                                               lvn_mesh_vertex *vertices, int vertexCapacity, int *numVertices,
                                               lvn_mesh_triangle *triangles, int triangleCapacity, int *numTriangles,
                                               lvn_seam_node_info *seamNodes, int seamCapacity, int *numSeamNodes)
{
    if (!g.initialised) return LVN_ERR_NOT_INITIALISED;
    if (!ctx || !numVertices || !numTriangles || !numSeamNodes) return LVN_ERR_INVALID_VALUE;
    *numVertices = *numTriangles = *numSeamNodes = 0;
    cudaStream_t st = ctx->stream;
    const Key key = make_key(min, clipmapNodeSize);
    auto oit = ctx->octrees.find(key);
    if (oit == ctx->octrees.end()) {
        FieldEntry *fe = nullptr;
        LV(load_density_field(ctx, min, clipmapNodeSize, false, &fe));
        int32_t ms[4] = {min[0], min[1], min[2], clipmapNodeSize};
        BatchOpts opts;
        LV(run_batch(ctx, 1, ms, opts));
        const ChunkHdr h = ctx->h_hdrs.p[0];
        if (h.E == 0) return LVN_SUCCESS;   // "no point in trying to construct the octree"
        OctreeEntry oe;
        oe.numNodes = h.N; oe.numQuads = h.Q; oe.numSeams = h.S;
        LV(oe.alloc(st));
        if (h.N) CU(cudaMemcpyAsync(oe.d_v, ctx->d_vertices.p + h.nodeBase, (size_t)h.N * sizeof(lvn_mesh_vertex), cudaMemcpyDeviceToDevice, st));
        if (h.Q) CU(cudaMemcpyAsync(oe.d_t, ctx->d_tris.p + (size_t)h.quadBase * 6, (size_t)h.Q * 6 * sizeof(int), cudaMemcpyDeviceToDevice, st));
        if (h.S) CU(cudaMemcpyAsync(oe.d_s, ctx->d_seams.p + h.seamBase, (size_t)h.S * sizeof(lvn_seam_node_info), cudaMemcpyDeviceToDevice, st));
        oit = ctx->octrees.emplace(key, oe).first;
    }
    const OctreeEntry &oe = oit->second;
    // no quads -> the mesh buffer stays empty (compute_octree.cpp:227-232); seams are still gathered
    const int nV = oe.numQuads > 0 ? oe.numNodes : 0, nT = 2 * oe.numQuads, nS = oe.numSeams;
    *numVertices = nV; *numTriangles = nT; *numSeamNodes = nS;
    if (nV > vertexCapacity || nT > triangleCapacity || nS > seamCapacity) return LVN_ERR_CAPACITY;
    if ((nV && !vertices) || (nT && !triangles) || (nS && !seamNodes)) return LVN_ERR_INVALID_VALUE;
    if (nV) CU(cudaMemcpyAsync(vertices, oe.d_v, (size_t)nV * sizeof(lvn_mesh_vertex), cudaMemcpyDeviceToHost, st));
    if (nT) CU(cudaMemcpyAsync(triangles, oe.d_t, (size_t)nT * sizeof(lvn_mesh_triangle), cudaMemcpyDeviceToHost, st));
    if (nS) CU(cudaMemcpyAsync(seamNodes, oe.d_s, (size_t)nS * sizeof(lvn_seam_node_info), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return LVN_SUCCESS;
}

extern "C" int lvn_debug_solve_qefs(int packed, int n, const float *qefs16, float *positions4)
{
    if (!g.initialised) return LVN_ERR_NOT_INITIALISED;
    if (n < 0 || (n > 0 && (!qefs16 || !positions4))) return LVN_ERR_INVALID_VALUE;
    if (n == 0) return LVN_SUCCESS;
    DevBuf<float> d_q;
    DevBuf<float4> d_p;
    int rc = LVN_SUCCESS;
    if ((rc = d_q.reserve((size_t)n * 16)) == LVN_SUCCESS && (rc = d_p.reserve(n)) == LVN_SUCCESS) {
        if (cudaMemcpy(d_q.p, qefs16, (size_t)n * 64, cudaMemcpyHostToDevice) != cudaSuccess) rc = LVN_ERR_CUDA;
        if (rc == LVN_SUCCESS) {
            launch_solve_debug(packed, n, d_q.p, d_p.p, 0);
            if (cudaMemcpy(positions4, d_p.p, (size_t)n * 16, cudaMemcpyDeviceToHost) != cudaSuccess) rc = LVN_ERR_CUDA;
        }
    }
    d_q.release(); d_p.release();
    return rc;
}

// ---------------------------------------------------------------------------
// per-stage dump
// ---------------------------------------------------------------------------
extern "C" int lvn_meshgen_debug_dump_chunk(lvn_meshgen *ctx, const int32_t min[3], int size, lvn_stage_dump *dump)
{
    if (!g.initialised) return LVN_ERR_NOT_INITIALISED;
    if (!ctx || !dump) return LVN_ERR_INVALID_VALUE;
    const Dims &d = ctx->dims;
    const size_t F3 = (size_t)d.F * d.F * d.F;
    cudaStream_t st = ctx->stream;
    FieldEntry *fe = nullptr;
    LV(load_density_field(ctx, min, size, false, &fe));
    int32_t ms[4] = {min[0], min[1], min[2], size};
    BatchOpts opts;
    opts.debug = true;
    LV(run_batch(ctx, 1, ms, opts));
    const ChunkHdr h = ctx->h_hdrs.p[0];
    dump->numEdges = h.E; dump->numNodes = h.N; dump->numTriangles = 2 * h.Q; dump->numSeamNodes = h.S;
    if (h.E > dump->edgeCapacity || h.N > dump->nodeCapacity) return LVN_ERR_CAPACITY;
    if (dump->materials) {
        if (fe) {
            CU(cudaMemcpyAsync(dump->materials, fe->d_field, F3, cudaMemcpyDeviceToHost, st));
        } else if (g.densityKind != 0) {
            CU(cudaMemcpyAsync(dump->materials, ctx->d_tmpFields.p, F3, cudaMemcpyDeviceToHost, st));
        } else {
            LV(ctx->d_tmpFields.reserve(F3));
            LV(ctx->d_fieldPtrs.reserve(1));
            CU(cudaMemcpyAsync(ctx->d_fieldPtrs.p, &ctx->d_tmpFields.p, sizeof(uint8_t *), cudaMemcpyHostToDevice, st));
            launch_field_from_heights(d, ctx->d_descs.p, 1, ctx->d_heights.p, g.defaultMaterial, ctx->d_fieldPtrs.p, st);
            CU(cudaMemcpyAsync(dump->materials, ctx->d_tmpFields.p, F3, cudaMemcpyDeviceToHost, st));
        }
    }
    if (h.E) {
        const int *keys = fe ? fe->d_keys : ctx->d_edgeKeys.p + h.edgeBase;
        const float4 *info = fe ? fe->d_info : ctx->d_edgeInfo.p + h.edgeBase;
        if (dump->edgeKeys) CU(cudaMemcpyAsync(dump->edgeKeys, keys, (size_t)h.E * 4, cudaMemcpyDeviceToHost, st));
        if (dump->edgeInfo) CU(cudaMemcpyAsync(dump->edgeInfo, info, (size_t)h.E * 16, cudaMemcpyDeviceToHost, st));
    }
    if (h.N) {
        const size_t b = (size_t)h.nodeBase, N = (size_t)h.N;
        if (dump->nodeCodes) CU(cudaMemcpyAsync(dump->nodeCodes, ctx->d_dbgCodes.p + b, N * 4, cudaMemcpyDeviceToHost, st));
        if (dump->nodeEdgeMasks) CU(cudaMemcpyAsync(dump->nodeEdgeMasks, ctx->d_dbgMasks.p + b, N * 4, cudaMemcpyDeviceToHost, st));
        if (dump->nodeMaterials) CU(cudaMemcpyAsync(dump->nodeMaterials, ctx->d_dbgMats.p + b, N * 4, cudaMemcpyDeviceToHost, st));
        if (dump->nodeQEFs) CU(cudaMemcpyAsync(dump->nodeQEFs, ctx->d_dbgQefs.p + b * 16, N * 64, cudaMemcpyDeviceToHost, st));
        if (dump->nodePositions) CU(cudaMemcpyAsync(dump->nodePositions, ctx->d_dbgPos.p + b, N * 16, cudaMemcpyDeviceToHost, st));
        if (dump->nodeNormals) CU(cudaMemcpyAsync(dump->nodeNormals, ctx->d_dbgNrm.p + b, N * 16, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return LVN_SUCCESS;
}

// ---------------------------------------------------------------------------
// utilities (a9, a15)
// ---------------------------------------------------------------------------
extern "C" int lvn_global_mesh_offsets(const int32_t *gathered, int worldSize, int perRank, int numChunks,
                                       int64_t *counts, int64_t *offsets, int64_t totals[3])
{
    if (!gathered || !counts || !offsets || !totals || worldSize < 1 || perRank < 0 || numChunks < 0 ||
        (int64_t)numChunks > (int64_t)worldSize * perRank) return LVN_ERR_INVALID_VALUE;
    int64_t run[3] = {0, 0, 0};
    for (int i = 0; i < numChunks; i++) {
        const int32_t *src = gathered + ((size_t)(i % worldSize) * 3) * (size_t)perRank + (size_t)(i / worldSize);
        for (int k = 0; k < 3; k++) {
            const int64_t c = src[(size_t)k * perRank];
            counts[3 * (size_t)i + k] = c;
            offsets[3 * (size_t)i + k] = run[k];
            run[k] += c;
        }
    }
    totals[0] = run[0]; totals[1] = run[1]; totals[2] = run[2];
    return LVN_SUCCESS;
}

extern "C" int lvn_find_next_prime(int n) { return host_find_next_prime(n); }

extern "C" int lvn_exclusive_scan(const int32_t *data, int32_t *scan, int count)
{
    if (!g.initialised) return LVN_ERR_NOT_INITIALISED;
    if (count < 0 || (count > 0 && (!data || !scan))) return LVN_ERR_INVALID_VALUE;
    if (count == 0) return 0;
    DevBuf<int> d_in, d_out, d_sums, d_total;
    int total = 0, rc = LVN_SUCCESS;
    auto body = [&]() -> int {
        LV(d_in.reserve(count)); LV(d_out.reserve(count));
        LV(d_sums.reserve(scan_block_sums_needed(count))); LV(d_total.reserve(1));
        CU(cudaMemcpy(d_in.p, data, (size_t)count * 4, cudaMemcpyHostToDevice));
        launch_exclusive_scan(d_in.p, d_out.p, count, d_sums.p, d_total.p, 0);
        CU(cudaMemcpy(scan, d_out.p, (size_t)count * 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&total, d_total.p, 4, cudaMemcpyDeviceToHost));
        return LVN_SUCCESS;
    };
    rc = body();
    d_in.release(); d_out.release(); d_sums.release(); d_total.release();
    return rc < 0 ? rc : total;
}

extern "C" int lvn_compact_index_array(const int32_t *values, const int32_t *valid, int count, int32_t *out)
{
    if (!g.initialised) return LVN_ERR_NOT_INITIALISED;
    if (count < 0 || (count > 0 && (!values || !valid || !out))) return LVN_ERR_INVALID_VALUE;
    if (count == 0) return 0;
    DevBuf<int> d_val, d_valid, d_scan, d_out, d_sums, d_total;
    int total = 0, rc = LVN_SUCCESS;
    auto body = [&]() -> int {
        LV(d_val.reserve(count)); LV(d_valid.reserve(count)); LV(d_scan.reserve(count)); LV(d_out.reserve(count));
        LV(d_sums.reserve(scan_block_sums_needed(count))); LV(d_total.reserve(1));
        CU(cudaMemcpy(d_val.p, values, (size_t)count * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d_valid.p, valid, (size_t)count * 4, cudaMemcpyHostToDevice));
        launch_exclusive_scan(d_valid.p, d_scan.p, count, d_sums.p, d_total.p, 0);
        launch_compact(d_val.p, d_valid.p, d_scan.p, count, d_out.p, 0);
        CU(cudaMemcpy(&total, d_total.p, 4, cudaMemcpyDeviceToHost));
        if (total > 0) CU(cudaMemcpy(out, d_out.p, (size_t)total * 4, cudaMemcpyDeviceToHost));
        return LVN_SUCCESS;
    };
    rc = body();
    d_val.release(); d_valid.release(); d_scan.release(); d_out.release(); d_sums.release(); d_total.release();
    return rc < 0 ? rc : total;
}

extern "C" int lvn_remove_duplicates(const int32_t *values, int count, int32_t *out)
{
    if (!g.initialised) return LVN_ERR_NOT_INITIALISED;
    if (count < 0 || (count > 0 && (!values || !out))) return LVN_ERR_INVALID_VALUE;
    if (count == 0) return 0;
    for (int i = 0; i < count; i++) if (values[i] == -1) return LVN_ERR_INVALID_VALUE;   // -1 marks an empty slot
    const unsigned int tableSize = (unsigned int)host_find_next_prime(count * 2);   // compute.cpp:462
    DevBuf<int> d_val, d_out;
    DevBuf<unsigned int> d_table, d_count;
    unsigned int unique = 0;
    int rc = LVN_SUCCESS;
    auto body = [&]() -> int {
        LV(d_val.reserve(count)); LV(d_out.reserve(count)); LV(d_table.reserve(tableSize)); LV(d_count.reserve(1));
        CU(cudaMemcpy(d_val.p, values, (size_t)count * 4, cudaMemcpyHostToDevice));
        CU(cudaMemset(d_table.p, 0xff, (size_t)tableSize * 4));
        CU(cudaMemset(d_count.p, 0, 4));
        launch_dedupe(d_val.p, count, d_table.p, tableSize, d_out.p, d_count.p, 0);
        CU(cudaMemcpy(&unique, d_count.p, 4, cudaMemcpyDeviceToHost));
        if (unique) CU(cudaMemcpy(out, d_out.p, (size_t)unique * 4, cudaMemcpyDeviceToHost));
        return LVN_SUCCESS;
    };
    rc = body();
    d_val.release(); d_out.release(); d_table.release(); d_count.release();
    return rc < 0 ? rc : (int)unique;
}

struct lvn_cuckoo {
    unsigned long long *d_table = nullptr;
    unsigned int prime = 0;
    unsigned int params[8] = {0};
    unsigned int tableSize = 0;
    int retries = 0;
    int insertedKeys = 0;
};

extern "C" lvn_cuckoo *lvn_cuckoo_create(unsigned int tableSize)
{
    if (!g.initialised) return nullptr;
    lvn_cuckoo *t = new lvn_cuckoo;
    t->tableSize = tableSize;
    t->prime = (unsigned int)host_find_next_prime((int)std::max(2048u, tableSize * 2u));
    if (cudaMalloc((void **)&t->d_table, (size_t)t->prime * 8) != cudaSuccess) { delete t; return nullptr; }
    launch_fill_u64(t->d_table, t->prime, ~0ull, 0);
    draw_cuckoo_params(t->params);
    return t;
}

extern "C" int lvn_cuckoo_insert_keys(lvn_cuckoo *t, const uint32_t *keys, unsigned int count)
{
    if (!t || (count && !keys)) return LVN_ERR_INVALID_VALUE;
    if (!count) return LVN_SUCCESS;
    DevBuf<unsigned int> d_keys, d_failed;
    int rc = LVN_CL_ERROR;
    auto body = [&]() -> int {
        LV(d_keys.reserve(count)); LV(d_failed.reserve(1));
        CU(cudaMemcpy(d_keys.p, keys, (size_t)count * 4, cudaMemcpyHostToDevice));
        for (int attempt = 0; attempt < 64; attempt++) {
            if (attempt) {   // rehash with fresh parameters, compute_cuckoo.cpp:93-109
                draw_cuckoo_params(t->params);
                launch_fill_u64(t->d_table, t->prime, ~0ull, 0);
                t->retries++;
            }
            CU(cudaMemset(d_failed.p, 0, 4));
            launch_cuckoo_insert(d_keys.p, count, t->d_table, t->prime, t->params, d_failed.p, 0);
            unsigned int failed = 0;
            CU(cudaMemcpy(&failed, d_failed.p, 4, cudaMemcpyDeviceToHost));
            if (!failed) { t->insertedKeys += (int)count; return LVN_SUCCESS; }
        }
        return LVN_CL_ERROR;
    };
    rc = body();
    d_keys.release(); d_failed.release();
    return rc;
}

extern "C" int lvn_cuckoo_find(const lvn_cuckoo *t, const uint32_t *keys, unsigned int count, uint32_t *values)
{
    if (!t || (count && (!keys || !values))) return LVN_ERR_INVALID_VALUE;
    if (!count) return LVN_SUCCESS;
    DevBuf<unsigned int> d_keys, d_vals;
    auto body = [&]() -> int {
        LV(d_keys.reserve(count)); LV(d_vals.reserve(count));
        CU(cudaMemcpy(d_keys.p, keys, (size_t)count * 4, cudaMemcpyHostToDevice));
        launch_cuckoo_find(d_keys.p, count, t->d_table, t->prime, t->params, d_vals.p, 0);
        CU(cudaMemcpy(values, d_vals.p, (size_t)count * 4, cudaMemcpyDeviceToHost));
        return LVN_SUCCESS;
    };
    const int rc = body();
    d_keys.release(); d_vals.release();
    return rc;
}

extern "C" int lvn_cuckoo_prime(const lvn_cuckoo *t) { return t ? (int)t->prime : 0; }
extern "C" int lvn_cuckoo_retries(const lvn_cuckoo *t) { return t ? t->retries : 0; }
extern "C" void lvn_cuckoo_destroy(lvn_cuckoo *t)
{
    if (!t) return;
    if (t->d_table) cudaFree(t->d_table);
    delete t;
}
