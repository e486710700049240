/*
 * lvn_oracle.c -- CPU ORACLE (test infrastructure, see lvn_oracle.h header).
 *
 * Plain C restatement of the reference's chunk-meshing path.  Citations are
 * relative to /root/reference/.  Build: oracle/Makefile
 *   gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC
 * -ffp-contract=off is REQUIRED: the arithmetic spec allows a fused
 * multiply-add only where the source calls fmaf() explicitly.
 */
#include "lvn_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>

#endif

/* dot(float4, float4): ONE built-in, one definition -- the same fma chain the simplex
 * dot products use, for every call site (qef.cl:25-27,149,182). */
static inline float dot4(float ax, float ay, float az, float aw, float bx, float by, float bz, float bw)
{
    return fmaf(aw, bw, fmaf(az, bz, fmaf(ay, by, ax * bx)));
}

#if defined(__x86_64__) && defined(__GNUC__)
/* fast fmaf on FMA-capable hosts, libm fmaf elsewhere: same result */
#define LVO_HOT __attribute__((target_clones("fma", "default")))
#else
#define LVO_HOT
#endif
#define LVO_INLINE static inline __attribute__((always_inline))

/* ------------------------------------------------------------------------ */
/* world state                                                              */
/* ------------------------------------------------------------------------ */

typedef struct {
    int      min[3], size;
    int      lastCSGOperation;
    int      numEdges;
    int32_t *edgeKeys;
    lvo_f4  *edgeInfo;
    int32_t *materials;
} lvo_field;   /* GPUDensityField, compute_local.h:28-38 */

typedef struct {
    int       min[3], size;
    int       numNodes;
    uint32_t *codes;
    int32_t  *edgeMasks;   /* kept for the stage dump only */
    int32_t  *matWords;
    lvo_qef  *qefs;        /* kept for the stage dump only */
    lvo_f4   *positions, *normals;
    /* snapshot of the field the octree was built from (stage dump only) */
    int       numEdges;
    int32_t  *materials, *edgeKeys;
    lvo_f4   *edgeInfo;
} lvo_octree;  /* GPUOctree, compute_local.h:44-50 */

struct lvo_world {
    uint8_t  image[256 * 256 * 4];
    int      defaultMaterial;
    int      V, H, F, shift, mask, depth;   /* compute.cpp:245-252,271 */
    int      densityKind;
    float    stressThreshold;
    /* g_storedOps / g_storedOpAABBs, compute_density_field.cpp:23-24 */
    lvo_csg_op *ops; int (*opAABB)[6]; int numOps, capOps;
    lvo_field  *fields;  int numFields, capFields;     /* densityFieldCache */
    lvo_octree *octrees; int numOctrees, capOctrees;   /* octreeCache */
};

/* ------------------------------------------------------------------------ */
/* a1: noise table -- compute_density_field.cpp:28-113                      */
/* ------------------------------------------------------------------------ */

static const int PERM256[256] = {151,160,137,91,90,15,
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

/* NB compute_density_field.cpp:28-53: perm[512] is NOT the table twice:
 * the second half starts again at entry 6 ("131,13,201,...") and is 6 short,
 * so the initialiser holds 256 + 250 values and the last 6 are zero. */
static void perm512(int out[512])
{
    int i;
    for (i = 0; i < 256; i++) out[i] = PERM256[i];
    for (i = 0; i < 250; i++) out[256 + i] = PERM256[6 + i];
    for (i = 506; i < 512; i++) out[i] = 0;
}

static const int GRAD3[16][3] = {{0,1,1},{0,1,-1},{0,-1,1},{0,-1,-1},
    {1,0,1},{1,0,-1},{-1,0,1},{-1,0,-1},
    {1,1,0},{1,-1,0},{-1,1,0},{-1,-1,0},
    {1,0,-1},{-1,0,-1},{0,-1,1},{0,1,1}};

/* compute_density_field.cpp:69-88 (Jenkins one-at-a-time over the 4 key bytes) */
uint32_t lvo_noise_hash(int x, int y, int seed)
{
    const uint32_t key = (((uint32_t)x << 24) | ((uint32_t)y << 16)) ^ (uint32_t)seed;
    uint32_t hash = 0;
    int i;
    for (i = 0; i < 4; i++) {
        hash += (key >> (8 * i)) & 0xffu;   /* little-endian keyBytes[i] */
        hash += (hash << 10);
        hash ^= (hash >> 6);
    }
    hash += (hash << 3);
    hash ^= (hash >> 11);
    hash += (hash << 15);
    return hash;
}

/* MT19937 (Matsumoto & Nishimura), == std::mt19937 */
typedef struct { uint32_t mt[624]; int idx; } lvo_mt;
static void mt_seed(lvo_mt *m, uint32_t s)
{
    int i;
    m->mt[0] = s;
    for (i = 1; i < 624; i++)
        m->mt[i] = 1812433253u * (m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) + (uint32_t)i;
    m->idx = 624;
}
static uint32_t mt_next(lvo_mt *m)
{
    uint32_t y;
    if (m->idx >= 624) {
        int i;
        for (i = 0; i < 624; i++) {
            y = (m->mt[i] & 0x80000000u) | (m->mt[(i + 1) % 624] & 0x7fffffffu);
            m->mt[i] = m->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->idx = 0;
    }
    y = m->mt[m->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* compute_density_field.cpp:90-113; shuffle = our documented generator */
void lvo_noise_image(int seed, uint8_t *rgba)
{
    int shuffled[512];
    lvo_mt mt;
    int i, j;
    perm512(shuffled);
    mt_seed(&mt, (uint32_t)seed);
    for (i = 511; i > 0; i--) {
        const int k = (int)(mt_next(&mt) % (uint32_t)(i + 1));
        const int tmp = shuffled[i]; shuffled[i] = shuffled[k]; shuffled[k] = tmp;
    }
    for (i = 0; i < 256; i++)
        for (j = 0; j < 256; j++) {
            const int offset = ((i * 256) + j) * 4;
            const unsigned char value = (unsigned char)shuffled[lvo_noise_hash(i, j, seed) & 0x1ff];
            rgba[offset + 0] = (uint8_t)(GRAD3[value & 0x0f][0] * 64 + 64);
            rgba[offset + 1] = (uint8_t)(GRAD3[value & 0x0f][1] * 64 + 64);
            rgba[offset + 2] = (uint8_t)(GRAD3[value & 0x0f][2] * 64 + 64);
            rgba[offset + 3] = value;
        }
}

/* ------------------------------------------------------------------------ */
/* a2: simplex noise, fractals, terrain                                     */
/* ------------------------------------------------------------------------ */

/* read_imagef(permTexture, permSampler, coord) with CLK_FILTER_NEAREST |
 * CLK_ADDRESS_REPEAT | CLK_NORMALIZED_COORDS_TRUE (simplex.cl:59) at the
 * coordinate Pi*ONE+ONEHALF: the arithmetic is exact for |Pi| < 2^14, so
 * the texel is (column = Pi.x mod 256, row = Pi.y mod 256); UNORM8 -> float
 * is byte/255 (correctly rounded). */
LVO_INLINE const uint8_t *texel(const lvo_world *w, int col, int row)
{
    return &w->image[(((row & 255) << 8) | (col & 255)) * 4];
}
LVO_INLINE float unorm_grad(uint8_t b)
{
    return ((float)b / 255.0f) * 4.f - 1.f;    /* simplex.cl:124 ".xy * 4.f - 1.f" */
}

#define SKEW_F2 0.366025403784f   /* simplex.cl:103 */
#define SKEW_G2 0.211324865405f   /* simplex.cl:105 */

/* simplex.cl:99-157.  dot(a,b) of float2 := fmaf(a.y, b.y, a.x*b.x). */
LVO_INLINE float snoise2_impl(const lvo_world *w, float px, float py)
{
    const float s = (px + py) * SKEW_F2;
    const float ix = floorf(px + s), iy = floorf(py + s);
    const float t = (ix + iy) * SKEW_G2;
    const float x0 = px - (ix - t), y0 = py - (iy - t);
    const int ii = (int)ix, jj = (int)iy;
    float o1x, o1y;
    const uint8_t *g;
    float gx, gy, t0, t1, t2, n0, n1, n2, x1, y1, x2, y2;

    if (x0 > y0) { o1x = 1.f; o1y = 0.f; } else { o1x = 0.f; o1y = 1.f; }

    g = texel(w, ii, jj); gx = unorm_grad(g[0]); gy = unorm_grad(g[1]);
    t0 = 0.5f - fmaf(y0, y0, x0 * x0);
    if (t0 < 0.f) n0 = 0.f;
    else { t0 *= t0; n0 = t0 * t0 * fmaf(gy, y0, gx * x0); }

    x1 = x0 - o1x + SKEW_G2; y1 = y0 - o1y + SKEW_G2;
    g = texel(w, ii + (int)o1x, jj + (int)o1y); gx = unorm_grad(g[0]); gy = unorm_grad(g[1]);
    t1 = 0.5f - fmaf(y1, y1, x1 * x1);
    if (t1 < 0.f) n1 = 0.f;
    else { t1 *= t1; n1 = t1 * t1 * fmaf(gy, y1, gx * x1); }

    x2 = x0 - (1.f - 2.f * SKEW_G2); y2 = y0 - (1.f - 2.f * SKEW_G2);
    g = texel(w, ii + 1, jj + 1); gx = unorm_grad(g[0]); gy = unorm_grad(g[1]);
    t2 = 0.5f - fmaf(y2, y2, x2 * x2);
    if (t2 < 0.f) n2 = 0.f;
    else { t2 *= t2; n2 = t2 * t2 * fmaf(gy, y2, gx * x2); }

    return 70.f * (n0 + n1 + n2);
}

#define SKEW_F3 0.333333333333f   /* simplex.cl:162 */
#define SKEW_G3 0.166666666667f   /* simplex.cl:163 */

/* second lookup of snoise3: x coordinate = perm (a UNORM byte v/255, NOT
 * texel-centred): REPEAT + NEAREST gives column v for v<255 and 0 for 255 */
LVO_INLINE int perm_col(uint8_t v) { return v == 255 ? 0 : (int)v; }

/* simplex.cl:72-97 + 159-230.  dot of float3 := fmaf(z,z', fmaf(y,y', x*x')). */
LVO_INLINE float corner3(const lvo_world *w, int ci, int cj, int ck, float x, float y, float z)
{
    const uint8_t perm = texel(w, ci, cj)[3];
    const uint8_t *g = texel(w, perm_col(perm), ck);
    const float gx = unorm_grad(g[0]), gy = unorm_grad(g[1]), gz = unorm_grad(g[2]);
    /* "0.6 - dot(Pf, Pf)" (simplex.cl:184,196,208,220): the unsuffixed 0.6 is a double, so the
     * subtraction happens in double and is rounded once; 0.6f != 0.6, unlike snoise2's 0.5 */
    float t = (float)(0.6 - (double)fmaf(z, z, fmaf(y, y, x * x)));
    if (t < 0.f) return 0.f;
    t *= t;
    return t * t * fmaf(gz, z, fmaf(gy, y, gx * x));
}

LVO_INLINE float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

LVO_INLINE float snoise3_impl(const lvo_world *w, float px, float py, float pz)
{
    const float s = (px + py + pz) * SKEW_F3;
    const float ix = floorf(px + s), iy = floorf(py + s), iz = floorf(pz + s);
    const float t = (ix + iy + iz) * SKEW_G3;
    const float x0 = px - (ix - t), y0 = py - (iy - t), z0 = pz - (iz - t);
    const int ii = (int)ix, jj = (int)iy, kk = (int)iz;
    /* simplex(): step(edge, x) = x < edge ? 0 : 1 */
    const float isXy = (x0 < y0) ? 0.f : 1.f;     /* step(P.y, P.x) */
    const float isXz = (x0 < z0) ? 0.f : 1.f;     /* step(P.z, P.x) */
    const float isY  = (y0 < z0) ? 0.f : 1.f;     /* step(P.z, P.y) */
    float ox = isXy + isXz, oy = 1.f - isXy, oz = 1.f - isXz;
    float o1x, o1y, o1z, o2x, o2y, o2z, n0, n1, n2, n3;
    oy += isY; oz += 1.f - isY;
    o2x = clamp01(ox); o2y = clamp01(oy); o2z = clamp01(oz);
    o1x = clamp01(ox - 1.f); o1y = clamp01(oy - 1.f); o1z = clamp01(oz - 1.f);

    n0 = corner3(w, ii, jj, kk, x0, y0, z0);
    n1 = corner3(w, ii + (int)o1x, jj + (int)o1y, kk + (int)o1z,
                 x0 - o1x + SKEW_G3, y0 - o1y + SKEW_G3, z0 - o1z + SKEW_G3);
    n2 = corner3(w, ii + (int)o2x, jj + (int)o2y, kk + (int)o2z,
                 x0 - o2x + 2.f * SKEW_G3, y0 - o2y + 2.f * SKEW_G3, z0 - o2z + 2.f * SKEW_G3);
    n3 = corner3(w, ii + 1, jj + 1, kk + 1,
                 x0 - (1.f - 3.f * SKEW_G3), y0 - (1.f - 3.f * SKEW_G3), z0 - (1.f - 3.f * SKEW_G3));
    return 32.f * (n0 + n1 + n2 + n3);
}

/* noise.cl:8-36 (NOISE_SCALE 1) */
LVO_INLINE float basic_fractal(const lvo_world *w, int octaves, float frequency,
                               float lacunarity, float persistence, float px, float py)
{
    float noise = 0.f, amplitude = 1.f;
    int i;
    px = px * 1.f; py = py * 1.f;
    px *= frequency; py *= frequency;
    for (i = 0; i < octaves; i++) {
        noise += snoise2_impl(w, px, py) * amplitude;
        px *= lacunarity; py *= lacunarity;
        amplitude *= persistence;
    }
    return noise;
}

/* noise.cl:42-81; pow(frequency, -1.f) := 1.f / frequency */
LVO_INLINE float ridged_multifractal(const lvo_world *w, int octaves, float lacunarity,
                                     float gain, float offset, float px, float py)
{
    float signal, noise, weight, frequency = 1.f;
    int i;
    px = px * 1.f; py = py * 1.f;
    signal = snoise2_impl(w, px, py);
    signal = fabsf(signal);
    signal = offset - signal;
    signal *= signal;
    noise = signal;
    for (i = 0; i < octaves; i++) {
        float exponent;
        px *= lacunarity; py *= lacunarity;
        weight = signal * gain;
        weight = clamp01(weight);
        signal = snoise2_impl(w, px, py);
        signal = fabsf(signal);
        signal = offset - signal;
        signal *= weight;
        exponent = 1.f / frequency;
        frequency *= lacunarity;
        noise += signal * exponent;
    }
    noise *= (1.f / (float)octaves);
    return noise;
}

/* noise.cl:205-223 */
LVO_INLINE float terrain_impl(const lvo_world *w, float x, float z)
{
    const float px = x * (1.f / 2000.f), py = z * (1.f / 2000.f);
    float ridged, billow, noise, b2;
    ridged = 0.8f * ridged_multifractal(w, 7, 2.114352f, 1.5241f, 1.f, px, py);
    ridged = clamp01(ridged);
    billow = 0.6f * basic_fractal(w, 4, 0.24f, 1.8754f, 0.433f, -4.33f * px, 7.98f * py);
    billow = (0.5f * billow) + 0.5f;
    noise = billow * ridged;
    b2 = 0.6f * basic_fractal(w, 2, 0.63f, 2.2f, 0.15f, px, py);
    b2 = (b2 * 0.5f) + 0.5f;
    noise += b2;
    return noise;
}

/* BASELINE config 4: ridged 3-D fBm from snoise3, 4 octaves, lacunarity 2,
 * base frequency 1/16 per voxel; solid where fBm > threshold. */
LVO_INLINE float stress_impl(const lvo_world *w, float x, float y, float z)
{
    float qx = x * (1.f / 16.f), qy = y * (1.f / 16.f), qz = z * (1.f / 16.f);
    float f = 0.f, amp = 1.f;
    int o;
    for (o = 0; o < 4; o++) {
        float r = 1.f - fabsf(snoise3_impl(w, qx, qy, qz));
        r *= r;
        f += r * amp;
        qx *= 2.f; qy *= 2.f; qz *= 2.f;
        amp *= 0.5f;
    }
    f *= (1.f / 1.875f);
    return w->stressThreshold - f;
}

/* noise.cl:225-268: position.y - (MAX_TERRAIN_HEIGHT * noise), height 900
 * (compute.cpp:254,267) */
LVO_INLINE float density_impl(const lvo_world *w, float x, float y, float z)
{
    if (w->densityKind == LVO_DENSITY_STRESS)
        return stress_impl(w, x, y, z);
    return y - (900.f * terrain_impl(w, x, z));
}

LVO_HOT float lvo_snoise2(const lvo_world *w, float x, float y) { return snoise2_impl(w, x, y); }
LVO_HOT float lvo_snoise3(const lvo_world *w, float x, float y, float z) { return snoise3_impl(w, x, y, z); }
LVO_HOT float lvo_terrain(const lvo_world *w, float x, float z) { return terrain_impl(w, x, z); }
LVO_HOT float lvo_density(const lvo_world *w, float x, float y, float z) { return density_impl(w, x, y, z); }

/* ------------------------------------------------------------------------ */
/* a3: GenerateDefaultField -- density_field.cl:11-37                       */
/* ------------------------------------------------------------------------ */

LVO_INLINE int field_index(const lvo_world *w, int x, int y, int z)
{
    return x + (y * w->F) + (z * w->F * w->F);   /* shared_constants.cl:24-27 */
}

LVO_INLINE int sample_scale(const lvo_world *w, int size)
{
    return size / (w->V * LVO_LEAF_SIZE_SCALE);   /* compute_density_field.cpp:149 */
}

LVO_HOT void lvo_generate_field(const lvo_world *w, const int min[3], int size, int32_t *materials)
{
    const int sampleScale = sample_scale(w, size);
    const int ox = min[0] / LVO_LEAF_SIZE_SCALE, oy = min[1] / LVO_LEAF_SIZE_SCALE,
              oz = min[2] / LVO_LEAF_SIZE_SCALE;   /* LeafScaleVec, compute.cpp:569-577 */
    int x, y, z;
    for (z = 0; z < w->F; z++)
        for (y = 0; y < w->F; y++)
            for (x = 0; x < w->F; x++) {
                const float wx = (float)((x * sampleScale) + ox);
                const float wy = (float)((y * sampleScale) + oy);
                const float wz = (float)((z * sampleScale) + oz);
                const float density = density_impl(w, wx, wy, wz);
                materials[field_index(w, x, y, z)] = density < 0.f ? w->defaultMaterial : LVO_MATERIAL_AIR;
            }
}

/* ------------------------------------------------------------------------ */
/* a15: scan / compact / dedupe                                             */
/* ------------------------------------------------------------------------ */

/* compute.cpp:343-395 + scan.cl: result is the plain exclusive prefix sum;
 * returned total = data[n-1] + scan[n-1] */
int lvo_exclusive_scan(const int32_t *data, int32_t *scan, int count)
{
    int i, acc = 0;
    if (count <= 0) return 0;
    for (i = 0; i < count; i++) { scan[i] = acc; acc += data[i]; }
    return data[count - 1] + scan[count - 1];
}

/* compact.cl:4-16 + compute.cpp:424-442 */
int lvo_compact(const int32_t *values, const int32_t *valid, int count, int32_t *out)
{
    int i, n = 0;
    for (i = 0; i < count; i++)
        if (valid[i]) out[n++] = values[i];
    return n;
}

/* duplicate.cl:4-28 (the seed argument is unused there too) */
static uint32_t murmur_hash(uint32_t value)
{
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u, r1 = 15, r2 = 13, m = 5, n = 0xe6546b64u;
    uint32_t hash = value;
    hash *= c1;
    hash = (hash << r1) | (hash >> (32 - r1));
    hash *= c2;
    hash ^= value;
    hash = ((hash << r2) | (hash >> (32 - r2))) * m + n;
    hash ^= (hash >> 16);
    hash *= 0x85ebca6bu;
    hash ^= (hash >> 13);
    hash *= 0xc2b2ae35u;
    hash ^= (hash >> 16);
    return hash;
}

/* compute.cpp:446-543 + duplicate.cl:40-114.  Work-items are serialised in
 * id order (table[hash] = id is last-writer-wins in the reference, so any
 * serialisation is a valid outcome).  Output order is therefore one of the
 * reference's possible orders; only the SET is defined. */
int lvo_remove_duplicates(const int32_t *values, int count, int32_t *out)
{
    int prime, numItems = count, resultsSize = 0, i;
    int32_t *sequence, *table, *losers;
    if (count <= 0) return 0;
    prime = lvo_find_next_prime(count * 2);
    sequence = (int32_t *)malloc(sizeof(int32_t) * (size_t)count);
    losers = (int32_t *)malloc(sizeof(int32_t) * (size_t)count);
    table = (int32_t *)malloc(sizeof(int32_t) * (size_t)prime);
    memcpy(sequence, values, sizeof(int32_t) * (size_t)count);
    while (numItems > 0) {
        int numLosers = 0;
        for (i = 0; i < prime; i++) table[i] = -1;
        for (i = 0; i < numItems; i++)                        /* MapSequenceIndices */
            table[murmur_hash((uint32_t)sequence[i]) % (uint32_t)prime] = i;
        for (i = 0; i < prime; i++)                           /* ExtractWinners + compact */
            if (table[i] != -1) {
                const int32_t winner = sequence[table[i]];
                out[resultsSize++] = winner;
            }
        for (i = 0; i < prime; i++)                           /* MapSequenceValues */
            if (table[i] != -1) table[i] = sequence[table[i]];
        for (i = 0; i < numItems; i++) {                      /* ExtractLosers + compact */
            const int32_t value = sequence[i];
            if (table[murmur_hash((uint32_t)value) % (uint32_t)prime] != value)
                losers[numLosers++] = value;
        }
        if (numLosers == 0) break;
        numItems = numLosers;
        memcpy(sequence, losers, sizeof(int32_t) * (size_t)numLosers);
        /* the xor value is threaded through but never used by MurmurHash */
    }
    free(sequence); free(losers); free(table);
    return resultsSize;
}

/* ------------------------------------------------------------------------ */
/* a4: FindFieldEdges + CompactEdges -- density_field.cl:41-92              */
/* ------------------------------------------------------------------------ */

int lvo_find_edges(const lvo_world *w, const int32_t *materials, int32_t *edgeKeys)
{
    const int H = w->H;
    int x, y, z, i, n = 0;
    /* ascending (x + H*y + H*H*z)*3 + axis == the scan/compact order */
    for (z = 0; z < H; z++)
        for (y = 0; y < H; y++)
            for (x = 0; x < H; x++) {
                const int corner[4] = {
                    materials[field_index(w, x, y, z)],
                    materials[field_index(w, x + 1, y, z)],
                    materials[field_index(w, x, y + 1, z)],
                    materials[field_index(w, x, y, z + 1)],
                };
                const int voxelIndex = x | (y << w->shift) | (z << (w->shift * 2));
                for (i = 0; i < 3; i++) {
                    const int e = 1 + i;
                    const int signChange =
                        ((corner[0] != LVO_MATERIAL_AIR && corner[e] == LVO_MATERIAL_AIR) ||
                         (corner[0] == LVO_MATERIAL_AIR && corner[e] != LVO_MATERIAL_AIR)) ? 1 : 0;
                    if (signChange) edgeKeys[n++] = (voxelIndex << 2) | i;
                }
            }
    return n;
}

/* ------------------------------------------------------------------------ */
/* a5: FindEdgeIntersectionInfo -- density_field.cl:96-151                  */
/* ------------------------------------------------------------------------ */

LVO_INLINE float mixf(float a, float b, float t) { return a + (b - a) * t; }  /* OpenCL mix */

/* normalize(float3) := v * (1 / sqrt((x*x + y*y) + z*z)); zero vector -> itself */
LVO_INLINE void normalize3(float *x, float *y, float *z)
{
    const float lenSq = ((*x) * (*x) + (*y) * (*y)) + (*z) * (*z);
    if (lenSq == 0.f) return;
    {
        const float inv = 1.f / sqrtf(lenSq);
        *x *= inv; *y *= inv; *z *= inv;
    }
}

static const int EDGE_END_OFFSETS[3][3] = {{1,0,0},{0,1,0},{0,0,1}};

LVO_HOT void lvo_edge_info(const lvo_world *w, const int min[3], int size,
                           const int32_t *edgeKeys, int numEdges, lvo_f4 *edgeInfo)
{
    const int sampleScale = sample_scale(w, size);
    const int off[3] = { min[0] / LVO_LEAF_SIZE_SCALE, min[1] / LVO_LEAF_SIZE_SCALE,
                         min[2] / LVO_LEAF_SIZE_SCALE };
    int index;
    for (index = 0; index < numEdges; index++) {
        const int edge = edgeKeys[index];
        const int axisIndex = edge & 3;
        const int hermiteIndex = edge >> 2;
        const int lp[3] = { (hermiteIndex >> (w->shift * 0)) & w->mask,
                            (hermiteIndex >> (w->shift * 1)) & w->mask,
                            (hermiteIndex >> (w->shift * 2)) & w->mask };
        float p0[3], p1[3], p[3];
        float minValue = FLT_MAX, currentT = 0.f, t = 0.f, dx, dy, dz;
        const float h = 0.001f;
        int i, k;
        for (k = 0; k < 3; k++) {
            const int wp = (sampleScale * lp[k]) + off[k];
            p0[k] = (float)wp;
            p1[k] = (float)(wp + (sampleScale * EDGE_END_OFFSETS[axisIndex][k]));
        }
        for (i = 0; i <= 16; i++) {                     /* FIND_EDGE_INFO_STEPS 16 */
            float d;
            for (k = 0; k < 3; k++) p[k] = mixf(p0[k], p1[k], currentT);
            d = fabsf(density_impl(w, p[0], p[1], p[2]));
            if (d < minValue) { t = currentT; minValue = d; }
            currentT += (1.f / 16.f);                   /* FIND_EDGE_INFO_INCREMENT */
        }
        for (k = 0; k < 3; k++) p[k] = mixf(p0[k], p1[k], t);
        dx = density_impl(w, p[0] + h, p[1], p[2]) - density_impl(w, p[0] - h, p[1], p[2]);
        dy = density_impl(w, p[0], p[1] + h, p[2]) - density_impl(w, p[0], p[1] - h, p[2]);
        dz = density_impl(w, p[0], p[1], p[2] + h) - density_impl(w, p[0], p[1], p[2] - h);
        normalize3(&dx, &dy, &dz);
        edgeInfo[index].x = dx; edgeInfo[index].y = dy; edgeInfo[index].z = dz;
        edgeInfo[index].w = t;
    }
}

/* ------------------------------------------------------------------------ */
/* a6/a7: FindActiveVoxels + CompactVoxels -- octree.cl:34-223              */
/* ------------------------------------------------------------------------ */

static const int EDGE_MAP[12][2] = {      /* shared_constants.cl:4-9 */
    {0,4},{1,5},{2,6},{3,7},
    {0,2},{1,3},{4,6},{5,7},
    {0,1},{2,3},{4,5},{6,7} };
static const int CHILD_MIN_OFFSETS[8][3] = {   /* shared_constants.cl:11-22 */
    {0,0,0},{0,0,1},{0,1,0},{0,1,1},{1,0,0},{1,0,1},{1,1,0},{1,1,1} };

/* octree.cl:34-47 with nodeDepth == MAX_OCTREE_DEPTH */
static uint32_t code_for_position(const lvo_world *w, int px, int py, int pz)
{
    uint32_t code = 1;
    int depth;
    for (depth = w->depth - 1; depth >= 0; depth--) {
        const int x = (px >> depth) & 1, y = (py >> depth) & 1, z = (pz >> depth) & 1;
        const int c = (x << 2) | (y << 1) | z;
        code = (code << 3) | (uint32_t)c;
    }
    return code;
}

/* octree.cl:51-74 */
static void position_for_code(const lvo_world *w, uint32_t code, int pos[3])
{
    int msb = 0, nodeDepth, i;
    { uint32_t c = code; while (c) { msb++; c >>= 1; } }   /* 32 - clz */
    nodeDepth = msb / 3;
    pos[0] = pos[1] = pos[2] = 0;
    for (i = w->depth - nodeDepth; i < w->depth; i++) {
        const uint32_t c = code & 7;
        code >>= 3;
        pos[0] |= (int)((c >> 2) & 1) << i;
        pos[1] |= (int)((c >> 1) & 1) << i;
        pos[2] |= (int)((c >> 0) & 1) << i;
    }
}

/* octree.cl:80-138 */
static int find_dominant_material(const int m[8])
{
    int data[8], i, j, current, count = 1, maxCount = 0, maxMaterial = 0;
    for (i = 0; i < 8; i++) data[i] = m[i];
    for (i = 1; i < 8; i++) {
        const int tmp = data[i];
        for (j = i; j >= 1 && tmp < data[j - 1]; j--) data[j] = data[j - 1];
        data[j] = tmp;
    }
    current = data[0];
    for (i = 1; i < 8; i++) {
        const int mi = data[i];
        if (mi == LVO_MATERIAL_AIR || mi == LVO_MATERIAL_NONE) continue;
        if (current != mi) {
            if (count > maxCount) { maxCount = count; maxMaterial = current; }
            current = mi;
            count = 1;
        } else {
            count++;
        }
    }
    if (count > maxCount) maxMaterial = current;
    return maxMaterial;
}

int lvo_find_active_voxels(const lvo_world *w, const int32_t *materials,
                           uint32_t *codes, int32_t *edgeMasks, int32_t *matWords)
{
    const int V = w->V;
    int x, y, z, i, n = 0;
    /* ascending x + V*y + V*V*z == the scan/compact order (octree.cl:205-223) */
    for (z = 0; z < V; z++)
        for (y = 0; y < V; y++)
            for (x = 0; x < V; x++) {
                int cornerMaterials[8], cornerValues = 0, edgeList = 0;
                for (i = 0; i < 8; i++) {
                    cornerMaterials[i] = materials[field_index(w, x + CHILD_MIN_OFFSETS[i][0],
                        y + CHILD_MIN_OFFSETS[i][1], z + CHILD_MIN_OFFSETS[i][2])];
                    cornerValues |= ((cornerMaterials[i] == LVO_MATERIAL_AIR ? 0 : 1) << i);
                }
                for (i = 0; i < 12; i++) {
                    const int edgeStart = (cornerValues >> EDGE_MAP[i][0]) & 1;
                    const int edgeEnd = (cornerValues >> EDGE_MAP[i][1]) & 1;
                    const int signChange = (!edgeStart && edgeEnd) || (edgeStart && !edgeEnd);
                    edgeList |= (signChange << i);
                }
                if (cornerValues != 0 && cornerValues != 255) {
                    const int materialIndex = find_dominant_material(cornerMaterials);
                    codes[n] = code_for_position(w, x, y, z);
                    edgeMasks[n] = edgeList;
                    matWords[n] = (materialIndex << 8) | cornerValues;
                    n++;
                }
            }
    return n;
}

/* ------------------------------------------------------------------------ */
/* a9: cuckoo -- cuckoo.cl:3-104, compute_cuckoo.cpp:46-138, primes.cpp     */
/* ------------------------------------------------------------------------ */

static int is_prime(int x)   /* primes.cpp:9-29 */
{
    int o = 4, i = 5;
    for (;;) {
        const int q = x / i;
        if (q < i) return 1;
        if (x == (q * i)) return 0;
        o ^= 6;
        i += o;
    }
}

int lvo_find_next_prime(int n)   /* primes.cpp:32-59 */
{
    int k, i, o, x;
    if (n <= 2) return 2;
    else if (n == 3) return 3;
    else if (n <= 5) return 5;
    k = n / 6;
    i = n - (6 * k);
    o = i < 2 ? 1 : 5;
    x = (6 * k) + o;
    for (i = (3 + o) / 2; !is_prime(x); x += i) i ^= 6;
    return x;
}

#define CUCKOO_EMPTY (~0ULL)          /* compute_cuckoo.h:6-10 */
#define CUCKOO_STASH_HASH_INDEX 4
#define CUCKOO_STASH_SIZE 101
#define CUCKOO_MAX_ITERATIONS 32

/* the reference draws hash parameters from a default-seeded std::mt19937
 * through uniform_int_distribution(1<<15, 1<<30) (compute_cuckoo.cpp:46-47);
 * the mapping is implementation-defined and the values are unobservable, so
 * any generator over the same range restates it. */
static lvo_mt g_cuckooRng; static int g_cuckooRngInit = 0;
static uint32_t cuckoo_param(void)
{
    const uint32_t lo = 1u << 15, hi = 1u << 30;
    uint32_t v;
#ifdef _OPENMP
#pragma omp critical(lvo_cuckoo_rng)
#endif
    {
        if (!g_cuckooRngInit) { mt_seed(&g_cuckooRng, 5489u); g_cuckooRngInit = 1; }
        v = lo + mt_next(&g_cuckooRng) % (hi - lo + 1u);
    }
    return v;
}

/* cuckoo.cl:18-24.  NB "unsigned long h = a * key" multiplies two 32-bit
 * uints, so the product wraps to 32 bits BEFORE it is widened. */
static uint32_t cuckoo_hash(int whichHash, uint32_t key, uint32_t a, uint32_t b, uint32_t prime)
{
    const uint32_t p = 4294967291u;
    const uint64_t h = (uint64_t)(uint32_t)(a * key);
    const uint32_t mod = whichHash < CUCKOO_STASH_HASH_INDEX ? prime : CUCKOO_STASH_SIZE;
    return (uint32_t)(((h + b) % p) % mod);
}

int lvo_cuckoo_init(lvo_cuckoo *c, uint32_t tableSize)   /* compute_cuckoo.cpp:49-74 */
{
    uint32_t i;
    const uint32_t want = tableSize * 2 > 2048u ? tableSize * 2 : 2048u;   /* MIN_TABLE_SIZE */
    memset(c, 0, sizeof(*c));
    c->prime = (uint32_t)lvo_find_next_prime((int)want);
    c->table = (uint64_t *)malloc(sizeof(uint64_t) * c->prime);
    for (i = 0; i < c->prime; i++) c->table[i] = CUCKOO_EMPTY;
    for (i = 0; i < CUCKOO_STASH_SIZE; i++) c->stash[i] = CUCKOO_EMPTY;
    for (i = 0; i < 10; i++) c->params[i] = cuckoo_param();
    return 0;
}

void lvo_cuckoo_free(lvo_cuckoo *c) { free(c->table); c->table = NULL; }

/* cuckoo.cl:26-71 for one work-item.  Work-items are serialised in index
 * order (a valid interleaving of the atom_xchg chain).  The stash insert of
 * the reference indexes stash[] with a hash taken mod prime (cuckoo.cl:67-69,
 * out of bounds for a 101-entry stash); the restatement reports such a key
 * as not inserted, which makes the host loop rehash, as it would after any
 * failed insert. */
static int cuckoo_insert_one(lvo_cuckoo *c, uint32_t key, uint32_t value, int *stashUsed)
{
    uint64_t entry = ((uint64_t)value << 32) | key;
    uint32_t h = cuckoo_hash(0, key, c->params[0], c->params[1], c->prime);
    int i;
    for (i = 0; i < CUCKOO_MAX_ITERATIONS; i++) {
        const uint64_t old = c->table[h];      /* atom_xchg */
        c->table[h] = entry;
        entry = old;
        if (entry == CUCKOO_EMPTY) { *stashUsed = 0; return 1; }
        key = (uint32_t)(entry & 0xffffffffu);
        {
            const uint32_t h0 = cuckoo_hash(0, key, c->params[0], c->params[1], c->prime);
            const uint32_t h1 = cuckoo_hash(1, key, c->params[2], c->params[3], c->prime);
            const uint32_t h2 = cuckoo_hash(2, key, c->params[4], c->params[5], c->prime);
            const uint32_t h3 = cuckoo_hash(3, key, c->params[6], c->params[7], c->prime);
            if (h == h0) h = h1;
            else if (h == h1) h = h2;
            else if (h == h2) h = h3;
            else if (h == h3) h = h0;
        }
    }
    *stashUsed = 1;
    return 0;
}

/* compute_cuckoo.cpp:78-138: retry with fresh parameters until every key is in */
int lvo_cuckoo_insert_keys(lvo_cuckoo *c, const uint32_t *keys, uint32_t count)
{
    uint32_t insertedCount = 0, i;
    int first = 1;
    do {
        if (!first) {
            c->retries++;
            for (i = 0; i < 10; i++) c->params[i] = cuckoo_param();
            for (i = 0; i < c->prime; i++) c->table[i] = CUCKOO_EMPTY;
            for (i = 0; i < CUCKOO_STASH_SIZE; i++) c->stash[i] = CUCKOO_EMPTY;
            if (c->retries > 64) return -1;
        }
        first = 0;
        insertedCount = 0;
        for (i = 0; i < count; i++) {
            int stashUsed = 0;
            insertedCount += (uint32_t)cuckoo_insert_one(c, keys[i], i, &stashUsed);
        }
    } while (insertedCount < count);
    c->insertedKeys += (int)insertedCount;
    return 0;
}

/* cuckoo.cl:73-104 (stashUsed is always 0 here, see cuckoo_insert_one) */
uint32_t lvo_cuckoo_find(const lvo_cuckoo *c, uint32_t key)
{
    int i;
    for (i = 0; i < CUCKOO_STASH_HASH_INDEX; i++) {
        const uint32_t h = cuckoo_hash(i, key, c->params[i * 2 + 0], c->params[i * 2 + 1], c->prime);
        const uint64_t entry = c->table[h];
        if ((uint32_t)(entry & 0xffffffffu) == key) return (uint32_t)(entry >> 32);
    }
    return ~0u;
}

/* leven/src/cuckoo.h:14-165, the CPU table test_cuckoo.cpp exercises */
struct lvo_cpu_cuckoo {
    uint64_t *data; uint32_t size;
    uint64_t  stash[101];
    uint32_t  params[5][2];
    int       stashUsed, insertedKeys;
};

static uint32_t cpu_cuckoo_hash(const lvo_cpu_cuckoo *c, int whichHash, uint64_t key)
{
    const uint32_t mod = whichHash < 4 ? c->size : 101u;
    const uint32_t a = c->params[whichHash][0], b = c->params[whichHash][1];
    const uint32_t p = 4294967291u;
    const uint64_t h = key * a;        /* cuckoo.h:149: 64-bit product */
    return (uint32_t)(((h + b) % p) % mod);
}

lvo_cpu_cuckoo *lvo_cpu_cuckoo_create(int size, uint32_t seed)   /* cuckoo.h:18-39 */
{
    lvo_cpu_cuckoo *c = (lvo_cpu_cuckoo *)calloc(1, sizeof(*c));
    lvo_mt mt;
    uint32_t i;
    int a, b;
    c->size = (uint32_t)lvo_find_next_prime((int)((float)size * 2.f));
    c->data = (uint64_t *)malloc(sizeof(uint64_t) * c->size);
    for (i = 0; i < c->size; i++) c->data[i] = CUCKOO_EMPTY;
    for (i = 0; i < 101; i++) c->stash[i] = CUCKOO_EMPTY;
    mt_seed(&mt, seed);
    for (a = 0; a < 5; a++)
        for (b = 0; b < 2; b++)
            c->params[a][b] = (1u << 10) + mt_next(&mt) % ((1u << 20) - (1u << 10) + 1u);
    return c;
}

int lvo_cpu_cuckoo_insert(lvo_cpu_cuckoo *c, uint32_t key, uint32_t value)   /* cuckoo.h:41-79 */
{
    uint64_t entry = ((uint64_t)value << 32) | key;
    uint32_t h = cpu_cuckoo_hash(c, 0, key);
    int i;
    for (i = 0; i < 32; i++) {
        const uint64_t tmp = c->data[h]; c->data[h] = entry; entry = tmp;
        if (entry == CUCKOO_EMPTY) { c->insertedKeys++; return 1; }
        {
            const uint32_t k = (uint32_t)(entry & 0xffffffffu);
            const uint32_t h0 = cpu_cuckoo_hash(c, 0, k), h1 = cpu_cuckoo_hash(c, 1, k),
                           h2 = cpu_cuckoo_hash(c, 2, k), h3 = cpu_cuckoo_hash(c, 3, k);
            if (h == h0) h = h1;
            else if (h == h1) h = h2;
            else if (h == h2) h = h3;
            else if (h == h3) h = h0;
        }
    }
    c->stashUsed = 1;
    h = cpu_cuckoo_hash(c, 4, (uint32_t)(entry & 0xffffffffu));
    if (c->stash[h] == CUCKOO_EMPTY) { c->stash[h] = entry; c->insertedKeys++; return 1; }
    return 0;
}

int lvo_cpu_cuckoo_find(const lvo_cpu_cuckoo *c, uint32_t key, uint32_t *value)   /* cuckoo.h:81-105 */
{
    int i;
    for (i = 0; i < 4; i++) {
        const uint32_t h = cpu_cuckoo_hash(c, i, key);
        if ((uint32_t)(c->data[h] & 0xffffffffu) == key) { *value = (uint32_t)(c->data[h] >> 32); return 1; }
    }
    if (c->stashUsed) {
        const uint32_t h = cpu_cuckoo_hash(c, 4, key);
        if ((uint32_t)(c->stash[h] & 0xffffffffu) == key) { *value = (uint32_t)(c->stash[h] >> 32); return 1; }
    }
    return 0;
}

void lvo_cpu_cuckoo_destroy(lvo_cpu_cuckoo *c) { if (c) { free(c->data); free(c); } }

/* ------------------------------------------------------------------------ */
/* a8: CreateLeafNodes -- octree.cl:227-312, qef.cl:170-191,283-303         */
/* ------------------------------------------------------------------------ */

static void qef_add_point(lvo_qef *qef, lvo_f4 n, lvo_f4 p)   /* qef.cl:170-191 */
{
    float b;
    qef->ATA[0] += n.x * n.x;
    qef->ATA[1] += n.x * n.y;
    qef->ATA[2] += n.x * n.z;
    qef->ATA[3] += n.y * n.y;
    qef->ATA[4] += n.y * n.z;
    qef->ATA[5] += n.z * n.z;
    b = dot4(p.x, p.y, p.z, p.w, n.x, n.y, n.z, n.w);
    qef->ATb.x += n.x * b;
    qef->ATb.y += n.y * b;
    qef->ATb.z += n.z * b;
    qef->masspoint.x += p.x;
    qef->masspoint.y += p.y;
    qef->masspoint.z += p.z;
    qef->masspoint.w += 1.f;
}

int lvo_create_leaf_nodes(const lvo_world *w, int sampleScale,
                          const uint32_t *codes, const int32_t *edgeMasks, int numNodes,
                          const int32_t *edgeKeys, const lvo_f4 *edgeInfo, int numEdges,
                          lvo_qef *qefs, lvo_f4 *normals)
{
    lvo_cuckoo table;
    int index, rc;
    /* compute_octree.cpp:107-109 */
    lvo_cuckoo_init(&table, (uint32_t)numEdges);
    rc = lvo_cuckoo_insert_keys(&table, (const uint32_t *)edgeKeys, (uint32_t)numEdges);
    if (rc < 0) { lvo_cuckoo_free(&table); return rc; }

    for (index = 0; index < numNodes; index++) {
        int position[3], i, edgeCount = 0;
        lvo_f4 edgePositions[12], edgeNormals[12], normal = {0.f, 0.f, 0.f, 0.f};
        lvo_qef qef;
        const int edgeList = edgeMasks[index];
        position_for_code(w, codes[index], position);
        for (i = 0; i < 12; i++) {
            const int active = (edgeList >> i) & 1;
            int e0, e1, axis, hx, hy, hz;
            uint32_t edgeIndex, dataIndex;
            float p0[3], p1[3];
            if (!active) continue;
            e0 = EDGE_MAP[i][0]; e1 = EDGE_MAP[i][1];   /* EDGE_VERTEX_MAP, octree.cl:227-232 */
            p0[0] = (float)position[0] + (float)CHILD_MIN_OFFSETS[e0][0];
            p0[1] = (float)position[1] + (float)CHILD_MIN_OFFSETS[e0][1];
            p0[2] = (float)position[2] + (float)CHILD_MIN_OFFSETS[e0][2];
            p1[0] = (float)position[0] + (float)CHILD_MIN_OFFSETS[e1][0];
            p1[1] = (float)position[1] + (float)CHILD_MIN_OFFSETS[e1][1];
            p1[2] = (float)position[2] + (float)CHILD_MIN_OFFSETS[e1][2];
            axis = i / 4;
            hx = position[0] + CHILD_MIN_OFFSETS[e0][0];
            hy = position[1] + CHILD_MIN_OFFSETS[e0][1];
            hz = position[2] + CHILD_MIN_OFFSETS[e0][2];
            edgeIndex = (((uint32_t)hx | ((uint32_t)hy << w->shift) | ((uint32_t)hz << (w->shift * 2))) << 2) | (uint32_t)axis;
            dataIndex = lvo_cuckoo_find(&table, edgeIndex);
            if (dataIndex != ~0u) {
                const lvo_f4 edgeData = edgeInfo[dataIndex];
                edgePositions[edgeCount].x = (float)sampleScale * mixf(p0[0], p1[0], edgeData.w);
                edgePositions[edgeCount].y = (float)sampleScale * mixf(p0[1], p1[1], edgeData.w);
                edgePositions[edgeCount].z = (float)sampleScale * mixf(p0[2], p1[2], edgeData.w);
                edgePositions[edgeCount].w = (float)sampleScale * mixf(0.f, 0.f, edgeData.w);
                edgeNormals[edgeCount].x = edgeData.x;
                edgeNormals[edgeCount].y = edgeData.y;
                edgeNormals[edgeCount].z = edgeData.z;
                edgeNormals[edgeCount].w = 0.f;
                edgeCount++;
            }
        }
        /* qef_create_from_points, qef.cl:283-303 */
        memset(&qef, 0, sizeof(qef));
        for (i = 0; i < edgeCount; i++) qef_add_point(&qef, edgeNormals[i], edgePositions[i]);
        {
            const float cnt = qef.masspoint.w;
            qef.masspoint.x /= cnt; qef.masspoint.y /= cnt; qef.masspoint.z /= cnt; qef.masspoint.w /= cnt;
        }
        qefs[index] = qef;
        for (i = 0; i < edgeCount; i++) {
            normal.x += edgeNormals[i].x; normal.y += edgeNormals[i].y;
            normal.z += edgeNormals[i].z; normal.w += edgeNormals[i].w;
            normal.w += 1.f;
        }
        normal.x /= normal.w; normal.y /= normal.w; normal.z /= normal.w;
        normal.w = 0.f;
        normals[index] = normal;
    }
    lvo_cuckoo_free(&table);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* a10: SolveQEFs -- octree.cl:316-331, qef.cl:16-144,239-256               */
/* ------------------------------------------------------------------------ */

#define SVD_NUM_SWEEPS 10
#define PSUEDO_INVERSE_THRESHOLD 0.1f

static void givens_coeffs_sym(float a_pp, float a_pq, float a_qq, float *c, float *s)   /* qef.cl:31-42 */
{
    float tau, stt, tan_;
    if (a_pq == 0.f) { *c = 1.f; *s = 0.f; return; }
    tau = (a_qq - a_pp) / (2.f * a_pq);
    stt = sqrtf(1.f + tau * tau);
    tan_ = 1.f / ((tau >= 0.f) ? (tau + stt) : (tau - stt));
    *c = 1.f / sqrtf(1.f + tan_ * tan_);     /* rsqrt := 1/sqrt */
    *s = tan_ * (*c);
}

static void svd_rotate_xy(float *x, float *y, float c, float s)   /* qef.cl:44-48 */
{
    const float u = *x, v = *y;
    *x = c * u - s * v;
    *y = s * u + c * v;
}

static void svd_rotateq_xy(float *x, float *y, float *a, float c, float s)   /* qef.cl:50-56 */
{
    const float cc = c * c, ss = s * s;
    /* "2.0 * c * s * (*a)": the literal is a double in OpenCL C */
    const float mx = (float)(2.0 * (double)c * (double)s * (double)(*a));
    const float u = *x, v = *y;
    *x = cc * u - mx + ss * v;
    *y = ss * u + mx + cc * v;
}

static void svd_rotate(float vtav[3][3], float v[3][3], int a, int b)   /* qef.cl:58-86 */
{
    float c, s, x, y, z;
    if (vtav[a][b] == 0.0f) return;
    givens_coeffs_sym(vtav[a][a], vtav[a][b], vtav[b][b], &c, &s);
    x = vtav[a][a]; y = vtav[b][b]; z = vtav[a][b];
    svd_rotateq_xy(&x, &y, &z, c, s);
    vtav[a][a] = x; vtav[b][b] = y; vtav[a][b] = z;
    x = vtav[0][3 - b]; y = vtav[1 - a][2];
    svd_rotate_xy(&x, &y, c, s);
    vtav[0][3 - b] = x; vtav[1 - a][2] = y;
    vtav[a][b] = 0.0f;
    x = v[0][a]; y = v[0][b]; svd_rotate_xy(&x, &y, c, s); v[0][a] = x; v[0][b] = y;
    x = v[1][a]; y = v[1][b]; svd_rotate_xy(&x, &y, c, s); v[1][a] = x; v[1][b] = y;
    x = v[2][a]; y = v[2][b]; svd_rotate_xy(&x, &y, c, s); v[2][a] = x; v[2][b] = y;
}

static float svd_invdet(float x, float tol)   /* qef.cl:107-109 (double 1.0/x) */
{
    const double inv = 1.0 / (double)x;
    return (fabsf(x) < tol || fabs(inv) < (double)tol) ? 0.0f : (float)inv;
}

static void svd_vmul_sym(lvo_f4 *result, const float A[6], lvo_f4 v)   /* qef.cl:146-152 */
{
    result->x = dot4(A[0], A[1], A[2], 0.f, v.x, v.y, v.z, v.w);   /* the x row is written with dot(), y and z are not */
    result->y = A[1] * v.x + A[3] * v.y + A[4] * v.z;
    result->z = A[2] * v.x + A[4] * v.y + A[5] * v.z;
}

static void svd_solve_ATA_ATb(const float ATA[6], lvo_f4 ATb, lvo_f4 *x)   /* qef.cl:88-144 */
{
    float V[3][3] = {{1.f,0.f,0.f},{0.f,1.f,0.f},{0.f,0.f,1.f}};
    float vtav[3][3], o[3][3], d0, d1, d2;
    int i;
    vtav[0][0] = ATA[0]; vtav[0][1] = ATA[1]; vtav[0][2] = ATA[2];
    vtav[1][0] = 0.f;    vtav[1][1] = ATA[3]; vtav[1][2] = ATA[4];
    vtav[2][0] = 0.f;    vtav[2][1] = 0.f;    vtav[2][2] = ATA[5];
    for (i = 0; i < SVD_NUM_SWEEPS; ++i) {
        svd_rotate(vtav, V, 0, 1);
        svd_rotate(vtav, V, 0, 2);
        svd_rotate(vtav, V, 1, 2);
    }
    d0 = svd_invdet(vtav[0][0], PSUEDO_INVERSE_THRESHOLD);
    d1 = svd_invdet(vtav[1][1], PSUEDO_INVERSE_THRESHOLD);
    d2 = svd_invdet(vtav[2][2], PSUEDO_INVERSE_THRESHOLD);
    {   /* svd_pseudoinverse, qef.cl:111-125 */
        int r, c;
        for (r = 0; r < 3; r++)
            for (c = 0; c < 3; c++)
                o[r][c] = V[r][0] * d0 * V[c][0] + V[r][1] * d1 * V[c][1] + V[r][2] * d2 * V[c][2];
    }
    /* svd_mul_matrix_vec, qef.cl:23-29 */
    x->x = dot4(o[0][0], o[0][1], o[0][2], 0.f, ATb.x, ATb.y, ATb.z, ATb.w);
    x->y = dot4(o[1][0], o[1][1], o[1][2], 0.f, ATb.x, ATb.y, ATb.z, ATb.w);
    x->z = dot4(o[2][0], o[2][1], o[2][2], 0.f, ATb.x, ATb.y, ATb.z, ATb.w);
    x->w = 0.f;
}

void lvo_solve_qefs(const int min[3], const lvo_qef *qefs, int numNodes, lvo_f4 *positions)
{
    int index;
    for (index = 0; index < numNodes; index++) {
        lvo_qef qef = qefs[index];
        lvo_f4 pos = {0.f, 0.f, 0.f, 0.f}, A_mp = {0.f, 0.f, 0.f, 0.f};
        /* qef_solve, qef.cl:239-256 */
        const float d = fmaxf(qef.masspoint.w, 1.f);
        qef.masspoint.x /= d; qef.masspoint.y /= d; qef.masspoint.z /= d; qef.masspoint.w /= d;
        svd_vmul_sym(&A_mp, qef.ATA, qef.masspoint);
        A_mp.x = qef.ATb.x - A_mp.x; A_mp.y = qef.ATb.y - A_mp.y;
        A_mp.z = qef.ATb.z - A_mp.z; A_mp.w = qef.ATb.w - A_mp.w;
        svd_solve_ATA_ATb(qef.ATA, A_mp, &pos);
        pos.x += qef.masspoint.x; pos.y += qef.masspoint.y; pos.z += qef.masspoint.z;
        /* octree.cl:327-328; worldSpaceOffset = chunk min in world units (compute_octree.cpp:131) */
        pos.x = (pos.x * (float)LVO_LEAF_SIZE_SCALE) + (float)min[0];
        pos.y = (pos.y * (float)LVO_LEAF_SIZE_SCALE) + (float)min[1];
        pos.z = (pos.z * (float)LVO_LEAF_SIZE_SCALE) + (float)min[2];
        pos.w = 1.f;
        positions[index] = pos;
    }
}

/* ------------------------------------------------------------------------ */
/* a11: GenerateMesh + CompactMeshTriangles -- octree.cl:335-464            */
/* ------------------------------------------------------------------------ */

static const int EDGE_NODE_OFFSETS[3][4][3] = {   /* octree.cl:376-381 */
    {{0,0,0},{0,0,1},{0,1,0},{0,1,1}},
    {{0,0,0},{1,0,0},{0,0,1},{1,0,1}},
    {{0,0,0},{0,1,0},{1,0,0},{1,1,0}} };

int lvo_generate_mesh(const lvo_world *w, const uint32_t *codes, const int32_t *matWords, int numNodes,
                      int32_t *indices)
{
    lvo_cuckoo table;
    int index, numQuads = 0;
    if (numNodes <= 0) return 0;
    /* compute_octree.cpp:143-144: node code -> node index */
    lvo_cuckoo_init(&table, (uint32_t)numNodes);
    if (lvo_cuckoo_insert_keys(&table, codes, (uint32_t)numNodes) < 0) { lvo_cuckoo_free(&table); return -1; }

    /* compaction keeps ascending (node, axis) order */
    for (index = 0; index < numNodes; index++) {
        int pos[3], axis;
        int nodeIndices[4] = { ~0, ~0, ~0, ~0 };
        position_for_code(w, codes[index], pos);
        for (axis = 0; axis < 3; axis++) {
            const int a = pos[(axis + 1) % 3], b = pos[(axis + 2) % 3];
            const int isEdgeVoxel = a == (w->V - 1) || b == (w->V - 1);
            int n;
            if (isEdgeVoxel) continue;
            nodeIndices[0] = index;
            for (n = 1; n < 4; n++) {
                const uint32_t c = code_for_position(w, pos[0] + EDGE_NODE_OFFSETS[axis][n][0],
                    pos[1] + EDGE_NODE_OFFSETS[axis][n][1], pos[2] + EDGE_NODE_OFFSETS[axis][n][2]);
                nodeIndices[n] = (int)lvo_cuckoo_find(&table, c);
            }
            if (nodeIndices[1] != ~0 && nodeIndices[2] != ~0 && nodeIndices[3] != ~0) {
                /* ProcessEdge, octree.cl:335-372 */
                const int edge = (axis * 4) + 3;
                const int c1 = EDGE_MAP[edge][0], c2 = EDGE_MAP[edge][1];
                const int corners = matWords[index] & 0xff;
                const int m1 = (corners >> c1) & 1, m2 = (corners >> c2) & 1;
                const int signChange = (m1 && !m2) || (!m1 && m2);
                if (signChange) {
                    static const int order[2][6] = {{0,1,3,0,3,2},{0,3,1,0,2,3}};
                    const int flip = m1 != 0 ? 1 : 0;
                    int k;
                    for (k = 0; k < 6; k++) indices[numQuads * 6 + k] = nodeIndices[order[flip][k]];
                    numQuads++;
                }
            }
        }
    }
    lvo_cuckoo_free(&table);
    return numQuads * 2;   /* compute_octree.cpp:234 */
}

/* clipmap.cpp:329-352, called with clipmapNodeSize / CLIPMAP_LEAF_SIZE (compute_octree.cpp:252) */
void lvo_colour_for_size(int size, float rgb[3])
{
    const int minLeafSize = size / (LVO_LEAF_SIZE_SCALE * 64);
    switch (minLeafSize) {
    case 1:  rgb[0] = 0.3f; rgb[1] = 0.1f; rgb[2] = 0.f;  break;
    case 2:  rgb[0] = 0.f;  rgb[1] = 0.f;  rgb[2] = 0.5f; break;
    case 4:  rgb[0] = 0.f;  rgb[1] = 0.5f; rgb[2] = 0.5f; break;
    case 8:  rgb[0] = 0.5f; rgb[1] = 0.f;  rgb[2] = 0.5f; break;
    case 16: rgb[0] = 0.f;  rgb[1] = 0.5f; rgb[2] = 0.f;  break;
    default: rgb[0] = 0.5f; rgb[1] = 0.f;  rgb[2] = 0.f;  break;
    }
}

/* a12: GenerateMeshVertexBuffer -- octree.cl:468-487 */
void lvo_vertex_buffer(const lvo_f4 *positions, const lvo_f4 *normals, const int32_t *matWords,
                       int numNodes, int size, lvo_vertex *out)
{
    float rgb[3];
    int i;
    lvo_colour_for_size(size, rgb);
    for (i = 0; i < numNodes; i++) {
        out[i].xyz = positions[i];
        out[i].normal = normals[i];
        out[i].colour.x = rgb[0]; out[i].colour.y = rgb[1]; out[i].colour.z = rgb[2];
        out[i].colour.w = (float)(matWords[i] >> 8);
    }
}

/* a14: FindSeamNodes + ExtractSeamNodeInfo -- octree.cl:506-551 */
int lvo_seam_nodes(const lvo_world *w, const uint32_t *codes, const int32_t *matWords,
                   const lvo_f4 *positions, const lvo_f4 *normals, int numNodes, lvo_seam_node *out)
{
    int i, n = 0;
    for (i = 0; i < numNodes; i++) {
        int p[3];
        position_for_code(w, codes[i], p);
        if ((p[0] == 0 || p[0] == w->V - 1) | (p[1] == 0 || p[1] == w->V - 1) | (p[2] == 0 || p[2] == w->V - 1)) {
            out[n].localspaceMin.x = p[0]; out[n].localspaceMin.y = p[1]; out[n].localspaceMin.z = p[2];
            out[n].localspaceMin.w = matWords[i];
            out[n].position = positions[i];
            out[n].normal = normals[i];
            n++;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------ */
/* a16: CSG -- apply_csg_operation.cl, compute_csg.cpp                      */
/* ------------------------------------------------------------------------ */

LVO_INLINE float length3(float x, float y, float z) { return sqrtf((x * x + y * y) + z * z); }

/* apply_csg_operation.cl:42-82; hg_sdf.glsl:164-166,218-221,460-463 */
float lvo_brush_density(float x, float y, float z, const lvo_csg_op *op)
{
    const float lx = x - op->origin[0], ly = y - op->origin[1], lz = z - op->origin[2];
    if (op->brushShape == 0) {
        /* pR(&pxz, a): p = cos(a)*p + sin(a)*(p.y, -p.x) on (x, z) */
        const float c = cosf(op->rotateY), s = sinf(op->rotateY);
        const float rx = c * lx + s * lz;
        const float rz = c * lz + s * (-lx);
        /* fBox: length(max(d,0)) + vmax3(min(d,0)) */
        const float dx = fabsf(rx) - op->dimensions[0];
        const float dy = fabsf(ly) - op->dimensions[1];
        const float dz = fabsf(rz) - op->dimensions[2];
        const float outside = length3(fmaxf(dx, 0.f), fmaxf(dy, 0.f), fmaxf(dz, 0.f));
        const float inside = fmaxf(fmaxf(fminf(dx, 0.f), fminf(dy, 0.f)), fminf(dz, 0.f));
        return outside + inside;
    }
    return length3(lx, ly, lz) - op->dimensions[0];   /* Density_Sphere, radius = dimensions.x */
}

/* apply_csg_operation.cl:114-140 */
static int brush_material(float x, float y, float z, int numOps, const lvo_csg_op *ops, int material)
{
    int m = material, i;
    for (i = 0; i < numOps; i++) {
        const int operationMaterial[2] = { ops[i].material, LVO_MATERIAL_AIR };
        const float d = lvo_brush_density(x, y, z, &ops[i]);
        if (d <= 0.f) m = operationMaterial[ops[i].type];
    }
    return m;
}

/* apply_csg_operation.cl:86-110 */
static float brush_zero_crossing(const float p0[3], const float p1[3], int numOps, const lvo_csg_op *ops)
{
    float minDensity = FLT_MAX, crossing = 0.f, t;
    for (t = 0.f; t <= 1.f; t += (1.f / 16.f)) {
        const float px = mixf(p0[0], p1[0], t), py = mixf(p0[1], p1[1], t), pz = mixf(p0[2], p1[2], t);
        int i;
        for (i = 0; i < numOps; i++) {
            const float d = fabsf(lvo_brush_density(px, py, pz, &ops[i]));
            if (d < minDensity) { crossing = t; minDensity = d; }
        }
    }
    return crossing;
}

/* apply_csg_operation.cl:144-175 */
static void brush_normal(float x, float y, float z, int numOps, const lvo_csg_op *ops, float n[3])
{
    int i;
    n[0] = n[1] = n[2] = 0.f;
    for (i = 0; i < numOps; i++) {
        const lvo_csg_op *op = &ops[i];
        const float d = lvo_brush_density(x, y, z, op);
        const float h = 0.001f;
        float dx, dy, dz, flip;
        if (d > 0.f) continue;
        dx = lvo_brush_density(x + h, y, z, op) - lvo_brush_density(x - h, y, z, op);
        dy = lvo_brush_density(x, y + h, z, op) - lvo_brush_density(x, y - h, z, op);
        dz = lvo_brush_density(x, y, z + h, op) - lvo_brush_density(x, y, z - h, op);
        flip = op->type == 0 ? 1.f : -1.f;
        normalize3(&dx, &dy, &dz);
        n[0] = flip * dx; n[1] = flip * dy; n[2] = flip * dz;
    }
}

/* ApplyCSGOperations, compute_csg.cpp:11-220.  Edges whose key names a
 * sample outside the Hermite grid (FindUpdatedEdges emits them for field
 * samples at index H, apply_csg_operation.cl:253-324, and FilterValidEdges
 * then reads past the row) are dropped: no voxel ever looks them up, so
 * they are unobservable.  RemoveDuplicates' output order is arbitrary in the
 * reference; here the unique edges are kept in first-seen order. */
static int apply_csg(const lvo_world *w, const lvo_csg_op *ops, int numOps,
                     const int min[3], int size, lvo_field *field)
{
    const int F = w->F, H = w->H;
    const int sampleScale = sample_scale(w, size);
    const int off[3] = { min[0] / LVO_LEAF_SIZE_SCALE, min[1] / LVO_LEAF_SIZE_SCALE,
                         min[2] / LVO_LEAF_SIZE_SCALE };
    const int fieldBufferSize = F * F * F;
    int32_t *updatedPos, *updatedMat, *generated, *invalidated, *created;
    int numUpdated = 0, numGenerated = 0, numInvalidated = 0, numCreated = 0;
    int x, y, z, i, k;

    if (numOps == 0) return 0;

    /* Apply: CSG_HermiteIndices + CompactPoints + UpdateFieldMaterials (:179-241) */
    updatedPos = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)fieldBufferSize);
    updatedMat = (int32_t *)malloc(sizeof(int32_t) * (size_t)fieldBufferSize);
    for (z = 0; z < F; z++)
        for (y = 0; y < F; y++)
            for (x = 0; x < F; x++) {
                const int index = field_index(w, x, y, z);
                const int oldMaterial = field->materials[index];
                const float wx = (float)(off[0] + sampleScale * x);
                const float wy = (float)(off[1] + sampleScale * y);
                const float wz = (float)(off[2] + sampleScale * z);
                const int material = brush_material(wx, wy, wz, numOps, ops, oldMaterial);
                if (material != oldMaterial) {
                    updatedPos[numUpdated * 3 + 0] = x; updatedPos[numUpdated * 3 + 1] = y;
                    updatedPos[numUpdated * 3 + 2] = z;
                    updatedMat[numUpdated] = material;
                    numUpdated++;
                }
            }
    if (numUpdated <= 0) { free(updatedPos); free(updatedMat); return 0; }
    for (i = 0; i < numUpdated; i++)
        field->materials[field_index(w, updatedPos[i * 3], updatedPos[i * 3 + 1], updatedPos[i * 3 + 2])] = updatedMat[i];

    /* Filter: FindUpdatedEdges (:253-324) + RemoveInvalidIndices + RemoveDuplicates */
    generated = (int32_t *)malloc(sizeof(int32_t) * 6 * (size_t)numUpdated);
    for (i = 0; i < numUpdated; i++) {
        const int p[3] = { updatedPos[i * 3], updatedPos[i * 3 + 1], updatedPos[i * 3 + 2] };
        for (k = 0; k < 3; k++) {   /* the three edges leaving the sample */
            const int q[3] = { p[0] + (k == 0), p[1] + (k == 1), p[2] + (k == 2) };
            if (p[0] < H && p[1] < H && p[2] < H && q[0] <= H && q[1] <= H && q[2] <= H)
                generated[numGenerated++] = ((p[0] | (p[1] << w->shift) | (p[2] << (w->shift * 2))) << 2) | k;
        }
        for (k = 0; k < 3; k++) {   /* the three edges arriving at the sample */
            if (p[k] > 0) {
                const int q[3] = { p[0] - (k == 0), p[1] - (k == 1), p[2] - (k == 2) };
                if (q[0] < H && q[1] < H && q[2] < H)
                    generated[numGenerated++] = ((q[0] | (q[1] << w->shift) | (q[2] << (w->shift * 2))) << 2) | k;
            }
        }
    }
    invalidated = (int32_t *)malloc(sizeof(int32_t) * (size_t)(numGenerated > 0 ? numGenerated : 1));
    numInvalidated = lvo_remove_duplicates(generated, numGenerated, invalidated);

    /* FilterValidEdges (:342-375) + compact = created edges */
    created = (int32_t *)malloc(sizeof(int32_t) * (size_t)(numInvalidated > 0 ? numInvalidated : 1));
    for (i = 0; i < numInvalidated; i++) {
        const int key = invalidated[i];
        const int axis = key & 3, e = key >> 2;
        const int px = e & w->mask, py = (e >> w->shift) & w->mask, pz = (e >> (w->shift * 2)) & w->mask;
        const int m0 = field->materials[field_index(w, px, py, pz)];
        const int m1 = field->materials[field_index(w, px + (axis == 0), py + (axis == 1), pz + (axis == 2))];
        const int signChange = (m0 == LVO_MATERIAL_AIR && m1 != LVO_MATERIAL_AIR) ||
                               (m1 == LVO_MATERIAL_AIR && m0 != LVO_MATERIAL_AIR);
        if (signChange) created[numCreated++] = key;
    }

    /* Prune: PruneFieldEdges + CompactFieldEdges (:400-437, compute_csg.cpp:142-177).
     * The reference keeps the old list when EVERY old edge was invalidated
     * (numPrunedEdges == 0 skips the swap, compute_csg.cpp:160), which leaves
     * duplicate keys with a hash-order-dependent winner.  Restated as the
     * evident intent: an empty surviving list.  See DESIGN.md "deviations". */
    if (numInvalidated > 0 && field->numEdges > 0) {
        int kept = 0;
        for (i = 0; i < field->numEdges; i++) {
            int invalid = 0;
            for (k = 0; k < numInvalidated; k++) invalid |= (invalidated[k] == field->edgeKeys[i]);
            if (!invalid) {
                field->edgeKeys[kept] = field->edgeKeys[i];
                field->edgeInfo[kept] = field->edgeInfo[i];
                kept++;
            }
        }
        field->numEdges = kept;
    }

    /* Create: CSG FindEdgeIntersectionInfo (:443-477) and append (compute_csg.cpp:180-217) */
    if (numCreated > 0) {
        const int newSize = field->numEdges + numCreated;
        field->edgeKeys = (int32_t *)realloc(field->edgeKeys, sizeof(int32_t) * (size_t)newSize);
        field->edgeInfo = (lvo_f4 *)realloc(field->edgeInfo, sizeof(lvo_f4) * (size_t)newSize);
        for (i = 0; i < numCreated; i++) {
            const int key = created[i];
            const int axis = key & 3, voxelIndex = key >> 2;
            const int lp[3] = { voxelIndex & w->mask, (voxelIndex >> w->shift) & w->mask,
                                (voxelIndex >> (w->shift * 2)) & w->mask };
            const int e0 = EDGE_MAP[4 * axis][0], e1 = EDGE_MAP[4 * axis][1];
            float p0[3], p1[3], n[3], t, px, py, pz;
            for (k = 0; k < 3; k++) {
                const int wp = (sampleScale * lp[k]) + off[k];
                p0[k] = (float)(wp + CHILD_MIN_OFFSETS[e0][k]);
                p1[k] = (float)(wp + (sampleScale * CHILD_MIN_OFFSETS[e1][k]));
            }
            t = brush_zero_crossing(p0, p1, numOps, ops);
            px = mixf(p0[0], p1[0], t); py = mixf(p0[1], p1[1], t); pz = mixf(p0[2], p1[2], t);
            brush_normal(px, py, pz, numOps, ops, n);
            field->edgeKeys[field->numEdges + i] = key;
            field->edgeInfo[field->numEdges + i].x = n[0];
            field->edgeInfo[field->numEdges + i].y = n[1];
            field->edgeInfo[field->numEdges + i].z = n[2];
            field->edgeInfo[field->numEdges + i].w = t;
        }
        field->numEdges = newSize;
    }
    free(updatedPos); free(updatedMat); free(generated); free(invalidated); free(created);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* a17 + orchestration                                                      */
/* ------------------------------------------------------------------------ */

lvo_world *lvo_world_create(const uint8_t *rgba, int defaultMaterial, int voxelsPerChunk)
{
    lvo_world *w = (lvo_world *)calloc(1, sizeof(*w));
    int l = 0;
    memcpy(w->image, rgba, sizeof(w->image));
    w->defaultMaterial = defaultMaterial;
    w->V = voxelsPerChunk; w->H = voxelsPerChunk + 1; w->F = voxelsPerChunk + 2;
    while ((1 << (l + 1)) <= voxelsPerChunk) l++;    /* glm::log2 (integer) */
    w->depth = l;                                    /* MAX_OCTREE_DEPTH, compute.cpp:271 */
    w->shift = l + 1;                                /* compute.cpp:251 */
    w->mask = (1 << w->shift) - 1;
    w->densityKind = LVO_DENSITY_TERRAIN;
    w->stressThreshold = 0.5f;
    return w;
}

void lvo_world_set_density(lvo_world *w, int kind, float stressThreshold)
{
    w->densityKind = kind; w->stressThreshold = stressThreshold;
}

static void field_release(lvo_field *f) { free(f->edgeKeys); free(f->edgeInfo); free(f->materials); }
static void octree_release(lvo_octree *o)
{
    free(o->codes); free(o->edgeMasks); free(o->matWords); free(o->qefs);
    free(o->positions); free(o->normals); free(o->materials); free(o->edgeKeys); free(o->edgeInfo);
}

void lvo_world_destroy(lvo_world *w)
{
    int i;
    if (!w) return;
    for (i = 0; i < w->numFields; i++) field_release(&w->fields[i]);
    for (i = 0; i < w->numOctrees; i++) octree_release(&w->octrees[i]);
    free(w->fields); free(w->octrees); free(w->ops); free(w->opAABB);
    free(w);
}

static int key_eq(const int a[3], int asize, const int b[3], int bsize)
{
    return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && asize == bsize;
}

static lvo_field *find_field(lvo_world *w, const int min[3], int size)
{
    int i;
    for (i = 0; i < w->numFields; i++)
        if (key_eq(w->fields[i].min, w->fields[i].size, min, size)) return &w->fields[i];
    return NULL;
}

static void field_copy(lvo_world *w, lvo_field *dst, const lvo_field *src)
{
    const size_t nF3 = (size_t)w->F * w->F * w->F;
    *dst = *src;
    dst->materials = (int32_t *)malloc(sizeof(int32_t) * nF3);
    memcpy(dst->materials, src->materials, sizeof(int32_t) * nF3);
    dst->edgeKeys = (int32_t *)malloc(sizeof(int32_t) * (size_t)(src->numEdges + 1));
    dst->edgeInfo = (lvo_f4 *)malloc(sizeof(lvo_f4) * (size_t)(src->numEdges + 1));
    if (src->numEdges) {
        memcpy(dst->edgeKeys, src->edgeKeys, sizeof(int32_t) * (size_t)src->numEdges);
        memcpy(dst->edgeInfo, src->edgeInfo, sizeof(lvo_f4) * (size_t)src->numEdges);
    }
}

/* StoreDensityField, compute_density_field.cpp:303-308 (takes ownership) */
static void store_field(lvo_world *w, lvo_field *field)
{
    lvo_field *slot = find_field(w, field->min, field->size);
    if (slot) {
        if (slot->materials != field->materials) { field_release(slot); }
        *slot = *field;
        return;
    }
    if (w->numFields == w->capFields) {
        w->capFields = w->capFields ? w->capFields * 2 : 16;
        w->fields = (lvo_field *)realloc(w->fields, sizeof(lvo_field) * (size_t)w->capFields);
    }
    w->fields[w->numFields++] = *field;
}

/* AABB::overlaps, aabb.h:24-33 */
static int aabb_overlaps(const int amin[3], const int amax[3], const int bmin[3], const int bmax[3])
{
    return !(amax[0] < bmin[0] || amax[1] < bmin[1] || amax[2] < bmin[2] ||
             amin[0] > bmax[0] || amin[1] > bmax[1] || amin[2] > bmax[2]);
}

/* GenerateDefaultDensityField + FindDefaultEdges, compute_density_field.cpp:138-231 */
static void generate_default(const lvo_world *w, const int min[3], int size, lvo_field *field)
{
    const size_t nF3 = (size_t)w->F * w->F * w->F;
    const size_t H3 = (size_t)w->H * w->H * w->H;
    int32_t *keys = (int32_t *)malloc(sizeof(int32_t) * 3 * H3);
    memset(field, 0, sizeof(*field));
    field->min[0] = min[0]; field->min[1] = min[1]; field->min[2] = min[2];
    field->size = size;
    field->materials = (int32_t *)malloc(sizeof(int32_t) * nF3);
    lvo_generate_field(w, min, size, field->materials);
    field->numEdges = lvo_find_edges(w, field->materials, keys);
    field->edgeKeys = (int32_t *)malloc(sizeof(int32_t) * (size_t)(field->numEdges + 1));
    field->edgeInfo = (lvo_f4 *)malloc(sizeof(lvo_f4) * (size_t)(field->numEdges + 1));
    memcpy(field->edgeKeys, keys, sizeof(int32_t) * (size_t)field->numEdges);
    free(keys);
    if (field->numEdges > 0)
        lvo_edge_info(w, min, size, field->edgeKeys, field->numEdges, field->edgeInfo);
}

/* LoadDensityField, compute_density_field.cpp:235-274.  *field is an owned
 * working copy; *stored says whether the cache now holds it. */
static int load_density_field(lvo_world *w, const int min[3], int size, lvo_field *field, int *stored)
{
    lvo_field *cached = find_field(w, min, size);
    const int fmax[3] = { min[0] + size, min[1] + size, min[2] + size };
    lvo_csg_op *replay;
    int i, numReplay = 0;
    *stored = 0;
    if (cached) field_copy(w, field, cached);
    else generate_default(w, min, size, field);

    replay = (lvo_csg_op *)malloc(sizeof(lvo_csg_op) * (size_t)(w->numOps + 1));
    for (i = field->lastCSGOperation; i < w->numOps; i++)
        if (aabb_overlaps(min, fmax, &w->opAABB[i][0], &w->opAABB[i][3])) replay[numReplay++] = w->ops[i];
    field->lastCSGOperation = w->numOps;
    if (numReplay > 0) {
        lvo_field copy;
        apply_csg(w, replay, numReplay, field->min, field->size, field);
        field_copy(w, &copy, field);
        store_field(w, &copy);
        *stored = 1;
    }
    free(replay);
    return 0;
}

int lvo_store_csg_operation(lvo_world *w, const lvo_csg_op *op, const int aabbMin[3], const int aabbMax[3])
{
    if (w->numOps == w->capOps) {
        w->capOps = w->capOps ? w->capOps * 2 : 16;
        w->ops = (lvo_csg_op *)realloc(w->ops, sizeof(lvo_csg_op) * (size_t)w->capOps);
        w->opAABB = (int (*)[6])realloc(w->opAABB, sizeof(int[6]) * (size_t)w->capOps);
    }
    w->ops[w->numOps] = *op;
    w->opAABB[w->numOps][0] = aabbMin[0]; w->opAABB[w->numOps][1] = aabbMin[1]; w->opAABB[w->numOps][2] = aabbMin[2];
    w->opAABB[w->numOps][3] = aabbMax[0]; w->opAABB[w->numOps][4] = aabbMax[1]; w->opAABB[w->numOps][5] = aabbMax[2];
    w->numOps++;
    return 0;
}

int lvo_clear_csg_operations(lvo_world *w) { w->numOps = 0; return 0; }

/* Compute_ApplyCSGOperations, compute_csg.cpp:224-242 */
int lvo_apply_csg_operations(lvo_world *w, const lvo_csg_op *ops, int numOps, const int min[3], int size)
{
    lvo_field field;
    int stored;
    load_density_field(w, min, size, &field, &stored);
    apply_csg(w, ops, numOps, min, size, &field);
    field.lastCSGOperation += numOps;
    store_field(w, &field);
    return 0;
}

static lvo_octree *find_octree(lvo_world *w, const int min[3], int size)
{
    int i;
    for (i = 0; i < w->numOctrees; i++)
        if (key_eq(w->octrees[i].min, w->octrees[i].size, min, size)) return &w->octrees[i];
    return NULL;
}

/* Compute_FreeChunkOctree, compute_octree.cpp:379-387 */
int lvo_free_chunk_octree(lvo_world *w, const int min[3], int size)
{
    lvo_octree *o = find_octree(w, min, size);
    if (o) {
        octree_release(o);
        *o = w->octrees[--w->numOctrees];
    }
    return 0;
}

/* Compute_ChunkIsEmpty, compute_density_field.cpp:278-299.  The reference
 * returns the inverted flag (isEmpty = numEdges > 0) from an unreachable
 * call site; restated with the evident meaning. */
int lvo_is_chunk_empty(lvo_world *w, const int min[3], int size, int *isEmpty)
{
    lvo_field *cached = find_field(w, min, size);
    lvo_field field;
    if (cached) { *isEmpty = cached->numEdges == 0; return 0; }
    generate_default(w, min, size, &field);
    *isEmpty = field.numEdges == 0;
    store_field(w, &field);
    return 0;
}

/* ConstructOctreeFromField, compute_octree.cpp:25-150 */
static int construct_octree(const lvo_world *w, const int min[3], const lvo_field *field, lvo_octree *octree)
{
    const size_t V3 = (size_t)w->V * w->V * w->V;
    uint32_t *codes = (uint32_t *)malloc(sizeof(uint32_t) * V3);
    int32_t *edgeMasks = (int32_t *)malloc(sizeof(int32_t) * V3);
    int32_t *matWords = (int32_t *)malloc(sizeof(int32_t) * V3);
    int n;
    memset(octree, 0, sizeof(*octree));
    octree->min[0] = min[0]; octree->min[1] = min[1]; octree->min[2] = min[2];
    octree->size = field->size;
    n = lvo_find_active_voxels(w, field->materials, codes, edgeMasks, matWords);
    octree->numNodes = n;
    if (n <= 0) { free(codes); free(edgeMasks); free(matWords); return 0; }
    octree->codes = (uint32_t *)realloc(codes, sizeof(uint32_t) * (size_t)n);
    octree->edgeMasks = (int32_t *)realloc(edgeMasks, sizeof(int32_t) * (size_t)n);
    octree->matWords = (int32_t *)realloc(matWords, sizeof(int32_t) * (size_t)n);
    octree->qefs = (lvo_qef *)malloc(sizeof(lvo_qef) * (size_t)n);
    octree->positions = (lvo_f4 *)malloc(sizeof(lvo_f4) * (size_t)n);
    octree->normals = (lvo_f4 *)malloc(sizeof(lvo_f4) * (size_t)n);
    lvo_create_leaf_nodes(w, sample_scale(w, field->size), octree->codes, octree->edgeMasks, n,
                          field->edgeKeys, field->edgeInfo, field->numEdges, octree->qefs, octree->normals);
    lvo_solve_qefs(min, octree->qefs, n, octree->positions);
    return 0;
}

static void snapshot_field(const lvo_world *w, lvo_octree *o, const lvo_field *f)
{
    const size_t nF3 = (size_t)w->F * w->F * w->F;
    o->numEdges = f->numEdges;
    o->materials = (int32_t *)malloc(sizeof(int32_t) * nF3);
    memcpy(o->materials, f->materials, sizeof(int32_t) * nF3);
    o->edgeKeys = (int32_t *)malloc(sizeof(int32_t) * (size_t)(f->numEdges + 1));
    o->edgeInfo = (lvo_f4 *)malloc(sizeof(lvo_f4) * (size_t)(f->numEdges + 1));
    memcpy(o->edgeKeys, f->edgeKeys, sizeof(int32_t) * (size_t)f->numEdges);
    memcpy(o->edgeInfo, f->edgeInfo, sizeof(lvo_f4) * (size_t)f->numEdges);
}

#define DUP(dst, src, n, T) do { (dst) = (T *)malloc(sizeof(T) * (size_t)((n) + 1)); \
    if ((n) > 0) memcpy((dst), (src), sizeof(T) * (size_t)(n)); } while (0)

/* Compute_GenerateChunkMesh, compute_octree.cpp:351-375 (+ LoadOctree :154-181) */
int lvo_generate_chunk_mesh(lvo_world *w, const int min[3], int size, lvo_chunk *out)
{
    lvo_octree *octree = find_octree(w, min, size);
    lvo_octree fresh;
    memset(out, 0, sizeof(*out));
    if (!octree) {
        lvo_field field;
        int stored;
        load_density_field(w, min, size, &field, &stored);
        memset(&fresh, 0, sizeof(fresh));
        if (field.numEdges == 0) {
            /* octree->numNodes = 0 and nothing is cached (compute_octree.cpp:167-171) */
            const size_t nF3 = (size_t)w->F * w->F * w->F;
            DUP(out->materials, field.materials, nF3, int32_t);
            field_release(&field);
            return 0;
        }
        construct_octree(w, min, &field, &fresh);
        snapshot_field(w, &fresh, &field);
        field_release(&field);
        if (w->numOctrees == w->capOctrees) {
            w->capOctrees = w->capOctrees ? w->capOctrees * 2 : 16;
            w->octrees = (lvo_octree *)realloc(w->octrees, sizeof(lvo_octree) * (size_t)w->capOctrees);
        }
        w->octrees[w->numOctrees++] = fresh;
        octree = &w->octrees[w->numOctrees - 1];
    }
    {
        const size_t nF3 = (size_t)w->F * w->F * w->F;
        const int n = octree->numNodes;
        out->numEdges = octree->numEdges;
        out->numNodes = n;
        DUP(out->materials, octree->materials, nF3, int32_t);
        DUP(out->edgeKeys, octree->edgeKeys, octree->numEdges, int32_t);
        DUP(out->edgeInfo, octree->edgeInfo, octree->numEdges, lvo_f4);
        if (n > 0) {
            DUP(out->codes, octree->codes, n, uint32_t);
            DUP(out->edgeMasks, octree->edgeMasks, n, int32_t);
            DUP(out->matWords, octree->matWords, n, int32_t);
            DUP(out->qefs, octree->qefs, n, lvo_qef);
            DUP(out->positions, octree->positions, n, lvo_f4);
            DUP(out->normals, octree->normals, n, lvo_f4);
            out->indices = (int32_t *)malloc(sizeof(int32_t) * 18 * (size_t)n);
            out->numTriangles = lvo_generate_mesh(w, octree->codes, octree->matWords, n, out->indices);
            out->vertices = (lvo_vertex *)malloc(sizeof(lvo_vertex) * (size_t)n);
            lvo_vertex_buffer(octree->positions, octree->normals, octree->matWords, n, size, out->vertices);
            out->seams = (lvo_seam_node *)malloc(sizeof(lvo_seam_node) * (size_t)n);
            out->numSeamNodes = lvo_seam_nodes(w, octree->codes, octree->matWords, octree->positions,
                                               octree->normals, n, out->seams);
        }
    }
    return 0;
}

void lvo_chunk_free(lvo_chunk *c)
{
    free(c->materials); free(c->edgeKeys); free(c->edgeInfo); free(c->codes); free(c->edgeMasks);
    free(c->matWords); free(c->qefs); free(c->positions); free(c->normals); free(c->vertices);
    free(c->indices); free(c->seams);
    memset(c, 0, sizeof(*c));
}

/* CPU-baseline helper: the per-chunk work of Compute_GenerateChunkMesh with
 * cold caches, OpenMP over independent chunks. */
/* torchrun exports OMP_NUM_THREADS=1 to its workers; the baseline leg asks for the host's cores */
int lvo_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

int lvo_generate_batch_counts(const lvo_world *w, int n, const int *minSize, int32_t *counts)
{
    int i, threads = 1;
#ifdef _OPENMP
    threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (i = 0; i < n; i++) {
        const int *ms = &minSize[4 * i];
        lvo_field field;
        lvo_octree octree;
        int T = 0, S = 0;
        generate_default(w, ms, ms[3], &field);
        memset(&octree, 0, sizeof(octree));
        if (field.numEdges > 0) {
            construct_octree(w, ms, &field, &octree);
            if (octree.numNodes > 0) {
                int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * 18 * (size_t)octree.numNodes);
                lvo_vertex *vb = (lvo_vertex *)malloc(sizeof(lvo_vertex) * (size_t)octree.numNodes);
                lvo_seam_node *sn = (lvo_seam_node *)malloc(sizeof(lvo_seam_node) * (size_t)octree.numNodes);
                T = lvo_generate_mesh(w, octree.codes, octree.matWords, octree.numNodes, idx);
                lvo_vertex_buffer(octree.positions, octree.normals, octree.matWords, octree.numNodes, ms[3], vb);
                S = lvo_seam_nodes(w, octree.codes, octree.matWords, octree.positions, octree.normals,
                                   octree.numNodes, sn);
                free(idx); free(vb); free(sn);
            }
        }
        counts[4 * i + 0] = field.numEdges;
        counts[4 * i + 1] = octree.numNodes;
        counts[4 * i + 2] = T;
        counts[4 * i + 3] = S;
        field_release(&field);
        octree_release(&octree);
    }
    return threads;
}

/* FNV-1a over a byte range: the per-chunk digests of lvo_generate_batch_digests (test
 * infrastructure: lets the full-size configurations be compared chunk by chunk without
 * holding every reference mesh in memory) */
static uint64_t fnv1a64(const void *data, size_t bytes)
{
    const uint8_t *p = (const uint8_t *)data;
    uint64_t h = 1469598103934665603ull;
    size_t i;
    for (i = 0; i < bytes; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

uint64_t lvo_fnv1a64(const void *data, size_t bytes) { return fnv1a64(data, bytes); }

/* Like lvo_generate_batch_counts, and additionally digests[3*i..] = FNV-1a of chunk i's vertex
 * buffer (numNodes * 48 B), triangle indices in order (T * 12 B) and seam nodes (S * 48 B) --
 * the three arrays generateChunkMesh hands out (compute_octree.cpp:326-347, :275-322). */
int lvo_generate_batch_digests(const lvo_world *w, int n, const int *minSize, int32_t *counts, uint64_t *digests)
{
    int i, threads = 1;
#ifdef _OPENMP
    threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (i = 0; i < n; i++) {
        const int *ms = &minSize[4 * i];
        lvo_field field;
        lvo_octree octree;
        int T = 0, S = 0;
        generate_default(w, ms, ms[3], &field);
        memset(&octree, 0, sizeof(octree));
        digests[3 * i + 0] = digests[3 * i + 1] = digests[3 * i + 2] = fnv1a64(NULL, 0);
        if (field.numEdges > 0) {
            construct_octree(w, ms, &field, &octree);
            if (octree.numNodes > 0) {
                int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * 18 * (size_t)octree.numNodes);
                lvo_vertex *vb = (lvo_vertex *)malloc(sizeof(lvo_vertex) * (size_t)octree.numNodes);
                lvo_seam_node *sn = (lvo_seam_node *)malloc(sizeof(lvo_seam_node) * (size_t)octree.numNodes);
                T = lvo_generate_mesh(w, octree.codes, octree.matWords, octree.numNodes, idx);
                lvo_vertex_buffer(octree.positions, octree.normals, octree.matWords, octree.numNodes, ms[3], vb);
                S = lvo_seam_nodes(w, octree.codes, octree.matWords, octree.positions, octree.normals,
                                   octree.numNodes, sn);
                digests[3 * i + 0] = fnv1a64(vb, sizeof(lvo_vertex) * (size_t)octree.numNodes);
                digests[3 * i + 1] = fnv1a64(idx, sizeof(int32_t) * 3 * (size_t)T);
                digests[3 * i + 2] = fnv1a64(sn, sizeof(lvo_seam_node) * (size_t)S);
                free(idx); free(vb); free(sn);
            }
        }
        counts[4 * i + 0] = field.numEdges;
        counts[4 * i + 1] = octree.numNodes;
        counts[4 * i + 2] = T;
        counts[4 * i + 3] = S;
        field_release(&field);
        octree_release(&octree);
    }
    return threads;
}
