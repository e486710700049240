// ref_simplify.cpp -- the reference's mesh simplifier (leven/src/ng_mesh_simplify.cpp) with its SSE
// 4-D QEF solver (leven/src/qef_simd.h), compiled for the host from where they lie, behind one
// plain C entry point.  It runs on every chunk mesh right after export (clipmap.cpp:449-465).
//
// TEST INFRASTRUCTURE (oracle/_ref): the checker of the GPU simplifier.
// Both files are compiled unmodified except for one syntax rewrite by ref_shim/translate.py:
// `x.m128_f32[i]` (MSVC's __m128 is a union) -> `x[i]` (gcc's is a vector type).  What this file
// has to DEFINE, because the reference leaves it to the platform:
//   _mm_rsqrt_ps     x86's ~12-bit reciprocal square root estimate, whose bits differ between CPU
//                    vendors; := 1 / sqrt(x) with correctly rounded _mm_sqrt_ps / _mm_div_ps, the same
//                    choice the arithmetic spec makes for OpenCL's rsqrt
//   std::uniform_int_distribution / std::mt19937(42)   the candidate-edge sampling
//                    (ng_mesh_simplify.cpp:195-205); mt19937 is standardised, the distribution is
//                    not: this build uses libstdc++'s (Lemire's multiply-shift with rejection)
//   glm::dot / glm::length2 on vec4  GLM 0.9.3's left-to-right sum (ref_shim/miniglm)
//   __declspec(align(16))            -> __attribute__((aligned(16)))
// Compiled with -ffp-contract=off: every SSE arithmetic intrinsic is one IEEE binary32 operation per lane.
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <stdint.h>

#include <immintrin.h>
#include <xmmintrin.h>

#include <glm/glm.hpp>

static inline __m128 lvn_exact_rsqrt_ps(__m128 x) { return _mm_div_ps(_mm_set1_ps(1.f), _mm_sqrt_ps(x)); }
#define _mm_rsqrt_ps lvn_exact_rsqrt_ps
#define __declspec(x) __attribute__((x))
#define align(n) aligned(n)

#define QEF_INCLUDE_IMPL
#include "qef_simd.h.inc"
#include "ng_mesh_simplify.cpp.inc"

extern "C" {

// ngMeshSimplifier on one mesh, in place.  vertices: MeshVertex[*numVertices] (12 floats each),
// triangles: MeshTriangle[*numTriangles].  Returns 0, or -1 when the mesh exceeds MeshBuffer's
// fixed capacity (MAX_MESH_VERTICES / MAX_MESH_TRIANGLES with LEVEN defined).
int ref_mesh_simplify(float *vertices12, int *numVertices, int *triangles3, int *numTriangles, const float *worldSpaceOffset4,
                      float edgeFraction, int maxIterations, float targetPercentage, float maxError, float maxEdgeSize,
                      float minAngleCosine)
{
    if (*numVertices > MAX_MESH_VERTICES || *numTriangles > MAX_MESH_TRIANGLES) return -1;
    MeshBuffer *mesh = new MeshBuffer;
    mesh->numVertices = *numVertices;
    mesh->numTriangles = *numTriangles;
    std::memcpy(mesh->vertices, vertices12, sizeof(MeshVertex) * (size_t)*numVertices);
    std::memcpy(mesh->triangles, triangles3, sizeof(MeshTriangle) * (size_t)*numTriangles);
    MeshSimplificationOptions options;
    options.edgeFraction = edgeFraction;
    options.maxIterations = maxIterations;
    options.targetPercentage = targetPercentage;
    options.maxError = maxError;
    options.maxEdgeSize = maxEdgeSize;
    options.minAngleCosine = minAngleCosine;
    ngMeshSimplifier(mesh, glm::vec4(worldSpaceOffset4[0], worldSpaceOffset4[1], worldSpaceOffset4[2], worldSpaceOffset4[3]), options);
    *numVertices = mesh->numVertices;
    *numTriangles = mesh->numTriangles;
    std::memcpy(vertices12, mesh->vertices, sizeof(MeshVertex) * (size_t)mesh->numVertices);
    std::memcpy(triangles3, mesh->triangles, sizeof(MeshTriangle) * (size_t)mesh->numTriangles);
    delete mesh;
    return 0;
}

// the raw candidate-edge sample of one FindValidCollapses call, for tests of the device-side sampler
void ref_random_edges(int numEdges, int count, int *out)
{
    std::mt19937 prng;
    prng.seed(42);
    std::uniform_int_distribution<int> distribution(0, numEdges - 1);
    for (int i = 0; i < count; i++) out[i] = distribution(prng);
}

}  // extern "C"
