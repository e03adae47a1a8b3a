// scene_flatten.cpp -- host side of the real-scene front end: indexed, instanced meshes -> the flat
// per-corner streams the kernels read.
//
// The reference keeps a scene object as up to three levels of OptiX traversables (pipeline_helper.cpp:268-272) whose leaf
// GASes carry an SBT record with INDEXED shading attributes (HitGroupData, pipeline_helper.cpp:182-193); its closest-hit
// program transforms the hit triangle's corners and normals to world space per hit (kernel.cu:205-240:
// optixTransformPointFromObjectToWorldSpace / optixTransformNormalFromObjectToWorldSpace) and gathers normals and texture
// coordinates through two index arrays. Here that work is done ONCE at upload: every (instance, triangle) becomes a
// world-space triangle with three world-space normals and three texture coordinates, stored contiguously per primitive.
// B200 has 180 GB of HBM; the per-hit cost drops to 36 + 24 contiguous bytes and the BVH is a single level.
//
// Arithmetic (the oracle restates it independently, bit for bit):
//   corner   p_w = M p          per row  fma(m2, z, fma(m1, y, fma(m0, x, m3)))                      (binary32)
//   normal   n_w = A^-T n       A = upper 3x3 of M; A^-T = cofactors / det evaluated in binary64 in the order written
//                               below, each entry rounded to binary32; per row fma(g2, nz, fma(g1, ny, g0 * nx))
// Built with -ffp-contract=off.
#include <cmath>
#include <cstring>

#include "scene_flatten.h"

namespace dprt {

namespace {

bool mesh_ok(const dprt_mesh_desc& m) {
    if (m.ntris < 0 || m.nPositions < 0 || m.nNormals < 0 || m.nTexCoords < 0) return false;
    if (m.ntris == 0) return true;
    if (!m.positions || !m.indices || !m.normals || !m.normalIndices) return false;
    if ((m.texCoords == nullptr) != (m.texCoordIndices == nullptr)) return false;
    if (m.materialID < 0 || m.materialID >= DPRT_MAX_MATERIALS) return false;
    return true;
}

}  // namespace

int64_t flatten_count(const dprt_mesh_desc* meshes, int nMeshes, const dprt_instance_desc* instances, int64_t nInstances) {
    if (!meshes || !instances || nMeshes <= 0 || nInstances <= 0) return -1;
    for (int m = 0; m < nMeshes; m++) if (!mesh_ok(meshes[m])) return -1;
    int64_t total = 0;
    for (int64_t i = 0; i < nInstances; i++) {
        const int mi = instances[i].mesh;
        if (mi < 0 || mi >= nMeshes) return -1;
        total += meshes[mi].ntris;
    }
    return total;
}

int flatten_instances(const dprt_mesh_desc* meshes, int nMeshes, const dprt_instance_desc* instances, int64_t nInstances,
                      float* verts9, float* normals9, float* uv6, int32_t* matIds, int* hasUv) {
    const int64_t total = flatten_count(meshes, nMeshes, instances, nInstances);
    if (total <= 0 || !verts9 || !normals9) return -1;
    bool anyUv = false;
    for (int m = 0; m < nMeshes; m++) anyUv = anyUv || (meshes[m].ntris > 0 && meshes[m].texCoords != nullptr);
    if (hasUv) *hasUv = anyUv ? 1 : 0;
    // index ranges, once per mesh
    for (int m = 0; m < nMeshes; m++) {
        const dprt_mesh_desc& me = meshes[m];
        for (int64_t k = 0; k < 3 * me.ntris; k++) {
            if (me.indices[k] < 0 || me.indices[k] >= me.nPositions) return -1;
            if (me.normalIndices[k] < 0 || me.normalIndices[k] >= me.nNormals) return -1;
            if (me.texCoordIndices && (me.texCoordIndices[k] < 0 || me.texCoordIndices[k] >= me.nTexCoords)) return -1;
        }
    }
    int64_t base = 0;
    for (int64_t i = 0; i < nInstances; i++) {
        const dprt_mesh_desc& me = meshes[instances[i].mesh];
        const float* M = instances[i].objectToWorld;
        const double a00 = M[0], a01 = M[1], a02 = M[2], a10 = M[4], a11 = M[5], a12 = M[6], a20 = M[8], a21 = M[9], a22 = M[10];
        const double c00 = a11 * a22 - a12 * a21, c01 = -(a10 * a22 - a12 * a20), c02 = a10 * a21 - a11 * a20;
        const double c10 = -(a01 * a22 - a02 * a21), c11 = a00 * a22 - a02 * a20, c12 = -(a00 * a21 - a01 * a20);
        const double c20 = a01 * a12 - a02 * a11, c21 = -(a00 * a12 - a02 * a10), c22 = a00 * a11 - a01 * a10;
        const double det = a00 * c00 + a01 * c01 + a02 * c02;
        if (!(det != 0.0) || !std::isfinite(det)) return -1;
        const double inv = 1.0 / det;
        const float g[9] = {(float)(c00 * inv), (float)(c01 * inv), (float)(c02 * inv), (float)(c10 * inv), (float)(c11 * inv),
                            (float)(c12 * inv), (float)(c20 * inv), (float)(c21 * inv), (float)(c22 * inv)};
#pragma omp parallel for schedule(static) if (me.ntris > 4096)
        for (int64_t t = 0; t < me.ntris; t++) {
            const int64_t p = base + t;
            for (int c = 0; c < 3; c++) {
                const float* v = me.positions + 3 * (int64_t)me.indices[3 * t + c];
                float* o = verts9 + 9 * p + 3 * c;
                o[0] = fmaf(M[2], v[2], fmaf(M[1], v[1], fmaf(M[0], v[0], M[3])));
                o[1] = fmaf(M[6], v[2], fmaf(M[5], v[1], fmaf(M[4], v[0], M[7])));
                o[2] = fmaf(M[10], v[2], fmaf(M[9], v[1], fmaf(M[8], v[0], M[11])));
                const float* n = me.normals + 3 * (int64_t)me.normalIndices[3 * t + c];
                float* q = normals9 + 9 * p + 3 * c;
                q[0] = fmaf(g[2], n[2], fmaf(g[1], n[1], g[0] * n[0]));
                q[1] = fmaf(g[5], n[2], fmaf(g[4], n[1], g[3] * n[0]));
                q[2] = fmaf(g[8], n[2], fmaf(g[7], n[1], g[6] * n[0]));
                if (uv6) {
                    float* u = uv6 + 6 * p + 2 * c;
                    if (me.texCoords) {
                        const float* tc = me.texCoords + 2 * (int64_t)me.texCoordIndices[3 * t + c];
                        u[0] = tc[0]; u[1] = tc[1];
                    } else { u[0] = 0.0f; u[1] = 0.0f; }
                }
            }
            if (matIds) matIds[p] = me.materialID;
        }
        base += me.ntris;
    }
    return 0;
}

}  // namespace dprt
