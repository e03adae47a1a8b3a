// oracle.cpp -- CPU ORACLE of the PG2024 data-parallel ray tracer's per-bounce inner loop.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it; libdprt.so never does.
//
// Scalar C++17 (+OpenMP over independent paths) restatement of the reference's device and host logic:
//   optix/random.hpp:31-67                      tea<4>, lcg, rnd               -> tea4/lcg/rnd
//   optix/sample.hpp:7-17                       uniformHemisphere              -> uniform_hemisphere
//   optix/path_gen_kernel.cu:46-105             PathGen raygen                 -> Rank::path_gen
//   optix/distributed_traversal_kernel.cu:215-340  TraRay raygen               -> Rank::traverse
//   optix/kernel.cu:50-162,171-300,362-466      MainRay raygen + closest hit   -> Rank::shade
//   optix/bsdfs/lambertian.hpp:10-32, water.hpp:12-94                          -> sample_lambertian/water
//   optix/shadow_ray_kernel.cu:150-355          ShadowRay raygen               -> Rank::shadow_trace
//   optix/secondary_ray_kernel.cu:172-369       SecondaryRay raygen            -> Rank::secondary_trace
//   src/cuda/cuda_compaction.cu:8-35,140-198,352-617  stable bucket partition  -> Rank::partition / bucket_queries
//   src/cuda/frame_buffer_update.cu:31-127,172-192,222-324                     -> frame/depth/target updates
//   src/render/renderer.cpp:1212-1318,1320-1452,1457-1574,2031-2052            -> World::render_sample, image
//   trainingcode/module.py:36-45,755-837        4Res256/6Res256 proxy MLP      -> mlp_forward_row (fp32)
//   optix/vis_ray_kernel.cu:98-161              Vis pipeline (training samples) -> orc_gen_train_data
//   optix/precom_ray_kernel.cu:193-299          Precom pipeline (training samples) -> orc_gen_precom_data
//   optix/kernel.cu:190-292,311-359 (+ the __anyhit__ah of every pipeline), pipeline_helper.cpp:182-193,
//   renderer.cpp:1621-1721, kernel.cu:28-48      indexed / instanced meshes, albedo + opacity maps, alpha cut-out,
//                                                environment map                 -> flatten_object, Texture, alpha_ignored, env_radiance
//
// PARITY PINNING. The reference ships no tests or golden vectors and cannot be built (README.md:5). What is
// pinned against reference code executed in the build container: tea<4>/lcg/rnd against optix/random.hpp
// compiled with g++ (tests/golden/rng_kat.json), and the MLP against trainingcode/module.py on PyTorch-CPU
// (tests/golden/mlp_*.npz). Ray/triangle intersection (OptiX runtime), Camera/Frame/Triangle::sample/
// Coordinates (missing moana headers) have no reference implementation in the tree: for those stages
// PARITY IS UNPINNED and this oracle is the specification (DESIGN.md "Arithmetic specification").
//
// Arithmetic: every float operation is an explicit IEEE-754 binary32 add/mul/fma/div/sqrt in a fixed order;
// build with -ffp-contract=off so the compiler neither fuses nor splits them. Closest-hit results are
// independent of the acceleration structure (watertight test per triangle, ties in t -> lower primitive id),
// so this file uses its own median-split binary BVH and also offers brute force.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <immintrin.h>
#include <omp.h>
#include <vector>

#include "dprt_types.h"

namespace {

struct V3 { float x, y, z; };
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 scale(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline float comp(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }
inline float dot(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
inline V3 cross(V3 a, V3 b) {
    return v3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
inline float length(V3 a) { return sqrtf(dot(a, a)); }
inline V3 normalized(V3 a) { float inv = 1.0f / length(a); return scale(a, inv); }
inline V3 at(V3 o, V3 d, float t) { return v3(fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z)); }

// ---- optix/random.hpp:31-67 ----
inline uint32_t tea4(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (int n = 0; n < 4; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
inline uint32_t lcg(uint32_t& prev) { prev = 1664525u * prev + 1013904223u; return prev & 0x00FFFFFFu; }
inline float rnd(uint32_t& prev) { return (float)lcg(prev) / (float)0x01000000; }

// ---- deterministic transcendentals (specification; Cephes single-precision polynomials) ----
void det_sincos2pi(float x, float* s, float* c) {
    float r = x * 4.0f;
    float qf = floorf(r + 0.5f);
    float f = r - qf;
    int q = (int)qf & 3;
    float a = f * 1.57079632679489661923f;
    float z = a * a;
    float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f) * z, a, a);
    float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f), z * z,
                    fmaf(-0.5f, z, 1.0f));
    float ss = (q & 1) ? cp : sp;
    float cc = (q & 1) ? sp : cp;
    if (q == 1 || q == 2) cc = -cc;
    if (q == 2 || q == 3) ss = -ss;
    *s = ss; *c = cc;
}
float det_asin_poly(float x) {
    float z = x * x;
    float p = fmaf(fmaf(fmaf(fmaf(4.2163199048e-2f, z, 2.4181311049e-2f), z, 4.5470025998e-2f), z, 7.4953002686e-2f), z,
                   1.6666752422e-1f);
    return fmaf(p * z, x, x);
}
float det_acos(float x) {
    const float PI = 3.14159265358979323846f, PIO2 = 1.57079632679489661923f;
    if (x > 0.5f) { float s = sqrtf(0.5f * (1.0f - x)); return 2.0f * det_asin_poly(s); }
    if (x < -0.5f) { float s = sqrtf(0.5f * (1.0f + x)); return PI - 2.0f * det_asin_poly(s); }
    return PIO2 - det_asin_poly(x);
}
float det_atan_pos(float t) {
    const float PIO2 = 1.57079632679489661923f, PIO4 = 0.78539816339744830962f;
    float y0, x;
    if (t > 2.414213562373095f) { y0 = PIO2; x = -(1.0f / t); }
    else if (t > 0.4142135623730950f) { y0 = PIO4; x = (t - 1.0f) / (t + 1.0f); }
    else { y0 = 0.0f; x = t; }
    float z = x * x;
    float p = fmaf(fmaf(fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f);
    return y0 + fmaf(p * z, x, x);
}
float det_atan2(float y, float x) {
    const float PI = 3.14159265358979323846f, PIO2 = 1.57079632679489661923f;
    if (x == 0.0f) { if (y > 0.0f) return PIO2; if (y < 0.0f) return -PIO2; return 0.0f; }
    float a = det_atan_pos(fabsf(y) / fabsf(x));
    if (x < 0.0f) a = PI - a;
    return (y < 0.0f) ? -a : a;
}
// Coordinates::cartesianToSpherical / ...ForTrain (convention of src/cuda/bvh_intersection.cu:18-26)
void cartesian_to_spherical(V3 d, float* phi, float* theta) {
    float p = det_atan2(d.y, d.x);
    if (p < 0.0f) p += 6.28318530717958647692f;
    *phi = p;
    *theta = det_acos(fminf(1.0f, fmaxf(-1.0f, d.z)));
}
// optix/sample.hpp:7-17
V3 uniform_hemisphere(float xi1, float xi2) {
    float z = xi1;
    float r = sqrtf(fmaxf(0.0f, fmaf(-z, z, 1.0f)));
    float s, c; det_sincos2pi(xi2, &s, &c);
    return v3(r * c, r * s, z);
}
// moana Frame (missing header): branchless orthonormal basis of Duff et al. 2017
struct Frame { V3 s, t, n; };
Frame make_frame(V3 n) {
    float sign = copysignf(1.0f, n.z);
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    Frame f;
    f.s = v3(fmaf(sign * n.x, n.x * a, 1.0f), sign * b, -(sign * n.x));
    f.t = v3(b, fmaf(n.y, n.y * a, sign), -n.y);
    f.n = n;
    return f;
}
V3 to_world(const Frame& f, V3 w) {
    return v3(fmaf(f.n.x, w.z, fmaf(f.t.x, w.y, f.s.x * w.x)), fmaf(f.n.y, w.z, fmaf(f.t.y, w.y, f.s.y * w.x)),
              fmaf(f.n.z, w.z, fmaf(f.t.z, w.y, f.s.z * w.x)));
}
V3 to_local(const Frame& f, V3 w) { return v3(dot(f.s, w), dot(f.t, w), dot(f.n, w)); }
V3 xform_point(const float* m, V3 p) {
    return v3(fmaf(m[2], p.z, fmaf(m[1], p.y, fmaf(m[0], p.x, m[3]))), fmaf(m[6], p.z, fmaf(m[5], p.y, fmaf(m[4], p.x, m[7]))),
              fmaf(m[10], p.z, fmaf(m[9], p.y, fmaf(m[8], p.x, m[11]))));
}
V3 xform_vector(const float* m, V3 v) {
    return v3(fmaf(m[2], v.z, fmaf(m[1], v.y, m[0] * v.x)), fmaf(m[6], v.z, fmaf(m[5], v.y, m[4] * v.x)),
              fmaf(m[10], v.z, fmaf(m[9], v.y, m[8] * v.x)));
}
inline uint16_t f2h(float f) { return (uint16_t)_cvtss_sh(f, _MM_FROUND_TO_NEAREST_INT); }
inline float h2f(uint16_t h) { return _cvtsh_ss(h); }

// ---- watertight ray/triangle test (Woop, Benthin, Wald 2013): stands in for the OptiX intersector ----
struct Shear { int kx, ky, kz; float Sx, Sy, Sz; };
Shear make_shear(V3 d) {
    Shear r;
    int kz = 0; float m = fabsf(d.x);
    if (fabsf(d.y) > m) { kz = 1; m = fabsf(d.y); }
    if (fabsf(d.z) > m) { kz = 2; }
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    float dz = comp(d, kz);
    if (dz < 0.0f) std::swap(kx, ky);
    r.kx = kx; r.ky = ky; r.kz = kz;
    r.Sx = comp(d, kx) / dz; r.Sy = comp(d, ky) / dz; r.Sz = 1.0f / dz;
    return r;
}
bool tri_hit(const Shear& rs, V3 o, const float* tv, float tmin, float tmax, float* t_out, float* alpha, float* beta) {
    V3 A = sub(v3(tv[0], tv[1], tv[2]), o), B = sub(v3(tv[3], tv[4], tv[5]), o), C = sub(v3(tv[6], tv[7], tv[8]), o);
    float Akz = comp(A, rs.kz), Bkz = comp(B, rs.kz), Ckz = comp(C, rs.kz);
    float Ax = fmaf(-rs.Sx, Akz, comp(A, rs.kx)), Ay = fmaf(-rs.Sy, Akz, comp(A, rs.ky));
    float Bx = fmaf(-rs.Sx, Bkz, comp(B, rs.kx)), By = fmaf(-rs.Sy, Bkz, comp(B, rs.ky));
    float Cx = fmaf(-rs.Sx, Ckz, comp(C, rs.kx)), Cy = fmaf(-rs.Sy, Ckz, comp(C, rs.ky));
    float U = fmaf(Cx, By, -(Cy * Bx));
    float V = fmaf(Ax, Cy, -(Ay * Cx));
    float W = fmaf(Bx, Ay, -(By * Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    float det = (U + V) + W;
    if (det == 0.0f) return false;
    float Az = rs.Sz * Akz, Bz = rs.Sz * Bkz, Cz = rs.Sz * Ckz;
    float T = fmaf(U, Az, fmaf(V, Bz, W * Cz));
    float t = T / det;
    if (!(t > tmin && t < tmax)) return false;
    *t_out = t; *alpha = V / det; *beta = W / det;
    return true;
}
// proxy AABB in object space: optixTrace(AS.aabbHandle) + optixIsFrontFaceHit
bool aabb_hit(V3 ol, V3 dl, const float* mn, const float* mx, float tmin, float tmax, float* t_out, bool* inside) {
    float ix = 1.0f / dl.x, iy = 1.0f / dl.y, iz = 1.0f / dl.z;
    float x0 = (mn[0] - ol.x) * ix, x1 = (mx[0] - ol.x) * ix;
    float y0 = (mn[1] - ol.y) * iy, y1 = (mx[1] - ol.y) * iy;
    float z0 = (mn[2] - ol.z) * iz, z1 = (mx[2] - ol.z) * iz;
    float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
    if (!(tn <= tf)) return false;
    if (tn > tmin && tn < tmax) { *t_out = tn; *inside = false; return true; }
    if (tf > tmin && tf < tmax) { *t_out = tf; *inside = true; return true; }
    return false;
}

// ---- the oracle's own accelerator: median-split binary BVH, double-precision padded slab test ----
struct Hit { float t; int prim; float alpha, beta; };

struct World;
struct Mesh;
// the __anyhit__ah program (kernel.cu:311-359 and its copies in the other pipelines): true = optixIgnoreIntersection()
bool alpha_ignored(const World& w, const Mesh& m, int prim, float alpha, float beta);

struct Mesh {
    std::vector<float> verts;     // 9 per tri
    std::vector<float> normals;   // 9 per tri
    std::vector<float> uvs;       // 6 per tri (u0 v0 u1 v1 u2 v2), empty = the object has no texture coordinates
    std::vector<int32_t> mats;
    int ntris = 0;
    const World* world = nullptr; // textures + material table for the alpha cut-out test (set when the object is added)
    bool cut(int prim, float al, float be) const { return world && !uvs.empty() && alpha_ignored(*world, *this, prim, al, be); }
    struct Node { double lo[3], hi[3]; int left, right, first, count; };
    std::vector<Node> nodes;
    std::vector<int> order;
    double pad = 0;

    void build() {
        order.resize(ntris);
        for (int i = 0; i < ntris; i++) order[i] = i;
        double m = 1.0;
        for (size_t i = 0; i < verts.size(); i++) m = std::max(m, (double)fabsf(verts[i]));
        pad = std::ldexp(m, -16);
        nodes.clear(); nodes.reserve(ntris);
        nodes.push_back(Node());
        build_node(0, 0, ntris);
    }
    void centroid(int p, double* c) const {
        const float* v = &verts[9 * (size_t)p];
        for (int a = 0; a < 3; a++) c[a] = ((double)v[a] + v[3 + a] + v[6 + a]) / 3.0;
    }
    void build_node(int ni, int first, int count) {
        Node n;
        for (int a = 0; a < 3; a++) { n.lo[a] = 1e300; n.hi[a] = -1e300; }
        double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
        for (int i = first; i < first + count; i++) {
            const float* v = &verts[9 * (size_t)order[i]];
            for (int k = 0; k < 3; k++) for (int a = 0; a < 3; a++) {
                n.lo[a] = std::min(n.lo[a], (double)v[3 * k + a]); n.hi[a] = std::max(n.hi[a], (double)v[3 * k + a]);
            }
            double c[3]; centroid(order[i], c);
            for (int a = 0; a < 3; a++) { clo[a] = std::min(clo[a], c[a]); chi[a] = std::max(chi[a], c[a]); }
        }
        for (int a = 0; a < 3; a++) { n.lo[a] -= pad; n.hi[a] += pad; }
        n.left = n.right = -1; n.first = first; n.count = count;
        if (count > 4) {
            int ax = 0;
            if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
            if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
            int mid = first + count / 2;
            std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count, [&](int a, int b) {
                double ca[3], cb[3]; centroid(a, ca); centroid(b, cb);
                return ca[ax] < cb[ax] || (ca[ax] == cb[ax] && a < b);
            });
            n.count = 0;
            n.left = (int)nodes.size(); nodes.push_back(Node());
            n.right = (int)nodes.size(); nodes.push_back(Node());
            nodes[ni] = n;
            build_node(n.left, first, mid - first);
            build_node(nodes[ni].right, mid, first + count - mid);
            return;
        }
        nodes[ni] = n;
    }
    static bool slab(const Node& n, const double* o, const double* inv, double tmin, double tmax, double* tn_out) {
        double tn = tmin, tf = tmax;
        for (int a = 0; a < 3; a++) {
            double t0 = (n.lo[a] - o[a]) * inv[a], t1 = (n.hi[a] - o[a]) * inv[a];
            if (t0 > t1) std::swap(t0, t1);
            if (t0 != t0 || t1 != t1) continue;   // 0*inf: origin on the slab plane of a parallel ray -> no constraint
            tn = std::max(tn, t0); tf = std::min(tf, t1);
        }
        *tn_out = tn;
        return tn <= tf * (1.0 + 1e-9) + 1e-12;
    }
    // closest (any=false) or first-found (any=true) hit with tmin < t < tmax
    bool trace(V3 o, V3 d, float tmin, float tmax, bool any, Hit& hit) const {
        Shear rs = make_shear(d);
        double od[3] = {o.x, o.y, o.z}, inv[3] = {1.0 / (double)d.x, 1.0 / (double)d.y, 1.0 / (double)d.z};
        float tbest = tmax; int bestPrim = 0x7fffffff; bool found = false; float ba = 0, bb = 0;
        int stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const Node& n = nodes[stack[--sp]];
            double tn;
            if (!slab(n, od, inv, (double)tmin, (double)tbest, &tn)) continue;
            if (n.count == 0) { stack[sp++] = n.left; stack[sp++] = n.right; continue; }
            for (int i = n.first; i < n.first + n.count; i++) {
                int p = order[i]; float t, al, be;
                if (tri_hit(rs, o, &verts[9 * (size_t)p], tmin, tmax, &t, &al, &be)) {
                    if (cut(p, al, be)) continue;
                    if (any) { hit.t = t; hit.prim = p; hit.alpha = al; hit.beta = be; return true; }
                    if (t < tbest || (t == tbest && p < bestPrim)) { tbest = t; bestPrim = p; ba = al; bb = be; found = true; }
                }
            }
        }
        hit.t = tbest; hit.prim = found ? bestPrim : -1; hit.alpha = ba; hit.beta = bb;
        return found;
    }
    bool trace_brute(V3 o, V3 d, float tmin, float tmax, Hit& hit) const {
        Shear rs = make_shear(d);
        float tbest = tmax; int bestPrim = 0x7fffffff; bool found = false; float ba = 0, bb = 0;
        for (int p = 0; p < ntris; p++) {
            float t, al, be;
            if (tri_hit(rs, o, &verts[9 * (size_t)p], tmin, tmax, &t, &al, &be) && !cut(p, al, be))
                if (t < tbest || (t == tbest && p < bestPrim)) { tbest = t; bestPrim = p; ba = al; bb = be; found = true; }
        }
        hit.t = tbest; hit.prim = found ? bestPrim : -1; hit.alpha = ba; hit.beta = bb;
        return found;
    }
};

// ---- proxy MLP, fp32: trainingcode/module.py:755-837 (NeuralVisNetworkWith{4,6}Res256SingleOutput) ----
struct Mlp {
    int width = 0, nres = 0, half = 0, head = 0;   // head: 0 = LeakyReLU output, 1 = Sigmoid (module.py:880-958)
    std::vector<float> w;   // packed blob body
    const float *e3w0, *e3b0, *e3w1, *e3b1, *e2w0, *e2b0, *e2w1, *e2b1, *pw0, *pb0, *pw1, *pb1;
    std::vector<const float*> rw, rb;
    bool load(const void* blob, size_t bytes) {
        if (bytes < 16) return false;
        const uint32_t* h = (const uint32_t*)blob;
        if (h[0] != 0x50524d4cu) return false;     // 'LMRP'
        width = (int)h[1]; nres = (int)h[2]; half = width / 2; head = (int)(h[3] & 1u);
        size_t need = (size_t)(32 * 3 + 32 + half * 32 + half) + (size_t)(32 * 2 + 32 + half * 32 + half) +
                      (size_t)nres * ((size_t)width * width + width) + (size_t)(64 * width + 64) + 64 + 1;
        if (bytes != 16 + need * 4) return false;
        w.assign((const float*)(h + 4), (const float*)(h + 4) + need);
        const float* p = w.data();
        e3w0 = p; p += 32 * 3; e3b0 = p; p += 32; e3w1 = p; p += half * 32; e3b1 = p; p += half;
        e2w0 = p; p += 32 * 2; e2b0 = p; p += 32; e2w1 = p; p += half * 32; e2b1 = p; p += half;
        rw.resize(nres); rb.resize(nres);
        for (int i = 0; i < nres; i++) { rw[i] = p; p += (size_t)width * width; rb[i] = p; p += width; }
        pw0 = p; p += 64 * width; pb0 = p; p += 64; pw1 = p; p += 64; pb1 = p; p += 1;
        return true;
    }
    static float lrelu(float x) { return x > 0.f ? x : 0.01f * x; }   // F.leaky_relu default slope
    float forward_row(const float* x5) const {
        float h3[32], h2[32]; std::vector<float> a(width), b(width), out1(width);
        for (int j = 0; j < 32; j++) { float s = e3b0[j]; for (int k = 0; k < 3; k++) s += e3w0[j * 3 + k] * x5[k]; h3[j] = lrelu(s); }
        for (int j = 0; j < 32; j++) { float s = e2b0[j]; for (int k = 0; k < 2; k++) s += e2w0[j * 2 + k] * x5[3 + k]; h2[j] = lrelu(s); }
        for (int j = 0; j < half; j++) { float s = e3b1[j]; for (int k = 0; k < 32; k++) s += e3w1[j * 32 + k] * h3[k]; a[j] = lrelu(s); }
        for (int j = 0; j < half; j++) { float s = e2b1[j]; for (int k = 0; k < 32; k++) s += e2w1[j * 32 + k] * h2[k]; a[half + j] = lrelu(s); }
        out1 = a;
        for (int l = 0; l < nres; l++) {
            for (int j = 0; j < width; j++) {
                float s = rb[l][j]; const float* wr = rw[l] + (size_t)j * width;
                for (int k = 0; k < width; k++) s += wr[k] * a[k];
                b[j] = lrelu(a[j] + s);
            }
            a.swap(b);
        }
        for (int j = 0; j < width; j++) a[j] += out1[j];
        float z[64];
        for (int j = 0; j < 64; j++) { float s = pb0[j]; const float* wr = pw0 + (size_t)j * width; for (int k = 0; k < width; k++) s += wr[k] * a[k]; z[j] = lrelu(s); }
        float s = pb1[0]; for (int k = 0; k < 64; k++) s += pw1[k] * z[k];
        return head ? 1.0f / (1.0f + expf(-s)) : lrelu(s);
    }
};

// ---- textures: tex2D<float4> on a linear-filtered, normalised-coordinate texture object (renderer.cpp:1700-1710), restated in
// binary32 (the texture unit's fixed-point weights are not reproducible on a CPU; DESIGN.md "Arithmetic specification"):
// sample position x = u W - 0.5 between texel centres floor(x) and floor(x) + 1, weight frac(x); lerp(a, b, f) = fma(f, b - a, a)
// along u in both rows, then along v. u always wraps; v wraps (albedo maps) or clamps (lat-long environment map).
struct Texture {
    int w = 0, h = 0;
    std::vector<float> texels;    // RGBA, row-major, row 0 = v 0
    static int wrap_index(int i, int n) { if (i < 0) i += n; if (i >= n) i -= n; return i; }
    void sample(float u, float v, bool clampV, float* out4) const {
        if (!(fabsf(u) < 1e30f)) u = 0.0f;
        if (!(fabsf(v) < 1e30f)) v = 0.0f;
        u -= floorf(u);
        if (clampV) v = fminf(fmaxf(v, 0.0f), 1.0f); else v -= floorf(v);
        const float x = fmaf(u, (float)w, -0.5f), y = fmaf(v, (float)h, -0.5f);
        const float xf = floorf(x), yf = floorf(y);
        const float fx = x - xf, fy = y - yf;
        const int xi = (int)xf, yi = (int)yf;
        const int xa = wrap_index(xi, w), xb = wrap_index(xi + 1, w);
        const int ya = clampV ? std::min(std::max(yi, 0), h - 1) : wrap_index(yi, h);
        const int yb = clampV ? std::min(std::max(yi + 1, 0), h - 1) : wrap_index(yi + 1, h);
        const float* r0 = &texels[4 * ((size_t)ya * w)];
        const float* r1 = &texels[4 * ((size_t)yb * w)];
        for (int c = 0; c < 4; c++) {
            const float top = fmaf(fx, r0[4 * xb + c] - r0[4 * xa + c], r0[4 * xa + c]);
            const float bot = fmaf(fx, r1[4 * xb + c] - r1[4 * xa + c], r1[4 * xa + c]);
            out4[c] = fmaf(fy, bot - top, top);
        }
    }
};

struct World;
struct Rank;
// ---- world: all scene objects + W simulated ranks ----
struct Object {
    dprt_object_desc desc{};
    bool present = false;
    Mesh mesh;
    Mlp vis, depth; bool hasVis = false, hasDepth = false;
    // optional copy of the PRODUCT's BVH8 blob (orc_world_set_bvh8): never used for results, only walked a second time by a
    // scalar exact-tbest walker to count the nodes / triangles the algorithm has to touch per ray (bench.py roofline)
    std::vector<dprt_bvh8_node> nodes8; std::vector<dprt_bvh8_tri> tris8;
};

struct World;

struct Rank {
    World* w = nullptr; int id = 0;
    std::vector<dprt_path_record> paths, transfer;
    std::vector<int32_t> transferOffset, sceneOffset, hitPrim;
    std::vector<float> direct, env, contribution, occlusion;
    std::vector<dprt_half> nnInput, nnPackedInput, pred;
    std::vector<dprt_nn_query> nnQuery, nnPackedQuery;
    int pathSize = 0, shadowPathSize = 0, queryTotal = 0;
    dprt_stats stats{};
    int64_t bvhNodes = 0, bvhTris = 0;
    // per dprt_stage_id: BVH8 nodes fetched / triangles tested / rays walked by the counting walker
    int64_t cntNodes[DPRT_STAGE_COUNT] = {0}, cntTris[DPRT_STAGE_COUNT] = {0}, cntRays[DPRT_STAGE_COUNT] = {0};
};

struct World {
    dprt_config cfg{}; int W = 1; int N = 0; int sample = 0;
    std::vector<Object> objects;
    std::vector<dprt_material> materials;
    std::vector<dprt_light_tri> lights;
    dprt_camera cam{};
    std::vector<Rank> ranks;
    bool countBvh8 = false;          // orc_count_bvh8: walk the product's BVH8 beside every trace, for the counters only
    // real-scene front end: params.albedoTextures (renderer.cpp:1621-1721), HitGroupData.textureIndex per material
    // (pipeline_helper.cpp:185), params.envLightTexture (renderer.cpp:1851)
    std::vector<Texture> textures;
    std::vector<int32_t> matTex;
    Texture envMap; float envRotation = 0.0f;

    bool is_proxy(int rank, int i) const { return objects[i].desc.nodeID != rank; }
    // texture of a material, or null
    const Texture* texture_of(int matID) const {
        if (matID < 0 || matID >= (int)matTex.size()) return nullptr;
        const int t = matTex[matID];
        if (t < 0 || t >= (int)textures.size() || textures[t].texels.empty()) return nullptr;
        return &textures[t];
    }
    V3 env_radiance(V3 d) const {
        if (!envMap.texels.empty()) {
            // calculateEnvironmentLighting (kernel.cu:28-48): lat-long map at (phi / 2pi, theta / pi), phi rotated
            float phi, theta;
            cartesian_to_spherical(d, &phi, &theta);
            phi += envRotation;
            if (phi > 6.28318530717958647692f) phi -= 6.28318530717958647692f;
            float c[4];
            envMap.sample(phi / 6.28318530717958647692f, theta / 3.14159265358979323846f, true, c);
            return v3(c[0], c[1], c[2]);
        }
        float wv = fmaf(0.5f, d.z, 0.5f);
        return v3(cfg.envColor[0] * wv, cfg.envColor[1] * wv, cfg.envColor[2] * wv);
    }
};

// texture coordinate at the hit: gamma t0 + alpha t1 + beta t2 (kernel.cu:264-265), gamma = 1 - alpha - beta
inline float interp3(float t0, float t1, float t2, float alpha, float beta) {
    const float gamma = 1.0f - alpha - beta;
    return fmaf(beta, t2, fmaf(alpha, t1, gamma * t0));
}
inline void hit_uv(const Mesh& m, int prim, float alpha, float beta, float* u, float* v) {
    const float* t = &m.uvs[6 * (size_t)prim];
    *u = interp3(t[0], t[2], t[4], alpha, beta);
    *v = interp3(t[1], t[3], t[5], alpha, beta);
}
bool alpha_ignored(const World& w, const Mesh& m, int prim, float alpha, float beta) {
    const Texture* T = w.texture_of(m.mats[prim]);
    if (!T) return false;
    float u, v, c[4];
    hit_uv(m, prim, alpha, beta, &u, &v);
    T->sample(u, v, false, c);
    return c[3] < 0.05f;          // opacity below the cut-out threshold: kernel.cu:349-355
}

void add_env(World& w, Rank& r, dprt_path_record& p) {
    V3 e = w.env_radiance(v3(p.direction[0], p.direction[1], p.direction[2]));
    p.throughput[0] *= e.x; p.throughput[1] *= e.y; p.throughput[2] *= e.z;
    const int px = p.pixelIndex * 3;
    r.env[px + 0] += p.throughput[0]; r.env[px + 1] += p.throughput[1]; r.env[px + 2] += p.throughput[2];
}

// ---- scalar walker over the PRODUCT's BVH8 blob: per-ray node/triangle counters for the roofline ----
struct Bvh8Count { int64_t nodes, tris; };
// any = true: any-hit semantics, the walk stops at the first accepted triangle (near-first depth-first order)
bool bvh8_walk(const dprt_bvh8_node* nodes, const dprt_bvh8_tri* tris, V3 o, V3 d, float tmin, float tmax, Hit& hit, Bvh8Count& cnt, bool any = false,
               const Mesh* cutouts = nullptr) {
    Shear rs = make_shear(d);
    const float dxs = fabsf(d.x) > 1e-20f ? d.x : copysignf(1e-20f, d.x);
    const float dys = fabsf(d.y) > 1e-20f ? d.y : copysignf(1e-20f, d.y);
    const float dzs = fabsf(d.z) > 1e-20f ? d.z : copysignf(1e-20f, d.z);
    const float idir[3] = {1.0f / dxs, 1.0f / dys, 1.0f / dzs};
    const bool negd[3] = {dxs < 0.f, dys < 0.f, dzs < 0.f};
    const uint32_t octinv = 7u - ((negd[0] ? 1u : 0u) | (negd[1] ? 2u : 0u) | (negd[2] ? 4u : 0u));
    const float oo[3] = {o.x, o.y, o.z};
    float tbest = tmax; int bestPrim = 0x7fffffff; bool found = false; float ba = 0, bb = 0;
    struct Grp { uint32_t base, bits; };
    Grp stack[64]; int sp = 0;
    Grp ng{0u, 0x80000000u};
    uint32_t tmask = 0u;
    for (;;) {
        Grp tg{0u, 0u};
        if (ng.bits & 0xff000000u) {
            uint32_t bit = 31u - (uint32_t)__builtin_clz(ng.bits);
            uint32_t slot = (bit - 24u) ^ octinv;
            uint32_t rel = (uint32_t)__builtin_popcount(ng.bits & 0xffu & ((1u << slot) - 1u));
            ng.bits &= ~(1u << bit);
            const dprt_bvh8_node& n = nodes[ng.base + rel];
            if (ng.bits & 0xff000000u) stack[sp++] = ng;
            cnt.nodes++;
            float adj[3], org[3];
            for (int a = 0; a < 3; a++) {
                uint32_t eb = (uint32_t)n.e[a] << 23; float sc; std::memcpy(&sc, &eb, 4);
                adj[a] = sc * idir[a]; org[a] = (n.p[a] - oo[a]) * idir[a];
            }
            uint32_t hitmask = 0;
            const uint8_t* qlo[3] = {n.qlox, n.qloy, n.qloz}; const uint8_t* qhi[3] = {n.qhix, n.qhiy, n.qhiz};
            for (int c = 0; c < 8; c++) {
                float tn = tmin, tf = tbest;
                for (int a = 0; a < 3; a++) {
                    const float lo = (float)(negd[a] ? qhi[a][c] : qlo[a][c]), hi = (float)(negd[a] ? qlo[a][c] : qhi[a][c]);
                    tn = fmaxf(tn, fmaf(lo, adj[a], org[a])); tf = fminf(tf, fmaf(hi, adj[a], org[a]));
                }
                if (tn <= tf) {
                    if ((n.imask >> c) & 1u) hitmask |= 1u << (24u + ((uint32_t)c ^ octinv));   // internal child: octant-ordered slot
                    else hitmask |= (7u << (3 * c)) & n.tmask;                                  // leaf child: its (<= 3) triangles
                }
            }
            ng = Grp{n.childBase, (hitmask & 0xff000000u) | n.imask};
            tg = Grp{n.triBase, hitmask & 0x00ffffffu};
            tmask = n.tmask;
        }
        while (tg.bits) {
            uint32_t b = (uint32_t)__builtin_ctz(tg.bits); tg.bits &= tg.bits - 1u;
            const uint32_t k = (uint32_t)__builtin_popcount(tmask & ((1u << b) - 1u));
            const dprt_bvh8_tri& t = tris[tg.base + k];
            cnt.tris++;
            float tv[9] = {t.v0[0], t.v0[1], t.v0[2], t.v1[0], t.v1[1], t.v1[2], t.v2[0], t.v2[1], t.v2[2]};
            float tt, al, be;
            if (tri_hit(rs, o, tv, tmin, tmax, &tt, &al, &be)) {
                if (cutouts && cutouts->cut(t.primID, al, be)) continue;
                if (any) { hit.t = tt; hit.prim = t.primID; hit.alpha = al; hit.beta = be; return true; }
                if (tt < tbest || (tt == tbest && t.primID < bestPrim)) { tbest = tt; bestPrim = t.primID; ba = al; bb = be; found = true; }
            }
        }
        if ((ng.bits & 0xff000000u) == 0u) { if (sp == 0) break; ng = stack[--sp]; }
    }
    hit.t = tbest; hit.prim = found ? bestPrim : -1; hit.alpha = ba; hit.beta = bb;
    return found;
}

// Counting walk of one ray over the rank's local objects (same object order, same shrinking tMax as trace_local): what a
// scalar walker with an exact tbest fetches from the product's BVH8. Results are discarded; objects without a blob are skipped.
void count_local(const World& w, Rank& r, int stage, V3 o, V3 d, float tmin, float tMax, uint32_t visited, bool skipVisited, bool any);

// does a ray with this visited set walk any local BVH on this rank? (dprt_stats.rays_walked)
bool has_local_work(const World& w, int rank, uint32_t visited, bool skipVisited) {
    for (int i = 0; i < (int)w.objects.size(); i++) {
        const Object& ob = w.objects[i];
        if (!ob.present || w.is_proxy(rank, i)) continue;
        if (skipVisited && ((visited >> ob.desc.nodeID) & 1u)) continue;
        return true;
    }
    return false;
}

// closest hit over the rank's local objects, in scene order, tMax shrinking (strict < across objects)
bool trace_local(const World& w, int rank, V3 o, V3 d, float tmin, float& tMax, uint32_t visited, bool skipVisited, Hit& best,
                 int& bestObj) {
    bool any = false;
    for (int i = 0; i < (int)w.objects.size(); i++) {
        const Object& ob = w.objects[i];
        if (!ob.present || w.is_proxy(rank, i)) continue;
        if (skipVisited && ((visited >> ob.desc.nodeID) & 1u)) continue;
        Hit h;
        if (ob.mesh.trace(o, d, tmin, tMax, false, h)) { tMax = h.t; best = h; bestObj = i; any = true; }
    }
    return any;
}

void count_local(const World& w, Rank& r, int stage, V3 o, V3 d, float tmin, float tMax, uint32_t visited, bool skipVisited, bool any) {
    if (!w.countBvh8) return;
    Bvh8Count c{0, 0}; bool walkedAny = false;
    for (int i = 0; i < (int)w.objects.size(); i++) {
        const Object& ob = w.objects[i];
        if (!ob.present || w.is_proxy(r.id, i) || ob.nodes8.empty()) continue;
        if (skipVisited && ((visited >> ob.desc.nodeID) & 1u)) continue;
        walkedAny = true;
        Hit h;
        const bool hit = bvh8_walk(ob.nodes8.data(), ob.tris8.data(), o, d, tmin, tMax, h, c, any, &ob.mesh);
        if (hit) { if (any) break; tMax = h.t; }
    }
    if (!walkedAny) return;
    __atomic_fetch_add(&r.cntNodes[stage], c.nodes, __ATOMIC_RELAXED);
    __atomic_fetch_add(&r.cntTris[stage], c.tris, __ATOMIC_RELAXED);
    __atomic_fetch_add(&r.cntRays[stage], (int64_t)1, __ATOMIC_RELAXED);
}

void path_gen(World& w, Rank& r) {
    const int n = r.pathSize;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const int pixel = w.cfg.pathGenMode == 1 ? i * w.W + r.id : i;
        const int width = w.cam.width;
        const int row = pixel / width, col = pixel - row * width;
        uint32_t seed = tea4((uint32_t)pixel, (uint32_t)w.sample);
        const float xi1 = rnd(seed), xi2 = rnd(seed);
        const float a = fmaf(2.0f, ((float)col + xi1) / (float)width, -1.0f);
        const float b = fmaf(-2.0f, ((float)row + xi2) / (float)w.cam.height, 1.0f);
        const dprt_camera& c = w.cam;
        V3 dir = v3(fmaf(c.U[0], a, fmaf(c.V[0], b, c.W[0])), fmaf(c.U[1], a, fmaf(c.V[1], b, c.W[1])),
                    fmaf(c.U[2], a, fmaf(c.V[2], b, c.W[2])));
        dir = normalized(dir);
        dprt_path_record p; std::memset(&p, 0, sizeof(p));
        p.origin[0] = c.origin[0]; p.origin[1] = c.origin[1]; p.origin[2] = c.origin[2];
        p.direction[0] = dir.x; p.direction[1] = dir.y; p.direction[2] = dir.z;
        p.tMax = FLT_MAX; p.throughput[0] = p.throughput[1] = p.throughput[2] = 1.f;
        p.pixelIndex = pixel; p.shadowPathID = -1; p.visitedMask = 0; p.currentNode = -1; p.targetNode = -1;
        p.isValid = 1;
        r.paths[i] = p;
    }
}

void traverse(World& w, Rank& r) {
    const int n = r.pathSize;
    int64_t walked = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : walked)
    for (int i = 0; i < n; i++) {
        dprt_path_record p = r.paths[i];
        if (!r.hitPrim.empty()) r.hitPrim[i] = -1;
        if (!p.isValid) continue;
        walked += has_local_work(w, r.id, p.visitedMask, true);
        const V3 o = v3(p.origin[0], p.origin[1], p.origin[2]), d = v3(p.direction[0], p.direction[1], p.direction[2]);
        Hit h; h.prim = -1; int hobj = -1; float tMax = p.tMax;
        count_local(w, r, DPRT_STAGE_TRAVERSE, o, d, DPRT_EPSILON, p.tMax, p.visitedMask, true, false);
        if (trace_local(w, r.id, o, d, DPRT_EPSILON, tMax, p.visitedMask, true, h, hobj)) {
            p.tMax = tMax; p.isHit = 1; p.currentNode = r.id;
        }
        if (!r.hitPrim.empty()) r.hitPrim[i] = hobj >= 0 ? h.prim : -1;
        p.visitedMask |= (1u << r.id);
        float tProxy = p.tMax; bool proxyHit = false;
        for (int k = 0; k < (int)w.objects.size(); k++) {
            const Object& ob = w.objects[k];
            if (!ob.present || !w.is_proxy(r.id, k)) continue;
            if ((p.visitedMask >> ob.desc.nodeID) & 1u) continue;
            const V3 ol = xform_point(ob.desc.worldToObject, o), dl = xform_vector(ob.desc.worldToObject, d);
            float t; bool inside;
            if (aabb_hit(ol, dl, ob.desc.aabbMin, ob.desc.aabbMax, DPRT_EPSILON, tProxy, &t, &inside)) {
                tProxy = t; proxyHit = true; p.targetNode = ob.desc.nodeID;
            }
        }
        if (!proxyHit) p.targetNode = p.currentNode;
        if (!proxyHit && !p.isHit) { add_env(w, r, p); p.isValid = 0; }
        r.paths[i] = p;
    }
    r.stats.rays_traverse += n; r.stats.rays_walked += walked; r.stats.walked_traverse += walked;
}

// Work_Efficient_Scan: stable partition of valid paths by targetNode (bucket-major, index order inside)
void partition(World& w, Rank& r) {
    const int n = r.pathSize;
    std::fill(r.transferOffset.begin(), r.transferOffset.end(), 0);
    int out = 0;
    for (int b = 0; b < w.W; b++) {
        r.transferOffset[b] = out;
        for (int i = 0; i < n; i++)
            if (r.paths[i].isValid && r.paths[i].targetNode == b) r.transfer[out++] = r.paths[i];
    }
    r.transferOffset[w.W] = out;
}

// MPI_Alltoall + MPI_Alltoallv + MPI_Allreduce(LAND): renderer.cpp:1254-1314
bool exchange(World& w) {
    long offdiag = 0;
    std::vector<std::vector<dprt_path_record>> recv(w.W);
    for (int d = 0; d < w.W; d++)
        for (int s = 0; s < w.W; s++) {
            const Rank& rs = w.ranks[s];
            const int a = rs.transferOffset[d], b = rs.transferOffset[d + 1];
            if (s != d) { offdiag += b - a; w.ranks[s].stats.paths_sent_offrank += b - a; }
            recv[d].insert(recv[d].end(), rs.transfer.begin() + a, rs.transfer.begin() + b);
        }
    for (int d = 0; d < w.W; d++) {
        Rank& r = w.ranks[d];
        std::copy(recv[d].begin(), recv[d].end(), r.paths.begin());
        r.pathSize = (int)recv[d].size();
        r.stats.exchange_iters++;
    }
    return offdiag == 0;
}

struct Bsdf { V3 wiLocal; float weight; bool isDelta; };
Bsdf sample_lambertian(float xi1, float xi2) { Bsdf s; s.wiLocal = uniform_hemisphere(xi1, xi2); s.weight = 2.0f; s.isDelta = false; return s; }
float fresnel_dielectric(float cosI, float etaI, float etaT) {
    const float eta = etaI / etaT;
    const float sin2t = (eta * eta) * fmaxf(0.0f, fmaf(-cosI, cosI, 1.0f));
    if (sin2t >= 1.0f) return 1.0f;
    const float cosT = sqrtf(1.0f - sin2t);
    const float rParl = (etaT * cosI - etaI * cosT) / (etaT * cosI + etaI * cosT);
    const float rPerp = (etaI * cosI - etaT * cosT) / (etaI * cosI + etaT * cosT);
    return 0.5f * (rParl * rParl + rPerp * rPerp);
}
Bsdf sample_water(float xi1, V3 normal, V3 woWorld, bool isInside) {
    const V3 wo = to_local(make_frame(normal), woWorld);
    float etaI = 1.0f, etaT = 1.33f;
    if (isInside) std::swap(etaI, etaT);
    V3 wi = v3(0, 0, 0);
    {
        const float eta = etaI / etaT;
        const float sin2t = (eta * eta) * fmaxf(0.0f, fmaf(-wo.z, wo.z, 1.0f));
        if (sin2t < 1.0f) { const float cosT = sqrtf(1.0f - sin2t); wi = v3(-(eta * wo.x), -(eta * wo.y), -cosT); }
    }
    const float fr = fresnel_dielectric(fabsf(wo.z), etaI, etaT);
    Bsdf s; s.isDelta = true;
    if (xi1 < fr) {
        wi = v3(-wo.x, -wo.y, wo.z);
        const float ct = fabsf(wi.z);
        const float thr = ct == 0.0f ? 0.0f : fr / ct;
        s.wiLocal = wi; s.weight = thr / fr;
    } else {
        const float ft = 1.0f - fr;
        const float ct = fabsf(wi.z);
        const float thr = ct == 0.0f ? 0.0f : ft / ct;
        const float corr = (etaI / etaT) * (etaI / etaT);
        s.wiLocal = wi; s.weight = thr * corr / ft;
    }
    return s;
}

void shade(World& w, Rank& r) {
    const int n = r.pathSize, spc = w.cfg.shadowPathCount;
    r.shadowPathSize = spc * n;
    const dprt_path_record zero{};
    int64_t walked = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : walked)
    for (int i = 0; i < n; i++) {
        dprt_path_record path = r.paths[i];
        if (!r.hitPrim.empty()) r.hitPrim[i] = -1;
        if (!path.isValid) continue;
        walked += has_local_work(w, r.id, 0u, false);     // the oracle always re-traces (kernel.cu:382-413)
        const V3 o = v3(path.origin[0], path.origin[1], path.origin[2]), d = v3(path.direction[0], path.direction[1], path.direction[2]);
        Hit h{}; int hobj = -1; float tMax = FLT_MAX;
        {   // counted with the bound a walker may use without changing the result: "hit on this rank at tMax" -> next float up
            float tc = FLT_MAX;
            if (path.isHit && path.currentNode == r.id && path.tMax > 0.0f && path.tMax < FLT_MAX) {
                uint32_t b; std::memcpy(&b, &path.tMax, 4); b += 1u; std::memcpy(&tc, &b, 4);
            }
            count_local(w, r, DPRT_STAGE_SHADE, o, d, DPRT_EPSILON, tc, 0u, false, false);
        }
        const bool isHit = trace_local(w, r.id, o, d, DPRT_EPSILON, tMax, 0u, false, h, hobj);
        if (!r.hitPrim.empty()) r.hitPrim[i] = isHit ? h.prim : -1;
        if (!isHit) {
            add_env(w, r, path);
            r.paths[i] = zero;
            for (int s = 0; s < spc; s++) r.paths[(size_t)i * spc + s + n] = zero;
            continue;
        }
        const Object& ob = w.objects[hobj];
        const dprt_material mat = w.materials[ob.mesh.mats[h.prim]];
        V3 albedo = v3(mat.baseColor[0], mat.baseColor[1], mat.baseColor[2]);
        if (!ob.mesh.uvs.empty()) {      // kernel.cu:251-281: textured material -> base colour from the albedo map
            if (const Texture* T = w.texture_of(ob.mesh.mats[h.prim])) {
                float u, v, c[4];
                hit_uv(ob.mesh, h.prim, h.alpha, h.beta, &u, &v);
                T->sample(u, v, false, c);
                albedo = v3(c[0], c[1], c[2]);
            }
        }
        const V3 point = at(o, d, h.t);
        const V3 woWorld = neg(d);
        const float* nn = &ob.mesh.normals[9 * (size_t)h.prim];
        const V3 n0 = normalized(v3(nn[0], nn[1], nn[2])), n1 = normalized(v3(nn[3], nn[4], nn[5])), n2 = normalized(v3(nn[6], nn[7], nn[8]));
        const float alpha = h.alpha, beta = h.beta, gamma = 1.0f - alpha - beta;
        V3 normal = v3(fmaf(beta, n2.x, fmaf(alpha, n1.x, gamma * n0.x)), fmaf(beta, n2.y, fmaf(alpha, n1.y, gamma * n0.y)),
                       fmaf(beta, n2.z, fmaf(alpha, n1.z, gamma * n0.z)));
        normal = normalized(normal);
        bool isInside = false;
        if (dot(normal, woWorld) < 0.0f) { normal = neg(normal); isInside = true; }
        uint32_t seed = tea4((uint32_t)path.pixelIndex, (uint32_t)w.sample);
        const float xi1 = rnd(seed), xi2 = rnd(seed);
        const Bsdf bs = mat.bsdfType == 1 ? sample_water(xi1, normal, woWorld, isInside) : sample_lambertian(xi1, xi2);

        dprt_path_record next{};
        const V3 nd = normalized(to_world(make_frame(normal), bs.wiLocal));
        next.origin[0] = point.x; next.origin[1] = point.y; next.origin[2] = point.z;
        next.direction[0] = nd.x; next.direction[1] = nd.y; next.direction[2] = nd.z;
        next.tMax = FLT_MAX;
        const float cosThetaWi = fabsf(bs.wiLocal.z);
        const V3 thr = mul(scale(scale(v3(path.throughput[0], path.throughput[1], path.throughput[2]), bs.weight), cosThetaWi), albedo);
        next.throughput[0] = thr.x; next.throughput[1] = thr.y; next.throughput[2] = thr.z;
        next.pixelIndex = path.pixelIndex; next.shadowPathID = -1; next.visitedMask = 0; next.currentNode = -1; next.targetNode = -1;
        next.isValid = 1;
        r.paths[i] = next;

        for (int s = 0; s < spc; s++) {
            dprt_path_record& slot = r.paths[(size_t)i * spc + s + n];
            if (bs.isDelta) { slot = zero; continue; }
            uint32_t sseed = tea4((uint32_t)(path.pixelIndex * spc + s), (uint32_t)w.sample);
            const float x1 = rnd(sseed), x2 = rnd(sseed), x3 = rnd(sseed);
            const int li = (int)floorf(x1 * (float)w.lights.size());
            const dprt_light_tri& L = w.lights[li];
            const V3 p0 = v3(L.p0[0], L.p0[1], L.p0[2]), p1 = v3(L.p1[0], L.p1[1], L.p1[2]), p2 = v3(L.p2[0], L.p2[1], L.p2[2]);
            const float su = sqrtf(x2);
            const float b0 = 1.0f - su, b1 = x3 * su, b2 = 1.0f - b0 - b1;
            const V3 lp = v3(fmaf(b2, p2.x, fmaf(b1, p1.x, b0 * p0.x)), fmaf(b2, p2.y, fmaf(b1, p1.y, b0 * p0.y)),
                             fmaf(b2, p2.z, fmaf(b1, p1.z, b0 * p0.z)));
            const V3 cr = cross(sub(p1, p0), sub(p2, p0));
            const float crl = length(cr);
            const V3 ln = scale(cr, 1.0f / crl);
            float areaPDF = 1.0f / (0.5f * crl);
            areaPDF = areaPDF * (1.0f / (float)w.lights.size());
            const V3 ldir = sub(lp, point);
            const V3 wi = normalized(ldir);
            const float stMax = length(ldir);
            const float f1 = fmaxf(0.0f, dot(ln, neg(wi)));
            const float f2 = fmaxf(0.0f, dot(wi, normal));
            V3 c = mul(mul(v3(L.Le[0], L.Le[1], L.Le[2]), v3(path.throughput[0], path.throughput[1], path.throughput[2])), albedo);
            c = scale(scale(c, f1), f2);
            const float tt = stMax * stMax;
            c = v3(c.x / areaPDF / tt, c.y / areaPDF / tt, c.z / areaPDF / tt);
            c = scale(c, 0.318309886183790671538f);
            dprt_path_record sh{};
            sh.origin[0] = point.x; sh.origin[1] = point.y; sh.origin[2] = point.z;
            sh.direction[0] = wi.x; sh.direction[1] = wi.y; sh.direction[2] = wi.z;
            sh.tMax = stMax; sh.throughput[0] = c.x; sh.throughput[1] = c.y; sh.throughput[2] = c.z;
            sh.pixelIndex = path.pixelIndex; sh.shadowPathID = s; sh.visitedMask = 0; sh.currentNode = -1; sh.targetNode = -1;
            sh.isShadowRay = 1; sh.isValid = 1;
            slot = sh;
        }
    }
    r.stats.rays_shade += n; r.stats.rays_walked += walked; r.stats.walked_shade += walked;
}

// proxy-AABB march shared by ShadowRay / SecondaryRay; returns -1 when nothing was in the way
int proxy_march(World& w, Rank& r, const dprt_path_record& path, int threadIndex, float tMaxPath, bool secondary) {
    const V3 o = v3(path.origin[0], path.origin[1], path.origin[2]), d = v3(path.direction[0], path.direction[1], path.direction[2]);
    const int mc = w.cfg.maxCount;
    bool isHit = true, isInside = false, nothing = false; float tMin = 0.0f; int count = 0, hitIdx = -1;
    while (isHit && count < mc) {
        isHit = false;
        float tMax = tMaxPath;
        V3 pl = v3(0, 0, 0), oloc = v3(0, 0, 0), dloc = v3(0, 0, 0);
        for (int k = 0; k < (int)w.objects.size(); k++) {
            const Object& ob = w.objects[k];
            if (!ob.present || !w.is_proxy(r.id, k)) continue;
            const V3 ol = xform_point(ob.desc.worldToObject, o), dl = xform_vector(ob.desc.worldToObject, d);
            float t; bool ins;
            if (aabb_hit(ol, dl, ob.desc.aabbMin, ob.desc.aabbMax, tMin + DPRT_EPSILON, tMax, &t, &ins)) {
                tMax = t; isHit = true; hitIdx = k; isInside = ins; oloc = ol; dloc = dl;
                pl = xform_point(ob.desc.worldToObject, at(o, d, t));
            }
        }
        if (isHit) tMin = tMax;
        if (isHit) {
            const dprt_object_desc& od = w.objects[hitIdx].desc;
            if (isInside) {
                bool skip = false;
                for (int q = 0; q < count; q++) {
                    const dprt_nn_query& e = r.nnQuery[(size_t)threadIndex * mc + q];
                    if (e.hitAABBID == hitIdx + 1 && e.instanceID == hitIdx) skip = true;
                }
                if (skip && count) continue;
            }
            const V3 dirL = isInside ? neg(dloc) : dloc;
            float phi, theta;
            cartesian_to_spherical(normalized(dirL), &phi, &theta);
            dprt_half* in = &r.nnInput[((size_t)threadIndex * mc + count) * 5];
            in[0] = f2h((pl.x - od.aabbMin[0]) / (od.aabbMax[0] - od.aabbMin[0]));
            in[1] = f2h((pl.y - od.aabbMin[1]) / (od.aabbMax[1] - od.aabbMin[1]));
            in[2] = f2h((pl.z - od.aabbMin[2]) / (od.aabbMax[2] - od.aabbMin[2]));
            in[3] = f2h(phi / 6.28318530717958647692f);
            in[4] = f2h(theta / 3.14159265358979323846f);
            dprt_nn_query e{};
            e.pixelIndex = path.pixelIndex; e.hitSequence = count; e.hitAABBID = hitIdx + 1; e.isValid = 1;
            e.instanceID = hitIdx; e.isInside = isInside ? 1 : 0;
            const float dist = length(sub(oloc, pl));
            if (secondary) {
                e.throughput[0] = tMax; e.throughput[1] = od.maxLength; e.throughput[2] = tMax / dist;
                e.shadowPathID = 0; e.pathIndex = od.nodeID;
                e.normalizedT = isInside ? tMax / od.maxLength : 0.0f;
            } else {
                e.throughput[0] = path.throughput[0]; e.throughput[1] = path.throughput[1]; e.throughput[2] = path.throughput[2];
                e.shadowPathID = path.shadowPathID;
                e.pathIndex = isInside ? threadIndex * mc + count : 0;
                e.normalizedT = isInside ? dist / od.maxLength : 0.0f;
            }
            r.nnQuery[(size_t)threadIndex * mc + count] = e;
            count++;
        } else if (count == 0) {
            nothing = true;
        }
    }
    for (int q = count; q < mc; q++) r.nnQuery[(size_t)threadIndex * mc + q].hitAABBID = 0;
    return nothing ? -1 : count;
}

void clear_slots(World& w, Rank& r, int threadIndex) {
    for (int q = 0; q < w.cfg.maxCount; q++) r.nnQuery[(size_t)threadIndex * w.cfg.maxCount + q].hitAABBID = 0;
}

void shadow_trace(World& w, Rank& r) {
    const int n = r.shadowPathSize, spc = w.cfg.shadowPathCount;
    int64_t walked = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : walked)
    for (int i = 0; i < n; i++) {
        dprt_path_record& path = r.paths[(size_t)r.pathSize + i];
        if (!path.isValid) { if (w.cfg.proxyMode) clear_slots(w, r, i); continue; }
        walked += has_local_work(w, r.id, 0u, false);
        const V3 o = v3(path.origin[0], path.origin[1], path.origin[2]), d = v3(path.direction[0], path.direction[1], path.direction[2]);
        bool occluded = false;
        count_local(w, r, DPRT_STAGE_SHADOW_TRACE, o, d, DPRT_EPSILON, path.tMax, 0u, false, true);
        for (int k = 0; k < (int)w.objects.size() && !occluded; k++) {
            const Object& ob = w.objects[k];
            if (!ob.present || w.is_proxy(r.id, k)) continue;
            Hit h;
            if (ob.mesh.trace(o, d, DPRT_EPSILON, path.tMax, true, h)) occluded = true;
        }
        if (occluded) { path.isHit = 1; path.isValid = 0; if (w.cfg.proxyMode) clear_slots(w, r, i); continue; }
        int res = -1;
        if (w.cfg.proxyMode) res = proxy_march(w, r, path, i, path.tMax, false);
        if (res < 0) {
            // distinct (pixel, shadowPathID) per thread: race-free like the reference's per-plane layout
            const size_t px = ((size_t)w.N * path.shadowPathID + path.pixelIndex) * 3;
            const float inv = (float)spc;
            r.direct[px + 0] += path.throughput[0] / inv;
            r.direct[px + 1] += path.throughput[1] / inv;
            r.direct[px + 2] += path.throughput[2] / inv;
        }
    }
    r.stats.rays_shadow += n; r.stats.rays_walked += walked; r.stats.walked_shadow += walked;
}

void secondary_trace(World& w, Rank& r) {
    const int n = r.pathSize;
    int64_t walked = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : walked)
    for (int i = 0; i < n; i++) {
        dprt_path_record path = r.paths[i];
        if (!r.hitPrim.empty()) r.hitPrim[i] = -1;
        if (!path.isValid) { clear_slots(w, r, i); continue; }
        walked += has_local_work(w, r.id, 0u, false);
        for (size_t k = 0; k < w.objects.size(); k++) if (w.objects[k].present) path.visitedMask |= (1u << w.objects[k].desc.nodeID);
        const V3 o = v3(path.origin[0], path.origin[1], path.origin[2]), d = v3(path.direction[0], path.direction[1], path.direction[2]);
        Hit h; int hobj = -1; float tMax = path.tMax;
        count_local(w, r, DPRT_STAGE_SECONDARY_TRACE, o, d, DPRT_EPSILON, path.tMax, 0u, false, false);
        if (trace_local(w, r.id, o, d, DPRT_EPSILON, tMax, 0u, false, h, hobj)) { path.tMax = tMax; path.isHit = 1; path.currentNode = r.id; }
        if (!r.hitPrim.empty()) r.hitPrim[i] = hobj >= 0 ? h.prim : -1;
        const int res = proxy_march(w, r, path, i, path.tMax, true);
        if (res < 0 && !path.isHit) { add_env(w, r, path); path.isValid = 0; }
        r.paths[i] = path;
    }
    r.stats.rays_secondary += n; r.stats.rays_walked += walked; r.stats.walked_secondary += walked;
}

// Work_Efficient_Scan_For_NN(_HIT_INSIDE): stable bucket of queries by hitAABBID
int bucket_queries(World& w, Rank& r, int which, int insideOnly) {
    const int n = w.cfg.maxCount * (which == 0 ? r.shadowPathSize : r.pathSize);
    const int S = (int)w.objects.size();
    int out = 0;
    for (int b = 0; b < S; b++) {
        r.sceneOffset[b] = out;
        for (int i = 0; i < n; i++) {
            const dprt_nn_query& q = r.nnQuery[i];
            if (q.hitAABBID != b + 1) continue;
            if (insideOnly && !q.isInside) continue;
            r.nnPackedQuery[out] = q;
            for (int k = 0; k < 5; k++) r.nnPackedInput[(size_t)out * 5 + k] = r.nnInput[(size_t)i * 5 + k];
            out++;
        }
    }
    r.sceneOffset[S] = out; r.queryTotal = out;
    return out;
}

void proxy_infer(World& w, Rank& r, int kind, int predOffset) {
    const int S = (int)w.objects.size();
    for (int i = 0; i < r.queryTotal; i++) r.pred[(size_t)predOffset + i] = 0;
    for (int b = 0; b < S; b++) {
        const Object& ob = w.objects[b];
        const bool has = kind == 0 ? ob.hasVis : ob.hasDepth;
        if (!has) continue;
        const Mlp& m = kind == 0 ? ob.vis : ob.depth;
        const int a = r.sceneOffset[b], e = r.sceneOffset[b + 1];
#pragma omp parallel for schedule(static)
        for (int i = a; i < e; i++) {
            float x[5];
            for (int k = 0; k < 5; k++) x[k] = h2f(r.nnPackedInput[(size_t)i * 5 + k]);
            r.pred[(size_t)predOffset + i] = f2h(m.forward_row(x));
        }
        r.stats.nn_queries += e - a;
    }
}

void frame_buffer_update(World& w, Rank& r) {
    const int spc = w.cfg.shadowPathCount, mc = w.cfg.maxCount, N = w.N;
    for (int i = 0; i < (w.cfg.proxyMode ? r.queryTotal : 0); i++) {
        const dprt_nn_query& q = r.nnPackedQuery[i];
        if (!q.isValid) continue;
        const size_t slot = (size_t)q.pixelIndex * spc + q.shadowPathID;
        for (int k = 0; k < 3; k++) r.contribution[slot * 3 + k] = q.throughput[k];
        const float pv = h2f(r.pred[i]);
        float flag = pv > 0.5f ? 1.0f : 0.0f;
        if (q.isInside && pv > 0.5f) flag = q.normalizedT;
        r.occlusion[slot * mc + q.hitSequence] = flag;
    }
#pragma omp parallel for schedule(static)
    for (int px = 0; px < N; px++) {
        float d0 = r.direct[(size_t)px * 3 + 0], d1 = r.direct[(size_t)px * 3 + 1], d2 = r.direct[(size_t)px * 3 + 2];
        for (int s = 0; s < spc; s++) {
            const size_t slot = (size_t)px * spc + s;
            float maxOcc = 0.0f;
            for (int j = 0; j < mc; j++) { const float o = r.occlusion[slot * mc + j]; maxOcc = maxOcc > o ? maxOcc : o; }
            const float wv = 1.0f - maxOcc;
            d0 += r.contribution[slot * 3 + 0] * wv / (float)spc;
            d1 += r.contribution[slot * 3 + 1] * wv / (float)spc;
            d2 += r.contribution[slot * 3 + 2] * wv / (float)spc;
        }
        for (int s = 1; s < spc; s++) {
            const size_t pl = ((size_t)N * s + px) * 3;
            d0 += r.direct[pl + 0]; d1 += r.direct[pl + 1]; d2 += r.direct[pl + 2];
        }
        r.direct[(size_t)px * 3 + 0] = d0; r.direct[(size_t)px * 3 + 1] = d1; r.direct[(size_t)px * 3 + 2] = d2;
    }
}

void depth_buffer_update(World& w, Rank& r) {
    for (int i = 0; i < r.queryTotal; i++) {
        const dprt_nn_query& q = r.nnPackedQuery[i];
        r.nnQuery[q.pathIndex].normalizedT = h2f(r.pred[i]) > q.normalizedT ? 0.0f : 1.0f;
    }
}

void target_node_update(World& w, Rank& r) {
    const int mc = w.cfg.maxCount, size = r.queryTotal;
    for (int i = 0; i < size; i++) {
        const dprt_nn_query& q = r.nnPackedQuery[i];
        if (!q.isValid) continue;
        const size_t ti = ((size_t)q.pixelIndex * mc + q.hitSequence) * 2;
        float v = 0.0f;
        if (h2f(r.pred[i]) > 0.5f) {
            const float predMax = q.throughput[2] * q.throughput[1] * h2f(r.pred[(size_t)i + size]);
            const float aabbMax = q.throughput[0];
            if (q.isInside) v = predMax > aabbMax ? 0.0f : (aabbMax - predMax);
            else v = aabbMax + predMax;
        }
        r.occlusion[ti] = v; r.occlusion[ti + 1] = (float)q.pathIndex;
    }
    for (int i = 0; i < r.pathSize; i++) {
        dprt_path_record& p = r.paths[i];
        if (!p.isValid) continue;
        float tMax = p.tMax; int cur = p.currentNode;
        for (int j = 0; j < mc; j++) {
            const size_t ti = ((size_t)p.pixelIndex * mc + j) * 2;
            const float t = r.occlusion[ti];
            if (t < FLT_EPSILON) continue;
            if (tMax > t) { tMax = t; cur = (int)r.occlusion[ti + 1]; }
        }
        if (cur >= 0) { p.currentNode = cur; p.targetNode = cur; p.isHit = 1; p.tMax = tMax; }
        else { p.targetNode = r.id; p.tMax = 0.0f; p.isHit = 0; p.isValid = 1; }
    }
}

void reset_nn(World& w, Rank& r) {
    std::fill(r.occlusion.begin(), r.occlusion.end(), 0.f);
    std::fill(r.contribution.begin(), r.contribution.end(), 0.f);
    if (w.cfg.shadowPathCount > 1) std::fill(r.direct.begin() + (size_t)w.N * 3, r.direct.end(), 0.f);
}

void begin_sample(World& w, int sample) {
    w.sample = sample;
    for (Rank& r : w.ranks) {
        if (w.cfg.pathGenMode == 1) r.pathSize = (w.N - r.id + w.W - 1) / w.W;
        else r.pathSize = r.id == 0 ? w.N : 0;
        r.shadowPathSize = 0;
    }
}

void shadow_module(World& w, Rank& r) {
    shadow_trace(w, r);
    if (w.cfg.proxyMode) {
        bucket_queries(w, r, 0, 1); proxy_infer(w, r, 1, 0); depth_buffer_update(w, r);
        bucket_queries(w, r, 0, 0); proxy_infer(w, r, 0, 0);
    } else r.queryTotal = 0;
    frame_buffer_update(w, r);
}
void secondary_module(World& w, Rank& r) {
    secondary_trace(w, r);
    const int total = bucket_queries(w, r, 1, 0);
    proxy_infer(w, r, 0, 0); proxy_infer(w, r, 1, total);
    target_node_update(w, r);
}

void render_sample(World& w, int sample) {
    begin_sample(w, sample);
    for (Rank& r : w.ranks) path_gen(w, r);
    for (int bounce = 0; bounce <= w.cfg.bounces; bounce++) {
        if (bounce > 0 && w.cfg.proxyMode) for (Rank& r : w.ranks) { reset_nn(w, r); secondary_module(w, r); }
        for (;;) {
            for (Rank& r : w.ranks) { traverse(w, r); partition(w, r); }
            if (exchange(w)) break;
        }
        for (Rank& r : w.ranks) { shade(w, r); reset_nn(w, r); shadow_module(w, r); }
    }
}

Rank* get_rank(World* w, int rank) { return (w && rank >= 0 && rank < w->W) ? &w->ranks[rank] : nullptr; }

}  // namespace

// ================================================================================================
extern "C" {

// RNG known-answer access: seeds = tea4(val0,val1); out[k] = k-th rnd()
uint32_t orc_tea4(uint32_t val0, uint32_t val1) { return tea4(val0, val1); }
void orc_rnd_sequence(uint32_t seed, int n, float* out) { for (int i = 0; i < n; i++) out[i] = rnd(seed); }
void orc_sincos2pi(const float* x, int n, float* s, float* c) { for (int i = 0; i < n; i++) det_sincos2pi(x[i], s + i, c + i); }
void orc_acos(const float* x, int n, float* y) { for (int i = 0; i < n; i++) y[i] = det_acos(x[i]); }
void orc_atan2(const float* y, const float* x, int n, float* r) { for (int i = 0; i < n; i++) r[i] = det_atan2(y[i], x[i]); }
void orc_f2h(const float* x, int n, uint16_t* y) { for (int i = 0; i < n; i++) y[i] = f2h(x[i]); }

void* orc_world_create(const dprt_config* cfg, int W) {
    World* w = new World();
    w->cfg = *cfg; w->W = W; w->N = cfg->width * cfg->height;
    w->objects.resize(cfg->sceneSize);
    w->ranks.resize(W);
    const size_t N = w->N, spc = cfg->shadowPathCount, mc = cfg->maxCount;
    const size_t Q = cfg->proxyMode ? N * mc * spc : 1;
    for (int k = 0; k < W; k++) {
        Rank& r = w->ranks[k];
        r.w = w; r.id = k;
        r.paths.assign((1 + spc) * N, dprt_path_record{}); r.transfer.assign(N, dprt_path_record{});
        r.transferOffset.assign(64, 0); r.sceneOffset.assign(64, 0);
        r.direct.assign(spc * 3 * N, 0.f); r.env.assign(3 * N, 0.f);
        r.contribution.assign(3 * N * spc, 0.f); r.occlusion.assign(N * mc * std::max<size_t>(spc, 2), 0.f);
        r.nnInput.assign(Q * 5, 0); r.nnPackedInput.assign(Q * 5, 0); r.pred.assign(Q * 4, 0);
        r.nnQuery.assign(Q, dprt_nn_query{}); r.nnPackedQuery.assign(Q, dprt_nn_query{});
    }
    return w;
}
void orc_world_destroy(void* wp) { delete (World*)wp; }

int orc_world_add_object(void* wp, int si, const dprt_object_desc* desc, const float* verts9, const float* normals9,
                         const int32_t* mat_ids, int64_t ntris) {
    World* w = (World*)wp;
    if (!w || si < 0 || si >= (int)w->objects.size() || !desc) return -1;
    Object& ob = w->objects[si];
    ob.desc = *desc; ob.present = true;
    ob.mesh.ntris = (int)ntris;
    ob.mesh.verts.assign(verts9, verts9 + 9 * ntris);
    if (normals9) ob.mesh.normals.assign(normals9, normals9 + 9 * ntris); else ob.mesh.normals.assign(9 * ntris, 0.f);
    if (mat_ids) ob.mesh.mats.assign(mat_ids, mat_ids + ntris); else ob.mesh.mats.assign(ntris, 0);
    ob.mesh.uvs.clear(); ob.mesh.world = w;
    ob.mesh.build();
    return 0;
}
// dprt_upload_chunk_uv: per-corner texture coordinates beside the geometry (uv6 may be NULL)
int orc_world_add_object_uv(void* wp, int si, const dprt_object_desc* desc, const float* verts9, const float* normals9, const float* uv6,
                            const int32_t* mat_ids, int64_t ntris) {
    if (orc_world_add_object(wp, si, desc, verts9, normals9, mat_ids, ntris)) return -1;
    Mesh& m = ((World*)wp)->objects[si].mesh;
    if (uv6) m.uvs.assign(uv6, uv6 + 6 * ntris); else m.uvs.clear();
    return 0;
}

// Indexed, instanced meshes -> flat per-corner streams: the oracle's own restatement of what the reference does per hit
// (kernel.cu:205-240: corners through optixTransformPointFromObjectToWorldSpace, normals through
// optixTransformNormalFromObjectToWorldSpace = inverse transpose, attributes through normalIndices / texCoordsIndices),
// applied once per (instance, triangle). Primitive id = running index over instances, then triangles.
// Arithmetic: corner row = fma(m2, z, fma(m1, y, fma(m0, x, m3))); normal matrix = cofactors(A) / det(A) in binary64
// (products and differences in the order written), rounded to binary32; normal row = fma(g2, nz, fma(g1, ny, g0 nx)).
static int64_t flat_count(const dprt_mesh_desc* meshes, int nm, const dprt_instance_desc* inst, int64_t ni) {
    if (!meshes || !inst || nm < 1 || ni < 1) return -1;
    int64_t total = 0;
    for (int64_t i = 0; i < ni; i++) {
        if (inst[i].mesh < 0 || inst[i].mesh >= nm) return -1;
        const dprt_mesh_desc& m = meshes[inst[i].mesh];
        if (m.ntris < 0) return -1;
        total += m.ntris;
    }
    return total;
}
static int flat_fill(const dprt_mesh_desc* meshes, int nm, const dprt_instance_desc* inst, int64_t ni, std::vector<float>& verts,
                     std::vector<float>& normals, std::vector<float>& uvs, std::vector<int32_t>& mats, bool& anyUv) {
    const int64_t total = flat_count(meshes, nm, inst, ni);
    if (total <= 0) return -1;
    verts.resize(9 * (size_t)total); normals.resize(9 * (size_t)total); uvs.assign(6 * (size_t)total, 0.0f); mats.resize((size_t)total);
    anyUv = false;
    size_t p = 0;
    for (int64_t i = 0; i < ni; i++) {
        const dprt_mesh_desc& m = meshes[inst[i].mesh];
        const float* X = inst[i].objectToWorld;
        double A[3][3], C[3][3];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) A[r][c] = (double)X[4 * r + c];
        C[0][0] = A[1][1] * A[2][2] - A[1][2] * A[2][1];
        C[0][1] = -(A[1][0] * A[2][2] - A[1][2] * A[2][0]);
        C[0][2] = A[1][0] * A[2][1] - A[1][1] * A[2][0];
        C[1][0] = -(A[0][1] * A[2][2] - A[0][2] * A[2][1]);
        C[1][1] = A[0][0] * A[2][2] - A[0][2] * A[2][0];
        C[1][2] = -(A[0][0] * A[2][1] - A[0][1] * A[2][0]);
        C[2][0] = A[0][1] * A[1][2] - A[0][2] * A[1][1];
        C[2][1] = -(A[0][0] * A[1][2] - A[0][2] * A[1][0]);
        C[2][2] = A[0][0] * A[1][1] - A[0][1] * A[1][0];
        const double det = A[0][0] * C[0][0] + A[0][1] * C[0][1] + A[0][2] * C[0][2];
        if (det == 0.0 || det != det || std::isinf(det)) return -1;
        const double rdet = 1.0 / det;
        float G[3][3];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) G[r][c] = (float)(C[r][c] * rdet);
        if (m.ntris > 0 && m.texCoords) anyUv = true;
        for (int64_t t = 0; t < m.ntris; t++, p++) {
            for (int k = 0; k < 3; k++) {
                const int vi = m.indices[3 * t + k], nidx = m.normalIndices[3 * t + k];
                if (vi < 0 || vi >= m.nPositions || nidx < 0 || nidx >= m.nNormals) return -1;
                const V3 q = xform_point(X, v3(m.positions[3 * (size_t)vi], m.positions[3 * (size_t)vi + 1], m.positions[3 * (size_t)vi + 2]));
                verts[9 * p + 3 * k] = q.x; verts[9 * p + 3 * k + 1] = q.y; verts[9 * p + 3 * k + 2] = q.z;
                const float nx = m.normals[3 * (size_t)nidx], ny = m.normals[3 * (size_t)nidx + 1], nz = m.normals[3 * (size_t)nidx + 2];
                for (int r = 0; r < 3; r++) normals[9 * p + 3 * k + r] = fmaf(G[r][2], nz, fmaf(G[r][1], ny, G[r][0] * nx));
                if (m.texCoords) {
                    const int ti = m.texCoordIndices[3 * t + k];
                    if (ti < 0 || ti >= m.nTexCoords) return -1;
                    uvs[6 * p + 2 * k] = m.texCoords[2 * (size_t)ti]; uvs[6 * p + 2 * k + 1] = m.texCoords[2 * (size_t)ti + 1];
                }
            }
            mats[p] = m.materialID;
        }
    }
    return 0;
}
// the flat streams themselves (tests compare them with dprt_flatten_instances bit for bit); buffers sized by orc_flatten_count
int64_t orc_flatten_count(const dprt_mesh_desc* meshes, int nm, const dprt_instance_desc* inst, int64_t ni) { return flat_count(meshes, nm, inst, ni); }
int orc_flatten(const dprt_mesh_desc* meshes, int nm, const dprt_instance_desc* inst, int64_t ni, float* verts9, float* normals9, float* uv6,
                int32_t* mats, int* hasUv) {
    std::vector<float> v, n, u; std::vector<int32_t> m; bool any = false;
    if (flat_fill(meshes, nm, inst, ni, v, n, u, m, any)) return -1;
    std::memcpy(verts9, v.data(), v.size() * 4); std::memcpy(normals9, n.data(), n.size() * 4);
    if (uv6) std::memcpy(uv6, u.data(), u.size() * 4);
    if (mats) std::memcpy(mats, m.data(), m.size() * 4);
    if (hasUv) *hasUv = any ? 1 : 0;
    return 0;
}
// dprt_upload_instanced_chunk
int orc_world_add_instanced_object(void* wp, int si, const dprt_object_desc* desc, const dprt_mesh_desc* meshes, int nm,
                                   const dprt_instance_desc* inst, int64_t ni) {
    std::vector<float> v, n, u; std::vector<int32_t> m; bool any = false;
    if (flat_fill(meshes, nm, inst, ni, v, n, u, m, any)) return -1;
    return orc_world_add_object_uv(wp, si, desc, v.data(), n.data(), any ? u.data() : nullptr, m.data(), (int64_t)m.size());
}
// dprt_set_texture / dprt_set_material_textures / dprt_set_env_map
int orc_world_set_texture(void* wp, int ti, const float* rgba, int width, int height) {
    World* w = (World*)wp;
    if (!w || ti < 0 || ti >= DPRT_MAX_TEXTURES) return -1;
    if ((int)w->textures.size() <= ti) w->textures.resize(ti + 1);
    Texture& T = w->textures[ti];
    if (!rgba) { T = Texture(); return 0; }
    if (width < 1 || height < 1) return -1;
    T.w = width; T.h = height; T.texels.assign(rgba, rgba + 4 * (size_t)width * height);
    return 0;
}
int orc_world_set_material_textures(void* wp, const int32_t* tex, int n) {
    World* w = (World*)wp;
    if (!w || !tex || n < 1) return -1;
    w->matTex.assign(tex, tex + n);
    return 0;
}
int orc_world_set_env_map(void* wp, const float* rgba, int width, int height, float rotation) {
    World* w = (World*)wp;
    if (!w) return -1;
    w->envMap = Texture(); w->envRotation = rotation;
    if (rgba) {
        if (width < 1 || height < 1) return -1;
        w->envMap.w = width; w->envMap.h = height; w->envMap.texels.assign(rgba, rgba + 4 * (size_t)width * height);
    }
    return 0;
}
// the two look-ups on their own (tests: against dprt_spec_texture_sample / dprt_spec_env_lookup and a numpy model)
int orc_texture_sample(const float* rgba, int width, int height, const float* u, const float* v, int64_t n, int clampV, float* out4) {
    if (!rgba || width < 1 || height < 1) return -1;
    Texture T; T.w = width; T.h = height; T.texels.assign(rgba, rgba + 4 * (size_t)width * height);
    for (int64_t i = 0; i < n; i++) T.sample(u[i], v[i], clampV != 0, out4 + 4 * i);
    return 0;
}
int orc_env_lookup(const float* rgba, int width, int height, float rotation, const float* dirs3, int64_t n, float* out3) {
    if (!rgba || width < 1 || height < 1) return -1;
    World w; w.envRotation = rotation;
    w.envMap.w = width; w.envMap.h = height; w.envMap.texels.assign(rgba, rgba + 4 * (size_t)width * height);
    for (int64_t i = 0; i < n; i++) {
        const V3 e = w.env_radiance(v3(dirs3[3 * i], dirs3[3 * i + 1], dirs3[3 * i + 2]));
        out3[3 * i] = e.x; out3[3 * i + 1] = e.y; out3[3 * i + 2] = e.z;
    }
    return 0;
}

int orc_world_set_model(void* wp, int si, int kind, const void* blob, size_t bytes) {
    World* w = (World*)wp;
    if (!w || si < 0 || si >= (int)w->objects.size()) return -1;
    Object& ob = w->objects[si];
    if (kind == 0) { ob.hasVis = ob.vis.load(blob, bytes); return ob.hasVis ? 0 : -1; }
    ob.hasDepth = ob.depth.load(blob, bytes); return ob.hasDepth ? 0 : -1;
}
// a copy of the product's BVH8 blob of scene object si (dprt_bvh8_copy): only the counting walker reads it
int orc_world_set_bvh8(void* wp, int si, const dprt_bvh8_node* nodes, int64_t nnodes, const dprt_bvh8_tri* tris, int64_t ntris) {
    World* w = (World*)wp;
    if (!w || si < 0 || si >= (int)w->objects.size() || !nodes || !tris || nnodes < 1 || ntris < 0) return -1;
    w->objects[si].nodes8.assign(nodes, nodes + nnodes); w->objects[si].tris8.assign(tris, tris + ntris);
    return 0;
}
int orc_count_bvh8(void* wp, int enable) { World* w = (World*)wp; if (!w) return -1; w->countBvh8 = enable != 0; return 0; }
// out: DPRT_STAGE_COUNT x {nodes, tris, rays} of rank `rank` since the last reset
int orc_get_bvh8_counters(void* wp, int rank, int64_t* out, int reset) {
    Rank* r = get_rank((World*)wp, rank); if (!r || !out) return -1;
    for (int s = 0; s < DPRT_STAGE_COUNT; s++) {
        out[3 * s] = r->cntNodes[s]; out[3 * s + 1] = r->cntTris[s]; out[3 * s + 2] = r->cntRays[s];
        if (reset) r->cntNodes[s] = r->cntTris[s] = r->cntRays[s] = 0;
    }
    return 0;
}
int orc_world_set_materials(void* wp, const dprt_material* m, int n) { World* w = (World*)wp; w->materials.assign(m, m + n); return 0; }
int orc_world_set_lights(void* wp, const dprt_light_tri* l, int n) { World* w = (World*)wp; w->lights.assign(l, l + n); return 0; }
int orc_world_set_camera(void* wp, const dprt_camera* c) { World* w = (World*)wp; w->cam = *c; return 0; }
int orc_enable_hit_prim(void* wp, int enable) {
    World* w = (World*)wp;
    for (Rank& r : w->ranks) { if (enable) r.hitPrim.assign((size_t)(1 + w->cfg.shadowPathCount) * w->N, -1); else r.hitPrim.clear(); }
    return 0;
}

int orc_reset_frame(void* wp) {
    World* w = (World*)wp;
    for (Rank& r : w->ranks) {
        std::fill(r.direct.begin(), r.direct.end(), 0.f); std::fill(r.env.begin(), r.env.end(), 0.f);
        std::fill(r.contribution.begin(), r.contribution.end(), 0.f); std::fill(r.occlusion.begin(), r.occlusion.end(), 0.f);
    }
    return 0;
}
int orc_begin_sample(void* wp, int sample) { begin_sample(*(World*)wp, sample); return 0; }
int orc_path_gen(void* wp, int rank) { World* w = (World*)wp; Rank* r = get_rank(w, rank); if (!r) return -1; path_gen(*w, *r); return 0; }
int orc_traverse(void* wp, int rank) { World* w = (World*)wp; Rank* r = get_rank(w, rank); if (!r) return -1; traverse(*w, *r); return 0; }
int orc_partition(void* wp, int rank) { World* w = (World*)wp; Rank* r = get_rank(w, rank); if (!r) return -1; partition(*w, *r); return 0; }
int orc_exchange(void* wp, int* done) { World* w = (World*)wp; bool d = exchange(*w); if (done) *done = d ? 1 : 0; return 0; }
int orc_shade(void* wp, int rank) { World* w = (World*)wp; Rank* r = get_rank(w, rank); if (!r) return -1; shade(*w, *r); return 0; }
int orc_reset_nn(void* wp, int rank) { World* w = (World*)wp; Rank* r = get_rank(w, rank); if (!r) return -1; reset_nn(*w, *r); return 0; }
int orc_shadow_trace(void* wp, int rank) { World* w = (World*)wp; Rank* r = get_rank(w, rank); if (!r) return -1; shadow_trace(*w, *r); return 0; }
int orc_secondary_trace(void* wp, int rank) { World* w = (World*)wp; Rank* r = get_rank(w, rank); if (!r) return -1; secondary_trace(*w, *r); return 0; }
int orc_bucket_queries(void* wp, int rank, int which, int insideOnly, int* total) {
    World* w = (World*)wp; Rank* r = get_rank(w, rank); if (!r) return -1;
    int t = bucket_queries(*w, *r, which, insideOnly); if (total) *total = t; return 0;
}
int orc_proxy_infer(void* wp, int rank, int kind, int predOffset) { World* w = (World*)wp; Rank* r = get_rank(w, rank); if (!r) return -1; proxy_infer(*w, *r, kind, predOffset); return 0; }
int orc_frame_buffer_update(void* wp, int rank) { World* w = (World*)wp; Rank* r = get_rank(w, rank); if (!r) return -1; frame_buffer_update(*w, *r); return 0; }
int orc_depth_buffer_update(void* wp, int rank) { World* w = (World*)wp; Rank* r = get_rank(w, rank); if (!r) return -1; depth_buffer_update(*w, *r); return 0; }
int orc_target_node_update(void* wp, int rank) { World* w = (World*)wp; Rank* r = get_rank(w, rank); if (!r) return -1; target_node_update(*w, *r); return 0; }
int orc_render_sample(void* wp, int sample) { render_sample(*(World*)wp, sample); return 0; }

// (direct+env)/spp summed over ranks in rank order: renderer.cpp:2031-2052
int orc_image(void* wp, float* out) {
    World* w = (World*)wp;
    const size_t n3 = (size_t)w->N * 3;
    std::fill(out, out + n3, 0.f);
    for (Rank& r : w->ranks)
        for (size_t i = 0; i < n3; i++) out[i] += (r.direct[i] + r.env[i]) / (float)w->cfg.spp;
    return 0;
}

int orc_get_path_size(void* wp, int rank, int* ps, int* sps) {
    Rank* r = get_rank((World*)wp, rank); if (!r) return -1;
    if (ps) *ps = r->pathSize;
    if (sps) *sps = r->shadowPathSize;
    return 0;
}
int orc_set_path_size(void* wp, int rank, int ps) { Rank* r = get_rank((World*)wp, rank); if (!r) return -1; r->pathSize = ps; return 0; }
int orc_set_query_total(void* wp, int rank, int t) { Rank* r = get_rank((World*)wp, rank); if (!r) return -1; r->queryTotal = t; return 0; }
int orc_get_stats(void* wp, int rank, dprt_stats* out) { Rank* r = get_rank((World*)wp, rank); if (!r) return -1; *out = r->stats; return 0; }

static void* buffer_ptr(Rank& r, int id, size_t* bytes) {
    switch (id) {
        case DPRT_BUF_PATHS: *bytes = r.paths.size() * sizeof(dprt_path_record); return r.paths.data();
        case DPRT_BUF_TRANSFER: *bytes = r.transfer.size() * sizeof(dprt_path_record); return r.transfer.data();
        case DPRT_BUF_TRANSFER_OFFSET: *bytes = r.transferOffset.size() * 4; return r.transferOffset.data();
        case DPRT_BUF_DIRECT: *bytes = r.direct.size() * 4; return r.direct.data();
        case DPRT_BUF_ENV: *bytes = r.env.size() * 4; return r.env.data();
        case DPRT_BUF_NN_INPUT: *bytes = r.nnInput.size() * 2; return r.nnInput.data();
        case DPRT_BUF_NN_QUERY: *bytes = r.nnQuery.size() * sizeof(dprt_nn_query); return r.nnQuery.data();
        case DPRT_BUF_NN_PACKED_INPUT: *bytes = r.nnPackedInput.size() * 2; return r.nnPackedInput.data();
        case DPRT_BUF_NN_PACKED_QUERY: *bytes = r.nnPackedQuery.size() * sizeof(dprt_nn_query); return r.nnPackedQuery.data();
        case DPRT_BUF_SCENE_OFFSET: *bytes = r.sceneOffset.size() * 4; return r.sceneOffset.data();
        case DPRT_BUF_PRED: *bytes = r.pred.size() * 2; return r.pred.data();
        case DPRT_BUF_OCCLUSION: *bytes = r.occlusion.size() * 4; return r.occlusion.data();
        case DPRT_BUF_CONTRIBUTION: *bytes = r.contribution.size() * 4; return r.contribution.data();
        case DPRT_BUF_HIT_PRIM: *bytes = r.hitPrim.size() * 4; return r.hitPrim.data();
    }
    *bytes = 0; return nullptr;
}
int orc_download(void* wp, int rank, int id, size_t off, void* host, size_t bytes) {
    Rank* r = get_rank((World*)wp, rank); if (!r) return -1;
    size_t total; char* p = (char*)buffer_ptr(*r, id, &total);
    if (!p || off + bytes > total) return -1;
    std::memcpy(host, p + off, bytes); return 0;
}
int orc_upload(void* wp, int rank, int id, size_t off, const void* host, size_t bytes) {
    Rank* r = get_rank((World*)wp, rank); if (!r) return -1;
    size_t total; char* p = (char*)buffer_ptr(*r, id, &total);
    if (!p || off + bytes > total) return -1;
    std::memcpy(p + off, host, bytes); return 0;
}

// closest hit of n rays against rank's local objects. mode 0 = oracle BVH, 1 = brute force.
int orc_trace_closest(void* wp, int rank, const dprt_ray* rays, int64_t n, dprt_hit* hits, int mode) {
    World* w = (World*)wp;
    if (!get_rank(w, rank)) return -1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
        const dprt_ray& ry = rays[i];
        const V3 o = v3(ry.origin[0], ry.origin[1], ry.origin[2]), d = v3(ry.direction[0], ry.direction[1], ry.direction[2]);
        float tMax = ry.tMax; int prim = -1;
        for (int k = 0; k < (int)w->objects.size(); k++) {
            const Object& ob = w->objects[k];
            if (!ob.present || w->is_proxy(rank, k)) continue;
            Hit h;
            const bool hit = mode == 1 ? ob.mesh.trace_brute(o, d, ry.tMin, tMax, h) : ob.mesh.trace(o, d, ry.tMin, tMax, false, h);
            if (hit) { tMax = h.t; prim = h.prim; }
        }
        hits[i].t = tMax; hits[i].primID = prim;
    }
    return 0;
}

// Vis pipeline (optix/vis_ray_kernel.cu:98-161): training samples of scene object si -- closest hit against that object
// only, features ((o - aabbMin)/(aabbMax - aabbMin), phi/2pi, theta/pi) in object space, label t/maxLength or 1.0 (miss).
int orc_gen_train_data(void* wp, int si, const dprt_ray* rays, int64_t n, float* feat, float* label) {
    World* w = (World*)wp;
    if (!w || si < 0 || si >= (int)w->objects.size() || !w->objects[si].present) return -1;
    const Object& ob = w->objects[si];
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
        const dprt_ray& ry = rays[i];
        const V3 o = v3(ry.origin[0], ry.origin[1], ry.origin[2]), d = v3(ry.direction[0], ry.direction[1], ry.direction[2]);
        Hit h;
        const bool hit = ob.mesh.trace(o, d, ry.tMin, ry.tMax, false, h);
        const V3 ol = xform_point(ob.desc.worldToObject, o), dl = xform_vector(ob.desc.worldToObject, d);
        float phi, theta;
        cartesian_to_spherical(normalized(dl), &phi, &theta);
        float* f = feat + 5 * i;
        f[0] = (ol.x - ob.desc.aabbMin[0]) / (ob.desc.aabbMax[0] - ob.desc.aabbMin[0]);
        f[1] = (ol.y - ob.desc.aabbMin[1]) / (ob.desc.aabbMax[1] - ob.desc.aabbMin[1]);
        f[2] = (ol.z - ob.desc.aabbMin[2]) / (ob.desc.aabbMax[2] - ob.desc.aabbMin[2]);
        f[3] = phi / 6.28318530717958647692f;
        f[4] = theta / 3.14159265358979323846f;
        label[i] = hit ? h.t / ob.desc.maxLength : 1.0f;
    }
    return 0;
}

// Precom pipeline (optix/precom_ray_kernel.cu:193-299): features at the proxy-AABB hit of each ray, label = depth of the
// original geometry behind the AABB surface / maxLength (1.0: AABB hit, geometry missed; valid 0: AABB missed).
int orc_gen_precom_data(void* wp, int si, const dprt_ray* rays, int64_t n, float* feat, float* label, uint8_t* valid) {
    World* w = (World*)wp;
    if (!w || si < 0 || si >= (int)w->objects.size() || !w->objects[si].present) return -1;
    const Object& ob = w->objects[si];
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
        const dprt_ray& ry = rays[i];
        const V3 o = v3(ry.origin[0], ry.origin[1], ry.origin[2]), d = v3(ry.direction[0], ry.direction[1], ry.direction[2]);
        const V3 ol = xform_point(ob.desc.worldToObject, o), dl = xform_vector(ob.desc.worldToObject, d);
        float* f = feat + 5 * i;
        float ta; bool inside;
        const bool aabb = aabb_hit(ol, dl, ob.desc.aabbMin, ob.desc.aabbMax, ry.tMin, ry.tMax, &ta, &inside);
        if (aabb) {
            const V3 pl = xform_point(ob.desc.worldToObject, at(o, d, ta));
            const V3 dirL = inside ? neg(dl) : dl;
            float phi, theta;
            cartesian_to_spherical(normalized(dirL), &phi, &theta);
            f[0] = (pl.x - ob.desc.aabbMin[0]) / (ob.desc.aabbMax[0] - ob.desc.aabbMin[0]);
            f[1] = (pl.y - ob.desc.aabbMin[1]) / (ob.desc.aabbMax[1] - ob.desc.aabbMin[1]);
            f[2] = (pl.z - ob.desc.aabbMin[2]) / (ob.desc.aabbMax[2] - ob.desc.aabbMin[2]);
            f[3] = phi / 6.28318530717958647692f;
            f[4] = theta / 3.14159265358979323846f;
        } else {
            f[0] = f[1] = f[2] = f[3] = f[4] = 0.0f;
        }
        Hit h;
        const bool geo = ob.mesh.trace(o, d, ry.tMin, FLT_MAX, false, h);
        label[i] = aabb ? (geo ? (h.t - ta) / ob.desc.maxLength : 1.0f) : 1.0f;
        valid[i] = aabb ? 1 : 0;
    }
    return 0;
}

// walk the product's BVH8 blob: hits + per-ray counters (nodes fetched, triangles tested)
int orc_bvh8_trace(const dprt_bvh8_node* nodes, const dprt_bvh8_tri* tris, const dprt_ray* rays, int64_t n, dprt_hit* hits,
                   int64_t* nodes_visited, int64_t* tris_tested) {
    int64_t tn = 0, tt = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : tn, tt)
    for (int64_t i = 0; i < n; i++) {
        const dprt_ray& ry = rays[i];
        Hit h; Bvh8Count c{0, 0};
        bvh8_walk(nodes, tris, v3(ry.origin[0], ry.origin[1], ry.origin[2]), v3(ry.direction[0], ry.direction[1], ry.direction[2]),
                  ry.tMin, ry.tMax, h, c);
        if (hits) { hits[i].t = h.t; hits[i].primID = h.prim; }
        tn += c.nodes; tt += c.tris;
    }
    if (nodes_visited) *nodes_visited = tn;
    if (tris_tested) *tris_tested = tt;
    return 0;
}

// standalone fp32 MLP forward: x [n,5] fp16 -> y_f32 [n] (unrounded) and y_f16 [n] (may be NULL)
int orc_mlp_forward(const void* blob, size_t bytes, const uint16_t* x, int64_t n, float* y_f32, uint16_t* y_f16) {
    Mlp m;
    if (!m.load(blob, bytes)) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        float xf[5];
        for (int k = 0; k < 5; k++) xf[k] = h2f(x[i * 5 + k]);
        const float y = m.forward_row(xf);
        if (y_f32) y_f32[i] = y;
        if (y_f16) y_f16[i] = f2h(y);
    }
    return 0;
}

// all host threads for the timed CPU arms, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
int orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
    return 0;
}

int orc_num_threads(void) {
    int n = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
#pragma omp single
        n = omp_get_num_threads();
    }
#endif
    return n;
}

}  // extern "C"
