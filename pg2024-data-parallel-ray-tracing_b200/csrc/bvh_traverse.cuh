// bvh_traverse.cuh -- per-thread traversal of the 80-byte compressed 8-wide BVH (device).
//
// This is what replaces optixTrace() (call sites: optix/kernel.cu:394, distributed_traversal_kernel.cu:245,
// shadow_ray_kernel.cu:177, secondary_ray_kernel.cu:200). Closest-hit results are independent of the BVH:
// every candidate triangle goes through the watertight test of dprt_math.cuh and ties in t resolve to the
// lower primitive id, so any conservative traversal order yields the same (t, primID).
#pragma once
#include "dprt_math.cuh"

namespace dprt {

struct TraceHit {
    float t;        // closest accepted t (tmax when nothing was hit)
    int   prim;     // original primitive id, -1 = none
    int   tri;      // index into the leaf-ordered triangle array
    float alpha, beta;
};

DPRT_D float q2f(uint32_t w, uint32_t sel) {
    // byte `sel & 3` of w -> float, via the 2^23 mantissa trick (exact for 0..255)
    return __uint_as_float(__byte_perm(w, 0x4B000000u, sel)) - 8388608.0f;
}

#define DPRT_STACK 40

// instrumentation (dprt_enable_counters): BVH8 nodes fetched and triangles tested by one thread
struct TraceCount { uint32_t nodes, tris; };

// ANY = true: return as soon as one triangle is hit in (tmin, tmax) (shadow rays).
template <bool ANY, bool COUNT>
DPRT_D bool bvh8_trace(const uint4* __restrict__ nodes, const float4* __restrict__ tris,
                       V3 o, V3 d, float tmin, float tmax, TraceHit& hit, TraceCount& cnt) {
    const RayShear rs = ray_shear(d);
    const float dxs = fabsf(d.x) > 1e-20f ? d.x : copysignf(1e-20f, d.x);
    const float dys = fabsf(d.y) > 1e-20f ? d.y : copysignf(1e-20f, d.y);
    const float dzs = fabsf(d.z) > 1e-20f ? d.z : copysignf(1e-20f, d.z);
    const float idx = 1.0f / dxs, idy = 1.0f / dys, idz = 1.0f / dzs;
    const bool nx = dxs < 0.0f, ny = dys < 0.0f, nz = dzs < 0.0f;
    const uint32_t octinv = 7u - ((nx ? 1u : 0u) | (ny ? 2u : 0u) | (nz ? 4u : 0u));

    float tbest = tmax; int bestPrim = 0x7fffffff; int bestTri = -1; float ba = 0.f, bb = 0.f;

    uint2 stack[DPRT_STACK];
    int sp = 0;
    uint2 ng = make_uint2(0u, 0x80000000u);

    for (;;) {
        uint2 tg = make_uint2(0u, 0u);
        if (ng.y & 0xff000000u) {
            const uint32_t bit = 31u - __clz(ng.y);
            const uint32_t slot = (bit - 24u) ^ octinv;
            const uint32_t rel = __popc(ng.y & 0xffu & ((1u << slot) - 1u));
            ng.y &= ~(1u << bit);
            const uint32_t ni = ng.x + rel;
            if (ng.y & 0xff000000u) { if (sp < DPRT_STACK) stack[sp++] = ng; }

            if (COUNT) cnt.nodes++;
            const uint4 n0 = __ldg(nodes + 5 * (size_t)ni + 0);
            const uint4 n1 = __ldg(nodes + 5 * (size_t)ni + 1);
            const uint4 n2 = __ldg(nodes + 5 * (size_t)ni + 2);
            const uint4 n3 = __ldg(nodes + 5 * (size_t)ni + 3);
            const uint4 n4 = __ldg(nodes + 5 * (size_t)ni + 4);

            const float adjx = __uint_as_float((n0.w & 0xffu) << 23) * idx;
            const float adjy = __uint_as_float(((n0.w >> 8) & 0xffu) << 23) * idy;
            const float adjz = __uint_as_float(((n0.w >> 16) & 0xffu) << 23) * idz;
            const float orgx = (__uint_as_float(n0.x) - o.x) * idx;
            const float orgy = (__uint_as_float(n0.y) - o.y) * idy;
            const float orgz = (__uint_as_float(n0.z) - o.z) * idz;

            uint32_t hitmask = 0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t meta4 = h ? n1.w : n1.z;
                const uint32_t lox = h ? (nx ? n3.w : n2.y) : (nx ? n3.z : n2.x);
                const uint32_t hix = h ? (nx ? n2.y : n3.w) : (nx ? n2.x : n3.z);
                const uint32_t loy = h ? (ny ? n4.y : n2.w) : (ny ? n4.x : n2.z);
                const uint32_t hiy = h ? (ny ? n2.w : n4.y) : (ny ? n2.z : n4.x);
                const uint32_t loz = h ? (nz ? n4.w : n3.y) : (nz ? n4.z : n3.x);
                const uint32_t hiz = h ? (nz ? n3.y : n4.w) : (nz ? n3.x : n4.z);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t sel = 0x7650u | (uint32_t)j;
                    const float tnx = fmaf(q2f(lox, sel), adjx, orgx);
                    const float tny = fmaf(q2f(loy, sel), adjy, orgy);
                    const float tnz = fmaf(q2f(loz, sel), adjz, orgz);
                    const float tfx = fmaf(q2f(hix, sel), adjx, orgx);
                    const float tfy = fmaf(q2f(hiy, sel), adjy, orgy);
                    const float tfz = fmaf(q2f(hiz, sel), adjz, orgz);
                    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
                    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tbest));
                    if (tn <= tf) {
                        const uint32_t meta = (meta4 >> (8 * j)) & 0xffu;
                        const uint32_t inner = ((meta & 0x18u) == 0x18u) ? octinv : 0u;
                        const uint32_t bidx = (meta ^ inner) & 31u;
                        hitmask |= (meta >> 5) << bidx;
                    }
                }
            }
            ng = make_uint2(n1.x, (hitmask & 0xff000000u) | (n0.w >> 24));
            tg = make_uint2(n1.y, hitmask & 0x00ffffffu);
        }

        while (tg.y) {
            const uint32_t k = __ffs(tg.y) - 1u;
            tg.y &= tg.y - 1u;
            const uint32_t ti = tg.x + k;
            if (COUNT) cnt.tris++;
            const float4 a = __ldg(tris + 3 * (size_t)ti + 0);
            const float4 b = __ldg(tris + 3 * (size_t)ti + 1);
            const float4 c = __ldg(tris + 3 * (size_t)ti + 2);
            float t, al, be;
            if (tri_intersect(rs, o, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), tmin, tmax, &t, &al, &be)) {
                if (ANY) { hit.t = t; hit.prim = __float_as_int(a.w); hit.tri = (int)ti; hit.alpha = al; hit.beta = be; return true; }
                const int prim = __float_as_int(a.w);
                if (t < tbest || (t == tbest && prim < bestPrim)) {
                    tbest = t; bestPrim = prim; bestTri = (int)ti; ba = al; bb = be;
                }
            }
        }

        if ((ng.y & 0xff000000u) == 0u) {
            if (sp == 0) break;
            ng = stack[--sp];
        }
    }
    hit.t = tbest; hit.prim = bestTri >= 0 ? bestPrim : -1; hit.tri = bestTri; hit.alpha = ba; hit.beta = bb;
    return bestTri >= 0;
}

}  // namespace dprt
