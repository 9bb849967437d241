// forge3d_b200/csrc/f3d_wavefront.cuh
// Wavefront multi-bounce path tracer for sm_100a: what the reference's raygen / intersect / shade /
// shadow / scatter dispatches (src/path_tracing/wavefront/render.rs:87-209) do to a path, restructured for the device:
//
//   * one kernel per bounce instead of four stages: a thread owns a path for the whole bounce -- closest hit, the NEE samples with
//     their any-hit shadow rays traced in place, the continuation sample, Russian roulette -- so hit / shadow / scatter records never
//     travel through HBM; only the 48-byte path state does, once per bounce;
//   * survivors are stream-compacted into the next bounce's queue with one warp-aggregated atomic per warp (ballot + popc), the stage the
//     reference removed (wavefront/dispatch.rs:130-138); queue lengths stay on the device (the reference reads the 16-byte header back
//     and stalls every bounce, queues/types.rs:166-213);
//   * the wide bounces (depth < kWfWideDepth) run as compacted waves; whatever is still alive then (a few percent of the paths) is
//     finished by one tail kernel in which each thread walks its path to the end, so a batch of frames is kWfWideDepth + 2 launches, not 64 per frame;
//   * consecutive frames are traced together in batches (a 512 x 512 frame is 262 k paths at depth 0 and 24 k at depth 3: one frame
//     cannot fill 148 SMs); each frame sums its own radiance per pixel in push order (emissive, environment, directional, area, miss) and
//     the batch adds those sums to the accumulator in frame order, so the image is deterministic and independent of the batch size;
//     the reference's non-atomic adds from different threads of one dispatch are neither.
//
// Arithmetic follows pt_raygen.wgsl:74-224, pt_intersect.wgsl:84-216,374-558, pt_shade.wgsl:44-196,342-470,478-862,
// pt_shadow.wgsl:1-58,161-294 and pt_scatter.wgsl:77-132 under the numerics contract of DESIGN.md section 4 (no FMA contraction,
// pinned sin/cos/exp2/log2, pow by small integers written as products).  Not implemented: hair segments, ReSTIR reservoirs, the fog
// medium, anisotropic GGX, the debug AOV preview.
#pragma once
#include "f3d_aether.cuh"   // exp2_pinned
#include "f3d_math.cuh"

namespace f3d {

#ifndef F3D_WF_MIN_CTAS
#define F3D_WF_MIN_CTAS 2   // resident CTAs per SM the bounce kernels are compiled for (3 costs ~300 B of spills per thread: A/B on the GPU)
#endif
constexpr int kWfThreads = 256;
constexpr uint32_t kWfWideDepth = 4;    // default: bounces 0..3 are compacted waves, the tail kernel takes over at depth 4 (F3D_B200_WF_WIDE_DEPTH)
constexpr uint32_t kWfMaxDepth = 16;    // (h.depth + 1) < 16, pt_shade.wgsl:831; MAX_DEPTH * 2 iterations, render.rs:115

struct WfMesh {   // passed by value to the (non-inlined) traversal so that the kernel parameters are never copied to local memory
    const float* xyz; const uint32_t* idx; uint32_t ntris;
    const float4* bvh_nodes; const uint32_t* bvh_tris;   // nullptr: sweep the triangles in index order
};

struct WfParams {
    uint32_t w, h;
    uint32_t part_rank, part_world, part_rows, local_rows;   // image partition in interleaved blocks of part_rows rows (world 1: one block)
    float cam_origin[3], cam_forward[3], cam_right[3], cam_up[3];
    float fov_y_rad, aspect;
    const float* spheres; uint32_t nsph;        // 20 floats each
    const float* dirl; uint32_t ndir;           // 8 floats each
    const float* areal; uint32_t narea;         // 12 floats each
    const float* imp; uint32_t nimp;
    float env[16];
    WfMesh mesh;
    const float* inst; uint32_t ninst;          // 36 words each
    float4* accum;                              // running sum over frames, per pixel
    float4* fsum;                               // [frame in batch][pixel] radiance sum of one frame
    float4* qa[2]; float4* qb[2]; float4* qc[2];   // path queues, ping-pong: (o, pixel) (d, rng_hi) (throughput, frame in batch)
};

struct WfFrame {
    uint32_t frame, seed_hi, seed_lo;
    float u1, u2;                // sobol2(frame), pt_raygen.wgsl:122-153 (computed on the host: integer work)
};

// A batch of consecutive frames traced together: small images do not fill 148 SMs with one frame's paths, least of all at depth >= 2.
constexpr uint32_t kWfMaxBatch = 16;
struct WfBatch {
    uint32_t first_frame, nframes;
    uint32_t seed_hi[kWfMaxBatch], seed_lo[kWfMaxBatch];
    float u1[kWfMaxBatch], u2[kWfMaxBatch];
    uint32_t* qcount;            // this batch's kWfMaxDepth + 1 queue lengths
    __device__ __forceinline__ WfFrame frame(uint32_t fl) const {
        WfFrame F;
        F.frame = first_frame + fl; F.seed_hi = seed_hi[fl]; F.seed_lo = seed_lo[fl]; F.u1 = u1[fl]; F.u2 = u2[fl];
        return F;
    }
};

struct WfPath { v3 o, d, thr; uint32_t pixel, rng_hi; float tmin; };

__device__ __forceinline__ v3 ld3(const float* p) { return V3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }
__device__ __forceinline__ v3 ar3(const float* p) { return V3(p[0], p[1], p[2]); }
__device__ __forceinline__ float satf(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
__device__ __forceinline__ v3 mix3(v3 a, v3 b, float t) { return V3(mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)); }
__device__ __forceinline__ float vget(v3 a, uint32_t i) { return i == 0u ? a.x : (i == 1u ? a.y : a.z); }
__device__ __forceinline__ float pow2f(float x) { return x * x; }
__device__ __forceinline__ float pow5f(float x) { const float x2 = x * x; return (x2 * x2) * x; }
__device__ __forceinline__ float pow16f(float x) { float a = x * x; a = a * a; a = a * a; return a * a; }

// Pinned log2 for normal positive x (Cephes log2f kernel); x <= 0 or subnormal -> -inf, +inf -> +inf, NaN -> NaN.
__device__ __forceinline__ float log2_pinned(float x) {
    if (x != x) return x;
    if (!(x >= 1.17549435e-38f)) return __int_as_float(0xff800000);
    if (x == __int_as_float(0x7f800000)) return x;
    uint32_t b = __float_as_uint(x);
    int e = (int)(b >> 23) - 126;
    float m = __uint_as_float((b & 0x007FFFFFu) | 0x3F000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; } else { m = m - 1.0f; }
    const float z = m * m;
    float p = 7.0376836292e-2f;
    p = p * m + -1.1514610310e-1f;
    p = p * m + 1.1676998740e-1f;
    p = p * m + -1.2420140846e-1f;
    p = p * m + 1.4249322787e-1f;
    p = p * m + -1.6668057665e-1f;
    p = p * m + 2.0000714765e-1f;
    p = p * m + -2.4999993993e-1f;
    p = p * m + 3.3333331174e-1f;
    float y = (m * z) * p;
    y = y - 0.5f * z;
    float r = y * 0.44269504088896340736f;
    r = r + m * 0.44269504088896340736f;
    r = r + y;
    r = r + m;
    r = r + (float)e;
    return r;
}
__device__ __forceinline__ float pow_pinned(float x, float y) { return exp2_pinned(y * log2_pinned(x)); }

__device__ __forceinline__ float wf_cp_rotate(float u, float r) { const float x = u + r; return x - floorf(x); }   // pt_raygen.wgsl:155-159

struct WfBasis { v3 t, b, n; };   // columns of make_tangent_basis, pt_shade.wgsl:352-360
__device__ __forceinline__ WfBasis wf_basis(v3 n) {
    const float sign = n.z < 0.0f ? -1.0f : 1.0f;
    const float a = fdiv(-1.0f, sign + n.z);
    const float b = (n.x * n.y) * a;
    WfBasis r;
    r.t = V3(1.0f + ((sign * n.x) * n.x) * a, sign * b, -sign * n.x);
    r.b = V3(b, sign + (n.y * n.y) * a, -n.y);
    r.n = n;
    return r;
}
__device__ __forceinline__ v3 wf_to_world(const WfBasis& B, v3 v) { return (B.t * v.x + B.b * v.y) + B.n * v.z; }
__device__ __forceinline__ v3 wf_cosine_hemisphere(float u1, float u2) {   // :363-370
    const float r = fsqrt(u1), phi = (2.0f * 3.14159265358979323846f) * u2;
    float s, c;
    sincos_pinned(phi, s, c);
    return V3(r * c, r * s, fsqrt(fmaxf(0.0f, 1.0f - u1)));
}
__device__ __forceinline__ v3 wf_reflect(v3 i, v3 n) { return i - n * (2.0f * dot3(n, i)); }
__device__ __forceinline__ v3 wf_refract(v3 i, v3 n, float eta) {
    const float d = dot3(n, i);
    const float k = 1.0f - (eta * eta) * (1.0f - d * d);
    if (k < 0.0f) return V3(0, 0, 0);
    return i * eta - n * (eta * d + fsqrt(k));
}

// ray_sphere, pt_intersect.wgsl:374-386
__device__ __forceinline__ float wf_sphere_t(v3 ro, v3 rd, v3 c, float r) {
    const v3 oc = ro - c;
    const float b = dot3(oc, rd);
    const float cterm = dot3(oc, oc) - r * r;
    const float disc = b * b - cterm;
    if (disc <= 0.0f) return 1e30f;
    const float s = fsqrt(disc);
    const float t0 = -b - s, t1 = -b + s;
    if (t0 > 1e-3f) return t0;
    if (t1 > 1e-3f) return t1;
    return 1e30f;
}
// ray_sphere, pt_shadow.wgsl:161-174
__device__ __forceinline__ bool wf_sphere_any(v3 ro, v3 rd, v3 c, float r, float tmin, float tmax) {
    const v3 oc = ro - c;
    const float b = dot3(oc, rd);
    const float cterm = dot3(oc, oc) - r * r;
    const float disc = b * b - cterm;
    if (disc <= 0.0f) return false;
    const float s = fsqrt(disc);
    const float t0 = -b - s, t1 = -b + s;
    return (t0 > tmin && t0 < tmax) || (t1 > tmin && t1 < tmax);
}

struct WfShear { uint32_t kx, ky, kz; float Sx, Sy, Sz; };   // the per-ray part of the watertight test, pt_intersect.wgsl:112-122
__device__ __forceinline__ WfShear wf_shear(v3 d) {
    const float adx = fabsf(d.x), ady = fabsf(d.y), adz = fabsf(d.z);
    WfShear s;
    s.kz = 2u; s.kx = 0u; s.ky = 1u;
    if (adx > ady && adx > adz) { s.kz = 0u; s.kx = 1u; s.ky = 2u; }
    else if (ady > adz) { s.kz = 1u; s.kx = 2u; s.ky = 0u; }
    s.Sz = frcp(vget(d, s.kz));
    s.Sx = vget(d, s.kx) * s.Sz;
    s.Sy = vget(d, s.ky) * s.Sz;
    return s;
}
// ray_triangle_intersect (watertight), pt_intersect.wgsl:101-166; the closest hit keeps the lowest triangle index among equal t
__device__ __forceinline__ void wf_tri_closest(const WfMesh& P, v3 ro, const WfShear& sh, float tmin, float tmax, uint32_t tri, float& best,
                                               uint32_t& best_tri) {
    const v3 v0 = ld3(P.xyz + 3 * (size_t)__ldg(P.idx + 3 * tri)), v1 = ld3(P.xyz + 3 * (size_t)__ldg(P.idx + 3 * tri + 1)),
             v2 = ld3(P.xyz + 3 * (size_t)__ldg(P.idx + 3 * tri + 2));
    const v3 A = v0 - ro, B = v1 - ro, C = v2 - ro;
    const float Akz = vget(A, sh.kz), Bkz = vget(B, sh.kz), Ckz = vget(C, sh.kz);
    const float ax = vget(A, sh.kx) - sh.Sx * Akz, ay = vget(A, sh.ky) - sh.Sy * Akz;
    const float bx = vget(B, sh.kx) - sh.Sx * Bkz, by = vget(B, sh.ky) - sh.Sy * Bkz;
    const float cx = vget(C, sh.kx) - sh.Sx * Ckz, cy = vget(C, sh.ky) - sh.Sy * Ckz;
    const float az = Akz * sh.Sz, bz = Bkz * sh.Sz, cz = Ckz * sh.Sz;
    const float U = (bx * cy) - (by * cx);
    const float V = (cx * ay) - (cy * ax);
    const float W = (ax * by) - (ay * bx);
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return;
    const float det = (U + V) + W;
    if (det == 0.0f) return;
    const float T = (U * az + V * bz) + W * cz;
    const float t = fdiv(T, det);
    if (t > tmin && t < tmax && (t < best || (t == best && tri < best_tri))) { best = t; best_tri = tri; }
}
// Moeller-Trumbore any-hit, pt_shadow.wgsl:31-46
__device__ __forceinline__ bool wf_tri_any(const WfMesh& P, v3 ro, v3 rd, float tmin, float tmax, uint32_t tri) {
    const v3 v0 = ld3(P.xyz + 3 * (size_t)__ldg(P.idx + 3 * tri)), v1 = ld3(P.xyz + 3 * (size_t)__ldg(P.idx + 3 * tri + 1)),
             v2 = ld3(P.xyz + 3 * (size_t)__ldg(P.idx + 3 * tri + 2));
    const v3 e1 = v1 - v0, e2 = v2 - v0;
    const v3 h = cross3(rd, e2);
    const float a = dot3(e1, h);
    if (fabsf(a) < 1e-7f) return false;
    const float f = frcp(a);
    const v3 s = ro - v0;
    const float u = f * dot3(s, h);
    if (u < 0.0f || u > 1.0f) return false;
    const v3 q = cross3(s, e1);
    const float v = f * dot3(rd, q);
    if (v < 0.0f || u + v > 1.0f) return false;
    const float t = f * dot3(e2, q);
    return t > tmin && t < tmax;
}

// bvh_intersect_mesh(_desc), pt_intersect.wgsl:168-216,301-371.  The BVH (median split or LBVH, padded leaf boxes, f3d_backend.cu) prunes
// the index-order sweep; with the lowest-index tie rule the closest hit is the sweep's whatever the tree looks like.
template <bool ANY>
__device__ __noinline__ bool wf_mesh(const WfMesh P, v3 ro, v3 rd, float tmin, float tmax, float& t_out, v3& n_out) {
    float best = tmax;
    uint32_t best_tri = 0xFFFFFFFFu;
    WfShear sh;
    if (!ANY) sh = wf_shear(rd);
    if (P.bvh_nodes == nullptr) {
        for (uint32_t k = 0; k < P.ntris; k++) {
            if (ANY) { if (wf_tri_any(P, ro, rd, tmin, tmax, k)) return true; }
            else wf_tri_closest(P, ro, sh, tmin, tmax, k, best, best_tri);
        }
    } else {
        const float ix = frcp(rd.x), iy = frcp(rd.y), iz = frcp(rd.z);
        uint32_t stack[64];
        uint32_t sp = 0;
        stack[sp++] = 0u;
        while (sp) {
            const uint32_t ni = stack[--sp];
            const float4 n0 = __ldg(P.bvh_nodes + 2 * (size_t)ni), n1 = __ldg(P.bvh_nodes + 2 * (size_t)ni + 1);
            const float tx0 = (n0.x - ro.x) * ix, tx1 = (n1.x - ro.x) * ix;
            const float ty0 = (n0.y - ro.y) * iy, ty1 = (n1.y - ro.y) * iy;
            const float tz0 = (n0.z - ro.z) * iz, tz1 = (n1.z - ro.z) * iz;
            const float tnear = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), tmin));
            const float tfar = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), best));
            if (!(tnear <= tfar)) continue;
            const uint32_t a = __float_as_uint(n0.w), b = __float_as_uint(n1.w);
            if (a & 0x80000000u) {
                const uint32_t first = a & 0x7FFFFFFFu;
                for (uint32_t k = 0; k < b; k++) {
                    const uint32_t tri = __ldg(P.bvh_tris + first + k);
                    if (ANY) { if (wf_tri_any(P, ro, rd, tmin, tmax, tri)) return true; }
                    else wf_tri_closest(P, ro, sh, tmin, tmax, tri, best, best_tri);
                }
            } else if (sp + 2u <= 64u) {
                stack[sp++] = a;
                stack[sp++] = b;
            }
        }
    }
    if (ANY || best_tri == 0xFFFFFFFFu) return false;
    const v3 v0 = ld3(P.xyz + 3 * (size_t)__ldg(P.idx + 3 * best_tri)), v1 = ld3(P.xyz + 3 * (size_t)__ldg(P.idx + 3 * best_tri + 1)),
             v2 = ld3(P.xyz + 3 * (size_t)__ldg(P.idx + 3 * best_tri + 2));
    t_out = best;
    n_out = normalize3(cross3(v1 - v0, v2 - v0));
    return true;
}

// mat4x4 (column-major) * vec4(p, w), pt_intersect.wgsl:273-287
__device__ __forceinline__ v3 wf_xform(const float* m, v3 p, float w) {
    return V3(((__ldg(m + 0) * p.x + __ldg(m + 4) * p.y) + __ldg(m + 8) * p.z) + __ldg(m + 12) * w,
              ((__ldg(m + 1) * p.x + __ldg(m + 5) * p.y) + __ldg(m + 9) * p.z) + __ldg(m + 13) * w,
              ((__ldg(m + 2) * p.x + __ldg(m + 6) * p.y) + __ldg(m + 10) * p.z) + __ldg(m + 14) * w);
}
// transpose(world_to_object) * vec4(n, 0), pt_intersect.wgsl:289-294
__device__ __forceinline__ v3 wf_xform_normal(const float* m, v3 n) {
    return V3((__ldg(m + 0) * n.x + __ldg(m + 1) * n.y) + __ldg(m + 2) * n.z, (__ldg(m + 4) * n.x + __ldg(m + 5) * n.y) + __ldg(m + 6) * n.z,
              (__ldg(m + 8) * n.x + __ldg(m + 9) * n.y) + __ldg(m + 10) * n.z);
}

// pt_shadow.wgsl:248-293
__device__ __forceinline__ bool wf_occluded(const WfParams& P, v3 ro, v3 rd, float tmin, float tmax) {
    for (uint32_t i = 0; i < P.nsph; i++) {
        const float* s = P.spheres + 20 * (size_t)i;
        if (wf_sphere_any(ro, rd, ld3(s), __ldg(s + 3), tmin, tmax)) return true;
    }
    float t;
    v3 n;
    if (P.ninst == 0u) return P.mesh.ntris ? wf_mesh<true>(P.mesh, ro, rd, tmin, tmax, t, n) : false;
    for (uint32_t ii = 0; ii < P.ninst; ii++) {
        const float* w2o = P.inst + 36 * (size_t)ii + 16;
        if (__float_as_uint(__ldg(P.inst + 36 * (size_t)ii + 32)) != 0u) continue;   // only BLAS 0 exists: mesh_any_hit_desc returns false
        if (wf_mesh<true>(P.mesh, wf_xform(w2o, ro, 1.0f), normalize3(wf_xform(w2o, rd, 0.0f)), tmin, tmax, t, n)) return true;
    }
    return false;
}

struct WfBsdf { v3 f; float pdf; };
// bsdf_eval_pdf, pt_shade.wgsl:44-96 (isotropic branch)
__device__ __forceinline__ WfBsdf wf_bsdf(v3 wo, v3 wi, v3 n, v3 albedo, float metallic, float roughness) {
    const float PI_F = 3.14159265358979323846f;
    WfBsdf r;
    const float ndl = fmaxf(dot3(n, wi), 0.0f), ndv = fmaxf(dot3(n, wo), 0.0f);
    if (ndl <= 0.0f || ndv <= 0.0f) { r.f = V3(0, 0, 0); r.pdf = 0.0f; return r; }
    const float kd = satf(1.0f - metallic);
    const v3 fd = V3(fdiv(albedo.x, PI_F), fdiv(albedo.y, PI_F), fdiv(albedo.z, PI_F)) * kd;
    const float pdf_d = fdiv(ndl, PI_F);
    const float m = fmaxf(0.02f, roughness * roughness);
    const v3 h = normalize3(wi + wo);
    const float ndh = fmaxf(dot3(n, h), 0.0f), vdh = fmaxf(dot3(wo, h), 0.0f);
    const float a2 = m * m;
    const float D = fdiv(a2, fmaxf(PI_F * pow2f((ndh * ndh) * (a2 - 1.0f) + 1.0f), 1e-6f));
    const float k = fdiv(pow2f(m + 1.0f), 8.0f);
    const float G = fdiv(ndl, ndl * (1.0f - k) + k) * fdiv(ndv, ndv * (1.0f - k) + k);
    const float sm = satf(metallic);
    const v3 F0 = V3(mixf(0.04f, albedo.x, sm), mixf(0.04f, albedo.y, sm), mixf(0.04f, albedo.z, sm));
    const float fw = pow5f(1.0f - satf(vdh));
    const v3 F = V3(F0.x + (1.0f - F0.x) * fw, F0.y + (1.0f - F0.y) * fw, F0.z + (1.0f - F0.z) * fw);
    const float spec = fdiv(D * G, fmaxf((4.0f * ndl) * ndv, 1e-6f));
    const float pdf_s = fdiv(D * ndh, fmaxf(4.0f * vdh, 1e-6f));
    const float ks = 1.0f - kd;
    r.f = fd + F * spec;
    r.pdf = fmaxf(kd * pdf_d + ks * pdf_s, 1e-8f);
    return r;
}
__device__ __forceinline__ float wf_power_cosine_pdf_up(v3 w) {   // pt_shade.wgsl:157-161, m = 16
    const float c = fmaxf(dot3(V3(0, 1, 0), normalize3(w)), 0.0f);
    return fdiv(17.0f * pow16f(c), 2.0f * 3.14159265358979323846f);
}

// pt_raygen.wgsl:161-224, one sample per pixel per frame (spp = 1, adjudication.rs:196)
__device__ __forceinline__ WfPath wf_raygen(const WfParams& P, const WfFrame& F, uint32_t pix) {
    const uint32_t px = pix % P.w, py = pix / P.w;
    uint32_t rr = F.seed_lo ^ (px * 9781u) ^ (py * 6271u) ^ (F.seed_hi * 13007u);
    const float r1 = xorshift32(rr), r2 = xorshift32(rr);
    const float jx = tent_offset(wf_cp_rotate(F.u1, r1)) * 0.5f, jy = tent_offset(wf_cp_rotate(F.u2, r2)) * 0.5f;
    const float ndc_x = fdiv(((float)px + 0.5f) + jx, (float)P.w) * 2.0f - 1.0f;
    const float ndc_y = (1.0f - fdiv(((float)py + 0.5f) + jy, (float)P.h)) * 2.0f - 1.0f;
    float hs, hc;
    sincos_pinned(0.5f * P.fov_y_rad, hs, hc);
    const float half_h = fdiv(hs, hc), half_w = P.aspect * half_h;
    v3 rd = normalize3(V3(ndc_x * half_w, ndc_y * half_h, -1.0f));
    rd = normalize3((ar3(P.cam_right) * rd.x + ar3(P.cam_up) * rd.y) + (-ar3(P.cam_forward)) * rd.z);
    WfPath p;
    p.o = ar3(P.cam_origin); p.d = rd; p.thr = V3(1, 1, 1);
    p.pixel = pix; p.rng_hi = F.seed_hi ^ (pix * 9781u) ^ (F.frame * 6271u); p.tmin = 1e-4f;
    return p;
}

// One bounce of one path: pt_intersect -> pt_shade (+ its shadow rays through pt_shadow) -> pt_scatter.  `acc` is the pixel's
// accumulator, held in registers by the caller.  Returns true when the path continues (p is then the scattered ray).
__device__ __forceinline__ bool wf_bounce(const WfParams& P, const WfFrame& F, WfPath& p, uint32_t depth, float4& acc) {
    const float PI_F = 3.14159265358979323846f;
    // ---- pt_intersect.wgsl:389-557 ----
    float t_best = 1e30f;
    v3 n_hit = V3(0, 1, 0);
    uint32_t mat = 0;
    for (uint32_t i = 0; i < P.nsph; i++) {
        const float* s = P.spheres + 20 * (size_t)i;
        const v3 c = ld3(s);
        const float t = wf_sphere_t(p.o, p.d, c, __ldg(s + 3));
        if (t >= p.tmin && t < fminf(t_best, 1e30f)) {
            t_best = t;
            n_hit = normalize3((p.o + p.d * t) - c);
            mat = i;
        }
    }
    if (P.ninst == 0u) {
        float t; v3 n;
        if (P.mesh.ntris && wf_mesh<false>(P.mesh, p.o, p.d, p.tmin, 1e30f, t, n) && t < t_best) { t_best = t; n_hit = n; mat = 0u; }
    } else {
        for (uint32_t ii = 0; ii < P.ninst; ii++) {
            const float* w2o = P.inst + 36 * (size_t)ii + 16;
            if (__float_as_uint(__ldg(P.inst + 36 * (size_t)ii + 32)) != 0u) continue;
            float t; v3 n;
            if (wf_mesh<false>(P.mesh, wf_xform(w2o, p.o, 1.0f), normalize3(wf_xform(w2o, p.d, 0.0f)), p.tmin, 1e30f, t, n) && t < t_best) {
                t_best = t;
                n_hit = normalize3(wf_xform_normal(w2o, n));
                const uint32_t mid = __float_as_uint(__ldg(P.inst + 36 * (size_t)ii + 33));
                mat = P.nsph ? min(mid, P.nsph - 1u) : 0u;
            }
        }
    }
    if (!(t_best < 1e20f)) {   // miss, pt_scatter.wgsl:108-131
        const v3 sky = mix3(ar3(P.env + 8), ar3(P.env + 12), 0.5f * (p.d.y + 1.0f));
        const v3 c = p.thr * sky;
        acc.x += c.x; acc.y += c.y; acc.z += c.z;
        return false;
    }
    const v3 hp = p.o + p.d * t_best;
    const v3 wo_raw = normalize3(-p.d);

    // ---- pt_shade.wgsl:478-861 ----
    const uint32_t mi = mat < P.nsph ? mat : 0u;
    const float* M = P.spheres + 20 * (size_t)mi;
    const v3 albedo = ld3(M + 4);
    const float metallic = __ldg(M + 7), roughness = __ldg(M + 8), ior = __ldg(M + 9);
    const v3 emissive = ld3(M + 12);
    if (emissive.x > 0.0f || emissive.y > 0.0f || emissive.z > 0.0f) {
        const v3 c = p.thr * emissive;
        acc.x += c.x; acc.y += c.y; acc.z += c.z;
    }
    uint32_t rng = p.rng_hi ^ (p.pixel * 26699u) ^ (F.frame * 30977u);
    const v3 n = normalize3(n_hit), wo = normalize3(wo_raw);
    const float ndv = fmaxf(dot3(n, wo), 0.0f);
    const WfBasis basis = wf_basis(n);
    const float a = fmaxf(0.02f, roughness * roughness);
    const float sm = satf(metallic);
    const v3 F0 = V3(mixf(0.04f, albedo.x, sm), mixf(0.04f, albedo.y, sm), mixf(0.04f, albedo.z, sm));
    const float imp = mi < P.nimp ? __ldg(P.imp + mi) : 1.0f;
    const v3 so = hp + n * 1e-3f;

    {   // environment NEE, :584-614; sample_env_mixture :166-183
        const float e1 = xorshift32(rng), e2 = xorshift32(rng), e3 = xorshift32(rng);
        v3 wi;
        if (e1 < 0.5f) {   // sample_power_cosine_about_up :146-155
            const float phi = (2.0f * PI_F) * e3;
            const float ct = pow_pinned(1.0f - e2, fdiv(1.0f, 16.0f + 1.0f));
            const float st = fsqrt(fmaxf(0.0f, 1.0f - ct * ct));
            float s, c;
            sincos_pinned(phi, s, c);
            wi = V3(st * c, ct, st * s);
        } else {
            wi = wf_to_world(basis, wf_cosine_hemisphere(e2, e3));
        }
        const float pdf_up = wf_power_cosine_pdf_up(wi);
        const float cos_surf = fmaxf(dot3(n, wi), 0.0f);
        const float pdf_cos = fdiv(cos_surf, PI_F);
        const float pdf_light = 0.5f * pdf_up + (1.0f - 0.5f) * pdf_cos;
        if (cos_surf > 0.0f) {
            const WfBsdf br = wf_bsdf(wo, wi, n, albedo, metallic, roughness);
            const float w_mis = fdiv(pdf_light, fmaxf(pdf_light + br.pdf, 1e-8f));
            const v3 L_env = mix3(ar3(P.env), ar3(P.env + 4), 0.5f * (wi.y + 1.0f));
            const v3 c = ((((p.thr * br.f) * L_env) * fdiv(cos_surf, fmaxf(pdf_light, 1e-8f))) * w_mis) * imp;
            if (!wf_occluded(P, so, wi, 1e-3f, 1e30f)) { acc.x += c.x; acc.y += c.y; acc.z += c.z; }
        }
    }
    if (P.ndir) {   // delta lights, :617-657
        float sum_imp = 0.0f;
        for (uint32_t i = 0; i < P.ndir; i++) sum_imp = sum_imp + fmaxf(__ldg(P.dirl + 8 * (size_t)i + 7), 0.0f);
        uint32_t idx = 0;
        const float u = xorshift32(rng);
        if (sum_imp > 0.0f) {
            const float rsel = u * sum_imp;
            float run = 0.0f;
            for (uint32_t i = 0; i < P.ndir; i++) { run = run + fmaxf(__ldg(P.dirl + 8 * (size_t)i + 7), 0.0f); if (rsel <= run) { idx = i; break; } }
        } else {
            idx = (uint32_t)floorf(u * (float)P.ndir);
        }
        const float* L = P.dirl + 8 * (size_t)min(idx, P.ndir - 1u);
        const v3 wi = normalize3(-ld3(L));
        const float cos_surf = fmaxf(dot3(n, wi), 0.0f);
        if (cos_surf > 0.0f) {
            const WfBsdf br = wf_bsdf(wo, wi, n, albedo, metallic, roughness);
            const v3 Li = ld3(L + 4) * __ldg(L + 3);
            const float p_sel = sum_imp > 0.0f ? fdiv(fmaxf(__ldg(L + 7), 0.0f), fmaxf(sum_imp, 1e-8f)) : fdiv(1.0f, (float)P.ndir);
            const v3 c = (((p.thr * br.f) * Li) * fdiv(cos_surf, fmaxf(p_sel, 1e-8f))) * imp;
            if (!wf_occluded(P, so, wi, 1e-3f, 1e30f)) { acc.x += c.x; acc.y += c.y; acc.z += c.z; }
        }
    }
    if (P.narea) {   // disc lights, :660-707; sample_area_light_disc :114-143
        float sum_imp = 0.0f;
        for (uint32_t i = 0; i < P.narea; i++) sum_imp = sum_imp + fmaxf(__ldg(P.areal + 12 * (size_t)i + 11), 0.0f);
        uint32_t idx = 0;
        const float u = xorshift32(rng);
        if (sum_imp > 0.0f) {
            const float rsel = u * sum_imp;
            float run = 0.0f;
            for (uint32_t i = 0; i < P.narea; i++) { run = run + fmaxf(__ldg(P.areal + 12 * (size_t)i + 11), 0.0f); if (rsel <= run) { idx = i; break; } }
        } else {
            idx = (uint32_t)floorf(u * (float)P.narea);
        }
        const float* L = P.areal + 12 * (size_t)min(idx, P.narea - 1u);
        const float a1 = xorshift32(rng), a2 = xorshift32(rng);
        const v3 nL = normalize3(ld3(L + 4));
        const WfBasis bl = wf_basis(nL);
        // the shader reads basisL[0][0], basisL[1][0], basisL[2][0]: the first ROW of the (t, b, n) matrix, :118-119
        const v3 tL = V3(bl.t.x, bl.b.x, bl.n.x), bL = V3(bl.t.y, bl.b.y, bl.n.y);
        const float rad = fmaxf(__ldg(L + 3), 1e-6f);
        const float r = fsqrt(a1) * rad, phi = (2.0f * PI_F) * a2;
        float s, c;
        sincos_pinned(phi, s, c);
        const v3 X = (ld3(L) + tL * (r * c)) + bL * (r * s);
        const v3 dir = X - hp;
        const float d = fsqrt(dot3(dir, dir));
        if (d > 1e-6f) {
            const v3 wi = V3(fdiv(dir.x, d), fdiv(dir.y, d), fdiv(dir.z, d));
            const float cos_s = fmaxf(dot3(n, wi), 0.0f), cos_l = fmaxf(dot3(nL, -wi), 0.0f);
            if (cos_s > 0.0f && cos_l > 0.0f) {
                const float area = (PI_F * rad) * rad;
                const float pdf = fdiv(frcp(area) * (d * d), fmaxf(cos_l, 1e-6f));
                if (pdf > 0.0f) {
                    const WfBsdf br = wf_bsdf(wo, wi, n, albedo, metallic, roughness);
                    const float p_sel = sum_imp > 0.0f ? fdiv(fmaxf(__ldg(L + 11), 0.0f), fmaxf(sum_imp, 1e-8f)) : fdiv(1.0f, (float)P.narea);
                    const float pdf_light = p_sel * pdf;
                    const float w_mis = fdiv(pdf_light, fmaxf(pdf_light + br.pdf, 1e-8f));
                    const v3 Li = ld3(L + 8) * __ldg(L + 7);
                    const v3 cc = ((((p.thr * br.f) * Li) * fdiv(cos_s, fmaxf(pdf_light, 1e-8f))) * w_mis) * imp;
                    if (!wf_occluded(P, so, wi, 1e-3f, d - 1e-3f)) { acc.x += cc.x; acc.y += cc.y; acc.z += cc.z; }
                }
            }
        }
    }

    // continuation, :738-808
    v3 wi, nthr;
    if (metallic > 0.5f) {
        const float m1 = xorshift32(rng), m2 = xorshift32(rng);
        const float a2 = a * a;   // sample_ggx_isotropic :404-413
        const float ch = fsqrt(fdiv(1.0f - m1, 1.0f + (a2 - 1.0f) * m1));
        const float sh = fsqrt(fmaxf(0.0f, 1.0f - ch * ch));
        const float phi = (2.0f * PI_F) * m2;
        float s, c;
        sincos_pinned(phi, s, c);
        const v3 hw = normalize3(wf_to_world(basis, V3(sh * c, sh * s, ch)));
        wi = normalize3(wf_reflect(-wo, hw));
        const float ndl = fmaxf(dot3(n, wi), 0.0f), ndh = fmaxf(dot3(n, hw), 0.0f), vdh = fmaxf(dot3(wo, hw), 0.0f);
        if (!(ndl > 0.0f && ndv > 0.0f)) return false;   // invalid sample: the shader thread moves on to its next hit, :771-774
        const float D = fdiv(a2, fmaxf(PI_F * pow2f((ndh * ndh) * (a2 - 1.0f) + 1.0f), 1e-6f));
        const float k = fdiv(pow2f(a + 1.0f), 8.0f);
        const float G = fdiv(ndl, ndl * (1.0f - k) + k) * fdiv(ndv, ndv * (1.0f - k) + k);
        const float fw = pow5f(1.0f - satf(vdh));
        const v3 Fr = V3(F0.x + (1.0f - F0.x) * fw, F0.y + (1.0f - F0.y) * fw, F0.z + (1.0f - F0.z) * fw);
        const v3 spec = Fr * fdiv(D * G, fmaxf((4.0f * ndl) * ndv, 1e-6f));
        const float pdf = fdiv(D * ndh, fmaxf(4.0f * vdh, 1e-6f));
        nthr = (p.thr * spec) * fdiv(ndl, fmaxf(pdf, 1e-6f));
    } else if (ior > 1.01f) {
        const float cosi = satf(dot3(n, wo));
        const float F0s = pow2f(fdiv(ior - 1.0f, ior + 1.0f));
        const float Fr = F0s + (1.0f - F0s) * pow5f(1.0f - cosi);
        const float u = xorshift32(rng);
        if (u < Fr) {
            wi = normalize3(wf_reflect(-wo, n));
        } else {
            const bool entering = dot3(n, wo) > 0.0f;
            const float eta = entering ? fdiv(1.0f, ior) : fdiv(ior, 1.0f);
            const v3 N = entering ? n : -n;
            wi = normalize3(wf_refract(-wo, N, eta));
            if (!(dot3(wi, wi) >= 1e-12f)) wi = normalize3(wf_reflect(-wo, n));   // total internal reflection (normalize(0) is NaN)
        }
        nthr = p.thr * V3(fmaxf(albedo.x, 0.0f), fmaxf(albedo.y, 0.0f), fmaxf(albedo.z, 0.0f));
    } else {
        const float l1 = xorshift32(rng), l2 = xorshift32(rng);
        wi = normalize3(wf_to_world(basis, wf_cosine_hemisphere(l1, l2)));
        const float ct = fmaxf(0.0f, dot3(n, wi));
        const float pdf = fdiv(ct, PI_F) + 1e-8f;
        nthr = (p.thr * V3(fdiv(albedo.x, PI_F), fdiv(albedo.y, PI_F), fdiv(albedo.z, PI_F))) * fdiv(ct, pdf);
    }
    // Russian roulette, :811-829 (adaptive threshold 0.25, wavefront/mod.rs:104)
    float rr_scale = 1.0f;
    if (depth >= 4u) {
        const float max_c = fmaxf(nthr.x, fmaxf(nthr.y, nthr.z));
        float q = fminf(fmaxf(1.0f - max_c, 0.0f), 0.95f);
        const float q_extra = fminf(fmaxf(1.0f - fdiv(max_c, 0.25f), 0.0f), 0.90f);
        q = fminf(fmaxf(q + q_extra, 0.0f), 0.95f);
        const float u = xorshift32(rng);
        if (u < q) return false;
        rr_scale = fdiv(1.0f, 1.0f - q);
    }
    if (!(depth + 1u < kWfMaxDepth)) return false;
    // scatter, :831-848; pt_scatter.wgsl:77-106
    p.o = hp + normalize3(n_hit) * 1e-3f;
    p.d = wi;
    p.tmin = 1e-3f;
    p.thr = nthr * rr_scale;
    p.rng_hi = rng;
    return true;
}

// Bounce `depth` of every path in the queue, for a batch of frames at once (depth 0: one primary ray per owned pixel per frame,
// generated in place).  Survivors are compacted into the other queue with one atomic per warp.  Launched with a fixed grid; the queue
// length is read from device memory.  A path adds into its own frame's radiance sum (fsum), never into another frame's.
template <bool PRIMARY>
__global__ void __launch_bounds__(kWfThreads, F3D_WF_MIN_CTAS) k_wf_bounce(WfParams P, WfBatch B, uint32_t depth) {
    const uint32_t per_frame = P.w * P.local_rows;
    const uint32_t count = PRIMARY ? per_frame * B.nframes : B.qcount[depth];
    const uint32_t in = depth & 1u, out = in ^ 1u;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * blockDim.x;
    const size_t npx = (size_t)P.w * P.h;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < count; base += stride) {   // warp-uniform trip count
        const uint32_t i = base + lane;
        bool alive = false, skip = i >= count;
        uint32_t fl = 0;
        WfPath p;
        if (!skip) {
            if (PRIMARY) {
                fl = i / per_frame;
                const uint32_t li = i - fl * per_frame;
                // local row lr of this rank -> global row: its blocks are rank, rank + world, rank + 2 world, ...
                const uint32_t lr = li / P.w, x = li - lr * P.w;
                const uint32_t row = ((lr / P.part_rows) * P.part_world + P.part_rank) * P.part_rows + lr % P.part_rows;
                skip = row >= P.h;
                if (!skip) p = wf_raygen(P, B.frame(fl), row * P.w + x);
            } else {
                const float4 a = P.qa[in][i], b = P.qb[in][i], c = P.qc[in][i];
                p.o = V3(a.x, a.y, a.z); p.pixel = __float_as_uint(a.w);
                p.d = V3(b.x, b.y, b.z); p.rng_hi = __float_as_uint(b.w);
                p.thr = V3(c.x, c.y, c.z); fl = __float_as_uint(c.w); p.tmin = 1e-3f;
            }
            if (!skip) {
                float4* sum = P.fsum + fl * npx + p.pixel;
                float4 acc = PRIMARY ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : *sum;
                alive = wf_bounce(P, B.frame(fl), p, depth, acc);
                acc.w = __uint_as_float(depth + 1u);   // rays this path has traced so far: k_wf_merge turns these into the per-frame counts
                *sum = acc;
            }
        }
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, alive);
        if (mask) {
            uint32_t slot = 0;
            if (lane == 0u) slot = atomicAdd(B.qcount + depth + 1u, (uint32_t)__popc(mask));
            slot = __shfl_sync(0xFFFFFFFFu, slot, 0) + (uint32_t)__popc(mask & ((1u << lane) - 1u));
            if (alive) {
                P.qa[out][slot] = make_float4(p.o.x, p.o.y, p.o.z, __uint_as_float(p.pixel));
                P.qb[out][slot] = make_float4(p.d.x, p.d.y, p.d.z, __uint_as_float(p.rng_hi));
                P.qc[out][slot] = make_float4(p.thr.x, p.thr.y, p.thr.z, __uint_as_float(fl));
            }
        }
    }
}

// The thin tail: every path still alive at `depth` is walked to its end by one thread (a few percent of the pixels are left by then,
// and they die off geometrically under Russian roulette, so compacting them bounce by bounce would cost more launches than work).
__global__ void __launch_bounds__(kWfThreads, F3D_WF_MIN_CTAS) k_wf_tail(WfParams P, WfBatch B, uint32_t depth) {
    const uint32_t count = B.qcount[depth];
    const uint32_t in = depth & 1u;
    const size_t npx = (size_t)P.w * P.h;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const float4 a = P.qa[in][i], b = P.qb[in][i], c = P.qc[in][i];
        WfPath p;
        p.o = V3(a.x, a.y, a.z); p.pixel = __float_as_uint(a.w);
        p.d = V3(b.x, b.y, b.z); p.rng_hi = __float_as_uint(b.w);
        p.thr = V3(c.x, c.y, c.z); p.tmin = 1e-3f;
        const uint32_t fl = __float_as_uint(c.w);
        const WfFrame F = B.frame(fl);
        float4* sum = P.fsum + fl * npx + p.pixel;
        float4 acc = *sum;
        uint32_t d = depth;
        while (wf_bounce(P, F, p, d, acc)) d++;          // wf_bounce ends every path at depth kWfMaxDepth - 1
        acc.w = __uint_as_float(d + 1u);
        *sum = acc;
    }
}

// End of a batch: every owned pixel adds its frames' radiance sums to the accumulator IN FRAME ORDER, so the image does not depend on
// how many frames shared a batch.  The same pass bins the paths by the number of rays they traced (left in the sums' fourth lane by the
// bounce kernels) into a per-frame histogram - shared-memory atomics, one flush per CTA - from which the host derives the per-frame,
// per-depth ray counts the reference's two frame rules need; the bounce kernels themselves carry no counting atomics.
__global__ void __launch_bounds__(kWfThreads) k_wf_merge(WfParams P, uint32_t first_frame, uint32_t nframes, uint32_t* __restrict__ hist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];      // nframes * (kWfMaxDepth + 1) counters
    uint32_t* bins = reinterpret_cast<uint32_t*>(smem_raw);
    const uint32_t nbins = nframes * (kWfMaxDepth + 1u);
    for (uint32_t t = threadIdx.x; t < nbins; t += blockDim.x) bins[t] = 0u;
    __syncthreads();
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li < P.w * P.local_rows) {
        const uint32_t lr = li / P.w, x = li - lr * P.w;
        const uint32_t row = ((lr / P.part_rows) * P.part_world + P.part_rank) * P.part_rows + lr % P.part_rows;
        if (row < P.h) {
            const size_t npx = (size_t)P.w * P.h, pix = (size_t)row * P.w + x;
            float4 a = P.accum[pix];
            for (uint32_t f = 0; f < nframes; f++) {
                const float4 s = P.fsum[f * npx + pix];
                a.x += s.x; a.y += s.y; a.z += s.z;
                atomicAdd(bins + f * (kWfMaxDepth + 1u) + min(__float_as_uint(s.w), kWfMaxDepth), 1u);
            }
            P.accum[pix] = a;
        }
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < nbins; t += blockDim.x)
        if (bins[t]) atomicAdd(hist + (size_t)first_frame * (kWfMaxDepth + 1u) + t, bins[t]);
}

// adjudication.rs:318-331 (mean over frames, alpha 1) + resolve_reference_hdr_to_rgba8, src/core/tonemap.rs:11-32
__global__ void k_wf_resolve(const float4* __restrict__ accum, uint32_t npx, float inv_spp, float exposure, float4* __restrict__ hdr,
                             uchar4* __restrict__ rgba8) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npx) return;
    const float4 a = accum[i];
    const float v[3] = {a.x * inv_spp, a.y * inv_spp, a.z * inv_spp};
    if (hdr) hdr[i] = make_float4(v[0], v[1], v[2], 1.0f);
    if (rgba8) {
        uint32_t o[3];
        for (int c = 0; c < 3; c++) {
            const float x = fmaxf(v[c], 0.0f) * exposure;
            const float t = fdiv(x, 1.0f + x);
            const float e = t <= 0.0031308f ? 12.92f * t : 1.055f * pow_pinned(t, fdiv(1.0f, 2.4f)) - 0.055f;
            const float s = fminf(fmaxf(e, 0.0f), 1.0f);
            o[c] = (uint32_t)(s * 255.0f + 0.5f);
        }
        rgba8[i] = make_uchar4((unsigned char)o[0], (unsigned char)o[1], (unsigned char)o[2], 255);
    }
}

}  // namespace f3d
