// forge3d_b200/csrc/f3d_trace.cuh
// Scene parameters + the LITERAL restatement of the WGSL traversal (terrain_trace: kept as the on-device
// cross-check of the production traversal in f3d_trace_fast.cuh, reachable through f3d_trace_rays
// variant 1), the triangle-mesh test with its BVH, and the environment lookup.
// Replaces /root/reference/src/shaders/hybrid_terrain_traversal.wgsl:88-383 and
// hybrid_traversal.wgsl:86-259.  Arithmetic follows the numerics contract (f3d_math.cuh) so the
// results are bit-identical to oracle/f3d_oracle.c; the data layout is B200-native:
//   * leaf cells are pre-packed (h00,h10,h01,h11)*exaggeration -> ONE 128-bit load per leaf
//     (the reference issues four R32F textureLoads, :149-156, and four more for the normal, :240);
//   * min-max levels are plain float2 arrays pre-multiplied by the exaggeration (one 64-bit load
//     per node, the reference multiplies per visit, :301-302).  min/max commute with the
//     multiplication by a positive finite factor under round-to-nearest, so values are identical.
#pragma once
#include "f3d_math.cuh"

namespace f3d {

constexpr int kMaxLevels = 16;       // DEM <= 8192 cells per axis (13-bit node packing, :143-146)
constexpr uint32_t kStackSize = 64;  // TERRAIN_STACK_SIZE, :73

struct SceneParams {
    // TerrainPtUniforms (terrain_heightfield.rs:31-38,348-369)
    float ox, oz, sx, sz;
    float albedo[3];
    float env_intensity;
    uint32_t cell_w, cell_h, mip_count;
    const float4* cells;             // cell_w * cell_h packed corner heights (pre-exaggerated)
    const float2* mm[kMaxLevels];    // per level [min,max] (pre-exaggerated), row pitch mm_pitch[l]
    uint32_t mm_pitch[kMaxLevels];
    // EarthCurvatureUniforms (terrain_heightfield.rs:44-84)
    float inv_two_r_prime;
    uint32_t curvature_enabled;
    // environment (RGBA32F texels, terrain_heightfield.rs:443-482)
    const float4* env;
    uint32_t env_w, env_h;
    // mesh (HybridUniforms, hybrid_traversal.wgsl:9-17)
    const float4* mesh_v;
    const uint32_t* mesh_i;
    uint32_t mesh_index_count, mesh_nverts;
    // mesh BVH (csrc/f3d_backend.cu::build_mesh_bvh); NULL = index-order sweep
    const float4* bvh_nodes;         // 2 x float4 per node: (bmin, left|first) (bmax, count)
    const uint32_t* bvh_tris;        // triangle ids referenced by the leaves
    uint32_t traversal_mode;         // 0 hybrid, 3 terrain only
};

struct Ray { v3 o; float tmin; v3 d; float tmax; };
struct Hit { float t; v3 point; v3 normal; uint32_t hit_type; uint32_t hit; };

__device__ __forceinline__ float safe_inv(float d) {  // terrain_safe_inv :88-91
    float ad = fmaxf(fabsf(d), 1e-12f);
    return d < 0.0f ? -frcp(ad) : frcp(ad);
}

// Per-ray constants hoisted out of the node loop (the WGSL recomputes them per slab call with
// identical results).
struct RayCtx {
    float inv_x, inv_z;   // safe reciprocals
    float hd2;            // dot(dir.xz, dir.xz)
};

template <bool CURV>
__device__ __forceinline__ float curved_height(const SceneParams& S, const Ray& r, const RayCtx& c, float t) {  // :95-103
    float corr = 0.0f;
    if (CURV) corr = (t * t * c.hd2) * S.inv_two_r_prime;
    return r.o.y + t * r.d.y + corr;
}

template <bool CURV>
__device__ __forceinline__ void curved_height_range(const SceneParams& S, const Ray& r, const RayCtx& c,
                                                    float t0, float t1, float& lo, float& hi) {  // :108-127
    float y0 = curved_height<CURV>(S, r, c, t0);
    float y1 = curved_height<CURV>(S, r, c, t1);
    float minimum = fminf(y0, y1);
    if (CURV) {
        float a = c.hd2 * S.inv_two_r_prime;
        if (a > 0.0f) {
            float vertex = fdiv(-r.d.y, 2.0f * a);
            if (vertex >= t0 && vertex <= t1) minimum = fminf(minimum, curved_height<true>(S, r, c, vertex));
        }
    }
    lo = minimum;
    hi = fmaxf(y0, y1);
}

__device__ __forceinline__ void slab_xz(const Ray& r, const RayCtx& c, float x0, float x1, float z0, float z1,
                                        float& te, float& tx) {  // :131-141
    float tx0 = (x0 - r.o.x) * c.inv_x, tx1 = (x1 - r.o.x) * c.inv_x;
    if (tx0 > tx1) { float tmp = tx0; tx0 = tx1; tx1 = tmp; }
    float tz0 = (z0 - r.o.z) * c.inv_z, tz1 = (z1 - r.o.z) * c.inv_z;
    if (tz0 > tz1) { float tmp = tz0; tz0 = tz1; tz1 = tmp; }
    te = fmaxf(tx0, tz0);
    tx = fminf(tx1, tz1);
}

__device__ __forceinline__ uint32_t pack_node(uint32_t level, uint32_t x, uint32_t y) {  // :144-146
    return (level << 26) | (y << 13) | x;
}

// terrain_leaf_intersect :167-235.  h = (h00,h10,h01,h11) of the cell.
template <bool ANY_HIT, bool CURV>
__device__ __forceinline__ bool leaf_intersect(const SceneParams& S, const Ray& r, const RayCtx& c, float4 h,
                                               uint32_t cx, uint32_t cz, float t0, float t1, float& t_out) {
    float tm = 0.5f * (t0 + t1);
    float d3[3];
    const float fcx = (float)cx, fcz = (float)cz;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float t = i == 0 ? t0 : (i == 1 ? tm : t1);
        float px = r.o.x + t * r.d.x;
        float pz = r.o.z + t * r.d.z;
        float u = clampf(fdiv(px - S.ox, S.sx) - fcx, 0.0f, 1.0f);
        float v = clampf(fdiv(pz - S.oz, S.sz) - fcz, 0.0f, 1.0f);
        float hh = mixf(mixf(h.x, h.y, u), mixf(h.z, h.w, u), v);
        d3[i] = curved_height<CURV>(S, r, c, t) - hh;
    }
    float cc = d3[0];
    float a = 2.0f * d3[2] + 2.0f * d3[0] - 4.0f * d3[1];
    float b = d3[2] - d3[0] - a;
    float s_hit = 1e30f;
    if (ANY_HIT && cc <= 0.0f) s_hit = 0.0f;
    else if (fabsf(a) < 1e-12f) {
        if (fabsf(b) > 1e-12f) {
            float s = fdiv(-cc, b);
            if (s >= 0.0f && s <= 1.0f) s_hit = s;
        }
    } else {
        float disc = b * b - 4.0f * a * cc;
        if (disc >= 0.0f) {
            float sq = fsqrt(disc);
            float q = -0.5f * (b + (b >= 0.0f ? sq : -sq));
            float r0 = fdiv(q, a);
            float r1 = fabsf(q) < 1e-30f ? 1e30f : fdiv(cc, q);
            if (r0 > r1) { float tmp = r0; r0 = r1; r1 = tmp; }
            if (r0 >= 0.0f && r0 <= 1.0f) s_hit = r0;
            else if (r1 >= 0.0f && r1 <= 1.0f) s_hit = r1;
        }
    }
    if (s_hit <= 1.0f) {
        float t = t0 + s_hit * (t1 - t0);
        if (t > r.tmin && t < r.tmax) { t_out = t; return true; }
    }
    return false;
}

// terrain_normal_at :239-248
__device__ __forceinline__ v3 normal_at(const SceneParams& S, float4 h, v3 p, uint32_t cx, uint32_t cz) {
    float u = clampf(fdiv(p.x - S.ox, S.sx) - (float)cx, 0.0f, 1.0f);
    float v = clampf(fdiv(p.z - S.oz, S.sz) - (float)cz, 0.0f, 1.0f);
    float dh_du = mixf(h.y - h.x, h.w - h.z, v);
    float dh_dv = mixf(h.z - h.x, h.w - h.y, u);
    return normalize3(V3(fdiv(-dh_du, S.sx), 1.0f, fdiv(-dh_dv, S.sz)));
}

// terrain_trace :254-372.  `nodes` counts stack pops (parity statistic).
template <bool ANY_HIT, bool CURV>
__device__ __noinline__ Hit terrain_trace(const SceneParams& S, const Ray& r, uint32_t& nodes) {
    Hit res;
    res.hit = 0u;
    res.t = r.tmax;
    res.hit_type = 3u;
    res.point = V3(0, 0, 0);
    res.normal = V3(0, 0, 0);
    const uint32_t cell_w = S.cell_w, cell_h = S.cell_h;
    const float ox = S.ox, oz = S.oz, sx = S.sx, sz = S.sz;
    RayCtx c;
    c.inv_x = safe_inv(r.d.x);
    c.inv_z = safe_inv(r.d.z);
    c.hd2 = dot2(r.d.x, r.d.z, r.d.x, r.d.z);

    uint32_t stack[kStackSize];
    uint32_t sp = 0;
    stack[sp++] = pack_node(S.mip_count - 1u, 0u, 0u);
    while (sp != 0u) {
        sp--;
        nodes++;
        const uint32_t node = stack[sp];
        const uint32_t level = node >> 26, ny = (node >> 13) & 0x1FFFu, nx = node & 0x1FFFu;
        const uint32_t cx0 = nx << level, cz0 = ny << level;
        if (cx0 >= cell_w || cz0 >= cell_h) continue;
        const uint32_t cx1 = min((nx + 1u) << level, cell_w);
        const uint32_t cz1 = min((ny + 1u) << level, cell_h);
        float s0, s1;
        slab_xz(r, c, ox + (float)cx0 * sx, ox + (float)cx1 * sx, oz + (float)cz0 * sz, oz + (float)cz1 * sz, s0, s1);
        const float t_lo = fmaxf(s0, r.tmin);
        const float t_hi = fminf(s1, fminf(r.tmax, res.t));
        if (t_lo > t_hi) continue;
        const float2 mm = __ldg(S.mm[level] + (size_t)ny * S.mm_pitch[level] + nx);
        float ry_lo, ry_hi;
        curved_height_range<CURV>(S, r, c, t_lo, t_hi, ry_lo, ry_hi);
        if (ry_lo > mm.y || ry_hi < mm.x) continue;
        if (level == 0u) {
            const float4 h = __ldg(S.cells + (size_t)cz0 * cell_w + cx0);
            float lt;
            if (leaf_intersect<ANY_HIT, CURV>(S, r, c, h, cx0, cz0, t_lo, t_hi, lt) && lt < res.t) {
                res.hit = 1u;
                res.t = lt;
                res.point = r.o + r.d * lt;
                res.normal = normal_at(S, h, res.point, cx0, cz0);
                res.hit_type = 3u;
                if (ANY_HIT) return res;
            }
            continue;
        }
        const uint32_t child_level = level - 1u;
        float child_t[4];
        uint32_t child_id[4];
        uint32_t child_count = 0;
#pragma unroll
        for (uint32_t cy = 0; cy < 2u; cy++) {
#pragma unroll
            for (uint32_t cxi = 0; cxi < 2u; cxi++) {
                const uint32_t ccx = nx * 2u + cxi, ccy = ny * 2u + cy;
                const uint32_t gx0 = ccx << child_level, gz0 = ccy << child_level;
                if (gx0 >= cell_w || gz0 >= cell_h) continue;
                const uint32_t gx1 = min((ccx + 1u) << child_level, cell_w);
                const uint32_t gz1 = min((ccy + 1u) << child_level, cell_h);
                float c0, c1;
                slab_xz(r, c, ox + (float)gx0 * sx, ox + (float)gx1 * sx, oz + (float)gz0 * sz, oz + (float)gz1 * sz, c0, c1);
                const float ct_lo = fmaxf(c0, t_lo), ct_hi = fminf(c1, t_hi);
                if (ct_lo > ct_hi) continue;
                child_t[child_count] = ct_lo;
                child_id[child_count] = pack_node(child_level, ccx, ccy);
                child_count++;
            }
        }
        // insertion sort, descending t_enter, stable (:351-363)
        for (uint32_t i = 1; i < child_count; i++) {
            const float kt = child_t[i];
            const uint32_t kid = child_id[i];
            uint32_t j = i;
            while (!(j == 0u || child_t[j - 1u] >= kt)) {
                child_t[j] = child_t[j - 1u];
                child_id[j] = child_id[j - 1u];
                j--;
            }
            child_t[j] = kt;
            child_id[j] = kid;
        }
        for (uint32_t i = 0; i < child_count; i++)
            if (sp < kStackSize) stack[sp++] = child_id[i];
    }
    return res;
}

// ray_triangle_intersect, hybrid_traversal.wgsl:86-132
__device__ __forceinline__ bool ray_triangle(const Ray& r, v3 v0, v3 v1, v3 v2, float& t_out, v3& n_out) {
    v3 e1 = v1 - v0, e2 = v2 - v0;
    v3 h = cross3(r.d, e2);
    float a = dot3(e1, h);
    if (fabsf(a) < 1e-7f) return false;
    float f = frcp(a);
    v3 s = r.o - v0;
    float u = f * dot3(s, h);
    if (u < 0.0f || u > 1.0f) return false;
    v3 q = cross3(s, e1);
    float v = f * dot3(r.d, q);
    if (v < 0.0f || u + v > 1.0f) return false;
    float t = f * dot3(e2, q);
    if (t > r.tmin && t < r.tmax) {
        t_out = t;
        n_out = normalize3(cross3(e1, e2));
        return true;
    }
    return false;
}

// One triangle of the mesh against the ray; keeps the lexicographically smallest (t, triangle index),
// which is exactly what the reference's index-order sweep with a strict '<' produces (:150-169).
__device__ __forceinline__ void mesh_test_triangle(const SceneParams& S, const Ray& r, uint32_t tri_id, Hit& res, uint32_t& best_id) {
    const uint32_t tri = tri_id * 3u;
    const uint32_t i0 = __ldg(S.mesh_i + tri), i1 = __ldg(S.mesh_i + tri + 1), i2 = __ldg(S.mesh_i + tri + 2);
    if (i0 >= S.mesh_nverts || i1 >= S.mesh_nverts || i2 >= S.mesh_nverts) return;
    const float4 a = __ldg(S.mesh_v + i0), b = __ldg(S.mesh_v + i1), c = __ldg(S.mesh_v + i2);
    float t;
    v3 n;
    if (ray_triangle(r, V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), t, n) &&
        (t < res.t || (t == res.t && res.hit && tri_id < best_id))) {
        res.hit = 1u;
        res.t = t;
        res.point = r.o + r.d * t;
        res.normal = n;
        best_id = tri_id;
    }
}

// intersect_mesh, hybrid_traversal.wgsl:137-172.  The reference sweeps every triangle in index order
// (its uploaded BVH is never traversed); here a BVH prunes the sweep.  Each candidate triangle runs the
// identical Moeller-Trumbore test against the same ray, boxes are padded so that rounding can never cull
// the winning triangle, and ties resolve to the lowest triangle index, so the closest hit is identical.
__device__ __noinline__ Hit intersect_mesh(const SceneParams& S, const Ray& r) {
    Hit res;
    res.hit = 0u;
    res.t = r.tmax;
    res.hit_type = 0u;
    res.point = V3(0, 0, 0);
    res.normal = V3(0, 0, 0);
    const uint32_t ic = S.mesh_index_count;
    if (ic < 3u) return res;
    uint32_t best_id = 0xFFFFFFFFu;
    if (S.bvh_nodes == nullptr) {
        for (uint32_t tri = 0; tri + 2u < ic; tri += 3u) mesh_test_triangle(S, r, tri / 3u, res, best_id);
        return res;
    }
    const float ix = frcp(r.d.x), iy = frcp(r.d.y), iz = frcp(r.d.z);   // +-inf for axis-parallel rays: handled by fmin/fmax
    uint32_t stack[64];                      // a Karras tree over 62-bit unique keys is at most 62 deep
    uint32_t sp = 0;
    stack[sp++] = 0u;
    while (sp) {
        const uint32_t ni = stack[--sp];
        const float4 n0 = __ldg(S.bvh_nodes + 2 * (size_t)ni), n1 = __ldg(S.bvh_nodes + 2 * (size_t)ni + 1);
        const float tx0 = (n0.x - r.o.x) * ix, tx1 = (n1.x - r.o.x) * ix;
        const float ty0 = (n0.y - r.o.y) * iy, ty1 = (n1.y - r.o.y) * iy;
        const float tz0 = (n0.z - r.o.z) * iz, tz1 = (n1.z - r.o.z) * iz;
        const float tnear = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), r.tmin));
        const float tfar = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), res.t));
        if (!(tnear <= tfar)) continue;
        // leaf: n0.w = 0x80000000 | first index into bvh_tris, n1.w = count; internal: n0.w / n1.w = left / right child
        const uint32_t a = __float_as_uint(n0.w), b = __float_as_uint(n1.w);
        if (a & 0x80000000u) {
            const uint32_t first = a & 0x7FFFFFFFu;
            for (uint32_t k = 0; k < b; k++) mesh_test_triangle(S, r, __ldg(S.bvh_tris + first + k), res, best_id);
        } else if (sp + 2u <= 64u) {
            stack[sp++] = a;
            stack[sp++] = b;
        }
    }
    return res;
}

// terrain_env_radiance :392-405
__device__ __forceinline__ v3 env_radiance(const SceneParams& S, v3 dir) {
    const float I = S.env_intensity;
    const uint32_t ew = S.env_w, eh = S.env_h;
    if (ew == 0u || eh == 0u) return V3(I, I, I);
    const float PI_F = 3.14159265358979323846f;
    v3 d = normalize3(dir);
    float uu = fdiv(atan2_pinned(d.z, d.x), 2.0f * PI_F) + 0.5f;
    float vv = fdiv(acos_pinned(clampf(d.y, -1.0f, 1.0f)), PI_F);
    float fx = uu * (float)ew, fy = vv * (float)eh;
    uint32_t px = fx <= 0.0f ? 0u : (fx >= 4294967040.0f ? 0xFFFFFFFFu : (uint32_t)fx);
    uint32_t py = fy <= 0.0f ? 0u : (fy >= 4294967040.0f ? 0xFFFFFFFFu : (uint32_t)fy);
    px = min(px, ew - 1u);
    py = min(py, eh - 1u);
    const float4 t = __ldg(S.env + (size_t)py * ew + px);
    return V3(t.x * I, t.y * I, t.z * I);
}

}  // namespace f3d
