// CPU emulation of the PRODUCTION traversal (forge3d_b200/csrc/f3d_trace_fast.cuh) as a one-lane warp.
// Test infrastructure (see cuda_runtime.h in this directory).  Builds the B200 data layout on the host exactly as
// k_build_level0 / k_reduce_level / k_pack_quads + fill_fast_scene do (csrc/f3d_kernels.cuh, csrc/f3d_backend.cu),
// then runs trace_fast<ANY_HIT,CURV> (variant 0) or the literal terrain_trace (variant 1) over a ray batch.
#include "f3d_trace_fast.cuh"

#include <vector>

using namespace f3d;

namespace {

struct HostTerrain {
    std::vector<float4> cells;
    std::vector<std::vector<float2>> plain;   // per level, pitch lw[l]
    std::vector<std::vector<float2>> quads;   // per level l: grouped by level l+1 parent
    std::vector<uint32_t> lw, lh, ppitch;
    SceneParams S;
    FastScene F;
};

void build(HostTerrain& T, const float* heights, uint32_t w, uint32_t h, const float spacing[2], const float origin[2],
           float ex, float k, int curv_enabled) {
    const uint32_t cw = w - 1u, ch = h - 1u;
    uint32_t pw = 1u, ph = 1u;
    while (pw < cw) pw <<= 1;
    while (ph < ch) ph <<= 1;
    const float inf = __int_as_float(0x7f800000), ninf = __int_as_float((int32_t)0xff800000);
    T.cells.assign((size_t)cw * ch, make_float4(0, 0, 0, 0));
    T.plain.emplace_back((size_t)pw * ph, make_float2(inf, ninf));
    T.lw.push_back(pw); T.lh.push_back(ph);
    for (uint32_t y = 0; y < ch; y++)
        for (uint32_t x = 0; x < cw; x++) {
            const size_t i = (size_t)y * w + x;
            const float a = heights[i] * ex, b = heights[i + 1] * ex, c = heights[i + w] * ex, d = heights[i + w + 1] * ex;
            T.cells[(size_t)y * cw + x] = make_float4(a, b, c, d);
            T.plain[0][(size_t)y * pw + x] = make_float2(fminf(fminf(fminf(a, b), c), d), fmaxf(fmaxf(fmaxf(a, b), c), d));
        }
    while (T.lw.back() > 1u || T.lh.back() > 1u) {
        const uint32_t lw = T.lw.back(), lh = T.lh.back();
        const uint32_t nw = lw / 2u > 1u ? lw / 2u : 1u, nh = lh / 2u > 1u ? lh / 2u : 1u;
        std::vector<float2> next((size_t)nw * nh);
        const std::vector<float2>& prev = T.plain.back();
        for (uint32_t y = 0; y < nh; y++)
            for (uint32_t x = 0; x < nw; x++) {
                float mn = inf, mx = ninf;
                for (uint32_t dy = 0; dy < 2u; dy++)
                    for (uint32_t dx = 0; dx < 2u; dx++) {
                        const float2 v = prev[(size_t)min(2u * y + dy, lh - 1u) * lw + min(2u * x + dx, lw - 1u)];
                        mn = fminf(mn, v.x);
                        mx = fmaxf(mx, v.y);
                    }
                next[(size_t)y * nw + x] = make_float2(mn, mx);
            }
        T.plain.push_back(next);
        T.lw.push_back(nw); T.lh.push_back(nh);
    }
    const uint32_t mips = (uint32_t)T.plain.size();
    T.quads.resize(mips);
    T.ppitch.assign(mips, 0u);
    for (uint32_t l = 0; l + 1u < mips; l++) {
        const uint32_t pp = T.lw[l + 1], phh = T.lh[l + 1];
        T.ppitch[l] = pp;
        T.quads[l].assign((size_t)pp * phh * 4u, make_float2(inf, ninf));
        for (uint32_t y = 0; y < 2u * phh; y++)
            for (uint32_t x = 0; x < 2u * pp; x++)
                if (x < T.lw[l] && y < T.lh[l])
                    T.quads[l][((size_t)(y >> 1) * pp + (x >> 1)) * 4u + ((y & 1u) * 2u + (x & 1u))] = T.plain[l][(size_t)y * T.lw[l] + x];
    }
    memset(&T.S, 0, sizeof T.S);
    memset(&T.F, 0, sizeof T.F);
    T.S.ox = origin[0]; T.S.oz = origin[1]; T.S.sx = spacing[0]; T.S.sz = spacing[1];
    T.S.cell_w = cw; T.S.cell_h = ch; T.S.mip_count = mips; T.S.cells = T.cells.data();
    for (uint32_t l = 0; l < mips; l++) { T.S.mm[l] = T.plain[l].data(); T.S.mm_pitch[l] = T.lw[l]; }
    T.S.inv_two_r_prime = k; T.S.curvature_enabled = curv_enabled ? 1u : 0u; T.S.traversal_mode = 3u;
    T.F.ox = origin[0]; T.F.oz = origin[1]; T.F.sx = spacing[0]; T.F.sz = spacing[1];
    T.F.cell_w = cw; T.F.cell_h = ch; T.F.mip_count = mips; T.F.cells = T.cells.data();
    for (uint32_t l = 0; l + 1u < mips; l++) { T.F.q.lv[l] = T.quads[l].data(); T.F.q.parent_pitch[l] = T.ppitch[l]; }
    T.F.root_mm = T.plain[mips - 1u][0];
    T.F.inv_two_r_prime = k;
    fast_scene_finish(T.F);
#ifdef F3D_EMU_FILL_EXTRA
    F3D_EMU_FILL_EXTRA(T.F)
#endif
}

template <bool ANY, bool CURV>
void run_fast(const HostTerrain& T, const float* rays, uint64_t n, uint8_t* hit, float* t, float* nrm, uint64_t* nodes_total) {
    uint32_t stack[kStackSize];
    SmemStack st{stack, 1u};
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; i++) {
        const float* r = rays + 8 * i;
        Ray ray; ray.o = V3(r[0], r[1], r[2]); ray.tmin = r[3]; ray.d = V3(r[4], r[5], r[6]); ray.tmax = r[7];
        uint32_t nodes = 0;
        const FastHit fh = trace_fast<ANY, CURV>(T.F, ray, true, st, nodes);
        total += nodes;
        hit[i] = fh.hit ? 1 : 0;
        t[i] = fh.t;
        v3 p = V3(0, 0, 0), nn = V3(0, 0, 0);
        if (fh.hit) finish_hit(T.F, ray, fh, p, nn);
        nrm[3 * i] = nn.x; nrm[3 * i + 1] = nn.y; nrm[3 * i + 2] = nn.z;
    }
    if (nodes_total) *nodes_total = total;
}

#if F3D_CULL_FAST
// Bottom-up any-hit traversal exactly as k_ascent + k_trace run it (one lane): the ray's own cell, then the seeds of
// ascent_seeds, level-0 seeds as leaves, the others through the stack with the conservative expansion.
template <bool CURV>
void run_bottom_up(const HostTerrain& T, const float* rays, uint64_t n, uint8_t* hit, float* t, uint64_t* nodes_total) {
    uint32_t stack[kStackSize];
    SmemStack st{stack, 1u};
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; i++) {
        const float* r = rays + 8 * i;
        Ray ray; ray.o = V3(r[0], r[1], r[2]); ray.tmin = r[3]; ray.d = V3(r[4], r[5], r[6]); ray.tmax = r[7];
        TraceState S;
        ray_setup<CURV>(T.F, ray, S);
        const uint32_t cell0 = origin_cell(T.F, ray.o);
        const bool asc = ray.d.y >= 0.0f && ray.tmin >= 0.0f;
        bool h = leaf_node<true, CURV>(T.F, S, cell0);
        total++;
        if (!h) {
            unsigned long long seeds = asc ? ascent_seeds<CURV, true>(T.F, S, cell0) : ascent_seeds<CURV, false>(T.F, S, cell0);
            const uint32_t sib0 = cell0 & ~(1u | (1u << 13));
            for (uint32_t q = 0; q < 4u && !h; q++)
                if ((seeds >> q) & 1ull) { total++; h = leaf_node<true, CURV>(T.F, S, sib0 | (q & 1u) | ((q >> 1) << 13)); }
            seeds &= ~15ull;
            S.sp = 0u;
            while (!h && (S.sp > 0u || seeds != 0ull)) {
                uint32_t node;
                if (S.sp > 0u) node = st.at(--S.sp);
                else {
                    const uint32_t b = (uint32_t)__ffsll((long long)seeds) - 1u;
                    seeds &= seeds - 1ull;
                    const uint32_t L = b >> 2, q = b & 3u;
                    node = pack_node(L, (((cell0 & 0x1FFFu) >> (L + 1u)) << 1) | (q & 1u), (((cell0 >> 13) >> (L + 1u)) << 1) | (q >> 1));
                }
                uint32_t bid;
                const uint32_t okm = asc ? expand_core<true, CURV, true>(T.F, S, node, bid) : expand_core<true, CURV, false>(T.F, S, node, bid);
                total++;
                if (((bid >> 26) & 15u) == 0u) {
                    for (uint32_t j = 0; j < 4u && !h; j++)
                        if ((okm >> j) & 1u) { total++; h = leaf_node<true, CURV>(T.F, S, bid ^ (j & 1u) ^ ((j >> 1) << 13)); }
                } else {
                    if (okm & 8u) st.at(S.sp++) = bid ^ (1u | (1u << 13));
                    if (okm & 4u) st.at(S.sp++) = bid ^ (1u << 13);
                    if (okm & 2u) st.at(S.sp++) = bid ^ 1u;
                    if (okm & 1u) st.at(S.sp++) = bid;
                }
            }
        }
        hit[i] = h ? 1 : 0;
        t[i] = S.best_t;
    }
    if (nodes_total) *nodes_total = total;
}
#endif

}  // namespace

#if F3D_CULL_FAST
extern "C" int emu_trace_rays_bottom_up(const float* heights, uint32_t w, uint32_t h, const float spacing[2], const float origin[2],
                                        float exaggeration, float inv_two_r_prime, int32_t curvature_enabled, const float* rays,
                                        uint64_t n, int32_t apply_curvature, uint8_t* hit, float* t, uint64_t* nodes) {
    if (w < 2u || h < 2u) return 1;
    HostTerrain T;
    build(T, heights, w, h, spacing, origin, exaggeration, inv_two_r_prime, curvature_enabled);
    if (apply_curvature && curvature_enabled) run_bottom_up<true>(T, rays, n, hit, t, nodes);
    else run_bottom_up<false>(T, rays, n, hit, t, nodes);
    return 0;
}
#endif

extern "C" int emu_trace_rays(const float* heights, uint32_t w, uint32_t h, const float spacing[2], const float origin[2],
                              float exaggeration, float inv_two_r_prime, int32_t curvature_enabled, const float* rays,
                              uint64_t n, int32_t any_hit, int32_t apply_curvature, uint8_t* hit, float* t, float* normal,
                              uint64_t* nodes) {
    if (w < 2u || h < 2u) return 1;
    HostTerrain T;
    build(T, heights, w, h, spacing, origin, exaggeration, inv_two_r_prime, curvature_enabled);
    const bool curv = apply_curvature && curvature_enabled;
    if (any_hit) { if (curv) run_fast<true, true>(T, rays, n, hit, t, normal, nodes); else run_fast<true, false>(T, rays, n, hit, t, normal, nodes); }
    else         { if (curv) run_fast<false, true>(T, rays, n, hit, t, normal, nodes); else run_fast<false, false>(T, rays, n, hit, t, normal, nodes); }
    return 0;
}
