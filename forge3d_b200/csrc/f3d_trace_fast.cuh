// forge3d_b200/csrc/f3d_trace_fast.cuh
// Production traversal of the min-max quadtree: same visit order, same arithmetic and therefore the
// same hits (bit for bit) as the literal loop in f3d_trace.cuh / the WGSL
// (/root/reference/src/shaders/hybrid_terrain_traversal.wgsl:254-372), restructured for the SM:
//
//  * CULL AT PUSH.  The WGSL pushes every child whose clipped span is non-empty and tests its
//    height band only when it is popped (one dependent RG32F load per pop, :301).  Here the four
//    children of a node are tested when the node is expanded: their [min,max] pairs sit in ONE
//    32-byte quad (two 128-bit loads, one sector), and the pop-time tests
//        t_lo = max(c0, tmin), t_hi = min(c1, min(tmax, best_t)), band(t_lo, t_hi)
//    are evaluated right there on the same operands the pop would use.  A child that fails is never
//    pushed.  This is exact as long as best_t is unchanged between push and pop.  any-hit rays
//    return at their first hit, so for them it always is.  A closest-hit ray marks every entry
//    already on the stack as STALE when it records a hit; stale entries are re-tested at pop with
//    the new best_t (their own [min,max] is re-read), exactly what the WGSL does for every entry.
//    (Closest-hit rays never carry the curvature term on this path -- hybrid_traversal.wgsl:248-259,
//    hybrid_terrain_traversal.wgsl:374-376 -- so the production closest-hit instance is CURV=false.)
//  * Each node's six slab planes (x0, xm, x1, z0, zm, z1) are computed once; the 4 children's
//    spans are min/max combinations of those six ray parameters -- the identical float
//    expressions the WGSL evaluates 16 times (4 children x 4 planes).
//  * The stack holds 32-bit node ids in SHARED memory, laid out [depth][thread] (conflict-free),
//    not in a per-thread local array.
//  * The traversal is a resumable state machine (TraceState + expand_top / leaf_top) so that the
//    persistent k_trace kernel can refill idle lanes with new rays between steps.
//  * The hit point / normal are computed after the loop from the winning cell (same expressions).
#pragma once
#include "f3d_trace.cuh"

// Compile-time variants of the traversal.  Every variant must leave the outputs bit-identical to the oracle: checked
// on the CPU by tests/test_traversal_emulation.py (this header compiled for the host) before it is measured on the GPU.
//
// F3D_ANYHIT_SIGN_ORDER = 1 (default; 0 restores the reference's sort for any-hit rays too): any-hit rays push the surviving children far-to-near by the SIGNS of the ray direction
//   instead of sorting them by entry parameter.  The occlusion flag cannot depend on the visit order: until the first
//   hit an any-hit ray never changes res.t (hybrid_terrain_traversal.wgsl:297,306-314), so the set of nodes that pass
//   the span / band tests is a fixed tree and the flag is the OR of the leaf tests over that tree.  (Which hit is
//   found FIRST - its t - may differ in exact ties of the sort key; no caller reads the t of an any-hit ray:
//   hybrid_traversal.wgsl:248-259 compare it with 1e30 only.)  The sign order is a valid front-to-back order of a
//   quadtree's children, so early exit is as early as with the sort.  Closest-hit rays keep the reference order.
#ifndef F3D_ANYHIT_SIGN_ORDER
#define F3D_ANYHIT_SIGN_ORDER 1
#endif
// F3D_PUSH_CLIP_FOLDED = 1 (default; 0 evaluates the WGSL's push test literally): the push test of :342-344 clips a
//   child's span [c0, c1] to the PARENT's clipped span, ct = [max(c0, t_lo), min(c1, t_hi)], the pop test of :295-297
//   clips it to the ray, tl/th = [max(c0, tmin), min(c1, min(tmax, best_t))].  They are the same two numbers:
//   t_lo = max(n0, tmin) and t_hi = min(n1, min(tmax, best_t)) with [n0, n1] the parent's slab span, and c0 >= n0,
//   c1 <= n1 hold EXACTLY in f32 because every slab plane is the same monotone expression
//   ((origin + f32(cell) * spacing) - o) * inv of its integer cell coordinate (each rounding step preserves weak
//   order) and the child's planes lie between the parent's.  Hence max(c0, t_lo) = max(c0, tmin) and
//   min(c1, t_hi) = min(c1, tcap): one clipped span serves as push test, pop test and sort key, and the parent's own
//   span is only needed when a stale closest-hit entry is re-tested.
#ifndef F3D_PUSH_CLIP_FOLDED
#define F3D_PUSH_CLIP_FOLDED 1
#endif

// F3D_CULL_FAST = 1 (default; 0 restores the exact cull-at-push expansion of round 1): the hierarchy is walked with a CHEAP
//   CONSERVATIVE test and only the leaves run the reference's arithmetic.  Why the outputs cannot change:
//   (1) A leaf (level-0 node) is solved by the reference iff it is popped, and it is popped iff every ancestor passed its
//       own pop tests.  Those tests are implied by the leaf's own: a child's clipped span lies inside its parent's
//       EXACTLY in f32 (see F3D_PUSH_CLIP_FOLDED), a child's [min,max] lies inside its parent's, and the ray height
//       o.y + t*d.y (+ t*t*hd2*k for an ascending curved ray) is a monotone f32 expression of t, so its range over the
//       child's span lies inside its range over the parent's.  Hence: the set of leaves the reference solves is exactly the
//       set of cells that pass their OWN span + band test (with the best_t current at that moment).
//   (2) leaf_node() evaluates that own test and the patch solve with the reference's arithmetic (it recomputes the
//       cell's slab span from the integer cell coordinates, as terrain_trace does at every pop).  Any traversal that
//       visits a SUPERSET of those cells therefore reproduces an any-hit ray's flag (an OR over a fixed set).
//   (3) A closest-hit ray also needs the leaves in the reference's order, because a hit shrinks best_t and best_t clips the
//       span the next solve samples.  The reference pops children by ascending entry parameter.  The exact spans of the
//       four children partition the parent's span, so the children with a non-degenerate span are strictly ordered along
//       the ray: near, (at most one) side, far == the order given by the signs of the direction.  A child with an empty
//       span is never solved; one with a single-point span samples the patch three times at the same t, gets a = b = 0
//       and cannot produce a closest hit (leaf_intersect :196-204), so where it sits in the order is immaterial.
//   The conservative test: plane parameters as ONE fma each, t(c) = fma(f32(c), sx*inv, (ox-o)*inv), widened by a per-ray
//   bound on the difference to the reference's four-rounding chain (ex, ez below); ray heights as fma chains widened by a
//   per-ray bound (ey).  Children that cannot exist hold the (+inf,-inf) sentinel and fail the band test by themselves.
//   Curved rays that DESCEND (sun below the horizon, d.y < 0) are not monotone; the host keeps the round-1 expansion for
//   them (f3d_backend.cu picks the kernel instance), so (1) never rests on the vertex term.
#ifndef F3D_CULL_FAST
#define F3D_CULL_FAST 1
#endif

namespace f3d {

struct FastHit { float t; uint32_t cx, cz; bool hit; };

// Shared-memory stack accessor: entry `depth` of this thread at base[depth * stride].
struct SmemStack {
    uint32_t* base;      // already offset to this thread's column
    uint32_t stride;     // threads per CTA
    __device__ __forceinline__ uint32_t& at(uint32_t depth) const { return base[depth * stride]; }
};

// Quad-packed min-max level: the four children of parent (px, py) are contiguous.
//   index = ((py * parent_pitch + px) * 4 + (cy * 2 + cx))
struct QuadLevels {
    const float2* lv[kMaxLevels];   // lv[l] holds level l nodes grouped by their level l+1 parent
    uint32_t parent_pitch[kMaxLevels];
};

struct FastScene {
    float ox, oz, sx, sz;
    uint32_t cell_w, cell_h, mip_count;
    const float4* cells;
    QuadLevels q;
    float2 root_mm;
    float inv_two_r_prime;
    // F3D_CULL_FAST: magnitudes that scale the per-ray padding (fast_scene_finish)
    float mag_x, mag_z;      // |ox| + cell_w*sx, |oz| + cell_h*sz: bound on every plane coordinate
    float mag_y;             // max(|root min|, |root max|)
};

// Fills the padding magnitudes from the fields above (host side; also used by the CPU emulation of this header).
inline void fast_scene_finish(FastScene& F) {
    F.mag_x = fabsf(F.ox) + (float)F.cell_w * F.sx;
    F.mag_z = fabsf(F.oz) + (float)F.cell_h * F.sz;
    const float a = fabsf(F.root_mm.x), b = fabsf(F.root_mm.y);
    F.mag_y = (a > b ? a : b);
    if (!(F.mag_y < 1e30f)) F.mag_y = 1e30f;     // flat (+inf,-inf) roots cannot occur (>= 1 cell), keep it finite anyway
}

// Per-ray traversal state.
struct TraceState {
    v3 o, d;
    float tmin, tmax;
    float inv_x, inv_z;      // terrain_safe_inv of the direction (:88-91)
    float hd2;               // dot(dir.xz, dir.xz)
    float vertex;            // parameter of the curved ray's lowest point (:120), CURV only
    float y_vertex;          // terrain_curved_height(ray, vertex, true) (:122), constant per ray
    bool use_vertex;         // CURV && a > 0 (:119)
    float best_t;            // res.t
    uint32_t best_cx, best_cz;
    bool hit;
    uint32_t sp;             // entries on the stack
    uint32_t stale_sp;       // closest-hit only: entries below this index predate the last hit
#if F3D_CULL_FAST
    // conservative culling (see F3D_CULL_FAST): t(c) = fma(f32(c), kx, bx) approximates the reference's plane parameter
    // within ex; heights fma(t*t, kc, fma(t, d.y, o.y)) approximate ray_height within ey for |height| <= mag_y.
    float kx, bx, kz, bz, ex, ez, ey, kc;
#endif
};

template <bool CURV>
__device__ __forceinline__ float ray_height(const FastScene& S, const TraceState& T, float t) {   // :95-103
    float corr = 0.0f;
    if (CURV) corr = (t * t * T.hd2) * S.inv_two_r_prime;
    return T.o.y + t * T.d.y + corr;
}

template <bool CURV>
__device__ __forceinline__ bool band_ok(const FastScene& S, const TraceState& T, float tl, float th, float2 mm) {  // :303-304, :108-127
    const float y0 = ray_height<CURV>(S, T, tl), y1 = ray_height<CURV>(S, T, th);
    float lo = fminf(y0, y1);
    if (CURV) {
        if (T.use_vertex && T.vertex >= tl && T.vertex <= th) lo = fminf(lo, T.y_vertex);
    }
    const float hi = fmaxf(y0, y1);
    return !(lo > mm.y || hi < mm.x);
}

#if F3D_CULL_FAST
// Per-ray constants of the conservative tests (see F3D_CULL_FAST).  Needs T.o, T.d, T.inv_x/z, T.hd2.
// Reference chain per plane: fl(fl(fl(ox + fl(c*sx)) - o) * inv): four roundings; here one fma over two rounded
// constants.  With u = 2^-24 and M = |ox| + cell_w*sx the two differ by at most |inv| * u * (8 M + 5 |o|)
// (absolute terms 2uM + u|X-o| from the chain, u(3M + 2|o|) from kx/bx, 3u|t| with |t| <= |inv| (M + |o|));
// the pad below is 4x that.  The height pad covers both evaluations of a height of magnitude <= mag_y
// (relative error 3u each on |o.y| + |t d.y| + corr, with |t d.y| <= |y| + |o.y| + corr) with the same margin.
template <bool CURV>
__device__ __forceinline__ void cull_setup(const FastScene& S, TraceState& T) {
    T.kx = S.sx * T.inv_x; T.bx = (S.ox - T.o.x) * T.inv_x;
    T.kz = S.sz * T.inv_z; T.bz = (S.oz - T.o.z) * T.inv_z;
    const float rx = S.mag_x + fabsf(T.o.x), rz = S.mag_z + fabsf(T.o.z);
    T.ex = fabsf(T.inv_x) * rx * 1.9073486328125e-6f;      // 2^-19
    T.ez = fabsf(T.inv_z) * rz * 1.9073486328125e-6f;
    T.kc = CURV ? T.hd2 * S.inv_two_r_prime : 0.0f;
    const float reach = 2.0f * (rx + rz);                  // bound on the horizontal distance travelled inside the DEM
    const float corr_max = CURV ? fabsf(S.inv_two_r_prime) * reach * reach : 0.0f;
    T.ey = (fabsf(T.o.y) + S.mag_y + corr_max) * 3.814697265625e-6f;   // 2^-18
}
#endif

// Sets up a ray and performs the root's pop-time tests (:288-304); leaves the root on the stack if it
// survives.
template <bool CURV>
__device__ __forceinline__ void ray_setup(const FastScene& S, const Ray& r, TraceState& T) {
    T.o = r.o; T.d = r.d; T.tmin = r.tmin; T.tmax = r.tmax;
    T.inv_x = safe_inv(r.d.x); T.inv_z = safe_inv(r.d.z);
    T.hd2 = dot2(r.d.x, r.d.z, r.d.x, r.d.z);
    T.use_vertex = false; T.vertex = 0.0f; T.y_vertex = 0.0f;
    if (CURV) {
        const float a = T.hd2 * S.inv_two_r_prime;
        T.use_vertex = a > 0.0f;
        if (T.use_vertex) { T.vertex = fdiv(-r.d.y, 2.0f * a); T.y_vertex = ray_height<true>(S, T, T.vertex); }
    }
    T.best_t = r.tmax; T.best_cx = 0u; T.best_cz = 0u; T.hit = false;
    T.sp = 0u; T.stale_sp = 0u;
#if F3D_CULL_FAST
    cull_setup<CURV>(S, T);
#endif
}

// The part of ray_setup a patch solve reads (leaf_node): no culling constants.
template <bool CURV>
__device__ __forceinline__ void leaf_ray_setup(const FastScene& S, const Ray& r, TraceState& T) {
    T.o = r.o; T.d = r.d; T.tmin = r.tmin; T.tmax = r.tmax;
    T.inv_x = safe_inv(r.d.x); T.inv_z = safe_inv(r.d.z);
    T.hd2 = dot2(r.d.x, r.d.z, r.d.x, r.d.z);
    T.use_vertex = false; T.vertex = 0.0f; T.y_vertex = 0.0f;
    if (CURV) {
        const float a = T.hd2 * S.inv_two_r_prime;
        T.use_vertex = a > 0.0f;
        if (T.use_vertex) { T.vertex = fdiv(-r.d.y, 2.0f * a); T.y_vertex = ray_height<true>(S, T, T.vertex); }
    }
    T.best_t = r.tmax; T.best_cx = 0u; T.best_cz = 0u; T.hit = false;
    T.sp = 0u; T.stale_sp = 0u;
}

template <bool CURV>
__device__ __forceinline__ void trace_begin(const FastScene& S, const Ray& r, TraceState& T, const SmemStack st) {
    ray_setup<CURV>(S, r, T);
    const uint32_t level = S.mip_count - 1u;
    const uint32_t cx1 = min(1u << level, S.cell_w), cz1 = min(1u << level, S.cell_h);
    // ox + f32(0) * sx == ox exactly
    const float a0 = (S.ox - T.o.x) * T.inv_x, a1 = ((S.ox + (float)cx1 * S.sx) - T.o.x) * T.inv_x;
    const float b0 = (S.oz - T.o.z) * T.inv_z, b1 = ((S.oz + (float)cz1 * S.sz) - T.o.z) * T.inv_z;
    const float s0 = fmaxf(fminf(a0, a1), fminf(b0, b1)), s1 = fminf(fmaxf(a0, a1), fmaxf(b0, b1));
    const float tl = fmaxf(s0, T.tmin), th = fminf(s1, fminf(T.tmax, T.best_t));
    if (tl <= th && band_ok<CURV>(S, T, tl, th, S.root_mm)) { st.at(0) = pack_node(level, 0u, 0u); T.sp = 1u; }
}

// Precondition: T.sp > 0.
__device__ __forceinline__ bool top_is_leaf(const TraceState& T, const SmemStack st) {
    return (st.at(T.sp - 1u) >> 26) == 0u;
}

// One sorting-network comparator for the (t desc, original index asc) total order that the
// WGSL's stable insertion sort (:351-363) realises.
__device__ __forceinline__ void cmpswap(float& ta, uint32_t& ia, float& tb, uint32_t& ib) {
    const bool sw = (tb > ta) || (tb == ta && ib < ia);
    const float t0 = sw ? tb : ta, t1 = sw ? ta : tb;
    const uint32_t i0 = sw ? ib : ia, i1 = sw ? ia : ib;
    ta = t0; tb = t1; ia = i0; ib = i1;
}

// Pops the internal node on top of the stack and pushes its surviving children (:320-369 with the
// children's pop tests of :281-304 folded in).  Precondition: sp > 0 and the top is not a leaf.
template <bool ANY_HIT, bool CURV>
__device__ __forceinline__ void expand_top(const FastScene& S, TraceState& T, const SmemStack st) {
    T.sp--;
    const uint32_t node = st.at(T.sp);
    const uint32_t level = node >> 26, ny = (node >> 13) & 0x1FFFu, nx = node & 0x1FFFu;
    const uint32_t cl = level - 1u;
    const uint32_t cell_w = S.cell_w, cell_h = S.cell_h;
    // integer cell planes of this node and its mid-lines (the children's gx0/gx1, :330-334)
    const uint32_t cx0 = nx << level, cz0 = ny << level;
    const uint32_t cx1 = min((nx + 1u) << level, cell_w), cz1 = min((ny + 1u) << level, cell_h);
    const uint32_t mxu = (2u * nx + 1u) << cl, mzu = (2u * ny + 1u) << cl;
    const bool has_x1 = mxu < cell_w, has_z1 = mzu < cell_h;       // child column / row 1 exists (:332)
    const uint32_t cxm = min(mxu, cell_w), czm = min(mzu, cell_h);
    // children's [min,max]: one 32-byte quad, issued before the arithmetic that hides its latency
    const float4* qp = reinterpret_cast<const float4*>(S.q.lv[cl] + ((size_t)ny * S.q.parent_pitch[cl] + nx) * 4u);
    const float4 q01 = __ldg(qp), q23 = __ldg(qp + 1);
    // ray parameters at the six planes
    const float a0 = ((S.ox + (float)cx0 * S.sx) - T.o.x) * T.inv_x;
    const float am = ((S.ox + (float)cxm * S.sx) - T.o.x) * T.inv_x;
    const float a1 = ((S.ox + (float)cx1 * S.sx) - T.o.x) * T.inv_x;
    const float b0 = ((S.oz + (float)cz0 * S.sz) - T.o.z) * T.inv_z;
    const float bm = ((S.oz + (float)czm * S.sz) - T.o.z) * T.inv_z;
    const float b1 = ((S.oz + (float)cz1 * S.sz) - T.o.z) * T.inv_z;
    // this node's own clipped span (:288-297); needed to clip the children (:342-343)
    const float tcap = fminf(T.tmax, T.best_t);
#if !F3D_PUSH_CLIP_FOLDED
    const float t_lo = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), T.tmin);
    const float t_hi = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), tcap);
#endif
    if (!ANY_HIT) {
        if (T.sp < T.stale_sp) {              // entry predates the last hit: redo the pop tests (:297-304)
            T.stale_sp = T.sp;
#if F3D_PUSH_CLIP_FOLDED
            const float t_lo = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), T.tmin);
            const float t_hi = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), tcap);
#endif
            if (t_lo > t_hi) return;
            float2 mm;
            if (level == S.mip_count - 1u) mm = S.root_mm;
            else mm = __ldg(S.q.lv[level] + (((size_t)(ny >> 1) * S.q.parent_pitch[level] + (nx >> 1)) * 4u + ((ny & 1u) * 2u + (nx & 1u))));
            if (!band_ok<CURV>(S, T, t_lo, t_hi, mm)) return;
        }
    }
    // per-axis child spans
    const float xlo0 = fminf(a0, am), xhi0 = fmaxf(a0, am), xlo1 = fminf(am, a1), xhi1 = fmaxf(am, a1);
    const float zlo0 = fminf(b0, bm), zhi0 = fmaxf(b0, bm), zlo1 = fminf(bm, b1), zhi1 = fmaxf(bm, b1);
    float kt[4];
    bool ok[4];
#pragma unroll
    for (uint32_t c = 0; c < 4u; c++) {
        const uint32_t cxi = c & 1u, cy = c >> 1;
        const float c0 = fmaxf(cxi ? xlo1 : xlo0, cy ? zlo1 : zlo0);
        const float c1 = fminf(cxi ? xhi1 : xhi0, cy ? zhi1 : zhi0);
        const float tl = fmaxf(c0, T.tmin), th = fminf(c1, tcap);         // the child's own pop span (:295-297)
#if F3D_PUSH_CLIP_FOLDED
        const float ct_lo = tl, ct_hi = th;                               // == the push test's span, see above
#else
        const float ct_lo = fmaxf(c0, t_lo), ct_hi = fminf(c1, t_hi);     // push test (:342-344)
#endif
        const float2 mm = c == 0u ? make_float2(q01.x, q01.y) : c == 1u ? make_float2(q01.z, q01.w)
                        : c == 2u ? make_float2(q23.x, q23.y) : make_float2(q23.z, q23.w);
        const bool exists = (cxi ? has_x1 : true) && (cy ? has_z1 : true);
        const bool band = band_ok<CURV>(S, T, tl, th, mm);                 // evaluated unconditionally: no divergence
        const bool v = exists & (ct_lo <= ct_hi) & (tl <= th) & band;
        ok[c] = v;
        kt[c] = v ? ct_lo : __int_as_float(0x7f800000);   // rejected children sort to the front, never pushed
    }
#if F3D_ANYHIT_SIGN_ORDER
    if (ANY_HIT) {
        // nearest child = the one on the side the ray comes from; push far first so that it is popped first.
        // Child j of the sign-mirrored node is child j ^ flip of the real one; node ids of the four children differ
        // from base_id (even x, even y) only in the low bit of the x and y fields, so mirroring is one XOR.
        const uint32_t fx = T.d.x < 0.0f ? 1u : 0u, fz = T.d.z < 0.0f ? 1u : 0u;
        uint32_t okm = (ok[0] ? 1u : 0u) | (ok[1] ? 2u : 0u) | (ok[2] ? 4u : 0u) | (ok[3] ? 8u : 0u);
        if (fx) okm = ((okm & 0x5u) << 1) | ((okm >> 1) & 0x5u);     // swap columns
        if (fz) okm = ((okm & 0x3u) << 2) | (okm >> 2);              // swap rows
        const uint32_t bid = pack_node(cl, nx * 2u, ny * 2u) ^ (fx | (fz << 13));
        if (okm & 8u) { st.at(T.sp) = bid ^ (1u | (1u << 13)); T.sp++; }
        if (okm & 4u) { st.at(T.sp) = bid ^ (1u << 13); T.sp++; }
        if (okm & 2u) { st.at(T.sp) = bid ^ 1u; T.sp++; }
        if (okm & 1u) { st.at(T.sp) = bid; T.sp++; }
        return;
    }
#endif
    // order: descending t_enter, ties by original child index (== the stable insertion sort)
    float t0 = kt[0], t1 = kt[1], t2 = kt[2], t3 = kt[3];
    uint32_t i0 = 0u, i1 = 1u, i2 = 2u, i3 = 3u;
    cmpswap(t0, i0, t1, i1); cmpswap(t2, i2, t3, i3);
    cmpswap(t0, i0, t2, i2); cmpswap(t1, i1, t3, i3);
    cmpswap(t1, i1, t2, i2);
    const uint32_t okmask = (ok[0] ? 1u : 0u) | (ok[1] ? 2u : 0u) | (ok[2] ? 4u : 0u) | (ok[3] ? 8u : 0u);
    const uint32_t base_id = pack_node(cl, nx * 2u, ny * 2u);
    const uint32_t ord[4] = {i0, i1, i2, i3};
#pragma unroll
    for (uint32_t k = 0; k < 4u; k++) {
        const uint32_t c = ord[k];
        if ((okmask >> c) & 1u) {
            st.at(T.sp) = base_id + (c & 1u) + ((c >> 1) << 13);
            T.sp++;
        }
    }
}

#if F3D_CULL_FAST
// Conservative test of the four children of an internal node (see F3D_CULL_FAST): ~3x fewer instructions than expand_top
// (8 fma for the six planes, no sort, no integer clamps: a far plane beyond the DEM only widens a span).  Children are
// numbered in the order given by the signs of the direction: bit0 = far half in x, bit1 = far half in z; child j is the
// real child j ^ flip.  Returns the mask of children that MAY pass the reference's pop tests; `bid` receives the node id
// of child 0 (the id of child j is bid ^ (j & 1) ^ ((j >> 1) << 13)).
// ASC = the ray height is non-decreasing in t (d.y >= 0: an ascending ray, curved or not): the range over a span is
// [y(t_lo), y(t_hi)], no min/max and no vertex term.
template <bool ANY_HIT, bool CURV, bool ASC>
__device__ __forceinline__ uint32_t expand_core(const FastScene& S, const TraceState& T, const uint32_t node, uint32_t& bid) {
    const uint32_t level = (node >> 26) & 15u, ny = (node >> 13) & 0x1FFFu, nx = node & 0x1FFFu;
    const uint32_t cl = level - 1u;
    const bool fx = T.inv_x < 0.0f, fz = T.inv_z < 0.0f;
    const uint32_t flip = (fx ? 1u : 0u) | (fz ? 2u : 0u);
    // the children's [min,max]: four 8-byte loads from one 32-byte sector, already in sign order
    const float2* qp = S.q.lv[cl] + (ny * S.q.parent_pitch[cl] + nx) * 4u;
    const float2 m0 = __ldg(qp + flip), m1 = __ldg(qp + (flip ^ 1u)), m2 = __ldg(qp + (flip ^ 2u)), m3 = __ldg(qp + (flip ^ 3u));
    // cell coordinates of the three planes per axis, exact in f32 (integers <= 2^14)
    const float hs = __uint_as_float((127u + cl) << 23);                    // 2^cl
    const float fx0 = (float)(nx << level), fz0 = (float)(ny << level);
    const float fxm = fx0 + hs, fzm = fz0 + hs, fx1 = fxm + hs, fz1 = fzm + hs;
    const float A0 = __fmaf_rn(fx0, T.kx, T.bx), Am = __fmaf_rn(fxm, T.kx, T.bx), A1 = __fmaf_rn(fx1, T.kx, T.bx);
    const float B0 = __fmaf_rn(fz0, T.kz, T.bz), Bm = __fmaf_rn(fzm, T.kz, T.bz), B1 = __fmaf_rn(fz1, T.kz, T.bz);
    // near / far half per axis, widened
    const float xl0 = fminf(A0, A1) - T.ex, xh0 = Am + T.ex, xl1 = Am - T.ex, xh1 = fmaxf(A0, A1) + T.ex;
    const float zl0 = fminf(B0, B1) - T.ez, zh0 = Bm + T.ez, zl1 = Bm - T.ez, zh1 = fmaxf(B0, B1) + T.ez;
    const float tcap = fminf(T.tmax, T.best_t);
    bid = pack_node(cl, nx * 2u, ny * 2u) ^ ((fx ? 1u : 0u) | (fz ? (1u << 13) : 0u));
    if (!ANY_HIT) {
        if (fmaxf(xl0, zl0) > tcap) return 0u;        // whole node beyond the closest hit found since it was pushed
    }
    const float oy_lo = T.o.y - T.ey, oy_hi = T.o.y + T.ey;
    uint32_t okm = 0u;
#pragma unroll
    for (uint32_t j = 0; j < 4u; j++) {
        const float tl = fmaxf(fmaxf((j & 1u) ? xl1 : xl0, (j & 2u) ? zl1 : zl0), T.tmin);
        const float th = fminf(fminf((j & 1u) ? xh1 : xh0, (j & 2u) ? zh1 : zh0), tcap);
        const float2 mm = j == 0u ? m0 : j == 1u ? m1 : j == 2u ? m2 : m3;
        bool ok;
        if (ASC) {
            float lo = __fmaf_rn(tl, T.d.y, oy_lo), hi = __fmaf_rn(th, T.d.y, oy_hi);
            if (CURV) { lo = __fmaf_rn(tl * tl, T.kc, lo); hi = __fmaf_rn(th * th, T.kc, hi); }
            ok = (tl <= th) & !(lo > mm.y || hi < mm.x);
        } else {
            float y0 = __fmaf_rn(tl, T.d.y, T.o.y), y1 = __fmaf_rn(th, T.d.y, T.o.y);
            if (CURV) { y0 = __fmaf_rn(tl * tl, T.kc, y0); y1 = __fmaf_rn(th * th, T.kc, y1); }
            float lo = fminf(y0, y1) - T.ey;
            const float hi = fmaxf(y0, y1) + T.ey;
            if (CURV) {
                if (T.use_vertex && T.vertex >= tl && T.vertex <= th) lo = fminf(lo, T.y_vertex - T.ey);
            }
            ok = (tl <= th) & !(lo > mm.y || hi < mm.x);
        }
        okm |= ok ? (1u << j) : 0u;
    }
    return okm;
}

// Stack form: pops the internal node on top of the stack and pushes, far child first, the children that may pass.
// Precondition: sp > 0 and the top is not a leaf.
template <bool ANY_HIT, bool CURV>
__device__ __forceinline__ void expand_cull(const FastScene& S, TraceState& T, const SmemStack st) {
    T.sp--;
    uint32_t bid;
    const uint32_t okm = expand_core<ANY_HIT, CURV, false>(S, T, st.at(T.sp), bid);
    if (okm & 8u) { st.at(T.sp) = bid ^ (1u | (1u << 13)); T.sp++; }
    if (okm & 4u) { st.at(T.sp) = bid ^ (1u << 13); T.sp++; }
    if (okm & 2u) { st.at(T.sp) = bid ^ 1u; T.sp++; }
    if (okm & 1u) { st.at(T.sp) = bid; T.sp++; }
}

// ---------------------------------------------------------------------------------------------------------------
// Bottom-up start for rays that begin ON the terrain (every sun / IBL ray).  A top-down descent spends one expansion per
// level (11 for a 2048^2 DEM) just to reach the cell the ray starts in, inside the divergent traversal loop.  Instead, let
// A_0 be a cell, A_1 .. A_top its ancestors.  The DEM is the disjoint union of A_0 and, for every level L < top, the three
// siblings of A_L inside A_{L+1}.  ascent_seeds tests those siblings level by level with the conservative child test
// (expand_core) - a fixed-trip loop with independent loads, run by k_ascent at full lane occupancy - and returns the
// survivors as SEEDS: 4 bits per level, bit 4L + r = the level-L child r (real index: bit0 = x, bit1 = z) of A_{L+1}
// may pass its pop tests.  The tracer solves A_0 and the level-0 seeds as leaves and runs its stack traversal from the
// other seeds.  Coverage: every cell is A_0 or lies below exactly one sibling; a sibling is dropped only by the same
// conservative test an expansion of A_{L+1} would apply, or because it lies BEHIND the ray: if the ray starts strictly
// inside A_0's slabs (widened entry parameter < tmin; then also inside every A_L's, the near planes only move back), a
// sibling in the half A_L's near plane cuts off ends where A_L begins, before tmin, and its clipped span is empty.
// Any choice of A_0 is valid; if the ray does not start inside it, all three siblings of every level are tested.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t origin_cell(const FastScene& S, v3 o) {          // (cz << 13) | cx, clamped into the DEM
    const float fx = floorf(fdiv(o.x - S.ox, S.sx)), fz = floorf(fdiv(o.z - S.oz, S.sz));
    const uint32_t cx = fx > 0.0f ? min((uint32_t)fminf(fx, 16383.0f), S.cell_w - 1u) : 0u;    // NaN -> 0
    const uint32_t cz = fz > 0.0f ? min((uint32_t)fminf(fz, 16383.0f), S.cell_h - 1u) : 0u;
    return (cz << 13) | cx;
}

// Conservative span + band test of one sibling (see expand_core; same widening).
template <bool CURV, bool ASC>
__device__ __forceinline__ bool sibling_may_pass(const TraceState& T, float tl, float th, float2 mm) {
    if (ASC) {
        float lo = __fmaf_rn(tl, T.d.y, T.o.y - T.ey), hi = __fmaf_rn(th, T.d.y, T.o.y + T.ey);
        if (CURV) { lo = __fmaf_rn(tl * tl, T.kc, lo); hi = __fmaf_rn(th * th, T.kc, hi); }
        return (tl <= th) & !(lo > mm.y || hi < mm.x);
    }
    float y0 = __fmaf_rn(tl, T.d.y, T.o.y), y1 = __fmaf_rn(th, T.d.y, T.o.y);
    if (CURV) { y0 = __fmaf_rn(tl * tl, T.kc, y0); y1 = __fmaf_rn(th * th, T.kc, y1); }
    float lo = fminf(y0, y1) - T.ey;
    const float hi = fmaxf(y0, y1) + T.ey;
    if (CURV) {
        if (T.use_vertex && T.vertex >= tl && T.vertex <= th) lo = fminf(lo, T.y_vertex - T.ey);
    }
    return (tl <= th) & !(lo > mm.y || hi < mm.x);
}

// Needs ray_setup() done on T (tmin/tmax as the tracer will use them, or looser).
// Going up one level adds ONE new plane per axis (A_{L+1} extends A_L on its low or its high side), and only an
// extension on the side the ray travels to can be met: per level 2 fma for the new far planes, three 8-byte loads and up
// to three sibling tests whose spans follow from the far planes alone:
//   x-sibling  [far_x(L), min(far_x(L+1), far_z(L))]      (shares A_L's z-slab; entered through A_L's far x plane)
//   z-sibling  [far_z(L), min(far_z(L+1), far_x(L))]
//   diagonal   [max(far_x(L), far_z(L)), min(far_x(L+1), far_z(L+1))]
// each bound widened by the per-axis pad, clipped to [tmin, min(tmax, best_t)].
template <bool CURV, bool ASC>
__device__ __forceinline__ unsigned long long ascent_seeds(const FastScene& S, const TraceState& T, const uint32_t cell0) {
    const uint32_t cx0 = cell0 & 0x1FFFu, cz0 = cell0 >> 13;
    const uint32_t top = S.mip_count - 1u;
    const bool sgx = T.inv_x < 0.0f, sgz = T.inv_z < 0.0f;
    float lox = (float)cx0, hix = lox + 1.0f, loz = (float)cz0, hiz = loz + 1.0f;       // A_L in cell units, exact in f32
    const float txl = __fmaf_rn(lox, T.kx, T.bx), txh = __fmaf_rn(hix, T.kx, T.bx);
    const float tzl = __fmaf_rn(loz, T.kz, T.bz), tzh = __fmaf_rn(hiz, T.kz, T.bz);
    float fx = sgx ? txl : txh, fz = sgz ? tzl : tzh;                                   // far planes of A_L
    const bool inside = fmaxf((sgx ? txh : txl) + T.ex, (sgz ? tzh : tzl) + T.ez) < T.tmin;
    unsigned long long seeds = 0ull;
    if (!inside) {                                   // rare (origin within the pad of a cell border, or outside the DEM):
        for (uint32_t L = 0; L < top; L++) {         // every sibling that EXISTS is a seed, nothing is tested
            const uint32_t own = (((cz0 >> L) & 1u) << 1) | ((cx0 >> L) & 1u);
            for (uint32_t r = 0; r < 4u; r++) {
                const uint32_t gx = ((((cx0 >> (L + 1u)) << 1) | (r & 1u)) << L), gz = ((((cz0 >> (L + 1u)) << 1) | (r >> 1)) << L);
                if (r != own && gx < S.cell_w && gz < S.cell_h) seeds |= 1ull << (4u * L + r);      // :336 gx0 >= cell_w: no such child
            }
        }
        return seeds;
    }
    const float tcap = fminf(T.tmax, T.best_t);
    float size = 1.0f;
    for (uint32_t L = 0; L < top; L++) {
        const uint32_t bx = (cx0 >> L) & 1u, bz = (cz0 >> L) & 1u, own = (bz << 1) | bx;
        const float2* qp = S.q.lv[L] + ((cz0 >> (L + 1u)) * S.q.parent_pitch[L] + (cx0 >> (L + 1u))) * 4u;
        const float2 mx = __ldg(qp + (own ^ 1u)), mz = __ldg(qp + (own ^ 2u)), md = __ldg(qp + (own ^ 3u));
        // the new plane of A_{L+1} per axis; it is a FAR plane iff the extension is on the side the ray travels to
        const float nx = bx ? lox - size : hix + size, nz = bz ? loz - size : hiz + size;
        if (bx) lox = nx; else hix = nx;
        if (bz) loz = nz; else hiz = nz;
        const bool ax = (bx != 0u) == sgx, az = (bz != 0u) == sgz;
        const float fx1 = ax ? __fmaf_rn(nx, T.kx, T.bx) : fx, fz1 = az ? __fmaf_rn(nz, T.kz, T.bz) : fz;
        const float xl = fx - T.ex, zl = fz - T.ez, xh = fx + T.ex, zh = fz + T.ez, xh1 = fx1 + T.ex, zh1 = fz1 + T.ez;
        const bool okx = ax && sibling_may_pass<CURV, ASC>(T, fmaxf(xl, T.tmin), fminf(fminf(xh1, zh), tcap), mx);
        const bool okz = az && sibling_may_pass<CURV, ASC>(T, fmaxf(zl, T.tmin), fminf(fminf(zh1, xh), tcap), mz);
        const bool okd = ax && az && sibling_may_pass<CURV, ASC>(T, fmaxf(fmaxf(xl, zl), T.tmin), fminf(fminf(xh1, zh1), tcap), md);
        const uint32_t m = (okx ? (1u << (own ^ 1u)) : 0u) | (okz ? (1u << (own ^ 2u)) : 0u) | (okd ? (1u << (own ^ 3u)) : 0u);
        seeds |= (unsigned long long)m << (4u * L);
        fx = fx1; fz = fz1;
        size = size + size;
    }
    return seeds;
}
#endif

// Pops the leaf on top of the stack and runs the exact ray / bilinear-patch solve (:167-235).
// Returns true when the ray is finished (any-hit rays stop at their first hit).
template <bool ANY_HIT, bool CURV>
__device__ __forceinline__ bool leaf_node(const FastScene& S, TraceState& T, const uint32_t node);

template <bool ANY_HIT, bool CURV>
__device__ __forceinline__ bool leaf_top(const FastScene& S, TraceState& T, const SmemStack st) {
    T.sp--;
    return leaf_node<ANY_HIT, CURV>(S, T, st.at(T.sp));
}

// The leaf solve for an already popped level-0 node (T.sp is the stack height AFTER the pop: the stale test of a
// closest-hit ray compares against it).
template <bool ANY_HIT, bool CURV>
__device__ __forceinline__ bool leaf_node(const FastScene& S, TraceState& T, const uint32_t node) {
    const uint32_t cz = (node >> 13) & 0x1FFFu, cx = node & 0x1FFFu;
    const float4 h = __ldg(S.cells + (size_t)cz * S.cell_w + cx);
    const float a0 = ((S.ox + (float)cx * S.sx) - T.o.x) * T.inv_x;
    const float a1 = ((S.ox + (float)min(cx + 1u, S.cell_w) * S.sx) - T.o.x) * T.inv_x;
    const float b0 = ((S.oz + (float)cz * S.sz) - T.o.z) * T.inv_z;
    const float b1 = ((S.oz + (float)min(cz + 1u, S.cell_h) * S.sz) - T.o.z) * T.inv_z;
    const float t_lo = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), T.tmin);
    const float t_hi = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fminf(T.tmax, T.best_t));
#if F3D_CULL_FAST
    {   // the cell's own pop tests with the reference's arithmetic (:288-304): what decides whether it is solved
        if (t_lo > t_hi) return false;
        const float2 mm = make_float2(fminf(fminf(fminf(h.x, h.y), h.z), h.w), fmaxf(fmaxf(fmaxf(h.x, h.y), h.z), h.w));
        if (!band_ok<CURV>(S, T, t_lo, t_hi, mm)) return false;
    }
#else
    if (!ANY_HIT) {
        if (T.sp < T.stale_sp) {
            T.stale_sp = T.sp;
            if (t_lo > t_hi) return false;
            const float2 mm = make_float2(fminf(fminf(fminf(h.x, h.y), h.z), h.w), fmaxf(fmaxf(fmaxf(h.x, h.y), h.z), h.w));
            if (!band_ok<CURV>(S, T, t_lo, t_hi, mm)) return false;
        }
    }
#endif
    const float tm = 0.5f * (t_lo + t_hi);
    float d3[3];
    const float fcx = (float)cx, fcz = (float)cz;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float t = i == 0 ? t_lo : (i == 1 ? tm : t_hi);
        const float px = T.o.x + t * T.d.x, pz = T.o.z + t * T.d.z;
        const float u = clampf(fdiv(px - S.ox, S.sx) - fcx, 0.0f, 1.0f);
        const float v = clampf(fdiv(pz - S.oz, S.sz) - fcz, 0.0f, 1.0f);
        const float hh = mixf(mixf(h.x, h.y, u), mixf(h.z, h.w, u), v);
        d3[i] = ray_height<CURV>(S, T, t) - hh;
    }
    const float cc = d3[0];
    const float a = 2.0f * d3[2] + 2.0f * d3[0] - 4.0f * d3[1];
    const float b = d3[2] - d3[0] - a;
    float s_hit = 1e30f;
    if (ANY_HIT && cc <= 0.0f) s_hit = 0.0f;
    else if (fabsf(a) < 1e-12f) {
        if (fabsf(b) > 1e-12f) {
            const float s = fdiv(-cc, b);
            if (s >= 0.0f && s <= 1.0f) s_hit = s;
        }
    } else {
        const float disc = b * b - 4.0f * a * cc;
        if (disc >= 0.0f) {
            const float sq = fsqrt(disc);
            const float q = -0.5f * (b + (b >= 0.0f ? sq : -sq));
            float r0 = fdiv(q, a);
            float r1 = fabsf(q) < 1e-30f ? 1e30f : fdiv(cc, q);
            if (r0 > r1) { const float tmp = r0; r0 = r1; r1 = tmp; }
            if (r0 >= 0.0f && r0 <= 1.0f) s_hit = r0;
            else if (r1 >= 0.0f && r1 <= 1.0f) s_hit = r1;
        }
    }
    if (s_hit <= 1.0f) {
        const float t = t_lo + s_hit * (t_hi - t_lo);
        if (t > T.tmin && t < T.tmax && t < T.best_t) {
            T.hit = true;
            T.best_t = t;
            T.best_cx = cx; T.best_cz = cz;
            if (ANY_HIT) return true;
            T.stale_sp = T.sp;            // everything still stacked was tested against the old best_t
        }
    }
    return false;
}

// The expansion every caller uses.  EXACT_CULL = true keeps the round-1 expansion (required for descending curved rays).
template <bool ANY_HIT, bool CURV, bool EXACT_CULL = false>
__device__ __forceinline__ void expand_node(const FastScene& S, TraceState& T, const SmemStack st) {
#if F3D_CULL_FAST
    if (!EXACT_CULL) { expand_cull<ANY_HIT, CURV>(S, T, st); return; }
#endif
    expand_top<ANY_HIT, CURV>(S, T, st);
}

// Whole-ray traversal, warp-cooperative: must be called by all 32 lanes of a converged warp
// (`valid` = this lane really has a ray).  Scheduling is a bounded while-while: lanes keep expanding
// internal nodes until at least kLeafBatch lanes of the warp hold a leaf on top of their stacks (or
// nobody can expand any more); then those lanes run the patch solve together.  A lane's own sequence
// of expansions and leaf tests is unchanged by the scheduling, so results do not depend on it.
#ifndef F3D_LEAF_BATCH
#define F3D_LEAF_BATCH 4
#endif
constexpr int kLeafBatch = F3D_LEAF_BATCH;
#ifndef F3D_LEAF_BATCH_COOP
#define F3D_LEAF_BATCH_COOP F3D_LEAF_BATCH
#endif
constexpr int kLeafBatchCoop = F3D_LEAF_BATCH_COOP;   // static-lane traversal (primary / G-buffer rays)

template <bool ANY_HIT, bool CURV>
__device__ __forceinline__ FastHit trace_fast(const FastScene& S, const Ray& r, bool valid, const SmemStack st, uint32_t& nodes) {
    TraceState T;
    T.sp = 0u; T.hit = false; T.best_t = r.tmax; T.best_cx = 0u; T.best_cz = 0u;
    if (valid) trace_begin<CURV>(S, r, T, st);
    bool busy = T.sp != 0u;
    while (__ballot_sync(0xFFFFFFFFu, busy) != 0u) {
        while (true) {
            const bool can_expand = busy && !top_is_leaf(T, st);
            if (can_expand) {
                // curved rays: only ascending any-hit rays take the conservative expansion (see F3D_CULL_FAST); this
                // per-lane choice exists for the KAT seam, the renderer's curved rays all share the sun's direction
                if (CURV) {
                    if (ANY_HIT && T.d.y >= 0.0f && T.tmin >= 0.0f) expand_node<ANY_HIT, CURV, false>(S, T, st);
                    else expand_node<ANY_HIT, CURV, true>(S, T, st);
                } else expand_node<ANY_HIT, CURV, false>(S, T, st);
                nodes++;
                if (T.sp == 0u) busy = false;
            }
            const bool expandable = busy && !top_is_leaf(T, st);
            const uint32_t m_exp = __ballot_sync(0xFFFFFFFFu, expandable);
            const uint32_t m_leaf = __ballot_sync(0xFFFFFFFFu, busy && !expandable);
            if (m_exp == 0u || __popc(m_leaf) >= kLeafBatchCoop) break;
        }
        if (busy && top_is_leaf(T, st)) {
            nodes++;
            if (leaf_top<ANY_HIT, CURV>(S, T, st) || T.sp == 0u) busy = false;
        }
    }
    FastHit res;
    res.hit = T.hit; res.t = T.best_t; res.cx = T.best_cx; res.cz = T.best_cz;
    return res;
}

// Hit point + normal of a closest hit (res.point / res.normal of :311-312).
__device__ __forceinline__ void finish_hit(const FastScene& S, const Ray& r, const FastHit& fh, v3& point, v3& normal) {
    point = r.o + r.d * fh.t;
    const float4 h = __ldg(S.cells + (size_t)fh.cz * S.cell_w + fh.cx);
    const float u = clampf(fdiv(point.x - S.ox, S.sx) - (float)fh.cx, 0.0f, 1.0f);
    const float v = clampf(fdiv(point.z - S.oz, S.sz) - (float)fh.cz, 0.0f, 1.0f);
    const float dh_du = mixf(h.y - h.x, h.w - h.z, v);
    const float dh_dv = mixf(h.z - h.x, h.w - h.y, u);
    normal = normalize3(V3(fdiv(-dh_du, S.sx), 1.0f, fdiv(-dh_dv, S.sz)));
}

}  // namespace f3d
