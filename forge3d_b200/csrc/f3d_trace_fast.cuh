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

// ---------------------------------------------------------------------------------------------------------------
// Sun horizon strips (F3D_SUN_HORIZON, default 1): an acceleration structure for the ONE direction every sun ray shares.
// In cell units a sun ray is the line b = w + m a (a = coordinate along the MAJOR axis of the direction, b along the
// minor one, |m| <= 1), so w = b - m a is constant along the ray: rays with floor(w) = j form STRIP j, and inside the
// column of cells a in [i, i+1] a ray of strip j can only touch rows floor(j - d + min(m i, m (i+1))) ..
// floor(j + 1 + d + max(m i, m (i+1))) (at most four; d = 1/64 cell absorbs every rounding of w).  Let Hs[j][q] be the
// largest corner height of those cells, columns numbered q = 0, 1, .. in TRAVEL order.  An ascending ray (d.y >= 0, the
// curvature term only lifts it) that starts at travel coordinate e0 and height y0 enters column q at height
// y0 + (q - e0) g, g = height gained per column, and that is its LOWEST height inside the column.  Hence with
//     S[q][j] = max over q' >= q of (Hs[j][q'] - q' g)         (one suffix maximum per strip, built once per session)
// the test  y0 - e0 g - pad > S[q][j]  proves: in every cell of every column >= q the ray's height range lies above the
// cell's maximum, i.e. the reference's own band test (hybrid_terrain_traversal.wgsl:303-304) rejects each of those leaves
// without solving it.  Such cells cannot contribute to the occlusion flag, so the tracer may skip every node that lies
// entirely in columns >= q ("cleared"): an INTEGER test on cell coordinates, no float pads.  What remains is the near
// field: the origin column and the next k-1 columns, k the smallest of kHzK that clears the ray.  On the C2 scene (sun
// elevation 24 deg) 69 % of the sun rays are cleared from the second column on, 86 % from the eighth; only 10 % are
// occluded at all.  Exactness: the flag is an OR over the leaves that pass their own tests (F3D_CULL_FAST (1)-(2)); a
// cleared leaf fails its band test in the reference's arithmetic - pad covers the f32 evaluation errors of both sides -
// so dropping it leaves the OR unchanged.
// ---------------------------------------------------------------------------------------------------------------
#ifndef F3D_SUN_HORIZON
#define F3D_SUN_HORIZON 1
#endif
constexpr int kHzN = 8;
__host__ __device__ inline uint32_t hz_k(uint32_t idx) {
    return idx == 0u ? 1u : idx == 1u ? 2u : idx == 2u ? 3u : idx == 3u ? 4u : idx == 4u ? 6u : idx == 5u ? 8u : idx == 6u ? 12u : 16u;
}
constexpr uint32_t kHzNone = 15u;
constexpr int32_t kHzNoClear = 0x7FFFFFFF;
struct SunHorizon {
    const float* S;          // [ncols][nstrips]; nullptr = no horizon (descending sun, vertical sun, switched off)
    uint32_t nstrips, ncols; // ncols = cells along the major axis
    int32_t j0;              // strip index = floor(w) - j0
    float m, g;              // minor-per-major slope in cell units; height gained per column of travel (>= 0)
    float pad_rel, mag;      // pad = pad_rel * (|y0| + e0 g + mag)
    uint32_t xmajor;         // 1: the major axis is x
    uint32_t forward;        // 1: the rays travel towards increasing major coordinate
};

// ---------------------------------------------------------------------------------------------------------------
// TMA staging of the top pyramid levels (F3D_TMA_STAGE, default 0: measured slower, see below).  The quad-packed levels live in ONE arena, finest
// level first, so the top K levels (everything from level `first` up to the root's children) are one contiguous block
// at its end: 4^0 + 4^1 + ... quads of 32 bytes, 10.9 KB for the top 5 levels of a 2048^2 DEM.  Every ray of a top-down
// traversal reads exactly these nodes first, every bottom-up ray reads them last.  Each CTA of the ray kernels copies
// the block into shared memory ONCE with a single 1-D bulk async copy (cp.async.bulk.shared::cluster.global, the TMA
// engine: SASS UBLKCP) that signals an mbarrier, and builds a table of per-level base pointers (generic addresses:
// staged levels point into shared memory, the others into the arena), so a node fetch is one table lookup + plain
// loads with no "is it staged" branch.  QGlobal / QStaged are the two ways expand_core & co. fetch a quad.
// ---------------------------------------------------------------------------------------------------------------
// MEASURED on the B200 (gpurun_out r02l, profiles/r02_tma_staging.md), C2 frame loop, bit-identical images:
//   staging on 0.961 ms/frame | table only (F3D_B200_TMA_STAGE=0) 0.904 | compiled out (F3D_TMA_STAGE=0) 0.883.
// The staged nodes are the ones every ray shares, i.e. exactly the lines that already sit in L1 (hit rate 69-75 % in
// k_ptrace); a per-CTA private copy takes 11-44 KB per CTA out of the SM's unified L1/shared array (k_ascent L1 hit
// 48 % -> 38 %) and out of the co-residency of the other batch set's kernels.  Default therefore 0; build with
// -DF3D_TMA_STAGE=1 (forge3d_b200/build.py defines="F3D_TMA_STAGE=1") for the staged variant, which stays bit-exact
// (tests/test_gpu_parity.py::test_tma_staged_variant_is_bit_identical).
#ifndef F3D_TMA_STAGE
#define F3D_TMA_STAGE 0
#endif
struct StageParams {
    const float2* src;       // first staged quad in the arena (16-byte aligned)
    uint32_t first;          // lowest staged level (levels first .. mip_count-2 are staged); >= mip_count-1: none
    uint32_t bytes;          // size of the staged block (multiple of 32); 0 = table only
};
struct LevelTable { const float2* lv[kMaxLevels]; };

struct QGlobal {
    __device__ __forceinline__ const float2* base(const FastScene& S, uint32_t cl) const { return S.q.lv[cl]; }
    __device__ __forceinline__ float2 ld(const float2* p) const { return __ldg(p); }
};
struct QStaged {
    const LevelTable* tab;   // in shared memory
    __device__ __forceinline__ const float2* base(const FastScene&, uint32_t cl) const { return tab->lv[cl]; }
    __device__ __forceinline__ float2 ld(const float2* p) const { return *p; }
};
__host__ __device__ inline size_t stage_smem_bytes(uint32_t staged_bytes) {      // data + table + mbarrier
    return (size_t)((staged_bytes + 15u) & ~15u) + sizeof(LevelTable) + 16u;
}

// Called by EVERY thread of the CTA at kernel entry, converged.  `area` = 16-byte aligned shared memory of
// stage_smem_bytes(sp.bytes).  Returns the accessor; the staged data is visible to all threads on return.
__device__ __forceinline__ QStaged stage_top_levels(const FastScene& F, const StageParams sp, unsigned char* area) {
    const uint32_t data_bytes = (sp.bytes + 15u) & ~15u;
    LevelTable* tab = reinterpret_cast<LevelTable*>(area + data_bytes);
    const uint32_t top = F.mip_count - 1u;                 // levels 0 .. top-1 have quads
    if (threadIdx.x < (uint32_t)kMaxLevels) {
        const uint32_t l = threadIdx.x;
        const float2* g = F.q.lv[l];
        const bool staged = sp.bytes != 0u && l >= sp.first && l < top;
        tab->lv[l] = staged ? reinterpret_cast<const float2*>(area) + (g - sp.src) : g;
    }
#if defined(__CUDA_ARCH__)
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(area + data_bytes + sizeof(LevelTable));
    const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(mbar);
    if (sp.bytes != 0u) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(sp.bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(area)), "l"(sp.src), "r"(sp.bytes), "r"(mbar_s) : "memory");
        }
        __syncthreads();                                    // the barrier is initialised for everybody (and the table written)
        uint32_t done = 0u;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mbar_s) : "memory");
        }
    } else __syncthreads();
#else
    if (threadIdx.x == 0)
        for (uint32_t i = 0; i < sp.bytes / 8u; i++) reinterpret_cast<float2*>(area)[i] = sp.src[i];
    __syncthreads();
#endif
    QStaged q;
    q.tab = tab;
    return q;
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
    int32_t hz_qclear;       // sun horizon: every cell in travel columns >= hz_qclear is cleared (kHzNoClear: none)
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
template <bool ANY_HIT, bool CURV, bool ASC, class Q = QGlobal, bool HZ = false>
__device__ __forceinline__ uint32_t expand_core(const FastScene& S, const TraceState& T, const uint32_t node, uint32_t& bid, const Q q = Q(),
                                                const SunHorizon* Z = nullptr) {
    const uint32_t level = (node >> 26) & 15u, ny = (node >> 13) & 0x1FFFu, nx = node & 0x1FFFu;
    const uint32_t cl = level - 1u;
    const bool fx = T.inv_x < 0.0f, fz = T.inv_z < 0.0f;
    const uint32_t flip = (fx ? 1u : 0u) | (fz ? 2u : 0u);
    // the children's [min,max]: four 8-byte loads from one 32-byte sector, already in sign order
    const float2* qp = q.base(S, cl) + (ny * S.q.parent_pitch[cl] + nx) * 4u;
    const float2 m0 = q.ld(qp + flip), m1 = q.ld(qp + (flip ^ 1u)), m2 = q.ld(qp + (flip ^ 2u)), m3 = q.ld(qp + (flip ^ 3u));
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
    if (HZ) {   // sun horizon: children that lie entirely in cleared columns (see SunHorizon) - an integer test
        const uint32_t nmaj = Z->xmajor ? nx : ny;
        const int32_t ts = Z->forward ? (int32_t)(nmaj << level) : (int32_t)Z->ncols - (int32_t)((nmaj + 1u) << level);   // first travel column of the node
        if (ts + (int32_t)(1u << cl) >= T.hz_qclear) okm &= Z->xmajor ? 0x5u : 0x3u;     // drop the far half along the major axis
    }
    return okm;
}

// Stack form: pops the internal node on top of the stack and pushes, far child first, the children that may pass.
// Precondition: sp > 0 and the top is not a leaf.
template <bool ANY_HIT, bool CURV, class Q = QGlobal>
__device__ __forceinline__ void expand_cull(const FastScene& S, TraceState& T, const SmemStack st, const Q q = Q()) {
    T.sp--;
    uint32_t bid;
    const uint32_t okm = expand_core<ANY_HIT, CURV, false, Q>(S, T, st.at(T.sp), bid, q);
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

// Sun horizon lookup for a ray that starts at `o` inside cell `cell0` (origin_cell): the index into kHzK of the smallest
// k for which every column from q0 + k on is cleared (see SunHorizon), or kHzNone.  All loads are issued together.
__device__ __forceinline__ uint32_t horizon_lookup(const SunHorizon& Z, const FastScene& S, v3 o, uint32_t cell0) {
    if (Z.S == nullptr) return kHzNone;
    const float ua = fdiv(o.x - S.ox, S.sx), ub = fdiv(o.z - S.oz, S.sz);
    const float a0 = Z.xmajor ? ua : ub, b0 = Z.xmajor ? ub : ua;
    const uint32_t c0maj = Z.xmajor ? (cell0 & 0x1FFFu) : (cell0 >> 13);
    const int32_t q0 = Z.forward ? (int32_t)c0maj : (int32_t)Z.ncols - 1 - (int32_t)c0maj;
    const float e0 = Z.forward ? a0 : (float)Z.ncols - a0;                  // travel coordinate of the origin, in [q0, q0 + 1]
    if (!(e0 >= (float)q0 - 0.01f && e0 <= (float)q0 + 1.01f)) return kHzNone;     // origin outside the DEM (cell0 is clamped): no claim
    const float w = b0 - Z.m * a0;
    const int32_t j = (int32_t)floorf(w) - Z.j0;
    if (j < 0 || j >= (int32_t)Z.nstrips) return kHzNone;
    const float eg = e0 * Z.g;
    const float Y = (o.y - eg) - Z.pad_rel * (fabsf(o.y) + eg + Z.mag);
    float sv[kHzN];
#pragma unroll
    for (int i = 0; i < kHzN; i++) {
        const int32_t q = q0 + (int32_t)hz_k((uint32_t)i);
        sv[i] = q < (int32_t)Z.ncols ? __ldg(Z.S + (size_t)q * Z.nstrips + (uint32_t)j) : -3.0e38f;      // beyond the DEM: nothing there
    }
    uint32_t idx = kHzNone;
#pragma unroll
    for (int i = kHzN - 1; i >= 0; i--)
        if (Y > sv[i]) idx = (uint32_t)i;
    return idx;
}

// The ray parameter at which a sun ray enters travel column q_clear, rounded UP by the culling pad: everything the ray
// meets after it lies in cleared columns.  Used as the cap of the CONSERVATIVE span tests only (expand_core, ascent_seeds:
// tcap = min(tmax, best_t)); a node is then dropped iff a lower bound of its entry parameter exceeds the cap, i.e. iff the
// ray is inside it only beyond the cleared boundary.  Needs cull_setup() done on T.
__device__ __forceinline__ float horizon_t_clear(const SunHorizon& Z, const TraceState& T, int32_t q_clear) {
    const float ap = (float)(Z.forward ? q_clear : (int32_t)Z.ncols - q_clear);      // cell coordinate of the boundary plane
    return Z.xmajor ? __fmaf_rn(ap, T.kx, T.bx) + T.ex : __fmaf_rn(ap, T.kz, T.bz) + T.ez;
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

// Conservative own test of ONE cell (span + band, same widening as expand_core): may the reference's pop tests of the
// level-0 node (cx, cz) pass for this ray?  Used by the near-field walks of k_ascent, which hand the survivors to the exact
// patch solve (leaf_node re-does the test with the reference's arithmetic).
template <bool CURV, bool ASC, class Q = QGlobal>
__device__ __forceinline__ bool cell_may_pass(const FastScene& S, const TraceState& T, const uint32_t cx, const uint32_t cz, const Q q = Q()) {
    const float2 mm = q.ld(q.base(S, 0u) + ((cz >> 1) * S.q.parent_pitch[0] + (cx >> 1)) * 4u + ((cz & 1u) * 2u + (cx & 1u)));
    const float fx = (float)cx, fz = (float)cz;
    const float a0 = __fmaf_rn(fx, T.kx, T.bx), a1 = __fmaf_rn(fx + 1.0f, T.kx, T.bx);
    const float b0 = __fmaf_rn(fz, T.kz, T.bz), b1 = __fmaf_rn(fz + 1.0f, T.kz, T.bz);
    const float tl = fmaxf(fmaxf(fminf(a0, a1) - T.ex, fminf(b0, b1) - T.ez), T.tmin);
    const float th = fminf(fminf(fmaxf(a0, a1) + T.ex, fmaxf(b0, b1) + T.ez), fminf(T.tmax, T.best_t));
    return sibling_may_pass<CURV, ASC>(T, tl, th, mm);
}

// ---------------------------------------------------------------------------------------------------------------
// Near-field walk (F3D_SUN_NEAR, default 1).  A sun ray whose horizon lookup says "cleared from travel column q0 + k on"
// with a small k can only be occluded inside the k columns it starts in.  In cell units the ray is the line b = w + m a
// (see SunHorizon), so in the column of cells a in [i, i+1] it can only touch rows floor(w - d + min(m i, m (i+1))) ..
// floor(w + d + max(m i, m (i+1))): at most three (|m| <= 1; d = 1/32 cell absorbs every rounding of w and of the
// reference's slab arithmetic, which moves a plane by ~1e-4 cell at most).  k_ascent walks those <= 3 k cells directly -
// no bottom-up seeds, no k_trace - tests each with cell_may_pass and queues the survivors for the exact patch solve.
// Exactness: the occlusion flag is the OR of the leaf solves over the cells that pass their OWN span + band test
// (F3D_CULL_FAST (1)-(2)); every such cell lies in a column the ray touches: columns >= q0 + k are cleared (SunHorizon),
// columns < q0 are behind the origin (empty clipped span, as for the siblings ascent_seeds drops), and inside the near
// columns the walk visits a superset of the rows the ray touches.
// ---------------------------------------------------------------------------------------------------------------
#ifndef F3D_SUN_NEAR
#define F3D_SUN_NEAR 1
#endif
#ifndef F3D_SUN_NEAR_MAX_IDX
#define F3D_SUN_NEAR_MAX_IDX 3          // index into hz_k: rays cleared within hz_k(3) = 4 columns take the walk
#endif
constexpr uint32_t kSunNearMaxIdx = F3D_SUN_NEAR_MAX_IDX;
constexpr uint32_t kSunNearMaxCols = 4u;    // >= hz_k(kSunNearMaxIdx)
static_assert(F3D_SUN_NEAR_MAX_IDX <= 3, "the walk visits at most kSunNearMaxCols columns");

// The walk may only replace the bottom-up start for a ray that starts strictly inside its origin cell's slabs (the same
// condition, with the same widening, under which ascent_seeds drops the siblings behind the ray): then every cell behind
// the origin cell's entry planes has an empty clipped span in the reference's arithmetic as well.
__device__ __forceinline__ bool starts_inside(const TraceState& T, const uint32_t cell0) {
    const float lox = (float)(cell0 & 0x1FFFu), loz = (float)(cell0 >> 13);
    const bool sgx = T.inv_x < 0.0f, sgz = T.inv_z < 0.0f;
    const float ex_in = __fmaf_rn(sgx ? lox + 1.0f : lox, T.kx, T.bx), ez_in = __fmaf_rn(sgz ? loz + 1.0f : loz, T.kz, T.bz);
    return fmaxf(ex_in + T.ex, ez_in + T.ez) < T.tmin;
}
// The line a walk follows, in cell units: b = w + m a with a along the major axis (sun rays: shared, from SunHorizon;
// escape-map rays: per ray).
struct WalkDir { float m; uint32_t xmajor, forward, ncols; };
__device__ __forceinline__ WalkDir walk_dir(const SunHorizon& Z) { WalkDir D; D.m = Z.m; D.xmajor = Z.xmajor; D.forward = Z.forward; D.ncols = Z.ncols; return D; }
__device__ __forceinline__ WalkDir walk_dir(const FastScene& S, v3 d) {
    WalkDir D;
    const float du = d.x * S.sz, dv = d.z * S.sx;                    // (d.x / sx, d.z / sz) scaled by sx sz > 0
    D.xmajor = fabsf(du) >= fabsf(dv) ? 1u : 0u;
    D.m = D.xmajor ? fdiv(dv, du) : fdiv(du, dv);                    // |m| <= 1 up to rounding; the walk's pad absorbs it
    D.forward = (D.xmajor ? d.x : d.z) > 0.0f ? 1u : 0u;
    D.ncols = D.xmajor ? S.cell_w : S.cell_h;
    return D;
}
// Per-ray constants of the walk: major-axis cell of the origin, line offset w.
struct NearWalk { float w; uint32_t c0maj; };
__device__ __forceinline__ NearWalk near_walk_setup(const WalkDir& Z, const FastScene& S, v3 o, uint32_t cell0) {
    const float ua = fdiv(o.x - S.ox, S.sx), ub = fdiv(o.z - S.oz, S.sz);
    NearWalk N;
    N.w = (Z.xmajor ? ub : ua) - Z.m * (Z.xmajor ? ua : ub);
    N.c0maj = Z.xmajor ? (cell0 & 0x1FFFu) : (cell0 >> 13);
    return N;
}
// Rows the ray can touch in travel column q0 + i (absolute major cell index returned in `cmaj`); false: the column lies
// outside the DEM.  r_lo .. r_hi may reach outside [0, nrows): the caller clips.
__device__ __forceinline__ bool near_walk_column(const WalkDir& Z, const NearWalk& N, uint32_t i, int32_t& cmaj, int32_t& r_lo, int32_t& r_hi) {
    cmaj = Z.forward ? (int32_t)N.c0maj + (int32_t)i : (int32_t)N.c0maj - (int32_t)i;
    if (cmaj < 0 || cmaj >= (int32_t)Z.ncols) return false;
    const float e0 = Z.m * (float)cmaj, e1 = Z.m * (float)(cmaj + 1);
    r_lo = (int32_t)floorf(N.w - 0.03125f + fminf(e0, e1));
    r_hi = (int32_t)floorf(N.w + 0.03125f + fmaxf(e0, e1));
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Escape map (F3D_ESCAPE, default 1): the directional generalisation of the sun horizon, for rays of ANY direction (the IBL
// rays).  Per DEM cell c and per octant o of the horizontal direction (sign of x, sign of z, which axis is the major one, in
// cell units) the map holds E[c][o] >= (hmax(c') + pad - hmin(c)) / D(c, c') for EVERY cell c' that (a) lies further than
// kEscR cells (Chebyshev) from c and (b) can be reached from a point of c along a direction of octant o; D = the smallest
// horizontal distance between a point of c (widened by 1/64 cell) and a point of c'.  An ascending ray (d.y > 0, no
// curvature term) that starts inside c at a height >= hmin(c) with slope d.y / |d.xz| > E[c][o] is, inside every such c',
// above hmax(c'): its clipped span there starts no earlier than D / |d.xz| along the ray (the slab planes are the
// reference's, moved by <= 1e-3 cell by its rounding), its height is non-decreasing, so the low end of the reference's
// band test (hybrid_terrain_traversal.wgsl:303-304) exceeds the cell's maximum and the cell is never solved.  What is left
// is the near field: the (2 kEscR + 1)^2 block around c, walked like the sun's near field (kEscR + 1 columns of <= 3 rows).
// The map is built once per session from the min-max pyramid (k_escape_build): level by level the ring of nodes at index
// distance kEscR+1 .. 2 kEscR+1 around c's ancestor - a level's inner block is covered by the next finer level's inner
// block plus ring, so the rings of all levels plus the finest inner block (the near field) cover the DEM.
// ---------------------------------------------------------------------------------------------------------------
#ifndef F3D_ESCAPE
#define F3D_ESCAPE 1
#endif
constexpr int kEscR = 2;
constexpr int kEscOuter = 2 * kEscR + 1;
constexpr uint32_t kEscNearCols = (uint32_t)kEscR + 1u;
struct EscapeMap { const float* E; };       // [cell][8]; nullptr = none
struct EscapeOctTable { uint8_t m[2 * kEscOuter + 1][2 * kEscOuter + 1]; };   // octant mask of a ring node by its index offset [oz][ox]

// Octants a ring node at index offset (ox, oz) (in units of its level's node size) can be reached in: the offsets between a
// point of the origin's node and a point of that node fill [ox-1, ox+1] x [oz-1, oz+1].  Octant = sx | sz << 1 | xmajor << 2.
inline EscapeOctTable escape_oct_table() {
    EscapeOctTable T{};
    for (int oz = -kEscOuter; oz <= kEscOuter; oz++)
        for (int ox = -kEscOuter; ox <= kEscOuter; ox++) {
            uint8_t mask = 0;
            for (int sx = 0; sx < 2; sx++)
                for (int sz = 0; sz < 2; sz++) {
                    int a0 = sx ? -(ox + 1) : ox - 1, a1 = sx ? -(ox - 1) : ox + 1;
                    int b0 = sz ? -(oz + 1) : oz - 1, b1 = sz ? -(oz - 1) : oz + 1;
                    if (a1 < 0 || b1 < 0) continue;
                    if (a0 < 0) a0 = 0;
                    if (b0 < 0) b0 = 0;
                    if (a1 >= b0) mask |= (uint8_t)(1u << (sx | (sz << 1) | 4));       // |dx| >= |dz| somewhere in the box
                    if (b1 >= a0) mask |= (uint8_t)(1u << (sx | (sz << 1)));
                }
            T.m[oz + kEscOuter][ox + kEscOuter] = mask;
        }
    return T;
}

// May the bottom-up start of this ray be replaced by the near-field walk?  (see EscapeMap; T after ray_setup)
__device__ __forceinline__ bool escape_cleared(const EscapeMap& M, const FastScene& S, const TraceState& T, const uint32_t cell0) {
    if (M.E == nullptr || !(T.d.y > 0.0f)) return false;
    const uint32_t cx = cell0 & 0x1FFFu, cz = cell0 >> 13;
    const float ua = fdiv(T.o.x - S.ox, S.sx), ub = fdiv(T.o.z - S.oz, S.sz);       // the origin really lies in cell0 (it is clamped)
    const float e = 0.0078125f, fx = (float)cx, fz = (float)cz;
    if (!(ua >= fx - e && ua <= fx + 1.0f + e && ub >= fz - e && ub <= fz + 1.0f + e)) return false;
    const float4 h = __ldg(S.cells + (size_t)cz * S.cell_w + cx);
    if (!(T.o.y >= fminf(fminf(h.x, h.y), fminf(h.z, h.w)))) return false;
    const uint32_t oct = (T.d.x < 0.0f ? 1u : 0u) | (T.d.z < 0.0f ? 2u : 0u) | (fabsf(T.d.x) * S.sz >= fabsf(T.d.z) * S.sx ? 4u : 0u);
    const float slope = (T.d.y * frsqrt(T.hd2)) * 0.999755859375f;                  // 1 - 2^-12: lower bound of d.y / |d.xz|
    return slope > __ldg(M.E + ((size_t)cz * S.cell_w + cx) * 8u + oct);
}

// Needs ray_setup() done on T (tmin/tmax as the tracer will use them, or looser).
// Going up one level adds ONE new plane per axis (A_{L+1} extends A_L on its low or its high side), and only an
// extension on the side the ray travels to can be met: per level 2 fma for the new far planes, three 8-byte loads and up
// to three sibling tests whose spans follow from the far planes alone:
//   x-sibling  [far_x(L), min(far_x(L+1), far_z(L))]      (shares A_L's z-slab; entered through A_L's far x plane)
//   z-sibling  [far_z(L), min(far_z(L+1), far_x(L))]
//   diagonal   [max(far_x(L), far_z(L)), min(far_x(L+1), far_z(L+1))]
// each bound widened by the per-axis pad, clipped to [tmin, min(tmax, best_t)].
template <bool CURV, bool ASC, class Q = QGlobal, bool HZ = false>
__device__ __forceinline__ unsigned long long ascent_seeds(const FastScene& S, const TraceState& T, const uint32_t cell0, const Q q = Q(),
                                                           const SunHorizon* Z = nullptr, bool* defer_all_siblings = nullptr) {
    const uint32_t cx0 = cell0 & 0x1FFFu, cz0 = cell0 >> 13;
    const uint32_t top = S.mip_count - 1u;
    const bool sgx = T.inv_x < 0.0f, sgz = T.inv_z < 0.0f;
    float lox = (float)cx0, hix = lox + 1.0f, loz = (float)cz0, hiz = loz + 1.0f;       // A_L in cell units, exact in f32
    const float txl = __fmaf_rn(lox, T.kx, T.bx), txh = __fmaf_rn(hix, T.kx, T.bx);
    const float tzl = __fmaf_rn(loz, T.kz, T.bz), tzh = __fmaf_rn(hiz, T.kz, T.bz);
    float fx = sgx ? txl : txh, fz = sgz ? tzl : tzh;                                   // far planes of A_L
    const bool inside = fmaxf((sgx ? txh : txl) + T.ex, (sgz ? tzh : tzl) + T.ez) < T.tmin;
    unsigned long long seeds = 0ull;
    // sun horizon: travel column of the origin cell and how many columns remain to the far end of A_L
    uint32_t hz_c0 = 0u;
    int32_t hz_k = kHzNoClear;                       // cleared from travel column q0 + hz_k on
    if (HZ) {
        const uint32_t c0maj = Z->xmajor ? cx0 : cz0;
        hz_c0 = Z->forward ? c0maj : ~c0maj;         // low L bits: offset inside A_L along the direction of travel
        const int32_t q0 = Z->forward ? (int32_t)c0maj : (int32_t)Z->ncols - 1 - (int32_t)c0maj;
        hz_k = T.hz_qclear == kHzNoClear ? kHzNoClear : T.hz_qclear - q0;
    }
    if (!inside) {                                   // rare (origin within the pad of a cell border, or outside the DEM):
        if (defer_all_siblings != nullptr) { *defer_all_siblings = true; return 0ull; }    // the caller runs all_sibling_seeds_warp
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
        const float2* qp = q.base(S, L) + ((cz0 >> (L + 1u)) * S.q.parent_pitch[L] + (cx0 >> (L + 1u))) * 4u;
        const float2 mx = q.ld(qp + (own ^ 1u)), mz = q.ld(qp + (own ^ 2u)), md = q.ld(qp + (own ^ 3u));
        // the new plane of A_{L+1} per axis; it is a FAR plane iff the extension is on the side the ray travels to
        const float nx = bx ? lox - size : hix + size, nz = bz ? loz - size : hiz + size;
        if (bx) lox = nx; else hix = nx;
        if (bz) loz = nz; else hiz = nz;
        const bool ax = (bx != 0u) == sgx, az = (bz != 0u) == sgz;
        const float fx1 = ax ? __fmaf_rn(nx, T.kx, T.bx) : fx, fz1 = az ? __fmaf_rn(nz, T.kz, T.bz) : fz;
        const float xl = fx - T.ex, zl = fz - T.ez, xh = fx + T.ex, zh = fz + T.ez, xh1 = fx1 + T.ex, zh1 = fz1 + T.ez;
        bool okx = ax && sibling_may_pass<CURV, ASC>(T, fmaxf(xl, T.tmin), fminf(fminf(xh1, zh), tcap), mx);
        bool okz = az && sibling_may_pass<CURV, ASC>(T, fmaxf(zl, T.tmin), fminf(fminf(zh1, xh), tcap), mz);
        bool okd = ax && az && sibling_may_pass<CURV, ASC>(T, fmaxf(fmaxf(xl, zl), T.tmin), fminf(fminf(xh1, zh1), tcap), md);
        if (HZ) {   // the siblings beyond A_L's far plane on the major axis start at travel column q0 + rem: cleared?
            const int32_t rem = (int32_t)((1u << L) - (hz_c0 & ((1u << L) - 1u)));
            if (rem >= hz_k) { okd = false; if (Z->xmajor) okx = false; else okz = false; }
        }
        const uint32_t m = (okx ? (1u << (own ^ 1u)) : 0u) | (okz ? (1u << (own ^ 2u)) : 0u) | (okd ? (1u << (own ^ 3u)) : 0u);
        seeds |= (unsigned long long)m << (4u * L);
        fx = fx1; fz = fz1;
        size = size + size;
    }
    return seeds;
}

// The "every sibling that exists" seeds of one ray (the !inside case above), computed by a whole warp: lane 4 L' + r owns
// sibling r of level L' (two passes cover 16 levels), the ballots ARE the seed bits.  ~1 % of the rays take this path;
// run per lane it cost 10 % of k_ascent's warp-instructions at 1.4 active lanes.  Called converged by all 32 lanes.
__device__ __forceinline__ unsigned long long all_sibling_seeds_warp(const FastScene& S, const uint32_t cell0) {
    const uint32_t lane = threadIdx.x & 31u, r = lane & 3u;
    const uint32_t cx0 = cell0 & 0x1FFFu, cz0 = cell0 >> 13, top = S.mip_count - 1u;
    unsigned long long seeds = 0ull;
#pragma unroll
    for (uint32_t pass = 0; pass < 2u; pass++) {
        const uint32_t L = pass * 8u + (lane >> 2);
        const uint32_t own = (((cz0 >> L) & 1u) << 1) | ((cx0 >> L) & 1u);
        const uint32_t gx = ((((cx0 >> (L + 1u)) << 1) | (r & 1u)) << L), gz = ((((cz0 >> (L + 1u)) << 1) | (r >> 1)) << L);
        const bool bit = L < top && r != own && gx < S.cell_w && gz < S.cell_h;
        seeds |= (unsigned long long)__ballot_sync(0xFFFFFFFFu, bit) << (32u * pass);
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
template <bool ANY_HIT, bool CURV, bool EXACT_CULL = false, class Q = QGlobal>
__device__ __forceinline__ void expand_node(const FastScene& S, TraceState& T, const SmemStack st, const Q q = Q()) {
#if F3D_CULL_FAST
    if (!EXACT_CULL) { expand_cull<ANY_HIT, CURV, Q>(S, T, st, q); return; }
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
#ifndef F3D_LEAF_BATCH_PARK
#define F3D_LEAF_BATCH_PARK 12
#endif
constexpr int kLeafBatchPark = F3D_LEAF_BATCH_PARK;   // F3D_PRIMARY_PARK: lanes that can do nothing but wait for their parked leaf

// F3D_PRIMARY_PARK = 1: a lane that meets a leaf PARKS it (one slot, a register) and keeps expanding until it meets its next
// leaf; the solve phase then finds more lanes with a leaf, and fewer lanes idle through the expansion steps.  Exact for
// closest-hit rays too: the clipped spans of the leaves a ray visits are disjoint and ordered (F3D_CULL_FAST (3)), so until the
// first hit every solve sees best_t = tmax and is independent of the others, and after the first hit at t* every later leaf
// has t_lo >= t* = best_t, i.e. an empty (or single-point) clipped span: the first hit in visit order IS the closest hit and
// the ray can stop there.  One parked leaf keeps the per-ray solve order (the parked leaf is always the oldest unsolved one).
#ifndef F3D_PRIMARY_PARK
#define F3D_PRIMARY_PARK 0
#endif
template <bool ANY_HIT, bool CURV, class Q = QGlobal>
__device__ __forceinline__ FastHit trace_fast(const FastScene& S, const Ray& r, bool valid, const SmemStack st, uint32_t& nodes, const Q q = Q()) {
    TraceState T;
    T.sp = 0u; T.hit = false; T.best_t = r.tmax; T.best_cx = 0u; T.best_cz = 0u;
    if (valid) trace_begin<CURV>(S, r, T, st);
    bool busy = T.sp != 0u;
#if F3D_PRIMARY_PARK
    if (!CURV || ANY_HIT) {      // (curved closest-hit rays keep the exact expansion with its stale re-tests: not used by the renderer)
        constexpr uint32_t kNone = 0xFFFFFFFFu;
        uint32_t parked = kNone;
        while (__ballot_sync(0xFFFFFFFFu, busy || parked != kNone) != 0u) {
            while (true) {
                if (busy && parked == kNone && top_is_leaf(T, st)) {       // park the leaf, go on with what lies behind it
                    T.sp--;
                    parked = st.at(T.sp);
                    if (T.sp == 0u) busy = false;
                }
                const bool can_expand = busy && !top_is_leaf(T, st);
                if (can_expand) {
                    if (CURV) {
                        if (ANY_HIT && T.d.y >= 0.0f && T.tmin >= 0.0f) expand_node<ANY_HIT, CURV, false, Q>(S, T, st, q);
                        else expand_node<ANY_HIT, CURV, true, Q>(S, T, st, q);
                    } else expand_node<ANY_HIT, CURV, false, Q>(S, T, st, q);
                    nodes++;
                    if (T.sp == 0u) busy = false;
                }
                // a lane can go on if it can expand, or if it can still park the leaf it just reached
                const bool more = busy && (!top_is_leaf(T, st) || parked == kNone);
                const uint32_t m_more = __ballot_sync(0xFFFFFFFFu, more);
                const uint32_t m_leaf = __ballot_sync(0xFFFFFFFFu, parked != kNone && !more);
                if (m_more == 0u || __popc(m_leaf) >= kLeafBatchPark) break;
            }
            if (parked != kNone) {
                nodes++;
                const bool done = leaf_node<ANY_HIT, CURV>(S, T, parked);
                parked = kNone;
                if (done || T.hit) { busy = false; T.sp = 0u; }       // first hit in visit order = the closest hit (see above)
            }
        }
        FastHit res;
        res.hit = T.hit; res.t = T.best_t; res.cx = T.best_cx; res.cz = T.best_cz;
        return res;
    }
#endif
    while (__ballot_sync(0xFFFFFFFFu, busy) != 0u) {
        while (true) {
            const bool can_expand = busy && !top_is_leaf(T, st);
            if (can_expand) {
                // curved rays: only ascending any-hit rays take the conservative expansion (see F3D_CULL_FAST); this
                // per-lane choice exists for the KAT seam, the renderer's curved rays all share the sun's direction
                if (CURV) {
                    if (ANY_HIT && T.d.y >= 0.0f && T.tmin >= 0.0f) expand_node<ANY_HIT, CURV, false, Q>(S, T, st, q);
                    else expand_node<ANY_HIT, CURV, true, Q>(S, T, st, q);
                } else expand_node<ANY_HIT, CURV, false, Q>(S, T, st, q);
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
