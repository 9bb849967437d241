// forge3d_b200/csrc/f3d_kernels.cuh
// CUDA kernels of the B200 terrain path tracer.  The reference runs four WGSL dispatches per frame
// (main_terrain -> pt_restir_temporal -> pt_restir_spatial, plus a one-off main_terrain_gbuffer) with
// one thread per pixel doing primary + shadow + IBL rays back to back.  Here a frame is a wavefront:
//
//   k_primary (per pixel)   prev  = spatial_reuse(out_{f-1})      [pt_restir_spatial.wgsl:158-224, frame f-1's pass]
//                           prev  = M-clamp(prev)                 [hybrid_terrain_traversal.wgsl:452-462]
//                           primary ray (closest hit), shading set-up, candidate reservoir      [:467-512]
//                           out_f = temporal_reuse(prev, cand)    [pt_restir_temporal.wgsl:54-109]
//                           pixels that need a sun / IBL ray append their index to two COMPACTED
//                           global ray lists (warp ballot + prefix sum, one atomic per warp)
//   k_trace (persistent, per ray: sun list then IBL list)  any-hit traversal [:514-545]; every lane pulls its
//                           next ray from the list as soon as its current one finishes, so sky pixels,
//                           back-facing hits and short rays never idle lanes next to long rays
//   k_accum (per pixel)     combine, accumulate, windowed Welford                               [:547-574]
//
// The `prev` and `curr` reservoir buffers of the reference never exist in HBM, and `out` is a 16-byte
// record per pixel instead of 80 bytes: on this path every populated LightSample is the one
// directional sun (direction == normalize(light_dir) bit-for-bit, light_index 0, intensity constant),
// `position` is never read by anything that reaches an output, and the receiving pixel's G-buffer test
// N.wi > 0 (pt_restir_spatial.wgsl:73-75) is a per-pixel constant folded into one bit by k_gbuffer.
// The RGBA16F beauty store of every frame (:576-579) is dead until the last frame and is done once by
// k_resolve.  All arithmetic that reaches an output follows the numerics contract, so results are
// bit-identical to the CPU oracle.
#pragma once
#include <cuda_fp16.h>

#include "f3d_aether.cuh"
#include "f3d_trace_fast.cuh"

namespace f3d {

constexpr int kTileW = 16, kTileH = 16;      // CTA pixel tile; warps own 8x4 sub-tiles

// Wavefront buffers of ONE step (one camera sample of one frame).  A launch of k_ascent / k_trace / k_accum serves a BATCH of
// consecutive steps (FrameParams::slot[0 .. n_batch)): k_primary of step i+1 only needs k_primary of step i (the reservoir
// chain), so the primaries of a batch run back to back and one launch of each later kernel handles all their rays - on an
// image partition (1/8 of a 1080p frame is ~230 k rays) a single step cannot fill 148 SMs, a batch can.
constexpr int kMaxBatch = 8;
constexpr uint32_t kQCounts = 12u;     // counters per BatchSlot (q_counts)
struct BatchSlot {
    float4* prim;              // split primary pass (k_ptrace -> k_shade): 2 x float4 per pixel, (d.xyz, t) (hit, cell, rng, -)
    float4* rec;               // 4 x float4 per pixel, see PixelRec
    uint8_t* occl_sun;         // per pixel: sun ray occluded
    uint8_t* occl_ibl;         // per pixel: IBL ray occluded
    uint32_t* q_sun;           // compacted pixel indices that need a sun ray
    uint32_t* q_ibl;           // compacted pixel indices that need an IBL ray
    uint32_t* q_counts;        // [0] n_sun [1] n_ibl [2] next_sun [3] next_ibl; stage 2: [4] n2_sun [5] n2_ibl [6] next2_sun [7] next2_ibl; far lists: [8] nf_sun [9] nf_ibl
    uint32_t* q2_sun;          // stage-2 lists: rays k_ascent could not decide (pixel index; seeds in qn_*)
    uint32_t* q2_ibl;
    unsigned long long* qn_sun;   // per stage-2 sun entry: seeds of the bottom-up start (ascent_seeds; k_ascent -> k_trace)
    unsigned long long* qn_ibl;   // same for the IBL list
    uint32_t* qf_sun;          // far lists (k_ascent pass 0 -> pass 1): undecided rays no near-field walk can decide
    uint32_t* qf_ibl;
};

struct FrameParams {
    SceneParams scene;          // env / mesh / albedo (+ the plain pyramid when the KAT seam keeps it)
    FastScene fast;             // quad-packed pyramid for the production traversal
    uint32_t stack_depth;       // shared-memory stack entries per thread (3 * mip_count + 2)
    SunHorizon hz;              // sun horizon strips (F3D_SUN_HORIZON); hz.S == nullptr: none
    EscapeMap esc;              // per-cell, per-octant horizon slopes (F3D_ESCAPE); esc.E == nullptr: none
    StageParams stage[3];       // TMA staging of the top pyramid levels (F3D_TMA_STAGE): [0] k_ptrace [1] k_ascent [2] k_trace
    uint32_t W, H, frame_index, spp, window;
    float cam_origin[3], cam_right[3], cam_up[3], cam_forward[3];
    float half_w, half_h, exposure;
    uint32_t seed_hi, seed_lo;
    float light_dir[3], light_color[3];
    // image-row partition: this device owns row blocks b with b % world == rank
    uint32_t part_rank, part_world, block_rows, tiles_per_block, nblocks;
    uint32_t part_mode;         // 1 = gather-only partition: spatial reuse never leaves the row block (no halo, no frame flags)
    // per-pixel state (full-image arrays, only owned rows + 3/4-row halos are touched)
    float4* accum;
    float2* welford;
    const float4* resv_in;     // out_{f-1}: (w_sum, weight, target_pdf, bits(m | type<<31))
    float4* resv_out;          // out_f
    const uint8_t* pixflags;   // bit0: G-buffer normal faces the sun; bits1-2: centre-ray hit type
    unsigned long long* counters;  // primary, shadow, ibl, nodes
    // wavefront state
    uint32_t sample_index;     // 0 .. spp-1 (one k_primary/k_trace/k_accum round per camera sample)
    BatchSlot cur;             // k_primary: the slot of the step it renders
    BatchSlot slot[kMaxBatch]; // k_ascent / k_trace / k_accum: the steps of the batch, in order; slot[0] is step
    uint32_t n_batch;          //   (frame_index, sample_index), step k is sample (sample_index + k) % spp of a later frame
    float4* sstate;            // spp > 1 only: 3 x float4 per pixel (rng+cand, prev, partial radiance)
    // NVLink halo push: peer images of resv_out for the rank above / below (NULL = none)
    float4* peer_up;
    float4* peer_down;
    // Cross-GPU frame barrier in peer memory (no host, no NCCL): sync_local[0] counts finished CTAs of the
    // current k_primary, sync_local[1] / [2] = frames completed by the rank above / below (written by THEM).
    // The last CTA of k_primary(f) stores f+1 into the neighbours' slots after a system-scope fence.
    uint32_t* sync_local;
    uint32_t* peer_sync_up;     // the up neighbour's sync array (I am its "below" rank -> slot 2)
    uint32_t* peer_sync_down;   // the down neighbour's sync array (I am its "above" rank -> slot 1)
    uint32_t* sync_error;       // set to 1 if a wait timed out (host raises)
    unsigned long long sync_timeout_ns;   // budget of one neighbour wait (F3D_B200_SYNC_TIMEOUT_MS, default 2000)
};

// Spin until both neighbours have completed `need` frames.  One thread per CTA polls local memory; a wall-clock budget
// (FrameParams::sync_timeout_ns, default 2 s, measured with %globaltimer) turns a dead peer into an error instead of a hung
// GPU: the flag is sticky, and f3d_session_variance / every resolve fail when it is set, so a frame shaded with stale halo
// rows can never be returned as a result.
__device__ __forceinline__ unsigned long long global_ns() {
#if defined(__CUDA_ARCH__)
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#else
    return 0ull;
#endif
}
__device__ __forceinline__ void wait_neighbours(const FrameParams& P, uint32_t need) {
    if (P.part_world > 1u && need > 0u && P.sync_local != nullptr && (P.peer_sync_up != nullptr || P.peer_sync_down != nullptr)) {
        if (threadIdx.x == 0) {
            const volatile uint32_t* f = P.sync_local;
            const unsigned long long t0 = global_ns();
            while (f[1] < need || f[2] < need) {
                if (global_ns() - t0 > P.sync_timeout_ns) { atomicExch(P.sync_error, 1u); break; }
                __nanosleep(200);
            }
            __threadfence_system();
        }
        __syncthreads();
    }
}

// Called by every CTA when its stores (incl. NVLink halo stores) are issued; the last CTA publishes
// "frame_index + 1 frames done" to both neighbours.
__device__ __forceinline__ void signal_neighbours(const FrameParams& P) {
    if (P.part_world > 1u && P.sync_local != nullptr) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const uint32_t total = gridDim.x * gridDim.y;
            const uint32_t old = atomicAdd(P.sync_local + 0, 1u);
            if (old == total - 1u) {
                P.sync_local[0] = 0u;
                __threadfence_system();
                if (P.peer_sync_up) *(volatile uint32_t*)(P.peer_sync_up + 2) = P.frame_index + 1u;
                if (P.peer_sync_down) *(volatile uint32_t*)(P.peer_sync_down + 1) = P.frame_index + 1u;
                __threadfence_system();
            }
        }
    }
}

__device__ __forceinline__ v3 ld3(const float* p) { return V3(p[0], p[1], p[2]); }

// Camera ray, hybrid_terrain_traversal.wgsl:480-485 (and :584-588, :628-632 with zero jitter).
__device__ __forceinline__ Ray camera_ray(const FrameParams& P, uint32_t gx, uint32_t gy, float jx, float jy) {
    float ndc_x = fdiv((float)gx + 0.5f + jx, (float)P.W) * 2.0f - 1.0f;
    float ndc_y = (1.0f - fdiv((float)gy + 0.5f + jy, (float)P.H)) * 2.0f - 1.0f;
    v3 rd = normalize3(V3(ndc_x * P.half_w, ndc_y * P.half_h, -1.0f));
    rd = normalize3((ld3(P.cam_right) * rd.x + ld3(P.cam_up) * rd.y) + (-ld3(P.cam_forward)) * rd.z);
    Ray r;
    r.o = ld3(P.cam_origin);
    r.tmin = 1e-3f;
    r.d = rd;
    r.tmax = 1e30f;
    return r;
}

// Maps (blockIdx, threadIdx) to a pixel of an owned row block.  Returns false when out of range.
__device__ __forceinline__ bool owned_pixel(const FrameParams& P, uint32_t& gx, uint32_t& gy) {
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5, lane = tid & 31u;
    const uint32_t lx = (warp & 1u) * 8u + (lane & 7u);
    const uint32_t ly = (warp >> 1) * 4u + (lane >> 3);
    const uint32_t owned_block = blockIdx.y / P.tiles_per_block;
    const uint32_t tile_in_block = blockIdx.y % P.tiles_per_block;
    const uint32_t b = owned_block * P.part_world + P.part_rank;
    const uint32_t in_block_y = tile_in_block * kTileH + ly;
    gx = blockIdx.x * kTileW + lx;
    gy = b * P.block_rows + in_block_y;
    return gx < P.W && gy < P.H && in_block_y < P.block_rows && b < P.nblocks;
}

struct Resv { float w_sum, weight, target_pdf; uint32_t m; bool type1; };

__device__ __forceinline__ Resv unpack_resv(float4 v) {
    Resv r;
    r.w_sum = v.x; r.weight = v.y; r.target_pdf = v.z;
    uint32_t b = __float_as_uint(v.w);
    r.m = b & 0x7FFFFFFFu;
    r.type1 = (b >> 31) != 0u;
    return r;
}
__device__ __forceinline__ float4 pack_resv(const Resv& r) {
    return make_float4(r.w_sum, r.weight, r.target_pdf, __uint_as_float(r.m | (r.type1 ? 0x80000000u : 0u)));
}

// consider_candidate (directional branch), pt_restir_spatial.wgsl:45-117.  One sun with importance 1:
// p_sel = 1/max(1,1e-8) = 1 exactly.
__device__ __forceinline__ void consider_candidate(const Resv& r, bool facing, float& Wsum, bool& chosen_type1,
                                                   float& chosen_pdf, uint32_t& seed) {
    if (r.m == 0u) return;
    if (!r.type1) return;
    if (!facing) return;
    const float p_curr = 1.0f;
    if (r.target_pdf <= 0.0f) return;
    float w = r.w_sum * fdiv(p_curr, fmaxf(r.target_pdf, 1e-6f));
    if (w <= 0.0f) return;
    Wsum = Wsum + w;
    float u = xorshift32(seed);
    if (u < fdiv(w, Wsum)) {
        chosen_type1 = true;
        chosen_pdf = p_curr;
    }
}

// pt_restir_spatial::main, pt_restir_spatial.wgsl:158-224, reading the compact `out` records.
__device__ __forceinline__ Resv spatial_reuse(const FrameParams& P, const float4* __restrict__ resv, uint32_t x,
                                              uint32_t y, bool facing, uint32_t pass_frame) {
    const uint32_t W = P.W, H = P.H;
    const uint32_t idx = y * W + x;
    uint32_t seed = (P.seed_hi ^ pass_frame) + idx * 1664525u + 1013904223u;
    const Resv r_self = unpack_resv(__ldcg(resv + idx));
    bool chosen_type1 = r_self.type1;
    float chosen_pdf = r_self.target_pdf;
    float Wsum = 0.0f;
    uint32_t m_total = 0u;
    consider_candidate(r_self, facing, Wsum, chosen_type1, chosen_pdf, seed);
    m_total += r_self.m;
#pragma unroll 1
    for (uint32_t i = 0; i < 8u; i++) {
        int rx = (int)floorf(xorshift32(seed) * 7.0f) - 3;
        int ry = (int)floorf(xorshift32(seed) * 7.0f) - 3;
        if (rx == 0 && ry == 0) continue;
        int nxi = min(max((int)x + rx, 0), (int)W - 1);
        int nyi = min(max((int)y + ry, 0), (int)H - 1);
        if (P.part_mode == 1u && P.part_world > 1u) {          // gather-only partition: clamp to the rows of this block
            const int y0 = (int)((y / P.block_rows) * P.block_rows);
            nyi = min(max(nyi, y0), min(y0 + (int)P.block_rows, (int)H) - 1);
        }
        const Resv rn = unpack_resv(__ldcg(resv + (uint32_t)nyi * W + (uint32_t)nxi));
        consider_candidate(rn, facing, Wsum, chosen_type1, chosen_pdf, seed);
        m_total += rn.m;
    }
    Resv o;
    o.type1 = chosen_type1;
    o.target_pdf = chosen_pdf;
    o.w_sum = Wsum;
    o.m = m_total;
    o.weight = (o.w_sum > 0.0f && o.target_pdf > 0.0f) ? fdiv(o.w_sum, (float)o.m * o.target_pdf) : 0.0f;
    return o;
}

// pt_restir_temporal::main, pt_restir_temporal.wgsl:54-109
__device__ __forceinline__ Resv temporal_reuse(const Resv& rp, const Resv& rc) {
    const bool prev_valid = rp.m > 0u && rp.weight > 0.0f && rp.target_pdf > 0.0f;
    const bool curr_valid = rc.m > 0u && rc.weight > 0.0f && rc.target_pdf > 0.0f;
    if (!prev_valid) return rc;
    if (!curr_valid) return rp;
    Resv ro;
    const bool choose_prev = rp.weight > rc.weight;
    ro.type1 = choose_prev ? rp.type1 : rc.type1;
    ro.target_pdf = choose_prev ? rp.target_pdf : rc.target_pdf;
    ro.m = rp.m + rc.m;
    ro.w_sum = rp.w_sum + rc.w_sum;
    ro.weight = (ro.w_sum > 0.0f && ro.target_pdf > 0.0f) ? fdiv(ro.w_sum, (float)ro.m * ro.target_pdf) : 0.0f;
    return ro;
}

__device__ __forceinline__ void warp_add_counters(unsigned long long* counters, uint32_t c0, uint32_t c1, uint32_t c2,
                                                  uint32_t c3) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        c0 += __shfl_xor_sync(0xFFFFFFFFu, c0, off);
        c1 += __shfl_xor_sync(0xFFFFFFFFu, c1, off);
        c2 += __shfl_xor_sync(0xFFFFFFFFu, c2, off);
        c3 += __shfl_xor_sync(0xFFFFFFFFu, c3, off);
    }
    if ((threadIdx.x & 31u) == 0u) {
        if (c0) atomicAdd(counters + 0, (unsigned long long)c0);
        if (c1) atomicAdd(counters + 1, (unsigned long long)c1);
        if (c2) atomicAdd(counters + 2, (unsigned long long)c2);
        if (c3) atomicAdd(counters + 3, (unsigned long long)c3);
    }
}

constexpr int kThreads = kTileW * kTileH;

// Dynamic shared memory of every traversal kernel: stack_depth x blockDim u32, [depth][thread].
__host__ __device__ inline size_t stack_smem_bytes(uint32_t stack_depth, int threads) {
    return (size_t)stack_depth * threads * 4;
}
// k_trace adds one leaf ring per warp behind the stacks (see F3D_TRACE_LEAF_QUEUE, kLeafQBU x kLeafQFields words).
__host__ __device__ inline size_t trace_smem_bytes_for(uint32_t stack_depth, int threads) {
    return stack_smem_bytes(stack_depth, threads) + (size_t)(threads / 32) * (256 * 3) * 4;
}

// Per-pixel record written by k_primary (4 x float4, 128-bit accesses):
//   rec[4*pix+0] = (shade_o.xyz, bits(flags))   shade_o = hit.point + n * 1e-3   (:526,:540)
//   rec[4*pix+1] = (ei.xyz, reuse_w)            IBL direction, clamped reuse weight (:521,:539)
//   rec[4*pix+2] = (sun_pre.xyz, ibl_pre.x)     sun_pre = albedo*light_color*nd (:531 before vis)
//   rec[4*pix+3] = (ibl_pre.yz, 0, 0)           ibl_pre = albedo*env(ei) (:545 before env_vis); for a
//                                               miss the slots hold the sky radiance (:489)
constexpr uint32_t kRecHit = 1u, kRecSun = 2u, kRecSunReuseDir = 4u;

// Cache policy switch for once-per-frame per-pixel state.  Streaming (evict-first) accesses were measured
// SLOWER (1.43 vs 1.36 ms/frame: the records are re-read from L2 by the next kernel), so the default is
// ordinary L2-cached accesses; pinning the pyramid with an L2 access-policy window made no difference.
#ifndef F3D_STREAMING_STATE
#define F3D_STREAMING_STATE 0
#endif
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
#if F3D_STREAMING_STATE
    __stcs(p, v);
#else
    *p = v;
#endif
}
__device__ __forceinline__ void st_stream(float2* p, float2 v) {
#if F3D_STREAMING_STATE
    __stcs(p, v);
#else
    *p = v;
#endif
}
__device__ __forceinline__ float4 ld_stream(const float4* p) {
#if F3D_STREAMING_STATE
    return __ldcs(p);
#else
    return __ldcg(p);
#endif
}
__device__ __forceinline__ float2 ld_stream(const float2* p) {
#if F3D_STREAMING_STATE
    return __ldcs(p);
#else
    return *p;
#endif
}

// intersect_hybrid (hybrid_traversal.wgsl:175-201) on the production traversal.
struct PrimaryHit { bool hit; uint32_t hit_type; float t; v3 point, normal; };

// Warp-cooperative: call from converged code; `valid` = this lane has a ray.
__device__ __forceinline__ PrimaryHit primary_hit(const FrameParams& P, const Ray& ray, bool valid, const SmemStack st, uint32_t& nodes) {
    PrimaryHit ph;
    ph.hit = false; ph.hit_type = 0u; ph.t = ray.tmax; ph.point = V3(0, 0, 0); ph.normal = V3(0, 0, 0);
    Ray tr = ray;
    if (valid && P.scene.traversal_mode == 0u) {
        const Hit mh = intersect_mesh(P.scene, ray);
        if (mh.hit && mh.t < ph.t) { ph.hit = true; ph.hit_type = 0u; ph.t = mh.t; ph.point = mh.point; ph.normal = mh.normal; }
        tr.tmax = ph.t;
    }
    const FastHit fh = trace_fast<false, false>(P.fast, tr, valid, st, nodes);
    if (fh.hit && fh.t < ph.t) {
        ph.hit = true; ph.hit_type = 3u; ph.t = fh.t;
        finish_hit(P.fast, tr, fh, ph.point, ph.normal);
    }
    return ph;
}

// ---------------------------------------------------------------------------------------------
// k_primary: per pixel, one camera sample.
// ---------------------------------------------------------------------------------------------
#ifndef F3D_PRIMARY_MIN_CTAS
#define F3D_PRIMARY_MIN_CTAS 4
#endif
// Everything of one camera sample that follows the primary traversal: reuse chain, shading set-up, candidate, temporal
// reuse, publication of out_f (+ NVLink halo rows), compaction of the secondary-ray requests.  Called converged by all
// threads of the CTA (it contains the neighbour wait / signal and warp ballots).
__device__ __forceinline__ void shade_and_publish(const FrameParams& P, const bool active, const uint32_t gx, const uint32_t gy,
                                                  const uint32_t pix, const Ray& ray, const PrimaryHit& hit, uint32_t rng,
                                                  uint32_t n_primary, uint32_t n_nodes) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t spp = max(P.spp, 1u), s = P.sample_index;
    const bool multi = spp > 1u;
    const SceneParams& S = P.scene;
    const v3 light_color = ld3(P.light_color);
    const v3 wi = normalize3(ld3(P.light_dir));
    bool want_sun = false, want_ibl = false;
    Resv prev_r, cand;
    prev_r.w_sum = 0.0f; prev_r.weight = 0.0f; prev_r.target_pdf = 0.0f; prev_r.m = 0u; prev_r.type1 = false;
    cand = prev_r;
    // ---- merged reservoir from last frame's reuse chain + M-clamp (:452-465).  Done AFTER the primary
    // traversal: only the shading below needs it, so the wait for the neighbour GPUs' halo rows (and the
    // 9 record loads) overlaps with the traversal instead of preceding it. ----
    if (s == 0u) wait_neighbours(P, P.frame_index);
    bool prev_valid = false;
    v3 sun_dir = wi;
    float reuse_w = 1.0f;
    if (active) {
        if (s == 0u) {
            const bool facing = (P.pixflags[pix] & 1u) != 0u;
            if (P.frame_index > 0u) prev_r = spatial_reuse(P, P.resv_in, gx, gy, facing, P.frame_index - 1u);
            if (prev_r.m > 512u) {
                const float scale = fdiv(512.0f, (float)prev_r.m);
                prev_r.w_sum = prev_r.w_sum * scale;
                prev_r.m = 512u;
                if (prev_r.target_pdf > 0.0f) prev_r.weight = fdiv(prev_r.w_sum, (float)prev_r.m * prev_r.target_pdf);
            }
            if (multi) P.sstate[3 * (size_t)pix + 1] = pack_resv(prev_r);
        } else {
            const float4 a = P.sstate[3 * (size_t)pix + 0];
            cand.w_sum = a.y;
            const uint32_t mb = __float_as_uint(a.z);
            cand.m = mb & 0x7FFFFFFFu; cand.type1 = (mb >> 31) != 0u;
            cand.target_pdf = a.w;
            prev_r = unpack_resv(P.sstate[3 * (size_t)pix + 1]);
        }
        prev_valid = P.frame_index > 0u && prev_r.m > 0u && prev_r.weight > 0.0f && prev_r.target_pdf > 0.0f && prev_r.type1;
        // every populated sample stores direction == wi; the shader re-normalises it (:520)
        sun_dir = prev_valid ? normalize3(wi) : wi;
        reuse_w = prev_valid ? clampf(prev_r.weight, 0.0f, 4.0f) : 1.0f;
        float4* rec = P.cur.rec + 4 * (size_t)pix;
        if (!hit.hit) {
            const v3 sky = env_radiance(S, ray.d);
            st_stream(rec + 0, make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u)));
            st_stream(rec + 2, make_float4(0.0f, 0.0f, 0.0f, sky.x));
            st_stream(rec + 3, make_float4(sky.y, sky.z, 0.0f, 0.0f));
        } else {
            // ---- shading set-up (:492-545 without the two visibility factors) ----
            const v3 n = hit.normal;
            const v3 albedo = hit.hit_type == 3u ? ld3(S.albedo) : V3(0.7f, 0.7f, 0.8f);
            const float ndotl = fmaxf(dot3(n, wi), 0.0f);
            const float target_pdf = luminance((albedo * light_color) * ndotl);
            if (target_pdf > 0.0f) {
                cand.type1 = true;
                cand.w_sum = cand.w_sum + target_pdf;
                cand.m = cand.m + 1u;
                cand.target_pdf = target_pdf;
            }
            const float nd = fmaxf(dot3(n, sun_dir), 0.0f);
            const v3 shade_o = hit.point + n * 1e-3f;
            v3 sun_pre = V3(0, 0, 0);
            uint32_t flags = kRecHit;
            if (nd > 0.0f) {
                want_sun = true;
                sun_pre = (albedo * light_color) * nd;
                flags |= kRecSun | (prev_valid ? kRecSunReuseDir : 0u);
            }
            const float u1 = xorshift32(rng);
            const float u2 = xorshift32(rng);
            const v3 ei = cosine_dir(n, u1, u2);
            want_ibl = true;
            const v3 ibl_pre = albedo * env_radiance(S, ei);
            __stcg(rec + 0, make_float4(shade_o.x, shade_o.y, shade_o.z, __uint_as_float(flags)));   // re-read by k_trace
            __stcg(rec + 1, make_float4(ei.x, ei.y, ei.z, reuse_w));
            st_stream(rec + 2, make_float4(sun_pre.x, sun_pre.y, sun_pre.z, ibl_pre.x));
            st_stream(rec + 3, make_float4(ibl_pre.y, ibl_pre.z, 0.0f, 0.0f));
        }

        if (s + 1u < spp) {
            P.sstate[3 * (size_t)pix + 0] = make_float4(__uint_as_float(rng), cand.w_sum,
                                                        __uint_as_float(cand.m | (cand.type1 ? 0x80000000u : 0u)), cand.target_pdf);
        } else {
            // ---- finalise candidate, temporal reuse, publish out_f (+ NVLink halo push) (:551-555) ----
            if (cand.m > 0u && cand.w_sum > 0.0f && cand.target_pdf > 0.0f)
                cand.weight = fdiv(cand.w_sum, (float)cand.m * cand.target_pdf);
            const float4 out_rec = pack_resv(temporal_reuse(prev_r, cand));
            __stcg(P.resv_out + pix, out_rec);
            if (P.part_world > 1u) {
                const uint32_t b = gy / P.block_rows, in_y = gy - b * P.block_rows;
                // The spatial pass draws offsets floor(u*7)-3 with u in [0,1]: xorshift32 can return exactly 1.0
                // (SURVEY section 9.3), so offsets span [-3, +4]: the block ABOVE needs our first FOUR rows, the
                // block below our last three.
                if (in_y < 4u && b > 0u && P.peer_up) __stcg(P.peer_up + pix, out_rec);
                if (in_y + 3u >= P.block_rows && b + 1u < P.nblocks && P.peer_down) __stcg(P.peer_down + pix, out_rec);
            }
        }
    }
    // ---- ray compaction: append this pixel to the global sun / IBL lists ----
    {
        const uint32_t m_sun = __ballot_sync(0xFFFFFFFFu, want_sun), m_ibl = __ballot_sync(0xFFFFFFFFu, want_ibl);
        uint32_t b_sun = 0u, b_ibl = 0u;
        if (lane == 0u) {
            if (m_sun) b_sun = atomicAdd(P.cur.q_counts + 0, (uint32_t)__popc(m_sun));
            if (m_ibl) b_ibl = atomicAdd(P.cur.q_counts + 1, (uint32_t)__popc(m_ibl));
        }
        b_sun = __shfl_sync(0xFFFFFFFFu, b_sun, 0);
        b_ibl = __shfl_sync(0xFFFFFFFFu, b_ibl, 0);
        const uint32_t lt = (1u << lane) - 1u;
        if (want_sun) P.cur.q_sun[b_sun + __popc(m_sun & lt)] = pix;
        if (want_ibl) P.cur.q_ibl[b_ibl + __popc(m_ibl & lt)] = pix;
    }
    warp_add_counters(P.counters, n_primary, 0u, 0u, n_nodes);
    if (s + 1u == spp) signal_neighbours(P);
}

// The camera ray of one sample and the state of its RNG stream after the two jitter draws (:467-485).
__device__ __forceinline__ Ray sample_ray(const FrameParams& P, uint32_t gx, uint32_t gy, uint32_t pix, uint32_t frame, uint32_t s, uint32_t& rng) {
    if (s == 0u) rng = P.seed_hi ^ (gx * 1664525u) ^ (gy * 1013904223u) ^ (frame * 92837111u) ^ P.seed_lo;
    else rng = __float_as_uint(P.sstate[3 * (size_t)pix + 0].x);
    const float jx = tent_offset(xorshift32(rng)) * 0.5f;
    const float jy = tent_offset(xorshift32(rng)) * 0.5f;
    return camera_ray(P, gx, gy, jx, jy);
}

__global__ void __launch_bounds__(kThreads, F3D_PRIMARY_MIN_CTAS) k_primary(const __grid_constant__ FrameParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemStack st;
    st.base = reinterpret_cast<uint32_t*>(smem_raw) + threadIdx.x;
    st.stride = kThreads;
    uint32_t gx, gy;
    const bool active = owned_pixel(P, gx, gy);
    const uint32_t pix = active ? gy * P.W + gx : 0u;
    uint32_t n_primary = 0, n_nodes = 0, rng = 0u;
    Ray ray;
    ray.o = V3(0, 0, 0); ray.d = V3(0, 0, 1); ray.tmin = 1e-3f; ray.tmax = 1e30f;
    if (active) {
        // ---- primary ray (:467-485): its RNG stream does not depend on the reuse chain ----
        ray = sample_ray(P, gx, gy, pix, P.frame_index, P.sample_index, rng);
        n_primary++;
    }
    const PrimaryHit hit = primary_hit(P, ray, active, st, n_nodes);   // warp-cooperative
    shade_and_publish(P, active, gx, gy, pix, ray, hit, rng, n_primary, n_nodes);
}

// ---------------------------------------------------------------------------------------------
// Split primary pass (terrain-only scenes, spp = 1): the primary TRAVERSAL of a frame depends on nothing but the camera
// and the frame number, only the reuse chain behind it is sequential.  k_ptrace traces the primary rays of a whole BATCH of
// frames in one launch (blockIdx.z = step of the batch) and leaves a 32-byte record per pixel; k_shade - no traversal, no
// stack, a few hundred instructions per pixel - then runs once per frame in order.  On an image partition this moves ~70 %
// of the per-frame chain (and its launch + tail) into a launch that is 4 frames big.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, F3D_PRIMARY_MIN_CTAS) k_ptrace(const __grid_constant__ FrameParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemStack st;
    st.base = reinterpret_cast<uint32_t*>(smem_raw) + threadIdx.x;
    st.stride = kThreads;
    uint32_t gx, gy;
    const bool active = owned_pixel(P, gx, gy);
    const uint32_t pix = active ? gy * P.W + gx : 0u;
    const BatchSlot& B = P.slot[blockIdx.z];
#if F3D_TMA_STAGE
    // top pyramid levels -> shared memory, one bulk async copy per CTA (behind the stacks)
    const QStaged qs = stage_top_levels(P.fast, P.stage[0], smem_raw + ((stack_smem_bytes(P.stack_depth, kThreads) + 15) & ~(size_t)15));
#endif
    uint32_t n_primary = 0, n_nodes = 0, rng = 0u;
    Ray ray;
    ray.o = V3(0, 0, 0); ray.d = V3(0, 0, 1); ray.tmin = 1e-3f; ray.tmax = 1e30f;
    if (active) {
        ray = sample_ray(P, gx, gy, pix, P.frame_index + blockIdx.z, 0u, rng);
        n_primary++;
    }
#if F3D_TMA_STAGE
    const FastHit fh = trace_fast<false, false, QStaged>(P.fast, ray, active, st, n_nodes, qs);   // warp-cooperative
#else
    const FastHit fh = trace_fast<false, false>(P.fast, ray, active, st, n_nodes);   // warp-cooperative
#endif
    if (active) {
        st_stream(B.prim + 2 * (size_t)pix, make_float4(ray.d.x, ray.d.y, ray.d.z, fh.t));
        st_stream(B.prim + 2 * (size_t)pix + 1, make_float4(__uint_as_float(fh.hit ? 1u : 0u), __uint_as_float(fh.cx | (fh.cz << 13)),
                                                             __uint_as_float(rng), 0.0f));
    }
    warp_add_counters(P.counters, n_primary, 0u, 0u, n_nodes);
}

__global__ void __launch_bounds__(kThreads) k_shade(const __grid_constant__ FrameParams P) {
    uint32_t gx, gy;
    const bool active = owned_pixel(P, gx, gy);
    const uint32_t pix = active ? gy * P.W + gx : 0u;
    uint32_t rng = 0u;
    Ray ray;
    ray.o = V3(0, 0, 0); ray.d = V3(0, 0, 1); ray.tmin = 1e-3f; ray.tmax = 1e30f;
    PrimaryHit hit;
    hit.hit = false; hit.hit_type = 0u; hit.t = ray.tmax; hit.point = V3(0, 0, 0); hit.normal = V3(0, 0, 0);
    if (active) {
        const float4 p0 = ld_stream(P.cur.prim + 2 * (size_t)pix), p1 = ld_stream(P.cur.prim + 2 * (size_t)pix + 1);
        ray.o = ld3(P.cam_origin);
        ray.d = V3(p0.x, p0.y, p0.z);
        rng = __float_as_uint(p1.z);
        if (__float_as_uint(p1.x) != 0u) {          // intersect_hybrid's terrain branch (:190-197), as primary_hit
            FastHit fh;
            fh.hit = true; fh.t = p0.w;
            const uint32_t c = __float_as_uint(p1.y);
            fh.cx = c & 0x1FFFu; fh.cz = c >> 13;
            hit.hit = true; hit.hit_type = 3u; hit.t = fh.t;
            finish_hit(P.fast, ray, fh, hit.point, hit.normal);
        }
    }
    shade_and_publish(P, active, gx, gy, pix, ray, hit, rng, 0u, 0u);
}

// Orders the final reuse pass (k_resolve) after the neighbours' last halo stores.
__global__ void k_wait_peers(const __grid_constant__ FrameParams P, uint32_t need) { wait_neighbours(P, need); }

// ---------------------------------------------------------------------------------------------
// k_trace<IS_SUN, CURV>: persistent any-hit traversal over one compacted ray list
// (intersect_shadow_ray / intersect_ibl_occlusion_ray, hybrid_traversal.wgsl:204-259).
// Each lane owns one ray at a time; when fewer than kRefillBelow lanes of a warp still traverse,
// the idle lanes pull fresh rays (one atomic per warp).
// ---------------------------------------------------------------------------------------------
#ifndef F3D_REFILL_BELOW
#define F3D_REFILL_BELOW 16
#endif
#ifndef F3D_TRACE_MIN_CTAS
#define F3D_TRACE_MIN_CTAS 6
#endif
#ifndef F3D_TRACE_THREADS
#define F3D_TRACE_THREADS 128
#endif
constexpr int kTraceCtaThreads = F3D_TRACE_THREADS;
constexpr int kRefillBelow = F3D_REFILL_BELOW;
// F3D_TRACE_DEFER_LEAVES = N > 0: an any-hit lane that finds a leaf on top of its stack parks it in one of N per-lane
// slots and keeps expanding; parked leaves are solved together once F3D_DEFER_LEAF_BATCH lanes hold one (or nobody can
// expand any more).  Exact for the occlusion flag, which is an OR over a fixed set of leaf tests (see
// F3D_ANYHIT_SIGN_ORDER in f3d_trace_fast.cuh): only the ORDER of those tests changes.  0 = solve leaves in stack order.
#ifndef F3D_TRACE_DEFER_LEAVES
#define F3D_TRACE_DEFER_LEAVES 0
#endif
#ifndef F3D_DEFER_LEAF_BATCH
#define F3D_DEFER_LEAF_BATCH 20
#endif
#ifndef F3D_DEFER_RULE
#define F3D_DEFER_RULE 0
#endif
#ifndef F3D_DEFER_LEAF_MIN
#define F3D_DEFER_LEAF_MIN 8
#endif
#ifndef F3D_DEFER_STALL
#define F3D_DEFER_STALL 8
#endif
constexpr int kDeferLeaves = F3D_TRACE_DEFER_LEAVES;
constexpr int kDeferLeafBatch = F3D_DEFER_LEAF_BATCH;
#if F3D_TRACE_DEFER_LEAVES
// The parked-leaf slots stay in registers: every access is a compile-time index selected by a predicate.
__device__ __forceinline__ void pend_push(uint32_t (&pend)[F3D_TRACE_DEFER_LEAVES], uint32_t& n, uint32_t v) {
#pragma unroll
    for (int k = 0; k < F3D_TRACE_DEFER_LEAVES; k++)
        if (n == (uint32_t)k) pend[k] = v;
    n++;
}
__device__ __forceinline__ uint32_t pend_pop(const uint32_t (&pend)[F3D_TRACE_DEFER_LEAVES], uint32_t& n) {
    n--;
    uint32_t v = pend[0];
#pragma unroll
    for (int k = 1; k < F3D_TRACE_DEFER_LEAVES; k++)
        if (n == (uint32_t)k) v = pend[k];
    return v;
}
#endif

// F3D_SCHED_STATS (test / tuning builds only): SIMT-efficiency counters of the persistent scheduler, one warp-level event
// each: [0] expansion steps [1] lanes expanding [2] leaf phases [3] lanes solving a leaf [4] refill steps [5] rays fetched.
// Read with f3d_debug_sched_stats(); the CPU emulator (real 32-lane warps) gives the same counts the GPU would.
#ifdef F3D_SCHED_STATS
__device__ unsigned long long g_sched_stats[40];    // [8 + kidx]: sun rays per horizon index, [24 + kidx]: their seeds
#define F3D_SCHED_STAT(i, mask)                                                                                  \
    do {                                                                                                          \
        const uint32_t _m = __ballot_sync(0xFFFFFFFFu, (mask));                                                   \
        if ((threadIdx.x & 31u) == 0u && _m != 0u) {                                                              \
            atomicAdd(&g_sched_stats[i], 1ull);                                                                   \
            atomicAdd(&g_sched_stats[(i) + 1], (unsigned long long)__popc(_m));                                   \
        }                                                                                                         \
    } while (0)
#else
#define F3D_SCHED_STAT(i, mask) do {} while (0)
#endif

// F3D_TRACE_LEAF_QUEUE = 1 (default; 0 = round-1 scheduling: a lane solves its own leaves, >= kLeafBatch lanes at a time).
//   The patch solve is the one expensive step left (IEEE divisions, ~450 instructions) and under the round-1 scheduler it ran
//   with ~8 of 32 lanes (ncu, profiles/r01_source_regions.txt).  Here leaves are WORK ITEMS of the warp: a lane that pops a
//   leaf appends (lane, cell) to a 64-entry ring in shared memory and keeps expanding; once 32 items wait (or nobody can
//   expand) all 32 lanes take one item each, fetch the owner's ray with shuffles and solve it, and the hits are OR-ed back
//   to the owners.  Exact for the occlusion flag: it is an OR over a fixed set of leaf tests (see F3D_CULL_FAST (1)-(2));
//   only the order of the tests and the lane that runs them change.  A lane keeps its ray until its last item has been
//   served (my_last <= q_head), so an item always finds its owner's ray in the owner's registers.
#ifndef F3D_TRACE_LEAF_QUEUE
#define F3D_TRACE_LEAF_QUEUE 1
#endif
#ifndef F3D_LEAFQ_WAIT_DRAIN
#define F3D_LEAFQ_WAIT_DRAIN 8      // serve a partial batch once this many lanes only wait for their leaves
#endif
constexpr uint32_t kLeafQ = 64u;    // ring entries per warp (power of two, >= 2 * 32 - 1)

template <bool IS_SUN, bool CURV, bool EXACT_CULL>
__device__ __forceinline__ void trace_list(const FrameParams& P, const BatchSlot& B, const SmemStack st, uint32_t* wq) {
    const uint32_t lane = threadIdx.x & 31u;
    const FastScene& F = P.fast;
    const uint32_t n = B.q_counts[IS_SUN ? 0 : 1];
    const uint32_t* __restrict__ queue = IS_SUN ? B.q_sun : B.q_ibl;
    uint32_t* next = B.q_counts + (IS_SUN ? 2 : 3);
    uint8_t* __restrict__ occl = IS_SUN ? B.occl_sun : B.occl_ibl;
    const v3 wi = normalize3(ld3(P.light_dir));
    const v3 wi_reuse = normalize3(wi);
    const bool has_mesh = P.scene.traversal_mode == 0u;

    TraceState T{};
    bool busy = false, mesh_occl = false;
    uint32_t pix = 0u;
    bool exhausted = false;
    uint32_t n_rays = 0, n_nodes = 0;
#if F3D_TRACE_LEAF_QUEUE
    uint32_t q_head = 0u, q_tail = 0u;      // absolute item counters, warp-uniform
    uint32_t my_last = 0u;                  // 1 + sequence number of this lane's newest item
    bool decided_hit = false;               // this lane's ray is already known to be occluded
    // Serves the first `count` (<= 32) queued leaves, one per lane.  Called by all 32 lanes, converged.
    auto serve = [&](uint32_t count) {
        const bool have = lane < count;
        const uint32_t item = have ? wq[(q_head + lane) & (kLeafQ - 1u)] : 0u;
        const int owner = (int)(item >> 26);
        const uint32_t dead = __ballot_sync(0xFFFFFFFFu, decided_hit);
        TraceState L;                       // the owner's ray, fetched from its registers
        L.o.x = __shfl_sync(0xFFFFFFFFu, T.o.x, owner); L.o.y = __shfl_sync(0xFFFFFFFFu, T.o.y, owner);
        L.o.z = __shfl_sync(0xFFFFFFFFu, T.o.z, owner); L.d.x = __shfl_sync(0xFFFFFFFFu, T.d.x, owner);
        L.d.y = __shfl_sync(0xFFFFFFFFu, T.d.y, owner); L.d.z = __shfl_sync(0xFFFFFFFFu, T.d.z, owner);
        L.tmax = __shfl_sync(0xFFFFFFFFu, T.tmax, owner);
        L.inv_x = __shfl_sync(0xFFFFFFFFu, T.inv_x, owner); L.inv_z = __shfl_sync(0xFFFFFFFFu, T.inv_z, owner);
        L.hd2 = __shfl_sync(0xFFFFFFFFu, T.hd2, owner);
        L.use_vertex = false; L.vertex = 0.0f; L.y_vertex = 0.0f;
        if (CURV) {
            L.vertex = __shfl_sync(0xFFFFFFFFu, T.vertex, owner); L.y_vertex = __shfl_sync(0xFFFFFFFFu, T.y_vertex, owner);
            L.use_vertex = __shfl_sync(0xFFFFFFFFu, T.use_vertex ? 1 : 0, owner) != 0;
        }
        L.tmin = 1e-3f; L.best_t = L.tmax; L.hit = false; L.sp = 0u; L.stale_sp = 0u; L.best_cx = 0u; L.best_cz = 0u;
        bool hit = false;
        const bool run = have && !((dead >> owner) & 1u);
        F3D_SCHED_STAT(2, run);
        if (run) {
            n_nodes++;
            hit = leaf_node<true, CURV>(F, L, item & 0x03FFFFFFu);
        }
        uint32_t hm = __ballot_sync(0xFFFFFFFFu, hit), mine = 0u;
        while (hm != 0u) {                  // OR the hits back to their owners (hits are rare: ~1 per batch)
            const int j = __ffs((int)hm) - 1;
            mine |= 1u << __shfl_sync(0xFFFFFFFFu, owner, j);
            hm &= hm - 1u;
        }
        if ((mine >> lane) & 1u) { decided_hit = true; T.sp = 0u; }
        q_head += count;
        __syncwarp();
    };
#endif
#if F3D_TRACE_DEFER_LEAVES
    uint32_t pend[F3D_TRACE_DEFER_LEAVES];
    uint32_t npend = 0u;
#endif

    while (true) {
        // ---- refill idle lanes ----
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !busy);
        if (idle != 0u && !exhausted) {
            const uint32_t n_idle = (uint32_t)__popc(idle);
            uint32_t base = 0u;
            if (lane == 0u) base = atomicAdd(next, n_idle);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (base + n_idle >= n) exhausted = true;
            F3D_SCHED_STAT(4, !busy && base + (uint32_t)__popc(idle & ((1u << lane) - 1u)) < n);
            if (!busy) {
                const uint32_t idx = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
                if (idx < n) {
                    pix = __ldg(queue + idx);
                    const float4 r0 = __ldcg(B.rec + 4 * (size_t)pix);
                    Ray r;
                    r.o = V3(r0.x, r0.y, r0.z);
                    r.tmin = 1e-3f;
                    r.tmax = 1e30f;
                    if (IS_SUN) r.d = (__float_as_uint(r0.w) & kRecSunReuseDir) ? wi_reuse : wi;
                    else { const float4 r1 = __ldcg(B.rec + 4 * (size_t)pix + 1); r.d = V3(r1.x, r1.y, r1.z); }
                    n_rays++;
                    mesh_occl = false;
                    bool decided = false;
                    if (has_mesh) {                      // intersect_hybrid_optimized :213-221
                        const Hit mh = intersect_mesh(P.scene, r);
                        if (mh.hit && mh.t < 0.01f) { occl[pix] = 1u; decided = true; }
                        else if (mh.hit && mh.t < r.tmax) { r.tmax = mh.t; mesh_occl = true; }
                    }
                    if (!decided) {
                        trace_begin<CURV>(F, r, T, st);
                        if (T.sp == 0u) occl[pix] = mesh_occl ? 1u : 0u;
                        else busy = true;
                    }
                }
            }
        }
        if (__ballot_sync(0xFFFFFFFFu, busy) == 0u) {
            if (exhausted) break;
            continue;
        }
#if F3D_TRACE_LEAF_QUEUE
        // ---- traverse; leaves go through the warp's work queue (see F3D_TRACE_LEAF_QUEUE) ----
        while (true) {
            // (1) move leaf tops into the queue (an expansion leaves at most 4 of them on top)
            while (true) {
                const bool is_leaf = busy && T.sp > 0u && top_is_leaf(T, st);
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, is_leaf);
                if (m == 0u) break;
                if (is_leaf) {
                    T.sp--;
                    const uint32_t slot = q_tail + (uint32_t)__popc(m & ((1u << lane) - 1u));
                    wq[slot & (kLeafQ - 1u)] = (lane << 26) | (st.at(T.sp) & 0x03FFFFFFu);
                    my_last = slot + 1u;
                }
                q_tail += (uint32_t)__popc(m);
                __syncwarp();
                if (q_tail - q_head >= 32u) serve(32u);
            }
            // (2) expand, or serve what is queued when nobody can expand / too many lanes only wait for leaves
            const uint32_t m_exp = __ballot_sync(0xFFFFFFFFu, busy && T.sp > 0u);
            const uint32_t m_wait = __ballot_sync(0xFFFFFFFFu, busy && T.sp == 0u && my_last > q_head);
            if (q_tail != q_head && (m_exp == 0u || __popc(m_wait) >= F3D_LEAFQ_WAIT_DRAIN)) serve(q_tail - q_head);
            const bool can_expand = busy && T.sp > 0u;        // after serve(): a hit empties the owner's stack
            F3D_SCHED_STAT(0, can_expand);
            if (can_expand) {
                expand_node<true, CURV, EXACT_CULL>(F, T, st);
                n_nodes++;
            }
            // (3) retire rays: stack empty and every queued leaf of this lane served
            if (busy && T.sp == 0u && my_last <= q_head) {
                occl[pix] = (decided_hit || mesh_occl) ? 1u : 0u;
                busy = false;
                decided_hit = false;
            }
            const uint32_t live = __ballot_sync(0xFFFFFFFFu, busy);
            if (live == 0u) break;
            if (!exhausted && __popc(live) < kRefillBelow) break;
        }
#elif F3D_TRACE_DEFER_LEAVES
        // ---- traverse with parked leaves (see F3D_TRACE_DEFER_LEAVES) ----
        while (true) {
            while (true) {
#pragma unroll
                for (int k = 0; k < kDeferLeaves; k++)      // park leaf tops
                    if (busy && T.sp > 0u && npend < (uint32_t)kDeferLeaves && top_is_leaf(T, st)) { T.sp--; pend_push(pend, npend, st.at(T.sp)); }
                const bool can_expand = busy && T.sp > 0u && !top_is_leaf(T, st);
                F3D_SCHED_STAT(0, can_expand);
                if (can_expand) {
                    expand_node<true, CURV, EXACT_CULL>(F, T, st);
                    n_nodes++;
                    if (T.sp == 0u && npend == 0u) { occl[pix] = mesh_occl ? 1u : 0u; busy = false; }
                }
#pragma unroll
                for (int k = 0; k < kDeferLeaves; k++)
                    if (busy && T.sp > 0u && npend < (uint32_t)kDeferLeaves && top_is_leaf(T, st)) { T.sp--; pend_push(pend, npend, st.at(T.sp)); }
                const bool expandable = busy && T.sp > 0u && !top_is_leaf(T, st);
                const uint32_t m_exp = __ballot_sync(0xFFFFFFFFu, expandable);
                const uint32_t m_pend = __ballot_sync(0xFFFFFFFFu, busy && npend > 0u);
#if F3D_DEFER_RULE == 1
                if (m_exp == 0u || __popc(m_pend) >= max(__popc(m_exp), F3D_DEFER_LEAF_MIN)) break;
#elif F3D_DEFER_RULE == 2
                const uint32_t m_stall = __ballot_sync(0xFFFFFFFFu, busy && !expandable);
                if (m_exp == 0u || __popc(m_stall) >= F3D_DEFER_STALL || __popc(m_pend) >= kDeferLeafBatch) break;
#else
                if (m_exp == 0u || __popc(m_pend) >= kDeferLeafBatch) break;
#endif
            }
            F3D_SCHED_STAT(2, busy && npend > 0u);
            if (busy && npend > 0u) {
                n_nodes++;
                const bool hit = leaf_node<true, CURV>(F, T, pend_pop(pend, npend));
                if (hit) { occl[pix] = 1u; busy = false; T.sp = 0u; npend = 0u; }
                else if (T.sp == 0u && npend == 0u) { occl[pix] = mesh_occl ? 1u : 0u; busy = false; }
            }
            const uint32_t live = __ballot_sync(0xFFFFFFFFu, busy);
            if (live == 0u) break;
            if (!exhausted && __popc(live) < kRefillBelow) break;
        }
#else
        // ---- traverse until enough lanes have gone idle (bounded while-while, see trace_fast) ----
        while (true) {
            while (true) {
                const bool can_expand = busy && !top_is_leaf(T, st);
                F3D_SCHED_STAT(0, can_expand);
                if (can_expand) {
                    expand_node<true, CURV, EXACT_CULL>(F, T, st);
                    n_nodes++;
                    if (T.sp == 0u) { occl[pix] = mesh_occl ? 1u : 0u; busy = false; }
                }
                const bool expandable = busy && !top_is_leaf(T, st);
                const uint32_t m_exp = __ballot_sync(0xFFFFFFFFu, expandable);
                const uint32_t m_leaf = __ballot_sync(0xFFFFFFFFu, busy && !expandable);
                if (m_exp == 0u || __popc(m_leaf) >= kLeafBatch) break;
            }
            F3D_SCHED_STAT(2, busy && top_is_leaf(T, st));
            if (busy && top_is_leaf(T, st)) {
                n_nodes++;
                const bool hit = leaf_top<true, CURV>(F, T, st);
                if (hit || T.sp == 0u) { occl[pix] = (hit || mesh_occl) ? 1u : 0u; busy = false; }
            }
            const uint32_t live = __ballot_sync(0xFFFFFFFFu, busy);
            if (live == 0u) break;
            if (!exhausted && __popc(live) < kRefillBelow) break;
        }
#endif
    }
    warp_add_counters(P.counters, 0u, IS_SUN ? n_rays : 0u, IS_SUN ? 0u : n_rays, n_nodes);
}

// ---------------------------------------------------------------------------------------------
// Bottom-up tracer (see ascent_need in f3d_trace_fast.cuh and F3D_TRACE_LEAF_QUEUE above): the production path of k_trace.
//   * a ray starts with the cell it begins in queued as a leaf and with the `need` mask k_ascent computed for it: its
//     next node is the top of its stack, or - when the stack is empty - the parent named by the lowest set bit of `need`
//     (expanded with the child that leads back to the ray's own cell masked out);
//   * the stack holds INTERNAL nodes only: the children of a level-1 node go straight to the warp's leaf queue;
//   * MODE_ASC: every ray of the list ascends (the sun list with the sun above the horizon): monotone height tests.
// Exact for the occlusion flag: every cell that passes its own span + band test is solved (coverage: ascent_need),
// with the reference's arithmetic (leaf_node), and the flag is the OR of those solves.
// ---------------------------------------------------------------------------------------------
#ifndef F3D_TRACE_BOTTOM_UP
#define F3D_TRACE_BOTTOM_UP 1
#endif
// Leaf work items of a warp: a ring of kLeafQBU self-contained records in shared memory, structure of arrays
// [field][slot]: cell id, pixel (bit 31: the sun ray uses the re-normalised direction), tmax.  The solving lane re-reads the
// ray from the pixel's record (L1/L2 hits), so the lane that found the leaf does not have to keep its ray until then.
constexpr uint32_t kLeafQBU = 256u;       // 31 waiting + 4 x 32 from one expansion step, power of two
constexpr uint32_t kLeafQFields = 3u;

// One patch solve of a queued leaf.  Not inlined: k_trace reaches it from several places and the solve (IEEE divisions and
// their slow paths, ~600 SASS instructions) must exist once per list, or the kernel outgrows the instruction cache
// (measured: 39 % of the stall samples were "no instruction" with the solve inlined at every site).
// The patch solve itself, once per CURV instance and kernel: solve_leaf_item and k_ascent's origin-cell solve share it.
template <bool CURV>
__device__ __noinline__ bool solve_cell_anyhit(const FastScene& F, float ox, float oy, float oz, float dx, float dy, float dz, float tmax, uint32_t cell) {
    Ray r;
    r.o = V3(ox, oy, oz); r.d = V3(dx, dy, dz); r.tmin = 1e-3f; r.tmax = tmax;
    TraceState L;
    leaf_ray_setup<CURV>(F, r, L);
    return leaf_node<true, CURV>(F, L, cell);
}
template <bool CURV>
__device__ __forceinline__ bool solve_origin_cell(const FastScene& F, const TraceState& T, uint32_t cell) {
    return solve_cell_anyhit<CURV>(F, T.o.x, T.o.y, T.o.z, T.d.x, T.d.y, T.d.z, T.tmax, cell);
}

template <bool IS_SUN, bool CURV>
__device__ __forceinline__ bool solve_leaf_item(const FrameParams& P, const float4* __restrict__ rec, uint32_t cell, uint32_t pixw, float tmax) {
    const uint32_t pix = pixw & 0x7FFFFFFFu;
    const float4 r0 = __ldcg(rec + 4 * (size_t)pix);
    Ray r;
    r.o = V3(r0.x, r0.y, r0.z);
    r.tmin = 1e-3f;
    r.tmax = tmax;
    if (IS_SUN) {
        const v3 wi = normalize3(ld3(P.light_dir));
        r.d = (pixw >> 31) ? normalize3(wi) : wi;
    } else { const float4 r1 = __ldcg(rec + 4 * (size_t)pix + 1); r.d = V3(r1.x, r1.y, r1.z); }
    return solve_cell_anyhit<CURV>(P.fast, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, r.tmax, cell);
}

template <bool IS_SUN, bool CURV, bool ASC, class Q = QGlobal>
__device__ __forceinline__ void trace_list_bu(const FrameParams& P, const BatchSlot& B, const SmemStack st, uint32_t* wq, const Q qsrc = Q()) {
    const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const FastScene& F = P.fast;
    const uint32_t n = B.q_counts[IS_SUN ? 4 : 5];
    const uint32_t* __restrict__ queue = IS_SUN ? B.q2_sun : B.q2_ibl;
    const unsigned long long* __restrict__ qseeds = IS_SUN ? B.qn_sun : B.qn_ibl;
    uint32_t* next = B.q_counts + (IS_SUN ? 6 : 7);
    uint8_t* __restrict__ occl = IS_SUN ? B.occl_sun : B.occl_ibl;
    const v3 wi = normalize3(ld3(P.light_dir));
    const v3 wi_reuse = normalize3(wi);
    const bool has_mesh = P.scene.traversal_mode == 0u;

    TraceState T{};
    bool busy = false, exhausted = false;
    uint32_t pixw = 0u;                     // pixel | reuse-direction flag << 31
    uint32_t n_nodes = 0;
    uint32_t q_head = 0u, q_tail = 0u;      // absolute item counters, warp-uniform

    // Solves the first `count` (<= 32) queued leaves, one per lane; a hit marks the pixel occluded (k_ascent wrote the
    // "not occluded" default) and stops the lane that still traverses that ray.  Called by all 32 lanes, converged.
    auto serve = [&](uint32_t count) {
        bool hit = false;
        uint32_t hpix = 0xFFFFFFFFu;
        F3D_SCHED_STAT(2, lane < count);
        if (lane < count) {
            const uint32_t slot = (q_head + lane) & (kLeafQBU - 1u);
            hpix = wq[kLeafQBU + slot];
            n_nodes++;
            hit = solve_leaf_item<IS_SUN, CURV>(P, B.rec, wq[slot], hpix, __uint_as_float(wq[2u * kLeafQBU + slot]));
            if (hit) occl[hpix & 0x7FFFFFFFu] = 1u;
        }
        uint32_t hm = __ballot_sync(0xFFFFFFFFu, hit);
        while (hm != 0u) {                  // hits are rare (~1 per batch): tell the lane that still traverses that ray
            const int j = __ffs((int)hm) - 1;
            const uint32_t hp = __shfl_sync(0xFFFFFFFFu, hpix, j);
            if (busy && pixw == hp) { busy = false; T.sp = 0u; }
            hm &= hm - 1u;
        }
        q_head += count;
        __syncwarp();
    };
    // Appends one leaf per lane of `m` (this lane's cell: `id`).
    auto enqueue = [&](uint32_t m, bool mine, uint32_t id) {
        if (mine) {
            const uint32_t slot = (q_tail + (uint32_t)__popc(m & lt)) & (kLeafQBU - 1u);
            wq[slot] = id & 0x03FFFFFFu;
            wq[kLeafQBU + slot] = pixw;
            wq[2u * kLeafQBU + slot] = __float_as_uint(T.tmax);
        }
        q_tail += (uint32_t)__popc(m);
    };

    while (true) {
        // ---- refill idle lanes ----
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !busy);
        if (idle != 0u && !exhausted) {
            const uint32_t n_idle = (uint32_t)__popc(idle);
            uint32_t base = 0u;
            if (lane == 0u) base = atomicAdd(next, n_idle);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (base + n_idle >= n) exhausted = true;
            F3D_SCHED_STAT(4, !busy && base + (uint32_t)__popc(idle & lt) < n);
            uint32_t leaf_seeds = 0u, cell0 = 0u;
            if (!busy) {
                const uint32_t idx = base + (uint32_t)__popc(idle & lt);
                if (idx < n) {
                    const uint32_t pix = __ldg(queue + idx);
                    unsigned long long seeds = __ldg(qseeds + idx);
                    const float4 r0 = __ldcg(B.rec + 4 * (size_t)pix);
                    Ray r;
                    r.o = V3(r0.x, r0.y, r0.z);
                    r.tmin = 1e-3f;
                    r.tmax = 1e30f;
                    pixw = pix;
                    if (IS_SUN) {
                        const bool reuse = (__float_as_uint(r0.w) & kRecSunReuseDir) != 0u;
                        r.d = reuse ? wi_reuse : wi;
                        pixw |= reuse ? 0x80000000u : 0u;
                    } else { const float4 r1 = __ldcg(B.rec + 4 * (size_t)pix + 1); r.d = V3(r1.x, r1.y, r1.z); }
                    if (has_mesh) {                      // intersect_hybrid_optimized :213-221, as k_ascent did (t < 0.01 was decided there)
                        const Hit mh = intersect_mesh(P.scene, r);
                        if (mh.hit && mh.t < r.tmax) r.tmax = mh.t;
                    }
                    ray_setup<CURV>(F, r, T);
                    cell0 = origin_cell(F, r.o);
#if F3D_SUN_HORIZON
                    if (IS_SUN && ASC) {        // the cleared distance k_ascent found (see SunHorizon)
                        const uint32_t kidx = (uint32_t)(seeds >> 60);
                        const uint32_t c0maj = P.hz.xmajor ? (cell0 & 0x1FFFu) : (cell0 >> 13);
                        const int32_t q0 = P.hz.forward ? (int32_t)c0maj : (int32_t)P.hz.ncols - 1 - (int32_t)c0maj;
                        T.hz_qclear = kidx == kHzNone ? kHzNoClear : q0 + (int32_t)hz_k(kidx);
                        if (kidx != kHzNone) T.best_t = fminf(T.best_t, horizon_t_clear(P.hz, T, T.hz_qclear));    // caps the conservative tests only
                        seeds &= 0x0FFFFFFFFFFFFFFFull;
                    }
#endif
                    leaf_seeds = (uint32_t)seeds & 15u;
                    seeds >>= 4;
                    // the other seeds are the roots of this ray's traversal: onto the stack, coarsest first
                    T.sp = 0u;
                    while (seeds != 0ull) {
                        const uint32_t b = 63u - (uint32_t)__clzll((long long)seeds);
                        seeds &= ~(1ull << b);
                        const uint32_t L = (b >> 2) + 1u, q = b & 3u;
                        st.at(T.sp) = pack_node(L, (((cell0 & 0x1FFFu) >> (L + 1u)) << 1) | (q & 1u), (((cell0 >> 13) >> (L + 1u)) << 1) | (q >> 1));
                        T.sp++;
                    }
                    busy = T.sp > 0u;
                }
            }
            // the level-0 seeds of the new rays are leaves (the cell a ray starts in was solved by k_ascent)
            if (__ballot_sync(0xFFFFFFFFu, leaf_seeds != 0u) != 0u) {
                const uint32_t sib0 = cell0 & ~(1u | (1u << 13));
#pragma unroll
                for (uint32_t q = 0; q < 4u; q++) {
                    const bool mine = ((leaf_seeds >> q) & 1u) != 0u;
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
                    if (m != 0u) enqueue(m, mine, sib0 | (q & 1u) | ((q >> 1) << 13));
                }
                __syncwarp();
            }
        }
        if (__ballot_sync(0xFFFFFFFFu, busy) == 0u) {
            if (q_tail != q_head) serve(q_tail - q_head);
            if (exhausted) break;
            continue;
        }
        while (true) {
            // expand the node on top of every busy lane's stack (the stack holds internal nodes only)
            F3D_SCHED_STAT(0, busy);
            uint32_t okm = 0u, bid = 0u;
            bool leaf_kids = false;
            if (busy) {
                T.sp--;
#if F3D_SUN_HORIZON
                okm = expand_core<true, CURV, ASC, Q, IS_SUN && ASC>(F, T, st.at(T.sp), bid, qsrc, &P.hz);
#else
                okm = expand_core<true, CURV, ASC, Q>(F, T, st.at(T.sp), bid, qsrc);
#endif
                n_nodes++;
                leaf_kids = ((bid >> 26) & 15u) == 0u;
                if (!leaf_kids) {
                    if (okm & 8u) { st.at(T.sp) = bid ^ (1u | (1u << 13)); T.sp++; }
                    if (okm & 4u) { st.at(T.sp) = bid ^ (1u << 13); T.sp++; }
                    if (okm & 2u) { st.at(T.sp) = bid ^ 1u; T.sp++; }
                    if (okm & 1u) { st.at(T.sp) = bid; T.sp++; }
                }
            }
            // children of level-1 nodes are leaves: straight into the queue
            if (__ballot_sync(0xFFFFFFFFu, leaf_kids && okm != 0u) != 0u) {
#pragma unroll
                for (uint32_t j = 0; j < 4u; j++) {
                    const bool mine = leaf_kids && ((okm >> j) & 1u);
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
                    if (m != 0u) enqueue(m, mine, bid ^ (j & 1u) ^ ((j >> 1) << 13));
                }
                __syncwarp();
            }
            while (q_tail - q_head >= 32u) serve(32u);     // the one place full batches are solved
            if (busy && T.sp == 0u) busy = false;          // nothing left to expand: its leaves travel on their own
            const uint32_t live = __ballot_sync(0xFFFFFFFFu, busy);
            if (live == 0u) break;
            if (!exhausted && __popc(live) < kRefillBelow) break;
        }
    }
    warp_add_counters(P.counters, 0u, 0u, 0u, n_nodes);
}

// k_ascent: stage 1 of the secondary rays, one thread per listed ray at full lane occupancy (the lists are compacted):
// the mesh test (hybrid scenes), the patch solve of the cell the ray starts in - every ray needs exactly that one - and the
// bottom-up start (ascent_seeds).  A ray whose own cell occludes it, or that has no seed left, is DECIDED here; the others
// are appended (warp ballot + prefix sum, one atomic per warp) to the stage-2 lists k_trace traverses.
// Leaf work items of one k_ascent warp (the near-field walk): same records as k_trace's ring, 64 entries (<= 31 waiting +
// 32 from one walk step).  All members are called by the 32 lanes of a converged warp.
// F3D_ASCENT_PREFETCH = 1: k_ascent requests the next iteration's ray record into L2 one iteration ahead (A/B: profiles/).
#ifndef F3D_ASCENT_PREFETCH
#define F3D_ASCENT_PREFETCH 0
#endif
constexpr uint32_t kWalkMaxCols = kSunNearMaxCols > kEscNearCols ? kSunNearMaxCols : kEscNearCols;
constexpr uint32_t kLeafQA = 64u;
constexpr size_t kAscentRingBytes = (256 / 32) * 3 * kLeafQA * sizeof(uint32_t);     // k_ascent runs 256-thread CTAs
struct AscentRing {
    uint32_t* wq;               // [3][kLeafQA] in shared memory: cell id, pixel word, tmax
    uint32_t head, tail;        // absolute item counters, warp-uniform
    __device__ __forceinline__ uint32_t size() const { return tail - head; }
    __device__ __forceinline__ void enqueue(uint32_t m, bool mine, uint32_t lt, uint32_t id, uint32_t pixw, float tmax) {
        if (mine) {
            const uint32_t slot = (tail + (uint32_t)__popc(m & lt)) & (kLeafQA - 1u);
            wq[slot] = id & 0x03FFFFFFu;
            wq[kLeafQA + slot] = pixw;
            wq[2u * kLeafQA + slot] = __float_as_uint(tmax);
        }
        tail += (uint32_t)__popc(m);
        __syncwarp();
    }
    template <bool IS_SUN, bool CURV>
    __device__ __forceinline__ void serve(const FrameParams& P, const BatchSlot& B, uint8_t* __restrict__ occl, uint32_t count, uint32_t lane, uint32_t& n_nodes) {
        if (lane < count) {
            const uint32_t slot = (head + lane) & (kLeafQA - 1u);
            const uint32_t hpix = wq[kLeafQA + slot];
            n_nodes++;
            if (solve_leaf_item<IS_SUN, CURV>(P, B.rec, wq[slot], hpix, __uint_as_float(wq[2u * kLeafQA + slot]))) occl[hpix & 0x7FFFFFFFu] = 1u;
        }
        head += count;
        __syncwarp();
    }
};

// PASS 0 = classify: origin-cell solve, then the rays a near-field walk can decide (sun horizon / escape map) are walked
//          right here and every other undecided ray goes to the FAR list (compacted: warp ballot + one atomic per warp);
// PASS 1 = the far list: bottom-up start (ascent_seeds) at full lane occupancy -> the stage-2 lists k_trace traverses;
// PASS 2 = single pass (no walk structure exists for this list): origin-cell solve + bottom-up start, as before the walks.
// Two passes because the walk and the 11-level seed loop are both long: run in one warp for a mixed set of rays, each
// executes at partial occupancy (measured: IBL list 16 lanes, +30 % warp-instructions; profiles/r02_walks.md).
template <bool IS_SUN, bool CURV, bool ASC, int PASS, class Q = QGlobal>
__device__ __forceinline__ void ascent_list(const FrameParams& P, const BatchSlot& B, uint32_t* wq_warp, const Q qsrc = Q()) {
    const FastScene& F = P.fast;
    const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const uint32_t n = B.q_counts[PASS == 1 ? (IS_SUN ? 8 : 9) : (IS_SUN ? 0 : 1)];
    const uint32_t* __restrict__ queue = PASS == 1 ? (IS_SUN ? B.qf_sun : B.qf_ibl) : (IS_SUN ? B.q_sun : B.q_ibl);
    uint32_t* __restrict__ queue_far = IS_SUN ? B.qf_sun : B.qf_ibl;
    uint32_t* __restrict__ queue2 = IS_SUN ? B.q2_sun : B.q2_ibl;
    unsigned long long* __restrict__ qseeds = IS_SUN ? B.qn_sun : B.qn_ibl;
    uint8_t* __restrict__ occl = IS_SUN ? B.occl_sun : B.occl_ibl;
    const v3 wi = normalize3(ld3(P.light_dir));
    const v3 wi_reuse = normalize3(wi);
    const bool has_mesh = P.scene.traversal_mode == 0u;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t n_rays = 0u, n_nodes = 0u;
    AscentRing ring;
    ring.wq = wq_warp; ring.head = 0u; ring.tail = 0u;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += stride) {       // warp-uniform trip count
        const uint32_t i = base + lane;
        bool want = false, far = false, all_siblings = false;
        uint32_t pix = 0u, cell0 = 0u;
        unsigned long long seeds = 0ull, kbits = 0ull;
        uint32_t near_cols = 0u, near_pixw = 0u;      // near-field walk: columns to visit, 0 = not a walk ray
        TraceState T{};
#if F3D_ASCENT_PREFETCH
        // the NEXT iteration's record (64 B per pixel, written by k_shade of up to 4 frames ago: DRAM, not L2) is requested now;
        // its list entry is loaded here and consumed by the prefetch below, after the patch solve has covered the load's latency
        uint32_t pix_next = 0xFFFFFFFFu;
        if (i + stride < n) pix_next = __ldg(queue + i + stride);
#endif
        if (i < n) {
            pix = __ldg(queue + i);
            const float4 r0 = __ldcg(B.rec + 4 * (size_t)pix);
            Ray r;
            r.o = V3(r0.x, r0.y, r0.z); r.tmin = 1e-3f; r.tmax = 1e30f;
            if (IS_SUN) r.d = (__float_as_uint(r0.w) & kRecSunReuseDir) ? wi_reuse : wi;
            else { const float4 r1 = __ldcg(B.rec + 4 * (size_t)pix + 1); r.d = V3(r1.x, r1.y, r1.z); }
            if (PASS != 1) n_rays++;
            bool mesh_occl = false, decided = false;
            if (has_mesh) {                      // intersect_hybrid_optimized :213-221 (PASS 1 only needs the clipped tmax again)
                const Hit mh = intersect_mesh(P.scene, r);
                if (mh.hit && mh.t < 0.01f) { if (PASS != 1) occl[pix] = 1u; decided = true; }
                else if (mh.hit && mh.t < r.tmax) { r.tmax = mh.t; mesh_occl = true; }
            }
            if (!decided) {
                ray_setup<CURV>(F, r, T);
                cell0 = origin_cell(F, r.o);
                bool open = true;                // the origin cell does not occlude the ray (PASS 1: known from PASS 0)
                if (PASS != 1) {
                    n_nodes++;
                    open = !solve_origin_cell<CURV>(F, T, cell0);
                    occl[pix] = (!open || mesh_occl) ? 1u : 0u;        // undecided rays: the default a later leaf hit overwrites
                }
                if (open) {
                    uint32_t kidx = kHzNone;
#if F3D_SUN_HORIZON
                    if (IS_SUN && ASC) {        // ascending sun rays: how far does the near field reach? (see SunHorizon)
                        kidx = horizon_lookup(P.hz, F, r.o, cell0);
                        const uint32_t c0maj = P.hz.xmajor ? (cell0 & 0x1FFFu) : (cell0 >> 13);
                        const int32_t q0 = P.hz.forward ? (int32_t)c0maj : (int32_t)P.hz.ncols - 1 - (int32_t)c0maj;
                        T.hz_qclear = kidx == kHzNone ? kHzNoClear : q0 + (int32_t)hz_k(kidx);
                        if (kidx != kHzNone) T.best_t = fminf(T.best_t, horizon_t_clear(P.hz, T, T.hz_qclear));    // after the solve of cell0
                    }
#endif
                    if (PASS == 0) {             // a walk ray is decided below and never reaches k_trace
#if F3D_SUN_HORIZON && F3D_SUN_NEAR
                        if (IS_SUN && ASC && kidx <= kSunNearMaxIdx && !mesh_occl && starts_inside(T, cell0)) {
                            near_cols = hz_k(kidx);
                            near_pixw = pix | ((__float_as_uint(r0.w) & kRecSunReuseDir) ? 0x80000000u : 0u);
                        }
#endif
#if F3D_ESCAPE
                        if (!IS_SUN && !mesh_occl && starts_inside(T, cell0) && escape_cleared(P.esc, F, T, cell0)) {
                            near_cols = kEscNearCols;
                            near_pixw = pix;
                        }
#endif
                        far = near_cols == 0u;
                    } else {
#if F3D_SUN_HORIZON
                        if (IS_SUN && ASC) {
                            seeds = ascent_seeds<CURV, ASC, Q, true>(F, T, cell0, qsrc, &P.hz, &all_siblings);
                            kbits = (unsigned long long)kidx << 60;           // travels to k_trace with the seeds
                        } else
#endif
                            seeds = ascent_seeds<CURV, ASC, Q>(F, T, cell0, qsrc, nullptr, &all_siblings);
                        want = true;
                    }
                }
            }
        }
#if F3D_ASCENT_PREFETCH
        if (pix_next != 0xFFFFFFFFu) prefetch_l2(B.rec + 4 * (size_t)pix_next);
#endif
        if (PASS == 0) {                         // far rays: compacted into the far list
            const uint32_t mf = __ballot_sync(0xFFFFFFFFu, far);
            if (mf != 0u) {
                uint32_t bf = 0u;
                if (lane == 0u) bf = atomicAdd(B.q_counts + (IS_SUN ? 8 : 9), (uint32_t)__popc(mf));
                bf = __shfl_sync(0xFFFFFFFFu, bf, 0);
                if (far) queue_far[bf + (uint32_t)__popc(mf & lt)] = pix;
            }
        } else {
            // rays that start within the pad of a cell border (or outside the DEM): every existing sibling is a seed (warp-cooperative)
            for (uint32_t fb = __ballot_sync(0xFFFFFFFFu, all_siblings); fb != 0u; fb &= fb - 1u) {
                const int src = __ffs((int)fb) - 1;
                const unsigned long long sd = all_sibling_seeds_warp(F, __shfl_sync(0xFFFFFFFFu, cell0, src));
                if ((int)lane == src) seeds = sd;
            }
            want = want && seeds != 0ull;
#ifdef F3D_SCHED_STATS
            if (i < n) {
                atomicAdd(&g_sched_stats[IS_SUN ? 6 : 7], (unsigned long long)__popcll(seeds) + (1ull << 32));   // low: seeds, high: rays
                if (IS_SUN) { atomicAdd(&g_sched_stats[8 + (kbits >> 60)], 1ull); atomicAdd(&g_sched_stats[24 + (kbits >> 60)], (unsigned long long)__popcll(seeds)); }
            }
#endif
            seeds |= kbits;
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, want);
            if (m != 0u) {
                uint32_t b2 = 0u;
                if (lane == 0u) b2 = atomicAdd(B.q_counts + (IS_SUN ? 4 : 5), (uint32_t)__popc(m));
                b2 = __shfl_sync(0xFFFFFFFFu, b2, 0);
                if (want) {
                    const uint32_t slot = b2 + (uint32_t)__popc(m & lt);
                    queue2[slot] = pix;
                    qseeds[slot] = seeds;
                }
            }
        }
#if (F3D_SUN_HORIZON && F3D_SUN_NEAR) || F3D_ESCAPE
        // near-field walk, warp-convergent: column by column every walk ray of the warp tests the (at most three) cells its line
        // touches there; survivors of the conservative test are queued and solved 32 at a time
        if (PASS == 0 && __ballot_sync(0xFFFFFFFFu, near_cols != 0u) != 0u) {
            constexpr bool WALK_ASC = IS_SUN ? ASC : true;        // escape-map rays ascend (escape_cleared)
            WalkDir wd{};
            NearWalk N{};
            if (near_cols != 0u) {
                wd = IS_SUN ? walk_dir(P.hz) : walk_dir(F, T.d);
                N = near_walk_setup(wd, F, T.o, cell0);
            }
            const int32_t nrows = (int32_t)(wd.xmajor ? F.cell_h : F.cell_w);
            const float2* __restrict__ lv0 = F.q.lv[0];            // the global arena (a staged copy needs no prefetch)
            // every [min,max] the walk will read is requested up front (one prefetch per column end: rows r_lo and r_hi sit in
            // the same or in adjacent 32-byte quads), so the column steps below do not each wait for their own L2 round trip
            for (uint32_t c = 0; c < kWalkMaxCols; c++) {
                int32_t cmaj = 0, r_lo = 0, r_hi = -1;
                if (c < near_cols && near_walk_column(wd, N, c, cmaj, r_lo, r_hi)) {
                    const uint32_t ra = (uint32_t)min(max(r_lo, 0), nrows - 1), rb = (uint32_t)min(max(r_hi, 0), nrows - 1);
                    const uint32_t xa = wd.xmajor ? (uint32_t)cmaj : ra, za = wd.xmajor ? ra : (uint32_t)cmaj;
                    const uint32_t xb = wd.xmajor ? (uint32_t)cmaj : rb, zb = wd.xmajor ? rb : (uint32_t)cmaj;
                    prefetch_l1(lv0 + ((za >> 1) * F.q.parent_pitch[0] + (xa >> 1)) * 4u);
                    prefetch_l1(lv0 + ((zb >> 1) * F.q.parent_pitch[0] + (xb >> 1)) * 4u);
                }
            }
            for (uint32_t c = 0; c < kWalkMaxCols; c++) {
                if (__ballot_sync(0xFFFFFFFFu, c < near_cols) == 0u) break;
                int32_t cmaj = 0, r_lo = 0, r_hi = -1;
                const bool col_on = c < near_cols && near_walk_column(wd, N, c, cmaj, r_lo, r_hi);
                uint32_t pass = 0u;
                if (col_on) {
#pragma unroll
                    for (int32_t rr = 0; rr < 3; rr++) {
                        const int32_t r = r_lo + rr;
                        if (r <= r_hi && r >= 0 && r < nrows) {
                            const uint32_t cx = wd.xmajor ? (uint32_t)cmaj : (uint32_t)r, cz = wd.xmajor ? (uint32_t)r : (uint32_t)cmaj;
                            if (((cz << 13) | cx) != cell0 && cell_may_pass<CURV, WALK_ASC, Q>(F, T, cx, cz, qsrc)) pass |= 1u << rr;
                        }
                    }
                }
                if (__ballot_sync(0xFFFFFFFFu, pass != 0u) == 0u) continue;
#pragma unroll
                for (int32_t rr = 0; rr < 3; rr++) {
                    const bool mine = ((pass >> rr) & 1u) != 0u;
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
                    if (m != 0u) {
                        const uint32_t r = (uint32_t)(r_lo + rr);
                        const uint32_t id = wd.xmajor ? ((r << 13) | (uint32_t)cmaj) : (((uint32_t)cmaj << 13) | r);
                        ring.enqueue(m, mine, lt, id, near_pixw, T.tmax);
                        if (ring.size() >= 32u) ring.template serve<IS_SUN, CURV>(P, B, occl, 32u, lane, n_nodes);
                    }
                }
            }
        }
#endif
    }
    if (ring.size() != 0u) ring.template serve<IS_SUN, CURV>(P, B, occl, ring.size(), lane, n_nodes);
    warp_add_counters(P.counters, 0u, IS_SUN ? n_rays : 0u, IS_SUN ? 0u : n_rays, n_nodes);
}

// SUN_MODE as for k_trace: 2 = the sun list is left to the top-down tracer.
#ifndef F3D_ASCENT_MIN_CTAS
#define F3D_ASCENT_MIN_CTAS 3       // 85 registers: the near-field walk keeps the ray's culling constants live
#endif
// One kernel per LIST (IS_SUN): the exact patch solve, the walk and the bottom-up start of one list are ~3 k SASS
// instructions; both lists in one kernel (6.2 k, 99 KB) outgrew the instruction cache - measured: 4.1 of 12 stall cycles per
// issued instruction were "no instruction", k_ascent 0.75 -> 1.16 ms per batch (profiles/r02_ascent_icache.md).
template <bool IS_SUN, bool CURV_SUN, int SUN_MODE, int PASS>
__global__ void __launch_bounds__(256, F3D_ASCENT_MIN_CTAS) k_ascent(const __grid_constant__ FrameParams P) {
    // dynamic shared memory: one AscentRing per warp (kAscentRingBytes), then the staged top levels (F3D_TMA_STAGE)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* wq_warp = reinterpret_cast<uint32_t*>(smem_raw) + (threadIdx.x >> 5) * (3u * kLeafQA);
#if F3D_TMA_STAGE
    const QStaged qs = stage_top_levels(P.fast, P.stage[1], smem_raw + kAscentRingBytes);
    for (uint32_t b = 0; b < P.n_batch; b++) {
        if (IS_SUN) ascent_list<true, CURV_SUN, SUN_MODE == 1, PASS, QStaged>(P, P.slot[b], wq_warp, qs);
        else ascent_list<false, false, false, PASS, QStaged>(P, P.slot[b], wq_warp, qs);
    }
#else
    for (uint32_t b = 0; b < P.n_batch; b++) {
        if (IS_SUN) ascent_list<true, CURV_SUN, SUN_MODE == 1, PASS>(P, P.slot[b], wq_warp);
        else ascent_list<false, false, false, PASS>(P, P.slot[b], wq_warp);
    }
#endif
}

// One persistent launch walks the sun list, then the IBL list: a warp that runs out of sun rays moves
// straight on to IBL rays, so there is no kernel-boundary tail between the two.
// SUN_MODE: 0 = bottom-up, generic height tests; 1 = bottom-up, ascending sun rays (sun above the horizon);
//           2 = round-1 top-down traversal with the exact expansion for the sun list (curved AND descending sun rays,
//               see F3D_CULL_FAST; also what F3D_TRACE_BOTTOM_UP=0 builds use for both lists).
template <bool CURV_SUN, int SUN_MODE>
__global__ void __launch_bounds__(kTraceCtaThreads, F3D_TRACE_MIN_CTAS) k_trace(const __grid_constant__ FrameParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemStack st;
    st.base = reinterpret_cast<uint32_t*>(smem_raw) + threadIdx.x;
    st.stride = kTraceCtaThreads;
    // per-warp leaf ring behind the stacks
    uint32_t* wq = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)P.stack_depth * kTraceCtaThreads + (threadIdx.x >> 5) * (kLeafQBU * kLeafQFields);
    // every sun list of the batch, then every IBL list: a warp that runs out of rays in one list moves straight on to the next
#if F3D_TRACE_BOTTOM_UP && F3D_TMA_STAGE
    const QStaged qs = stage_top_levels(P.fast, P.stage[2], smem_raw + ((trace_smem_bytes_for(P.stack_depth, kTraceCtaThreads) + 15) & ~(size_t)15));
    for (uint32_t b = 0; b < P.n_batch; b++) {
        if (SUN_MODE == 2) trace_list<true, CURV_SUN, true>(P, P.slot[b], st, wq);
        else trace_list_bu<true, CURV_SUN, SUN_MODE == 1, QStaged>(P, P.slot[b], st, wq, qs);
    }
    for (uint32_t b = 0; b < P.n_batch; b++) trace_list_bu<false, false, false, QStaged>(P, P.slot[b], st, wq, qs);
#elif F3D_TRACE_BOTTOM_UP
    for (uint32_t b = 0; b < P.n_batch; b++) {
        if (SUN_MODE == 2) trace_list<true, CURV_SUN, true>(P, P.slot[b], st, wq);
        else trace_list_bu<true, CURV_SUN, SUN_MODE == 1>(P, P.slot[b], st, wq);
    }
    for (uint32_t b = 0; b < P.n_batch; b++) trace_list_bu<false, false, false>(P, P.slot[b], st, wq);
#else
    for (uint32_t b = 0; b < P.n_batch; b++) trace_list<true, CURV_SUN, SUN_MODE == 2>(P, P.slot[b], st, wq);
    for (uint32_t b = 0; b < P.n_batch; b++) trace_list<false, false, false>(P, P.slot[b], st, wq);
#endif
}

// ---------------------------------------------------------------------------------------------
// k_accum: per pixel, combine one camera sample (:523-549) and, after the last sample, accumulate
// and update the windowed Welford statistics (:557-574).  Also re-arms the ray-list counters.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_accum(const __grid_constant__ FrameParams P) {
    uint32_t gx, gy;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < kQCounts * P.n_batch) P.slot[threadIdx.x / kQCounts].q_counts[threadIdx.x % kQCounts] = 0u;
    if (!owned_pixel(P, gx, gy)) return;
    const uint32_t pix = gy * P.W + gx;
    const uint32_t spp = max(P.spp, 1u), window = max(P.window, 2u);
    uint32_t s = P.sample_index, frame = P.frame_index;
    v3 fr = V3(0, 0, 0);
    if (s > 0u) { const float4 c = P.sstate[3 * (size_t)pix + 2]; fr = V3(c.x, c.y, c.z); }
    float4 acc = make_float4(0, 0, 0, 0);
    float2 wf = make_float2(0, 0);
    bool loaded = false;
    for (uint32_t k = 0; k < P.n_batch; k++) {          // the steps of the batch in order: the per-pixel sums are order-sensitive
        const BatchSlot& B = P.slot[k];
        const float4* rec = B.rec + 4 * (size_t)pix;
        const float4 r0 = ld_stream(rec), r2 = ld_stream(rec + 2), r3 = ld_stream(rec + 3);
        const uint32_t flags = __float_as_uint(r0.w);
        if (!(flags & kRecHit)) {
            fr = fr + V3(r2.w, r3.x, r3.y);                       // sky radiance (:489)
        } else {
            const float4 r1 = ld_stream(rec + 1);
            v3 sun = V3(0, 0, 0);
            if (flags & kRecSun) {
                const float vis = B.occl_sun[pix] ? 0.0f : 1.0f;
                sun = (V3(r2.x, r2.y, r2.z) * vis) * r1.w;           // albedo*light_color*nd*vis*reuse_w (:531)
            }
            const float env_vis = B.occl_ibl[pix] ? 0.0f : 1.0f;
            const v3 ibl = V3(r2.w, r3.x, r3.y) * env_vis;           // albedo*env(ei)*env_vis (:545)
            fr = (fr + sun) + ibl;
        }
        if (s + 1u == spp) {                               // last sample of its frame: accumulate + windowed Welford (:557-574)
            if (!loaded) { acc = ld_stream(P.accum + pix); wf = ld_stream(P.welford + pix); loaded = true; }
            const float fspp = (float)spp;
            fr = V3(fdiv(fr.x, fspp), fdiv(fr.y, fspp), fdiv(fr.z, fspp));
            acc.x = acc.x + fr.x;
            acc.y = acc.y + fr.y;
            acc.z = acc.z + fr.z;
            acc.w = acc.w + 1.0f;
            if (frame % window == 0u) wf = make_float2(0.0f, 0.0f);
            const float mean_lum = luminance(V3(fdiv(acc.x, acc.w), fdiv(acc.y, acc.w), fdiv(acc.z, acc.w)));
            const float kk = (float)(frame % window) + 1.0f;
            const float delta = mean_lum - wf.x;
            const float mean = wf.x + fdiv(delta, kk);
            const float m2 = wf.y + delta * (mean_lum - mean);
            wf = make_float2(mean, m2);
            fr = V3(0, 0, 0);
            s = 0u;
            frame++;
        } else s++;
    }
    if (s > 0u) P.sstate[3 * (size_t)pix + 2] = make_float4(fr.x, fr.y, fr.z, 0.0f);      // a frame continues in the next batch
    if (loaded) { st_stream(P.accum + pix, acc); st_stream(P.welford + pix, wf); }
}

// ---------------------------------------------------------------------------------------------
// k_gbuffer: unjittered centre ray per pixel.  Serves both main_terrain_gbuffer (:619-644) and the
// frame-0 AOV block of main_terrain (:583-609), which trace the identical ray.
// ---------------------------------------------------------------------------------------------
struct GbufferOut {
    uint8_t* pixflags;     // bit0 facing, bits1-2 hit type (0 miss, 1 terrain, 2 mesh)
    ushort4* aov_normal;   // RGBA16F normal (alpha unused)
    float* aov_depth;      // R32F, qNaN 0x7fc00000 on miss
};

__global__ void __launch_bounds__(kThreads) k_gbuffer(const __grid_constant__ FrameParams P, GbufferOut G) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemStack st;
    st.base = reinterpret_cast<uint32_t*>(smem_raw) + threadIdx.x;
    st.stride = kThreads;
    uint32_t gx, gy;
    const bool active = owned_pixel(P, gx, gy);
    const uint32_t pix = active ? gy * P.W + gx : 0u;
    uint32_t nodes = 0;
    Ray ray;
    ray.o = V3(0, 0, 0); ray.d = V3(0, 0, 1); ray.tmin = 1e-3f; ray.tmax = 1e30f;
    if (active) ray = camera_ray(P, gx, gy, 0.0f, 0.0f);
    const PrimaryHit hit = primary_hit(P, ray, active, st, nodes);   // warp-cooperative
    if (!active) return;
    // ReSTIR G-buffer record: hit -> (normal, 1); miss -> (0,0,1,1)  (:635-643)
    const v3 nr = hit.hit ? hit.normal : V3(0.0f, 0.0f, 1.0f);
    const v3 N = normalize3(nr);
    const v3 wi_s = normalize3(normalize3(ld3(P.light_dir)));   // normalize(r.sample.direction), sample dir = wi
    const bool facing = fmaxf(dot3(N, wi_s), 0.0f) > 0.0f;
    uint8_t flags = facing ? 1u : 0u;
    if (hit.hit) flags |= (hit.hit_type == 3u ? 1u : 2u) << 1;
    G.pixflags[pix] = flags;
    ushort4 n16;
    n16.x = __half_as_ushort(__float2half_rn(hit.hit ? hit.normal.x : 0.0f));
    n16.y = __half_as_ushort(__float2half_rn(hit.hit ? hit.normal.y : 0.0f));
    n16.z = __half_as_ushort(__float2half_rn(hit.hit ? hit.normal.z : 0.0f));
    n16.w = __half_as_ushort(__float2half_rn(1.0f));
    G.aov_normal[pix] = n16;
    G.aov_depth[pix] = hit.hit ? hit.t : __uint_as_float(0x7fc00000u);
}

// ---------------------------------------------------------------------------------------------
// k_variance: the convergence read-back of render_terrain.rs:1206-1233 reduced on the device:
// out[0] = bits(max over owned pixels of m2/(n-1)) (non-negative floats order as uints),
// out[1] != 0 when any m2 is non-finite.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_variance(const __grid_constant__ FrameParams P, float n_window,
                                                             uint32_t* __restrict__ out) {
    uint32_t gx, gy;
    float v = 0.0f;
    uint32_t bad = 0u;
    if (owned_pixel(P, gx, gy)) {
        const float m2 = P.welford[gy * P.W + gx].y;
        if (!isfinite(m2)) bad = 1u;
        else v = fmaxf(0.0f, fdiv(m2, n_window - 1.0f));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, off));
        bad |= __shfl_xor_sync(0xFFFFFFFFu, bad, off);
    }
    if ((threadIdx.x & 31u) == 0u) {
        if (v > 0.0f) atomicMax(out + 0, __float_as_uint(v));
        if (bad) atomicOr(out + 1, 1u);
    }
}

// ---------------------------------------------------------------------------------------------
// k_resolve: last frame's reuse pass + reservoir validity (render_terrain.rs:1313-1337), beauty
// resolve mean -> Reinhard -> RGBA16F -> u8 (hybrid_kernel.wgsl:109-112, render_terrain.rs:1358-1366)
// and AOV decode (albedo/normal f16 -> f32, depth f32).
// ---------------------------------------------------------------------------------------------
struct ResolveOut {
    uint8_t* rgba;       // W*H*4 or NULL
    float* albedo;       // W*H*3 or NULL
    float* normal;       // W*H*3 or NULL
    float* depth;        // W*H   or NULL
    const ushort4* aov_normal;
    const float* aov_depth;
    uint32_t* validity;  // [0] |= 1 non-finite bookkeeping, [1] |= 1 some valid reservoir
    uint32_t last_frame; // frames - 1
};

__device__ __forceinline__ float f16_round(float v) { return __half2float(__float2half_rn(v)); }

__global__ void __launch_bounds__(kThreads) k_resolve(const __grid_constant__ FrameParams P, ResolveOut R) {
    uint32_t gx, gy;
    uint32_t nonfinite = 0u, valid = 0u;
    if (owned_pixel(P, gx, gy)) {
        const uint32_t pix = gy * P.W + gx;
        const uint8_t flags = P.pixflags[pix];
        const Resv r = spatial_reuse(P, P.resv_in, gx, gy, (flags & 1u) != 0u, R.last_frame);
        if (!(isfinite(r.w_sum) && isfinite(r.weight) && isfinite(r.target_pdf))) nonfinite = 1u;
        if (r.m > 0u && r.weight > 0.0f && r.target_pdf > 0.0f) valid = 1u;
        if (R.rgba) {
            const float4 acc = P.accum[pix];
            const float ch[3] = {acc.x, acc.y, acc.z};
            uchar4 px;
            uint8_t* pc = &px.x;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float mean = fdiv(ch[c], acc.w);
                float exposed = mean * P.exposure;
                float ldr = fdiv(exposed, 1.0f + exposed);
                float v = f16_round(ldr);
                pc[c] = (uint8_t)(clampf(v, 0.0f, 1.0f) * 255.0f + 0.5f);
            }
            px.w = 255;
            reinterpret_cast<uchar4*>(R.rgba)[pix] = px;
        }
        const uint32_t hit_type = (flags >> 1) & 3u;
        if (R.albedo) {
            v3 a = hit_type == 1u ? ld3(P.scene.albedo) : (hit_type == 2u ? V3(0.7f, 0.7f, 0.8f) : V3(0, 0, 0));
            R.albedo[3 * (size_t)pix + 0] = f16_round(a.x);
            R.albedo[3 * (size_t)pix + 1] = f16_round(a.y);
            R.albedo[3 * (size_t)pix + 2] = f16_round(a.z);
        }
        if (R.normal) {
            const ushort4 n16 = R.aov_normal[pix];
            R.normal[3 * (size_t)pix + 0] = __half2float(__ushort_as_half(n16.x));
            R.normal[3 * (size_t)pix + 1] = __half2float(__ushort_as_half(n16.y));
            R.normal[3 * (size_t)pix + 2] = __half2float(__ushort_as_half(n16.z));
        }
        if (R.depth) R.depth[pix] = R.aov_depth[pix];
    }
    nonfinite = __any_sync(0xFFFFFFFFu, nonfinite);
    valid = __any_sync(0xFFFFFFFFu, valid);
    if ((threadIdx.x & 31u) == 0u) {
        if (nonfinite) atomicOr(R.validity + 0, 1u);
        if (valid) atomicOr(R.validity + 1, 1u);
    }
}

// The AOV half of k_resolve alone: albedo / normal / depth come from the one-shot G-buffer pass (frame 0, render_terrain.rs:1091-1121),
// so the one-call path decodes them right after set-up and copies them to the host WHILE the frames render (f3d_backend.cu
// early_aov_readback): 58 of the 66 MB of a 1080p read-back leave the critical path.  Same expressions as k_resolve.
__global__ void __launch_bounds__(kThreads) k_resolve_aovs(const __grid_constant__ FrameParams P, ResolveOut R) {
    uint32_t gx, gy;
    if (!owned_pixel(P, gx, gy)) return;
    const uint32_t pix = gy * P.W + gx;
    const uint32_t hit_type = (P.pixflags[pix] >> 1) & 3u;
    if (R.albedo) {
        v3 a = hit_type == 1u ? ld3(P.scene.albedo) : (hit_type == 2u ? V3(0.7f, 0.7f, 0.8f) : V3(0, 0, 0));
        R.albedo[3 * (size_t)pix + 0] = f16_round(a.x);
        R.albedo[3 * (size_t)pix + 1] = f16_round(a.y);
        R.albedo[3 * (size_t)pix + 2] = f16_round(a.z);
    }
    if (R.normal) {
        const ushort4 n16 = R.aov_normal[pix];
        R.normal[3 * (size_t)pix + 0] = __half2float(__ushort_as_half(n16.x));
        R.normal[3 * (size_t)pix + 1] = __half2float(__ushort_as_half(n16.y));
        R.normal[3 * (size_t)pix + 2] = __half2float(__ushort_as_half(n16.z));
    }
    if (R.depth) R.depth[pix] = R.aov_depth[pix];
}

// ---------------------------------------------------------------------------------------------
// k_aether: AETHER aerial-perspective post (f3d_aether.cuh) over the finished accumulation, fused with the
// beauty read-back: accum + frame-0 depth AOV + hit-type bits -> L_surface*T + L_inscatter -> Reinhard ->
// RGBA16F rounding -> u8.  Replaces prometheus_aerial.wgsl `main` (8x8 workgroups, aether_post.rs:336-339),
// its two texture copies (:296-331) and the out-texture read-back (render_terrain.rs:1358-1366).  One launch
// per render; ALU-bound (two quadrilinear LUT fetches + 16 altitude samples + 11 wavelengths per hit pixel);
// 16 B + 4 B + 1 B read and 4 B written per pixel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_aether(const __grid_constant__ FrameParams P, const AetherParams A,
                                                    const float* __restrict__ aov_depth, uint8_t* __restrict__ rgba) {
    uint32_t gx, gy;
    if (!owned_pixel(P, gx, gy)) return;
    const uint32_t pix = gy * P.W + gx;
    AetherView V;
    V.cam_origin = ld3(P.cam_origin); V.cam_right = ld3(P.cam_right); V.cam_up = ld3(P.cam_up); V.cam_forward = ld3(P.cam_forward);
    V.sun_dir = normalize3(ld3(P.light_dir));
    V.exposure = P.exposure;
    V.W = P.W; V.H = P.H;
    const bool visible = ((P.pixflags[pix] >> 1) & 3u) != 0u;      // frame-0 visibility AOV (:605-607)
    const v3 ldr = aether_pixel(A, V, gx, gy, P.accum[pix], aov_depth[pix], visible);
    uchar4 px;
    px.x = (uint8_t)(clampf(f16_round(ldr.x), 0.0f, 1.0f) * 255.0f + 0.5f);
    px.y = (uint8_t)(clampf(f16_round(ldr.y), 0.0f, 1.0f) * 255.0f + 0.5f);
    px.z = (uint8_t)(clampf(f16_round(ldr.z), 0.0f, 1.0f) * 255.0f + 0.5f);
    px.w = 255;
    reinterpret_cast<uchar4*>(rgba)[pix] = px;
}

// ---------------------------------------------------------------------------------------------
// Pyramid build on the GPU (replaces the CPU loops of build_minmax_mips,
// terrain_heightfield.rs:132-202).  min/max are exact, so the levels equal the CPU build bit for bit.
// ---------------------------------------------------------------------------------------------
__global__ void k_build_level0(const float* __restrict__ heights, uint32_t w, uint32_t h, uint32_t pw, uint32_t ph,
                               float ex, float4* __restrict__ cells, float2* __restrict__ mm0) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= pw || y >= ph) return;
    const uint32_t cw = w - 1u, ch = h - 1u;
    float2 mm = make_float2(__int_as_float(0x7f800000), __int_as_float(0xff800000));
    if (x < cw && y < ch) {
        const size_t i00 = (size_t)y * w + x;
        const float a = heights[i00] * ex, b = heights[i00 + 1] * ex, c = heights[i00 + w] * ex, d = heights[i00 + w + 1] * ex;
        cells[(size_t)y * cw + x] = make_float4(a, b, c, d);
        mm.x = fminf(fminf(fminf(a, b), c), d);
        mm.y = fmaxf(fmaxf(fmaxf(a, b), c), d);
    }
    mm0[(size_t)y * pw + x] = mm;
}

__global__ void k_reduce_level(const float2* __restrict__ prev, uint32_t lw, uint32_t lh, float2* __restrict__ next,
                               uint32_t nw, uint32_t nh) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nw || y >= nh) return;
    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
#pragma unroll
    for (uint32_t dy = 0; dy < 2u; dy++)
#pragma unroll
        for (uint32_t dx = 0; dx < 2u; dx++) {
            const uint32_t sx = min(2u * x + dx, lw - 1u), sy = min(2u * y + dy, lh - 1u);
            const float2 v = prev[(size_t)sy * lw + sx];
            mn = fminf(mn, v.x);
            mx = fmaxf(mx, v.y);
        }
    next[(size_t)y * nw + x] = make_float2(mn, mx);
}

// Quad-packs one plain level for the production traversal: slot (x, y) of the level goes to
// quads[((y>>1) * parent_pitch + (x>>1)) * 4 + ((y&1)*2 + (x&1))]; slots outside the level hold the
// (+inf, -inf) sentinel (never visited: their cells lie outside the DEM).
// ---------------------------------------------------------------------------------------------
// Sun horizon strips (see SunHorizon in f3d_trace_fast.cuh), built once per session.
// k_hz_build: one thread per (travel column q, strip j): F = Hs[j][q] - q g in double, rounded UP to f32.
// k_hz_suffix: one thread per strip: in-place suffix maximum over the columns (coalesced across strips).
// ---------------------------------------------------------------------------------------------
__global__ void k_hz_build(const float4* __restrict__ cells, uint32_t cell_w, uint32_t cell_h, const SunHorizon Z, double m, double g,
                           float* __restrict__ F) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x, q = blockIdx.y;
    if (j >= Z.nstrips) return;
    const uint32_t nrows = Z.xmajor ? cell_h : cell_w;
    const uint32_t i = Z.forward ? q : Z.ncols - 1u - q;
    const double jw = (double)((int32_t)j + Z.j0), d = 1.0 / 64.0;
    const double lo = jw - d + fmin(m * (double)i, m * ((double)i + 1.0)), hi = jw + 1.0 + d + fmax(m * (double)i, m * ((double)i + 1.0));
    long long r0 = (long long)floor(lo), r1 = (long long)floor(hi);
    if (r0 < 0) r0 = 0;
    if (r1 > (long long)nrows - 1) r1 = (long long)nrows - 1;
    float hs = -3.0e38f;
    for (long long r = r0; r <= r1; r++) {
        const uint32_t cx = Z.xmajor ? i : (uint32_t)r, cz = Z.xmajor ? (uint32_t)r : i;
        const float4 h = cells[(size_t)cz * cell_w + cx];
        hs = fmaxf(hs, fmaxf(fmaxf(h.x, h.y), fmaxf(h.z, h.w)));
    }
    F[(size_t)q * Z.nstrips + j] = hs > -1.0e38f ? __double2float_ru((double)hs - (double)q * g) : -3.0e38f;
}

// One WARP per strip (a thread per strip walked 2047 columns serially: 0.93 ms of every session's 1.07 ms set-up): each lane
// takes a contiguous chunk of columns, reduces it, the 32 chunk maxima are suffix-scanned with shuffles, then every lane
// rewrites its chunk seeded with the maximum of all later chunks.  max is exact in any order: same bytes as the serial scan.
__global__ void __launch_bounds__(256) k_hz_suffix(float* __restrict__ S, uint32_t nstrips, uint32_t ncols) {
    const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (j >= nstrips) return;                                  // warp-uniform
    const uint32_t chunk = (ncols + 31u) / 32u;
    const uint32_t q0 = min(lane * chunk, ncols), q1 = min(q0 + chunk, ncols);
    float mine = -3.0e38f;
    for (uint32_t q = q0; q < q1; q++) mine = fmaxf(mine, S[(size_t)q * nstrips + j]);
    float later = -3.0e38f;                                    // maximum over the chunks of the lanes above this one
    float run = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {                         // inclusive suffix maximum over lanes
        const float o = __shfl_down_sync(0xFFFFFFFFu, run, d);
        if (lane + (uint32_t)d < 32u) run = fmaxf(run, o);
    }
    later = __shfl_down_sync(0xFFFFFFFFu, run, 1);
    if (lane == 31u) later = -3.0e38f;
    float acc = later;
    for (uint32_t q = q1; q > q0; q--) {
        const size_t k = (size_t)(q - 1u) * nstrips + j;
        acc = fmaxf(acc, S[k]);
        S[k] = acc;
    }
}

// Escape map (see EscapeMap in f3d_trace_fast.cuh), built once per session: one thread per DEM cell, 32 consecutive cells of
// a row per warp (they share the coarse levels' ring nodes: broadcast loads).  Per level the ring of (2 kEscOuter + 1)^2 -
// (2 kEscR + 1)^2 nodes around the cell's ancestor; per node one load of its maximum, the smallest distance from the
// (widened) cell to the node's rectangle, one rsqrt, and a running maximum per octant the node can be reached in.
__global__ void __launch_bounds__(128) k_escape_build(const FastScene F, const EscapeOctTable O, const float pad_abs, float* __restrict__ E) {
    const uint32_t cx = blockIdx.x * blockDim.x + threadIdx.x, cz = blockIdx.y;
    if (cx >= F.cell_w || cz >= F.cell_h) return;
    const float4 h = F.cells[(size_t)cz * F.cell_w + cx];
    const float href = fminf(fminf(h.x, h.y), fminf(h.z, h.w)) - pad_abs;
    float e[8];
#pragma unroll
    for (int o = 0; o < 8; o++) e[o] = 0.0f;         // slopes <= 0 are never needed: only ascending rays ask
    for (uint32_t k = 0; k + 1u < F.mip_count; k++) {
        const int32_t lw = (int32_t)((F.cell_w + (1u << k) - 1u) >> k), lh = (int32_t)((F.cell_h + (1u << k) - 1u) >> k);
        if (lw <= kEscR + 1 && lh <= kEscR + 1) break;          // the inner block is the whole level: everything is covered
        const int32_t px = (int32_t)(cx >> k), pz = (int32_t)(cz >> k);
        const float2* __restrict__ lv = F.q.lv[k];
        const uint32_t pitch = F.q.parent_pitch[k];
        for (int32_t oz = -kEscOuter; oz <= kEscOuter; oz++) {
            const int32_t jz = pz + oz;
            if (jz < 0 || jz >= lh) continue;
            const int32_t z0 = jz << k, z1 = min((jz + 1) << k, (int32_t)F.cell_h);
            const float gz = fmaxf((float)max(max(z0 - (int32_t)(cz + 1u), (int32_t)cz - z1), 0) - 0.015625f, 0.0f) * F.sz;
            const bool inner_z = oz >= -kEscR && oz <= kEscR;
            for (int32_t ox = -kEscOuter; ox <= kEscOuter; ox++) {
                if (inner_z && ox >= -kEscR && ox <= kEscR) continue;
                const int32_t jx = px + ox;
                if (jx < 0 || jx >= lw) continue;
                const float mx = __ldg(&lv[((size_t)(jz >> 1) * pitch + (size_t)(jx >> 1)) * 4u + ((jz & 1) * 2 + (jx & 1))].y);
                const int32_t x0 = jx << k, x1 = min((jx + 1) << k, (int32_t)F.cell_w);
                const float gx = fmaxf((float)max(max(x0 - (int32_t)(cx + 1u), (int32_t)cx - x1), 0) - 0.015625f, 0.0f) * F.sx;
                const float num = mx - href;                     // pad_abs on both ends
                if (!(num > 0.0f)) continue;
                float sl = num * rsqrtf(gx * gx + gz * gz);      // D > 0: a ring node is >= kEscR whole nodes away on one axis
                sl = sl * 1.00001f;                              // rsqrt.approx (2^-22) and the products above, rounded up
                const uint32_t mask = O.m[oz + kEscOuter][ox + kEscOuter];
#pragma unroll
                for (int o = 0; o < 8; o++)
                    if ((mask >> o) & 1u) e[o] = fmaxf(e[o], sl);
            }
        }
    }
    float4* out = reinterpret_cast<float4*>(E + ((size_t)cz * F.cell_w + cx) * 8u);
    out[0] = make_float4(e[0], e[1], e[2], e[3]);
    out[1] = make_float4(e[4], e[5], e[6], e[7]);
}

__global__ void k_pack_quads(const float2* __restrict__ plain, uint32_t lw, uint32_t lh, float2* __restrict__ quads,
                             uint32_t parent_pitch, uint32_t parent_h) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= 2u * parent_pitch || y >= 2u * parent_h) return;
    float2 v = make_float2(__int_as_float(0x7f800000), __int_as_float(0xff800000));
    if (x < lw && y < lh) v = plain[(size_t)y * lw + x];
    quads[((size_t)(y >> 1) * parent_pitch + (x >> 1)) * 4u + ((y & 1u) * 2u + (x & 1u))] = v;
}

// Non-finite scan of the uploaded heightfield (trust boundary of build_minmax_mips, :144-148).
__global__ void k_check_finite(const float* __restrict__ v, size_t n, uint32_t* __restrict__ flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint32_t bad = 0u;
    for (; i < n; i += stride)
        if (!isfinite(v[i])) bad = 1u;
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31u) == 0u) atomicOr(flag, 1u);
}

// KAT seam: terrain_trace over a ray batch (terrain_heightfield.rs:1646-1671 test entry).
// variant 0 = production traversal (trace_fast), 1 = literal restatement of the WGSL loop.
// The production traversal serves closest-hit only without curvature (the only combination the
// renderer uses, hybrid_traversal.wgsl:248-259 / hybrid_terrain_traversal.wgsl:374-376); a
// closest-hit + curvature request is routed to the literal loop.
constexpr int kTraceThreads = 128;

__global__ void __launch_bounds__(kTraceThreads) k_trace_rays(SceneParams S, FastScene F, uint32_t stack_depth, int variant,
                                                              const float4* __restrict__ rays, uint64_t n, int any_hit,
                                                              int apply_curv, uint8_t* __restrict__ hit, float* __restrict__ t,
                                                              float* __restrict__ normal, unsigned long long* __restrict__ nodes_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    Ray r;
    r.o = V3(0, 0, 0); r.tmin = 0.0f; r.d = V3(0, 0, 1); r.tmax = 0.0f;
    if (valid) {
        const float4 a = rays[2 * i], b = rays[2 * i + 1];
        r.o = V3(a.x, a.y, a.z); r.tmin = a.w; r.d = V3(b.x, b.y, b.z); r.tmax = b.w;
    }
    uint32_t nodes = 0;
    const bool curv = apply_curv && S.curvature_enabled;
    bool h_hit;
    float h_t;
    v3 h_n = V3(0, 0, 0);
    if (variant == 0 && (any_hit || !curv)) {
        SmemStack st;
        st.base = reinterpret_cast<uint32_t*>(smem_raw) + threadIdx.x;
        st.stride = kTraceThreads;
        FastHit fh;
        if (any_hit) fh = curv ? trace_fast<true, true>(F, r, valid, st, nodes) : trace_fast<true, false>(F, r, valid, st, nodes);
        else fh = trace_fast<false, false>(F, r, valid, st, nodes);
        h_hit = fh.hit; h_t = fh.t;
        if (fh.hit) { v3 p; finish_hit(F, r, fh, p, h_n); }
    } else {
        Hit h;
        if (any_hit) h = curv ? terrain_trace<true, true>(S, r, nodes) : terrain_trace<true, false>(S, r, nodes);
        else h = curv ? terrain_trace<false, true>(S, r, nodes) : terrain_trace<false, false>(S, r, nodes);
        h_hit = h.hit != 0u; h_t = h.t; h_n = h.normal;
    }
    if (!valid) return;
    hit[i] = h_hit ? 1 : 0;
    t[i] = h_t;
    if (normal) { normal[3 * i] = h_n.x; normal[3 * i + 1] = h_n.y; normal[3 * i + 2] = h_n.z; }
    if (nodes_out) atomicAdd(nodes_out, (unsigned long long)nodes);
}

}  // namespace f3d
