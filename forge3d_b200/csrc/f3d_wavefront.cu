// forge3d_b200/csrc/f3d_wavefront.cu
// Host side of f3d_wavefront_render: the frame loop of render_pt_reference (/root/reference/src/path_tracing/adjudication.rs:76-331)
// over the bounce kernels of f3d_wavefront.cuh.  Per batch of frames: kWfWideDepth compacted bounce launches, one tail launch and the
// in-order merge into the accumulator, all queued
// without a host round trip (the reference reads a queue header back every bounce, wavefront/queues/types.rs:166-213); the per-frame
// path-length histograms stay in a device table that is read once at the end to apply the reference's two frame rules (>= 2 iterations,
// ray-queue capacity 4 * W * H).  No CPU fallback.
#include "f3d_host.h"
#include "f3d_wavefront.cuh"

using namespace f3d;
#define g_err g_f3d_err

namespace {

uint32_t splitmix32(uint32_t x) {   // adjudication.rs:226-232
    x += 0x9E3779B9u;
    uint32_t z = x;
    z = (z ^ (z >> 16)) * 0x21F0AAADu;
    z = (z ^ (z >> 15)) * 0x735A2D97u;
    return z ^ (z >> 15);
}
void sobol2(uint32_t i, float* ox, float* oy) {   // pt_raygen.wgsl:107-153 (integer work; the scale by 2^-32 is exact)
    uint32_t xb = 0, yb = 0, idx = i;
    for (uint32_t j = 0; j < 32u; j++) {
        if (idx & 1u) {
            const uint32_t base = 0x80000000u >> j;
            xb ^= base;
            yb ^= base ^ ((base >> 1) ^ (base >> 3));
        }
        idx >>= 1;
    }
    *ox = (float)xb * (1.0f / 4294967296.0f);
    *oy = (float)yb * (1.0f / 4294967296.0f);
}

struct WfRun {
    int device = 0;
    std::vector<void*> bufs;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ~WfRun() {
        cudaDeviceSynchronize();
        for (void* p : bufs) cached_free(p, device);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
    template <class T> int alloc(T** p, size_t count) {
        void* q = nullptr;
        CUDA_TRY(cached_malloc(&q, (count ? count : 1) * sizeof(T), device));
        bufs.push_back(q);
        *p = (T*)q;
        return 0;
    }
    template <class T> int upload(const T** p, const T* src, size_t count) {
        T* q = nullptr;
        int rc = alloc(&q, count);
        if (rc) return rc;
        if (count) CUDA_TRY(cudaMemcpy(q, src, count * sizeof(T), cudaMemcpyHostToDevice));
        *p = q;
        return 0;
    }
};

}  // namespace

extern "C" int f3d_wavefront_render(const f3d_wavefront_scene* sc, uint32_t width, uint32_t height, uint32_t spp_frames, int32_t device,
                                    float* hdr_rgba, uint8_t* rgba8, f3d_wavefront_stats* stats) {
    return f3d_wavefront_render_part(sc, width, height, spp_frames, device, nullptr, hdr_rgba, rgba8, stats);
}

extern "C" int f3d_wavefront_render_part(const f3d_wavefront_scene* sc, uint32_t width, uint32_t height, uint32_t spp_frames, int32_t device,
                                         const f3d_wavefront_part* part, float* hdr_rgba, uint8_t* rgba8, f3d_wavefront_stats* stats) {
    g_err[0] = 0;
    if (!sc) return fail(F3D_ERR_ARGUMENT, "null scene");
    if (part && (part->world == 0 || part->rank >= part->world || (part->world > 1 && part->block_rows == 0)))
        return fail(F3D_ERR_ARGUMENT, "invalid partition: rank %u of %u, block_rows %u", part->rank, part->world, part->block_rows);
    if (width == 0 || height == 0 || spp_frames == 0)
        return fail(F3D_ERR_RENDER, "adjudication PT reference requires non-zero width/height/spp");   // adjudication.rs:85-89
    if ((uint64_t)width * height > 0x3FFFFFFFull) return fail(F3D_ERR_ARGUMENT, "image too large for the 32-bit ray queues");
    if ((sc->nspheres && !sc->spheres) || (sc->ndir && !sc->dir_lights) || (sc->narea && !sc->area_lights) ||
        (sc->nimportance && !sc->importance) || (sc->mesh_ntris && (!sc->mesh_xyz || !sc->mesh_idx)) || (sc->ninstances && !sc->instances))
        return fail(F3D_ERR_ARGUMENT, "null scene buffer with a non-zero count");
    // the sphere array doubles as the material table (pt_shade.wgsl:487-493 reads slot 0 for every out-of-range id)
    if (sc->nspheres == 0) return fail(F3D_ERR_ARGUMENT, "the scene needs at least one sphere / material slot");
    for (uint32_t i = 0; i < sc->nspheres; i++) {
        const float* s = sc->spheres + 20 * (size_t)i;
        if (fabsf(fmaxf(0.002f, s[15]) - fmaxf(0.002f, s[16])) >= 1e-4f)
            return fail(F3D_ERR_ARGUMENT, "anisotropic GGX (ax != ay) is not supported");
    }
    for (size_t k = 0; k < 3 * (size_t)sc->mesh_ntris; k++)
        if (sc->mesh_idx[k] >= sc->mesh_nverts) return fail(F3D_ERR_ARGUMENT, "mesh index %u out of range", sc->mesh_idx[k]);
    int rc = select_device(device);
    if (rc) return rc;

    WfRun R;
    R.device = device;
    const uint32_t npx = width * height;
    WfParams P{};
    P.w = width; P.h = height;
    P.part_rank = part ? part->rank : 0u;
    P.part_world = part ? part->world : 1u;
    P.part_rows = P.part_world > 1u ? part->block_rows : height;
    uint32_t owned_rows = 0;   // rows of this rank inside the image; local_rows also counts the ragged last block's rows past the edge
    {
        const uint32_t nblocks = (height + P.part_rows - 1u) / P.part_rows;
        const uint32_t mine = nblocks > P.part_rank ? (nblocks - P.part_rank + P.part_world - 1u) / P.part_world : 0u;
        P.local_rows = mine * P.part_rows;
        for (uint32_t b = P.part_rank; b < nblocks; b += P.part_world) owned_rows += std::min(P.part_rows, height - b * P.part_rows);
    }
    const uint32_t n_primary = owned_rows * width;
    memcpy(P.cam_origin, sc->cam_origin, 12); memcpy(P.cam_forward, sc->cam_forward, 12);
    memcpy(P.cam_right, sc->cam_right, 12); memcpy(P.cam_up, sc->cam_up, 12);
    P.fov_y_rad = sc->fov_y_rad;
    P.aspect = (float)width / (float)height;   // adjudication.rs:203
    memcpy(P.env, sc->environment, sizeof P.env);
    P.nsph = sc->nspheres; P.ndir = sc->ndir; P.narea = sc->narea; P.nimp = sc->nimportance; P.mesh.ntris = sc->mesh_ntris; P.ninst = sc->ninstances;
    if ((rc = R.upload(&P.spheres, sc->spheres, 20 * (size_t)sc->nspheres))) return rc;
    if ((rc = R.upload(&P.dirl, sc->dir_lights, 8 * (size_t)sc->ndir))) return rc;
    if ((rc = R.upload(&P.areal, sc->area_lights, 12 * (size_t)sc->narea))) return rc;
    if ((rc = R.upload(&P.imp, sc->importance, (size_t)sc->nimportance))) return rc;
    if ((rc = R.upload(&P.mesh.xyz, sc->mesh_xyz, 3 * (size_t)sc->mesh_nverts))) return rc;
    if ((rc = R.upload(&P.mesh.idx, sc->mesh_idx, 3 * (size_t)sc->mesh_ntris))) return rc;
    if ((rc = R.upload(&P.inst, sc->instances, 36 * (size_t)sc->ninstances))) return rc;
    if (sc->mesh_ntris > 8u && !getenv("F3D_B200_NO_MESH_BVH")) {
        std::vector<float4> nodes;
        std::vector<uint32_t> tris;
        host_build_mesh_bvh(sc->mesh_xyz, sc->mesh_idx, sc->mesh_ntris, &nodes, &tris);
        if ((rc = R.upload(&P.mesh.bvh_nodes, nodes.data(), nodes.size()))) return rc;
        if ((rc = R.upload(&P.mesh.bvh_tris, tris.data(), tris.size()))) return rc;
    }
    if ((rc = R.alloc(&P.accum, npx))) return rc;
    constexpr uint32_t kSlots = kWfMaxDepth + 1u;
    // frames per batch: enough paths at depth 0 to keep every SM busy through the thin bounces (about 2 M), at most kWfMaxBatch
    const uint32_t per_frame = P.local_rows * width;
    uint32_t batch = per_frame ? (uint32_t)std::min<uint64_t>(kWfMaxBatch, std::max<uint64_t>(1u, (1ull << 21) / per_frame)) : 1u;
    if (const char* e = getenv("F3D_B200_WF_BATCH")) batch = (uint32_t)std::min<long>(kWfMaxBatch, std::max<long>(1, atol(e)));
    batch = std::min(batch, spp_frames);
    // bounces [0, wide_depth) run as compacted waves, the tail kernel walks the rest; any split gives the same image
    uint32_t wide_depth = kWfWideDepth;
    if (const char* e = getenv("F3D_B200_WF_WIDE_DEPTH")) wide_depth = (uint32_t)std::min<long>(kWfMaxDepth - 1u, std::max<long>(1, atol(e)));
    const uint32_t nbatches = (spp_frames + batch - 1u) / batch;
    if ((uint64_t)per_frame * batch > 0x7FFFFFFFull) return fail(F3D_ERR_ARGUMENT, "image too large for the 32-bit ray queues");
    const size_t qcap = std::max<size_t>(1, (size_t)per_frame * batch);
    for (int q = 0; q < 2; q++) {
        if ((rc = R.alloc(&P.qa[q], qcap))) return rc;
        if ((rc = R.alloc(&P.qb[q], qcap))) return rc;
        if ((rc = R.alloc(&P.qc[q], qcap))) return rc;
    }
    if ((rc = R.alloc(&P.fsum, (size_t)npx * batch))) return rc;
    uint32_t *d_counts = nullptr, *d_qcount = nullptr;
    if ((rc = R.alloc(&d_counts, (size_t)spp_frames * kSlots))) return rc;
    if ((rc = R.alloc(&d_qcount, (size_t)nbatches * kSlots))) return rc;
    float4* d_hdr = nullptr;
    uchar4* d_rgba = nullptr;
    if (hdr_rgba && (rc = R.alloc(&d_hdr, npx))) return rc;
    if (rgba8 && (rc = R.alloc(&d_rgba, npx))) return rc;
    CUDA_TRY(cudaMemset(P.accum, 0, (size_t)npx * sizeof(float4)));
    CUDA_TRY(cudaMemset(d_counts, 0, (size_t)spp_frames * kSlots * sizeof(uint32_t)));
    CUDA_TRY(cudaMemset(d_qcount, 0, (size_t)nbatches * kSlots * sizeof(uint32_t)));
    CUDA_TRY(cudaEventCreate(&R.ev0));
    CUDA_TRY(cudaEventCreate(&R.ev1));

    int sms = 148;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const uint32_t full = (npx + kWfThreads - 1u) / kWfThreads;
    const uint32_t owned = std::max(1u, (per_frame + kWfThreads - 1u) / kWfThreads);
    const uint32_t first = (uint32_t)std::max<size_t>(1, (qcap + kWfThreads - 1u) / kWfThreads);
    const uint32_t wide = std::min(first, (uint32_t)sms * 8u);   // grid-stride over the device-side queue length
    const uint32_t thin = std::min(first, (uint32_t)sms * 4u);
    uint32_t launches = 0;
    CUDA_TRY(cudaEventRecord(R.ev0, 0));
    for (uint32_t b = 0; b < nbatches; b++) {
        WfBatch B{};
        B.first_frame = b * batch;
        B.nframes = std::min(batch, spp_frames - B.first_frame);
        for (uint32_t f = 0; f < B.nframes; f++) {
            const uint32_t fr = B.first_frame + f;
            B.seed_hi[f] = splitmix32(sc->seed_hi ^ fr);                  // adjudication.rs:237-239
            B.seed_lo[f] = splitmix32(sc->seed_lo ^ (fr * 0x00009E3Du));
            sobol2(fr, &B.u1[f], &B.u2[f]);                               // sidx = sample + frame_index * max(1, spp), spp = 1
        }
        B.qcount = d_qcount + (size_t)b * kSlots;
        const uint32_t g0 = std::max(1u, (per_frame * B.nframes + kWfThreads - 1u) / kWfThreads);
        k_wf_bounce<true><<<g0, kWfThreads>>>(P, B, 0u);
        for (uint32_t d = 1; d < wide_depth; d++) k_wf_bounce<false><<<wide, kWfThreads>>>(P, B, d);
        k_wf_tail<<<thin, kWfThreads>>>(P, B, wide_depth);
        k_wf_merge<<<owned, kWfThreads, B.nframes * kSlots * sizeof(uint32_t)>>>(P, B.first_frame, B.nframes, d_counts);
        launches += wide_depth + 2u;
        if ((b & 15u) == 15u) CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaGetLastError());
    if (d_hdr || d_rgba) {
        k_wf_resolve<<<full, kWfThreads>>>(P.accum, npx, 1.0f / (float)spp_frames, sc->exposure, d_hdr, d_rgba);
        CUDA_TRY(cudaGetLastError());
        launches++;
    }
    CUDA_TRY(cudaEventRecord(R.ev1, 0));
    CUDA_TRY(cudaEventSynchronize(R.ev1));

    // the two frame rules of the reference, applied to the device-side queue lengths
    std::vector<uint32_t> counts((size_t)spp_frames * kSlots);
    CUDA_TRY(cudaMemcpy(counts.data(), d_counts, counts.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    const uint64_t capacity = 4ull * npx;   // wavefront/mod.rs:34,77
    uint64_t total = 0, max_rays = 0;
    uint32_t min_iters = 0xFFFFFFFFu;
    for (uint32_t fr = 0; fr < spp_frames; fr++) {
        // hist[w] = paths of this frame that traced exactly w rays (k_wf_merge); rays at depth k = paths with w > k
        const uint32_t* hist = counts.data() + (size_t)fr * kSlots;
        uint64_t alive = 0, cum = 0;
        for (uint32_t w = 1; w <= kWfMaxDepth; w++) alive += hist[w];
        if (alive != n_primary || hist[0])
            return fail(F3D_ERR_DEVICE, "wavefront frame %u: %llu paths accounted for, %u traced", fr, (unsigned long long)alive, n_primary);
        uint32_t executed = 0;
        for (uint32_t k = 0; k < kWfMaxDepth; k++) {
            const uint64_t rays = alive;               // paths still tracing at depth k
            if (!rays) break;
            cum += rays;
            if (!part && cum > capacity)   // render.rs:127-137
                return fail(F3D_ERR_RENDER, "wavefront frame %u: wavefront ray queue overflow: %llu rays pushed into capacity %llu", fr,
                            (unsigned long long)cum, (unsigned long long)capacity);
            executed++;
            alive -= hist[k + 1u];
        }
        if (!part && executed < 2u)   // adjudication.rs:259-265
            return fail(F3D_ERR_RENDER,
                        "adjudication PT frame %u executed %u wavefront iteration(s); a multi-bounce path-traced reference requires >= 2", fr,
                        executed);
        if (part && part->frame_iterations) part->frame_iterations[fr] = executed;
        if (part && part->frame_rays) part->frame_rays[fr] = cum;
        total += cum;
        max_rays = std::max(max_rays, cum);
        min_iters = std::min(min_iters, executed);
    }
    if (hdr_rgba) CUDA_TRY(cudaMemcpy(hdr_rgba, d_hdr, (size_t)npx * sizeof(float4), cudaMemcpyDeviceToHost));
    if (rgba8) CUDA_TRY(cudaMemcpy(rgba8, d_rgba, (size_t)npx * sizeof(uchar4), cudaMemcpyDeviceToHost));
    if (stats) {
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, R.ev0, R.ev1));
        stats->rays = total; stats->max_rays_per_frame = max_rays; stats->min_iterations = min_iters; stats->launches = launches;
        stats->kernel_ms = ms;
    }
    return 0;
}
