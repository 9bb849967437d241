/* Cooperative SIMT interpreter for the CPU emulation of the product's CUDA kernels (test infrastructure).
 * One CTA at a time; every CUDA thread is a fiber (own stack, hand-written x86-64 context switch); warp collectives
 * (__ballot_sync, __shfl_*_sync, __any_sync, __syncwarp) and __syncthreads are rendezvous points: a lane deposits its
 * operand, yields to the scheduler, and is resumed once every live lane of its warp (CTA) has arrived.  Deterministic,
 * single OS thread, so plain loads/stores implement atomics and fences.  Exited lanes do not take part in collectives. */
#pragma once
#if !defined(__x86_64__)
#error "the SIMT emulator's context switch is written for x86-64"
#endif
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };

extern uint3_emu threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

extern "C" void emu_switch(void** save_sp, void* load_sp);

namespace emu {

enum LaneState { READY = 0, WAIT_WARP = 1, WAIT_CTA = 2, DONE = 3 };

struct Lane {
    void* sp = nullptr;
    int state = DONE;
    uint32_t seq = 0;      // warp collectives executed so far
};

struct Cta {
    std::vector<Lane> lanes;
    std::vector<uint64_t> xbuf;   // [warp][parity][lane]
    unsigned nthreads = 0;
    const std::function<void()>* body = nullptr;
};

extern Cta g_cta;
extern unsigned g_cur;          // linear thread id of the running fiber
extern void* g_sched_sp;

inline void set_thread(unsigned tid) {
    g_cur = tid;
    threadIdx.x = tid % blockDim.x;
    threadIdx.y = (tid / blockDim.x) % blockDim.y;
    threadIdx.z = tid / (blockDim.x * blockDim.y);
}

inline void yield(int state) {
    const unsigned me = g_cur;
    g_cta.lanes[me].state = state;
    emu_switch(&g_cta.lanes[me].sp, g_sched_sp);
    set_thread(me);                                     // resumed: restore the built-in variables
}

inline unsigned lane_id() { return g_cur & 31u; }
inline unsigned warp_base() { return g_cur & ~31u; }
inline bool lane_live(unsigned tid) { return tid < g_cta.nthreads && g_cta.lanes[tid].state != DONE; }

// deposit `v`, rendezvous with the warp, return the slot array of this collective
inline const uint64_t* warp_exchange(uint64_t v) {
    Lane& L = g_cta.lanes[g_cur];
    uint64_t* slot = &g_cta.xbuf[((size_t)(g_cur >> 5) * 2u + (L.seq & 1u)) * 32u];
    slot[lane_id()] = v;
    L.seq++;
    yield(WAIT_WARP);
    return slot;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);

}  // namespace emu

// ---- warp / CTA collectives ----
static inline uint32_t __ballot_sync(uint32_t, int pred) {
    const unsigned base = emu::warp_base();
    const uint64_t* s = emu::warp_exchange(pred ? 1u : 0u);
    uint32_t r = 0;
    for (unsigned l = 0; l < 32u; l++)
        if (emu::lane_live(base + l) && s[l]) r |= 1u << l;
    return r;
}
static inline int __any_sync(uint32_t m, int pred) { return __ballot_sync(m, pred) != 0u; }
static inline int __all_sync(uint32_t m, int pred) { return __ballot_sync(m, !pred) == 0u; }
static inline void __syncwarp(uint32_t = 0xFFFFFFFFu) { emu::warp_exchange(0); }
static inline uint32_t __activemask() {
    uint32_t r = 0;
    for (unsigned l = 0; l < 32u; l++)
        if (emu::lane_live(emu::warp_base() + l)) r |= 1u << l;
    return r;
}
template <class T>
static inline T __shfl_sync(uint32_t, T v, int src, int = 32) {
    static_assert(sizeof(T) <= 8, "shuffle of a type wider than 64 bits");
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    const uint64_t* s = emu::warp_exchange(bits);
    T out;
    memcpy(&out, &s[(unsigned)src & 31u], sizeof(T));
    return out;
}
template <class T>
static inline T __shfl_xor_sync(uint32_t m, T v, int lane_mask, int w = 32) {
    return __shfl_sync(m, v, (int)(emu::lane_id() ^ (unsigned)lane_mask), w);
}
template <class T>
static inline T __shfl_down_sync(uint32_t m, T v, unsigned delta, int w = 32) {
    const unsigned src = emu::lane_id() + delta;
    return __shfl_sync(m, v, (int)(src < 32u ? src : emu::lane_id()), w);
}
template <class T>
static inline T __shfl_up_sync(uint32_t m, T v, unsigned delta, int w = 32) {
    const unsigned me = emu::lane_id();
    return __shfl_sync(m, v, (int)(me >= delta ? me - delta : me), w);
}
static inline void __syncthreads() { emu::yield(emu::WAIT_CTA); }
