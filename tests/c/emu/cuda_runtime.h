/* Host stand-in for <cuda_runtime.h>, used ONLY by tests/c/emu/trace_emu.cpp to compile the product's device headers
 * (forge3d_b200/csrc/f3d_{math,trace,trace_fast}.cuh) with g++ and run them on the CPU as a ONE-LANE warp.
 * Test infrastructure: lets `-m "not gpu"` tests check that a restructured traversal is still bit-identical to the
 * oracle before any GPU time is spent.  IEEE semantics are obtained with -O2 -ffp-contract=off -fno-fast-math on
 * x86-64 SSE (no excess precision); min/max follow the CUDA definitions (NaN -> other operand, -0 < +0). */
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef __cplusplus   /* libstdc++ spells attributes __noinline__ etc.: parse it before the CUDA keywords become macros */
#include <algorithm>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#endif

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __noinline__ __attribute__((noinline))

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

template <class T> static inline T __ldg(const T* p) { return *p; }

static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline int32_t __float_as_int(float f) { int32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int32_t u) { float f; memcpy(&f, &u, 4); return f; }

static inline float emu_fminf(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.0f && b == 0.0f) return signbit(a) ? a : b;
    return a < b ? a : b;
}
static inline float emu_fmaxf(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.0f && b == 0.0f) return signbit(a) ? b : a;
    return a > b ? a : b;
}
#define fminf emu_fminf
#define fmaxf emu_fmaxf

static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __double2float_ru(double x) { float f = (float)x; if ((double)f < x) f = nextafterf(f, INFINITY); return f; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }   /* glibc fmaf is correctly rounded */
static inline float __uint2float_rn(uint32_t x) { return (float)x; }

static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

#ifndef EMU_SIMT
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
#endif
#ifdef EMU_SIMT
#include "simt.h"          /* 32-lane warps and CTAs as fibers; defines the *_sync collectives and dim3 */
#include "cuda_fake_runtime.h"
#else
/* a warp of one lane */
static inline uint32_t __ballot_sync(uint32_t, int pred) { return pred ? 1u : 0u; }
static inline uint32_t __activemask() { return 1u; }
template <class T> static inline T __shfl_sync(uint32_t, T v, int, int = 32) { return v; }
static inline void __syncwarp(uint32_t = 0xFFFFFFFFu) {}
/* the CTA-wide helpers of f3d_trace_fast.cuh (stage_top_levels, all_sibling_seeds_warp) only have to COMPILE here:
 * trace_emu.cpp runs trace_fast, which calls neither */
static const struct { uint32_t x, y, z; } threadIdx = {0u, 0u, 0u};
static inline void __syncthreads() {}
#endif
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
