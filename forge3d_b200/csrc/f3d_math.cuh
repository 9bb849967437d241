// forge3d_b200/csrc/f3d_math.cuh
// Device math for the terrain path tracer under the numerics contract of DESIGN.md section 4:
// IEEE binary32, round-to-nearest, NO FMA contraction (this translation unit is compiled with
// -fmad=false; division/sqrt/reciprocal use the correctly rounded intrinsics explicitly), and the
// pinned definitions of the operations WGSL leaves to the driver:
//   dot3(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z,  normalize(v) = v * (1/sqrt(dot3(v,v))),
//   mix(a,b,t) = a*(1-t) + b*t,  sin/cos/atan2/acos = Cephes single-precision kernels.
// Reference: /root/reference/src/shaders/hybrid_terrain_traversal.wgsl (helpers :88-91,:392-431),
// hybrid_kernel.wgsl:78-85,109-112.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace f3d {

struct v3 { float x, y, z; };

__device__ __forceinline__ v3 V3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ v3 operator+(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ v3 operator-(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ v3 operator*(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ v3 operator*(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ v3 operator-(v3 a) { return V3(-a.x, -a.y, -a.z); }
// F3D_FAST_NUMERICS = 1 builds the THROUGHPUT variant of the library (libforge3d_b200_fast.so, forge3d_b200/build.py
// numerics="fast": also -fmad=true): division, reciprocal and square root are the SFU approximations (2 ulp / 1 ulp) instead
// of the correctly rounded sequences, normalize uses rsqrt, and the compiler may contract a*b+c.  Outputs are then no longer
// bit-identical to the oracle; the variant is accepted by the north-star tolerance instead (RGBA RMSE <= 1e-3 against the exact
// build, tests/test_gpu_parity.py).  0 (default) = the numerics contract above.
#ifndef F3D_FAST_NUMERICS
#define F3D_FAST_NUMERICS 0
#endif
#if F3D_FAST_NUMERICS && defined(__CUDA_ARCH__)
__device__ __forceinline__ float fdiv(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ float frcp(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float fsqrt(float a) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float frsqrt(float a) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
#else
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float frcp(float a) { return __frcp_rn(a); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float frsqrt(float a) { return __frcp_rn(__fsqrt_rn(a)); }
#endif
// Requests a line into L1 without a destination register (no-op on the host builds of this header).
__device__ __forceinline__ void prefetch_l1(const void* p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
__device__ __forceinline__ float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ float dot2(float ax, float ay, float bx, float by) { return ax * bx + ay * by; }
__device__ __forceinline__ v3 cross3(v3 a, v3 b) {
    return V3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ v3 normalize3(v3 a) { return a * frsqrt(dot3(a, a)); }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float luminance(v3 c) { return dot3(c, V3(0.2126f, 0.7152f, 0.0722f)); }

// xorshift32, hybrid_kernel.wgsl:78-85.  u32 -> f32 is round-to-nearest, /2^32 is exact.
__device__ __forceinline__ float xorshift32(uint32_t& st) {
    uint32_t x = st;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    st = x;
    return __uint2float_rn(x) * 2.3283064365386963e-10f;
}

// terrain_tent_offset, hybrid_terrain_traversal.wgsl:409-414
__device__ __forceinline__ float tent_offset(float u) {
    if (u < 0.5f) return fsqrt(2.0f * u) - 1.0f;
    return 1.0f - fsqrt(2.0f * (1.0f - u));
}

// Pinned sin/cos for phi in [0, ~2*pi] (Cephes sinf/cosf kernels, 3-term Cody-Waite pi/2).
__device__ __forceinline__ void sincos_pinned(float x, float& s, float& c) {
    int k = (int)(x * 0.636619772f + 0.5f);
    float fk = (float)k;
    float r = x - fk * 1.5703125f;
    r = r - fk * 4.837512969970703125e-4f;
    r = r - fk * 7.549789954891882e-8f;
    float z = r * r;
    float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z
               - 0.5f * z + 1.0f;
    switch (k & 3) {
        case 0: s = sp; c = cp; break;
        case 1: s = cp; c = -sp; break;
        case 2: s = -sp; c = -cp; break;
        default: s = -cp; c = sp; break;
    }
}

__device__ __forceinline__ float atan_pos(float x) {
    float y;
    if (x > 2.414213562373095f) { y = 1.5707963267948966f; x = -frcp(x); }
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483f; x = fdiv(x - 1.0f, x + 1.0f); }
    else { y = 0.0f; }
    float z = x * x;
    y = y + ((((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z
              - 3.33329491539e-1f) * z * x + x);
    return y;
}

__device__ __forceinline__ float atan2_pinned(float y, float x) {
    const float PI_F = 3.14159265358979323846f;
    const float HALF_PI_F = 1.5707963267948966f;
    if (x == 0.0f) {
        if (y > 0.0f) return HALF_PI_F;
        if (y < 0.0f) return -HALF_PI_F;
        return 0.0f;
    }
    float q = fdiv(y, x);
    float a = atan_pos(fabsf(q));
    if (q < 0.0f) a = -a;
    if (x < 0.0f) a = (y >= 0.0f) ? a + PI_F : a - PI_F;
    return a;
}

__device__ __forceinline__ float asin_core(float x) {
    float a = fabsf(x);
    float z, w;
    bool flag = false;
    if (a > 0.5f) { z = 0.5f * (1.0f - a); w = fsqrt(z); flag = true; }
    else { w = a; z = w * w; }
    float p = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z
                + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * w + w;
    if (flag) { p = p + p; p = 1.5707963267948966f - p; }
    return x < 0.0f ? -p : p;
}

__device__ __forceinline__ float acos_pinned(float x) {
    if (x < -0.5f) return 3.14159265358979323846f - 2.0f * asin_core(fsqrt(0.5f * (1.0f + x)));
    if (x > 0.5f) return 2.0f * asin_core(fsqrt(0.5f * (1.0f - x)));
    return 1.5707963267948966f - asin_core(x);
}

// terrain_cosine_dir, hybrid_terrain_traversal.wgsl:421-431
__device__ __forceinline__ v3 cosine_dir(v3 n, float u1, float u2) {
    const float PI_F = 3.14159265358979323846f;
    float sign = n.z < 0.0f ? -1.0f : 1.0f;
    float a = -frcp(sign + n.z);
    float b = n.x * n.y * a;
    v3 t = V3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    v3 bt = V3(b, sign + n.y * n.y * a, -n.y);
    float r = fsqrt(u1);
    float phi = 2.0f * PI_F * u2;
    float sphi, cphi;
    sincos_pinned(phi, sphi, cphi);
    v3 local = V3(r * cphi, r * sphi, fsqrt(fmaxf(0.0f, 1.0f - u1)));
    v3 d = (t * local.x + bt * local.y) + n * local.z;
    return normalize3(d);
}

}  // namespace f3d
