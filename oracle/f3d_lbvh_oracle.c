/*
 * oracle/f3d_lbvh_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C restatement of the reference's GPU LBVH build (second half):
 *   src/shaders/lbvh_morton.wgsl:24-65 + src/accel/lbvh_gpu/morton.rs:14-27   Morton code of the triangle centroid
 *   src/accel/types.rs:173-188,298-313                                        centroid, triangle box, scene box
 *   src/accel/lbvh_gpu/sort.rs, sort_bitonic.rs                               sort of (code, index) pairs
 *   src/shaders/lbvh_link.wgsl:35-181                                         delta, determine_range, find_split, link_nodes, init_leaves
 *   src/shaders/bvh_refit.wgsl                                                parent box = union of the children's boxes
 * Integer work: compared bit for bit with the CUDA build (tests/test_lbvh.py).
 *
 * Two documented definitions where the reference is loose: the sort is by (code, triangle index) - the order its delta()
 * tie-break assumes, which its unstable radix scatter does not guarantee; and `literal_split == 0` finds the split on the
 * 64-bit composite key (code << 32 | index), which equals the shader's find_split whenever the two end codes differ and, for
 * equal codes, replaces its midpoint rule (inconsistent with the ranges delta() yields) by the index prefix.  Leaf boxes carry
 * the traversal's safety pad (f3d_backend.cu::build_mesh_bvh); pass pad = 0 for the reference's raw triangle boxes.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "f3d_oracle.h"

static uint32_t expand_bits(uint32_t v) {   /* lbvh_morton.wgsl:24-31 */
    uint32_t x = v & 0x000003ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}
static int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

typedef struct { const uint64_t* keys; int n; } link_ctx;
static int delta(const link_ctx* L, int i, int j) {   /* lbvh_link.wgsl:35-55 */
    if (j < 0 || j >= L->n) return -1;
    if (i == j) return 32;
    uint32_t ci = (uint32_t)(L->keys[i] >> 32), cj = (uint32_t)(L->keys[j] >> 32);
    if (ci == cj) return 32 + clz32((uint32_t)L->keys[i] ^ (uint32_t)L->keys[j]);
    return clz32(ci ^ cj);
}

int f3do_lbvh_build(const float* xyz, uint32_t nverts, const uint32_t* idx, uint32_t ntris, int literal_split, int pad_boxes,
                    uint32_t* morton, uint32_t* order, uint32_t* left, uint32_t* right, uint32_t* parent, float* nodes) {
    (void)nverts;
    if (ntris == 0) return 1;
    const int n = (int)ntris;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t t = 0; t < (size_t)ntris * 3; t++)
        for (int a = 0; a < 3; a++) { float v = xyz[3 * (size_t)idx[t] + a]; mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
    float ext[3];
    for (int a = 0; a < 3; a++) ext[a] = fmaxf(mx[a] - mn[a], 1e-6f);
    uint64_t* keys = (uint64_t*)malloc((size_t)ntris * sizeof(uint64_t));
    float* boxes = (float*)malloc((size_t)ntris * 6 * sizeof(float));
    if (!keys || !boxes) { free(keys); free(boxes); return 1; }
    for (int t = 0; t < n; t++) {
        const float* v0 = xyz + 3 * (size_t)idx[3 * t], *v1 = xyz + 3 * (size_t)idx[3 * t + 1], *v2 = xyz + 3 * (size_t)idx[3 * t + 2];
        uint32_t g[3];
        float lo[3], hi[3];
        for (int a = 0; a < 3; a++) {
            float c = (v0[a] + v1[a] + v2[a]) / 3.0f;
            float nrm = (c - mn[a]) / fmaxf(ext[a], 1e-6f);
            float cl = fminf(fmaxf(nrm, 0.0f), 1.0f);
            uint32_t q = (uint32_t)(cl * 1023.0f);
            g[a] = q < 1023u ? q : 1023u;
            lo[a] = fminf(v0[a], fminf(v1[a], v2[a]));
            hi[a] = fmaxf(v0[a], fmaxf(v1[a], v2[a]));
        }
        uint32_t code = expand_bits(g[0]) | (expand_bits(g[1]) << 1) | (expand_bits(g[2]) << 2);
        keys[t] = ((uint64_t)code << 32) | (uint64_t)t;
        float pad = 0.0f;
        if (pad_boxes) {
            float e = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
            float s = fmaxf(fmaxf(fmaxf(fabsf(lo[0]), fabsf(hi[0])), fmaxf(fabsf(lo[1]), fabsf(hi[1]))), fmaxf(fabsf(lo[2]), fabsf(hi[2])));
            pad = 1e-3f * e + 1e-5f * s + 1e-6f;
        }
        for (int a = 0; a < 3; a++) { boxes[6 * t + a] = lo[a] - pad; boxes[6 * t + 3 + a] = hi[a] + pad; }
    }
    qsort(keys, (size_t)n, sizeof(uint64_t), cmp_u64);
    for (int i = 0; i < n; i++) {
        if (morton) morton[i] = (uint32_t)(keys[i] >> 32);
        if (order) order[i] = (uint32_t)keys[i];
    }
    const size_t nnodes = 2 * (size_t)ntris - 1;
    uint32_t* L = (uint32_t*)malloc((size_t)ntris * sizeof(uint32_t));
    uint32_t* R = (uint32_t*)malloc((size_t)ntris * sizeof(uint32_t));
    uint32_t* P = (uint32_t*)malloc(nnodes * sizeof(uint32_t));
    float* N = (float*)calloc(nnodes * 8, sizeof(float));
    memset(P, 0xFF, nnodes * sizeof(uint32_t));
    link_ctx C = {keys, n};
    for (int i = 0; i < n - 1; i++) {   /* link_nodes, :117-163 */
        int dd = delta(&C, i, i + 1) - delta(&C, i, i - 1);
        int d = dd > 0 ? 1 : (dd < 0 ? -1 : 0);
        int delta_min = delta(&C, i, i - d);
        int l_max = 2;
        while (delta(&C, i, i + l_max * d) > delta_min) l_max *= 2;
        int l = 0;
        for (int t = l_max / 2; t >= 1; t /= 2)
            if (delta(&C, i, i + (l + t) * d) > delta_min) l += t;
        int j = i + l * d;
        int first = i < j ? i : j, last = i < j ? j : i;
        int split;
        uint32_t fc = (uint32_t)(keys[first] >> 32), lc = (uint32_t)(keys[last] >> 32);
        if (literal_split) {   /* find_split, :90-115 */
            if (fc == lc) split = (first + last) >> 1;
            else {
                int common = clz32(fc ^ lc);
                split = first;
                int step = last - first;
                while (step > 1) {
                    step = (step + 1) >> 1;
                    int ns = split + step;
                    if (ns < last && clz32(fc ^ (uint32_t)(keys[ns] >> 32)) > common) split = ns;
                }
            }
        } else {
            int common = clz64(keys[first] ^ keys[last]);
            split = first;
            int step = last - first;
            while (step > 1) {
                step = (step + 1) >> 1;
                int ns = split + step;
                if (ns < last && clz64(keys[first] ^ keys[ns]) > common) split = ns;
            }
        }
        uint32_t lch = split == first ? (uint32_t)(n - 1 + split) : (uint32_t)split;
        uint32_t rch = split + 1 == last ? (uint32_t)(n - 1 + split + 1) : (uint32_t)(split + 1);
        L[i] = lch; R[i] = rch; P[lch] = (uint32_t)i; P[rch] = (uint32_t)i;
    }
    /* init_leaves (:165-181) in the traversal's node format, then boxes bottom-up (each internal node once both children are done) */
    for (int i = 0; i < n; i++) {
        float* nd = N + 8 * (size_t)(n - 1 + i);
        uint32_t tri = (uint32_t)keys[i], w0 = 0x80000000u | (uint32_t)i, w1 = 1u;
        memcpy(nd, boxes + 6 * tri, 12); memcpy(nd + 3, &w0, 4);
        memcpy(nd + 4, boxes + 6 * tri + 3, 12); memcpy(nd + 7, &w1, 4);
    }
    if (!literal_split && n > 1) {
        uint8_t* seen = (uint8_t*)calloc((size_t)n, 1);
        for (int i = 0; i < n; i++) {
            uint32_t node = (uint32_t)(n - 1 + i);
            for (;;) {
                uint32_t p = P[node];
                if (p == 0xFFFFFFFFu) break;
                if (!seen[p]++) break;
                float* pn = N + 8 * (size_t)p; const float* a = N + 8 * (size_t)L[p]; const float* b = N + 8 * (size_t)R[p];
                for (int k = 0; k < 3; k++) { pn[k] = fminf(a[k], b[k]); pn[4 + k] = fmaxf(a[4 + k], b[4 + k]); }
                memcpy(pn + 3, &L[p], 4); memcpy(pn + 7, &R[p], 4);
                node = p;
            }
        }
        free(seen);
    }
    if (left && n > 1) memcpy(left, L, (size_t)(n - 1) * 4);
    if (right && n > 1) memcpy(right, R, (size_t)(n - 1) * 4);
    if (parent) memcpy(parent, P, nnodes * 4);
    if (nodes) memcpy(nodes, N, nnodes * 8 * 4);
    free(keys); free(boxes); free(L); free(R); free(P); free(N);
    return 0;
}
