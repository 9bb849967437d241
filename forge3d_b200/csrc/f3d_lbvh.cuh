// forge3d_b200/csrc/f3d_lbvh.cuh
// GPU LBVH build for mesh geometry (second half; north-star "src/accel"): Morton codes ->
// sort -> Karras topology -> bottom-up boxes, all on the device.  Replaces
//   /root/reference/src/shaders/lbvh_morton.wgsl:24-65        expand_bits, morton3d, main
//   /root/reference/src/accel/lbvh_gpu/{morton,sort,sort_bitonic,topology,refit}.rs + src/shaders/radix_sort_pairs.wgsl,
//   bvh_refit.wgsl (sort + refit)
//   /root/reference/src/shaders/lbvh_link.wgsl:35-181         delta, determine_range, find_split, link_nodes, init_leaves
// The reference builds this tree but never traverses it on the path-traced snapshot path (its intersect_mesh sweeps every
// triangle, hybrid_traversal.wgsl:137-172); here the tree feeds intersect_mesh (f3d_trace.cuh), whose closest hit is
// independent of the tree's shape (ties resolve to the lowest triangle index), so renders stay bit-identical.
//
// Differences by design: (0) ranges AND splits use the composite key (see k_lbvh_link); (1) the sort key is the 64-bit composite (morton << 32 | triangle index), which makes the order
// total - the (code, index) order the reference's delta() tie-break assumes - so a bitonic network over u64 needs no
// stability argument (the reference's radix scatter is not stable for equal codes); (2) leaf boxes are padded like the
// host builder's (f3d_backend.cu::build_mesh_bvh) so that rounding in the slab test can never cull the triangle the
// Moeller-Trumbore test would hit; (3) nodes are written straight into the traversal format.
#pragma once
#include "f3d_math.cuh"

namespace f3d {

constexpr uint32_t kBvhLeafFlag = 0x80000000u;   // n0.w of a leaf: flag | first index into bvh_tris; n1.w = triangle count

__device__ __forceinline__ uint32_t lbvh_expand_bits(uint32_t v) {        // lbvh_morton.wgsl:24-31
    uint32_t x = v & 0x000003ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

// One thread per triangle: Morton code of the centroid (lbvh_morton.wgsl:40-64; centroid = Triangle::centroid,
// src/accel/types.rs:173-179) and the padded triangle box.
__global__ void k_lbvh_prims(const float4* __restrict__ verts, const uint32_t* __restrict__ idx, uint32_t ntris, v3 world_min,
                             v3 world_extent, unsigned long long* __restrict__ keys, uint32_t padded_n,
                             float4* __restrict__ tri_boxes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= padded_n) return;
    if (i >= ntris) { keys[i] = 0xFFFFFFFFFFFFFFFFull; return; }       // bitonic padding sorts to the end
    const float4 a = __ldg(verts + __ldg(idx + 3 * (size_t)i)), b = __ldg(verts + __ldg(idx + 3 * (size_t)i + 1)),
                 c = __ldg(verts + __ldg(idx + 3 * (size_t)i + 2));
    const float cx = fdiv(a.x + b.x + c.x, 3.0f), cy = fdiv(a.y + b.y + c.y, 3.0f), cz = fdiv(a.z + b.z + c.z, 3.0f);
    const float nx = clampf(fdiv(cx - world_min.x, fmaxf(world_extent.x, 1e-6f)), 0.0f, 1.0f);
    const float ny = clampf(fdiv(cy - world_min.y, fmaxf(world_extent.y, 1e-6f)), 0.0f, 1.0f);
    const float nz = clampf(fdiv(cz - world_min.z, fmaxf(world_extent.z, 1e-6f)), 0.0f, 1.0f);
    const uint32_t gx = min((uint32_t)(nx * 1023.0f), 1023u), gy = min((uint32_t)(ny * 1023.0f), 1023u), gz = min((uint32_t)(nz * 1023.0f), 1023u);
    const uint32_t code = lbvh_expand_bits(gx) | (lbvh_expand_bits(gy) << 1) | (lbvh_expand_bits(gz) << 2);
    keys[i] = ((unsigned long long)code << 32) | (unsigned long long)i;
    // padded triangle box (same pad as the host builder)
    const float lox = fminf(a.x, fminf(b.x, c.x)), hix = fmaxf(a.x, fmaxf(b.x, c.x));
    const float loy = fminf(a.y, fminf(b.y, c.y)), hiy = fmaxf(a.y, fmaxf(b.y, c.y));
    const float loz = fminf(a.z, fminf(b.z, c.z)), hiz = fmaxf(a.z, fmaxf(b.z, c.z));
    const float ext = fmaxf(fmaxf(hix - lox, hiy - loy), hiz - loz);
    const float scale = fmaxf(fmaxf(fmaxf(fabsf(lox), fabsf(hix)), fmaxf(fabsf(loy), fabsf(hiy))), fmaxf(fabsf(loz), fabsf(hiz)));
    const float pad = 1e-3f * ext + 1e-5f * scale + 1e-6f;
    tri_boxes[2 * (size_t)i] = make_float4(lox - pad, loy - pad, loz - pad, 0.0f);
    tri_boxes[2 * (size_t)i + 1] = make_float4(hix + pad, hiy + pad, hiz + pad, 0.0f);
}

// One compare-exchange stage (k, j) of the bitonic network over `n` (power of two) 64-bit keys.
__global__ void k_lbvh_bitonic(unsigned long long* __restrict__ keys, uint32_t n, uint32_t k, uint32_t j) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t partner = i ^ j;
    if (partner <= i) return;
    const unsigned long long a = keys[i], b = keys[partner];
    const bool ascending = (i & k) == 0u;
    if ((a > b) == ascending) { keys[i] = b; keys[partner] = a; }
}

// All stages of the network that stay inside one CTA's tile of kBitonicTile keys, in shared memory.
//   k_from == 2: the full sort of every tile (k = 2 .. tile);  otherwise: the tail j = tile/2 .. 1 of merge step k_from.
constexpr uint32_t kBitonicTile = 2048u;      // keys per CTA (16 KB of shared memory), 1024 threads
__global__ void __launch_bounds__(1024) k_lbvh_bitonic_tile(unsigned long long* __restrict__ keys, uint32_t n, uint32_t k_from) {
    extern __shared__ __align__(16) unsigned char smem_raw[];       // kBitonicTile * 8 bytes, passed at launch
    unsigned long long* tile = reinterpret_cast<unsigned long long*>(smem_raw);
    const uint32_t base = blockIdx.x * kBitonicTile;
    for (uint32_t t = threadIdx.x; t < kBitonicTile; t += blockDim.x) tile[t] = base + t < n ? keys[base + t] : 0xFFFFFFFFFFFFFFFFull;
    __syncthreads();
    const uint32_t k_first = k_from == 2u ? 2u : k_from, k_last = k_from == 2u ? kBitonicTile : k_from;
    for (uint32_t k = k_first; k <= k_last; k <<= 1) {
        for (uint32_t j = min(k >> 1, kBitonicTile >> 1); j > 0u; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < kBitonicTile / 2u; t += blockDim.x) {
                const uint32_t lo = 2u * t - (t & (j - 1u));            // index with bit j clear
                const uint32_t hi = lo | j;
                const unsigned long long a = tile[lo], b = tile[hi];
                const bool ascending = ((base + lo) & k) == 0u;
                if ((a > b) == ascending) { tile[lo] = b; tile[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (uint32_t t = threadIdx.x; t < kBitonicTile; t += blockDim.x)
        if (base + t < n) keys[base + t] = tile[t];
}

// delta(i, j) of lbvh_link.wgsl:35-55 on the composite keys: clz of the code XOR, or 32 + clz of the index XOR when the codes
// are equal - which is exactly the count of leading zeros of the 64-bit XOR.
__device__ __forceinline__ int lbvh_delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    if (i == j) return 32;                                  // (the shader's value for identical indices; never compared)
    return __clzll((long long)(keys[i] ^ keys[j]));
}

// link_nodes (lbvh_link.wgsl:117-163): one thread per internal node.  left / right hold NODE indices (internal i, leaf n-1+i).
__global__ void k_lbvh_link(const unsigned long long* __restrict__ keys, uint32_t nprims, uint32_t* __restrict__ left,
                            uint32_t* __restrict__ right, uint32_t* __restrict__ parent) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int n = (int)nprims;
    if (i >= n - 1) return;
    // determine_range, :63-88
    const int dd = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1);
    const int d = dd > 0 ? 1 : (dd < 0 ? -1 : 0);
    const int delta_min = lbvh_delta(keys, n, i, i - d);
    int l_max = 2;
    while (lbvh_delta(keys, n, i, i + l_max * d) > delta_min) l_max *= 2;
    int l = 0;
    for (int t = l_max / 2; t >= 1; t /= 2)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > delta_min) l += t;
    const int jj = i + l * d;
    const int first = min(i, jj), last = max(i, jj);
    // find_split, :90-115, on the composite keys.  For distinct Morton codes this is the shader's search; for equal codes the
    // shader splits at the midpoint, which is inconsistent with the index-prefix ranges its own delta() produces (children
    // would cover other ranges than their parents assume) - the unique 64-bit keys give a proper tree in every case.
    const unsigned long long first_key = keys[first];
    const int common_prefix = __clzll((long long)(first_key ^ keys[last]));
    int split = first;
    int current_step = last - first;
    while (current_step > 1) {
        current_step = (current_step + 1) >> 1;
        const int new_split = split + current_step;
        if (new_split < last && __clzll((long long)(first_key ^ keys[new_split])) > common_prefix) split = new_split;
    }
    const uint32_t lc = split == first ? nprims - 1u + (uint32_t)split : (uint32_t)split;
    const uint32_t rc = split + 1 == last ? nprims - 1u + (uint32_t)(split + 1) : (uint32_t)(split + 1);
    left[i] = lc;
    right[i] = rc;
    parent[lc] = (uint32_t)i;
    parent[rc] = (uint32_t)i;
    if (i == 0) parent[0] = 0xFFFFFFFFu;
}

// init_leaves (:165-181) + bottom-up boxes (bvh_refit.wgsl): one thread per leaf writes its node and climbs; the second
// thread to reach an internal node owns it (both children are complete) and writes the union.  min / max are exact, so the
// boxes do not depend on arrival order.
__global__ void k_lbvh_refit(const unsigned long long* __restrict__ keys, uint32_t nprims, const uint32_t* __restrict__ left,
                             const uint32_t* __restrict__ right, const uint32_t* __restrict__ parent, const float4* __restrict__ tri_boxes,
                             uint32_t* __restrict__ arrivals, float4* __restrict__ nodes, uint32_t* __restrict__ order) {
    const uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= nprims) return;
    const uint32_t tri = (uint32_t)(keys[leaf] & 0xFFFFFFFFull);
    order[leaf] = tri;
    uint32_t node = nprims - 1u + leaf;
    float4 lo = tri_boxes[2 * (size_t)tri], hi = tri_boxes[2 * (size_t)tri + 1];
    lo.w = __uint_as_float(kBvhLeafFlag | leaf);
    hi.w = __uint_as_float(1u);
    nodes[2 * (size_t)node] = lo;
    nodes[2 * (size_t)node + 1] = hi;
    if (nprims == 1u) return;
    while (true) {
        __threadfence();
        const uint32_t p = parent[node];
        if (p == 0xFFFFFFFFu) return;
        if (atomicAdd(arrivals + p, 1u) == 0u) return;       // first child to arrive: the sibling will finish the parent
        __threadfence();
        const uint32_t lc = left[p], rc = right[p];
        const volatile float4* vn = nodes;
        const float4 a0 = make_float4(vn[2 * (size_t)lc].x, vn[2 * (size_t)lc].y, vn[2 * (size_t)lc].z, 0.0f);
        const float4 a1 = make_float4(vn[2 * (size_t)lc + 1].x, vn[2 * (size_t)lc + 1].y, vn[2 * (size_t)lc + 1].z, 0.0f);
        const float4 b0 = make_float4(vn[2 * (size_t)rc].x, vn[2 * (size_t)rc].y, vn[2 * (size_t)rc].z, 0.0f);
        const float4 b1 = make_float4(vn[2 * (size_t)rc + 1].x, vn[2 * (size_t)rc + 1].y, vn[2 * (size_t)rc + 1].z, 0.0f);
        nodes[2 * (size_t)p] = make_float4(fminf(a0.x, b0.x), fminf(a0.y, b0.y), fminf(a0.z, b0.z), __uint_as_float(lc));
        nodes[2 * (size_t)p + 1] = make_float4(fmaxf(a1.x, b1.x), fmaxf(a1.y, b1.y), fmaxf(a1.z, b1.z), __uint_as_float(rc));
        node = p;
    }
}

}  // namespace f3d
