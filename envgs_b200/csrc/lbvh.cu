// lbvh.cu — LBVH construction over surfel quads on the GPU (Morton order -> radix sort -> Karras topology ->
// bottom-up refit).  Replaces optixAccelBuild (TRACE/trace_surfels.cpp:47-120); rebuilt every training
// iteration (EVC/utils/optix_utils.py:71-85), so the whole build is asynchronous on the caller's stream.
#include <cub/cub.cuh>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include "lbvh.cuh"

namespace envgs {

static size_t lbvh_sort_temp_bytes(int P) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, P, 0, 64);
    return bytes;
}

template <typename T>
static inline void carve_one(char*& p, T*& ptr, size_t count) {
    size_t off = (reinterpret_cast<uintptr_t>(p) + 255) & ~(size_t)255;
    ptr = reinterpret_cast<T*>(off);
    p = reinterpret_cast<char*>(ptr + count);
}

void lbvh_carve(Lbvh& t, char* base, int P) {
    char* p = base;
    size_t n = (size_t)(P > 0 ? P : 1);
    carve_one(p, t.nodes, n);
    carve_one(p, t.nodes4, n);
    carve_one(p, t.prims, n);
    carve_one(p, t.leaf_lo, 4 * n);
    carve_one(p, t.leaf_hi, 4 * n);
    carve_one(p, t.keys, n);
    carve_one(p, t.keys_sorted, n);
    carve_one(p, t.vals, n);
    carve_one(p, t.vals_sorted, n);
    carve_one(p, t.parent, 2 * n);
    carve_one(p, t.flags, n);
    carve_one(p, t.bounds, (size_t)8);
    t.sort_temp_bytes = lbvh_sort_temp_bytes((int)n);
    carve_one(p, t.sort_temp, t.sort_temp_bytes);
    for (int k = 0; k < 2; k++) {
        carve_one(p, t.cl_lo[k], n);
        carve_one(p, t.cl_hi[k], n);
        carve_one(p, t.cl_link[k], n);
        carve_one(p, t.cl_slot[k], n);
    }
    carve_one(p, t.cl_nn, n);
    carve_one(p, t.cl_valid, n);
    carve_one(p, t.cl_scan, n);
    carve_one(p, t.cl_counter, (size_t)64);
    t.scan_temp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t.scan_temp_bytes, (int*)nullptr, (int*)nullptr, (int)n);
    carve_one(p, t.scan_temp, t.scan_temp_bytes);
    t.storage_bytes = (size_t)(p - base) + 256;
}

size_t lbvh_storage_bytes(int P) {
    Lbvh t;
    lbvh_carve(t, nullptr, P);
    return t.storage_bytes + 256;
}

// order-preserving float <-> uint mapping for atomicMin/Max on floats
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void bounds_init_kernel(uint32_t* bounds) {
    if (threadIdx.x < 3) bounds[threadIdx.x] = 0xffffffffu;       // lo (min)
    else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;           // hi (max)
}

// One thread per surfel: quad record, leaf AABB, centroid bounds.
__global__ void __launch_bounds__(256) prim_kernel(int P, const float* __restrict__ vertices, QuadPrim* __restrict__ prims,
                                                   float4* __restrict__ leaf_lo, float4* __restrict__ leaf_hi,
                                                   uint32_t* __restrict__ bounds) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 cmin = f3(FLT_MAX, FLT_MAX, FLT_MAX), cmax = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    if (i < P) {
        const float* v = vertices + 12 * (size_t)i;
        float3 v0 = f3(v[0], v[1], v[2]), v1 = f3(v[3], v[4], v[5]), v2 = f3(v[6], v[7], v[8]), v3 = f3(v[9], v[10], v[11]);
        float3 eu = v3 - v1, ev = v0 - v1, n = cross(eu, ev);
        float lu = dot(eu, eu), lv = dot(ev, ev);
        bool valid = isfinite(v0.x + v0.y + v0.z + v1.x + v1.y + v1.z + v2.x + v2.y + v2.z + v3.x + v3.y + v3.z) &&
                     lu > 0.0f && lv > 0.0f && dot(n, n) > 0.0f;
        QuadPrim q;
        q.v1 = make_float4(v1.x, v1.y, v1.z, valid ? 1.0f / lu : 0.0f);
        q.eu = make_float4(eu.x, eu.y, eu.z, valid ? 1.0f / lv : 0.0f);
        q.ev = make_float4(ev.x, ev.y, ev.z, valid ? 1.0f : 0.0f);
        q.n = make_float4(n.x, n.y, n.z, 0.0f);
        prims[i] = q;
        float3 lo, hi;
        if (valid) {
            lo = f3(fminf(fminf(v0.x, v1.x), fminf(v2.x, v3.x)), fminf(fminf(v0.y, v1.y), fminf(v2.y, v3.y)),
                    fminf(fminf(v0.z, v1.z), fminf(v2.z, v3.z)));
            hi = f3(fmaxf(fmaxf(v0.x, v1.x), fmaxf(v2.x, v3.x)), fmaxf(fmaxf(v0.y, v1.y), fmaxf(v2.y, v3.y)),
                    fmaxf(fmaxf(v0.z, v1.z), fmaxf(v2.z, v3.z)));
            // pad so flat quads keep a non-zero slab and rounding in the plane test cannot escape the box
            float3 ext = hi - lo;
            // (slab-test rounding is ~2.4e-7 * |coordinate| per plane; keep an order of magnitude of head-room)
            float pad = 1e-4f * fmaxf(fmaxf(ext.x, ext.y), ext.z) +
                        4e-6f * fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))),
                                      fmaxf(fabsf(lo.z), fabsf(hi.z)));
            lo = lo - f3(pad, pad, pad);
            hi = hi + f3(pad, pad, pad);
            float3 c = 0.25f * (v0 + v1 + v2 + v3);
            cmin = c; cmax = c;
        } else {   // never hit
            lo = f3(FLT_MAX, FLT_MAX, FLT_MAX);
            hi = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
        }
        leaf_lo[i] = make_float4(lo.x, lo.y, lo.z, valid ? 1.0f : 0.0f);
        leaf_hi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    }
    // block reduction of centroid bounds
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cmin.x = fminf(cmin.x, __shfl_xor_sync(0xffffffffu, cmin.x, o));
        cmin.y = fminf(cmin.y, __shfl_xor_sync(0xffffffffu, cmin.y, o));
        cmin.z = fminf(cmin.z, __shfl_xor_sync(0xffffffffu, cmin.z, o));
        cmax.x = fmaxf(cmax.x, __shfl_xor_sync(0xffffffffu, cmax.x, o));
        cmax.y = fmaxf(cmax.y, __shfl_xor_sync(0xffffffffu, cmax.y, o));
        cmax.z = fmaxf(cmax.z, __shfl_xor_sync(0xffffffffu, cmax.z, o));
    }
    __shared__ float s_b[8][6];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_b[warp][0] = cmin.x; s_b[warp][1] = cmin.y; s_b[warp][2] = cmin.z; s_b[warp][3] = cmax.x; s_b[warp][4] = cmax.y; s_b[warp][5] = cmax.z; }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s_b[0][threadIdx.x];
        for (int w = 1; w < 8; w++) v = threadIdx.x < 3 ? fminf(v, s_b[w][threadIdx.x]) : fmaxf(v, s_b[w][threadIdx.x]);
        if (threadIdx.x < 3) { if (v < FLT_MAX) atomicMin(&bounds[threadIdx.x], f2ord(v)); }
        else { if (v > -FLT_MAX) atomicMax(&bounds[threadIdx.x], f2ord(v)); }
    }
}

__device__ __forceinline__ uint64_t spread21(uint64_t x) {   // insert two zero bits between each of the low 21 bits
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void __launch_bounds__(256) morton_kernel(int P, const float* __restrict__ vertices,
                                                     const float4* __restrict__ leaf_lo, const uint32_t* __restrict__ bounds,
                                                     uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float3 lo = f3(ord2f(bounds[0]), ord2f(bounds[1]), ord2f(bounds[2]));
    float3 hi = f3(ord2f(bounds[3]), ord2f(bounds[4]), ord2f(bounds[5]));
    uint64_t key = ~0ull >> 1;   // invalid surfels sort last
    if (leaf_lo[i].w != 0.0f) {
        const float* v = vertices + 12 * (size_t)i;
        float3 c = 0.25f * (f3(v[0], v[1], v[2]) + f3(v[3], v[4], v[5]) + f3(v[6], v[7], v[8]) + f3(v[9], v[10], v[11]));
        float3 ext = hi - lo;
        float sx = ext.x > 0 ? (c.x - lo.x) / ext.x : 0.0f, sy = ext.y > 0 ? (c.y - lo.y) / ext.y : 0.0f,
              sz = ext.z > 0 ? (c.z - lo.z) / ext.z : 0.0f;
        const float S = 2097151.0f;   // 2^21 - 1
        uint64_t ix = (uint64_t)fminf(fmaxf(sx * S, 0.0f), S), iy = (uint64_t)fminf(fmaxf(sy * S, 0.0f), S),
                 iz = (uint64_t)fminf(fmaxf(sz * S, 0.0f), S);
        key = (spread21(ix) << 2) | (spread21(iy) << 1) | spread21(iz);
    }
    keys[i] = key;
    vals[i] = (uint32_t)i;
}

// Karras 2012 "Maximizing parallelism in the construction of BVHs": one thread per internal node.
__device__ __forceinline__ int lbvh_delta(const uint64_t* __restrict__ keys, int P, int i, int j) {
    if (j < 0 || j >= P) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void __launch_bounds__(256) topology_kernel(int P, const uint64_t* __restrict__ keys,
                                                       const uint32_t* __restrict__ vals_sorted,
                                                       BvhNode* __restrict__ nodes, int* __restrict__ parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P - 1) return;
    int d = (lbvh_delta(keys, P, i, i + 1) - lbvh_delta(keys, P, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = lbvh_delta(keys, P, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, P, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(keys, P, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = lbvh_delta(keys, P, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(keys, P, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int c0, c1;
    if (lo == gamma) { c0 = ~(int)vals_sorted[gamma]; parent[P - 1 + gamma] = i; }
    else { c0 = gamma; parent[gamma] = i; }
    if (hi == gamma + 1) { c1 = ~(int)vals_sorted[gamma + 1]; parent[P - 1 + gamma + 1] = i | 0x80000000; }
    else { c1 = gamma + 1; parent[gamma + 1] = i | 0x80000000; }
    nodes[i].child0 = c0;
    nodes[i].child1 = c1;
    if (i == 0) parent[0] = -1;
}

// Bottom-up AABB propagation: the second thread to reach a node merges both child boxes and moves up.
__global__ void __launch_bounds__(256) refit_kernel(int P, const uint32_t* __restrict__ vals_sorted,
                                                    const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi,
                                                    BvhNode* nodes, const int* __restrict__ parent, int* flags) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    uint32_t sid = vals_sorted[k];
    float4 l4 = leaf_lo[sid], h4 = leaf_hi[sid];
    float3 lo = f3(l4.x, l4.y, l4.z), hi = f3(h4.x, h4.y, h4.z);
    int link = parent[P - 1 + k];
    while (true) {
        int n = link & 0x7fffffff;
        int which = (link >> 31) & 1;
        volatile float* nf = reinterpret_cast<volatile float*>(&nodes[n]);
        int base = which * 6;
        nf[base + 0] = lo.x; nf[base + 1] = lo.y; nf[base + 2] = lo.z;
        nf[base + 3] = hi.x; nf[base + 4] = hi.y; nf[base + 5] = hi.z;
        __threadfence();
        int old = atomicAdd(&flags[n], 1);
        if (old == 0) return;   // sibling not there yet
        __threadfence();
        int ob = (1 - which) * 6;
        lo = f3(fminf(lo.x, nf[ob + 0]), fminf(lo.y, nf[ob + 1]), fminf(lo.z, nf[ob + 2]));
        hi = f3(fmaxf(hi.x, nf[ob + 3]), fmaxf(hi.y, nf[ob + 4]), fmaxf(hi.z, nf[ob + 5]));
        if (n == 0) return;
        link = parent[n];
    }
}

__global__ void single_prim_kernel(const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi, BvhNode* nodes) {
    float4 l = leaf_lo[0], h = leaf_hi[0];
    float* nf = reinterpret_cast<float*>(&nodes[0]);
    nf[0] = l.x; nf[1] = l.y; nf[2] = l.z; nf[3] = h.x; nf[4] = h.y; nf[5] = h.z;
    nf[6] = FLT_MAX; nf[7] = FLT_MAX; nf[8] = FLT_MAX; nf[9] = -FLT_MAX; nf[10] = -FLT_MAX; nf[11] = -FLT_MAX;
    nodes[0].child0 = ~0;
    nodes[0].child1 = ~0;
}

// Collapse the binary tree to 4-wide nodes: node i adopts its grandchildren (a leaf child is adopted as is).
__global__ void __launch_bounds__(256) collapse4_kernel(int n_internal, const BvhNode* __restrict__ nodes, Bvh4Node* __restrict__ nodes4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_internal) return;
    const float* nf = reinterpret_cast<const float*>(&nodes[i]);
    const int c[2] = {nodes[i].child0, nodes[i].child1};
    Bvh4Node out;
    int k = 0;
    int leafmask = 0;   // which of the four slots hold surfels (replicated in every slot's b.w)
#pragma unroll
    for (int s = 0; s < 2; s++) {
        if (c[s] < 0) {
            const float* b = nf + 6 * s;
            out.s[k].a = make_float4(b[0], b[1], b[2], b[3]);
            out.s[k].b = make_float4(b[4], b[5], __int_as_float(c[s]), 0.0f);
            leafmask |= 1 << k;
            k++;
        } else {
            const float* cf = reinterpret_cast<const float*>(&nodes[c[s]]);
            const int g[2] = {nodes[c[s]].child0, nodes[c[s]].child1};
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const float* b = cf + 6 * q;
                out.s[k].a = make_float4(b[0], b[1], b[2], b[3]);
                out.s[k].b = make_float4(b[4], b[5], __int_as_float(g[q]), 0.0f);
                if (g[q] < 0) leafmask |= 1 << k;
                k++;
            }
        }
    }
    for (; k < 4; k++) {
        out.s[k].a = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX);
        out.s[k].b = make_float4(-FLT_MAX, -FLT_MAX, __int_as_float(~0), 0.0f);
        leafmask |= 1 << k;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) out.s[q].b.w = __int_as_float(leafmask);
    nodes4[i] = out;
}


// ------------------------------------------------------------------------------------------------------
// PLOC: parallel locally-ordered clustering (Meister & Bittner 2018) on the Morton-sorted leaves
// ------------------------------------------------------------------------------------------------------
// The Karras LBVH splits purely by Morton prefix; its sibling boxes overlap and a ray visits ~2x the nodes an
// SAH-quality tree needs (round-1 verdict).  PLOC builds bottom-up instead: every cluster looks at its 2r neighbours in
// Morton order, picks the one whose union box has the smallest surface area, and mutually-nearest pairs merge; repeat until
// one cluster is left.  ~35 iterations of 4 small kernels for 256 K surfels, which is why lbvh_build re-uses the
// topology for several calls and only refits in between (boxes always come from the current quads, so any valid topology
// gives exactly the same hits).
constexpr int PLOC_RADIUS = 8;

__device__ __forceinline__ float box_area(float3 lo, float3 hi) {
    const float3 e = hi - lo;
    return e.x * e.y + e.y * e.z + e.z * e.x;
}

__global__ void __launch_bounds__(256) ploc_init_kernel(int P, const uint32_t* __restrict__ vals_sorted, const float4* __restrict__ leaf_lo,
                                                        const float4* __restrict__ leaf_hi, float4* __restrict__ lo, float4* __restrict__ hi,
                                                        int* __restrict__ link, int* __restrict__ slot, int* __restrict__ counter) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) { counter[0] = 0; counter[1] = P; counter[2] = P; }
    if (k >= P) return;
    const uint32_t sid = vals_sorted[k];
    lo[k] = leaf_lo[sid];
    hi[k] = leaf_hi[sid];
    link[k] = ~(int)sid;
    slot[k] = P - 1 + k;
}

// The cluster count of an iteration lives on the device (`n_ptr`, ping-pong between counter[1] and counter[2]); the host only knows an
// upper bound (the count it last read back) and sizes the grids with it, so several iterations run between two read-backs.
__global__ void __launch_bounds__(256) ploc_nn_kernel(const int* __restrict__ n_ptr, const float4* __restrict__ lo, const float4* __restrict__ hi,
                                                      int* __restrict__ nn) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *n_ptr;
    if (i >= n) return;
    const float4 l4 = lo[i], h4 = hi[i];
    const float3 l = f3(l4.x, l4.y, l4.z), h = f3(h4.x, h4.y, h4.z);
    float best = FLT_MAX;
    int bj = -1;
    const int j0 = max(0, i - PLOC_RADIUS), j1 = min(n - 1, i + PLOC_RADIUS);
    for (int j = j0; j <= j1; j++) {
        if (j == i) continue;
        const float4 a = lo[j], b = hi[j];
        const float ar = box_area(f3(fminf(l.x, a.x), fminf(l.y, a.y), fminf(l.z, a.z)), f3(fmaxf(h.x, b.x), fmaxf(h.y, b.y), fmaxf(h.z, b.z)));
        if (ar < best || bj < 0) { best = ar; bj = j; }      // invalid (inverted / infinite) boxes still pair up with someone
    }
    nn[i] = bj;
}

__global__ void __launch_bounds__(256) ploc_merge_kernel(const int* __restrict__ n_ptr, int n_bound, int P, const int* __restrict__ nn,
                                                         float4* __restrict__ lo, float4* __restrict__ hi,
                                                         int* __restrict__ link, int* __restrict__ slot, int* __restrict__ valid,
                                                         BvhNode* __restrict__ nodes, int* __restrict__ parent, int* __restrict__ counter) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *n_ptr;
    if (i >= n) {
        if (i < n_bound) valid[i] = 0;      // the scan runs over the host's bound
        return;
    }
    const int j = nn[i];
    if (j < 0 || nn[j] != i) { valid[i] = 1; return; }
    if (i > j) { valid[i] = 0; return; }
    const int id = P - 2 - atomicAdd(&counter[0], 1);      // the last merge of the build gets id 0: the root
    nodes[id].child0 = link[i];
    nodes[id].child1 = link[j];
    parent[slot[i]] = id;
    parent[slot[j]] = id | 0x80000000;
    const float4 la = lo[i], ha = hi[i], lb = lo[j], hb = hi[j];
    lo[i] = make_float4(fminf(la.x, lb.x), fminf(la.y, lb.y), fminf(la.z, lb.z), 1.0f);
    hi[i] = make_float4(fmaxf(ha.x, hb.x), fmaxf(ha.y, hb.y), fmaxf(ha.z, hb.z), 0.0f);
    link[i] = id;
    slot[i] = id;
    valid[i] = 1;
    if (id == 0) parent[0] = -1;
}

__global__ void __launch_bounds__(256) ploc_compact_kernel(const int* __restrict__ n_ptr, const int* __restrict__ valid, const int* __restrict__ scan,
                                                           const float4* __restrict__ lo, const float4* __restrict__ hi,
                                                           const int* __restrict__ link, const int* __restrict__ slot,
                                                           float4* __restrict__ lo2, float4* __restrict__ hi2, int* __restrict__ link2,
                                                           int* __restrict__ slot2, int* __restrict__ n_next) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *n_ptr;
    if (i >= n) return;
    if (valid[i]) {
        const int d = scan[i];
        lo2[d] = lo[i]; hi2[d] = hi[i]; link2[d] = link[i]; slot2[d] = slot[i];
    }
    if (i == n - 1) *n_next = scan[i] + valid[i];
}

static int ploc_topology(Lbvh& t, int P, cudaStream_t stream) {
    ploc_init_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, t.vals_sorted, (const float4*)t.leaf_lo, (const float4*)t.leaf_hi, t.cl_lo[0],
                                                         t.cl_hi[0], t.cl_link[0], t.cl_slot[0], t.cl_counter);
    ENVGS_LAUNCH_CHECK("ploc_init", false, stream);
    // Iterations run in batches of PLOC_BATCH between two 4-byte read-backs of the cluster count: the kernels take the count from the
    // device, the grids are sized with the count at the start of the batch.  A single cluster is a fixed point (no neighbour, no
    // merge), so the surplus iterations of the last batch change nothing.  This build runs once every few calls, not in the per-step path.
    constexpr int PLOC_BATCH = 4;
    int n = P, cur = 0, iter = 0;
    while (n > 1 && iter < 4096) {
        const int blocks = (n + 255) / 256;
        for (int b = 0; b < PLOC_BATCH; b++, iter++) {
            const int* n_ptr = t.cl_counter + 1 + (iter & 1);
            int* n_next = t.cl_counter + 1 + ((iter + 1) & 1);
            ploc_nn_kernel<<<blocks, 256, 0, stream>>>(n_ptr, t.cl_lo[cur], t.cl_hi[cur], t.cl_nn);
            ploc_merge_kernel<<<blocks, 256, 0, stream>>>(n_ptr, n, P, t.cl_nn, t.cl_lo[cur], t.cl_hi[cur], t.cl_link[cur], t.cl_slot[cur], t.cl_valid,
                                                         t.nodes, t.parent, t.cl_counter);
            ENVGS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(t.scan_temp, t.scan_temp_bytes, t.cl_valid, t.cl_scan, n, stream));
            ploc_compact_kernel<<<blocks, 256, 0, stream>>>(n_ptr, t.cl_valid, t.cl_scan, t.cl_lo[cur], t.cl_hi[cur], t.cl_link[cur], t.cl_slot[cur],
                                                           t.cl_lo[cur ^ 1], t.cl_hi[cur ^ 1], t.cl_link[cur ^ 1], t.cl_slot[cur ^ 1], n_next);
            count_launch(5);
            cur ^= 1;
        }
        int alive = 0;
        ENVGS_CUDA_TRY(cudaMemcpyAsync(&alive, t.cl_counter + 1 + (iter & 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
        ENVGS_CUDA_TRY(cudaStreamSynchronize(stream));
        if (alive >= n || alive < 1) { set_error("PLOC made no progress (%d -> %d clusters)", n, alive); return ENVGS_ERR_STATE; }
        n = alive;
    }
    if (n != 1) { set_error("PLOC did not converge"); return ENVGS_ERR_STATE; }
    return ENVGS_OK;
}

int lbvh_build(Lbvh& t, const float* vertices, int P, bool rebuild, cudaStream_t stream) {
    if (P <= 0) { t.P = 0; return ENVGS_OK; }
    ENVGS_TIMED("bvh_build", stream);
    if (!rebuild && t.P != P) rebuild = true;   // topology can only be reused for the same surfel count
    uint32_t* bounds = reinterpret_cast<uint32_t*>(t.bounds);
    int blocks = (P + 255) / 256;
    bounds_init_kernel<<<1, 32, 0, stream>>>(bounds);
    ENVGS_LAUNCH_CHECK("bvh_bounds_init", false, stream);
    prim_kernel<<<blocks, 256, 0, stream>>>(P, vertices, t.prims, (float4*)t.leaf_lo, (float4*)t.leaf_hi, bounds);
    ENVGS_LAUNCH_CHECK("bvh_prims", false, stream);
    if (P == 1) {
        single_prim_kernel<<<1, 1, 0, stream>>>((const float4*)t.leaf_lo, (const float4*)t.leaf_hi, t.nodes);
        ENVGS_LAUNCH_CHECK("bvh_single", false, stream);
        collapse4_kernel<<<1, 32, 0, stream>>>(1, t.nodes, t.nodes4);
        ENVGS_LAUNCH_CHECK("bvh_collapse4", false, stream);
        t.P = 1;
        return ENVGS_OK;
    }
    // Topology policy.  The caller asks for a rebuild on every call (EnvGS rebuilds its GAS every iteration); the boxes below are
    // always recomputed from the current quads, so re-using the previous TOPOLOGY is exact — it can only cost traversal
    // efficiency once the surfels have moved far.  A full build (PLOC, ~1 ms) therefore runs every ENVGS_BVH_REBUILD_EVERY-th
    // call (default 16; 1 = every call) and whenever the surfel count changes; ENVGS_BVH_BUILDER=lbvh selects the one-pass Karras
    // topology instead.
    static int every = 0, use_ploc = -1;
    if (every == 0) { const char* e = getenv("ENVGS_BVH_REBUILD_EVERY"); every = e ? atoi(e) : 16; if (every < 1) every = 1; }
    if (use_ploc < 0) { const char* e = getenv("ENVGS_BVH_BUILDER"); use_ploc = (e && strcmp(e, "lbvh") == 0) ? 0 : 1; }
    if (rebuild && t.P == P && t.calls_since_full < every - 1) { rebuild = false; t.calls_since_full++; }
    else if (rebuild) { t.calls_since_full = 0; t.full_builds++; }
    if (rebuild) {
        morton_kernel<<<blocks, 256, 0, stream>>>(P, vertices, (const float4*)t.leaf_lo, bounds, t.keys, t.vals);
        ENVGS_LAUNCH_CHECK("bvh_morton", false, stream);
        ENVGS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(t.sort_temp, t.sort_temp_bytes, t.keys, t.keys_sorted, t.vals,
                                                       t.vals_sorted, P, 0, 63, stream));
        count_launch(10);
        if (use_ploc) {
            int rc = ploc_topology(t, P, stream);
            if (rc != ENVGS_OK) return rc;
        } else {
            topology_kernel<<<(P - 1 + 255) / 256, 256, 0, stream>>>(P, t.keys_sorted, t.vals_sorted, t.nodes, t.parent);
            ENVGS_LAUNCH_CHECK("bvh_topology", false, stream);
        }
    }
    ENVGS_CUDA_TRY(cudaMemsetAsync(t.flags, 0, sizeof(int) * (size_t)P, stream));
    refit_kernel<<<blocks, 256, 0, stream>>>(P, t.vals_sorted, (const float4*)t.leaf_lo, (const float4*)t.leaf_hi, t.nodes,
                                             t.parent, t.flags);
    ENVGS_LAUNCH_CHECK("bvh_refit", false, stream);
    collapse4_kernel<<<(P - 1 + 255) / 256, 256, 0, stream>>>(P - 1, t.nodes, t.nodes4);
    ENVGS_LAUNCH_CHECK("bvh_collapse4", false, stream);
    t.P = P;
    return ENVGS_OK;
}

}  // namespace envgs
