// trace_fast.cu — the tracer's fast path for EnvGS' own call shape (max_trace_depth = 0, start_from_first = false):
//
//   traverse_kernel   four lanes per ray on the 4-wide BVH (one child slot per lane), persistent warps with group-level
//                     refill, two phases per round (node visits with queued leaves, then quad tests four at a time), the
//                     per-ray stack and leaf queue in shared memory.  Finds ALL quad hits of the ray; a ray with more
//                     than 15 hits is handed to the general kernel of trace.cu (which implements the reference's 16-hit
//                     chunk restart exactly).  Hits are appended unsorted to hits[ray][16] (t, surfel) = one 128-byte line.
//   class_*_kernel    stable compaction of the rays into four lists by hit count (0 / 1-4 / 5-8 / 9-15), in pixel order.
//   shade_fwd_kernel  HIT-parallel: 4, 8 or 16 lanes per ray (by hit count), one hit per lane.  Per-hit evaluation (surfel
//                     frame, alpha, SH colour) runs in parallel; the front-to-back recurrence T_i = T_{i-1}(1-a_i) is a
//                     log2(L)-step segmented shuffle scan; every accumulated output is one transpose-reduce.
//   shade_bwd_kernel  same layout; the reference's suffix terms (F - F_i)/(1-a_i) come from inclusive scans.  It re-uses
//                     the forward's hit lists, so the backward pass does no BVH traversal at all.
//
// Semantics are those of TRACE/optix_tracer/forward.cu:219-450 and backward.cu:508-887 for a ray whose first 16-hit
// chunk is not full (payload.cnt < CHUNK_SIZE => the while loop ends after one optixTrace, forward.cu:434).
#include "lbvh.cuh"

namespace envgs {

constexpr int TF_THREADS = 128;
constexpr int TF_STACK = 32;
constexpr int TF_LSTACK = 64;
constexpr int T4_STACK = 48;                // per-ray traversal stack (shared by the ray's 4 lanes)
constexpr int T4_LEAFQ = 32;                // queued leaves per ray between two quad-test phases
constexpr int TF_MAXHITS = 16;               // list capacity; a usable list has at most 15 entries
constexpr int STACK_EMPTY = (int)0x80000000;

struct FastParams {
    const BvhNode* nodes;
    const Bvh4Node* nodes4;
    const QuadPrim* prims;
    envgs_trace_args a;
    envgs_trace_out o;
    envgs_trace_grads g;
    float2* hits;                 // [R][16] (t, surfel id as int bits)
    int* counts;                  // [R] number of hits, -1 = overflow (handled by the fallback kernel)
    unsigned int* overflow_count; // [1]
    int* overflow_list;           // [R]
    unsigned int* class_count;    // [4] rays with 0 / 1-4 / 5-8 / 9-15 hits
    unsigned int* block_counts;   // scratch of the class-list kernels
    int* class_list;              // [4][R]
    int R;
    unsigned int* work_counter;
    unsigned long long* stats;
    int n_items, tiles_x;
    int refill_threshold;         // leave an inner traversal loop when fewer lanes than this are still in it
    int mid_half;                 // 1: one segment of a two-segment trace: mid_val is pre-offset to this segment's 16 floats, the
                                  // other half of the row is left alone
};

constexpr int CL_THREADS = 512;               // rays per block of the class-list kernels
static inline size_t cl_blocks(size_t R) { return (R + CL_THREADS - 1) / CL_THREADS; }

size_t trace_cache_bytes(int H, int W) {
    size_t R = (size_t)H * W;
    return R * (TF_MAXHITS * sizeof(float2) + 4 + 4 + 16) + 1024 + 4 * cl_blocks(R) * sizeof(unsigned int);
}
static void carve_cache(char* base, int H, int W, FastParams& p) {
    size_t R = (size_t)H * W;
    p.R = (int)R;
    p.hits = reinterpret_cast<float2*>(base);
    p.counts = reinterpret_cast<int*>(base + R * TF_MAXHITS * sizeof(float2));
    p.overflow_list = p.counts + R;
    p.class_list = p.overflow_list + R;
    p.overflow_count = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(p.class_list + 4 * R) + 256);
    p.class_count = p.overflow_count + 4;
    p.block_counts = p.overflow_count + 64;     // [4][cl_blocks(R)] per-block class counts -> offsets
}

__device__ __forceinline__ float safe_inv_f(float d) { return 1.0f / (fabsf(d) > 1e-30f ? d : copysignf(1e-30f, d)); }

// ------------------------------------------------------------------------------------------------------
// traversal
// ------------------------------------------------------------------------------------------------------
// Persistent warps with GROUP-level work refill: a 4-lane group whose ray has finished takes the next ray from a
// global cursor (warp-aggregated atomic), so a warp never idles behind its longest ray — reflection rays off a noisy
// normal map are incoherent and their traversal lengths differ by an order of magnitude.  Ray k maps to pixel
// (8x4 block k/32, position k%32), so rays fetched together are still neighbours on screen.  Hits are appended
// unsorted to the ray's own 128-byte line in HBM/L2 (the shading kernels sort the 16 slots with a bitonic network).
__device__ __forceinline__ bool ray_of_index(const FastParams& p, unsigned int k, int& row, int& col) {
    const int H = p.a.H, W = p.a.W;
    if (p.tiles_x > 0) {
        const unsigned int item = k >> 5, l = k & 31;
        const int by = item / p.tiles_x, bx = item % p.tiles_x;
        row = by * 4 + (int)(l >> 3); col = bx * 8 + (int)(l & 7);
    } else {
        row = (int)(k / (unsigned)W); col = (int)(k % (unsigned)W);
    }
    return row < H && col < W;
}

// FOUR lanes per ray on a 4-wide BVH.  With one ray per lane every node visit is 32 scattered 16-byte loads per
// instruction = 32 L1 wavefronts, and the kernel was bound by the L1 wavefront rate (ncu: L1 hit 49 %, issue 59 %).
// Here lane j of a group loads child slot j (32 B) of the 128-byte node, so a visit costs the group ONE cache line
// (two wavefronts for 8 rays per instruction), each lane slab-tests one child, and the hit masks are combined with a
// group-wide ballot.  Reached leaves are queued and quad-tested four at a time (one per lane).
// Shared memory is addressed through explicit 32-bit shared-window addresses (st.shared / ld.shared): with ordinary
// array indexing ptxas re-derived the window base (S2UR CgaCtaId + ULEA) and the group offset (S2R tid) at every access,
// ~10 % of the inner loop's instructions.  Per group: [0, T4_STACK) traversal stack, [T4_STACK, T4_GROUP) leaf queue.
constexpr int T4_GROUP = T4_STACK + T4_LEAFQ;
__device__ __forceinline__ void sts32(unsigned int addr, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ int lds32(unsigned int addr) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

#ifndef T4_MINBLOCKS
#define T4_MINBLOCKS 9
#endif
__device__ __forceinline__ const Bvh4Slot* node_slot(const char* slot0, int node, bool active) {
    unsigned long long addr;
    asm("{\n\t.reg .pred p;\n\t.reg .u32 n;\n\tsetp.ne.u32 p, %2, 0;\n\tselp.u32 n, %1, 0, p;\n\tmad.wide.u32 %0, n, 128, %3;\n\t}"
        : "=l"(addr) : "r"(node), "r"((unsigned int)active), "l"(reinterpret_cast<unsigned long long>(slot0)));
    return reinterpret_cast<const Bvh4Slot*>(addr);
}

template <bool STATS>
__global__ void __launch_bounds__(TF_THREADS, T4_MINBLOCKS) traverse_kernel(const FastParams p) {
    __shared__ int s_buf[(TF_THREADS / 4) * T4_GROUP];
    const int lane = threadIdx.x & 31;
    const int sub = lane & 3;
    const int gshift = lane & ~3;
    const unsigned int lower = (1u << sub) - 1u;
    unsigned int sbase;   // byte address of this group's slice in the shared window, pinned in a register
    asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"((unsigned int)__cvta_generic_to_shared(s_buf) + (threadIdx.x >> 2) * (T4_GROUP * 4)));
    const int W = p.a.W;
    const char* __restrict__ slot0 = reinterpret_cast<const char*>(p.nodes4) + 32 * sub;   // this lane's child slot of node 0
    const QuadPrim* __restrict__ prims = p.prims;
    const unsigned int total = (unsigned int)p.n_items * 32u;

    // per-ray state, replicated in the four lanes of the group
    int node = STACK_EMPTY, nodeB = STACK_EMPTY;   // the (up to) two nodes of the next round; nodeB is only set when node is
    bool have_ray = false, exhausted = false, overflow = false;
    size_t ray = 0;
    float3 o = f3(0, 0, 0), d = f3(0, 0, 1), inv = f3(0, 0, 0), oi = f3(0, 0, 0);
    int cnt = 0, kept = 0, sp = 0, nleaf = 0;   // quad hits found / hits stored (inside their alpha >= 1/255 disc)
    unsigned int st_nodes = 0, st_prims = 0;

    while (true) {
        // ---- retire finished rays, refill idle groups (all 32 lanes are converged here) ----
        const bool idle = (node == STACK_EMPTY) && nleaf == 0;
        {
            const bool retire = idle && have_ray;
            if (retire && sub == 0) {
                if (overflow) {
                    p.counts[ray] = -1;
                    const unsigned int k = atomicAdd(p.overflow_count, 1u);
                    p.overflow_list[k] = (int)ray;
                } else {
                    p.counts[ray] = kept;
                }
                if (STATS) {
                    atomicAdd(&p.stats[0], 1ull); atomicAdd(&p.stats[1], (unsigned long long)st_nodes);
                    atomicAdd(&p.stats[2], (unsigned long long)st_prims); atomicAdd(&p.stats[3], (unsigned long long)kept);
                    atomicAdd(&p.stats[5], (unsigned long long)cnt);
                    atomicAdd(&p.stats[4], 1ull);
                }
            }
            if (retire) have_ray = false;
        }
        const bool gwant = idle && !exhausted;
        const unsigned int want = __ballot_sync(0xffffffffu, gwant && sub == 0);
        if (want) {
            const int leader = __ffs(want) - 1;
            unsigned int base = 0;
            if (lane == leader) base = atomicAdd(p.work_counter, (unsigned int)__popc(want));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (gwant) {
                const unsigned int k = base + __popc(want & ((1u << gshift) - 1u));
                int row = 0, col = 0;
                bool ok = false;
                if (k < total) ok = ray_of_index(p, k, row, col);   // padding slots of edge tiles are skipped
                else exhausted = true;
                if (ok) {
                    ray = (size_t)row * W + col;
                    o = ld3(p.a.ray_o, (int)ray); d = ld3(p.a.ray_d, (int)ray);
                    inv = f3(safe_inv_f(d.x), safe_inv_f(d.y), safe_inv_f(d.z));
                    oi = o * inv;
                    cnt = 0; kept = 0; sp = 0; nleaf = 0; overflow = false; st_nodes = 0; st_prims = 0;
                    have_ray = true;
                    node = isfinite(d.x + d.y + d.z + o.x + o.y + o.z) ? 0 : STACK_EMPTY;
                    nodeB = STACK_EMPTY;
                }
            }
        }
        if (__all_sync(0xffffffffu, exhausted && !have_ray)) break;

        // ---- phase 1: 4-wide internal nodes; hit leaves are queued, hit internal children pushed ----
        // The warp stays converged (full-mask votes, idle groups predicated off): per-group vote masks would split the
        // warp into eight 4-lane fragments that are issued one after the other.
        const bool can_refill = __any_sync(0xffffffffu, idle && !exhausted);
        // TWO nodes per group per round (the two topmost stack entries): their loads are independent, which doubles the
        // memory-level parallelism of the pointer chase and amortises the round's votes and loop control.
        while (true) {
            const bool act = node != STACK_EMPTY && nleaf <= T4_LEAFQ - 8 && sp <= T4_STACK - 8;
            const unsigned int am = __ballot_sync(0xffffffffu, act);
            if (am == 0u) break;
            if (can_refill && __popc(am) < p.refill_threshold) break;   // let idle groups fetch new rays (Aila & Laine 2009)
            const bool actB = act && nodeB != STACK_EMPTY;
            // idle groups re-test node 0; select-then-widen (selp + mad.wide) instead of the 64-bit selects ptxas derives itself
            const Bvh4Slot* __restrict__ slA = node_slot(slot0, node, act);
            const Bvh4Slot* __restrict__ slB = node_slot(slot0, nodeB, actB);
            const float4 A0 = __ldg(&slA->a), A1 = __ldg(&slA->b), B0 = __ldg(&slB->a), B1 = __ldg(&slB->b);
            if (STATS) st_nodes += (act ? 1u : 0u) + (actB ? 1u : 0u);
            bool hitA, hitB;
            {
                const float ax = fmaf(A0.x, inv.x, -oi.x), bx = fmaf(A0.w, inv.x, -oi.x);
                const float ay = fmaf(A0.y, inv.y, -oi.y), by = fmaf(A1.x, inv.y, -oi.y);
                const float az = fmaf(A0.z, inv.z, -oi.z), bz = fmaf(A1.y, inv.z, -oi.z);
                const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
                const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) * 1.0000008f;
                hitA = act && (tn <= tf) && (tf > 0.0f);
            }
            {
                const float ax = fmaf(B0.x, inv.x, -oi.x), bx = fmaf(B0.w, inv.x, -oi.x);
                const float ay = fmaf(B0.y, inv.y, -oi.y), by = fmaf(B1.x, inv.y, -oi.y);
                const float az = fmaf(B0.z, inv.z, -oi.z), bz = fmaf(B1.y, inv.z, -oi.z);
                const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
                const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) * 1.0000008f;
                hitB = actB && (tn <= tf) && (tf > 0.0f);
            }
            const unsigned int lmA = (unsigned int)__float_as_int(A1.w), lmB = (unsigned int)__float_as_int(B1.w);
            const unsigned int mA = (__ballot_sync(0xffffffffu, hitA) >> gshift) & 0xFu;
            const unsigned int mB = (__ballot_sync(0xffffffffu, hitB) >> gshift) & 0xFu;
            const unsigned int mlA = mA & lmA, miA = mA & ~lmA, mlB = mB & lmB, miB = mB & ~lmB;
            const int nlA = __popc(mlA), niA = __popc(miA);
            const bool leafA = (lmA >> sub) & 1u, leafB = (lmB >> sub) & 1u;
            const int atA = (leafA ? T4_STACK + nleaf : sp) + __popc((leafA ? mlA : miA) & lower);
            const int atB = (leafB ? T4_STACK + nleaf + nlA : sp + niA) + __popc((leafB ? mlB : miB) & lower);
            const unsigned int adA = sbase + 4u * (unsigned int)atA, adB = sbase + 4u * (unsigned int)atB;
            if (hitA) sts32(adA, __float_as_int(A1.z));
            if (hitB) sts32(adB, __float_as_int(B1.z));
            __syncwarp();
            nleaf += nlA + __popc(mlB);          // all masks are 0 for idle groups
            sp += niA + __popc(miB);
            const bool popA = act && sp > 0, popB = act && sp > 1;
            const int tA = lds32(sbase + 4u * (unsigned int)max(sp - 1, 0));
            const int tB = lds32(sbase + 4u * (unsigned int)max(sp - 2, 0));
            node = popA ? tA : (act ? STACK_EMPTY : node);
            nodeB = popB ? tB : (act ? STACK_EMPTY : nodeB);
            sp -= (popA ? 1 : 0) + (popB ? 1 : 0);
            __syncwarp();
        }
        if (sp > T4_STACK - 8 && node != STACK_EMPTY && nleaf <= T4_LEAFQ - 8) {
            // stack capacity reached (pathologically deep tree): hand the ray to the general kernel
            overflow = true; node = STACK_EMPTY; nodeB = STACK_EMPTY; sp = 0; nleaf = 0;
        }
        // ---- phase 2: queued quads, four at a time (warp-uniform trip count) ----
        for (int i0 = 0; __any_sync(0xffffffffu, i0 < nleaf); i0 += 4) {
            // branch-free quad test: lanes without a queued leaf test surfel 0 with the result masked off; a ray parallel
            // to the plane gives t = inf/nan, which fails the range test exactly like the reference's q != 0 guard
            const int i = i0 + sub;
            const bool valid = i < nleaf;
            const int qv = lds32(sbase + 4u * (unsigned int)(T4_STACK + min(i, T4_LEAFQ - 1)));
            const int sid = valid ? ~qv : 0;
            if (STATS) st_prims += valid ? 1u : 0u;
            const QuadPrim* __restrict__ pr = prims + sid;
            const float4 pn = __ldg(&pr->n), pv = __ldg(&pr->v1), pu = __ldg(&pr->eu), pw = __ldg(&pr->ev);
            const float3 n = f3(pn.x, pn.y, pn.z), v1 = f3(pv.x, pv.y, pv.z);
            const float q = dot(n, d);
            const float t = dot(n, v1 - o) / q;
            const float3 r = (o + t * d) - v1;
            const float a = dot(r, f3(pu.x, pu.y, pu.z)) * pv.w;
            const float b = dot(r, f3(pw.x, pw.y, pw.z)) * pu.w;
            const bool hit = valid && q != 0.0f && t > 0.0f && t < TR_DEPTH_INF && pw.w != 0.0f && a >= 0.0f && a <= 1.0f &&
                             b >= 0.0f && b <= 1.0f;
            // A quad hit only ever contributes where alpha = o exp(-(u^2+v^2)/2) >= 1/255 (forward.cu:380-382), i.e. inside the
            // disc u^2 + v^2 <= 2 ln(255 o) of the +-3 sigma square ((u, v) = (6a - 3, 6b - 3)).  Hits outside it (35-50 % of the
            // square for typical opacities) are COUNTED — a ray with >= 16 quad hits must still go through the reference's 16-hit
            // chunk loop — but not stored, so the shading kernels sort and evaluate only hits that can matter.  The margin (0.2 % +
            // 0.01 in rho) is far above the difference between this (a, b) and the shading kernels' exact (u, v): every borderline
            // hit is kept and decided there.  Comparisons with NaN (o <= 0) are false: such a surfel never contributes.
            const float opa = __ldg(p.a.opacities + sid);
            const float ua = fmaf(a, 6.0f, -3.0f), ub = fmaf(b, 6.0f, -3.0f);
            const bool keep = hit && (ua * ua + ub * ub <= 2.0f * __logf(255.0f * opa) * 1.002f + 0.01f);
            const unsigned int hm = (__ballot_sync(0xffffffffu, hit) >> gshift) & 0xFu;
            const unsigned int km = (__ballot_sync(0xffffffffu, keep) >> gshift) & 0xFu;
            if (keep) {
                const int slot = kept + __popc(km & lower);
                if (slot < TF_MAXHITS) __stcs(p.hits + ray * TF_MAXHITS + slot, make_float2(t, __int_as_float(sid)));
            }
            cnt += __popc(hm);
            kept += __popc(km);
            if (cnt >= TF_MAXHITS && !overflow) { overflow = true; sp = 0; node = STACK_EMPTY; nodeB = STACK_EMPTY; nleaf = 0; }   // full chunk: general kernel
        }
        nleaf = 0;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------
// class lists in pixel order
// ------------------------------------------------------------------------------------------------------
// The shading kernels take their rays from four lists (0 / 1-4 / 5-8 / 9-15 hits).  The lists are built AFTER the traversal
// by a stable three-kernel compaction over counts[] in ray-index order (count -> scan -> scatter, 16 MB of traffic), not
// appended in retirement order: consecutive list entries are then neighbouring pixels, so a warp's per-ray loads and stores
// (hit lines, mid_val rows, the AoS outputs, upstream gradients) share sectors, and neighbouring reflection rays fetch
// overlapping sets of surfels.
__device__ __forceinline__ int hit_class(int cnt) { return cnt < 0 ? -1 : (cnt == 0 ? 0 : (cnt <= 4 ? 1 : (cnt <= 8 ? 2 : 3))); }

__global__ void __launch_bounds__(CL_THREADS) class_count_kernel(const FastParams p, int nblocks) {
    __shared__ unsigned int s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int r = blockIdx.x * CL_THREADS + threadIdx.x;
    const int cls = r < p.R ? hit_class(p.counts[r]) : -1;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const unsigned int m = __ballot_sync(0xffffffffu, cls == c);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt[c], (unsigned int)__popc(m));
    }
    __syncthreads();
    if (threadIdx.x < 4) p.block_counts[(size_t)threadIdx.x * nblocks + blockIdx.x] = s_cnt[threadIdx.x];
}

__global__ void __launch_bounds__(1024) class_scan_kernel(const FastParams p, int nblocks) {   // one block; counts -> exclusive offsets
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c < 4; c++) {
        unsigned int* row = p.block_counts + (size_t)c * nblocks;
        if (threadIdx.x == 0) s_carry = 0;
        __syncthreads();
        for (int base = 0; base < nblocks; base += 1024) {
            const int i = base + threadIdx.x;
            const unsigned int v = i < nblocks ? row[i] : 0u;
            unsigned int x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) s_warp[warp] = x;
            __syncthreads();
            if (warp == 0) {
                unsigned int w = s_warp[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
                s_warp[lane] = w;
            }
            __syncthreads();
            const unsigned int incl = x + (warp > 0 ? s_warp[warp - 1] : 0u) + s_carry;
            if (i < nblocks) row[i] = incl - v;
            __syncthreads();
            if (threadIdx.x == 1023) s_carry = incl;
            __syncthreads();
        }
        if (threadIdx.x == 0) p.class_count[c] = s_carry;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(CL_THREADS) class_scatter_kernel(const FastParams p, int nblocks) {
    __shared__ unsigned int s_warp[4][CL_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * CL_THREADS + threadIdx.x;
    const int cls = r < p.R ? hit_class(p.counts[r]) : -1;
    unsigned int rank = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const unsigned int m = __ballot_sync(0xffffffffu, cls == c);
        if (cls == c) rank = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_warp[c][warp] = __popc(m);
    }
    __syncthreads();
    if (cls >= 0) {
        unsigned int off = p.block_counts[(size_t)cls * nblocks + blockIdx.x];
        for (int w = 0; w < warp; w++) off += s_warp[cls][w];
        p.class_list[(size_t)cls * p.R + off + rank] = r;
    }
}

static int build_class_lists(const FastParams& p, cudaStream_t stream) {
    const int nblocks = (int)cl_blocks((size_t)p.R);
    class_count_kernel<<<nblocks, CL_THREADS, 0, stream>>>(p, nblocks);
    ENVGS_LAUNCH_CHECK("trace_class_count", false, stream);
    class_scan_kernel<<<1, 1024, 0, stream>>>(p, nblocks);
    ENVGS_LAUNCH_CHECK("trace_class_scan", false, stream);
    class_scatter_kernel<<<nblocks, CL_THREADS, 0, stream>>>(p, nblocks);
    ENVGS_LAUNCH_CHECK("trace_class_scatter", false, stream);
    return ENVGS_OK;
}

// ------------------------------------------------------------------------------------------------------
// hit-parallel shading (L = 4, 8 or 16 lanes per ray, chosen from the ray's hit count)
// ------------------------------------------------------------------------------------------------------
// A ray's hits are evaluated one per lane.  With a fixed 16 lanes per ray only ~30 % of the lanes carried a hit (mean
// 4.9 hits per ray, 93 % of the rays have at most 8), so the rays are binned by hit count and the shading kernels are
// instantiated for 4, 8 and 16 lanes per ray; rays without hits are written by a trivial kernel.
template <int L>
__device__ __forceinline__ float seg_scan_mul(float v, int h) {   // inclusive product over the L lanes of a ray
#pragma unroll
    for (int o = 1; o < L; o <<= 1) {
        float u = __shfl_up_sync(0xffffffffu, v, o, L);
        if (h >= o) v *= u;
    }
    return v;
}
template <int L>
__device__ __forceinline__ float seg_scan_add(float v, int h) {   // inclusive sum over the L lanes of a ray
#pragma unroll
    for (int o = 1; o < L; o <<= 1) {
        float u = __shfl_up_sync(0xffffffffu, v, o, L);
        if (h >= o) v += u;
    }
    return v;
}
// Transpose-reduce of 16 per-lane values across the L lanes of a ray: afterwards lane h holds, in v[0 .. 16/L), the sums of
// the components reduce_base<L>(h) + j.  (L = 16: lane h holds component h.)
template <int L>
__device__ __forceinline__ void reduce16(float (&v)[16], int h) {
    int len = 16;
#pragma unroll
    for (int s = L / 2; s >= 1; s >>= 1) {
        len >>= 1;
        const bool up = (h & s) != 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i < len) {
                const float send = up ? v[i] : v[i + len];
                const float keep = up ? v[i + len] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s, L);
            }
        }
    }
}
template <int L>
__device__ __forceinline__ int reduce_base(int h) { return h * (16 / L); }
// NOTE: with stages s = L/2 ... 1 halving the vector each time, lane bit s selects the upper half of the current
// vector, so the surviving block of lane h starts at sum_s (h & s ? len_s : 0) = h * 16 / L.

// all-lanes sum over the L lanes of a ray
template <int L>
__device__ __forceinline__ float seg_sum(float v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, L);
    return v;
}

// ascending bitonic sort of (t, id) across the L lanes of a ray; empty slots carry t = +inf
template <int L>
__device__ __forceinline__ void seg_bitonic_sort(float& t, int& id, int h) {
#pragma unroll
    for (int k = 2; k <= L; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const float ot = __shfl_xor_sync(0xffffffffu, t, j, L);
            const int oi = __shfl_xor_sync(0xffffffffu, id, j, L);
            const bool asc = (h & k) == 0;
            const bool lower = (h & j) == 0;
            const bool mine_first = (t < ot) || (t == ot && id < oi);
            const bool keep_mine = (lower == asc) ? mine_first : !mine_first;
            if (!keep_mine) { t = ot; id = oi; }
        }
    }
}

// 256-bit read-only global load (32-byte aligned address)
__device__ __forceinline__ void ldg256(const float* p, float* v) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}

struct HitEval {
    bool contrib;          // passes every test and lies before the termination point
    float alpha, G, w, T_prev, cmd, opac;
    float3 normal;         // flipped world-space normal
    float nsign;
    float2 uv, scale;
    float4 rot;
    float3 mu, tu, tv, xyz, c;
    bool clamped[3];
    float oth[2];
    int gidx;
};

// Evaluate this lane's hit and run the transmittance scan.  Returns T_final (same for the L lanes of the ray).
// When `dL_drgb` is given (backward), shU[k] = w * sum_c sh[k][c] * dL/drgb_c (unclamped channels): the per-hit part of
// dL/d(direction) through the SH colour, formed while the coefficients are in registers.
template <int L>
__device__ __forceinline__ float eval_and_scan(const envgs_trace_args& a, bool valid, float tmx, int gidx, float3 ray_o, float3 ray_d,
                                               const float* __restrict__ basis, int nbasis, bool sh_vec4, float3 campos, int h,
                                               HitEval& e, const float* dL_drgb = nullptr, float* shU = nullptr) {
    e.gidx = gidx;
    bool cand = valid;
    e.alpha = 0.0f; e.G = 0.0f; e.cmd = 1.0f; e.opac = 0.0f; e.nsign = 1.0f;
    e.normal = f3(0, 0, 0); e.c = f3(0, 0, 0); e.oth[0] = e.oth[1] = 0.0f;
    e.clamped[0] = e.clamped[1] = e.clamped[2] = false;
    e.uv = make_float2(0, 0); e.scale = make_float2(1, 1); e.rot = make_float4(1, 0, 0, 0);
    e.mu = e.tu = e.tv = e.xyz = f3(0, 0, 0);
    if (valid) {
        const float dpt = tmx;   // first chunk: payload.dpt == 0 (forward.cu:300)
        e.xyz = ray_o + dpt * ray_d;
        e.mu = ld3(a.means3D, gidx);
        e.scale = reinterpret_cast<const float2*>(a.scales)[gidx];
        e.rot = reinterpret_cast<const float4*>(a.rotations)[gidx];
        Rot3 R = quat_to_rot(e.rot);
        e.tu = R.c0; e.tv = R.c1;
        const float3 nr = R.c2;
        const float Tx = dot(e.tu, e.mu), Ty = dot(e.tv, e.mu);
        const float u1 = e.tu.x * e.xyz.x + e.tu.y * e.xyz.y + e.tu.z * e.xyz.z + (-Tx);
        const float v1 = e.tv.x * e.xyz.x + e.tv.y * e.xyz.y + e.tv.z * e.xyz.z + (-Ty);
        e.uv = make_float2(px_div(u1, e.scale.x, px_rcp(e.scale.x)), px_div(v1, e.scale.y, px_rcp(e.scale.y)));
        const float rho3d = e.uv.x * e.uv.x + e.uv.y * e.uv.y;
        const float cosv = -sum3((e.mu - campos) * nr);
        if (cosv == 0.0f) cand = false;
        e.nsign = cosv > 0.0f ? 1.0f : -1.0f;
        e.normal = e.nsign * nr;
        e.cmd = dpt;
        if (e.cmd < TR_START_OFFSET) cand = false;
        const float power = -0.5f * rho3d;
        if (power > 0.0f) cand = false;
        e.opac = a.opacities[gidx];
        e.G = px_exp(power);
        e.alpha = fminf(0.99f, e.opac * e.G);
        if (e.alpha < 1.0f / 255.0f) cand = false;
    }
    const float one_m = cand ? 1.0f - e.alpha : 1.0f;
    const float T_incl = seg_scan_mul<L>(one_m, h);
    float T_excl = __shfl_up_sync(0xffffffffu, T_incl, 1, L);
    if (h == 0) T_excl = 1.0f;
    // termination: the first candidate whose T_next drops below 1e-4 and everything behind it are dropped (forward.cu:370-372)
    const unsigned int seg_shift = (threadIdx.x & 31) & ~(L - 1);
    const unsigned int seg_mask = L == 32 ? 0xffffffffu : ((1u << L) - 1u);
    const unsigned int stop_mask = (__ballot_sync(0xffffffffu, cand && (T_incl < 0.0001f)) >> seg_shift) & seg_mask;
    const int stop = stop_mask ? (__ffs(stop_mask) - 1) : L;
    e.contrib = cand && h < stop;
    e.T_prev = T_excl;
    e.w = e.contrib ? e.alpha * T_excl : 0.0f;
    // final transmittance = product over contributors = inclusive product at the last lane before `stop`
    const int last = stop - 1;
    float T_final = __shfl_sync(0xffffffffu, T_incl, last < 0 ? 0 : last, L);
    if (last < 0) T_final = 1.0f;

    if (e.contrib) {   // colour and auxiliaries only where they are used
        if (a.colors_precomp == nullptr) {
            const float* sh = a.shs + (size_t)gidx * a.sh_coeffs * 3;
            float3 res = f3(0, 0, 0);
            float gr = 0.0f, gg = 0.0f, gb = 0.0f;
            const bool sh_vec8 = sh_vec4 && ((a.sh_coeffs * 3) % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.shs) & 31) == 0);
            if (sh_vec8 && shU == nullptr) {
                // forward, 256-bit loads (LDG.E.256, new on sm_100): one request per 32-byte sector of the 192-byte row
                // instead of two 128-bit requests — the rows of a warp's 32 hits are scattered, so the kernel is bound by L1
                // request slots, not bytes.  Each chunk is consumed as it arrives: flat element f = 8q+i is coefficient f/3
                // of channel f%3 (same summation order per channel as the scalar path).
                float acc[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int q = 0; q < 6; q++) {
                    if (q * 8 < nbasis * 3) {
                        float v8[8];
                        ldg256(sh + 8 * q, v8);
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int k = (8 * q + i) / 3;
                            if (k < nbasis) acc[(8 * q + i) % 3] = fmaf(basis[k], v8[i], acc[(8 * q + i) % 3]);
                        }
                    }
                }
                res = f3(acc[0] + 0.5f, acc[1] + 0.5f, acc[2] + 0.5f);
            } else if (sh_vec4) {
                float tmp[48];
                if (sh_vec8) {
#pragma unroll
                    for (int q = 0; q < 6; q++) {
                        if (q * 8 < nbasis * 3) ldg256(sh + 8 * q, tmp + 8 * q);
                        else {
#pragma unroll
                            for (int i = 0; i < 8; i++) tmp[8 * q + i] = 0.0f;
                        }
                    }
                } else {
                    const float4* s4 = reinterpret_cast<const float4*>(sh);
#pragma unroll
                    for (int q = 0; q < 12; q++) {
                        if (q * 4 < nbasis * 3) { float4 v = __ldg(&s4[q]); tmp[4 * q] = v.x; tmp[4 * q + 1] = v.y; tmp[4 * q + 2] = v.z; tmp[4 * q + 3] = v.w; }
                        else { tmp[4 * q] = tmp[4 * q + 1] = tmp[4 * q + 2] = tmp[4 * q + 3] = 0.0f; }
                    }
                }
#pragma unroll
                for (int k = 0; k < 16; k++)
                    if (k < nbasis) res += basis[k] * f3(tmp[3 * k], tmp[3 * k + 1], tmp[3 * k + 2]);
                res.x += 0.5f; res.y += 0.5f; res.z += 0.5f;
                if (shU != nullptr) {
                    gr = res.x < 0 ? 0.0f : dL_drgb[0] * e.w; gg = res.y < 0 ? 0.0f : dL_drgb[1] * e.w; gb = res.z < 0 ? 0.0f : dL_drgb[2] * e.w;
#pragma unroll
                    for (int k = 0; k < 16; k++) shU[k] = k < nbasis ? tmp[3 * k] * gr + tmp[3 * k + 1] * gg + tmp[3 * k + 2] * gb : 0.0f;
                }
            } else {
                for (int k = 0; k < nbasis; k++) res += basis[k] * f3(sh[3 * k], sh[3 * k + 1], sh[3 * k + 2]);
                res.x += 0.5f; res.y += 0.5f; res.z += 0.5f;
                if (shU != nullptr) {
                    gr = res.x < 0 ? 0.0f : dL_drgb[0] * e.w; gg = res.y < 0 ? 0.0f : dL_drgb[1] * e.w; gb = res.z < 0 ? 0.0f : dL_drgb[2] * e.w;
#pragma unroll
                    for (int k = 0; k < 16; k++) shU[k] = k < nbasis ? sh[3 * k] * gr + sh[3 * k + 1] * gg + sh[3 * k + 2] * gb : 0.0f;
                }
            }
            e.clamped[0] = res.x < 0; e.clamped[1] = res.y < 0; e.clamped[2] = res.z < 0;
            e.c = f3(fmaxf(res.x, 0.0f), fmaxf(res.y, 0.0f), fmaxf(res.z, 0.0f));
        } else {
            e.c = ld3(a.colors_precomp, gidx);
        }
        if (a.others_precomp != nullptr) {
            e.oth[0] = a.others_precomp[2 * (size_t)gidx];
            e.oth[1] = a.others_precomp[2 * (size_t)gidx + 1];
        }
    }
    return T_final;
}

// rays without any hit: background colour, everything else zero (forward.cu:444-449 with an empty loop)
__device__ __forceinline__ void shade_empty_body(const FastParams& p) {
    const unsigned int n = p.class_count[0];
    const envgs_trace_args& a = p.a;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const size_t ray = (size_t)p.class_list[k];
        const float3 o = ld3(a.ray_o, (int)ray), d = ld3(a.ray_d, (int)ray);
        const float b0 = a.background[0], b1 = a.background[1], b2 = a.background[2];
        p.o.out_rgb[ray * 3] = b0; p.o.out_rgb[ray * 3 + 1] = b1; p.o.out_rgb[ray * 3 + 2] = b2;
        p.o.out_dpt[ray] = 0.0f; p.o.out_acc[ray] = 0.0f;
        p.o.out_norm[ray * 3] = 0.0f; p.o.out_norm[ray * 3 + 1] = 0.0f; p.o.out_norm[ray * 3 + 2] = 0.0f;
        p.o.out_dist[ray * 3] = 0.0f; p.o.out_dist[ray * 3 + 1] = 0.0f; p.o.out_dist[ray * 3 + 2] = 0.0f;
        p.o.out_aux[ray * 2] = 0.0f; p.o.out_aux[ray * 2 + 1] = 0.0f;
        float4* m4 = reinterpret_cast<float4*>(p.o.mid_val + ray * 32);
        m4[0] = make_float4(o.x, o.y, o.z, d.x);
        m4[1] = make_float4(d.y, d.z, 0.0f, 0.0f);
        m4[2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        m4[3] = make_float4(0.0f, b0, b1, b2);
        if (!p.mid_half) m4[4] = m4[5] = m4[6] = m4[7] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

#ifndef SFWD_MINBLOCKS
#define SFWD_MINBLOCKS 4
#endif
template <int L>
__device__ __forceinline__ void shade_fwd_body(const FastParams& p, int cls) {
    constexpr int RPW = 32 / L;              // rays per warp
    constexpr int VPL = 16 / L;              // reduced components per lane
    const envgs_trace_args& a = p.a;
    const int lane = threadIdx.x & 31;
    const int h = lane & (L - 1);
    const unsigned int n = p.class_count[cls];
    const int* __restrict__ list = p.class_list + (size_t)cls * p.R;
    const float3 campos = f3(a.campos[0], a.campos[1], a.campos[2]);
    const int nbasis = (a.sh_degree + 1) * (a.sh_degree + 1);
    const bool use_sh = a.colors_precomp == nullptr;
    const bool sh_vec4 = use_sh && ((a.sh_coeffs * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.shs) & 15) == 0);
    const unsigned int warps_total = (gridDim.x * blockDim.x) >> 5;
    const unsigned int n_items = (n + RPW - 1) / RPW;

    // Software pipeline over the warp's items: the chain list -> ray -> (count, hit, origin, direction) -> surfel -> SH row is
    // five dependent memory round trips; the first three are issued one and two items ahead so that only the surfel and SH
    // fetches are exposed (ncu before: 13.6 long-scoreboard stall cycles per issued instruction at 35 % issue utilisation).
    struct RayPre { int cnt; float2 hit; float3 o, d; };
    auto fetch_ray = [&](unsigned int it) -> int {
        const unsigned int li = it * RPW + (lane / L);
        return (it < n_items && li < n) ? list[li] : -1;
    };
    auto fetch_pre = [&](int r) -> RayPre {
        RayPre q;
        const size_t rr = r < 0 ? 0 : (size_t)r;
        // per-ray data is read once: streaming (evict-first) accesses keep the 60 MB surfel set resident in the 126 MB L2
        q.cnt = __ldcs(p.counts + rr);
        q.hit = __ldcs(p.hits + rr * TF_MAXHITS + h);          // slots >= cnt hold stale bytes; masked below
        q.o = f3(__ldcs(a.ray_o + 3 * rr), __ldcs(a.ray_o + 3 * rr + 1), __ldcs(a.ray_o + 3 * rr + 2));
        q.d = f3(__ldcs(a.ray_d + 3 * rr), __ldcs(a.ray_d + 3 * rr + 1), __ldcs(a.ray_d + 3 * rr + 2));
        return q;
    };
    unsigned int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int ray_cur = fetch_ray(item), ray_nxt = fetch_ray(item + warps_total);
    RayPre pre_cur = fetch_pre(ray_cur);
    for (; item < n_items; item += warps_total) {
        const int ray_nn = fetch_ray(item + 2 * warps_total);
        const RayPre pre_nxt = fetch_pre(ray_nxt);
        const bool live = ray_cur >= 0;
        const size_t ray = live ? (size_t)ray_cur : 0;
        const int cnt = live ? pre_cur.cnt : 0;
        const float3 ray_o = pre_cur.o, ray_d = pre_cur.d;
        float basis[16];
        if (use_sh) sh_basis(a.sh_degree, normalize3(ray_d), basis);
        const bool valid = live && h < cnt;
        float hit_t = TR_DEPTH_INF * 4.0f;
        int hit_id = 0x7fffffff;
        if (valid) { hit_t = pre_cur.hit.x; hit_id = __float_as_int(pre_cur.hit.y); }
        ray_cur = ray_nxt; ray_nxt = ray_nn; pre_cur = pre_nxt;
        seg_bitonic_sort<L>(hit_t, hit_id, h);       // slots [0,cnt) now hold the hits in ascending distance
        HitEval e;
        const float T_final = eval_and_scan<L>(a, valid, hit_t, hit_id, ray_o, ray_d, basis, nbasis, sh_vec4, campos, h, e);

        // distortion bookkeeping needs exclusive prefix sums of w*m and w*m^2 (forward.cu:405-409)
        const float m = FAR_N / (FAR_N - NEAR_N) * (1.0f - NEAR_N * px_rcp_v(e.cmd));
        const float wm = e.w * m, wmm = e.w * m * m;
        const float M1_ex = seg_scan_add<L>(wm, h) - wm, M2_ex = seg_scan_add<L>(wmm, h) - wmm;
        float v[16];
        v[0] = e.w * (m * m * (1.0f - e.T_prev) + M2_ex - 2.0f * m * M1_ex);
        v[1] = wm; v[2] = wmm; v[3] = 0.0f; v[4] = 0.0f; v[5] = 0.0f;
        v[6] = e.w * e.cmd; v[7] = 0.0f;
        v[8] = e.w * e.normal.x; v[9] = e.w * e.normal.y; v[10] = e.w * e.normal.z;
        v[11] = e.w * e.oth[0]; v[12] = e.w * e.oth[1];
        v[13] = e.w * e.c.x; v[14] = e.w * e.c.y; v[15] = e.w * e.c.z;
        reduce16<L>(v, h);   // v[j], j < VPL: ray-wide sum of component reduce_base(h) + j  (components follow the mid_val layout)

        if (live) {
            if (a.training && e.contrib) atomicAdd(&p.o.a_weights[e.gidx], e.w);
            float* mid = p.o.mid_val + ray * 32;
            const int base = reduce_base<L>(h);
            float mvs[VPL];
#pragma unroll
            for (int j = 0; j < VPL; j++) {
                const int c = base + j;    // mid_val slot (config.h:18-26): 0-2 ray_o, 3-5 ray_d, 6 D, 7 A, 8-10 N, 11-12 aux, 13-15 rgb (+T*bg)
                const float tot = v[j];
                float mv;
                if (c < 3) mv = c == 0 ? ray_o.x : (c == 1 ? ray_o.y : ray_o.z);
                else if (c < 6) mv = c == 3 ? ray_d.x : (c == 4 ? ray_d.y : ray_d.z);
                else if (c == 7) mv = 1.0f - T_final;
                else if (c >= 13) mv = tot + T_final * a.background[c - 13];
                else mv = tot;
                mvs[j] = mv;
                if (c < 3) __stcs(p.o.out_dist + ray * 3 + c, tot);                 // dist, M1, M2
                else if (c == 6) __stcs(p.o.out_dpt + ray, mv);
                else if (c == 7) __stcs(p.o.out_acc + ray, mv);
                else if (c >= 8 && c <= 10) __stcs(p.o.out_norm + ray * 3 + (c - 8), mv);
                else if (c == 11 || c == 12) __stcs(p.o.out_aux + ray * 2 + (c - 11), mv);
                else if (c >= 13) __stcs(p.o.out_rgb + ray * 3 + (c - 13), mv);     // max_trace_depth == 0: f_this = s_prod = 1
            }
            // the ray's 128-byte mid_val row (second segment zeroed) leaves as VPL-wide vector stores: the ray's L lanes
            // cover one contiguous 64-byte half each
            if (VPL == 4) {
                __stcs(reinterpret_cast<float4*>(mid + base), make_float4(mvs[0], mvs[1 % VPL], mvs[2 % VPL], mvs[3 % VPL]));
                if (!p.mid_half) __stcs(reinterpret_cast<float4*>(mid + 16 + base), make_float4(0.0f, 0.0f, 0.0f, 0.0f));
            } else if (VPL == 2) {
                __stcs(reinterpret_cast<float2*>(mid + base), make_float2(mvs[0], mvs[1 % VPL]));
                if (!p.mid_half) __stcs(reinterpret_cast<float2*>(mid + 16 + base), make_float2(0.0f, 0.0f));
            } else {
                __stcs(mid + base, mvs[0]);
                if (!p.mid_half) __stcs(mid + 16 + base, 0.0f);
            }
        }
    }
}

// ONE launch for all hit classes: a warp that runs out of 4-lane rays moves straight on to the 8- and 16-lane ones, so there
// are no idle tails and launch gaps between the classes (three short kernels back to back left the GPU partly idle 3 times).
__global__ void __launch_bounds__(256, SFWD_MINBLOCKS) shade_fwd_kernel(const FastParams p) {
    shade_empty_body(p);
    shade_fwd_body<4>(p, 1);
    shade_fwd_body<8>(p, 2);
    shade_fwd_body<16>(p, 3);
}

// d basis_k / d(x,y,z) of the SH polynomials (RAST backward.cu:58-122 collects the same terms per colour channel)
__device__ __forceinline__ float3 sh_basis_grad(int k, float3 d) {
    const float x = d.x, y = d.y, z = d.z;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    switch (k) {
        case 1: return f3(0.0f, -SH_C1, 0.0f);
        case 2: return f3(0.0f, 0.0f, SH_C1);
        case 3: return f3(-SH_C1, 0.0f, 0.0f);
        case 4: return SH_C2[0] * f3(y, x, 0.0f);
        case 5: return SH_C2[1] * f3(0.0f, z, y);
        case 6: return SH_C2[2] * f3(-2.0f * x, -2.0f * y, 4.0f * z);
        case 7: return SH_C2[3] * f3(z, 0.0f, x);
        case 8: return SH_C2[4] * f3(2.0f * x, -2.0f * y, 0.0f);
        case 9: return SH_C3[0] * f3(6.0f * xy, 3.0f * xx - 3.0f * yy, 0.0f);
        case 10: return SH_C3[1] * f3(yz, xz, xy);
        case 11: return SH_C3[2] * f3(-2.0f * xy, 4.0f * zz - xx - 3.0f * yy, 8.0f * yz);
        case 12: return SH_C3[3] * f3(-6.0f * xz, -6.0f * yz, 6.0f * zz - 3.0f * xx - 3.0f * yy);
        case 13: return SH_C3[4] * f3(4.0f * zz - 3.0f * xx - yy, -2.0f * xy, 8.0f * xz);
        case 14: return SH_C3[5] * f3(2.0f * xz, -2.0f * yz, xx - yy);
        case 15: return SH_C3[6] * f3(3.0f * xx - 3.0f * yy, -6.0f * xy, 0.0f);
        default: return f3(0.0f, 0.0f, 0.0f);
    }
}

constexpr int GR_STRIDE = 20;   // floats per staged hit-gradient record

#ifndef SBWD_THREADS
#define SBWD_THREADS 256
#endif
#ifndef SBWD_MINBLOCKS
#define SBWD_MINBLOCKS 2
#endif
template <int L>
__device__ __forceinline__ void shade_bwd_body(const FastParams& p, int cls, float* s_rec, float* s_basis) {
    // per-hit gradient records staged in shared memory so the 61 atomics of a contributing hit are issued by all lanes of
    // the ray instead of one lane after another (s_rec: one record per thread = per hit slot; s_basis: 8 x 16 floats per warp)
    constexpr int RPW = 32 / L;
    constexpr int VPL = 16 / L;
    const envgs_trace_args& a = p.a;
    const envgs_trace_grads& g = p.g;
    const int lane = threadIdx.x & 31;
    const int h = lane & (L - 1);
    const int ray_slot = threadIdx.x / L;                 // ray index within the block
    const unsigned int n = p.class_count[cls];
    const int* __restrict__ list = p.class_list + (size_t)cls * p.R;
    const float3 campos = f3(a.campos[0], a.campos[1], a.campos[2]);
    const int nbasis = (a.sh_degree + 1) * (a.sh_degree + 1);
    const bool use_sh = a.colors_precomp == nullptr;
    const bool sh_vec4 = use_sh && ((a.sh_coeffs * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.shs) & 15) == 0);
    const bool has_oth = a.others_precomp != nullptr;
    const unsigned int warps_total = (gridDim.x * blockDim.x) >> 5;
    const unsigned int n_items = (n + RPW - 1) / RPW;
    const unsigned int seg_shift = lane & ~(L - 1);
    const unsigned int seg_mask = (1u << L) - 1u;

    for (unsigned int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < n_items; item += warps_total) {
        const unsigned int li = item * RPW + (lane / L);
        const bool in_list = li < n;
        const size_t ray = in_list ? (size_t)list[li] : 0;
        const int cnt = in_list ? p.counts[ray] : 0;
        // saved forward state of the (only) segment
        const float* mid = g.mid_val + ray * 32;
        const float3 ray_o = f3(__ldcs(mid + MID_RAYO), __ldcs(mid + MID_RAYO + 1), __ldcs(mid + MID_RAYO + 2));
        const float3 ray_d = f3(__ldcs(mid + MID_RAYD), __ldcs(mid + MID_RAYD + 1), __ldcs(mid + MID_RAYD + 2));
        // backward.cu:952-957 with i = 0: s_prev = 1
        const bool live = in_list && !(length3(ray_d) == 0.0f || 1.0f <= a.specular_threshold);
        const float seg_rgb[3] = {__ldcs(mid + MID_RGB), __ldcs(mid + MID_RGB + 1), __ldcs(mid + MID_RGB + 2)};
        const float seg_dpt = __ldcs(mid + MID_DPT), seg_acc = __ldcs(mid + MID_ACC);
        const float3 seg_norm = f3(__ldcs(mid + MID_NORM), __ldcs(mid + MID_NORM + 1), __ldcs(mid + MID_NORM + 2));
        const float seg_aux[2] = {__ldcs(mid + MID_AUX), __ldcs(mid + MID_AUX + 1)};
        const float dL_drgb[3] = {__ldcs(g.dL_drgb + ray * 3), __ldcs(g.dL_drgb + ray * 3 + 1), __ldcs(g.dL_drgb + ray * 3 + 2)};
        const float dL_ddpt = __ldcs(g.dL_ddpt + ray);
        const float3 dL_dnorm = ld3(g.dL_dnorm, (int)ray);
        const float dL_daux[2] = {__ldcs(g.dL_daux + ray * 2), __ldcs(g.dL_daux + ray * 2 + 1)};
        const float T_final_saved = 1.0f - seg_acc;

        float basis[16];
        const float3 dirn = normalize3(ray_d);
        if (use_sh) sh_basis(a.sh_degree, dirn, basis);
        float* sb = s_basis + (threadIdx.x >> 5) * 128 + (lane / L) * 16;   // the scatter phase reads the basis from shared memory: written now so
        if (use_sh && h == 0) {                        // the 16 registers die after the colour evaluation
#pragma unroll
            for (int k = 0; k < 16; k++) sb[k] = k < nbasis ? basis[k] : 0.0f;
        }
        const bool valid = live && h < cnt;
        float hit_t = TR_DEPTH_INF * 4.0f;
        int hit_id = 0x7fffffff;
        if (valid) { const float2 hit = __ldcs(p.hits + ray * TF_MAXHITS + h); hit_t = hit.x; hit_id = __float_as_int(hit.y); }
        seg_bitonic_sort<L>(hit_t, hit_id, h);
        HitEval e;
        float shU[16];
#pragma unroll
        for (int k = 0; k < 16; k++) shU[k] = 0.0f;
        eval_and_scan<L>(a, valid, hit_t, hit_id, ray_o, ray_d, basis, nbasis, sh_vec4, campos, h, e, dL_drgb, use_sh ? shU : nullptr);
        // SH direction gradient (done first: it retires the 16 shU registers before the long gradient chain): dL/ddir = sum_k U_k * grad basis_k(dir), U_k = sum over hits of shU[k]
        float3 extra = f3(0, 0, 0);
        if (use_sh) {
            reduce16<L>(shU, h);
            const int base = reduce_base<L>(h);
            float sx = 0.0f, sy = 0.0f, sz = 0.0f;
#pragma unroll
            for (int j = 0; j < VPL; j++) {
                const int k = base + j;
                const float3 gb = sh_basis_grad(k < nbasis ? k : 0, dirn);
                sx += shU[j] * gb.x; sy += shU[j] * gb.y; sz += shU[j] * gb.z;
            }
            sx = seg_sum<L>(sx); sy = seg_sum<L>(sy); sz = seg_sum<L>(sz);
            if (live) extra = dnormvdv(ray_d, f3(sx, sy, sz));   // backward.cu:182-184
        }

        // inclusive prefix sums the reference keeps in C, O, D, N while walking front to back
        const float Cx = seg_scan_add<L>(e.w * e.c.x, h), Cy = seg_scan_add<L>(e.w * e.c.y, h), Cz = seg_scan_add<L>(e.w * e.c.z, h);
        const float O0 = seg_scan_add<L>(e.w * e.oth[0], h), O1 = seg_scan_add<L>(e.w * e.oth[1], h);
        const float Dp = seg_scan_add<L>(e.w * e.cmd, h);
        const float Nx = seg_scan_add<L>(e.w * e.normal.x, h), Ny = seg_scan_add<L>(e.w * e.normal.y, h), Nz = seg_scan_add<L>(e.w * e.normal.z, h);

        float3 g_ray_o = f3(0, 0, 0), g_ray_d = f3(0, 0, 0);
        if (e.contrib) {
            const int gidx = e.gidx;
            const float alpha = e.alpha, T_prev = e.T_prev, w = e.w, G = e.G, cmd = e.cmd;
            const float r1ma = px_rcp_v(1.0f - alpha);
            const float cc[3] = {e.c.x, e.c.y, e.c.z}, Cc[3] = {Cx, Cy, Cz}, O[2] = {O0, O1};
            float dL_dalpha = 0.0f;
            float dL_dcolor[3];
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                dL_dcolor[ch] = dL_drgb[ch] * w;
                dL_dalpha += dL_drgb[ch] * (T_prev * cc[ch] - (seg_rgb[ch] - Cc[ch]) * r1ma);
                dL_dalpha += dL_drgb[ch] * a.background[ch] * (-T_final_saved * r1ma);
            }
            float dL_doth[2] = {0.0f, 0.0f};
            if (has_oth) {
#pragma unroll
                for (int ch = 0; ch < 2; ch++) {
                    dL_doth[ch] = dL_daux[ch] * w;
                    dL_dalpha += dL_daux[ch] * (T_prev * e.oth[ch] - (seg_aux[ch] - O[ch]) * r1ma);
                }
            }
            const float dL_dD = dL_ddpt * w;
            dL_dalpha += dL_ddpt * (T_prev * cmd - (seg_dpt - Dp) * r1ma);
            const float3 dL_dN = dL_dnorm * w;
            dL_dalpha += sum3(dL_dnorm * (T_prev * e.normal - (seg_norm - f3(Nx, Ny, Nz)) * r1ma));
            const float dL_dG = dL_dalpha * e.opac;
            const float dL_dopac = dL_dalpha * G;   // straight-through cap (backward.cu:801-806)

            // compute_transmat_uv_backward (backward.cu:233-335), plane form t = m.(mu-o)/(m.d), m = t_u x t_v
            const float isx = px_rcp_v(e.scale.x), isy = px_rcp_v(e.scale.y);
            const float2 dL_duv = make_float2(dL_dG * -G * e.uv.x, dL_dG * -G * e.uv.y);
            const float3 xm = e.xyz - e.mu;
            float3 dR0 = (dL_duv.x * isx) * xm;
            float3 dR1 = (dL_duv.y * isy) * xm;
            const float3 dR2 = e.nsign * dL_dN;
            const float2 dL_dscale = make_float2(dL_dG * (G * e.uv.x * e.uv.x * isx), dL_dG * (G * e.uv.y * e.uv.y * isy));
            float3 dL_dmean = dL_dG * (G * ((e.uv.x * isx) * e.tu + (e.uv.y * isy) * e.tv));
            const float3 dL_dxyz = (dL_duv.x * isx) * e.tu + (dL_duv.y * isy) * e.tv;
            const float dL_dd = dL_dD + dot(dL_dxyz, ray_d);
            const float3 mvec = cross(e.tu, e.tv);
            const float iqd = px_rcp_v(dot(mvec, ray_d));
            const float3 cvec = e.mu - ray_o;
            const float pn = dot(mvec, cvec);
            const float3 dt_dm = (cvec - (pn * iqd) * ray_d) * iqd;
            dR0 += dL_dd * cross(e.tv, dt_dm);
            dR1 += dL_dd * cross(dt_dm, e.tu);
            dL_dmean += (dL_dd * iqd) * mvec;
            g_ray_o = dL_dxyz - (dL_dd * iqd) * mvec;
            g_ray_d = dL_dxyz * cmd - (dL_dd * pn * iqd * iqd) * mvec;
            Rot3 gR; gR.c0 = dR0; gR.c1 = dR1; gR.c2 = dR2;
            const float4 dL_drot = quat_to_rot_vjp(e.rot, gR);

            float3 dRGB = f3(0, 0, 0);
            if (use_sh) dRGB = f3(e.clamped[0] ? 0.0f : dL_dcolor[0], e.clamped[1] ? 0.0f : dL_dcolor[1], e.clamped[2] ? 0.0f : dL_dcolor[2]);
            float* rec = s_rec + threadIdx.x * GR_STRIDE;
            rec[0] = dL_dscale.x; rec[1] = dL_dscale.y;
            rec[2] = dL_drot.x; rec[3] = dL_drot.y; rec[4] = dL_drot.z; rec[5] = dL_drot.w;
            rec[6] = dL_dmean.x; rec[7] = dL_dmean.y; rec[8] = dL_dmean.z;
            rec[9] = dL_dcolor[0]; rec[10] = dL_dcolor[1]; rec[11] = dL_dcolor[2];
            rec[12] = dL_dopac; rec[13] = dL_doth[0]; rec[14] = dL_doth[1];
            rec[15] = dRGB.x; rec[16] = dRGB.y; rec[17] = dRGB.z;
            rec[18] = __int_as_float(gidx);
        }
        // ---- cooperative scatter: the ray's lanes walk its contributing hits; lane q takes components q, q+L, ... ----
        {
            __syncwarp();
            unsigned int cm = (__ballot_sync(0xffffffffu, e.contrib) >> seg_shift) & seg_mask;
            while (__any_sync(0xffffffffu, cm != 0u)) {
                if (cm != 0u) {
                    const int hit = __ffs(cm) - 1;
                    cm &= cm - 1u;
                    const float* rec = s_rec + (ray_slot * L + hit) * GR_STRIDE;
                    const size_t gi = (size_t)__float_as_int(rec[18]);
#pragma unroll
                    for (int q0 = 0; q0 < 16; q0 += L) {
                        const int q = q0 + h;
                        if (q < 15) {
                            float* dst;
                            if (q < 2) dst = g.dL_dscales + 2 * gi + q;
                            else if (q < 6) dst = g.dL_drotations + 4 * gi + (q - 2);
                            else if (q < 9) dst = g.dL_dmeans3D + 3 * gi + (q - 6);
                            else if (q < 12) dst = g.dL_dcolors + 3 * gi + (q - 9);
                            else if (q == 12) dst = g.dL_dopacities + gi;
                            else dst = has_oth ? g.dL_dothers + 2 * gi + (q - 13) : nullptr;
                            if (dst != nullptr) atomicAdd(dst, rec[q]);
                        }
                    }
                    if (use_sh) {
#pragma unroll
                        for (int k0 = 0; k0 < 16; k0 += L) {
                            const int k = k0 + h;
                            if (k < nbasis) {
                                const float bk = sb[k];
                                float* dsh = g.dL_dshs + gi * a.sh_coeffs * 3 + 3 * k;
                                atomicAdd(dsh + 0, bk * rec[15]);
                                atomicAdd(dsh + 1, bk * rec[16]);
                                atomicAdd(dsh + 2, bk * rec[17]);
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
        // ---- per-ray gradients: sum over the hits ----
        float v[16];
        v[0] = g_ray_o.x; v[1] = g_ray_o.y; v[2] = g_ray_o.z; v[3] = g_ray_d.x; v[4] = g_ray_d.y; v[5] = g_ray_d.z;
#pragma unroll
        for (int i = 6; i < 16; i++) v[i] = 0.0f;
        reduce16<L>(v, h);
        if (in_list) {
            const int base = reduce_base<L>(h);
#pragma unroll
            for (int j = 0; j < VPL; j++) {
                const int c = base + j;
                if (c < 3) g.dL_dray_o[ray * 3 + c] = live ? v[j] : 0.0f;
                else if (c < 6) g.dL_dray_d[ray * 3 + (c - 3)] = live ? v[j] + (c == 3 ? extra.x : (c == 4 ? extra.y : extra.z)) : 0.0f;
            }
        }
    }
}

// (one launch per class here: the merged form of the forward kernel spilled at 128 registers and ran 4 % slower)
template <int L>
__global__ void __launch_bounds__(SBWD_THREADS, SBWD_MINBLOCKS) shade_bwd_kernel(const FastParams p, int cls) {
    __shared__ float s_rec[SBWD_THREADS * GR_STRIDE];
    __shared__ float s_basis[(SBWD_THREADS / 32) * 128];
    shade_bwd_body<L>(p, cls, s_rec, s_basis);
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
int launch_trace(bool bwd, Lbvh& bvh, unsigned int* counter, unsigned long long* stats, const envgs_trace_args& a, const envgs_trace_out* o,
                 const envgs_trace_grads* g, const int* ray_list, const unsigned int* ray_count, cudaStream_t stream, int mid_half);

bool trace_fast_eligible(const envgs_trace_args& a, const void* cache) {
    // max_trace_depth 1 runs segment by segment (trace_forward_fast / trace_backward_fast dispatch on it); its second segment
    // only exists when the traced set carries a specular channel
    if (cache == nullptr || a.start_from_first != 0 || a.max_trace_depth > 1) return false;
    if (a.max_trace_depth == 1 && a.others_precomp == nullptr) return false;
    const char* f = getenv("ENVGS_TRACE_FORCE_MONOLITHIC");
    return !(f && f[0] == '1');
}

// The few rays with a full first chunk run in the general kernel on a side stream, concurrently with the shading kernel
// (they write disjoint rays); the caller's stream joins before returning control.
struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; };

static int trace_forward_segment(Lbvh& bvh, unsigned int* counter, unsigned long long* stats, const envgs_trace_args& a, const envgs_trace_out& o,
                                 void* cache, cudaStream_t stream, const SideStream& side, int mid_half) {
    static int sm_counts[64] = {0}, occs[64] = {0};   // per device id (the handle may live on any GPU of the process)
    int dev = 0;
    ENVGS_CUDA_TRY(cudaGetDevice(&dev));
    const int slot = dev & 63;
    if (sm_counts[slot] == 0) {
        int n = 0, oc = 0;
        ENVGS_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        ENVGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, traverse_kernel<false>, TF_THREADS, 0));
        occs[slot] = oc < 1 ? 1 : oc;
        sm_counts[slot] = n;
    }
    const int sm_count = sm_counts[slot], occ = occs[slot];
    FastParams p;
    memset(&p, 0, sizeof(p));
    p.nodes = bvh.nodes; p.nodes4 = bvh.nodes4; p.prims = bvh.prims; p.a = a; p.o = o;
    carve_cache((char*)cache, a.H, a.W, p);
    p.work_counter = counter;
    p.stats = stats;
    p.mid_half = mid_half;
    {
        const char* e = getenv("ENVGS_TRACE_REFILL");
        p.refill_threshold = e ? atoi(e) : 28;
    }
    if (a.H >= 4) { p.tiles_x = (a.W + 7) / 8; p.n_items = p.tiles_x * ((a.H + 3) / 4); }
    else { p.tiles_x = 0; p.n_items = (int)(((long long)a.H * a.W + 31) / 32); }
    const int R = a.H * a.W;
    ENVGS_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
    ENVGS_CUDA_TRY(cudaMemsetAsync(p.overflow_count, 0, 8 * sizeof(unsigned int), stream));   // overflow + 4 class counters
    int blocks = sm_count * occ;
    int need = (p.n_items * 32 + (TF_THREADS / 4) - 1) / (TF_THREADS / 4);
    if (blocks > need) blocks = need > 0 ? need : 1;
    {
        ENVGS_TIMED("trace_traverse", stream);
        if (stats) traverse_kernel<true><<<blocks, TF_THREADS, 0, stream>>>(p);
        else traverse_kernel<false><<<blocks, TF_THREADS, 0, stream>>>(p);
        ENVGS_LAUNCH_CHECK("trace_traverse", false, stream);
    }
    // rays whose first chunk was full: the reference's multi-chunk loop, run by the general kernel on the device-side list.  It is
    // forked BEFORE the (tiny) class-list kernels so that its few CTAs are resident when the persistent shading kernel, which
    // fills every SM for its whole duration, starts.
    ENVGS_CUDA_TRY(cudaEventRecord(side.fork, stream));
    ENVGS_CUDA_TRY(cudaStreamWaitEvent(side.stream, side.fork, 0));
    int rc = launch_trace(false, bvh, counter, stats, a, &o, nullptr, p.overflow_list, p.overflow_count, side.stream, mid_half);
    if (rc != ENVGS_OK) return rc;
    {
        ENVGS_TIMED("trace_class_lists", stream);
        int rcl = build_class_lists(p, stream);
        if (rcl != ENVGS_OK) return rcl;
    }
    ENVGS_CUDA_TRY(cudaEventRecord(side.join, side.stream));
    {
        ENVGS_TIMED("trace_shade_fwd", stream);
        // statically strided items: several CTA "waves" per SM let late waves fill the holes early finishers leave
        static int per_sm = 0;
        if (per_sm == 0) { const char* e = getenv("ENVGS_SHADE_GRID"); per_sm = e ? atoi(e) : 64; if (per_sm < 1) per_sm = 1; }
        const int sblocks = sm_count * per_sm;
        shade_fwd_kernel<<<sblocks, 256, 0, stream>>>(p);
        ENVGS_LAUNCH_CHECK("trace_shade_fwd", false, stream);
    }
    ENVGS_CUDA_TRY(cudaStreamWaitEvent(stream, side.join, 0));
    return ENVGS_OK;
}

static int trace_backward_segment(Lbvh& bvh, unsigned int* counter, unsigned long long* stats, const envgs_trace_args& a, const envgs_trace_grads& g,
                                  void* cache, cudaStream_t stream, const SideStream& side, int mid_half) {
    FastParams p;
    memset(&p, 0, sizeof(p));
    p.nodes = bvh.nodes; p.prims = bvh.prims; p.a = a; p.g = g;
    p.mid_half = mid_half;
    carve_cache((char*)cache, a.H, a.W, p);
    const size_t R = (size_t)a.H * a.W;
    int sm_count = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); }
    // rays without hits (class 0) and rays skipped by the kernels keep a zero ray gradient
    ENVGS_CUDA_TRY(cudaMemsetAsync(g.dL_dray_o, 0, R * 3 * sizeof(float), stream));
    ENVGS_CUDA_TRY(cudaMemsetAsync(g.dL_dray_d, 0, R * 3 * sizeof(float), stream));
    ENVGS_CUDA_TRY(cudaEventRecord(side.fork, stream));
    ENVGS_CUDA_TRY(cudaStreamWaitEvent(side.stream, side.fork, 0));
    int rc = launch_trace(true, bvh, counter, stats, a, nullptr, &g, p.overflow_list, p.overflow_count, side.stream, mid_half);
    if (rc != ENVGS_OK) return rc;
    ENVGS_CUDA_TRY(cudaEventRecord(side.join, side.stream));
    {
        ENVGS_TIMED("trace_shade_bwd", stream);
        static int per_sm_b = 0;   // CTAs per SM launched (items are statically strided; see trace_forward_fast)
        if (per_sm_b == 0) { const char* e = getenv("ENVGS_SHADE_GRID_BWD"); per_sm_b = e ? atoi(e) : SBWD_MINBLOCKS; if (per_sm_b < 1) per_sm_b = 1; }
        const int sblocks = sm_count * per_sm_b;
        shade_bwd_kernel<4><<<sblocks, SBWD_THREADS, 0, stream>>>(p, 1);
        ENVGS_LAUNCH_CHECK("trace_shade_bwd4", false, stream);
        shade_bwd_kernel<8><<<sblocks, SBWD_THREADS, 0, stream>>>(p, 2);
        ENVGS_LAUNCH_CHECK("trace_shade_bwd8", false, stream);
        shade_bwd_kernel<16><<<sblocks, SBWD_THREADS, 0, stream>>>(p, 3);
        ENVGS_LAUNCH_CHECK("trace_shade_bwd16", false, stream);
    }
    ENVGS_CUDA_TRY(cudaStreamWaitEvent(stream, side.join, 0));
    return ENVGS_OK;
}

// ------------------------------------------------------------------------------------------------------
// two segments (max_trace_depth = 1), segment by segment on the hit-parallel kernels
// ------------------------------------------------------------------------------------------------------
// tracePath (TRACE/optix_tracer/forward.cu:453-559, backward.cu:890-1086) restated at the granularity of whole segments:
//   forward : segment 0 on the given rays -> per ray: continue? (|N0| != 0 and spec0 > threshold) -> reflected ray
//             (d1 = d0 - 2 (n.d0) n, o1 = o0 + D0 d0 + 1e-5 d1) -> segment 1 on those rays -> rgb = C0 (1 - spec0) + spec0 C1
//   backward: segment 1 with dL/dC1 = spec0 dL/drgb -> its ray gradient is folded into segment 0's upstream
//             (dL/dD0, dL/dN0, dL/dspec0) and into the running ray gradient -> segment 0 -> ray gradient += carried part.
// Each segment is the single-segment fast path above (traverse -> class lists -> hit-parallel shading, general kernel for the
// >= 16-hit rays), with its own hit cache; rays that do not continue carry NaN into segment 1, find no hit and cost one box test.
// Scratch layout behind the two hit caches (floats): see trace_cache_bytes2.
struct Seg2Scratch {
    float *ray1_o, *ray1_d;                                   // R x 3 each
    float *o1_rgb, *o1_dpt, *o1_acc, *o1_norm, *o1_dist, *o1_aux;   // segment-1 outputs nobody reads (its mid_val half is what counts)
    float *g_rgb1, *zeros;                                    // backward: upstream of segment 1; R x 3 zeros for its other inputs
    float *g_o1, *g_d1;                                       // segment 1's ray gradient
    float *u_rgb, *u_dpt, *u_norm, *u_aux, *base_d;           // segment 0's upstream + carried direction gradient
    void *cache0, *cache1;
};

size_t trace_cache_bytes2(int H, int W, int max_trace_depth) {
    if (max_trace_depth <= 0) return trace_cache_bytes(H, W);
    const size_t R = (size_t)H * W, c = align_up(trace_cache_bytes(H, W), 256);
    return 2 * c + align_up(R * 24 * sizeof(float), 256) + 256;    // max(forward 19 R, backward 24 R) floats of scratch
}

static Seg2Scratch carve_seg2(void* cache, int H, int W) {
    Seg2Scratch s;
    const size_t R = (size_t)H * W, c = align_up(trace_cache_bytes(H, W), 256);
    char* b = (char*)cache;
    s.cache0 = b; s.cache1 = b + c;
    float* f = reinterpret_cast<float*>(b + 2 * c);
    // forward use                               // backward use of the same region
    s.ray1_o = f; s.ray1_d = f + 3 * R;          s.g_rgb1 = f; s.zeros = f + 3 * R;
    s.o1_rgb = f + 6 * R; s.o1_dpt = f + 9 * R;  s.g_o1 = f + 6 * R; s.g_d1 = f + 9 * R;
    s.o1_acc = f + 10 * R; s.o1_norm = f + 11 * R; s.o1_dist = f + 14 * R; s.o1_aux = f + 17 * R;   // ends at 19 R
    s.u_rgb = f + 12 * R; s.u_dpt = f + 15 * R; s.u_norm = f + 16 * R; s.u_aux = f + 19 * R; s.base_d = f + 21 * R;   // ends at 24 R
    return s;
}

__device__ __forceinline__ bool seg_continues(const float* mid0, float thr, float3& N0) {
    N0 = f3(mid0[MID_NORM], mid0[MID_NORM + 1], mid0[MID_NORM + 2]);
    return !(length3(N0) == 0.0f || mid0[MID_AUX] <= thr);                  // forward.cu:551
}

__global__ void __launch_bounds__(256) seg1_rays_kernel(int R, const float* __restrict__ mid_val, float thr, float* __restrict__ ray1_o,
                                                         float* __restrict__ ray1_d) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float* m = mid_val + (size_t)r * 32;
    float3 N0;
    float3 o1 = f3(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000), __int_as_float(0x7fc00000)), d1 = o1;   // NaN: no traversal
    if (seg_continues(m, thr, N0)) {
        const float3 ray_ot = f3(m[MID_RAYO], m[MID_RAYO + 1], m[MID_RAYO + 2]);
        float3 ray_dt = f3(m[MID_RAYD], m[MID_RAYD + 1], m[MID_RAYD + 2]);
        const float3 E = ray_ot + m[MID_DPT] * ray_dt;                      // forward.cu:447, :553-557
        const float3 n = normalize3(N0);
        ray_dt = ray_dt - 2.0f * dot(n, ray_dt) * n;
        o1 = E + TR_STEP_EPSILON * ray_dt;
        d1 = ray_dt;
    }
    ray1_o[3 * (size_t)r] = o1.x; ray1_o[3 * (size_t)r + 1] = o1.y; ray1_o[3 * (size_t)r + 2] = o1.z;
    ray1_d[3 * (size_t)r] = d1.x; ray1_d[3 * (size_t)r + 1] = d1.y; ray1_d[3 * (size_t)r + 2] = d1.z;
}

__global__ void __launch_bounds__(256) seg2_combine_kernel(int R, float* __restrict__ mid_val, float thr, float* __restrict__ out_rgb) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float* m = mid_val + (size_t)r * 32;
    float3 N0;
    const bool cont = seg_continues(m, thr, N0);
    float s_prod = 1.0f;
    const float f0 = 1.0f - m[MID_AUX];                                     // forward.cu:523-528 with max_trace_depth = 1
    float rgb[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) rgb[ch] = 0.0f + m[MID_RGB + ch] * f0 * s_prod;
    s_prod *= m[MID_AUX];
    if (cont) {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) rgb[ch] += m[16 + MID_RGB + ch] * 1.0f * s_prod;
    } else {
        float4* z = reinterpret_cast<float4*>(m + 16);                       // the path stopped: no second segment was recorded
        z[0] = z[1] = z[2] = z[3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    out_rgb[3 * (size_t)r] = rgb[0]; out_rgb[3 * (size_t)r + 1] = rgb[1]; out_rgb[3 * (size_t)r + 2] = rgb[2];
}

// backward.cu:915-1023 for two segments.  Segment 1's upstream colour gradient.
__global__ void __launch_bounds__(256) seg2_bwd_prep1_kernel(int R, const float* __restrict__ mid_val, const float* __restrict__ dLo_rgb,
                                                              float* __restrict__ g_rgb1, float* __restrict__ zeros) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float s_prod = mid_val[(size_t)r * 32 + MID_AUX];                 // s_prod after the bookkeeping loop = spec0; f_this = 1
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        g_rgb1[3 * (size_t)r + ch] = dLo_rgb[3 * (size_t)r + ch] * 1.0f * s_prod;
        zeros[3 * (size_t)r + ch] = 0.0f;
    }
}

// Segment 0's upstream (colour, depth, normal, aux) and the ray gradient carried over from segment 1.
__global__ void __launch_bounds__(256) seg2_bwd_prep0_kernel(int R, const float* __restrict__ mid_val, float thr, const float* __restrict__ out_rgb,
                                                              const float* __restrict__ dLo_rgb, const float* __restrict__ dLo_dpt,
                                                              const float* __restrict__ dLo_norm, const float* __restrict__ dLo_aux,
                                                              const float* __restrict__ g_o1, const float* __restrict__ g_d1,
                                                              float* __restrict__ u_rgb, float* __restrict__ u_dpt, float* __restrict__ u_norm,
                                                              float* __restrict__ u_aux, float* __restrict__ base_d) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float* m = mid_val + (size_t)r * 32;
    const float seg_aux0 = m[MID_AUX];
    const float3 seg_norm = f3(m[MID_NORM], m[MID_NORM + 1], m[MID_NORM + 2]);
    const float3 ray_dt = f3(m[MID_RAYD], m[MID_RAYD + 1], m[MID_RAYD + 2]);
    const float seg_dpt = m[MID_DPT];
    float3 g_ray_o = ld3(g_o1, r), g_ray_d = ld3(g_d1, r);
    float dL_ddpt = 0.0f, dL_daux0 = 0.0f;
    float3 dL_dnorm = f3(0, 0, 0);
    const float f_this = 1.0f - seg_aux0;
    // The reference divides its running specular product by spec0 only when it actually walks segment 1 (`s_prod /= s_prev` sits
    // behind the `continue`, backward.cu:955-957, :988): a ray that stopped after its first segment therefore back-propagates
    // its colour scaled by spec0.  Reproduced, not corrected (parity).
    float3 N0;
    const float s_prod = seg_continues(m, thr, N0) ? seg_aux0 / seg_aux0 : seg_aux0;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const float go = dLo_rgb[3 * (size_t)r + ch];
        u_rgb[3 * (size_t)r + ch] = go * f_this * s_prod;
        // C_prev after the suffix bookkeeping (backward.cu:915-935): out_rgb + 1 * (0 - C0)
        const float C_prev = out_rgb[3 * (size_t)r + ch] + 1.0f * (0.0f - m[MID_RGB + ch]);
        dL_daux0 += go * (C_prev / seg_aux0);
    }
    if (length3(seg_norm) != 0.0f) {                                        // backward.cu:1003-1021
        g_ray_d = g_ray_d + g_ray_o * TR_STEP_EPSILON;
        const float3 n = normalize3(seg_norm);
        dL_ddpt += sum3(g_ray_o * ray_dt);
        const float3 up = g_ray_d * -2.0f;
        const float dotnd = dot(n, ray_dt);
        const float3 dn = f3(up.x * (dotnd + ray_dt.x * n.x) + up.y * ray_dt.x * n.y + up.z * ray_dt.x * n.z,
                             up.x * ray_dt.y * n.x + up.y * (dotnd + ray_dt.y * n.y) + up.z * ray_dt.y * n.z,
                             up.x * ray_dt.z * n.x + up.y * ray_dt.z * n.y + up.z * (dotnd + ray_dt.z * n.z));
        dL_dnorm += dnormvdv(seg_norm, dn);
        const float3 dd = f3(up.x * n.x * n.x + up.y * n.x * n.y + up.z * n.x * n.z,
                             up.x * n.y * n.x + up.y * n.y * n.y + up.z * n.y * n.z,
                             up.x * n.z * n.x + up.y * n.z * n.y + up.z * n.z * n.z);
        g_ray_d = g_ray_o * seg_dpt + g_ray_d + dd;
    }
    u_dpt[r] = dL_ddpt + dLo_dpt[r];
    const float3 gn = dL_dnorm + ld3(dLo_norm, r);
    u_norm[3 * (size_t)r] = gn.x; u_norm[3 * (size_t)r + 1] = gn.y; u_norm[3 * (size_t)r + 2] = gn.z;
    u_aux[2 * (size_t)r] = dL_daux0 + dLo_aux[2 * (size_t)r];
    u_aux[2 * (size_t)r + 1] = dLo_aux[2 * (size_t)r + 1];
    base_d[3 * (size_t)r] = g_ray_d.x; base_d[3 * (size_t)r + 1] = g_ray_d.y; base_d[3 * (size_t)r + 2] = g_ray_d.z;
}

__global__ void __launch_bounds__(256) seg2_bwd_finish_kernel(int R, const float* __restrict__ g_o1, const float* __restrict__ base_d,
                                                               float* __restrict__ dL_dray_o, float* __restrict__ dL_dray_d) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)R * 3) return;
    dL_dray_o[e] += g_o1[e];
    dL_dray_d[e] += base_d[e];
}

int trace_forward_fast(Lbvh& bvh, unsigned int* counter, unsigned long long* stats, const envgs_trace_args& a, const envgs_trace_out& o,
                       void* cache, cudaStream_t stream, const SideStream& side) {
    if (a.max_trace_depth == 0) return trace_forward_segment(bvh, counter, stats, a, o, cache, stream, side, 0);
    const int R = a.H * a.W, nb = (R + 255) / 256;
    Seg2Scratch sc = carve_seg2(cache, a.H, a.W);
    envgs_trace_args a0 = a;
    a0.max_trace_depth = 0;
    int rc = trace_forward_segment(bvh, counter, stats, a0, o, sc.cache0, stream, side, 1);      // writes mid_val[0:16], out_* of segment 0
    if (rc != ENVGS_OK) return rc;
    {
        ENVGS_TIMED("trace_seg_glue", stream);
        seg1_rays_kernel<<<nb, 256, 0, stream>>>(R, o.mid_val, a.specular_threshold, sc.ray1_o, sc.ray1_d);
        ENVGS_LAUNCH_CHECK("seg1_rays", false, stream);
    }
    envgs_trace_args a1 = a0;
    a1.ray_o = sc.ray1_o; a1.ray_d = sc.ray1_d;
    envgs_trace_out o1 = o;
    o1.out_rgb = sc.o1_rgb; o1.out_dpt = sc.o1_dpt; o1.out_acc = sc.o1_acc; o1.out_norm = sc.o1_norm; o1.out_dist = sc.o1_dist;
    o1.out_aux = sc.o1_aux; o1.mid_val = o.mid_val + 16;                                         // a_weights: both segments accumulate
    rc = trace_forward_segment(bvh, counter, stats, a1, o1, sc.cache1, stream, side, 1);
    if (rc != ENVGS_OK) return rc;
    {
        ENVGS_TIMED("trace_seg_glue", stream);
        seg2_combine_kernel<<<nb, 256, 0, stream>>>(R, o.mid_val, a.specular_threshold, o.out_rgb);
        ENVGS_LAUNCH_CHECK("seg2_combine", false, stream);
    }
    return ENVGS_OK;
}

int trace_backward_fast(Lbvh& bvh, unsigned int* counter, unsigned long long* stats, const envgs_trace_args& a, const envgs_trace_grads& g,
                        void* cache, cudaStream_t stream, const SideStream& side) {
    if (a.max_trace_depth == 0) return trace_backward_segment(bvh, counter, stats, a, g, cache, stream, side, 0);
    const int R = a.H * a.W, nb = (R + 255) / 256;
    Seg2Scratch sc = carve_seg2(cache, a.H, a.W);
    envgs_trace_args a0 = a;
    a0.max_trace_depth = 0;
    {
        ENVGS_TIMED("trace_seg_glue", stream);
        seg2_bwd_prep1_kernel<<<nb, 256, 0, stream>>>(R, g.mid_val, g.dL_drgb, sc.g_rgb1, sc.zeros);
        ENVGS_LAUNCH_CHECK("seg2_bwd_prep1", false, stream);
    }
    // segment 1: rays and saved state come from its half of mid_val; only its colour has an upstream gradient
    envgs_trace_grads g1 = g;
    g1.mid_val = g.mid_val + 16;
    g1.dL_drgb = sc.g_rgb1; g1.dL_ddpt = sc.zeros; g1.dL_dacc = sc.zeros; g1.dL_dnorm = sc.zeros; g1.dL_ddist = sc.zeros; g1.dL_daux = sc.zeros;
    g1.dL_dray_o = sc.g_o1; g1.dL_dray_d = sc.g_d1;
    int rc = trace_backward_segment(bvh, counter, stats, a0, g1, sc.cache1, stream, side, 1);
    if (rc != ENVGS_OK) return rc;
    {
        ENVGS_TIMED("trace_seg_glue", stream);
        seg2_bwd_prep0_kernel<<<nb, 256, 0, stream>>>(R, g.mid_val, a.specular_threshold, g.out_rgb, g.dL_drgb, g.dL_ddpt, g.dL_dnorm, g.dL_daux, sc.g_o1, sc.g_d1,
                                                      sc.u_rgb, sc.u_dpt, sc.u_norm, sc.u_aux, sc.base_d);
        ENVGS_LAUNCH_CHECK("seg2_bwd_prep0", false, stream);
    }
    envgs_trace_grads g0 = g;
    g0.dL_drgb = sc.u_rgb; g0.dL_ddpt = sc.u_dpt; g0.dL_dnorm = sc.u_norm; g0.dL_daux = sc.u_aux;
    rc = trace_backward_segment(bvh, counter, stats, a0, g0, sc.cache0, stream, side, 1);
    if (rc != ENVGS_OK) return rc;
    {
        ENVGS_TIMED("trace_seg_glue", stream);
        seg2_bwd_finish_kernel<<<(unsigned)(((size_t)R * 3 + 255) / 256), 256, 0, stream>>>(R, sc.g_o1, sc.base_d, g.dL_dray_o, g.dL_dray_d);
        ENVGS_LAUNCH_CHECK("seg2_bwd_finish", false, stream);
    }
    return ENVGS_OK;
}

}  // namespace envgs
