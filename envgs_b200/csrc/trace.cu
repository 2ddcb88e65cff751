// trace.cu — environment-surfel ray tracer, forward and backward, hand-written for sm_100a.
//
// Replaces the OptiX programs of diff-surfel-tracing (TRACE/optix_tracer/forward.cu:219-660, backward.cu:508-1209):
//   * traversal: LBVH (lbvh.cu), persistent warps, one ray per lane, per-lane stack in shared memory.
//     A warp owns an 8x4 pixel block so neighbouring reflection rays walk the same nodes.
//   * ordering: the reference collects the 16 nearest hits per optixTrace in an any-hit k-buffer while visiting
//     EVERY primitive on the ray (any-hit + optixIgnoreIntersection never shrinks tmax).  Here the same ascending
//     16-entry buffer prunes subtrees whose entry distance is beyond the current 16th hit, and children are
//     visited near-to-far, so a chunk touches a few dozen nodes instead of the whole ray extent.
//   * the chunk restart (origin advanced by t_last + 1e-5, forward.cu:438-441), 256-contributor cap, alpha
//     thresholds, early termination and every accumulation follow the reference statement by statement.
//   * SH evaluation, compositing and (backward) all per-hit gradients are fused into the traversal kernel.
#include "lbvh.cuh"

namespace envgs {

constexpr int TR_THREADS = 128;
constexpr int TR_SSTACK = 32;    // per-lane stack entries in shared memory; deeper levels spill to local memory
constexpr int TR_LSTACK = 64;

struct TraceParams {
    const BvhNode* nodes;
    const QuadPrim* prims;
    envgs_trace_args a;
    envgs_trace_out o;
    envgs_trace_grads g;
    unsigned int* work_counter;
    unsigned long long* stats;   // optional debug counters: [rays, nodes, prim tests, hits inserted, chunks, contributions]
    int n_items;      // warp work items
    int tiles_x;      // 8x4 tiles per row (0 => linear mapping)
    const int* ray_list;             // optional: trace only these rays (overflow fallback of the fast path) ...
    const unsigned int* ray_count;   // ... whose number lives on the device
    int mid_half;                    // 1: this call is one segment of a segment-by-segment trace (trace_fast.cu): `mid_val` points at that
                                     // segment's 16-float half of the row and only those 16 floats are written
};

struct HitBuf {
    float t[TR_CHUNK];
    int id[TR_CHUNK];
    int cnt;
    int n_nodes, n_prims;   // debug statistics
};

__device__ __forceinline__ void hit_insert(HitBuf& hb, float t, int id) {
    // ascending insertion, identical to the any-hit program (forward.cu:627-656)
    hb.cnt += 1;
    float ct = t;
    int ci = id;
#pragma unroll
    for (int i = 0; i < TR_CHUNK; i++) {
        if (hb.t[i] > ct) {
            float tt = hb.t[i]; int ti = hb.id[i];
            hb.t[i] = ct; hb.id[i] = ci;
            ct = tt; ci = ti;
        }
    }
}

__device__ __forceinline__ float safe_inv(float d) {
    return 1.0f / (fabsf(d) > 1e-30f ? d : copysignf(1e-30f, d));
}

__device__ __forceinline__ void prim_test(const QuadPrim* __restrict__ prims, int sid, float3 o, float3 d, HitBuf& hb) {
    const float4 pn = __ldg(&prims[sid].n);
    const float4 pv = __ldg(&prims[sid].v1);
    float3 n = f3(pn.x, pn.y, pn.z);
    float q = dot(n, d);
    hb.n_prims++;
    if (q == 0.0f) return;
    float3 v1 = f3(pv.x, pv.y, pv.z);
    float t = dot(n, v1 - o) / q;
    if (!(t > 0.0f && t < hb.t[TR_CHUNK - 1])) return;
    const float4 pu = __ldg(&prims[sid].eu);
    const float4 pw = __ldg(&prims[sid].ev);
    if (pw.w == 0.0f) return;
    float3 r = (o + t * d) - v1;
    float a = dot(r, f3(pu.x, pu.y, pu.z)) * pv.w;
    float b = dot(r, f3(pw.x, pw.y, pw.z)) * pu.w;
    if (a < 0.0f || a > 1.0f || b < 0.0f || b > 1.0f) return;
    hit_insert(hb, t, sid);
}

// Collect the (up to) 16 nearest surfel hits with t in (0, 1e16) along o + t d, ascending in hb.
__device__ __forceinline__ void gather_chunk(const BvhNode* __restrict__ nodes, const QuadPrim* __restrict__ prims, float3 o,
                                             float3 d, HitBuf& hb, int* sstack /* [TR_SSTACK][TR_THREADS] */) {
#pragma unroll
    for (int i = 0; i < TR_CHUNK; i++) { hb.t[i] = TR_DEPTH_INF; hb.id[i] = 0; }
    hb.cnt = 0;
    hb.n_nodes = 0; hb.n_prims = 0;
    if (!isfinite(d.x + d.y + d.z + o.x + o.y + o.z)) return;
    const float3 inv = f3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
    int lstack[TR_LSTACK];
    int sp = 0;
    int node = 0;
    while (true) {
        // node >= 0: internal
        const float4 A = __ldg(&nodes[node].a), B = __ldg(&nodes[node].b), C = __ldg(&nodes[node].c);
        const int c0 = __ldg(&nodes[node].child0), c1 = __ldg(&nodes[node].child1);
        const float bound = hb.t[TR_CHUNK - 1];
        hb.n_nodes++;
        float tn0, tn1;
        bool h0, h1;
        {
            float ax = (A.x - o.x) * inv.x, bx = (A.w - o.x) * inv.x;
            float ay = (A.y - o.y) * inv.y, by = (B.x - o.y) * inv.y;
            float az = (A.z - o.z) * inv.z, bz = (B.y - o.z) * inv.z;
            tn0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
            float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) * 1.0000004f;
            h0 = (tn0 <= tf) && (tf > 0.0f) && (tn0 < bound);
        }
        {
            float ax = (B.z - o.x) * inv.x, bx = (C.y - o.x) * inv.x;
            float ay = (B.w - o.y) * inv.y, by = (C.z - o.y) * inv.y;
            float az = (C.x - o.z) * inv.z, bz = (C.w - o.z) * inv.z;
            tn1 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
            float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) * 1.0000004f;
            h1 = (tn1 <= tf) && (tf > 0.0f) && (tn1 < bound);
        }
        // leaves are tested immediately (nearest first so the bound tightens early)
        int next = -1, defer = -1;
        if (h0 && h1) {
            bool first0 = tn0 <= tn1;
            int ca = first0 ? c0 : c1, cb = first0 ? c1 : c0;
            if (ca < 0) prim_test(prims, ~ca, o, d, hb); else next = ca;
            if (cb < 0) prim_test(prims, ~cb, o, d, hb);
            else if (next < 0) next = cb;
            else defer = cb;
        } else if (h0) {
            if (c0 < 0) prim_test(prims, ~c0, o, d, hb); else next = c0;
        } else if (h1) {
            if (c1 < 0) prim_test(prims, ~c1, o, d, hb); else next = c1;
        }
        if (defer >= 0) {
            if (sp < TR_SSTACK) sstack[sp * TR_THREADS] = defer;
            else if (sp < TR_SSTACK + TR_LSTACK) lstack[sp - TR_SSTACK] = defer;
            sp = min(sp + 1, TR_SSTACK + TR_LSTACK);   // deeper than 96 levels cannot happen with 63-bit keys + index tie-break
        }
        if (next >= 0) { node = next; continue; }
        if (sp == 0) break;
        sp--;
        node = (sp < TR_SSTACK) ? sstack[sp * TR_THREADS] : lstack[sp - TR_SSTACK];
    }
}

__device__ __forceinline__ void rot_rows(float4 q, float3& r0, float3& r1, float3& r2) {
    // rows of R^T == columns of R (TRACE auxiliary.h:414-440)
    Rot3 R = quat_to_rot(q);
    r0 = R.c0; r1 = R.c1; r2 = R.c2;
}

// splat2pixel (3x3) and projected centre for the first-segment low-pass filter (TRACE forward.cu:120-173)
__device__ __forceinline__ bool transmat_xy(float3 mu, float2 scale, float4 rot, float mod, const float* __restrict__ pm,
                                            int W, int H, float Pm[3][4], float3& s0, float3& s1, float3& s2, float2& xy) {
    Rot3 R = quat_to_rot(rot);
    float3 L0 = R.c0 * (mod * scale.x), L1 = R.c1 * (mod * scale.y);
    float hw = float(W) / 2.0f, cw = float(W - 1) / 2.0f, hh = float(H) / 2.0f, ch = float(H - 1) / 2.0f;
#pragma unroll
    for (int c = 0; c < 4; c++) {   // P = ndc2pix * world2ndc, world2ndc[r][c] = pm[4*c + r]
        float w0 = pm[4 * c + 0], w1 = pm[4 * c + 1], w3 = pm[4 * c + 3];
        Pm[0][c] = hw * w0 + cw * w3;
        Pm[1][c] = hh * w1 + ch * w3;
        Pm[2][c] = w3;
    }
    auto row = [&](int i) {
        return f3(Pm[i][0] * L0.x + Pm[i][1] * L0.y + Pm[i][2] * L0.z, Pm[i][0] * L1.x + Pm[i][1] * L1.y + Pm[i][2] * L1.z,
                  Pm[i][0] * mu.x + Pm[i][1] * mu.y + Pm[i][2] * mu.z + Pm[i][3]);
    };
    s0 = row(0); s1 = row(1); s2 = row(2);
    float3 t = f3(9.0f, 9.0f, -1.0f);
    float d = dot(t, s2 * s2);
    if (d == 0.0f) return false;
    float3 f = t / d;
    xy = make_float2(sum3(f * s0 * s2), sum3(f * s1 * s2));
    return true;
}

template <bool BWD>
__global__ void __launch_bounds__(TR_THREADS) trace_kernel(const TraceParams p) {
    __shared__ int s_stack[TR_SSTACK * TR_THREADS];
    const int lane = threadIdx.x & 31;
    int* sstack = s_stack + threadIdx.x;
    const envgs_trace_args& a = p.a;
    const int H = a.H, W = a.W;
    const float3 campos = f3(a.campos[0], a.campos[1], a.campos[2]);
    const float3 bgc = f3(a.background[0], a.background[1], a.background[2]);
    const int nbasis = (a.sh_degree + 1) * (a.sh_degree + 1);
    const bool use_sh = a.colors_precomp == nullptr;
    const bool sh_vec4 = use_sh && ((a.sh_coeffs * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.shs) & 15) == 0);

    while (true) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(p.work_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        int row, col;
        bool listed = true;
        if (p.ray_list != nullptr) {
            // Listed rays are the long ones (>= 16 hits) and few: they are spread over as many warps as the grid has —
            // 32 of them in one warp would serialise 32 divergent traversals (measured: 107 rays took 1.1 ms that way).
            const unsigned int n = *p.ray_count;
            const unsigned int warps = gridDim.x * (TR_THREADS / 32);
            const unsigned int rpw = min(32u, (n + warps - 1u) / warps);   // rays per warp (>= 1 when n > 0)
            if (n == 0u || item >= (n + rpw - 1u) / rpw) break;
            const unsigned int k = item * rpw + lane;
            listed = lane < rpw && k < n;
            const int r = listed ? p.ray_list[k] : 0;
            row = r / W; col = r % W;
        } else if (p.tiles_x > 0) {
            if (item >= (unsigned)p.n_items) break;
            int by = item / p.tiles_x, bx = item % p.tiles_x;
            row = by * 4 + (lane >> 3); col = bx * 8 + (lane & 7);
        } else {
            if (item >= (unsigned)p.n_items) break;
            long long r = (long long)item * 32 + lane;
            row = (int)(r / W); col = (int)(r % W);
        }
        const bool active = listed && row < H && col < W;
        if (active) {
            const size_t tidx = (size_t)row * W + col;
            const float3 ray_o0 = ld3(a.ray_o, (int)tidx), ray_d0 = ld3(a.ray_d, (int)tidx);

            // ---------------- forward path bookkeeping (tracePath, forward.cu:453-559) ----------------
            float mid[TR_MID_CHANNELS * (TR_MAX_TRACE_DEPTH + 1)];
            float out_rgb[3] = {0, 0, 0}, out_dpt = 0, out_acc = 0, out_aux[2] = {0, 0};
            float3 out_norm = f3(0, 0, 0), out_dist = f3(0, 0, 0);
            float3 g_ray_o = f3(0, 0, 0), g_ray_d = f3(0, 0, 0);   // BWD: running dL/d(ray of the segment above)
            float C_prev[3] = {0, 0, 0}, c_prev[3] = {0, 0, 0}, s_prod = 1.0f;
            float dLo_rgb[3] = {0, 0, 0}, dLo_dpt = 0, dLo_acc = 0, dLo_aux[2] = {0, 0};
            float3 dLo_norm = f3(0, 0, 0);
            const int nseg = a.max_trace_depth + 1;
            if (BWD) {
                const envgs_trace_grads& g = p.g;
#pragma unroll
                for (int i = 0; i < TR_MID_CHANNELS * (TR_MAX_TRACE_DEPTH + 1); i++)
                    mid[i] = (i < nseg * TR_MID_CHANNELS) ? g.mid_val[tidx * 32 + i] : 0.0f;
                for (int ch = 0; ch < 3; ch++) { out_rgb[ch] = g.out_rgb[tidx * 3 + ch]; dLo_rgb[ch] = g.dL_drgb[tidx * 3 + ch]; }
                dLo_dpt = g.dL_ddpt[tidx]; dLo_acc = g.dL_dacc[tidx];
                dLo_norm = ld3(g.dL_dnorm, (int)tidx);
                dLo_aux[0] = g.dL_daux[tidx * 2]; dLo_aux[1] = g.dL_daux[tidx * 2 + 1];
                // suffix bookkeeping of the specular chain (backward.cu:915-935)
                for (int ch = 0; ch < 3; ch++) C_prev[ch] = out_rgb[ch];
                for (int i = 0; i < a.max_trace_depth; i++) {
                    for (int ch = 0; ch < 3; ch++) {
                        float c_curr = mid[ch + MID_RGB + i * TR_MID_CHANNELS];
                        C_prev[ch] = C_prev[ch] + s_prod * (c_prev[ch] - c_curr);
                        c_prev[ch] = c_curr;
                    }
                    s_prod *= mid[MID_AUX + i * TR_MID_CHANNELS];
                }
            } else {
#pragma unroll
                for (int i = 0; i < TR_MID_CHANNELS * (TR_MAX_TRACE_DEPTH + 1); i++) mid[i] = 0.0f;
            }

            float3 ray_ot = ray_o0, ray_dt = ray_d0;
            for (int it = 0; it < nseg; it++) {
                const int seg = BWD ? (a.max_trace_depth - it) : it;
                // ---- per-segment upstream (BWD) ----
                float seg_rgb[3] = {0, 0, 0}, seg_dpt = 0, seg_acc = 0, seg_aux[2] = {0, 0};
                float3 seg_norm = f3(0, 0, 0);
                float dL_drgb[3] = {0, 0, 0}, dL_ddpt = 0, dL_dacc = 0, dL_daux[2] = {0, 0};
                float3 dL_dnorm = f3(0, 0, 0);
                if (BWD) {
                    ray_ot = f3(mid[MID_RAYO + seg * 16], mid[MID_RAYO + 1 + seg * 16], mid[MID_RAYO + 2 + seg * 16]);
                    ray_dt = f3(mid[MID_RAYD + seg * 16], mid[MID_RAYD + 1 + seg * 16], mid[MID_RAYD + 2 + seg * 16]);
                    float s_prev = (seg > 0) ? mid[MID_AUX + (seg - 1) * 16] : 1.0f;
                    if (length3(ray_dt) == 0.0f || s_prev <= a.specular_threshold) continue;
                    for (int ch = 0; ch < 3; ch++) seg_rgb[ch] = mid[ch + MID_RGB + seg * 16];
                    seg_dpt = mid[MID_DPT + seg * 16]; seg_acc = mid[MID_ACC + seg * 16];
                    seg_norm = f3(mid[MID_NORM + seg * 16], mid[MID_NORM + 1 + seg * 16], mid[MID_NORM + 2 + seg * 16]);
                    seg_aux[0] = mid[MID_AUX + seg * 16]; seg_aux[1] = mid[MID_AUX + 1 + seg * 16];
                    float f_this = (seg == a.max_trace_depth) ? 1.0f : 1.0f - seg_aux[0];
                    for (int ch = 0; ch < 3; ch++) dL_drgb[ch] = dLo_rgb[ch] * f_this * s_prod;
                    s_prod /= s_prev;
                    if (seg != a.max_trace_depth) {   // backward.cu:998-1023
                        for (int ch = 0; ch < 3; ch++) {
                            dL_daux[0] += dLo_rgb[ch] * (C_prev[ch] / seg_aux[0]);
                            c_prev[ch] = (seg > 0) ? mid[ch + MID_RGB + (seg - 1) * 16] : 0.0f;
                            C_prev[ch] = C_prev[ch] - s_prod * (c_prev[ch] - seg_rgb[ch]);
                        }
                        if (length3(seg_norm) != 0.0f) {
                            g_ray_d = g_ray_d + g_ray_o * TR_STEP_EPSILON;
                            float3 n = normalize3(seg_norm);
                            dL_ddpt += sum3(g_ray_o * ray_dt);
                            float3 up = g_ray_d * -2.0f;
                            float dotnd = dot(n, ray_dt);
                            // ddotndndn / ddotndndd (TRACE auxiliary.h:245-263)
                            float3 dn = f3(up.x * (dotnd + ray_dt.x * n.x) + up.y * ray_dt.x * n.y + up.z * ray_dt.x * n.z,
                                           up.x * ray_dt.y * n.x + up.y * (dotnd + ray_dt.y * n.y) + up.z * ray_dt.y * n.z,
                                           up.x * ray_dt.z * n.x + up.y * ray_dt.z * n.y + up.z * (dotnd + ray_dt.z * n.z));
                            dL_dnorm += dnormvdv(seg_norm, dn);
                            float3 dd = f3(up.x * n.x * n.x + up.y * n.x * n.y + up.z * n.x * n.z,
                                           up.x * n.y * n.x + up.y * n.y * n.y + up.z * n.y * n.z,
                                           up.x * n.z * n.x + up.y * n.z * n.y + up.z * n.z * n.z);
                            g_ray_d = g_ray_o * seg_dpt + g_ray_d + dd;
                        }
                    }
                    if (seg == 0) {
                        dL_ddpt += dLo_dpt; dL_dacc += dLo_acc; dL_dnorm += dLo_norm;
                        dL_daux[0] += dLo_aux[0]; dL_daux[1] += dLo_aux[1];
                    }
                }

                // ---------------- traceRay (forward.cu:219-450 / backward.cu:508-887) ----------------
                const bool first = (seg == 0) && (a.start_from_first != 0);
                const float min_depth = first ? NEAR_N : TR_START_OFFSET;
                const float T_threshold = 0.0001f;
                float Cc[3] = {0, 0, 0}, D = 0, Wacc = 0, dist = 0, M1 = 0, M2 = 0, O[2] = {0, 0};
                float3 N = f3(0, 0, 0);
                float T_prev = 1.0f, T_next = 1.0f;
                const float T_final = 1.0f - seg_acc;   // BWD only
                int last_gidx = -1, contributor = 0;
                float payload_dpt = 0.0f, dpt = 0.0f;
                float3 o_chunk = ray_ot;
                // SH basis of the (normalised) segment direction, shared by every hit on this ray
                float basis[16];
                const float3 dirn = normalize3(ray_dt);
                if (use_sh) sh_basis(a.sh_degree, dirn, basis);
                float3 dL_ddir_sh = f3(0, 0, 0);   // BWD: accumulated dL/d(normalised dir) through SH
                unsigned int st_nodes = 0, st_prims = 0, st_hits = 0, st_chunks = 0;

                const int max_chunks = a.P / TR_CHUNK + 2;   // every productive chunk consumes 16 distinct surfels
                for (int chunk = 0; chunk < max_chunks; chunk++) {
                    HitBuf hb;
                    gather_chunk(p.nodes, p.prims, o_chunk, ray_dt, hb, sstack);
                    st_nodes += hb.n_nodes; st_prims += hb.n_prims; st_hits += hb.cnt; st_chunks++;
                    // spill the sorted chunk to the (now idle) stack area so the compositing loop can index it
                    const int nh = min(hb.cnt, TR_CHUNK);
#pragma unroll
                    for (int i = 0; i < TR_CHUNK; i++) {
                        sstack[(2 * i) * TR_THREADS] = __float_as_int(hb.t[i]);
                        sstack[(2 * i + 1) * TR_THREADS] = hb.id[i];
                    }
                    bool stop = false;
                    for (int i = 0; i < nh; i++) {
                        const float tmx = __int_as_float(sstack[(2 * i) * TR_THREADS]);
                        const int gidx = sstack[(2 * i + 1) * TR_THREADS];
                        if (gidx == last_gidx) continue;
                        dpt = tmx + payload_dpt + (payload_dpt == 0.0f ? 0.0f : TR_STEP_EPSILON);
                        const float3 xyz = ray_ot + dpt * ray_dt;

                        // compute_transmat_uv (forward.cu:77-117)
                        const float3 mu = ld3(a.means3D, gidx);
                        const float2 scale = reinterpret_cast<const float2*>(a.scales)[gidx];
                        const float4 rot = reinterpret_cast<const float4*>(a.rotations)[gidx];
                        float3 tu, tv, nr;
                        rot_rows(rot, tu, tv, nr);
                        const float3 Tm = f3(dot(tu, mu), dot(tv, mu), dot(nr, mu));
                        const float u1 = tu.x * xyz.x + tu.y * xyz.y + tu.z * xyz.z + (-Tm.x);
                        const float v1 = tv.x * xyz.x + tv.y * xyz.y + tv.z * xyz.z + (-Tm.y);
                        const float2 uv = make_float2(px_div(u1, scale.x, px_rcp(scale.x)), px_div(v1, scale.y, px_rcp(scale.y)));   // same helpers as trace_fast.cu
                        const float rho3d = uv.x * uv.x + uv.y * uv.y;
                        float3 normal = nr;
                        float nsign = 1.0f;
                        {
                            float3 dirc = mu - campos;
                            float cosv = -sum3(dirc * normal);
                            if (cosv == 0.0f) continue;
                            nsign = cosv > 0.0f ? 1.0f : -1.0f;
                            normal = nsign * normal;
                        }
                        float rho2d, cmd;
                        float Pm[3][4];
                        float3 s2p0, s2p1, s2p2;
                        float2 xy = make_float2(0, 0);
                        if (first) {
                            bool ok = transmat_xy(mu, scale, rot, a.scale_modifier, a.projmatrix, W, H, Pm, s2p0, s2p1, s2p2, xy);
                            xy = make_float2(xy.x - (float)col, xy.y - (float)row);
                            rho2d = (xy.x * xy.x + xy.y * xy.y) * FILTER_INV_SQUARE;
                            if (!ok && rho3d > rho2d) continue;
                            cmd = (rho3d <= rho2d) ? dpt : s2p2.z;
                        } else {
                            rho2d = rho3d + 1.0f;
                            cmd = dpt;
                        }
                        if (cmd < min_depth) continue;
                        const float power = -0.5f * fminf(rho3d, rho2d);
                        if (power > 0.0f) continue;
                        const float opac = a.opacities[gidx];
                        const float G = px_exp(power);
                        const float alpha = fminf(0.99f, opac * G);
                        if (alpha < 1.0f / 255.0f) continue;
                        T_next = T_prev * (1.0f - alpha);
                        if (T_next < T_threshold) { stop = true; break; }
                        contributor++;
                        if (contributor > TR_MAX_INTERSECTION) { stop = true; break; }
                        last_gidx = gidx;
                        const float w = alpha * T_prev;

                        // colour
                        float3 c;
                        bool clamped[3] = {false, false, false};
                        if (use_sh) {
                            const float* sh = a.shs + (size_t)gidx * a.sh_coeffs * 3;
                            float3 res = f3(0, 0, 0);
                            if (sh_vec4) {
                                const float4* s4 = reinterpret_cast<const float4*>(sh);
                                float tmp[48];
#pragma unroll
                                for (int q = 0; q < 12; q++) {
                                    if (q * 4 < nbasis * 3) { float4 v = __ldg(&s4[q]); tmp[4 * q] = v.x; tmp[4 * q + 1] = v.y; tmp[4 * q + 2] = v.z; tmp[4 * q + 3] = v.w; }
                                    else { tmp[4 * q] = tmp[4 * q + 1] = tmp[4 * q + 2] = tmp[4 * q + 3] = 0.0f; }
                                }
#pragma unroll
                                for (int k = 0; k < 16; k++)
                                    if (k < nbasis) res += basis[k] * f3(tmp[3 * k], tmp[3 * k + 1], tmp[3 * k + 2]);
                            } else {
                                for (int k = 0; k < nbasis; k++) res += basis[k] * f3(sh[3 * k], sh[3 * k + 1], sh[3 * k + 2]);
                            }
                            res.x += 0.5f; res.y += 0.5f; res.z += 0.5f;
                            clamped[0] = res.x < 0; clamped[1] = res.y < 0; clamped[2] = res.z < 0;
                            c = f3(fmaxf(res.x, 0.0f), fmaxf(res.y, 0.0f), fmaxf(res.z, 0.0f));
                        } else {
                            c = ld3(a.colors_precomp, gidx);
                        }
                        Cc[0] += w * c.x; Cc[1] += w * c.y; Cc[2] += w * c.z;
                        float oth[2] = {0, 0};
                        if (a.others_precomp != nullptr) {
                            oth[0] = a.others_precomp[2 * (size_t)gidx]; oth[1] = a.others_precomp[2 * (size_t)gidx + 1];
                            O[0] += w * oth[0]; O[1] += w * oth[1];
                        }
                        if (!BWD || first) {
                            float aa = 1.0f - T_prev;
                            float m = FAR_N / (FAR_N - NEAR_N) * (1.0f - NEAR_N / cmd);
                            dist += w * (m * m * aa + M2 - 2.0f * m * M1);
                            M1 += w * m;
                            M2 += w * m * m;
                            Wacc += w;
                        }
                        D += w * cmd;
                        N += w * normal;

                        if (!BWD) {
                            if (a.training) atomicAdd(&p.o.a_weights[gidx], w);
                        } else {
                            // ---------------- per-hit backward (backward.cu:753-872) ----------------
                            const envgs_trace_grads& g = p.g;
                            float dL_dalpha = 0.0f;
                            const float cc[3] = {c.x, c.y, c.z};
                            const float bgv[3] = {bgc.x, bgc.y, bgc.z};
                            float dL_dcolor[3];
#pragma unroll
                            for (int ch = 0; ch < 3; ch++) {
                                dL_dcolor[ch] = dL_drgb[ch] * w;
                                dL_dalpha += dL_drgb[ch] * (T_prev * cc[ch] - (seg_rgb[ch] - Cc[ch]) / (1.0f - alpha));
                                dL_dalpha += dL_drgb[ch] * bgv[ch] * (-T_final / (1.0f - alpha));
                            }
                            if (a.others_precomp != nullptr) {
#pragma unroll
                                for (int ch = 0; ch < 2; ch++) {
                                    atomicAdd(&g.dL_dothers[2 * (size_t)gidx + ch], dL_daux[ch] * w);
                                    dL_dalpha += dL_daux[ch] * (T_prev * oth[ch] - (seg_aux[ch] - O[ch]) / (1.0f - alpha));
                                }
                            }
                            const float dL_dD = dL_ddpt * w;
                            dL_dalpha += dL_ddpt * (T_prev * cmd - (seg_dpt - D) / (1.0f - alpha));
                            const float3 dL_dN = dL_dnorm * w;
                            dL_dalpha += sum3(dL_dnorm * (T_prev * normal - (seg_norm - N) / (1.0f - alpha)));
                            if (first) dL_dalpha += dL_dacc * (T_prev * 1.0f - (seg_acc - Wacc) / (1.0f - alpha));
                            const float dL_dG = dL_dalpha * opac;
                            atomicAdd(&g.dL_dopacities[gidx], dL_dalpha * G);   // straight-through cap (backward.cu:801-806)

                            float2 dL_dscale;
                            float4 dL_drot;
                            float3 dL_dmean;
                            if (rho3d <= rho2d || !first) {
                                // compute_transmat_uv_backward (backward.cu:233-335), hit distance through the surfel plane
                                // t = m.(mu - o) / (m.d), m = t_u x t_v (identical to the quad-vertex form of the reference)
                                const float2 dL_duv = make_float2(dL_dG * -G * uv.x, dL_dG * -G * uv.y);
                                const float3 xm = xyz - mu;
                                float3 dR0 = (dL_duv.x / scale.x) * xm;
                                float3 dR1 = (dL_duv.y / scale.y) * xm;
                                const float3 dR2 = nsign * dL_dN;
                                dL_dscale = make_float2(dL_dG * (G * uv.x * uv.x / scale.x), dL_dG * (G * uv.y * uv.y / scale.y));
                                dL_dmean = dL_dG * (G * ((uv.x / scale.x) * tu + (uv.y / scale.y) * tv));
                                const float3 dL_dxyz = (dL_duv.x / scale.x) * tu + (dL_duv.y / scale.y) * tv;
                                const float dL_dd = dL_dD + dot(dL_dxyz, ray_dt);
                                const float3 m = cross(tu, tv);
                                const float qd = dot(m, ray_dt);
                                const float3 cvec = mu - ray_ot;
                                const float pn = dot(m, cvec);
                                const float3 dt_dm = (cvec - (pn / qd) * ray_dt) / qd;   // d t / d m
                                dR0 += dL_dd * cross(tv, dt_dm);
                                dR1 += dL_dd * cross(dt_dm, tu);
                                dL_dmean += dL_dd * (m / qd);
                                g_ray_o += dL_dxyz + dL_dd * (-(m / qd));
                                g_ray_d += dL_dxyz * cmd + dL_dd * (-pn / (qd * qd)) * m;
                                Rot3 gR; gR.c0 = dR0; gR.c1 = dR1; gR.c2 = dR2;
                                dL_drot = quat_to_rot_vjp(rot, gR);
                            } else {
                                // compute_transmat_xy_backward (backward.cu:395-463)
                                const float2 dL_dxy = make_float2(dL_dG * -G * FILTER_INV_SQUARE * xy.x, dL_dG * -G * FILTER_INV_SQUARE * xy.y);
                                float3 d0 = f3(0, 0, 0), d1 = f3(0, 0, 0), d2 = f3(0, 0, dL_dD);
                                const float3 t = f3(9.0f, 9.0f, -1.0f);
                                const float dd = dot(t, s2p2 * s2p2);
                                const float3 f = t / dd;
                                d0 += dL_dxy.x * f * s2p2;
                                d1 += dL_dxy.y * f * s2p2;
                                d2 += dL_dxy.x * f * s2p0 + dL_dxy.y * f * s2p1;
                                const float3 dL_df = dL_dxy.x * s2p0 * s2p2 + dL_dxy.y * s2p1 * s2p2;
                                const float dL_ddd = dot(dL_df, f) * (-1.0f / dd);
                                d2 += dL_ddd * t * s2p2 * 2.0f;
                                // dL_dsplat2world = P^T (4x3) * dL_dsplat2pixel (3x3)
                                float3 dsw[4];
#pragma unroll
                                for (int r = 0; r < 4; r++)
                                    dsw[r] = f3(Pm[0][r] * d0.x + Pm[1][r] * d1.x + Pm[2][r] * d2.x, Pm[0][r] * d0.y + Pm[1][r] * d1.y + Pm[2][r] * d2.y,
                                                Pm[0][r] * d0.z + Pm[1][r] * d1.z + Pm[2][r] * d2.z);
                                Rot3 gR;
                                gR.c0 = f3(dsw[0].x, dsw[1].x, dsw[2].x) * scale.x;
                                gR.c1 = f3(dsw[0].y, dsw[1].y, dsw[2].y) * scale.y;
                                gR.c2 = nsign * dL_dN;
                                dL_drot = quat_to_rot_vjp(rot, gR);
                                Rot3 Rm = quat_to_rot(rot);
                                dL_dscale = make_float2(dsw[0].x * Rm.c0.x + dsw[1].x * Rm.c0.y + dsw[2].x * Rm.c0.z,
                                                        dsw[0].y * Rm.c1.x + dsw[1].y * Rm.c1.y + dsw[2].y * Rm.c1.z);
                                dL_dmean = f3(dsw[0].z, dsw[1].z, dsw[2].z);
                            }
                            atomicAdd(&g.dL_dscales[2 * (size_t)gidx + 0], dL_dscale.x);
                            atomicAdd(&g.dL_dscales[2 * (size_t)gidx + 1], dL_dscale.y);
                            atomicAdd(&g.dL_drotations[4 * (size_t)gidx + 0], dL_drot.x);
                            atomicAdd(&g.dL_drotations[4 * (size_t)gidx + 1], dL_drot.y);
                            atomicAdd(&g.dL_drotations[4 * (size_t)gidx + 2], dL_drot.z);
                            atomicAdd(&g.dL_drotations[4 * (size_t)gidx + 3], dL_drot.w);
                            atomicAdd(&g.dL_dmeans3D[3 * (size_t)gidx + 0], dL_dmean.x);
                            atomicAdd(&g.dL_dmeans3D[3 * (size_t)gidx + 1], dL_dmean.y);
                            atomicAdd(&g.dL_dmeans3D[3 * (size_t)gidx + 2], dL_dmean.z);
#pragma unroll
                            for (int ch = 0; ch < 3; ch++) atomicAdd(&g.dL_dcolors[3 * (size_t)gidx + ch], dL_dcolor[ch]);
                            if (use_sh) {
                                float3 dRGB = f3(clamped[0] ? 0.0f : dL_dcolor[0], clamped[1] ? 0.0f : dL_dcolor[1], clamped[2] ? 0.0f : dL_dcolor[2]);
                                float* dsh = g.dL_dshs + (size_t)gidx * a.sh_coeffs * 3;
                                for (int k = 0; k < nbasis; k++) {
                                    atomicAdd(&dsh[3 * k + 0], basis[k] * dRGB.x);
                                    atomicAdd(&dsh[3 * k + 1], basis[k] * dRGB.y);
                                    atomicAdd(&dsh[3 * k + 2], basis[k] * dRGB.z);
                                }
                                float3 gx, gy, gz;
                                sh_dir_grad(a.sh_degree, dirn, a.shs + (size_t)gidx * a.sh_coeffs * 3, gx, gy, gz);
                                dL_ddir_sh += f3(dot(gx, dRGB), dot(gy, dRGB), dot(gz, dRGB));
                            }
                        }
                        T_prev = T_next;
                    }
                    if (stop || T_next < T_threshold || hb.cnt < TR_CHUNK || contributor > TR_MAX_INTERSECTION) break;
                    payload_dpt = dpt;
                    o_chunk = ray_ot + (payload_dpt + TR_STEP_EPSILON) * ray_dt;
                }
                if (p.stats) {
                    atomicAdd(&p.stats[0], 1ull); atomicAdd(&p.stats[1], (unsigned long long)st_nodes);
                    atomicAdd(&p.stats[2], (unsigned long long)st_prims); atomicAdd(&p.stats[3], (unsigned long long)st_hits);
                    atomicAdd(&p.stats[4], (unsigned long long)st_chunks); atomicAdd(&p.stats[5], (unsigned long long)contributor);
                }
                if (BWD) {
                    if (use_sh) g_ray_d += dnormvdv(ray_dt, dL_ddir_sh);   // backward.cu:182-184
                    continue;
                }

                // ---- forward epilogue of traceRay + tracePath bookkeeping ----
                Cc[0] += T_prev * bgc.x; Cc[1] += T_prev * bgc.y; Cc[2] += T_prev * bgc.z;
                const float A = 1.0f - T_prev;
                const float3 E = ray_ot + D * ray_dt;
                if (seg == 0) {
                    out_dpt = D; out_acc = A; out_norm = N; out_dist = f3(dist, M1, M2);
                    out_aux[0] = O[0]; out_aux[1] = O[1];
                }
                const float f_this = (seg == a.max_trace_depth) ? 1.0f : 1.0f - O[0];
                for (int ch = 0; ch < 3; ch++) out_rgb[ch] += Cc[ch] * f_this * s_prod;
                s_prod *= O[0];
                mid[MID_RAYO + 0 + seg * 16] = ray_ot.x; mid[MID_RAYO + 1 + seg * 16] = ray_ot.y; mid[MID_RAYO + 2 + seg * 16] = ray_ot.z;
                mid[MID_RAYD + 0 + seg * 16] = ray_dt.x; mid[MID_RAYD + 1 + seg * 16] = ray_dt.y; mid[MID_RAYD + 2 + seg * 16] = ray_dt.z;
                mid[MID_RGB + 0 + seg * 16] = Cc[0]; mid[MID_RGB + 1 + seg * 16] = Cc[1]; mid[MID_RGB + 2 + seg * 16] = Cc[2];
                mid[MID_DPT + seg * 16] = D; mid[MID_ACC + seg * 16] = A;
                mid[MID_NORM + 0 + seg * 16] = N.x; mid[MID_NORM + 1 + seg * 16] = N.y; mid[MID_NORM + 2 + seg * 16] = N.z;
                mid[MID_AUX + 0 + seg * 16] = O[0]; mid[MID_AUX + 1 + seg * 16] = O[1];
                if (length3(N) == 0.0f || O[0] <= a.specular_threshold) break;
                const float3 n = normalize3(N);
                ray_dt = ray_dt - 2.0f * dot(n, ray_dt) * n;
                ray_ot = E + TR_STEP_EPSILON * ray_dt;
            }

            if (!BWD) {
                const envgs_trace_out& o = p.o;
                o.out_rgb[tidx * 3 + 0] = out_rgb[0]; o.out_rgb[tidx * 3 + 1] = out_rgb[1]; o.out_rgb[tidx * 3 + 2] = out_rgb[2];
                o.out_dpt[tidx] = out_dpt; o.out_acc[tidx] = out_acc;
                o.out_norm[tidx * 3 + 0] = out_norm.x; o.out_norm[tidx * 3 + 1] = out_norm.y; o.out_norm[tidx * 3 + 2] = out_norm.z;
                o.out_dist[tidx * 3 + 0] = out_dist.x; o.out_dist[tidx * 3 + 1] = out_dist.y; o.out_dist[tidx * 3 + 2] = out_dist.z;
                o.out_aux[tidx * 2 + 0] = out_aux[0]; o.out_aux[tidx * 2 + 1] = out_aux[1];
                float4* m4 = reinterpret_cast<float4*>(o.mid_val + tidx * 32);
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (q < 4 || !p.mid_half) m4[q] = make_float4(mid[4 * q], mid[4 * q + 1], mid[4 * q + 2], mid[4 * q + 3]);
            } else {
                const envgs_trace_grads& g = p.g;
                g.dL_dray_o[tidx * 3 + 0] = g_ray_o.x; g.dL_dray_o[tidx * 3 + 1] = g_ray_o.y; g.dL_dray_o[tidx * 3 + 2] = g_ray_o.z;
                g.dL_dray_d[tidx * 3 + 0] = g_ray_d.x; g.dL_dray_d[tidx * 3 + 1] = g_ray_d.y; g.dL_dray_d[tidx * 3 + 2] = g_ray_d.z;
            }
        }
        __syncwarp();
    }
}

int launch_trace(bool bwd, Lbvh& bvh, unsigned int* counter, unsigned long long* stats, const envgs_trace_args& a, const envgs_trace_out* o,
                 const envgs_trace_grads* g, const int* ray_list, const unsigned int* ray_count, cudaStream_t stream, int mid_half) {
    // launch geometry is a property of the CURRENT device: cached per device id, not once per process
    static int sm_counts[64] = {0}, occ_fs[64] = {0}, occ_bs[64] = {0};
    int dev = 0;
    ENVGS_CUDA_TRY(cudaGetDevice(&dev));
    const int slot = dev & 63;
    if (sm_counts[slot] == 0) {
        int n = 0, f = 0, b = 0;
        ENVGS_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        ENVGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&f, trace_kernel<false>, TR_THREADS, 0));
        ENVGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, trace_kernel<true>, TR_THREADS, 0));
        occ_fs[slot] = f < 1 ? 1 : f;
        occ_bs[slot] = b < 1 ? 1 : b;
        sm_counts[slot] = n;
    }
    const int sm_count = sm_counts[slot], occ_f = occ_fs[slot], occ_b = occ_bs[slot];
    TraceParams p;
    p.nodes = bvh.nodes; p.prims = bvh.prims; p.a = a;
    if (o) p.o = *o; else memset(&p.o, 0, sizeof(p.o));
    if (g) p.g = *g; else memset(&p.g, 0, sizeof(p.g));
    p.work_counter = counter;
    p.stats = stats;
    p.ray_list = ray_list;
    p.ray_count = ray_count;
    p.mid_half = mid_half;
    if (a.H >= 4) {
        p.tiles_x = (a.W + 7) / 8;
        p.n_items = p.tiles_x * ((a.H + 3) / 4);
    } else {
        p.tiles_x = 0;
        p.n_items = (int)(((long long)a.H * a.W + 31) / 32);
    }
    ENVGS_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
    int blocks = sm_count * (bwd ? occ_b : occ_f);
    int need = (p.n_items + (TR_THREADS / 32) - 1) / (TR_THREADS / 32);
    if (blocks > need) blocks = need > 0 ? need : 1;
    if (ray_list != nullptr) blocks = sm_count;   // count is only known on the device; idle CTAs exit at once
    ENVGS_TIMED(ray_list ? (bwd ? "trace_bwd_fallback" : "trace_fwd_fallback") : (bwd ? "trace_bwd" : "trace_fwd"), stream);
    if (bwd) trace_kernel<true><<<blocks, TR_THREADS, 0, stream>>>(p);
    else trace_kernel<false><<<blocks, TR_THREADS, 0, stream>>>(p);
    ENVGS_LAUNCH_CHECK(bwd ? "trace_bwd" : "trace_fwd", false, stream);
    return ENVGS_OK;
}

int trace_forward_impl(Lbvh& bvh, unsigned int* counter, unsigned long long* stats, const envgs_trace_args& a, const envgs_trace_out& o, cudaStream_t stream) {
    return launch_trace(false, bvh, counter, stats, a, &o, nullptr, nullptr, nullptr, stream, 0);
}
int trace_backward_impl(Lbvh& bvh, unsigned int* counter, unsigned long long* stats, const envgs_trace_args& a, const envgs_trace_grads& g, cudaStream_t stream) {
    return launch_trace(true, bvh, counter, stats, a, nullptr, &g, nullptr, nullptr, stream, 0);
}

}  // namespace envgs
