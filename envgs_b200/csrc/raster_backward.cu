// raster_backward.cu — backward pass of the surfel rasterizer for sm_100a.
//
//   render_bwd_kernel      per tile, back-to-front re-walk of the instance list (RAST5/cuda_rasterizer/backward.cu:143-440).
//                          Per-(pixel,surfel) partial gradients are reduced across the warp's 32 pixels (split-reduce) and
//                          each component leaves as ONE fire-and-forget global reduction per (warp, surfel) instead of
//                          the reference's per-pixel atomics.  (Staging the 8 warps' sums in shared memory first cost
//                          more than it saved: shared fp32 atomics are CAS loops on sm_100.)
//   preprocess_bwd_kernel  per surfel chain rule dL/dT, dL/dnormal, dL/dmean2D -> mean/scale/quat (+SH)
//                          (backward.cu:443-630), including the densification-signal overwrite (:626-629).
#include "raster_state.cuh"
#include "glm_order.cuh"
#include "tma.cuh"

namespace envgs {

#ifndef RBWD_BATCH
#define RBWD_BATCH 256
#endif
constexpr int BB = RBWD_BATCH;   // surfels staged per batch (<= CTA size 256)
constexpr int NG_FIX = 15;       // reduced values besides colour: normal(3) opacity(1) dT(9) dmean2D(2)

#ifndef RBWD_MINBLOCKS
#define RBWD_MINBLOCKS 2
#endif
template <int C>
__device__ __forceinline__ void issue_stage_bwd(TileStage<C>* st, uint64_t* bar, const BinningState& b, int first, int nb) {
    // same copy set as the forward kernel's issue_stage (raster_forward.cu); entries [first, first + nb) of the tile list
    const int a0 = first & ~3;
    const uint32_t cnt = (uint32_t)(((first & 3) + nb + 3) & ~3);
    mbar_arrive_expect_tx(bar, cnt * (uint32_t)(64 + 8 + 4 + 4 * C));
    bulk_g2s(st->r0, b.srec[0] + a0, cnt * 16u, bar);
    bulk_g2s(st->r1, b.srec[1] + a0, cnt * 16u, bar);
    bulk_g2s(st->r2, b.srec[2] + a0, cnt * 16u, bar);
    bulk_g2s(st->r3, b.srec[3] + a0, cnt * 16u, bar);
    bulk_g2s(st->box, b.sbox + a0, cnt * 8u, bar);
    bulk_g2s(st->id, b.sid + a0, cnt * 4u, bar);
#pragma unroll
    for (int ch = 0; ch < C; ch++) bulk_g2s(st->feat[ch], b.sfeat + (size_t)ch * b.cap + a0, cnt * 4u, bar);
}

template <int C>
__global__ void __launch_bounds__(256, RBWD_MINBLOCKS) render_bwd_kernel(
    const uint2* __restrict__ ranges, const BinningState b, int W, int H, int tiles_x, const float* __restrict__ bg,
    const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
    const float* __restrict__ dL_dpixels, const float* __restrict__ dL_dothers, float* __restrict__ dL_dtransMat,
    float* __restrict__ dL_dmean2D, float* __restrict__ dL_dnormal3D, float* __restrict__ dL_dopacity,
    float* __restrict__ dL_dcolors) {
    static_assert(BB == 256, "the bulk-copy staging assumes 256-entry batches (TileStage)");
    constexpr int NG = C + NG_FIX;   // per-surfel reduced values: [0,C) colour, C..C+2 normal, C+3 opacity, C+4..C+12 dT, C+13,C+14 dmean2D
    // two staging buffers filled by bulk asynchronous copies (tma.cuh): batch i+1 is in flight while batch i is processed
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileStage<C>* stages = reinterpret_cast<TileStage<C>*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * sizeof(TileStage<C>));
    __shared__ uint32_t s_max;

    const int tile = blockIdx.x;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wx0 = tx * TILE_X + (warp & 1) * 8, wy0 = ty * TILE_Y + (warp >> 1) * 4;
    const int wx1 = wx0 + 7, wy1 = wy0 + 3;
    const int pix_x = wx0 + (lane & 7), pix_y = wy0 + (lane >> 3);
    const bool inside = pix_x < W && pix_y < H;
    const size_t HW = (size_t)W * H;
    const size_t pix_id = (size_t)W * pix_y + pix_x;
    const float pxf = (float)pix_x, pyf = (float)pix_y;
    const uint2 range = ranges[tile];

    const float T_final = inside ? final_Ts[pix_id] : 0.0f;
    float T = T_final;
    const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0u;
    const uint32_t median_contributor = inside ? n_contrib[pix_id + HW] : 0u;

    if (tid == 0) {
        s_max = 0;
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init_fence();
    }
    __syncthreads();
    {
        uint32_t m = __reduce_max_sync(0xffffffffu, last_contributor);
        if (lane == 0) atomicMax(&s_max, m);
    }
    __syncthreads();
    const int total = (int)s_max;   // entries beyond the tile-wide last contributor are never visited
    if (total == 0) return;
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last_contributor);
    const int my_comp = warp_split_component<NG>(lane);   // the component whose warp sum this lane ends up with
    // destination of this lane's component: base pointer + per-surfel stride
    float* my_dst = nullptr;
    int my_stride = 0;
    {
        const int q = my_comp;
        if (q >= 0 && q < C) { my_dst = dL_dcolors + q; my_stride = C; }
        else if (q < C + 3) { my_dst = dL_dnormal3D + (q - C); my_stride = 3; }
        else if (q == C + 3) { my_dst = dL_dopacity; my_stride = 1; }
        else if (q < C + 13) { my_dst = dL_dtransMat + (q - C - 4); my_stride = 9; }
        else { my_dst = dL_dmean2D + (q - C - 13); my_stride = 3; }
    }

    float accum_rec[C], dL_dpixel[C], last_color[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) { accum_rec[ch] = 0; last_color[ch] = 0; dL_dpixel[ch] = 0; }
    float dL_dreg = 0, dL_ddepth = 0, dL_daccum = 0, dL_dn0 = 0, dL_dn1 = 0, dL_dn2 = 0, dL_dmedian_depth = 0;
    if (inside) {
#pragma unroll
        for (int ch = 0; ch < C; ch++) dL_dpixel[ch] = dL_dpixels[ch * HW + pix_id];
        dL_ddepth = dL_dothers[DEPTH_OFFSET * HW + pix_id];
        dL_daccum = dL_dothers[ALPHA_OFFSET * HW + pix_id];
        dL_dreg = dL_dothers[DISTORTION_OFFSET * HW + pix_id];
        dL_dn0 = dL_dothers[(NORMAL_OFFSET + 0) * HW + pix_id];
        dL_dn1 = dL_dothers[(NORMAL_OFFSET + 1) * HW + pix_id];
        dL_dn2 = dL_dothers[(NORMAL_OFFSET + 2) * HW + pix_id];
        dL_dmedian_depth = dL_dothers[MIDDEPTH_OFFSET * HW + pix_id];
    }
    float bg_dot_dpixel = 0.0f;
#pragma unroll
    for (int ch = 0; ch < C; ch++) bg_dot_dpixel += bg[ch] * dL_dpixel[ch];

    float last_depth = 0, accum_depth_rec = 0, accum_alpha_rec = 0, last_alpha = 0, last_dL_dT = 0;
    float ln0 = 0, ln1 = 0, ln2 = 0, an0 = 0, an1 = 0, an2 = 0;
    const float final_D = inside ? final_Ts[pix_id + HW] : 0.0f;
    const float final_D2 = inside ? final_Ts[pix_id + 2 * HW] : 0.0f;
    const float final_A = 1.0f - T_final;

    const int rounds = (total + BB - 1) / BB;
    // batch i covers the 1-based list positions pos = total - i*BB - t, t = 0..nb-1 (descending), i.e. the ascending index
    // range [range.x + total - i*BB - nb, range.x + total - i*BB) of the sorted arrays; entry t sits at lead + nb - 1 - t
    if (tid == 0) { const int nb0 = min(BB, total); issue_stage_bwd<C>(&stages[0], &bars[0], b, (int)range.x + total - nb0, nb0); }
    for (int i = 0; i < rounds; i++) {
        const int nb = min(BB, total - i * BB);
        __syncthreads();   // everyone is past batch i-1: its buffer (the one batch i+1 goes into) may be overwritten
        if (tid == 0 && i + 1 < rounds) {
            const int nbn = min(BB, total - (i + 1) * BB);
            issue_stage_bwd<C>(&stages[(i + 1) & 1], &bars[(i + 1) & 1], b, (int)range.x + total - (i + 1) * BB - nbn, nbn);
        }
        mbar_wait(&bars[i & 1], (i >> 1) & 1);
        const TileStage<C>& st = stages[i & 1];
        const int top = (((int)range.x + total - i * BB - nb) & 3) + nb - 1;   // staged slot of entry t is top - t

        // Candidate compaction: 32 staged surfels are box-tested at once (one per lane) against this warp's 8x4 pixel
        // block and the ballot is walked bit by bit, so a rejected surfel costs no loop iteration at all (63 % of the
        // (warp, surfel) pairs are rejected; the former one-by-one scan spent 20 % of the kernel's instructions on them).
        for (int c0 = 0; c0 < nb; c0 += 32) {
        bool cand = false;
        {
            const int jj = c0 + lane;
            if (jj < nb && (uint32_t)(total - i * BB - jj) <= warp_last) {
                const short4 bb = st.box[top - jj];
                cand = !(bb.x > wx1 || bb.z < wx0 || bb.y > wy1 || bb.w < wy0);
            }
        }
        unsigned int cmask = __ballot_sync(0xffffffffu, cand);
        while (cmask) {
            const int j = c0 + __ffs(cmask) - 1;
            cmask &= cmask - 1;
            const uint32_t pos = (uint32_t)(total - i * BB - j);
            const int sj = top - j;
            const float4 r0 = st.r0[sj], r1 = st.r1[sj], r2 = st.r2[sj];
            const float3 Tu = f3(r0.x, r0.y, r0.z), Tv = f3(r0.w, r1.x, r1.y), Tw = f3(r1.z, r1.w, r2.x);

            float g[32];
#pragma unroll
            for (int q = 0; q < 32; q++) g[q] = 0.0f;
            bool contrib = false;

            if (pos <= last_contributor) {
                float3 k = pxf * Tw - Tu;
                float3 l = pyf * Tw - Tv;
                float3 p = cross(k, l);
                float dx = r2.y - pxf, dy = r2.z - pyf;
                float rho2d = FILTER_INV_SQUARE * (dx * dx + dy * dy);
                const float tau = st.r3[sj].w;                   // the forward kernel's early reject, same expression, same bits
                const bool reject = (p.x * p.x + p.y * p.y > tau * (p.z * p.z)) && (rho2d > tau);
                if (!reject && p.z != 0.0f) {
                    const float rz = px_rcp_v(p.z);              // value-only use below (chain rule through s = p.xy / p.z)
                    float sx = px_div(p.x, p.z, rz), sy = px_div(p.y, p.z, rz);   // the forward kernel's own division: same alpha, same branch
                    float rho3d = sx * sx + sy * sy;
                    float rho = fminf(rho3d, rho2d);
                    float c_d = (rho3d <= rho2d) ? (sx * Tw.x + sy * Tw.y) + Tw.z : Tw.z;
                    float power = -0.5f * rho;
                    if (c_d >= NEAR_N && power <= 0.0f) {
                        const float opa = r2.w;
                        const float G = px_exp(power);
                        const float alpha = fminf(ALPHA_MAX, opa * G);
                        if (alpha >= ALPHA_MIN) {
                            contrib = true;
                            const float4 r3 = st.r3[sj];
                            const float r1ma = px_rcp_v(1.0f - alpha);
                            T = T * r1ma;
                            const float w = alpha * T;
                            float dL_dalpha = 0.0f;
#pragma unroll
                            for (int ch = 0; ch < C; ch++) {
                                const float c = st.feat[ch][sj];
                                accum_rec[ch] = last_alpha * last_color[ch] + (1.0f - last_alpha) * accum_rec[ch];
                                last_color[ch] = c;
                                dL_dalpha += (c - accum_rec[ch]) * dL_dpixel[ch];
                                g[ch] = w * dL_dpixel[ch];
                            }
                            float dL_dz = 0.0f, dL_dweight = 0.0f;
                            const float rcd = px_rcp_v(c_d);
                            const float m_d = FAR_N / (FAR_N - NEAR_N) * (1.0f - NEAR_N * rcd);
                            const float dmd_dd = (FAR_N * NEAR_N) / (FAR_N - NEAR_N) * rcd * rcd;
                            if (pos == median_contributor) dL_dz += dL_dmedian_depth;
                            dL_dweight += (final_D2 + m_d * m_d * final_A - 2.0f * m_d * final_D) * dL_dreg;
                            dL_dalpha += dL_dweight - last_dL_dT;
                            last_dL_dT = dL_dweight * alpha + (1.0f - alpha) * last_dL_dT;
                            const float dL_dmd = 2.0f * (T * alpha) * (m_d * final_A - final_D) * dL_dreg;
                            dL_dz += dL_dmd * dmd_dd;

                            accum_depth_rec = last_alpha * last_depth + (1.0f - last_alpha) * accum_depth_rec;
                            last_depth = c_d;
                            dL_dalpha += (c_d - accum_depth_rec) * dL_ddepth;
                            accum_alpha_rec = last_alpha * 1.0f + (1.0f - last_alpha) * accum_alpha_rec;
                            dL_dalpha += (1.0f - accum_alpha_rec) * dL_daccum;

                            an0 = last_alpha * ln0 + (1.0f - last_alpha) * an0; ln0 = r3.x;
                            dL_dalpha += (r3.x - an0) * dL_dn0; g[C + 0] = w * dL_dn0;
                            an1 = last_alpha * ln1 + (1.0f - last_alpha) * an1; ln1 = r3.y;
                            dL_dalpha += (r3.y - an1) * dL_dn1; g[C + 1] = w * dL_dn1;
                            an2 = last_alpha * ln2 + (1.0f - last_alpha) * an2; ln2 = r3.z;
                            dL_dalpha += (r3.z - an2) * dL_dn2; g[C + 2] = w * dL_dn2;

                            dL_dalpha *= T;
                            last_alpha = alpha;
                            dL_dalpha += (-T_final * r1ma) * bg_dot_dpixel;

                            const float dL_dG = opa * dL_dalpha;   // alpha cap is straight-through (backward.cu:391)
                            dL_dz += w * dL_ddepth;
                            g[C + 3] = G * dL_dalpha;
                            if (rho3d <= rho2d) {
                                const float dLdsx = dL_dG * -G * sx + dL_dz * Tw.x;
                                const float dLdsy = dL_dG * -G * sy + dL_dz * Tw.y;
                                const float dsx_pz = dLdsx * rz, dsy_pz = dLdsy * rz;
                                const float3 dL_dp = f3(dsx_pz, dsy_pz, -(dsx_pz * sx + dsy_pz * sy));
                                const float3 dL_dk = cross(l, dL_dp);
                                const float3 dL_dl = cross(dL_dp, k);
                                g[C + 4] = -dL_dk.x; g[C + 5] = -dL_dk.y; g[C + 6] = -dL_dk.z;
                                g[C + 7] = -dL_dl.x; g[C + 8] = -dL_dl.y; g[C + 9] = -dL_dl.z;
                                g[C + 10] = pxf * dL_dk.x + pyf * dL_dl.x + dL_dz * sx;
                                g[C + 11] = pxf * dL_dk.y + pyf * dL_dl.y + dL_dz * sy;
                                g[C + 12] = pxf * dL_dk.z + pyf * dL_dl.z + dL_dz * 1.0f;
                            } else {
                                const float dG_ddelx = -G * FILTER_INV_SQUARE * dx;
                                const float dG_ddely = -G * FILTER_INV_SQUARE * dy;
                                g[C + 13] = dL_dG * dG_ddelx;
                                g[C + 14] = dL_dG * dG_ddely;
                                g[C + 12] = dL_dz;
                            }
                        }
                    }
                }
            }
            const unsigned int cb = __ballot_sync(0xffffffffu, contrib);
            if (cb == 0u) continue;
            // split-reduce: each component's warp sum lands in one lane, which adds it to the surfel's gradient row (REDG.ADD.F32)
            const float tot = warp_split_reduce<NG>(g, lane);
            if (my_comp >= 0 && tot != 0.0f) atomicAdd(my_dst + (size_t)st.id[sj] * my_stride, tot);
        }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) preprocess_bwd_kernel(
    int P, int D, int M, const float* __restrict__ means3D, const int* __restrict__ radii,
    const float* __restrict__ shs, const uint8_t* __restrict__ clamped, const float2* __restrict__ scales,
    const float4* __restrict__ rotations, const float* __restrict__ transMat_precomp, const float4* __restrict__ rec,
    const float* __restrict__ viewmatrix, const float* __restrict__ pm, float focal_x, float focal_y, float tan_fovx,
    float tan_fovy, const float* __restrict__ campos, float* __restrict__ dL_dtransMats,
    const float* __restrict__ dL_dnormal3Ds, const float* __restrict__ dL_dcolors, float* __restrict__ dL_dshs,
    float* __restrict__ dL_dmean2Ds, float* __restrict__ dL_dmean3Ds, float* __restrict__ dL_dscales,
    float* __restrict__ dL_drots) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P || !(radii[idx] > 0)) return;
    // the reference recomputes the image size from the focal length (backward.cu:607-608)
    const int W = int(focal_x * tan_fovx * 2);
    const int H = int(focal_y * tan_fovy * 2);

    float* dT = dL_dtransMats + 9 * (size_t)idx;
    const float dT_render2 = dT[2], dT_render5 = dT[5];   // densification signal uses the render-only part
    float3 gTu = f3(dT[0], dT[1], dT[2]), gTv = f3(dT[3], dT[4], dT[5]), gTw = f3(dT[6], dT[7], dT[8]);

    float3 Tu, Tv, Tw, normal = f3(0, 0, 0);
    float3 p_orig = ld3(means3D, idx);
    float2 scale = make_float2(0, 0);
    float4 rot = make_float4(1, 0, 0, 0);
    Rot3 R;
    GMat<3, 4> Pm;   // P = world2ndc * ndc2pix (backward.cu:496), reference GLM operation order (glm_order.cuh)
    const bool precomp = (scales == nullptr);
    if (precomp) {
        const float* t = transMat_precomp + 9 * (size_t)idx;
        Tu = f3(t[0], t[1], t[2]); Tv = f3(t[3], t[4], t[5]); Tw = f3(t[6], t[7], t[8]);
    } else {
        scale = scales[idx];
        rot = rotations[idx];
        R = quat_to_rot(rot);
        GMat<3, 3> Rm = {{{R.c0.x, R.c0.y, R.c0.z}, {R.c1.x, R.c1.y, R.c1.z}, {R.c2.x, R.c2.y, R.c2.z}}};
        GMat<3, 3> S = {{{1.0f * scale.x, 0.0f, 0.0f}, {0.0f, 1.0f * scale.y, 0.0f}, {0.0f, 0.0f, 1.0f}}};   // scale_modifier ignored (backward.cu:481)
        GMat<3, 3> L = gmul<3, 3, 3>(Rm, S);
        GMat<3, 4> M = {{{L.v[0][0], L.v[0][1], L.v[0][2], 0.0f}, {L.v[1][0], L.v[1][1], L.v[1][2], 0.0f},
                         {p_orig.x, p_orig.y, p_orig.z, 1.0f}}};
        GMat<4, 4> world2ndc = {{{pm[0], pm[4], pm[8], pm[12]}, {pm[1], pm[5], pm[9], pm[13]}, {pm[2], pm[6], pm[10], pm[14]},
                                 {pm[3], pm[7], pm[11], pm[15]}}};
        GMat<3, 4> ndc2pix = {{{float(W) / 2.0f, 0.0f, 0.0f, float(W - 1) / 2.0f},
                               {0.0f, float(H) / 2.0f, 0.0f, float(H - 1) / 2.0f}, {0.0f, 0.0f, 0.0f, 1.0f}}};
        Pm = gmul<4, 4, 3>(world2ndc, ndc2pix);
        GMat<3, 3> T = gmul<4, 3, 3>(gtranspose(M), Pm);
        Tu = f3(T.v[0][0], T.v[0][1], T.v[0][2]);
        Tv = f3(T.v[1][0], T.v[1][1], T.v[1][2]);
        Tw = f3(T.v[2][0], T.v[2][1], T.v[2][2]);
        normal = xform_vec_4x3(f3(L.v[2][0], L.v[2][1], L.v[2][2]), viewmatrix);
    }

    const float m2x = dL_dmean2Ds[3 * (size_t)idx + 0], m2y = dL_dmean2Ds[3 * (size_t)idx + 1];
    if (m2x != 0.0f || m2y != 0.0f) {   // backward.cu:515-543
        float3 t_vec = f3(9.0f, 9.0f, -1.0f);
        float d = dot(t_vec, Tw * Tw);
        float3 f_vec = t_vec * (1.0f / d);
        float3 dT0 = m2x * f_vec * Tw;
        float3 dT1 = m2y * f_vec * Tw;
        float3 dT3 = m2x * f_vec * Tu + m2y * f_vec * Tv;
        float3 dL_df = m2x * Tu * Tw + m2y * Tv * Tw;
        float dL_dd = (float)((double)dot(dL_df, f_vec) * (-1.0 / (double)d));   // the reference's `-1.0 / d` is a double expression (backward.cu:525)
        float3 dd_dT3 = t_vec * Tw * 2.0f;
        dT3 += dL_dd * dd_dT3;
        gTu += dT0; gTv += dT1; gTw += dT3;
        if (precomp) {
            dT[0] = gTu.x; dT[1] = gTu.y; dT[2] = gTu.z; dT[3] = gTv.x; dT[4] = gTv.y; dT[5] = gTv.z;
            dT[6] = gTw.x; dT[7] = gTw.y; dT[8] = gTw.z;
        }
    }

    if (!precomp) {
        // dL_dM = P * transpose(dL_dT) (backward.cu:547): columns d/dL0, d/dL1, d/d(p,1)
        GMat<3, 3> dLdT = {{{gTu.x, gTu.y, gTu.z}, {gTv.x, gTv.y, gTv.z}, {gTw.x, gTw.y, gTw.z}}};
        GMat<3, 4> dM = gmul<3, 4, 3>(Pm, gtranspose(dLdT));
        float dS[4][3];
#pragma unroll
        for (int r = 0; r < 4; r++) { dS[r][0] = dM.v[0][r]; dS[r][1] = dM.v[1][r]; dS[r][2] = dM.v[2][r]; }
        float3 dL_dtn = xform_vec_4x3_T(ld3(dL_dnormal3Ds, idx), viewmatrix);
        float3 p_view = xform_point_4x3(p_orig, viewmatrix);
        float cosv = -sum3(p_view * normal);
        float mult = cosv > 0.0f ? 1.0f : -1.0f;
        dL_dtn = mult * dL_dtn;
        float3 dRS0 = f3(dS[0][0], dS[1][0], dS[2][0]), dRS1 = f3(dS[0][1], dS[1][1], dS[2][1]);
        Rot3 dR;
        dR.c0 = dRS0 * scale.x; dR.c1 = dRS1 * scale.y; dR.c2 = dL_dtn;
        float4 dq = quat_to_rot_vjp(rot, dR);
        dL_drots[4 * (size_t)idx + 0] = dq.x; dL_drots[4 * (size_t)idx + 1] = dq.y;
        dL_drots[4 * (size_t)idx + 2] = dq.z; dL_drots[4 * (size_t)idx + 3] = dq.w;
        dL_dscales[2 * (size_t)idx + 0] = dot(dRS0, R.c0);
        dL_dscales[2 * (size_t)idx + 1] = dot(dRS1, R.c1);
        dL_dmean3Ds[3 * (size_t)idx + 0] = dS[0][2];
        dL_dmean3Ds[3 * (size_t)idx + 1] = dS[1][2];
        dL_dmean3Ds[3 * (size_t)idx + 2] = dS[2][2];
    }

    if (shs != nullptr) {   // SH backward (backward.cu:20-139); C == 3
        float3 cam = f3(campos[0], campos[1], campos[2]);
        float3 dir_orig = p_orig - cam;
        float3 dir = dir_orig / length3(dir_orig);
        const float* sh = shs + (size_t)idx * M * 3;
        float3 dRGB = ld3(dL_dcolors, idx);
        dRGB.x *= clamped[3 * idx + 0] ? 0.0f : 1.0f;
        dRGB.y *= clamped[3 * idx + 1] ? 0.0f : 1.0f;
        dRGB.z *= clamped[3 * idx + 2] ? 0.0f : 1.0f;
        float b[16];
        sh_basis(D, dir, b);
        float* dsh = dL_dshs + (size_t)idx * M * 3;
        int nb = (D + 1) * (D + 1);
        for (int k = 0; k < nb; k++) { dsh[3 * k] = b[k] * dRGB.x; dsh[3 * k + 1] = b[k] * dRGB.y; dsh[3 * k + 2] = b[k] * dRGB.z; }
        float3 gx, gy, gz;
        sh_dir_grad(D, dir, sh, gx, gy, gz);
        float3 dL_ddir = f3(dot(gx, dRGB), dot(gy, dRGB), dot(gz, dRGB));
        float3 dm = dnormvdv(dir_orig, dL_ddir);
        dL_dmean3Ds[3 * (size_t)idx + 0] += dm.x;
        dL_dmean3Ds[3 * (size_t)idx + 1] += dm.y;
        dL_dmean3Ds[3 * (size_t)idx + 2] += dm.z;
    }

    // densification signal (backward.cu:626-629): depth = forward transMat[8] = Tw.z of the stored record
    float depth = precomp ? transMat_precomp[9 * (size_t)idx + 8] : rec[4 * (size_t)idx + 2].x;
    // with a precomputed transMat the reference has already folded the centre term into dL_dTs at this point (:532-543)
    const bool folded = precomp && (m2x != 0.0f || m2y != 0.0f);
    dL_dmean2Ds[3 * (size_t)idx + 0] = (folded ? gTu.z : dT_render2) * depth * 0.5f * float(W);
    dL_dmean2Ds[3 * (size_t)idx + 1] = (folded ? gTv.z : dT_render5) * depth * 0.5f * float(H);
}

template <int C>
static int raster_backward_impl(const envgs_raster_args& a, const int* radii_in, int R, char* geom_buffer,
                                char* binning_buffer, char* img_buffer, const envgs_raster_grads& gr,
                                cudaStream_t stream) {
    const int P = a.P, W = a.width, H = a.height;
    const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
    const bool debug = a.debug != 0;
    GeomState g;
    GeomState::carve(geom_buffer, P, &g, scan_temp_bytes(P));
    ImageState im;
    ImageState::carve(img_buffer, W, H, &im);
    BinningState b;
    BinningState::carve(binning_buffer, R, &b, sort_temp_bytes(R), C);
    const int* radii = radii_in ? radii_in : g.internal_radii;
    const float focal_y = H / (2.0f * a.tan_fovy), focal_x = W / (2.0f * a.tan_fovx);
    const float* feat = a.colors_precomp ? a.colors_precomp : g.rgb;

    if (R > 0) {
        constexpr size_t smem = 2 * sizeof(TileStage<C>) + 2 * sizeof(uint64_t);
        static bool attr_set = false;
        if (!attr_set) {
            ENVGS_CUDA_TRY(cudaFuncSetAttribute(render_bwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        ENVGS_TIMED("render_bwd", stream);
        render_bwd_kernel<C><<<gx * gy, 256, smem, stream>>>(im.ranges, b, W, H, gx, a.background, im.final_T, im.n_contrib,
                                                             gr.dL_dout_color, gr.dL_dout_others, gr.dL_dtransMat, gr.dL_dmeans2D,
                                                             gr.dL_dnormal, gr.dL_dopacity, gr.dL_dcolors);
        ENVGS_LAUNCH_CHECK("render_bwd", debug, stream);
    }
    ENVGS_TIMED("preprocess_bwd", stream);
    preprocess_bwd_kernel<C><<<(P + 255) / 256, 256, 0, stream>>>(
        P, a.sh_degree, a.sh_coeffs, a.means3D, radii, a.shs, g.clamped, (const float2*)a.scales,
        (const float4*)a.rotations, a.transMat_precomp, g.rec, a.viewmatrix, a.projmatrix, focal_x, focal_y, a.tan_fovx,
        a.tan_fovy, a.campos, gr.dL_dtransMat, gr.dL_dnormal, gr.dL_dcolors, gr.dL_dsh, gr.dL_dmeans2D, gr.dL_dmeans3D,
        gr.dL_dscales, gr.dL_drotations);
    ENVGS_LAUNCH_CHECK("preprocess_bwd", debug, stream);
    return ENVGS_OK;
}

int raster_backward_dispatch(const envgs_raster_args& a, const int* radii, int R, void* geom, void* binning, void* img,
                             const envgs_raster_grads& gr, cudaStream_t stream) {
    switch (a.C) {
        case 3: return raster_backward_impl<3>(a, radii, R, (char*)geom, (char*)binning, (char*)img, gr, stream);
        case 5: return raster_backward_impl<5>(a, radii, R, (char*)geom, (char*)binning, (char*)img, gr, stream);
        case 7: return raster_backward_impl<7>(a, radii, R, (char*)geom, (char*)binning, (char*)img, gr, stream);
        default: set_error("unsupported channel count C=%d (3, 5 or 7)", a.C); return ENVGS_ERR_INVALID;
    }
}

}  // namespace envgs
