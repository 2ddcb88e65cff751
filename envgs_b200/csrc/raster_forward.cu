// raster_forward.cu — forward pass of the 2D-surfel tile rasterizer, hand-written for sm_100a.
//
// Pipeline (one CUDA stream, no host sync except the 4-byte instance-count read-back):
//   preprocess_kernel  -> per surfel: homography T, normal, projected centre, 3-sigma radius (reference
//                         semantics, RAST5/cuda_rasterizer/forward.cu:75-251) + exact opacity-aware tile box
//   cub InclusiveSum   -> instance offsets            (library primitive, same as rasterizer_impl.cu:279)
//   duplicate_kernel   -> (tile | depth-bits) keys    (rasterizer_impl.cu:70-111)
//   cub SortPairs      -> per-tile depth order        (library primitive, rasterizer_impl.cu:305-310)
//   tile_ranges_kernel -> [start,end) per tile        (rasterizer_impl.cu:116-138)
//   render_fwd_kernel  -> per-pixel front-to-back compositing (forward.cu:256-445)
#include <cub/cub.cuh>
#include "raster_state.cuh"
#include "glm_order.cuh"
#include "tma.cuh"

namespace envgs {

// ------------------------------------------------------------------------------------------------------
// state carving
// ------------------------------------------------------------------------------------------------------
size_t scan_temp_bytes(int P) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, P);
    return bytes;
}
size_t sort_temp_bytes(int N) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, N);
    return bytes;
}

size_t GeomState::carve(char* base, int P, GeomState* s, size_t scan_bytes) {
    char* p = base;
    GeomState tmp;
    GeomState& g = s ? *s : tmp;
    obtain(p, g.rec, (size_t)4 * P);
    obtain(p, g.pixbox, (size_t)P);
    obtain(p, g.tilebox, (size_t)P);
    obtain(p, g.depths, (size_t)P);
    obtain(p, g.tiles_touched, (size_t)P);
    obtain(p, g.point_offsets, (size_t)P);
    obtain(p, g.rgb, (size_t)3 * P);
    obtain(p, g.clamped, (size_t)3 * P);
    obtain(p, g.internal_radii, (size_t)P);
    obtain(p, g.scan_temp, scan_bytes);
    g.scan_temp_bytes = scan_bytes;
    return (size_t)(p - base) + 256;
}
size_t ImageState::carve(char* base, int W, int H, ImageState* s) {
    char* p = base;
    ImageState tmp;
    ImageState& g = s ? *s : tmp;
    size_t HW = (size_t)W * H;
    size_t tiles = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
    obtain(p, g.final_T, 3 * HW);
    obtain(p, g.n_contrib, 2 * HW);
    obtain(p, g.ranges, tiles);
    return (size_t)(p - base) + 256;
}
size_t BinningState::carve(char* base, int N, BinningState* s, size_t sort_bytes, int C) {
    char* p = base;
    BinningState tmp;
    BinningState& g = s ? *s : tmp;
    obtain(p, g.keys_unsorted, (size_t)N);
    obtain(p, g.keys, (size_t)N);
    obtain(p, g.vals_unsorted, (size_t)N);
    obtain(p, g.point_list, (size_t)N);
    g.cap = ((size_t)N + 8 + 63) / 64 * 64;
    for (int k = 0; k < 4; k++) obtain(p, g.srec[k], g.cap);
    obtain(p, g.sbox, g.cap);
    obtain(p, g.sid, g.cap);
    obtain(p, g.sfeat, (size_t)C * g.cap);
    obtain(p, g.sort_temp, sort_bytes);
    g.sort_temp_bytes = sort_bytes;
    return (size_t)(p - base) + 256;
}

// ------------------------------------------------------------------------------------------------------
// preprocess
// ------------------------------------------------------------------------------------------------------

// 3x3 homography rows (Tu,Tv,Tw) = rows of N * PV * S  (forward.cu:75-115), evaluated in the reference's
// GLM operation order `T = transpose(splat2world) * world2ndc * ndc2pix` (see glm_order.cuh for why the order matters).
__device__ __forceinline__ void compute_transmat(float3 p, float2 scale, float mod, float4 rot, const float* pm,
                                                 const float* vm, int W, int H, float3& Tu, float3& Tv, float3& Tw,
                                                 float3& normal) {
    Rot3 Rq = quat_to_rot(rot);
    GMat<3, 3> R = {{{Rq.c0.x, Rq.c0.y, Rq.c0.z}, {Rq.c1.x, Rq.c1.y, Rq.c1.z}, {Rq.c2.x, Rq.c2.y, Rq.c2.z}}};
    GMat<3, 3> S = {{{mod * scale.x, 0.0f, 0.0f}, {0.0f, mod * scale.y, 0.0f}, {0.0f, 0.0f, 1.0f}}};   // scale_to_mat
    GMat<3, 3> L = gmul<3, 3, 3>(R, S);
    GMat<3, 4> splat2world = {{{L.v[0][0], L.v[0][1], L.v[0][2], 0.0f}, {L.v[1][0], L.v[1][1], L.v[1][2], 0.0f}, {p.x, p.y, p.z, 1.0f}}};
    GMat<4, 4> world2ndc = {{{pm[0], pm[4], pm[8], pm[12]}, {pm[1], pm[5], pm[9], pm[13]}, {pm[2], pm[6], pm[10], pm[14]},
                             {pm[3], pm[7], pm[11], pm[15]}}};
    GMat<3, 4> ndc2pix = {{{float(W) / 2.0f, 0.0f, 0.0f, float(W - 1) / 2.0f}, {0.0f, float(H) / 2.0f, 0.0f, float(H - 1) / 2.0f},
                           {0.0f, 0.0f, 0.0f, 1.0f}}};
    GMat<3, 3> T = gmul<4, 3, 3>(gmul<4, 3, 4>(gtranspose(splat2world), world2ndc), ndc2pix);
    Tu = f3(T.v[0][0], T.v[0][1], T.v[0][2]);
    Tv = f3(T.v[1][0], T.v[1][1], T.v[1][2]);
    Tw = f3(T.v[2][0], T.v[2][1], T.v[2][2]);
    normal = xform_vec_4x3(f3(L.v[2][0], L.v[2][1], L.v[2][2]), vm);
}

// centre + half extent of the screen-space conic {u^2+v^2 <= cutoff^2} (forward.cu:119-145).
// Written with explicit round-to-nearest intrinsics in the operation order the reference build executes (read off the SASS
// of its preprocessCUDA<5>, nvcc 12.9 -O3 sm_100a): which product of each three-term dot stays a separately rounded
// multiply and which ones are fused is otherwise at the mercy of common-subexpression sharing with neighbouring code, and
// a 1-ulp change of d moves the centre on ~20 % of the surfels and the integer radius on ~0.15 % of them.
//   d      = fma(-Tw.z, Tw.z, fma(Tw.x^2, 9, (Tw.y^2) * 9))
//   f      = (F, F, -rd),  rd = 1 / d,  F = rd * 9
//   p.x    = fma(Tw.z Tu.z, -rd, fma(F, Tw.x Tu.x, F * (Tw.y Tu.y)))
//   h0.x   = fma(p.x, p.x, fma(Tu.z^2, rd, -fma(F, Tu.x^2, F * Tu.y^2)))
#ifndef ENVGS_AABB_VARIANT
#define ENVGS_AABB_VARIANT 0
#endif
__device__ __forceinline__ float dot3_ref_order(float F, float rd, float mx, float my, float mz) {
    // F*mx + F*my + (-rd)*mz in the reference's fused order
#if ENVGS_AABB_VARIANT == 0
    return __fmaf_rn(mz, -rd, __fmaf_rn(F, mx, __fmul_rn(F, my)));
#else
    return __fmaf_rn(mz, -rd, __fmaf_rn(F, my, __fmul_rn(F, mx)));
#endif
}
__device__ __forceinline__ bool compute_aabb(float3 Tu, float3 Tv, float3 Tw, float cutoff2, float2& center,
                                             float2& extent) {
    const float wx2 = __fmul_rn(Tw.x, Tw.x), wy2 = __fmul_rn(Tw.y, Tw.y);
#if ENVGS_AABB_VARIANT == 0
    const float d = __fmaf_rn(-Tw.z, Tw.z, __fmaf_rn(wx2, cutoff2, __fmul_rn(wy2, cutoff2)));
#else
    const float d = __fmaf_rn(-Tw.z, Tw.z, __fmaf_rn(wy2, cutoff2, __fmul_rn(wx2, cutoff2)));
#endif
    if (d == 0.0f) return false;
    const float rd = __fdiv_rn(1.0f, d);
    const float F = __fmul_rn(rd, cutoff2);
    const float px = dot3_ref_order(F, rd, __fmul_rn(Tu.x, Tw.x), __fmul_rn(Tu.y, Tw.y), __fmul_rn(Tu.z, Tw.z));
    const float py = dot3_ref_order(F, rd, __fmul_rn(Tv.x, Tw.x), __fmul_rn(Tv.y, Tw.y), __fmul_rn(Tv.z, Tw.z));
    const float du = dot3_ref_order(F, rd, __fmul_rn(Tu.x, Tu.x), __fmul_rn(Tu.y, Tu.y), __fmul_rn(Tu.z, Tu.z));
    const float dv = dot3_ref_order(F, rd, __fmul_rn(Tv.x, Tv.x), __fmul_rn(Tv.y, Tv.y), __fmul_rn(Tv.z, Tv.z));
    const float h0x = __fmaf_rn(px, px, -du), h0y = __fmaf_rn(py, py, -dv);
    center = make_float2(px, py);
    extent = make_float2(sqrtf(fmaxf(1e-4f, h0x)), sqrtf(fmaxf(1e-4f, h0y)));
    return true;
}

// reference tile rectangle (auxiliary.h:67-77)
__device__ __forceinline__ void get_rect(float2 p, int max_radius, uint2& rmin, uint2& rmax, int gx, int gy) {
    rmin.x = min(gx, max(0, (int)((p.x - max_radius) / TILE_X)));
    rmin.y = min(gy, max(0, (int)((p.y - max_radius) / TILE_Y)));
    rmax.x = min(gx, max(0, (int)((p.x + max_radius + TILE_X - 1) / TILE_X)));
    rmax.y = min(gy, max(0, (int)((p.y + max_radius + TILE_Y - 1) / TILE_Y)));
}

// Inclusive pixel box of the region where this surfel can reach alpha >= 1/255:
//   { min(rho3d, rho2d) <= tau2 },  tau2 = 2 ln(255 o)   (forward.cu:361-382)
// = conic {u^2+v^2 <= tau2}  U  disc {|pix - c|^2 <= tau2 / 2}.  Conservative (margins) and only ever
// removes (pixel, surfel) pairs the reference `continue`s on, so rendered values are unchanged.
// Returns false when the surfel cannot contribute anywhere.
__device__ __forceinline__ bool contributing_pixbox(float3 Tu, float3 Tv, float3 Tw, float2 c, float opacity,
                                                    float& x0, float& y0, float& x1, float& y1) {
    const float BIG = 30000.0f;
    x0 = y0 = -BIG; x1 = y1 = BIG;
    float a = 255.0f * opacity;
    if (!(a >= 1.0f)) return false;                 // alpha <= o < 1/255 everywhere (also rejects NaN)
    float tau2 = 2.0f * __logf(a) + 1e-3f;           // +margin for the approximate log
    // translate pixel coordinates so the projected centre is the origin: better conditioned than forward.cu:119-145
    float3 U = Tu - c.x * Tw, V = Tv - c.y * Tw;
    float3 t = f3(tau2, tau2, -1.0f);
    float d = dot(t, Tw * Tw);
    // ellipse only if d < 0 with head-room; otherwise (hyperbola / grazing) keep the reference rect
    if (!(d < -1e-2f * Tw.z * Tw.z)) return true;
    float3 f = (1.0f / d) * t;
    float px = dot(f, U * Tw), py = dot(f, V * Tw);
    float hx = px * px - dot(f, U * U), hy = py * py - dot(f, V * V);
    if (!(hx >= 0.0f) || !(hy >= 0.0f) || !isfinite(px) || !isfinite(py)) return true;
    hx = sqrtf(hx); hy = sqrtf(hy);
    float r2d = sqrtf(0.5f * tau2);
    float lo_x = fminf(px - hx, -r2d), hi_x = fmaxf(px + hx, r2d);
    float lo_y = fminf(py - hy, -r2d), hi_y = fmaxf(py + hy, r2d);
    float mx = 0.02f + 1e-3f * (hi_x - lo_x), my = 0.02f + 1e-3f * (hi_y - lo_y);
    x0 = fmaxf(c.x + lo_x - mx, -BIG); x1 = fminf(c.x + hi_x + mx, BIG);
    y0 = fmaxf(c.y + lo_y - my, -BIG); y1 = fminf(c.y + hi_y + my, BIG);
    if (!(x0 <= x1) || !(y0 <= y1)) { x0 = y0 = -BIG; x1 = y1 = BIG; }   // NaN guard: no tightening
    return true;
}

template <int C>
__global__ void __launch_bounds__(256) preprocess_kernel(
    int P, int D, int M, const float* __restrict__ means3D, const float2* __restrict__ scales, float scale_modifier,
    const float4* __restrict__ rotations, const float* __restrict__ opacities, const float* __restrict__ shs,
    const float* __restrict__ transMat_precomp, const float* __restrict__ colors_precomp,
    const float* __restrict__ viewmatrix, const float* __restrict__ projmatrix, const float* __restrict__ campos, int W,
    int H, int gx, int gy, int* __restrict__ radii, GeomState g) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    radii[idx] = 0;
    g.tiles_touched[idx] = 0;

    float3 p_orig = ld3(means3D, idx);
    float3 p_view = xform_point_4x3(p_orig, viewmatrix);
    if (p_view.z <= 0.2f) return;   // in_frustum, auxiliary.h:185-210

    float3 Tu, Tv, Tw, normal;
    if (transMat_precomp == nullptr) {
        compute_transmat(p_orig, scales[idx], scale_modifier, rotations[idx], projmatrix, viewmatrix, W, H, Tu, Tv, Tw,
                         normal);
    } else {
        const float* t = transMat_precomp + 9 * (size_t)idx;
        Tu = f3(t[0], t[1], t[2]); Tv = f3(t[3], t[4], t[5]); Tw = f3(t[6], t[7], t[8]);
        normal = f3(0.0f, 0.0f, 1.0f);
    }
    // DUAL_VISIABLE (forward.cu:209-214)
    float cosv = -sum3(p_view * normal);
    if (cosv == 0.0f) return;
    float mult = cosv > 0.0f ? 1.0f : -1.0f;
    normal = mult * normal;

    float2 center, extent;
    if (!compute_aabb(Tu, Tv, Tw, 9.0f, center, extent)) return;
    float radius = ceilf(fmaxf(fmaxf(extent.x, extent.y), 3.0f * FILTER_SIZE));
    uint2 rmin, rmax;
    get_rect(center, (int)radius, rmin, rmax, gx, gy);
    if ((rmax.x - rmin.x) * (rmax.y - rmin.y) == 0) return;

    if (colors_precomp == nullptr) {   // in-kernel SH (C == 3 only; checked on the host)
        float3 dir = p_orig - f3(campos[0], campos[1], campos[2]);
        dir = dir / length3(dir);
        float b[16];
        sh_basis(D, dir, b);
        const float* sh = shs + (size_t)idx * M * 3;
        float3 res = f3(0, 0, 0);
        int nb = (D + 1) * (D + 1);
        for (int k = 0; k < nb; k++) res += b[k] * f3(sh[3 * k], sh[3 * k + 1], sh[3 * k + 2]);
        res.x += 0.5f; res.y += 0.5f; res.z += 0.5f;
        g.clamped[3 * idx + 0] = res.x < 0; g.clamped[3 * idx + 1] = res.y < 0; g.clamped[3 * idx + 2] = res.z < 0;
        g.rgb[3 * idx + 0] = fmaxf(res.x, 0.0f); g.rgb[3 * idx + 1] = fmaxf(res.y, 0.0f); g.rgb[3 * idx + 2] = fmaxf(res.z, 0.0f);
    }

    float opacity = opacities[idx];
    g.depths[idx] = p_view.z;
    radii[idx] = (int)radius;
    g.rec[4 * (size_t)idx + 0] = make_float4(Tu.x, Tu.y, Tu.z, Tv.x);
    g.rec[4 * (size_t)idx + 1] = make_float4(Tv.y, Tv.z, Tw.x, Tw.y);
    g.rec[4 * (size_t)idx + 2] = make_float4(Tw.z, center.x, center.y, opacity);
    // .w: conservative threshold of the tile kernels' early reject.  A pair contributes iff min(rho3d, rho2d) <= 2 ln(255 o);
    // the margin (0.2 % + 4e-3) is three orders above the fp32 noise of either side's evaluation, so the reject only ever drops
    // pairs the exact test `alpha < 1/255` drops (forward.cu:380-382) and every borderline pair still takes the exact path.
    g.rec[4 * (size_t)idx + 3] = make_float4(normal.x, normal.y, normal.z, 2.0f * logf(255.0f * opacity) * 1.002f + 4e-3f);

    // exact culling: intersect the reference rect with the tiles the contributing region can touch
    float bx0, by0, bx1, by1;
    if (!contributing_pixbox(Tu, Tv, Tw, center, opacity, bx0, by0, bx1, by1)) return;   // radii stays set (reference semantics)
    int px0 = (int)ceilf(bx0), px1 = (int)floorf(bx1), py0 = (int)ceilf(by0), py1 = (int)floorf(by1);
    // pixels actually covered by the reference rect
    px0 = max(px0, (int)rmin.x * TILE_X); px1 = min(px1, min((int)rmax.x * TILE_X, W) - 1);
    py0 = max(py0, (int)rmin.y * TILE_Y); py1 = min(py1, min((int)rmax.y * TILE_Y, H) - 1);
    if (px0 > px1 || py0 > py1) return;
    int tx0 = px0 / TILE_X, tx1 = px1 / TILE_X + 1, ty0 = py0 / TILE_Y, ty1 = py1 / TILE_Y + 1;
    g.pixbox[idx] = make_short4((short)px0, (short)py0, (short)px1, (short)py1);
    g.tilebox[idx] = make_ushort4((unsigned short)tx0, (unsigned short)ty0, (unsigned short)tx1, (unsigned short)ty1);
    g.tiles_touched[idx] = (uint32_t)((tx1 - tx0) * (ty1 - ty0));
}

// ------------------------------------------------------------------------------------------------------
// binning
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) duplicate_kernel(int P, GeomState g, uint64_t* __restrict__ keys,
                                                        uint32_t* __restrict__ vals, int gx, uint32_t cap) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    uint32_t n = g.tiles_touched[idx];
    if (n == 0) return;
    uint32_t off = (idx == 0) ? 0 : g.point_offsets[idx - 1];
    ushort4 tb = g.tilebox[idx];
    uint32_t dbits = __float_as_uint(g.depths[idx]);
    for (int y = tb.y; y < tb.w; y++)
        for (int x = tb.x; x < tb.z; x++) {
            uint64_t key = (uint64_t)(y * gx + x);
            key = (key << 32) | dbits;
            if (off < cap) {       // sync-free mode: the buffer holds `cap` instances; anything beyond is dropped and flagged
                keys[off] = key;
                vals[off] = (uint32_t)idx;
            }
            off++;
        }
}

// Sync-free mode: slots [count, cap) get the all-ones key (sorts behind every real instance) and the true count / overflow
// flag are published for the caller.
__global__ void __launch_bounds__(256) pad_keys_kernel(uint32_t cap, const uint32_t* __restrict__ count_ptr, uint64_t* __restrict__ keys,
                                                       uint32_t* __restrict__ vals, int32_t* __restrict__ count_out) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = *count_ptr;
    if (idx == 0 && count_out) { count_out[0] = (int32_t)n; count_out[1] = n > cap ? 1 : 0; }
    if (idx >= cap || idx < n) return;
    keys[idx] = ~0ull;
    vals[idx] = 0u;
}

__global__ void __launch_bounds__(256) tile_ranges_kernel(int L, const uint64_t* __restrict__ keys,
                                                          uint2* __restrict__ ranges, uint32_t n_tiles) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    uint32_t cur = (uint32_t)(keys[idx] >> 32);
    const bool cur_ok = cur < n_tiles;                  // padding entries of the sync-free mode carry tile id 0xffffffff
    if (idx == 0) { if (cur_ok) ranges[cur].x = 0; }
    else {
        uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
        if (cur != prev) { if (prev < n_tiles) ranges[prev].y = idx; if (cur_ok) ranges[cur].x = idx; }
    }
    if (idx == L - 1 && cur_ok) ranges[cur].y = L;
}

// Gather the per-instance render state into sorted structure-of-arrays order (BinningState::srec/sbox/sid/sfeat): one
// 64-byte record gather + the feature row per instance, written back fully coalesced.  render_fwd and render_bwd then both
// stream their tile's range with bulk asynchronous copies instead of each repeating the scattered gather.
template <int C>
__global__ void __launch_bounds__(256) gather_sorted_kernel(int N, const uint32_t* __restrict__ count_ptr,
                                                            const uint32_t* __restrict__ point_list,
                                                            const float4* __restrict__ rec, const short4* __restrict__ pixbox,
                                                            const float* __restrict__ features, BinningState b) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N || (count_ptr && (uint32_t)k >= *count_ptr)) return;   // sync-free mode: N is the capacity, the real count is on the device
    const int id = (int)point_list[k];
#pragma unroll
    for (int q = 0; q < 4; q++) b.srec[q][k] = rec[4 * (size_t)id + q];
    b.sbox[k] = pixbox[id];
    b.sid[k] = id;
#pragma unroll
    for (int ch = 0; ch < C; ch++) b.sfeat[(size_t)ch * b.cap + k] = features[(size_t)id * C + ch];
}

// one thread: arm the stage's barrier and issue the bulk copies of list entries [first, first + nb) (start rounded down /
// count rounded up to multiples of 4 entries so every copy is 16-byte aligned; the consumer indexes with lead = first & 3)
template <int C>
__device__ __forceinline__ void issue_stage(TileStage<C>* st, uint64_t* bar, const BinningState& b, int first, int nb) {
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

// ------------------------------------------------------------------------------------------------------
// render forward
// ------------------------------------------------------------------------------------------------------
// One CTA (256 threads) per 16x16 tile; warp w owns the 8x4 pixel block (w%2, w/2) so that a per-warp
// bounding-box reject can skip surfels that cannot reach any of its 32 pixels.  The tile's instance list is
// streamed through shared memory in batches of 256 records (64 B record + 8 B box + C features each).
constexpr int RB = 256;   // batch size == CTA size

template <int C>
__global__ void __launch_bounds__(256) render_fwd_kernel(
    const uint2* __restrict__ ranges, const BinningState b, int W, int H, int tiles_x, const float* __restrict__ bg,
    float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
    float* __restrict__ out_others, float* __restrict__ out_weight) {
    // two staging buffers filled by bulk asynchronous copies (tma.cuh): batch i+1 is in flight while batch i is composited
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileStage<C>* stages = reinterpret_cast<TileStage<C>*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * sizeof(TileStage<C>));

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
    const int total = (int)(range.y - range.x);
    const int rounds = (total + RB - 1) / RB;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init_fence();
    }
    __syncthreads();
    if (tid == 0 && rounds > 0) issue_stage<C>(&stages[0], &bars[0], b, (int)range.x, min(RB, total));

    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0, median_contributor = 0;
    float Cacc[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) Cacc[ch] = 0.0f;
    float N0 = 0, N1 = 0, N2 = 0, Dacc = 0, M1 = 0, M2 = 0, distortion = 0, median_depth = 0;

    int toDo = total;
    for (int i = 0; i < rounds; i++, toDo -= RB) {
        // everyone is past batch i-1 here, so its buffer (the one batch i+1 goes into) may be overwritten
        if (__syncthreads_count(done) == RB) {
            if (tid == 0) mbar_wait(&bars[i & 1], (i >> 1) & 1);   // batch i is still in flight: it must land before the CTA exits
            break;
        }
        if (tid == 0 && i + 1 < rounds)
            issue_stage<C>(&stages[(i + 1) & 1], &bars[(i + 1) & 1], b, (int)range.x + (i + 1) * RB, min(RB, toDo - RB));
        mbar_wait(&bars[i & 1], (i >> 1) & 1);
        const TileStage<C>& st = stages[i & 1];
        const int lead = ((int)range.x + i * RB) & 3;

        const int nb = min(RB, toDo);
        bool warp_done = __all_sync(0xffffffffu, done);
        // Candidate compaction: 32 staged surfels are box-tested at once (one per lane) against this warp's 8x4 pixel
        // block and the ballot is walked bit by bit, so a rejected surfel costs no loop iteration at all.
        for (int c0 = 0; c0 < nb && !warp_done; c0 += 32) {
        bool cand = false;
        {
            const int jj = c0 + lane;
            if (jj < nb) {
                const short4 bb = st.box[lead + jj];
                cand = !(bb.x > wx1 || bb.z < wx0 || bb.y > wy1 || bb.w < wy0);
            }
        }
        unsigned int cmask = __ballot_sync(0xffffffffu, cand);
        while (cmask && !warp_done) {
            const int j = c0 + __ffs(cmask) - 1;
            cmask &= cmask - 1;
            const int sj = lead + j;
            const float4 r0 = st.r0[sj], r1 = st.r1[sj], r2 = st.r2[sj];
            const float3 Tu = f3(r0.x, r0.y, r0.z), Tv = f3(r0.w, r1.x, r1.y), Tw = f3(r1.z, r1.w, r2.x);
            float w = 0.0f;
            bool contrib = false;
            if (!done) {
                // ray-splat intersection through two homogeneous planes (forward.cu:355-366)
                float3 k = pxf * Tw - Tu;
                float3 l = pyf * Tw - Tv;
                float3 p = cross(k, l);
                float dx = r2.y - pxf, dy = r2.z - pyf;
                float rho2d = FILTER_INV_SQUARE * (dx * dx + dy * dy);
                // Early reject without the two divisions and the exponential: rho3d = (p.x^2 + p.y^2) / p.z^2 > tau and
                // rho2d > tau  =>  alpha < 1/255.  ~80 % of the (pixel, surfel) pairs that survive the warp's box test end here.
                // (NaNs compare false and fall through to the exact path.)
                const float tau = st.r3[sj].w;
                const bool reject = (p.x * p.x + p.y * p.y > tau * (p.z * p.z)) && (rho2d > tau);
                if (!reject && p.z != 0.0f) {
                    const float rz = px_rcp(p.z);                // fast-math build: one reciprocal instead of two divisions
                    float sx = px_div(p.x, p.z, rz), sy = px_div(p.y, p.z, rz);
                    float rho3d = sx * sx + sy * sy;
                    float rho = fminf(rho3d, rho2d);
                    float depth = (rho3d <= rho2d) ? (sx * Tw.x + sy * Tw.y) + Tw.z : Tw.z;
                    float power = -0.5f * rho;
                    if (depth >= NEAR_N && power <= 0.0f) {
                        float alpha = fminf(ALPHA_MAX, r2.w * px_exp(power));
                        if (alpha >= ALPHA_MIN) {
                            float test_T = T * (1.0f - alpha);
                            if (test_T < T_STOP) {
                                done = true;
                            } else {
                                contrib = true;
                                w = alpha * T;
                                const float4 r3 = st.r3[sj];
                                float A = 1.0f - T;
                                float m = FAR_N / (FAR_N - NEAR_N) * (1.0f - NEAR_N * px_rcp_v(depth));
                                distortion += (m * m * A + M2 - 2.0f * m * M1) * w;
                                Dacc += depth * w;
                                M1 += m * w;
                                M2 += m * m * w;
                                uint32_t pos = (uint32_t)(i * RB + j + 1);
                                if (T > 0.5f) { median_depth = depth; median_contributor = pos; }
                                N0 += r3.x * w; N1 += r3.y * w; N2 += r3.z * w;
#pragma unroll
                                for (int ch = 0; ch < C; ch++) Cacc[ch] += st.feat[ch][sj] * w;
                                T = test_T;
                                last_contributor = pos;
                            }
                        }
                    }
                }
            }
            // per-surfel accumulated blend weight (forward.cu:420): warp-reduce, then one global reduction per (warp, surfel)
            if (__any_sync(0xffffffffu, contrib)) {
                float ws = warp_sum(w);
                if (lane == 0) atomicAdd(&out_weight[st.id[sj]], ws);
            }
            warp_done = __all_sync(0xffffffffu, done);
        }
        }
    }

    if (inside) {
        final_T[pix_id] = T;
        final_T[pix_id + HW] = M1;
        final_T[pix_id + 2 * HW] = M2;
        n_contrib[pix_id] = last_contributor;
        n_contrib[pix_id + HW] = median_contributor;
#pragma unroll
        for (int ch = 0; ch < C; ch++) out_color[ch * HW + pix_id] = Cacc[ch] + T * bg[ch];
        out_others[pix_id + DEPTH_OFFSET * HW] = Dacc;
        out_others[pix_id + ALPHA_OFFSET * HW] = 1.0f - T;
        out_others[pix_id + (NORMAL_OFFSET + 0) * HW] = N0;
        out_others[pix_id + (NORMAL_OFFSET + 1) * HW] = N1;
        out_others[pix_id + (NORMAL_OFFSET + 2) * HW] = N2;
        out_others[pix_id + MIDDEPTH_OFFSET * HW] = median_depth;
        out_others[pix_id + DISTORTION_OFFSET * HW] = distortion;
    }
}

// ------------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------------
static uint32_t higher_msb(uint32_t n) {   // rasterizer_impl.cu:35-50
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

template <int C>
static int raster_forward_impl(const envgs_raster_args& a, const envgs_raster_out& o, envgs_alloc_fn geom_alloc,
                               envgs_alloc_fn binning_alloc, envgs_alloc_fn img_alloc, void* user,
                               cudaStream_t stream) {
    const int P = a.P, W = a.width, H = a.height;
    const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
    const bool debug = a.debug != 0;

    size_t scan_bytes = scan_temp_bytes(P);
    size_t geom_bytes = GeomState::carve(nullptr, P, nullptr, scan_bytes);
    char* geom_ptr = (char*)geom_alloc(user, geom_bytes);
    if (!geom_ptr) { set_error("geometry allocator returned NULL for %zu bytes", geom_bytes); return ENVGS_ERR_ALLOC; }
    GeomState g;
    GeomState::carve(geom_ptr, P, &g, scan_bytes);

    size_t img_bytes = ImageState::carve(nullptr, W, H, nullptr);
    char* img_ptr = (char*)img_alloc(user, img_bytes);
    if (!img_ptr) { set_error("image allocator returned NULL for %zu bytes", img_bytes); return ENVGS_ERR_ALLOC; }
    ImageState im;
    ImageState::carve(img_ptr, W, H, &im);

    int* radii = o.radii ? o.radii : g.internal_radii;

    {
        ENVGS_TIMED("preprocess_fwd", stream);
        preprocess_kernel<C><<<(P + 255) / 256, 256, 0, stream>>>(
            P, a.sh_degree, a.sh_coeffs, a.means3D, (const float2*)a.scales, a.scale_modifier, (const float4*)a.rotations,
            a.opacities, a.shs, a.transMat_precomp, a.colors_precomp, a.viewmatrix, a.projmatrix, a.campos, W, H, gx, gy,
            radii, g);
        ENVGS_LAUNCH_CHECK("preprocess", debug, stream);
    }
    {
        ENVGS_TIMED("scan", stream);
        ENVGS_CUDA_TRY(cub::DeviceScan::InclusiveSum(g.scan_temp, g.scan_temp_bytes, g.tiles_touched, g.point_offsets, P, stream));
        count_launch(2);
    }

    const bool sync_free = a.instance_capacity > 0;
    int num_rendered = 0;
    if (sync_free) {
        num_rendered = a.instance_capacity;             // buffer size; the true count never leaves the device
    } else {
        ENVGS_CUDA_TRY(cudaMemcpyAsync(&num_rendered, g.point_offsets + P - 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
        ENVGS_CUDA_TRY(cudaStreamSynchronize(stream));
        if (num_rendered < 0) {   // the inclusive sum is a uint32: more than INT_MAX instances would read back as an error code
            set_error("%u (tile, surfel) instances exceed the 2^31 - 1 this interface can return", (unsigned)num_rendered);
            return ENVGS_ERR_INVALID;
        }
        if (a.instance_count_out) {
            const int32_t two[2] = {num_rendered, 0};
            ENVGS_CUDA_TRY(cudaMemcpyAsync(a.instance_count_out, two, sizeof(two), cudaMemcpyHostToDevice, stream));
        }
    }
    const uint32_t* count_ptr = sync_free ? g.point_offsets + P - 1 : nullptr;

    size_t sort_bytes = sort_temp_bytes(num_rendered);
    size_t bin_bytes = BinningState::carve(nullptr, num_rendered, nullptr, sort_bytes, C);
    char* bin_ptr = (char*)binning_alloc(user, bin_bytes);
    if (!bin_ptr) { set_error("binning allocator returned NULL for %zu bytes", bin_bytes); return ENVGS_ERR_ALLOC; }
    BinningState b;
    BinningState::carve(bin_ptr, num_rendered, &b, sort_bytes, C);

    ENVGS_CUDA_TRY(cudaMemsetAsync(im.ranges, 0, (size_t)gx * gy * sizeof(uint2), stream));
    if (num_rendered > 0) {
        {
            ENVGS_TIMED("duplicate", stream);
            duplicate_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, g, b.keys_unsorted, b.vals_unsorted, gx, (uint32_t)num_rendered);
            ENVGS_LAUNCH_CHECK("duplicate", debug, stream);
            if (sync_free) {
                pad_keys_kernel<<<(num_rendered + 255) / 256, 256, 0, stream>>>((uint32_t)num_rendered, count_ptr, b.keys_unsorted, b.vals_unsorted,
                                                                                a.instance_count_out);
                ENVGS_LAUNCH_CHECK("pad_keys", debug, stream);
            }
        }
        int bit = (int)higher_msb((uint32_t)(gx * gy));
        {
            ENVGS_TIMED("sort", stream);
            ENVGS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(b.sort_temp, b.sort_temp_bytes, b.keys_unsorted, b.keys,
                                                           b.vals_unsorted, b.point_list, num_rendered, 0, 32 + bit, stream));
            count_launch(8);
        }
        {
            ENVGS_TIMED("tile_ranges", stream);
            tile_ranges_kernel<<<(num_rendered + 255) / 256, 256, 0, stream>>>(num_rendered, b.keys, im.ranges, (uint32_t)(gx * gy));
            ENVGS_LAUNCH_CHECK("tile_ranges", debug, stream);
        }
    }

    const float* feat = a.colors_precomp ? a.colors_precomp : g.rgb;
    if (num_rendered > 0) {
        ENVGS_TIMED("gather_sorted", stream);
        gather_sorted_kernel<C><<<(num_rendered + 255) / 256, 256, 0, stream>>>(num_rendered, count_ptr, b.point_list, g.rec, g.pixbox, feat, b);
        ENVGS_LAUNCH_CHECK("gather_sorted", debug, stream);
    }
    {
        constexpr size_t smem = 2 * sizeof(TileStage<C>) + 2 * sizeof(uint64_t);
        static bool attr_set = false;   // > 48 KB of dynamic shared memory needs the opt-in once per kernel instance
        if (!attr_set) {
            ENVGS_CUDA_TRY(cudaFuncSetAttribute(render_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        ENVGS_TIMED("render_fwd", stream);
        render_fwd_kernel<C><<<gx * gy, 256, smem, stream>>>(im.ranges, b, W, H, gx, a.background, im.final_T, im.n_contrib,
                                                             o.out_color, o.out_others, o.out_weight);
        ENVGS_LAUNCH_CHECK("render_fwd", debug, stream);
    }
    return num_rendered;
}

int raster_forward_dispatch(const envgs_raster_args& a, const envgs_raster_out& o, envgs_alloc_fn ga, envgs_alloc_fn ba,
                            envgs_alloc_fn ia, void* user, cudaStream_t stream) {
    switch (a.C) {
        case 3: return raster_forward_impl<3>(a, o, ga, ba, ia, user, stream);
        case 5: return raster_forward_impl<5>(a, o, ga, ba, ia, user, stream);
        case 7: return raster_forward_impl<7>(a, o, ga, ba, ia, user, stream);
        default: set_error("unsupported channel count C=%d (3, 5 or 7)", a.C); return ENVGS_ERR_INVALID;
    }
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ viewmatrix,
                                    uint8_t* __restrict__ present) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float3 p_view = xform_point_4x3(ld3(means3D, idx), viewmatrix);
    present[idx] = p_view.z > 0.2f;
}

int mark_visible_impl(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t stream) {
    if (P <= 0) return ENVGS_OK;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, viewmatrix, present);
    ENVGS_LAUNCH_CHECK("mark_visible", false, stream);
    return ENVGS_OK;
}

}  // namespace envgs
