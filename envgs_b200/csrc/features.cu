// features.cu — fused "SH -> view-dependent RGB, clamp, append specular/roughness" for the base surfels, with backward.
//
// SURVEY.md §8(f) rank 1 (glue on the caller's side of the rasterizer): EnvGS evaluates the base colours in PyTorch
// (EVC/utils/gaussian2d_utils.py:1071-1084: eval_sh + clamp_min(+0.5) + cat[specular, roughness]) because the C = 5
// rasterizer only takes precomputed colours.  In eager PyTorch that is ~60 element-wise launches forward and, in backward,
// sixteen select_backward + add_ passes over the (P,16,3) coefficient tensor (~9 GB of traffic at 1 M surfels).
// Here: one pass each way, 192 B of coefficients read once per surfel.
#include "common.cuh"

namespace envgs {

__device__ __forceinline__ float3 sh_basis_grad_k(int k, float3 d) {
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

// Both kernels: one surfel per thread, but the (P, M, 3) coefficient rows of a warp's 32 consecutive surfels are ONE
// contiguous 32*M*12-byte block, so the warp moves it with coalesced 16-byte accesses through a padded shared-memory tile
// (row stride M*3+1 words: conflict-free per-thread row walks).  A thread-per-row walk straight to global memory touched 32
// different 128-byte lines per instruction and ran at 1.2 TB/s.
constexpr int SHF_THREADS = 128;
constexpr int SHF_MAXROW = 48;            // M * 3 floats per surfel (M <= 16)

__device__ __forceinline__ void shf_tile_load(const float* __restrict__ shs, int P, int row_floats, int warp_first, int lane, float* tile) {
    const size_t base = (size_t)warp_first * row_floats;
    const int n_rows = min(32, P - warp_first);
    const int total = n_rows * row_floats;
    if ((row_floats & 3) == 0 && (reinterpret_cast<uintptr_t>(shs) & 15) == 0) {
        const float4* src = reinterpret_cast<const float4*>(shs + base);
        for (int q = lane; q < total / 4; q += 32) {
            const float4 v = __ldg(src + q);
            const int f = 4 * q, r = f / row_floats, c = f - r * row_floats;   // a float4 never straddles rows (row_floats % 4 == 0)
            float* t = tile + r * (row_floats + 1) + c;
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
        }
    } else {
        for (int f = lane; f < total; f += 32) { const int r = f / row_floats, c = f - r * row_floats; tile[r * (row_floats + 1) + c] = shs[base + f]; }
    }
}
__device__ __forceinline__ void shf_tile_store(float* __restrict__ dst, int P, int row_floats, int warp_first, int lane, const float* tile) {
    const size_t base = (size_t)warp_first * row_floats;
    const int n_rows = min(32, P - warp_first);
    const int total = n_rows * row_floats;
    if ((row_floats & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        float4* out = reinterpret_cast<float4*>(dst + base);
        for (int q = lane; q < total / 4; q += 32) {
            const int f = 4 * q, r = f / row_floats, c = f - r * row_floats;
            const float* t = tile + r * (row_floats + 1) + c;
            out[q] = make_float4(t[0], t[1], t[2], t[3]);
        }
    } else {
        for (int f = lane; f < total; f += 32) { const int r = f / row_floats, c = f - r * row_floats; dst[base + f] = tile[r * (row_floats + 1) + c]; }
    }
}

// out (P, 3 + E): rgb | extra (E columns copied from `extra`, e.g. specular, roughness)
__global__ void __launch_bounds__(SHF_THREADS) sh_features_fwd_kernel(int P, int D, int M, int E, const float* __restrict__ means,
                                                                      const float* __restrict__ shs, const float* __restrict__ extra,
                                                                      const float* __restrict__ campos, float* __restrict__ out) {
    __shared__ float s_tile[(SHF_THREADS / 32) * 32 * (SHF_MAXROW + 1)];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warp_first = blockIdx.x * SHF_THREADS + warp * 32;
    if (warp_first >= P) return;
    const int row = M * 3;
    float* tile = s_tile + warp * 32 * (SHF_MAXROW + 1);
    shf_tile_load(shs, P, row, warp_first, lane, tile);
    __syncwarp();
    const int i = warp_first + lane;
    if (i >= P) return;
    float3 dir = ld3(means, i) - f3(campos[0], campos[1], campos[2]);
    dir = dir / fmaxf(length3(dir), 1e-12f);   // F.normalize's eps
    float b[16];
    sh_basis(D, dir, b);
    const int nb = (D + 1) * (D + 1);
    const float* sh = tile + lane * (row + 1);
    float3 res = f3(0, 0, 0);
    for (int k = 0; k < nb; k++) res += b[k] * f3(sh[3 * k], sh[3 * k + 1], sh[3 * k + 2]);
    float* o = out + (size_t)i * (3 + E);
    o[0] = fmaxf(res.x + 0.5f, 0.0f); o[1] = fmaxf(res.y + 0.5f, 0.0f); o[2] = fmaxf(res.z + 0.5f, 0.0f);
    for (int e = 0; e < E; e++) o[3 + e] = extra[(size_t)i * E + e];
}

__global__ void __launch_bounds__(SHF_THREADS) sh_features_bwd_kernel(int P, int D, int M, int E, const float* __restrict__ means,
                                                                      const float* __restrict__ shs, const float* __restrict__ campos,
                                                                      const float* __restrict__ dL_dout, float* __restrict__ dL_dmeans,
                                                                      float* __restrict__ dL_dshs, float* __restrict__ dL_dextra,
                                                                      float* __restrict__ dL_dfactor) {
    __shared__ float s_tile[(SHF_THREADS / 32) * 32 * (SHF_MAXROW + 1)];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warp_first = blockIdx.x * SHF_THREADS + warp * 32;
    if (warp_first >= P) return;
    const int row = M * 3;
    float* tile = s_tile + warp * 32 * (SHF_MAXROW + 1);
    shf_tile_load(shs, P, row, warp_first, lane, tile);
    __syncwarp();
    const int i = warp_first + lane;
    if (i < P) {
        const float3 dir_orig = ld3(means, i) - f3(campos[0], campos[1], campos[2]);
        const float len = fmaxf(length3(dir_orig), 1e-12f);
        const float3 dir = dir_orig / len;
        float b[16];
        sh_basis(D, dir, b);
        const int nb = (D + 1) * (D + 1);
        float* sh = tile + lane * (row + 1);     // coefficients in, their gradients out (same slot, this thread only)
        float3 res = f3(0, 0, 0);
        for (int k = 0; k < nb; k++) res += b[k] * f3(sh[3 * k], sh[3 * k + 1], sh[3 * k + 2]);
        const float* g = dL_dout + (size_t)i * (3 + E);
        // clamp_min(x + 0.5, 0): gradient passes where the pre-clamp value is >= 0
        const float3 dRGB = f3(res.x + 0.5f >= 0.0f ? g[0] : 0.0f, res.y + 0.5f >= 0.0f ? g[1] : 0.0f, res.z + 0.5f >= 0.0f ? g[2] : 0.0f);
        float3 dL_ddir = f3(0, 0, 0);
        for (int k = 0; k < M; k++) {
            const float bk = k < nb ? b[k] : 0.0f;
            if (k < nb) {
                const float u = sh[3 * k] * dRGB.x + sh[3 * k + 1] * dRGB.y + sh[3 * k + 2] * dRGB.z;
                dL_ddir += u * sh_basis_grad_k(k, dir);
            }
            sh[3 * k] = bk * dRGB.x; sh[3 * k + 1] = bk * dRGB.y; sh[3 * k + 2] = bk * dRGB.z;
        }
        if (dL_dfactor) { dL_dfactor[3 * (size_t)i] = dRGB.x; dL_dfactor[3 * (size_t)i + 1] = dRGB.y; dL_dfactor[3 * (size_t)i + 2] = dRGB.z; }
        const float3 dm = dnormvdv(dir_orig, dL_ddir);
        dL_dmeans[3 * (size_t)i] = dm.x; dL_dmeans[3 * (size_t)i + 1] = dm.y; dL_dmeans[3 * (size_t)i + 2] = dm.z;
        for (int e = 0; e < E; e++) dL_dextra[(size_t)i * E + e] = g[3 + e];
    }
    __syncwarp();
    if (dL_dshs) shf_tile_store(dL_dshs, P, row, warp_first, lane, tile);
}

// ------------------------------------------------------------------------------------------------------
// Factored exchange of the SH gradient (view-parallel training)
// ------------------------------------------------------------------------------------------------------
// One view's gradient of a surfel's (M,3) SH coefficients is RANK ONE: basis(dir_i) (M values, a function of the surfel's
// position and that view's camera centre) times the clamp-masked colour gradient (3 values).  Summing it over N ranks with an
// all-reduce moves 4*M*3 = 192 B per surfel both ways; all-GATHERING the 12-byte factors and rebuilding the sum locally moves
// 12*N bytes one way — 4x less at N = 8 for the 80 % of the base set's gradient bytes that are SH.  This kernel is the
// receiving end fused with the optimizer: per surfel, sum_r basis(normalize(mean_i - campos_r)) (x) factor_r,i in fixed rank
// order (every rank computes bit-identical sums, so the replicas stay in sync), then the Adam update of the 3M coefficients
// (same arithmetic as adam_step_kernel).  Coefficients and moments move through padded shared-memory tiles like above.
constexpr int SHA_THREADS = 128;
constexpr int SHA_STRIDE = SHF_MAXROW + 4;   // 52 floats: a lane's 16-byte stores of its row land in distinct bank groups
__device__ __forceinline__ void adam_one(float& p, float& m, float& v, float g, float lr_over_bc1, float inv_sqrt_bc2, float b1, float b2, float eps) {
    m = b1 * m + (1.0f - b1) * g;
    v = b2 * v + (1.0f - b2) * g * g;
    p -= lr_over_bc1 * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
}
// Layout: a warp owns 32 consecutive surfels.  Lane i rebuilds surfel i's 3M gradient values in registers and parks them in a
// shared tile; then the warp streams the surfels' contiguous 32 * 3M floats of coefficients and both moments straight from / to
// global memory with 16-byte accesses (no staging of the 3 x 6 KB themselves: 6.6 KB of shared memory per warp instead of 18.8,
// and every lane has 3 independent 16-byte loads in flight per iteration).
__global__ void __launch_bounds__(SHA_THREADS) sh_adam_factored_kernel(int P, int D, int M, int N, const float* __restrict__ means,
                                                                        const float* __restrict__ campos_all, const float* __restrict__ factors_all,
                                                                        float* __restrict__ shs, float* __restrict__ exp_avg,
                                                                        float* __restrict__ exp_avg_sq, float lr_over_bc1, float inv_sqrt_bc2,
                                                                        float b1, float b2, float eps, float grad_scale) {
    __shared__ __align__(16) float s_g[(SHA_THREADS / 32) * 32 * SHA_STRIDE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warp_first = blockIdx.x * SHA_THREADS + warp * 32;
    if (warp_first >= P) return;
    const int row = M * 3;
    float* tg = s_g + warp * 32 * SHA_STRIDE;
    const int i = warp_first + lane;
    if (i < P) {
        const float3 mean = ld3(means, i);
        const int nb = (D + 1) * (D + 1);
        float g[SHF_MAXROW];
#pragma unroll
        for (int q = 0; q < SHF_MAXROW; q++) g[q] = 0.0f;
        for (int r = 0; r < N; r++) {
            const float3 dir_orig = mean - f3(campos_all[3 * r], campos_all[3 * r + 1], campos_all[3 * r + 2]);
            const float3 dir = dir_orig / fmaxf(length3(dir_orig), 1e-12f);
            float b[16];
            sh_basis(D, dir, b);
            const float* f = factors_all + ((size_t)r * P + i) * 3;
            const float fx = f[0], fy = f[1], fz = f[2];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const float bk = k < nb ? b[k] : 0.0f;
                g[3 * k] += bk * fx; g[3 * k + 1] += bk * fy; g[3 * k + 2] += bk * fz;
            }
        }
        float4* dst = reinterpret_cast<float4*>(tg + lane * SHA_STRIDE);
#pragma unroll
        for (int q = 0; q < SHF_MAXROW / 4; q++)
            dst[q] = make_float4(g[4 * q] * grad_scale, g[4 * q + 1] * grad_scale, g[4 * q + 2] * grad_scale, g[4 * q + 3] * grad_scale);
    }
    __syncwarp();
    const int n_valid = min(32, P - warp_first);
    const size_t base = (size_t)warp_first * row;
    if ((row & 3) == 0) {
        const int row4 = row >> 2, total4 = n_valid * row4;
        float4* p4 = reinterpret_cast<float4*>(shs + base);
        float4* m4 = reinterpret_cast<float4*>(exp_avg + base);
        float4* v4 = reinterpret_cast<float4*>(exp_avg_sq + base);
#pragma unroll 2
        for (int e = lane; e < total4; e += 32) {
            const int sidx = e / row4, q4 = e - sidx * row4;
            float4 pp = p4[e], mm = m4[e], vv = v4[e];
            const float4 gg = *reinterpret_cast<const float4*>(tg + sidx * SHA_STRIDE + 4 * q4);
            adam_one(pp.x, mm.x, vv.x, gg.x, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
            adam_one(pp.y, mm.y, vv.y, gg.y, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
            adam_one(pp.z, mm.z, vv.z, gg.z, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
            adam_one(pp.w, mm.w, vv.w, gg.w, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
            p4[e] = pp; m4[e] = mm; v4[e] = vv;
        }
    } else {
        const int total = n_valid * row;
        for (int e = lane; e < total; e += 32) {
            const int sidx = e / row, q = e - sidx * row;
            float pp = shs[base + e], mm = exp_avg[base + e], vv = exp_avg_sq[base + e];
            adam_one(pp, mm, vv, tg[sidx * SHA_STRIDE + q], lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
            shs[base + e] = pp; exp_avg[base + e] = mm; exp_avg_sq[base + e] = vv;
        }
    }
}

int sh_features_forward_impl(int P, int D, int M, int E, const float* means, const float* shs, const float* extra, const float* campos,
                             float* out, cudaStream_t stream) {
    if (P <= 0) return ENVGS_OK;
    ENVGS_TIMED("sh_features_fwd", stream);
    sh_features_fwd_kernel<<<(P + SHF_THREADS - 1) / SHF_THREADS, SHF_THREADS, 0, stream>>>(P, D, M, E, means, shs, extra, campos, out);
    ENVGS_LAUNCH_CHECK("sh_features_fwd", false, stream);
    return ENVGS_OK;
}
int sh_features_backward_impl(int P, int D, int M, int E, const float* means, const float* shs, const float* campos, const float* dL_dout,
                              float* dL_dmeans, float* dL_dshs, float* dL_dextra, float* dL_dfactor, cudaStream_t stream) {
    if (P <= 0) return ENVGS_OK;
    ENVGS_TIMED("sh_features_bwd", stream);
    sh_features_bwd_kernel<<<(P + SHF_THREADS - 1) / SHF_THREADS, SHF_THREADS, 0, stream>>>(P, D, M, E, means, shs, campos, dL_dout, dL_dmeans, dL_dshs,
                                                                                            dL_dextra, dL_dfactor);
    ENVGS_LAUNCH_CHECK("sh_features_bwd", false, stream);
    return ENVGS_OK;
}

int sh_adam_factored_impl(int P, int D, int M, int N, const float* means, const float* campos_all, const float* factors_all, float* shs,
                          float* exp_avg, float* exp_avg_sq, int step, float lr, float b1, float b2, float eps, float grad_scale,
                          cudaStream_t stream) {
    if (P <= 0) return ENVGS_OK;
    ENVGS_TIMED("sh_adam_factored", stream);
    const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
    sh_adam_factored_kernel<<<(P + SHA_THREADS - 1) / SHA_THREADS, SHA_THREADS, 0, stream>>>(P, D, M, N, means, campos_all, factors_all, shs, exp_avg,
                                                                                             exp_avg_sq, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)),
                                                                                             b1, b2, eps, grad_scale);
    ENVGS_LAUNCH_CHECK("sh_adam_factored", false, stream);
    return ENVGS_OK;
}

}  // namespace envgs
