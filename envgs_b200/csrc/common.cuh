// common.cuh — shared device helpers for the sm_100a EnvGS hot path (rasterizer + tracer).
// Constants restate RAST/cuda_rasterizer/auxiliary.h:21-60 and TRACE/optix_tracer/{auxiliary.h:31-54,config.h:16-44}.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/envgs_b200.h"

namespace envgs {

// ---- compile-time constants carried over from the reference ----
constexpr int TILE_X = 16, TILE_Y = 16, TILE_PIX = TILE_X * TILE_Y;
constexpr float NEAR_N = 0.2f, FAR_N = 100.0f;
constexpr float FILTER_SIZE = 0.707106f;      // sqrt(2)/2
constexpr float FILTER_INV_SQUARE = 2.0f;
constexpr float ALPHA_MIN = 1.0f / 255.0f, ALPHA_MAX = 0.99f, T_STOP = 0.0001f;
constexpr int DEPTH_OFFSET = 0, ALPHA_OFFSET = 1, NORMAL_OFFSET = 2, MIDDEPTH_OFFSET = 5, DISTORTION_OFFSET = 6;
// tracer
constexpr int TR_CHUNK = 16, TR_MAX_INTERSECTION = 256, TR_MID_CHANNELS = 16, TR_MAX_TRACE_DEPTH = 1;
constexpr float TR_START_OFFSET = 0.06f, TR_STEP_EPSILON = 1e-5f, TR_DEPTH_INF = 1e16f;
constexpr int MID_RAYO = 0, MID_RAYD = 3, MID_DPT = 6, MID_ACC = 7, MID_NORM = 8, MID_AUX = 11, MID_RGB = 13;

__device__ __constant__ const float SH_C0 = 0.28209479177387814f;
__device__ __constant__ const float SH_C1 = 0.4886025119029199f;
__device__ __constant__ const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                                -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                                0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                                -0.5900435899266435f};

// ---- error plumbing ----
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Optional per-kernel timing with CUDA events on the launching stream (envgs_profile_enable); a no-op otherwise.
struct KernelTimer {
    KernelTimer(const char* name, cudaStream_t stream);
    ~KernelTimer();
    int slot;
    cudaStream_t stream;
};
#define ENVGS_TIMED(name, stream) ::envgs::KernelTimer _envgs_kt_##__LINE__(name, stream)

#define ENVGS_CUDA_TRY(expr)                                                                      \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::envgs::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return ENVGS_ERR_CUDA;                                                                \
        }                                                                                         \
    } while (0)

#define ENVGS_LAUNCH_CHECK(name, debug, stream)                                                   \
    do {                                                                                          \
        ::envgs::count_launch();                                                                  \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e == cudaSuccess && (debug)) _e = cudaStreamSynchronize(stream);                     \
        if (_e != cudaSuccess) {                                                                  \
            ::envgs::set_error("kernel %s failed: %s", name, cudaGetErrorString(_e));             \
            return ENVGS_ERR_CUDA;                                                                \
        }                                                                                         \
    } while (0)

// ---- tiny vector algebra (float3/float2) ----
__host__ __device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__host__ __device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
__host__ __device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(s * a.x, s * a.y, s * a.z); }
__host__ __device__ __forceinline__ float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
__host__ __device__ __forceinline__ float3& operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
__host__ __device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ float sum3(float3 a) { return a.x + a.y + a.z; }
__host__ __device__ __forceinline__ float3 cross(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__host__ __device__ __forceinline__ float length3(float3 a) { return sqrtf(dot(a, a)); }
__host__ __device__ __forceinline__ float3 normalize3(float3 a) { return a / sqrtf(dot(a, a)); }

__device__ __forceinline__ float3 ld3(const float* p, int i) { return f3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

// column-major 4x4 (the reference passes W2C^T row-major == W2C column-major; auxiliary.h:79-118)
__device__ __forceinline__ float3 xform_point_4x3(float3 p, const float* m) {
    return f3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
              m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
__device__ __forceinline__ float3 xform_vec_4x3(float3 p, const float* m) {
    return f3(m[0] * p.x + m[4] * p.y + m[8] * p.z, m[1] * p.x + m[5] * p.y + m[9] * p.z,
              m[2] * p.x + m[6] * p.y + m[10] * p.z);
}
__device__ __forceinline__ float3 xform_vec_4x3_T(float3 p, const float* m) {
    return f3(m[0] * p.x + m[1] * p.y + m[2] * p.z, m[4] * p.x + m[5] * p.y + m[6] * p.z,
              m[8] * p.x + m[9] * p.y + m[10] * p.z);
}

// d normalize(v) / dv applied to dv  (auxiliary.h:128-138)
__device__ __forceinline__ float3 dnormvdv(float3 v, float3 dv) {
    float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    float3 r;
    r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return r;
}

// Rotation from an (unnormalised) wxyz quaternion, R = [c0 c1 c2] columns (auxiliary.h:213-235).
struct Rot3 { float3 c0, c1, c2; };
__device__ __forceinline__ Rot3 quat_to_rot(float4 q) {  // q = (w,x,y,z) stored in (.x,.y,.z,.w)
    float s = rsqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    float w = q.x * s, x = q.y * s, y = q.z * s, z = q.w * s;
    Rot3 R;
    R.c0 = f3(1.f - 2.f * (y * y + z * z), 2.f * (x * y + w * z), 2.f * (x * z - w * y));
    R.c1 = f3(2.f * (x * y - w * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + w * x));
    R.c2 = f3(2.f * (x * z + w * y), 2.f * (y * z - w * x), 1.f - 2.f * (x * x + y * y));
    return R;
}
// vjp of quat_to_rot wrt the normalised quaternion components, returned as-is (no normalisation
// back-prop), exactly like the reference (auxiliary.h:238-282).  g.cK = dL/d(column K of R).
__device__ __forceinline__ float4 quat_to_rot_vjp(float4 q, const Rot3& g) {
    float s = rsqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    float w = q.x * s, x = q.y * s, y = q.z * s, z = q.w * s;
    // v_R[c][r] in the reference == g.c<c>.<r>
    float4 v;
    v.x = 2.f * (x * (g.c1.z - g.c2.y) + y * (g.c2.x - g.c0.z) + z * (g.c0.y - g.c1.x));
    v.y = 2.f * (-2.f * x * (g.c1.y + g.c2.z) + y * (g.c0.y + g.c1.x) + z * (g.c0.z + g.c2.x) + w * (g.c1.z - g.c2.y));
    v.z = 2.f * (x * (g.c0.y + g.c1.x) - 2.f * y * (g.c0.x + g.c2.z) + z * (g.c1.z + g.c2.y) + w * (g.c2.x - g.c0.z));
    v.w = 2.f * (x * (g.c0.z + g.c2.x) + y * (g.c1.z + g.c2.y) - 2.f * z * (g.c0.x + g.c1.y) + w * (g.c0.y - g.c1.x));
    return v;
}

// ---- spherical harmonics (forward.cu:20-71 / TRACE forward.cu:29-72) ----
// basis[k], k < (deg+1)^2, for a UNIT direction; colour = sum_k basis[k]*sh[k] + 0.5, clamped at 0.
__device__ __forceinline__ void sh_basis(int deg, float3 d, float* b) {
    b[0] = SH_C0;
    if (deg > 0) {
        float x = d.x, y = d.y, z = d.z;
        b[1] = -SH_C1 * y; b[2] = SH_C1 * z; b[3] = -SH_C1 * x;
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = SH_C2[0] * xy; b[5] = SH_C2[1] * yz; b[6] = SH_C2[2] * (2.0f * zz - xx - yy);
            b[7] = SH_C2[3] * xz; b[8] = SH_C2[4] * (xx - yy);
            if (deg > 2) {
                b[9] = SH_C3[0] * y * (3.0f * xx - yy);
                b[10] = SH_C3[1] * xy * z;
                b[11] = SH_C3[2] * y * (4.0f * zz - xx - yy);
                b[12] = SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                b[13] = SH_C3[4] * x * (4.0f * zz - xx - yy);
                b[14] = SH_C3[5] * z * (xx - yy);
                b[15] = SH_C3[6] * x * (xx - 3.0f * yy);
            }
        }
    }
}

// d colour / d dir (per channel vector form): returns dRGB/dx, dRGB/dy, dRGB/dz given sh coefficients
// (backward.cu:58-122).  `sh` points at this surfel's (M,3) coefficients.
__device__ __forceinline__ void sh_dir_grad(int deg, float3 d, const float* sh, float3& gx, float3& gy, float3& gz) {
    gx = gy = gz = f3(0, 0, 0);
    if (deg < 1) return;
    float x = d.x, y = d.y, z = d.z;
#define SHV(k) f3(sh[3 * (k)], sh[3 * (k) + 1], sh[3 * (k) + 2])
    gx = -SH_C1 * SHV(3);
    gy = -SH_C1 * SHV(1);
    gz = SH_C1 * SHV(2);
    if (deg > 1) {
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        gx += SH_C2[0] * y * SHV(4) + SH_C2[2] * 2.f * -x * SHV(6) + SH_C2[3] * z * SHV(7) + SH_C2[4] * 2.f * x * SHV(8);
        gy += SH_C2[0] * x * SHV(4) + SH_C2[1] * z * SHV(5) + SH_C2[2] * 2.f * -y * SHV(6) + SH_C2[4] * 2.f * -y * SHV(8);
        gz += SH_C2[1] * y * SHV(5) + SH_C2[2] * 2.f * 2.f * z * SHV(6) + SH_C2[3] * x * SHV(7);
        if (deg > 2) {
            gx += SH_C3[0] * SHV(9) * 3.f * 2.f * xy + SH_C3[1] * SHV(10) * yz + SH_C3[2] * SHV(11) * -2.f * xy +
                  SH_C3[3] * SHV(12) * -3.f * 2.f * xz + SH_C3[4] * SHV(13) * (-3.f * xx + 4.f * zz - yy) +
                  SH_C3[5] * SHV(14) * 2.f * xz + SH_C3[6] * SHV(15) * 3.f * (xx - yy);
            gy += SH_C3[0] * SHV(9) * 3.f * (xx - yy) + SH_C3[1] * SHV(10) * xz +
                  SH_C3[2] * SHV(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SHV(12) * -3.f * 2.f * yz +
                  SH_C3[4] * SHV(13) * -2.f * xy + SH_C3[5] * SHV(14) * -2.f * yz + SH_C3[6] * SHV(15) * -3.f * 2.f * xy;
            gz += SH_C3[1] * SHV(10) * xy + SH_C3[2] * SHV(11) * 4.f * 2.f * yz +
                  SH_C3[3] * SHV(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SHV(13) * 4.f * 2.f * xz +
                  SH_C3[5] * SHV(14) * (xx - yy);
        }
    }
#undef SHV
}

// Per-(pixel, surfel) / per-hit division and exponential of the render and shading kernels.
// DEFAULT: the reference's own IEEE forms (`a / b`, `expf`; it is built without -use_fast_math).  Together with the
// reference-ordered per-surfel preprocess (glm_order.cuh, compute_aabb) this makes the rasterizer's forward outputs agree
// with the reference build to 2e-7 and its gradients to 4e-7 relative L2 at 1 M surfels / 1080p
// (profiles/round2_parity_experiments.md), because every alpha >= 1/255, T < 1e-4, T > 0.5 and rho3d <= rho2d decision is
// then taken on identical bits.  Forward and backward kernels use the same helpers, so they always agree with each other.
// -DENVGS_FAST_MATH selects the 1-instruction MUFU forms (rcp.approx / ex2.approx, <= 2 ulp): render kernels 14-17 % faster,
// gradients still within 8e-5 relative L2, but ~3 of 2 M pixels flip an alpha >= 1/255 decision (colour error up to 1e-3).
// px_rcp_v: reciprocal for VALUE-only uses (distortion depth mapping, the backward's T / (1 - alpha), 1 / p.z in the chain rule):
// no alpha / T / branch decision reads these, so the 1-ulp MUFU form only moves gradients and the distortion map by ~1e-7 relative.
// Every quantity a decision depends on (s = p.xy / p.z -> rho -> alpha) goes through px_div / px_exp below.
__device__ __forceinline__ float px_rcp_v(float x) { return __fdividef(1.0f, x); }
#ifdef ENVGS_FAST_MATH
__device__ __forceinline__ float px_rcp(float x) { return __fdividef(1.0f, x); }
__device__ __forceinline__ float px_exp(float x) { return __expf(x); }
__device__ __forceinline__ float px_div(float a, float b, float rb) { return a * rb; }   // rb = px_rcp(b)
#else
__device__ __forceinline__ float px_rcp(float x) { return 1.0f / x; }
__device__ __forceinline__ float px_exp(float x) { return expf(x); }
__device__ __forceinline__ float px_div(float a, float b, float) { return a / b; }       // the reference's own a / b
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Warp "split-reduce": every lane holds N (<= 32) partial values v[0..N-1]; afterwards each of the N components' warp-wide
// sum sits in exactly one lane (warp_split_component<N>(lane) tells which, -1 = none).  Recursive halving over the lane
// bits 16, 8, 4, 2, 1: at each stage the lower half of the lanes keeps the first ceil(n/2) values and the upper half the
// rest, so a stage costs ceil(n/2) shuffles and the whole reduction N/2 + N/4 + ... (21 shuffles for N = 20) instead of
// the 5*N of N separate butterfly sums or the 31 of a power-of-two transpose.  The reduction step of render_bwd.
template <int N, int H>
struct WarpSplitReduce {
    static __device__ __forceinline__ float run(float (&v)[32], int lane) {
        constexpr int NL = (N + 1) / 2;
        const bool up = (lane & H) != 0;
#pragma unroll
        for (int i = 0; i < NL; i++) {
            const float hi = (NL + i < N) ? v[NL + i] : 0.0f;
            const float send = up ? v[i] : hi;
            const float keep = up ? hi : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, H);
        }
        return WarpSplitReduce<NL, H / 2>::run(v, lane);
    }
};
template <int N>
struct WarpSplitReduce<N, 0> {
    static __device__ __forceinline__ float run(float (&v)[32], int) { return v[0]; }
};
template <int N>
__device__ __forceinline__ float warp_split_reduce(float (&v)[32], int lane) {
    static_assert(N >= 1 && N <= 32, "at most one component per lane");
    return WarpSplitReduce<N, 16>::run(v, lane);
}
// which component's sum lane `lane` holds after warp_split_reduce<N> (-1: none)
template <int N>
__device__ __forceinline__ int warp_split_component(int lane) {
    // n: values every lane of the current group carries (the upper half is zero-padded to ceil(n/2)); real: how many of them
    // are components rather than padding
    int lo = 0, n = N, real = N;
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
        const int nl = (n + 1) / 2;
        if (lane & h) { lo += nl; real -= nl; } else { real = real < nl ? real : nl; }
        n = nl;
    }
    return real >= 1 ? lo : -1;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace envgs
