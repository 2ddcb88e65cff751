// glue.cu — fused per-pixel glue between the two operators ("callers either side of the path").
//
// In the reference these are ~150 eager PyTorch launches per view on 25 MB maps: allmap split, depth = D/alpha with
// nan_to_num, normal view->world (gaussian2d_utils.py:1117-1131), dpt2norm (:1158-1206), camera rays
// (ray_utils.py:24-78), reflection (envgs_sampler.py:420-431), blend (:474) and the image / normal-consistency loss.
// Here: two forward and two backward kernels, each one pass over the frame.
//
//   reflect_fwd / reflect_bwd          allmap -> reflected rays (ref_o, ref_d)             [before the tracer]
//   blend_loss_fwd / blend_loss_bwd    rgb = (1-s) base + s env;  loss = mean|rgb - target| + lambda * mean(1 - <n, n_surf>)
#include "common.cuh"

namespace envgs {

struct GlueCam {
    const float* K;       // (3,3) row-major intrinsics used for the camera rays (pixel centres at +0.5)
    const float* R;       // (3,3) row-major world->camera rotation
    const float* campos;  // (3,)
    float tanfovx, tanfovy;
    int H, W;
};

__device__ __forceinline__ float3 rot_T(const float* R, float3 v) {   // R^T v
    return f3(R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z, R[2] * v.x + R[5] * v.y + R[8] * v.z);
}
__device__ __forceinline__ float3 rot(const float* R, float3 v) {     // R v
    return f3(R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z);
}
// camera ray through the centre of pixel (row i, col j): direction with camera-space z = 1 (ray_utils.py, z_depth=True)
__device__ __forceinline__ float3 camera_ray_dir(const GlueCam& c, int i, int j) {
    const float fx = c.K[0], fy = c.K[4], cx = c.K[2], cy = c.K[5];
    return rot_T(c.R, f3((j + 0.5f - cx) / fx, (i + 0.5f - cy) / fy, 1.0f));
}
// ray used by dpt2norm: integer pixel coordinates, focal from the field of view, principal point W/2,H/2
__device__ __forceinline__ float3 backproject_dir(const GlueCam& c, int i, int j) {
    const float fx = c.W / (2.0f * c.tanfovx), fy = c.H / (2.0f * c.tanfovy);
    return rot_T(c.R, f3((j - c.W * 0.5f) / fx, (i - c.H * 0.5f) / fy, 1.0f));
}
__device__ __forceinline__ float safe_depth(float D, float a) {   // nan_to_num(D / a, 0, 0, 0)
    const float d = D / a;
    return isfinite(d) ? d : 0.0f;
}
__device__ __forceinline__ float3 surf_point(const GlueCam& c, const float* __restrict__ allmap, size_t HW, int i, int j) {
    const size_t p = (size_t)i * c.W + j;
    const float d = safe_depth(allmap[p], allmap[HW + p]);
    return d * backproject_dir(c, i, j) + f3(c.campos[0], c.campos[1], c.campos[2]);
}
// pseudo surface normal of dpt2norm before the alpha factor; zero on the one-pixel border
__device__ __forceinline__ float3 surf_normal_raw(const GlueCam& c, const float* __restrict__ allmap, size_t HW, int i, int j,
                                                  float3& dx, float3& dy, float3& cr) {
    dx = dy = cr = f3(0, 0, 0);
    if (i < 1 || j < 1 || i > c.H - 2 || j > c.W - 2) return f3(0, 0, 0);
    dx = surf_point(c, allmap, HW, i + 1, j) - surf_point(c, allmap, HW, i - 1, j);
    dy = surf_point(c, allmap, HW, i, j + 1) - surf_point(c, allmap, HW, i, j - 1);
    cr = cross(dx, dy);
    return cr / fmaxf(length3(cr), 1e-12f);
}

// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reflect_fwd_kernel(GlueCam c, const float* __restrict__ allmap, float* __restrict__ ref_o,
                                                          float* __restrict__ ref_d) {
    const size_t HW = (size_t)c.H * c.W;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int i = (int)(p / c.W), j = (int)(p % c.W);
    const float depth = safe_depth(allmap[p], allmap[HW + p]);
    const float3 nw = rot_T(c.R, f3(allmap[2 * HW + p], allmap[3 * HW + p], allmap[4 * HW + p]));
    const float3 nn = nw / (length3(nw) + 1e-8f);
    const float3 d = camera_ray_dir(c, i, j);
    const float3 rd = d - 2.0f * dot(d, nn) * nn;
    const float3 ro = f3(c.campos[0], c.campos[1], c.campos[2]) + depth * d;
    ref_o[3 * p] = ro.x; ref_o[3 * p + 1] = ro.y; ref_o[3 * p + 2] = ro.z;
    ref_d[3 * p] = rd.x; ref_d[3 * p + 1] = rd.y; ref_d[3 * p + 2] = rd.z;
}

__global__ void __launch_bounds__(256) reflect_bwd_kernel(GlueCam c, const float* __restrict__ allmap, const float* __restrict__ g_ro,
                                                          const float* __restrict__ g_rd, float* __restrict__ dL_dallmap) {
    const size_t HW = (size_t)c.H * c.W;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int i = (int)(p / c.W), j = (int)(p % c.W);
    const float D = allmap[p], a = allmap[HW + p];
    const float3 nw = rot_T(c.R, f3(allmap[2 * HW + p], allmap[3 * HW + p], allmap[4 * HW + p]));
    const float r = length3(nw), s = 1.0f / (r + 1e-8f);
    const float3 nn = s * nw;
    const float3 d = camera_ray_dir(c, i, j);
    const float3 go = f3(g_ro[3 * p], g_ro[3 * p + 1], g_ro[3 * p + 2]), gd = f3(g_rd[3 * p], g_rd[3 * p + 1], g_rd[3 * p + 2]);
    const float g_depth = dot(d, go);
    const float3 g_nn = -2.0f * (dot(gd, nn) * d + dot(d, nn) * gd);
    float3 g_nw = s * g_nn;
    if (r > 0.0f) g_nw = g_nw - (dot(g_nn, nw) / (r * (r + 1e-8f) * (r + 1e-8f))) * nw;
    const float3 g_nv = rot(c.R, g_nw);
    const bool fin = isfinite(D / a);
    dL_dallmap[p] = fin ? g_depth / a : 0.0f;
    dL_dallmap[HW + p] = fin ? -g_depth * D / (a * a) : 0.0f;
    dL_dallmap[2 * HW + p] = g_nv.x; dL_dallmap[3 * HW + p] = g_nv.y; dL_dallmap[4 * HW + p] = g_nv.z;
    dL_dallmap[5 * HW + p] = 0.0f; dL_dallmap[6 * HW + p] = 0.0f;
}

// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) blend_loss_fwd_kernel(GlueCam c, const float* __restrict__ image /* (5,H,W) */,
                                                             const float* __restrict__ allmap, const float* __restrict__ env_rgb /* (H,W,3) or null */,
                                                             const float* __restrict__ target /* (H,W,3) */, float lambda_n,
                                                             float* __restrict__ rgb_out, float* __restrict__ loss /* [2]: l1 sum, ncons sum */) {
    const size_t HW = (size_t)c.H * c.W;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float l1 = 0.0f, nc = 0.0f;
    if (p < HW) {
        const int i = (int)(p / c.W), j = (int)(p % c.W);
        const float spec = env_rgb ? image[3 * HW + p] : 0.0f;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float base = image[ch * HW + p];
            const float v = env_rgb ? (1.0f - spec) * base + spec * env_rgb[3 * p + ch] : base;
            rgb_out[3 * p + ch] = v;
            l1 += fabsf(v - target[3 * p + ch]);
        }
        const float3 nw = rot_T(c.R, f3(allmap[2 * HW + p], allmap[3 * HW + p], allmap[4 * HW + p]));
        float3 dx, dy, cr;
        const float3 sn = allmap[HW + p] * surf_normal_raw(c, allmap, HW, i, j, dx, dy, cr);
        nc = 1.0f - dot(nw, sn);
    }
    l1 = warp_sum(l1); nc = warp_sum(nc);
    __shared__ float s_l1[8], s_nc[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_l1[warp] = l1; s_nc[warp] = nc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.0f, b = 0.0f;
        for (int w = 0; w < 8; w++) { a += s_l1[w]; b += s_nc[w]; }
        atomicAdd(&loss[0], a / (3.0f * (float)HW));
        atomicAdd(&loss[1], lambda_n * b / (float)HW);
    }
}

// a_q = dL/d dx_q and b_q = dL/d dy_q of the surface-normal term at pixel q (zero outside the interior)
__device__ __forceinline__ void surf_normal_adjoint(const GlueCam& c, const float* __restrict__ allmap, size_t HW, int i, int j, float k,
                                                    float3& a_q, float3& b_q) {
    a_q = b_q = f3(0, 0, 0);
    if (i < 1 || j < 1 || i > c.H - 2 || j > c.W - 2) return;
    const size_t q = (size_t)i * c.W + j;
    float3 dx, dy, cr;
    surf_normal_raw(c, allmap, HW, i, j, dx, dy, cr);
    const float3 nw = rot_T(c.R, f3(allmap[2 * HW + q], allmap[3 * HW + q], allmap[4 * HW + q]));
    const float3 G = (-k * allmap[HW + q]) * nw;           // dL/d(normalised cross), alpha detached
    const float len = length3(cr);
    float3 v;
    if (len > 1e-12f) v = dnormvdv(cr, G);
    else v = G / 1e-12f;
    a_q = cross(dy, v);
    b_q = cross(v, dx);
}

__global__ void __launch_bounds__(256) blend_loss_bwd_kernel(GlueCam c, const float* __restrict__ image, const float* __restrict__ allmap,
                                                             const float* __restrict__ env_rgb, const float* __restrict__ target,
                                                             const float* __restrict__ rgb, float lambda_n, const float* __restrict__ g_loss /* [1] */,
                                                             const float* __restrict__ g_rgb /* (H,W,3) or null */, float* __restrict__ dL_dimage,
                                                             float* __restrict__ dL_dallmap, float* __restrict__ dL_denv) {
    const size_t HW = (size_t)c.H * c.W;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int i = (int)(p / c.W), j = (int)(p % c.W);
    const float gl = g_loss[0];
    const float spec = env_rgb ? image[3 * HW + p] : 0.0f;
    float d_spec = 0.0f;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const float diff = rgb[3 * p + ch] - target[3 * p + ch];
        float drgb = gl * (diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f)) / (3.0f * (float)HW);
        if (g_rgb) drgb += g_rgb[3 * p + ch];
        if (env_rgb) {
            const float e = env_rgb[3 * p + ch], b = image[ch * HW + p];
            dL_denv[3 * p + ch] = spec * drgb;
            dL_dimage[ch * HW + p] = (1.0f - spec) * drgb;
            d_spec += drgb * (e - b);
        } else {
            dL_dimage[ch * HW + p] = drgb;
        }
    }
    dL_dimage[3 * HW + p] = d_spec;
    dL_dimage[4 * HW + p] = 0.0f;

    // normal-consistency term: lambda * mean(1 - <n_w, alpha * normalize(dx x dy)>)
    const float k = lambda_n * gl / (float)HW;
    float3 dx, dy, cr;
    const float alpha = allmap[HW + p];
    const float3 sn = alpha * surf_normal_raw(c, allmap, HW, i, j, dx, dy, cr);
    const float3 g_nv = rot(c.R, (-k) * sn);
    // depth gradient: this pixel's surface point enters dx of (i-1,j),(i+1,j) and dy of (i,j-1),(i,j+1)
    float3 a0, b0, a1, b1, tmp;
    surf_normal_adjoint(c, allmap, HW, i - 1, j, k, a0, tmp);
    surf_normal_adjoint(c, allmap, HW, i + 1, j, k, a1, tmp);
    surf_normal_adjoint(c, allmap, HW, i, j - 1, k, tmp, b0);
    surf_normal_adjoint(c, allmap, HW, i, j + 1, k, tmp, b1);
    const float3 g_pt = (a0 - a1) + (b0 - b1);
    const float g_depth = dot(g_pt, backproject_dir(c, i, j));
    const float D = allmap[p];
    const bool fin = isfinite(D / alpha);
    dL_dallmap[p] = fin ? g_depth / alpha : 0.0f;
    dL_dallmap[HW + p] = fin ? -g_depth * D / (alpha * alpha) : 0.0f;
    dL_dallmap[2 * HW + p] = g_nv.x; dL_dallmap[3 * HW + p] = g_nv.y; dL_dallmap[4 * HW + p] = g_nv.z;
    dL_dallmap[5 * HW + p] = 0.0f; dL_dallmap[6 * HW + p] = 0.0f;
}

// ------------------------------------------------------------------------------------------------------
static GlueCam make_cam(const envgs_glue_camera& cam) {
    GlueCam c;
    c.K = cam.K; c.R = cam.R; c.campos = cam.campos; c.tanfovx = cam.tanfovx; c.tanfovy = cam.tanfovy; c.H = cam.H; c.W = cam.W;
    return c;
}

int glue_reflect_forward(const envgs_glue_camera& cam, const float* allmap, float* ref_o, float* ref_d, cudaStream_t stream) {
    const size_t HW = (size_t)cam.H * cam.W;
    ENVGS_TIMED("glue_reflect_fwd", stream);
    reflect_fwd_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, stream>>>(make_cam(cam), allmap, ref_o, ref_d);
    ENVGS_LAUNCH_CHECK("glue_reflect_fwd", false, stream);
    return ENVGS_OK;
}
int glue_reflect_backward(const envgs_glue_camera& cam, const float* allmap, const float* g_ro, const float* g_rd, float* dL_dallmap,
                          cudaStream_t stream) {
    const size_t HW = (size_t)cam.H * cam.W;
    ENVGS_TIMED("glue_reflect_bwd", stream);
    reflect_bwd_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, stream>>>(make_cam(cam), allmap, g_ro, g_rd, dL_dallmap);
    ENVGS_LAUNCH_CHECK("glue_reflect_bwd", false, stream);
    return ENVGS_OK;
}
int glue_blend_loss_forward(const envgs_glue_camera& cam, const float* image, const float* allmap, const float* env_rgb, const float* target,
                            float lambda_n, float* rgb_out, float* loss2, cudaStream_t stream) {
    const size_t HW = (size_t)cam.H * cam.W;
    ENVGS_CUDA_TRY(cudaMemsetAsync(loss2, 0, 2 * sizeof(float), stream));
    ENVGS_TIMED("glue_blend_loss_fwd", stream);
    blend_loss_fwd_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, stream>>>(make_cam(cam), image, allmap, env_rgb, target, lambda_n, rgb_out, loss2);
    ENVGS_LAUNCH_CHECK("glue_blend_loss_fwd", false, stream);
    return ENVGS_OK;
}
int glue_blend_loss_backward(const envgs_glue_camera& cam, const float* image, const float* allmap, const float* env_rgb, const float* target,
                             const float* rgb, float lambda_n, const float* g_loss, const float* g_rgb, float* dL_dimage, float* dL_dallmap,
                             float* dL_denv, cudaStream_t stream) {
    const size_t HW = (size_t)cam.H * cam.W;
    ENVGS_TIMED("glue_blend_loss_bwd", stream);
    blend_loss_bwd_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, stream>>>(make_cam(cam), image, allmap, env_rgb, target, rgb, lambda_n, g_loss,
                                                                           g_rgb, dL_dimage, dL_dallmap, dL_denv);
    ENVGS_LAUNCH_CHECK("glue_blend_loss_bwd", false, stream);
    return ENVGS_OK;
}

// ------------------------------------------------------------------------------------------------------
// surfel -> quad corners at +-3 sigma, reference order [(-,+),(-,-),(+,+),(+,-)] (EVC/utils/optix_utils.py:39-69): the
// vertex buffer build_acceleration_structure takes, in one launch instead of ~25 element-wise PyTorch ops
__global__ void __launch_bounds__(256) quad_vertices_kernel(int P, const float* __restrict__ means, const float2* __restrict__ scales,
                                                            const float4* __restrict__ rots, float* __restrict__ verts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float4 q = rots[i];
    const float inv = 1.0f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x *= inv; q.y *= inv; q.z *= inv; q.w *= inv;
    const Rot3 R = quat_to_rot(q);
    const float2 s = scales[i];
    const float3 mu = ld3(means, i), eu = (3.0f * s.x) * R.c0, ev = (3.0f * s.y) * R.c1;
    const float su[4] = {-1.0f, -1.0f, 1.0f, 1.0f}, sv[4] = {1.0f, -1.0f, 1.0f, -1.0f};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float3 v = mu + su[k] * eu + sv[k] * ev;
        float* o = verts + (size_t)(4 * i + k) * 3;
        o[0] = v.x; o[1] = v.y; o[2] = v.z;
    }
}
int glue_quad_vertices(int P, const float* means, const float* scales, const float* rots, float* verts, cudaStream_t stream) {
    if (P <= 0) return ENVGS_OK;
    ENVGS_TIMED("glue_quad_vertices", stream);
    quad_vertices_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means, reinterpret_cast<const float2*>(scales),
                                                              reinterpret_cast<const float4*>(rots), verts);
    ENVGS_LAUNCH_CHECK("glue_quad_vertices", false, stream);
    return ENVGS_OK;
}

}  // namespace envgs
