// optim.cu — the parameter update of the view-parallel step ("fused all-reduce + Adam"), first half: ONE pass over
// the flat parameter / gradient / moment buffers that applies Adam, the 1/world_size averaging of the all-reduced gradient and
// the zero-fill of the gradient buffer for the next step (torch: a scale pass + 4 multi-tensor Adam launches + a memset).
// Arithmetic follows torch.optim.Adam(fused=True, amsgrad=False, weight_decay=0, maximize=False):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
#include "common.cuh"

namespace envgs {

__global__ void __launch_bounds__(256) adam_step_kernel(size_t n4, size_t n, float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                                        float4* __restrict__ v, float lr_over_bc1, float inv_sqrt_bc2, float b1, float b2,
                                                        float eps, float grad_scale, int zero_grad) {
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= grad_scale;
        mm = b1 * mm + (1.0f - b1) * gg;
        vv = b2 * vv + (1.0f - b2) * gg * gg;
        pp -= lr_over_bc1 * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
    };
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = p[i], G = __ldcs(g + i), M = m[i], V = v[i];
        upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
        p[i] = P; m[i] = M; v[i] = V;
        if (zero_grad) g[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    // tail (n not a multiple of 4)
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = n4 * 4 + threadIdx.x;
        float* ps = reinterpret_cast<float*>(p); float* gs = reinterpret_cast<float*>(g);
        float* ms = reinterpret_cast<float*>(m); float* vs = reinterpret_cast<float*>(v);
        float P = ps[i], M = ms[i], V = vs[i];
        upd(P, gs[i], M, V);
        ps[i] = P; ms[i] = M; vs[i] = V;
        if (zero_grad) gs[i] = 0.0f;
    }
}

int adam_step_impl(size_t n, float* p, float* g, float* m, float* v, int step, float lr, float b1, float b2, float eps, float grad_scale,
                   int zero_grad, cudaStream_t stream) {
    if (n == 0) return ENVGS_OK;
    ENVGS_TIMED("adam_step", stream);
    const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t n4 = n / 4;
    size_t blocks = (n4 + 255) / 256;
    if (blocks > (size_t)sms * 16) blocks = (size_t)sms * 16;
    if (blocks == 0) blocks = 1;
    adam_step_kernel<<<(unsigned)blocks, 256, 0, stream>>>(n4, n, (float4*)p, (float4*)g, (float4*)m, (float4*)v, (float)(lr / bc1),
                                                           (float)(1.0 / sqrt(bc2)), b1, b2, eps, grad_scale, zero_grad);
    ENVGS_LAUNCH_CHECK("adam_step", false, stream);
    return ENVGS_OK;
}

}  // namespace envgs
