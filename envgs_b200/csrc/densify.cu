// densify.cu — device-side densification / pruning of a surfel set, sm_100a.
//
// The reference (easyvolcap/utils/gaussian2d_utils.py) runs densify_and_prune (:867-899) as ~30 boolean-mask gathers and
// torch.cat re-allocations per stage over 8 parameter tensors, their two Adam moments and 4 statistics buffers, with a
// `.item()` sync per stage.  Here one `plan` kernel evaluates the clone / split / prune rules of all three stages for every
// ORIGINAL surfel at once (every new entry's attributes are functions of its source row, so the stage order can be resolved
// per row), a four-lane scan assigns destination rows in the reference's final order, and one `apply` launch per array
// scatters rows (parameters, Adam moments, statistics) into freshly sized packed buffers:
//
//   final order = [kept originals | kept clones | children copy 0 (of split originals, then of split clones) | copy 1 | ...]
//
// which is exactly what clone (:718-730) -> split (:732-760, children appended block-wise by `.repeat(N, 1)`, parents pruned)
// -> prune_min_opacity_and_gradients (:785-818) produce.  Statistics follow densify_stats (:689-703).
#include <cub/cub.cuh>
#include "common.cuh"

namespace envgs {

struct Quad { uint32_t a, b, c, d; };
struct QuadSum {
    __host__ __device__ Quad operator()(const Quad& x, const Quad& y) const { return Quad{x.a + y.a, x.b + y.b, x.c + y.c, x.d + y.d}; }
};

// add_densification_stats (:901-911) + the max_radii2D update of envgs_sampler.py:250-253
__global__ void __launch_bounds__(256) densify_stats_kernel(int P, const float* __restrict__ vs_grad, const int* __restrict__ radii,
                                                            const float* __restrict__ weight, float* __restrict__ grad_accum,
                                                            float* __restrict__ denom, float* __restrict__ weight_accum,
                                                            float* __restrict__ max_radii2D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || !(radii[i] > 0)) return;
    denom[i] += 1.0f;
    const float gx = vs_grad[3 * (size_t)i], gy = vs_grad[3 * (size_t)i + 1], gz = vs_grad[3 * (size_t)i + 2];
    grad_accum[i] += sqrtf(gx * gx + gy * gy + gz * gz);
    if (weight) weight_accum[i] += weight[i];
    if (max_radii2D) max_radii2D[i] = fmaxf(max_radii2D[i], (float)radii[i]);
}

__device__ __forceinline__ float avg_or_zero(float accum, float den) {
    const float v = accum / den;          // IEEE, like torch; NaN (0/0) -> 0 (get_xyz_gradient_avg, :661-664)
    return isnan(v) ? 0.0f : v;
}

// flags per original row: bit0 keep self, bit1 keep clone, bit2 children of the original, bit3 children of the clone
__global__ void __launch_bounds__(256) densify_flags_kernel(envgs_densify_params p, const float* __restrict__ scaling_act,
                                                            const float* __restrict__ opacity_act, const float* __restrict__ grad_accum,
                                                            const float* __restrict__ denom, const float* __restrict__ max_radii2D,
                                                            Quad* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.P) return;
    const float den = denom[i], acc = grad_accum[i];
    const float g = avg_or_zero(acc, den);
    const bool high = g >= p.grad_threshold;
    const float smax = fmaxf(scaling_act[2 * (size_t)i], scaling_act[2 * (size_t)i + 1]);
    const bool small = smax <= p.size_threshold;
    const bool radbig = p.split_screen_threshold >= 0.0f && max_radii2D[i] > p.split_screen_threshold;
    const bool c = small && high;                                  // densify_and_clone (:762-777)
    const bool s_o = ((smax > p.size_threshold) || radbig) && high;   // densify_and_split (:779-800), evaluated after the clone
    const bool s_c = c && radbig;                                   // the clone carries its source's statistics
    const bool occ = p.min_opacity >= 0.0f && opacity_act[i] < p.min_opacity;   // prune_min_opacity_and_gradients (:802-836)
    const float r = 1.0f / (p.split_ratio * (float)p.n_split);
    const float g_child = avg_or_zero(acc * r, den);
    const bool mg = p.min_gradient >= 0.0f && g <= p.min_gradient && den != 0.0f;
    const bool mg_child = p.min_gradient >= 0.0f && g_child <= p.min_gradient && den != 0.0f;
    Quad f;
    f.a = (!s_o && !(occ || mg)) ? 1u : 0u;
    f.b = (c && !s_c && !(occ || mg)) ? 1u : 0u;
    f.c = (s_o && !(occ || mg_child)) ? 1u : 0u;
    f.d = (c && s_c && !(occ || mg_child)) ? 1u : 0u;
    flags[i] = f;
}

// plan[0*P+i] = destination row of the original (or -1); [1*P+i] of its clone; [2*P+i] rank among split originals;
// [3*P+i] rank among split clones.  counts[4] = totals of the four kinds.
__global__ void __launch_bounds__(256) densify_plan_kernel(int P, const Quad* __restrict__ flags, const Quad* __restrict__ scan,
                                                           int32_t* __restrict__ plan, int32_t* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const Quad f = flags[i], s = scan[i];
    const int n_self = (int)(scan[P - 1].a + flags[P - 1].a);
    plan[i] = f.a ? (int)s.a : -1;
    plan[(size_t)P + i] = f.b ? n_self + (int)s.b : -1;
    plan[2 * (size_t)P + i] = f.c ? (int)s.c : -1;
    plan[3 * (size_t)P + i] = f.d ? (int)s.d : -1;
    if (i == P - 1) { counts[0] = n_self; counts[1] = (int)(s.b + f.b); counts[2] = (int)(s.c + f.c); counts[3] = (int)(s.d + f.d); }
}

// One launch per array: element e = (row i, column c) of the source goes to every destination row the plan assigns to i.
__global__ void __launch_bounds__(256) densify_apply_kernel(int P, int n_split, float split_ratio, const int32_t* __restrict__ plan,
                                                            int n_self, int n_clone, int n_co, int n_cc, envgs_row_array arr,
                                                            const float* __restrict__ xyz, const float* __restrict__ scaling_act,
                                                            const float* __restrict__ rotation, const float* __restrict__ noise,
                                                            const float* __restrict__ weight_max) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = arr.width;
    if (e >= (size_t)P * w) return;
    const int i = (int)(e / w), c = (int)(e % w);
    const float v = arr.src[e];
    const int d_self = plan[i], d_clone = plan[(size_t)P + i], r_o = plan[2 * (size_t)P + i], r_c = plan[3 * (size_t)P + i];
    const float r = 1.0f / (split_ratio * (float)n_split);
    // weight statistics of appended entries are scaled by the running maximum (densify_stats :693): m0 over the originals for
    // the clones, m1 = max over originals and clones for the children
    const float m0 = weight_max ? weight_max[0] : 1.0f, m1 = weight_max ? weight_max[1] : 1.0f;
    if (d_self >= 0) arr.dst[(size_t)d_self * w + c] = v;
    if (d_clone >= 0) {
        float o = v;
        if (arr.role == ENVGS_ROW_OPT_STATE) o = 0.0f;               // cat_tensors_to_optimizer appends zero moments (:573-574)
        else if (arr.role == ENVGS_ROW_STAT_WEIGHT) o = v * m0;
        arr.dst[(size_t)d_clone * w + c] = o;
    }
    if (r_o < 0 && r_c < 0) return;
    for (int k = 0; k < n_split; k++) {
        const size_t base = (size_t)n_self + n_clone + (size_t)k * (n_co + n_cc);
#pragma unroll
        for (int from_clone = 0; from_clone < 2; from_clone++) {
            const int rank = from_clone ? r_c : r_o;
            if (rank < 0) continue;
            const size_t d = base + (from_clone ? n_co : 0) + rank;
            float o = v;
            switch (arr.role) {
                case ENVGS_ROW_XYZ: {    // new_xyz = R(q) (n0 s_u, n1 s_v, 0) + xyz  (split :733-739, build_rotation :143-178)
                    const float4 q4 = reinterpret_cast<const float4*>(rotation)[i];
                    const float qn = sqrtf(q4.x * q4.x + q4.y * q4.y + q4.z * q4.z + q4.w * q4.w);
                    const float qr = q4.x / qn, qx = q4.y / qn, qy = q4.z / qn, qz = q4.w / qn;
                    const float* nz = noise + (((size_t)k * 2 + from_clone) * P + i) * 2;
                    const float a = nz[0] * scaling_act[2 * (size_t)i], b = nz[1] * scaling_act[2 * (size_t)i + 1];
                    float R0, R1;
                    if (c == 0) { R0 = 1.0f - 2.0f * (qy * qy + qz * qz); R1 = 2.0f * (qx * qy - qr * qz); }
                    else if (c == 1) { R0 = 2.0f * (qx * qy + qr * qz); R1 = 1.0f - 2.0f * (qx * qx + qz * qz); }
                    else { R0 = 2.0f * (qx * qz - qr * qy); R1 = 2.0f * (qy * qz + qr * qx); }
                    o = (R0 * a + R1 * b) + v;
                    break;
                }
                case ENVGS_ROW_SCALING: o = logf(scaling_act[2 * (size_t)i + c] * r); break;   // torch divides by a host scalar as a multiply by its fp32 inverse
                case ENVGS_ROW_OPT_STATE: o = 0.0f; break;
                case ENVGS_ROW_STAT_GRAD: o = v * r; break;
                case ENVGS_ROW_STAT_RADII: o = v * r; break;
                case ENVGS_ROW_STAT_WEIGHT: o = (from_clone ? v * m0 : v) * m1; break;
                default: break;
            }
            arr.dst[d * w + c] = o;
        }
    }
}

__global__ void __launch_bounds__(256) rows_gather_kernel(size_t M, const int64_t* __restrict__ index, envgs_row_array arr) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = arr.width;
    if (e >= M * w) return;
    const size_t j = e / w;
    arr.dst[e] = arr.src[(size_t)index[j] * w + (e % w)];
}

}  // namespace envgs

using namespace envgs;

extern "C" {

int envgs_densify_stats(int32_t P, const float* viewspace_grad, const int32_t* radii, const float* weight, float* grad_accum,
                        float* denom, float* weight_accum, float* max_radii2D, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (P <= 0) return ENVGS_OK;
    if (!viewspace_grad || !radii || !grad_accum || !denom || (weight && !weight_accum)) {
        set_error("envgs_densify_stats: null argument");
        return ENVGS_ERR_INVALID;
    }
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, viewspace_grad, radii, weight, grad_accum, denom, weight_accum, max_radii2D);
    ENVGS_LAUNCH_CHECK("densify_stats", false, stream);
    return ENVGS_OK;
}

size_t envgs_densify_plan_scratch_bytes(int32_t P) {
    size_t scan = 0;
    cub::DeviceScan::ExclusiveScan(nullptr, scan, (Quad*)nullptr, (Quad*)nullptr, QuadSum(), Quad{0, 0, 0, 0}, P);
    return align_up(scan, 256) + 2 * align_up((size_t)P * sizeof(Quad), 256) + 256;
}

/* Evaluates the three stages for every row and writes the destination plan (4*P int32) and the four totals (device, 4 int32);
 * the totals are also copied to `counts_host` (this call synchronises the stream: the caller sizes the new buffers from them). */
int envgs_densify_plan(const envgs_densify_params* p, const float* scaling_act, const float* opacity_act, const float* grad_accum,
                       const float* denom, const float* max_radii2D, int32_t* plan, int32_t* counts_dev, int32_t counts_host[4],
                       void* scratch, size_t scratch_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int P = p->P;
    if (P <= 0) { for (int k = 0; k < 4; k++) counts_host[k] = 0; return ENVGS_OK; }
    if (p->n_split < 1 || p->n_split > 8) { set_error("envgs_densify_plan: n_split must be in [1, 8]"); return ENVGS_ERR_INVALID; }
    if (scratch_bytes < envgs_densify_plan_scratch_bytes(P)) { set_error("envgs_densify_plan: scratch too small"); return ENVGS_ERR_INVALID; }
    char* s = (char*)scratch;
    Quad* flags = (Quad*)s; s += align_up((size_t)P * sizeof(Quad), 256);
    Quad* scan = (Quad*)s; s += align_up((size_t)P * sizeof(Quad), 256);
    size_t scan_bytes = scratch_bytes - (size_t)(s - (char*)scratch);
    densify_flags_kernel<<<(P + 255) / 256, 256, 0, stream>>>(*p, scaling_act, opacity_act, grad_accum, denom, max_radii2D, flags);
    ENVGS_LAUNCH_CHECK("densify_flags", false, stream);
    ENVGS_CUDA_TRY(cub::DeviceScan::ExclusiveScan(s, scan_bytes, flags, scan, QuadSum(), Quad{0, 0, 0, 0}, P, stream));
    count_launch(2);
    densify_plan_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, flags, scan, plan, counts_dev);
    ENVGS_LAUNCH_CHECK("densify_plan", false, stream);
    ENVGS_CUDA_TRY(cudaMemcpyAsync(counts_host, counts_dev, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    ENVGS_CUDA_TRY(cudaStreamSynchronize(stream));
    return ENVGS_OK;
}

int envgs_densify_apply(const envgs_densify_params* p, const int32_t* plan, const int32_t counts[4], int32_t n_arrays,
                        const envgs_row_array* arrays, const float* xyz, const float* scaling_act, const float* rotation,
                        const float* noise, const float* weight_max, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int P = p->P;
    if (P <= 0) return ENVGS_OK;
    for (int a = 0; a < n_arrays; a++) {
        const envgs_row_array& arr = arrays[a];
        if (!arr.src || !arr.dst || arr.width <= 0) { set_error("envgs_densify_apply: array %d is invalid", a); return ENVGS_ERR_INVALID; }
        if ((arr.role == ENVGS_ROW_XYZ && (arr.width != 3 || !rotation || !scaling_act || (!noise && counts[2] + counts[3] > 0))) ||
            (arr.role == ENVGS_ROW_SCALING && (arr.width != 2 || !scaling_act))) {
            set_error("envgs_densify_apply: array %d lacks the inputs its role needs", a);
            return ENVGS_ERR_INVALID;
        }
        const size_t n = (size_t)P * arr.width;
        densify_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(P, p->n_split, p->split_ratio, plan, counts[0], counts[1],
                                                                              counts[2], counts[3], arr, xyz, scaling_act, rotation, noise,
                                                                              weight_max);
        ENVGS_LAUNCH_CHECK("densify_apply", false, stream);
    }
    return ENVGS_OK;
}

/* dst[a][j, :] = src[a][index[j], :] for every array: prune_points / _prune_optimizer / prune_stats (:538-559, :676-687) in one
 * launch per array, for masks that need a global order statistic first (prune_visibility :849-865). */
int envgs_rows_gather(int64_t M, const int64_t* index, int32_t n_arrays, const envgs_row_array* arrays, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (M <= 0) return ENVGS_OK;
    for (int a = 0; a < n_arrays; a++) {
        const envgs_row_array& arr = arrays[a];
        if (!arr.src || !arr.dst || arr.width <= 0) { set_error("envgs_rows_gather: array %d is invalid", a); return ENVGS_ERR_INVALID; }
        const size_t n = (size_t)M * arr.width;
        rows_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((size_t)M, index, arr);
        ENVGS_LAUNCH_CHECK("rows_gather", false, stream);
    }
    return ENVGS_OK;
}

}  // extern "C"
