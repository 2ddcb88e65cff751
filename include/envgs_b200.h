/*
 * envgs_b200.h — C ABI of libenvgs_b200.so: the B200-native (sm_100a) replacement for the two native
 * submodules of zju3dv/EnvGS' per-view rendering hot path.
 *
 *   surfel rasterizer  (replaces diff-surfel-rasterization-wet[-ch05|-ch07], pybind11 module `_C`)
 *   surfel tracer      (replaces diff-surfel-tracing: OptiX GAS + raygen/anyhit programs)
 *
 * Every pointer is caller-owned DEVICE memory (contiguous fp32 / int32), every `stream` is a cudaStream_t
 * passed as void*.  No torch types cross this boundary.  Functions return 0 (or a non-negative count) on
 * success and a negative code on failure; envgs_last_error() then holds a message.  Nothing here
 * synchronises the stream except the 4-byte read-back of the instance count in envgs_raster_forward
 * (the reference does the same: rasterizer_impl.cu:283) — and not even that with envgs_raster_args.instance_capacity set.
 *
 * Reference interfaces replaced (paths relative to /root/reference/submodules):
 *   envgs_raster_forward   <- RasterizeGaussiansCUDA            diff-surfel-rasterizations/diff-surfel-rasterization-wet-ch05/rasterize_points.cu:39-136
 *                             CudaRasterizer::Rasterizer::forward  .../cuda_rasterizer/rasterizer_impl.cu:198-344
 *   envgs_raster_backward  <- RasterizeGaussiansBackwardCUDA    .../rasterize_points.cu:138-235, rasterizer_impl.cu:348-450
 *   envgs_mark_visible     <- markVisible                       .../rasterize_points.cu:237-256
 *   envgs_tracer_create/destroy <- OptiXStateWrapper            diff-surfel-tracing/optix_tracer/optix_wrapper.cpp:177-233
 *   envgs_tracer_build     <- BuildAccelerationStructure        diff-surfel-tracing/trace_surfels.cpp:47-170
 *   envgs_trace_forward    <- TraceSurfelsCUDA                  diff-surfel-tracing/trace_surfels.cpp:173-315
 *   envgs_trace_backward   <- TraceSurfelsBackwardCUDA          diff-surfel-tracing/trace_surfels.cpp:318-471
 */
#ifndef ENVGS_B200_H_
#define ENVGS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ENVGS_OK 0
#define ENVGS_ERR_INVALID -1   /* bad argument (shape, null pointer, unsupported channel count ...) */
#define ENVGS_ERR_CUDA -2      /* a CUDA runtime call failed */
#define ENVGS_ERR_ALLOC -3     /* an allocator callback returned NULL */
#define ENVGS_ERR_STATE -4     /* e.g. trace before build */

/* Scratch allocator: must return >= nbytes of device memory, 256-byte aligned, that stays valid until the
 * matching backward call finished (the reference uses resizable torch byte tensors the same way:
 * rasterize_points.cu:31-37). */
typedef void* (*envgs_alloc_fn)(void* user, size_t nbytes);

const char* envgs_last_error(void);
int envgs_version(void);

/* ------------------------------------------------------------------------------------------------ */
/* Rasterizer                                                                                         */
/* ------------------------------------------------------------------------------------------------ */

typedef struct envgs_raster_args {
    int32_t P;              /* number of surfels */
    int32_t C;              /* feature channels: 3 (wet), 5 (wet_ch05), 7 (wet_ch07) */
    int32_t sh_degree;      /* active SH degree D (0..3); used only when shs != NULL */
    int32_t sh_coeffs;      /* M = coefficients stored per surfel in `shs` */
    int32_t width, height;
    float tan_fovx, tan_fovy;
    float scale_modifier;
    int32_t prefiltered, debug;
    const float* background;       /* (C,)           */
    const float* means3D;          /* (P,3)          */
    const float* shs;              /* (P,M,3) or NULL (only legal when C == 3) */
    const float* colors_precomp;   /* (P,C)  or NULL */
    const float* opacities;        /* (P,1)          */
    const float* scales;           /* (P,2)  or NULL when transMat_precomp given */
    const float* rotations;        /* (P,4) wxyz or NULL */
    const float* transMat_precomp; /* (P,9)  or NULL */
    const float* viewmatrix;       /* (4,4) W2C^T, contiguous        */
    const float* projmatrix;       /* (4,4) (Proj*W2C)^T, contiguous */
    const float* campos;           /* (3,)           */
    /* Host-sync-free mode (0 / NULL = off).  With instance_capacity > 0 the forward call sizes the binning buffer for that many
     * (tile, surfel) instances, never reads the true count back (the reference blocks on it, rasterizer_impl.cu:283) and
     * returns instance_capacity — the value to hand to envgs_raster_backward as num_rendered.  The true count and an overflow
     * flag (count > capacity: the excess instances were dropped from this frame) are written to instance_count_out[0..1]
     * (device int32) for the caller to inspect later, e.g. after an asynchronous copy. */
    int32_t instance_capacity;
    int32_t* instance_count_out;
} envgs_raster_args;

typedef struct envgs_raster_out {
    float* out_color;   /* (C,H,W)  written                         */
    float* out_others;  /* (7,H,W)  written: [depth*a, a, nx,ny,nz, median depth, distortion] */
    float* out_weight;  /* (P,1)    accumulated (must be zeroed by caller) */
    int32_t* radii;     /* (P,)     written                         */
} envgs_raster_out;

/* Returns num_rendered (instances in the binned list, >= 0) or a negative error.  The three scratch
 * buffers are opaque; hand the same pointers to envgs_raster_backward. */
int envgs_raster_forward(const envgs_raster_args* args, const envgs_raster_out* out,
                         envgs_alloc_fn geom_alloc, envgs_alloc_fn binning_alloc, envgs_alloc_fn img_alloc,
                         void* alloc_user, void* stream);

typedef struct envgs_raster_grads {
    const float* dL_dout_color;   /* (C,H,W) */
    const float* dL_dout_others;  /* (7,H,W) */
    /* outputs, all must be zero-filled by the caller */
    float* dL_dmeans2D;    /* (P,3) densification signal (backward.cu:626-629) */
    float* dL_dcolors;     /* (P,C) */
    float* dL_dopacity;    /* (P,1) */
    float* dL_dmeans3D;    /* (P,3) */
    float* dL_dtransMat;   /* (P,9) */
    float* dL_dsh;         /* (P,M,3) or NULL */
    float* dL_dscales;     /* (P,2) */
    float* dL_drotations;  /* (P,4) */
    float* dL_dnormal;     /* (P,3) scratch */
} envgs_raster_grads;

int envgs_raster_backward(const envgs_raster_args* args, const int32_t* radii, int32_t num_rendered,
                          void* geom_buffer, void* binning_buffer, void* img_buffer,
                          const envgs_raster_grads* grads, void* stream);

/* present[i] = view-space z of means3D[i] > 0.2 (auxiliary.h:185-210) */
int envgs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                       uint8_t* present, void* stream);

/* sizes of the fixed-size scratch regions, for callers that pre-allocate */
size_t envgs_raster_geom_bytes(int32_t P);
size_t envgs_raster_img_bytes(int32_t width, int32_t height);
size_t envgs_raster_binning_bytes(int32_t num_rendered);

/* ------------------------------------------------------------------------------------------------ */
/* Densification / pruning of a surfel set on the device.                     */
/* Replaces the mask-gather / torch.cat implementation of                                            */
/*   easyvolcap/utils/gaussian2d_utils.py:517-592 (optimizer state), :623-716 (statistics),           */
/*   :718-760 (clone / split), :762-836 (selection rules), :867-911 (densify_and_prune, statistics).  */
/* ------------------------------------------------------------------------------------------------ */
typedef struct envgs_densify_params {
    int32_t P;                    /* rows before the update */
    int32_t n_split;              /* children per split surfel (densify_and_split: 2) */
    float split_ratio;            /* 0.8: child scale = scale / (ratio * n_split) */
    float grad_threshold;         /* densify_grad_threshold */
    float size_threshold;         /* densify_size_threshold * spatial_scale */
    float split_screen_threshold; /* < 0: disabled (None) */
    float min_opacity;            /* < 0: disabled */
    float min_gradient;           /* < 0: disabled */
} envgs_densify_params;

enum { ENVGS_ROW_PLAIN = 0, ENVGS_ROW_XYZ = 1, ENVGS_ROW_SCALING = 2, ENVGS_ROW_OPT_STATE = 4, ENVGS_ROW_STAT_GRAD = 5,
       ENVGS_ROW_STAT_DENOM = 6, ENVGS_ROW_STAT_RADII = 7, ENVGS_ROW_STAT_WEIGHT = 8 };
typedef struct envgs_row_array {  /* one (rows, width) fp32 array that follows the surfel rows */
    const float* src;
    float* dst;
    int32_t width;
    int32_t role;                 /* ENVGS_ROW_*: how new rows (clones, split children) are derived from the source row */
} envgs_row_array;

/* add_densification_stats (:901-911) + max_radii2D update (envgs_sampler.py:250-253) for rows with radii > 0 */
int envgs_densify_stats(int32_t P, const float* viewspace_grad, const int32_t* radii, const float* weight, float* grad_accum,
                        float* denom, float* weight_accum, float* max_radii2D, void* stream);
size_t envgs_densify_plan_scratch_bytes(int32_t P);
/* clone -> split -> prune_min_opacity_and_gradients resolved per source row; plan = 4*P int32, counts = {kept originals, kept
 * clones, split originals, split clones}; new row count = counts[0] + counts[1] + n_split * (counts[2] + counts[3]).
 * Synchronises `stream` (the caller allocates the new buffers from counts_host). */
int envgs_densify_plan(const envgs_densify_params* p, const float* scaling_act, const float* opacity_act, const float* grad_accum,
                       const float* denom, const float* max_radii2D, int32_t* plan, int32_t* counts_dev, int32_t counts_host[4],
                       void* scratch, size_t scratch_bytes, void* stream);
/* scatter every array's rows to their planned destinations (noise: (n_split, 2, P, 2) standard normals; weight_max: device
 * {max weight over the originals, max over originals and clones} or NULL) */
int envgs_densify_apply(const envgs_densify_params* p, const int32_t* plan, const int32_t counts[4], int32_t n_arrays,
                        const envgs_row_array* arrays, const float* xyz, const float* scaling_act, const float* rotation,
                        const float* noise, const float* weight_max, void* stream);
/* dst[a][j,:] = src[a][index[j],:] for all arrays (prune_points + _prune_optimizer + prune_stats for an index list) */
int envgs_rows_gather(int64_t M, const int64_t* index, int32_t n_arrays, const envgs_row_array* arrays, void* stream);

/* ------------------------------------------------------------------------------------------------ */
/* Tracer                                                                                             */
/* ------------------------------------------------------------------------------------------------ */

typedef struct envgs_tracer envgs_tracer;

envgs_tracer* envgs_tracer_create(void);
void envgs_tracer_destroy(envgs_tracer* h);

/* Build (rebuild != 0) or refit (rebuild == 0, same P as the last build) the LBVH over the P surfel quads
 * given as 4P vertices in the reference's order [(-,+),(-,-),(+,+),(+,-)] (optix_utils.py:39-69).  BVH
 * storage comes from `alloc` (called at most once per build when capacity must grow; the handle keeps the
 * pointer).  Pass alloc == NULL to let the library cudaMallocAsync/cudaFreeAsync on `stream`. */
int envgs_tracer_build(envgs_tracer* h, const float* vertices, int32_t num_vertices, int32_t rebuild,
                       envgs_alloc_fn alloc, void* alloc_user, void* stream);
size_t envgs_tracer_bvh_bytes(int32_t P);

/* Debug aid (synchronises the device): read and clear the traversal counters
 * [rays, nodes visited, quad tests, hits stored, 16-hit chunks, quad hits found (fast path) / contributions (general kernel), 0, 0]
 * and switch counting on/off. */
int envgs_debug_trace_stats(envgs_tracer* h, int32_t enable, uint64_t out[8]);

typedef struct envgs_trace_args {
    int32_t P;          /* surfels (must equal the built P) */
    int32_t H, W;       /* rays are (H,W,3) */
    int32_t sh_degree, sh_coeffs;
    int32_t training;   /* accumulate per-surfel weights */
    int32_t start_from_first;
    int32_t max_trace_depth;     /* 0 or 1 (MAX_TRACE_DEPTH, config.h:38) */
    float specular_threshold;
    float scale_modifier;
    const float* ray_o;          /* (H,W,3) */
    const float* ray_d;          /* (H,W,3) */
    const float* background;     /* (3,) */
    const float* means3D;        /* (P,3) */
    const float* shs;            /* (P,M,3) or NULL */
    const float* colors_precomp; /* (P,3) or NULL */
    const float* others_precomp; /* (P,2) or NULL */
    const float* opacities;      /* (P,1) */
    const float* scales;         /* (P,2) */
    const float* rotations;      /* (P,4) */
    const float* viewmatrix;     /* (4,4) */
    const float* projmatrix;     /* (4,4) */
    const float* campos;         /* (3,) */
} envgs_trace_args;

typedef struct envgs_trace_out {   /* all written for every ray; a_weights accumulated (zeroed by caller) */
    float* out_rgb;   /* (H,W,3)  */
    float* out_dpt;   /* (H,W,1)  */
    float* out_acc;   /* (H,W,1)  */
    float* out_norm;  /* (H,W,3)  */
    float* out_dist;  /* (H,W,3)  [distortion, M1, M2] */
    float* out_aux;   /* (H,W,2)  */
    float* mid_val;   /* (H,W,32) fully written */
    float* a_weights; /* (P,1) or NULL when !training */
    void* hit_cache;  /* optional scratch of envgs_trace_cache_bytes2(H, W, max_trace_depth) bytes: enables the hit-parallel fast
                         path for !start_from_first (one segment, or two segments traced one after the other when
                         max_trace_depth == 1 and others_precomp is given) and carries the per-ray hit lists to the backward
                         call; NULL selects the general (monolithic) kernel */
} envgs_trace_out;

size_t envgs_trace_cache_bytes(int32_t H, int32_t W);                              /* = envgs_trace_cache_bytes2(H, W, 0) */
size_t envgs_trace_cache_bytes2(int32_t H, int32_t W, int32_t max_trace_depth);

int envgs_trace_forward(envgs_tracer* h, const envgs_trace_args* args, const envgs_trace_out* out, void* stream);

typedef struct envgs_trace_grads {
    /* saved forward outputs */
    const float* out_rgb; const float* out_dpt; const float* out_acc; const float* out_norm;
    const float* out_dist; const float* out_aux; const float* mid_val;
    /* upstream */
    const float* dL_drgb;  /* (H,W,3) */
    const float* dL_ddpt;  /* (H,W,1) */
    const float* dL_dacc;  /* (H,W,1) */
    const float* dL_dnorm; /* (H,W,3) */
    const float* dL_ddist; /* (H,W,1) */
    const float* dL_daux;  /* (H,W,2) */
    /* outputs: ray grads written, per-surfel grads accumulated (zeroed by caller) */
    float* dL_dray_o;      /* (H,W,3) */
    float* dL_dray_d;      /* (H,W,3) */
    float* dL_dmeans3D;    /* (P,3) */
    float* dL_dshs;        /* (P,M,3) or NULL */
    float* dL_dcolors;     /* (P,3) */
    float* dL_dothers;     /* (P,2) or NULL */
    float* dL_dopacities;  /* (P,1) */
    float* dL_dscales;     /* (P,2) */
    float* dL_drotations;  /* (P,4) */
    void* hit_cache;       /* the buffer the forward call filled, or NULL */
} envgs_trace_grads;

int envgs_trace_backward(envgs_tracer* h, const envgs_trace_args* args, const envgs_trace_grads* grads,
                         void* stream);

/* ------------------------------------------------------------------------------------------------ */
/* Caller-side glue: base-surfel features for the C = 5 / 7 rasterizer          */
/* ------------------------------------------------------------------------------------------------ */
/* out (P, 3+E) = [ max(SH_D(normalize(mean - campos)) + 0.5, 0) | extra (P,E) ]  — what EasyVolcap computes in PyTorch
 * at easyvolcap/utils/gaussian2d_utils.py:1071-1084 (eval_sh, clamp_min, cat[specular, roughness]). */
int envgs_sh_features_forward(int32_t P, int32_t sh_degree, int32_t sh_coeffs, int32_t E, const float* means3D, const float* shs,
                              const float* extra, const float* campos, float* out, void* stream);
/* all three gradient outputs are fully written (no zero-fill needed) */
int envgs_sh_features_backward(int32_t P, int32_t sh_degree, int32_t sh_coeffs, int32_t E, const float* means3D, const float* shs,
                               const float* campos, const float* dL_dout, float* dL_dmeans3D, float* dL_dshs, float* dL_dextra,
                               void* stream);

/* View-parallel training: one view's SH gradient is rank one — basis(normalize(mean - campos)) (sh_coeffs values) times the
 * clamp-masked colour gradient (3 values).  `_factored` writes only that (P,3) factor instead of the (P,sh_coeffs,3) gradient;
 * envgs_sh_adam_factored rebuilds sum over n_views of basis (x) factor from the all-gathered factors (n_views,P,3) and camera
 * centres (n_views,3) in fixed view order and applies the Adam update (arithmetic of envgs_adam_step) to the coefficients.
 * `means3D` must be the positions the forward pass saw (call it before the positions are updated). */
int envgs_sh_features_backward_factored(int32_t P, int32_t sh_degree, int32_t sh_coeffs, int32_t E, const float* means3D, const float* shs,
                                        const float* campos, const float* dL_dout, float* dL_dmeans3D, float* dL_dfactor, float* dL_dextra,
                                        void* stream);
int envgs_sh_adam_factored(int32_t P, int32_t sh_degree, int32_t sh_coeffs, int32_t n_views, const float* means3D, const float* campos_all,
                           const float* factors_all, float* shs, float* exp_avg, float* exp_avg_sq, int32_t step, float lr, float beta1,
                           float beta2, float eps, float grad_scale, void* stream);

/* Per-pixel glue between the two operators, fused (reference: ~150 eager PyTorch launches per view —
 * gaussian2d_utils.py:1117-1153 + :1158-1206, ray_utils.py:24-78, envgs_sampler.py:420-431 + :474, plus the loss). */
typedef struct envgs_glue_camera {
    int32_t H, W;
    float tanfovx, tanfovy;
    const float* K;       /* (3,3) row-major, device: intrinsics of the camera rays (pixel centres at +0.5) */
    const float* R;       /* (3,3) row-major, device: world->camera rotation */
    const float* campos;  /* (3,)  device */
} envgs_glue_camera;

/* Environment surfels -> (4P,3) quad corners at +-3 sigma in the order [(-,+),(-,-),(+,+),(+,-)]: the vertex buffer of
 * envgs_tracer_build.  Replaces the PyTorch expressions of HardwareRendering.get_disks (EVC/utils/optix_utils.py:39-69);
 * rotations are normalised first, as there.  No gradient (the reference builds the GAS under no_grad). */
int envgs_glue_quad_vertices(int32_t P, const float* means3D, const float* scales, const float* rotations, float* vertices, void* stream);

/* allmap (7,H,W) -> reflected rays ref_o, ref_d (H,W,3):  depth = nan_to_num(D/alpha), n = R^T n_view / (|.|+1e-8),
 * d' = d - 2 (d.n) n, o' = campos + depth d */
int envgs_glue_reflect_forward(const envgs_glue_camera* cam, const float* allmap, float* ref_o, float* ref_d, void* stream);
/* writes all 7 channels of dL_dallmap */
int envgs_glue_reflect_backward(const envgs_glue_camera* cam, const float* allmap, const float* dL_dref_o, const float* dL_dref_d,
                                float* dL_dallmap, void* stream);
/* rgb (H,W,3) = (1-spec) base + spec env (env_rgb may be NULL: rgb = base);
 * loss2[0] = mean|rgb - target|, loss2[1] = lambda_n * mean(1 - <n_world, alpha * dpt2norm(depth)>) */
int envgs_glue_blend_loss_forward(const envgs_glue_camera* cam, const float* image, const float* allmap, const float* env_rgb,
                                  const float* target, float lambda_n, float* rgb_out, float* loss2, void* stream);
/* g_loss: (1,) upstream gradient of loss2[0]+loss2[1]; g_rgb optional (H,W,3); all outputs fully written */
int envgs_glue_blend_loss_backward(const envgs_glue_camera* cam, const float* image, const float* allmap, const float* env_rgb,
                                   const float* target, const float* rgb, float lambda_n, const float* g_loss, const float* g_rgb,
                                   float* dL_dimage, float* dL_dallmap, float* dL_denv, void* stream);

/* Parameter update of the view-parallel step over FLAT fp32 buffers (n elements each, 16-byte aligned): Adam with bias
 * correction at step `step` (1-based), the gradient pre-multiplied by `grad_scale` (1/world_size after the sum all-reduce) and,
 * if zero_grad != 0, the gradient buffer cleared for the next step — one pass instead of torch's scale + multi-tensor Adam +
 * memset.  Same arithmetic as torch.optim.Adam(fused=True) without weight decay / amsgrad.  Replaces: the optimizer step the
 * reference runs through torch (EVC/runners/optimizers.py:163-201, volumetric_video_runner.py:443). */
int envgs_adam_step(uint64_t n, float* params, float* grads, float* exp_avg, float* exp_avg_sq, int32_t step, float lr, float beta1,
                    float beta2, float eps, float grad_scale, int32_t zero_grad, void* stream);

/* Kernel-launch counter (every kernel this library launches increments it); bench.py reports it. */
uint64_t envgs_launch_count(void);

/* Per-kernel device timing (CUDA events recorded on the launching stream around each kernel / library primitive).
 * Off by default.  envgs_profile_report synchronises the recorded events, writes one line per kernel name
 * ("name count total_ms\n") into buf (NUL-terminated, truncated to n) and clears the records; returns the
 * number of bytes the full report needs. */
void envgs_profile_enable(int on);
size_t envgs_profile_report(char* buf, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* ENVGS_B200_H_ */
