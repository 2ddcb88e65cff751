// capi.cu — the extern "C" surface declared in include/envgs_b200.h.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <map>
#include <string>
#include <vector>
#include <string.h>
#include "lbvh.cuh"
#include "raster_state.cuh"

namespace envgs {

static thread_local char g_error[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

struct ProfRecord { const char* name; cudaEvent_t start, stop; };
static bool g_profile = false;
static std::mutex g_prof_mutex;
static std::vector<ProfRecord> g_prof_records;

KernelTimer::KernelTimer(const char* name, cudaStream_t s) : slot(-1), stream(s) {
    if (!g_profile) return;
    ProfRecord r;
    r.name = name;
    if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
    cudaEventRecord(r.start, s);
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof_records.push_back(r);
    slot = (int)g_prof_records.size() - 1;
}
KernelTimer::~KernelTimer() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    cudaEventRecord(g_prof_records[slot].stop, stream);
}

int raster_forward_dispatch(const envgs_raster_args& a, const envgs_raster_out& o, envgs_alloc_fn ga, envgs_alloc_fn ba,
                            envgs_alloc_fn ia, void* user, cudaStream_t stream);
int raster_backward_dispatch(const envgs_raster_args& a, const int* radii, int R, void* geom, void* binning, void* img,
                             const envgs_raster_grads& gr, cudaStream_t stream);
int mark_visible_impl(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t stream);
int trace_forward_impl(Lbvh& bvh, unsigned int* counter, unsigned long long* stats, const envgs_trace_args& a, const envgs_trace_out& o, cudaStream_t stream);
int trace_backward_impl(Lbvh& bvh, unsigned int* counter, unsigned long long* stats, const envgs_trace_args& a, const envgs_trace_grads& g, cudaStream_t stream);
size_t trace_cache_bytes(int H, int W);
size_t trace_cache_bytes2(int H, int W, int max_trace_depth);
int glue_quad_vertices(int P, const float* means, const float* scales, const float* rots, float* verts, cudaStream_t stream);
int glue_reflect_forward(const envgs_glue_camera& cam, const float* allmap, float* ref_o, float* ref_d, cudaStream_t stream);
int glue_reflect_backward(const envgs_glue_camera& cam, const float* allmap, const float* g_ro, const float* g_rd, float* dL_dallmap, cudaStream_t stream);
int glue_blend_loss_forward(const envgs_glue_camera& cam, const float* image, const float* allmap, const float* env_rgb, const float* target, float lambda_n, float* rgb_out, float* loss2, cudaStream_t stream);
int glue_blend_loss_backward(const envgs_glue_camera& cam, const float* image, const float* allmap, const float* env_rgb, const float* target, const float* rgb, float lambda_n, const float* g_loss, const float* g_rgb, float* dL_dimage, float* dL_dallmap, float* dL_denv, cudaStream_t stream);
int adam_step_impl(size_t n, float* p, float* g, float* m, float* v, int step, float lr, float b1, float b2, float eps, float grad_scale,
                   int zero_grad, cudaStream_t stream);
int sh_features_forward_impl(int P, int D, int M, int E, const float* means, const float* shs, const float* extra, const float* campos, float* out, cudaStream_t stream);
int sh_features_backward_impl(int P, int D, int M, int E, const float* means, const float* shs, const float* campos, const float* dL_dout, float* dL_dmeans, float* dL_dshs, float* dL_dextra, float* dL_dfactor, cudaStream_t stream);
int sh_adam_factored_impl(int P, int D, int M, int N, const float* means, const float* campos_all, const float* factors_all, float* shs, float* exp_avg, float* exp_avg_sq, int step, float lr, float b1, float b2, float eps, float grad_scale, cudaStream_t stream);
bool trace_fast_eligible(const envgs_trace_args& a, const void* cache);
struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; };
int trace_forward_fast(Lbvh& bvh, unsigned int* counter, unsigned long long* stats, const envgs_trace_args& a, const envgs_trace_out& o, void* cache, cudaStream_t stream, const SideStream& side);
int trace_backward_fast(Lbvh& bvh, unsigned int* counter, unsigned long long* stats, const envgs_trace_args& a, const envgs_trace_grads& g, void* cache, cudaStream_t stream, const SideStream& side);

}  // namespace envgs

struct envgs_tracer {
    envgs::Lbvh bvh;
    unsigned int* counter = nullptr;       // [0]: persistent-warp work counter; stats live 64 bytes further
    bool stats_on = false;
    envgs::SideStream side{nullptr, nullptr, nullptr};
    unsigned long long* stats() { return stats_on ? reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(counter) + 64) : nullptr; }
};

using namespace envgs;

extern "C" {

const char* envgs_last_error(void) { return g_error; }
int envgs_version(void) { return 100; }
uint64_t envgs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void envgs_profile_enable(int on) { g_profile = on != 0; }

size_t envgs_profile_report(char* buf, size_t n) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    std::map<std::string, std::pair<int, double>> acc;
    for (auto& r : g_prof_records) {
        float ms = 0.0f;
        if (cudaEventSynchronize(r.stop) == cudaSuccess && cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
            auto& e = acc[r.name];
            e.first += 1;
            e.second += ms;
        }
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    g_prof_records.clear();
    std::string out;
    char line[256];
    for (auto& kv : acc) {
        snprintf(line, sizeof(line), "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (buf && n > 0) {
        size_t m = out.size() < n - 1 ? out.size() : n - 1;
        memcpy(buf, out.data(), m);
        buf[m] = 0;
    }
    return out.size() + 1;
}

static int check_raster_args(const envgs_raster_args* a) {
    if (!a) { set_error("args is NULL"); return ENVGS_ERR_INVALID; }
    if (a->P < 0 || a->width <= 0 || a->height <= 0) { set_error("bad sizes P=%d W=%d H=%d", a->P, a->width, a->height); return ENVGS_ERR_INVALID; }
    if (a->width > 32000 || a->height > 32000) { set_error("image larger than 32000 px is not supported"); return ENVGS_ERR_INVALID; }
    if (a->C != 3 && a->C != 5 && a->C != 7) { set_error("unsupported channel count C=%d (3, 5 or 7)", a->C); return ENVGS_ERR_INVALID; }
    if (a->P == 0) return ENVGS_OK;
    if (!a->means3D || !a->opacities || !a->background || !a->viewmatrix || !a->projmatrix || !a->campos) {
        set_error("a required pointer (means3D/opacities/background/viewmatrix/projmatrix/campos) is NULL");
        return ENVGS_ERR_INVALID;
    }
    if (a->colors_precomp == nullptr) {
        if (a->C != 3) { set_error("For non-RGB, provide precomputed Gaussian colors!"); return ENVGS_ERR_INVALID; }   // rasterizer_impl.cu:244-247
        if (a->shs == nullptr) { set_error("either shs or colors_precomp must be given"); return ENVGS_ERR_INVALID; }
        if (a->sh_degree < 0 || a->sh_degree > 3 || (a->sh_degree + 1) * (a->sh_degree + 1) > a->sh_coeffs) {
            set_error("sh_degree %d does not fit %d coefficients", a->sh_degree, a->sh_coeffs);
            return ENVGS_ERR_INVALID;
        }
    }
    if (a->transMat_precomp == nullptr && (a->scales == nullptr || a->rotations == nullptr)) {
        set_error("either scales+rotations or transMat_precomp must be given");
        return ENVGS_ERR_INVALID;
    }
    return ENVGS_OK;
}

int envgs_raster_forward(const envgs_raster_args* args, const envgs_raster_out* out, envgs_alloc_fn geom_alloc,
                         envgs_alloc_fn binning_alloc, envgs_alloc_fn img_alloc, void* alloc_user, void* stream) {
    int rc = check_raster_args(args);
    if (rc != ENVGS_OK) return rc;
    if (args->P == 0) return 0;   // reference leaves the zero-filled outputs untouched (rasterize_points.cu:100)
    if (!out || !out->out_color || !out->out_others || !out->out_weight) { set_error("output pointer is NULL"); return ENVGS_ERR_INVALID; }
    if (!geom_alloc || !binning_alloc || !img_alloc) { set_error("allocator callback is NULL"); return ENVGS_ERR_INVALID; }
    return raster_forward_dispatch(*args, *out, geom_alloc, binning_alloc, img_alloc, alloc_user, (cudaStream_t)stream);
}

int envgs_raster_backward(const envgs_raster_args* args, const int32_t* radii, int32_t num_rendered, void* geom_buffer,
                          void* binning_buffer, void* img_buffer, const envgs_raster_grads* grads, void* stream) {
    int rc = check_raster_args(args);
    if (rc != ENVGS_OK) return rc;
    if (args->P == 0) return ENVGS_OK;
    if (!grads || !geom_buffer || !img_buffer || (num_rendered > 0 && !binning_buffer)) { set_error("buffer pointer is NULL"); return ENVGS_ERR_INVALID; }
    if (!grads->dL_dout_color || !grads->dL_dout_others || !grads->dL_dmeans2D || !grads->dL_dcolors || !grads->dL_dopacity ||
        !grads->dL_dmeans3D || !grads->dL_dtransMat || !grads->dL_dscales || !grads->dL_drotations || !grads->dL_dnormal) {
        set_error("gradient pointer is NULL");
        return ENVGS_ERR_INVALID;
    }
    if (args->shs && !grads->dL_dsh) { set_error("dL_dsh is NULL although shs were given"); return ENVGS_ERR_INVALID; }
    return raster_backward_dispatch(*args, radii, num_rendered, geom_buffer, binning_buffer, img_buffer, *grads, (cudaStream_t)stream);
}

int envgs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present,
                       void* stream) {
    (void)projmatrix;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) { set_error("bad arguments to envgs_mark_visible"); return ENVGS_ERR_INVALID; }
    return mark_visible_impl(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

size_t envgs_raster_geom_bytes(int32_t P) { return GeomState::carve(nullptr, P, nullptr, scan_temp_bytes(P)); }
size_t envgs_raster_img_bytes(int32_t width, int32_t height) { return ImageState::carve(nullptr, width, height, nullptr); }
size_t envgs_raster_binning_bytes(int32_t n) { return BinningState::carve(nullptr, n, nullptr, sort_temp_bytes(n), 7); }   // upper bound over C = 3, 5, 7

envgs_tracer* envgs_tracer_create(void) {
    envgs_tracer* h = new (std::nothrow) envgs_tracer();
    if (!h) { set_error("out of host memory"); return nullptr; }
    if (cudaMalloc(&h->counter, 256) != cudaSuccess || cudaMemset(h->counter, 0, 256) != cudaSuccess) {
        set_error("cudaMalloc failed for the tracer work counter: %s", cudaGetErrorString(cudaGetLastError()));
        delete h;
        return nullptr;
    }
    if (cudaStreamCreateWithFlags(&h->side.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->side.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->side.join, cudaEventDisableTiming) != cudaSuccess) {
        set_error("could not create the tracer's side stream: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(h->counter);
        delete h;
        return nullptr;
    }
    return h;
}

void envgs_tracer_destroy(envgs_tracer* h) {
    if (!h) return;
    if (h->bvh.owns_storage && h->bvh.storage) cudaFree(h->bvh.storage);
    if (h->counter) cudaFree(h->counter);
    if (h->side.stream) { cudaStreamSynchronize(h->side.stream); cudaStreamDestroy(h->side.stream); }
    if (h->side.fork) cudaEventDestroy(h->side.fork);
    if (h->side.join) cudaEventDestroy(h->side.join);
    delete h;
}

size_t envgs_tracer_bvh_bytes(int32_t P) { return lbvh_storage_bytes(P); }

size_t envgs_trace_cache_bytes(int32_t H, int32_t W) { return trace_cache_bytes(H, W); }
size_t envgs_trace_cache_bytes2(int32_t H, int32_t W, int32_t max_trace_depth) { return trace_cache_bytes2(H, W, max_trace_depth); }

static int check_glue_cam(const envgs_glue_camera* c) {
    if (!c || c->H <= 0 || c->W <= 0 || !c->K || !c->R || !c->campos) { set_error("bad glue camera"); return ENVGS_ERR_INVALID; }
    return ENVGS_OK;
}
int envgs_glue_quad_vertices(int32_t P, const float* means3D, const float* scales, const float* rotations, float* vertices, void* stream) {
    if (P < 0 || (P > 0 && (!means3D || !scales || !rotations || !vertices))) { set_error("bad quad_vertices arguments"); return ENVGS_ERR_INVALID; }
    if ((reinterpret_cast<uintptr_t>(scales) & 7) || (reinterpret_cast<uintptr_t>(rotations) & 15)) { set_error("scales / rotations must be 8 / 16-byte aligned"); return ENVGS_ERR_INVALID; }
    return glue_quad_vertices(P, means3D, scales, rotations, vertices, (cudaStream_t)stream);
}
int envgs_glue_reflect_forward(const envgs_glue_camera* cam, const float* allmap, float* ref_o, float* ref_d, void* stream) {
    int rc = check_glue_cam(cam);
    if (rc != ENVGS_OK) return rc;
    if (!allmap || !ref_o || !ref_d) { set_error("glue pointer is NULL"); return ENVGS_ERR_INVALID; }
    return glue_reflect_forward(*cam, allmap, ref_o, ref_d, (cudaStream_t)stream);
}
int envgs_glue_reflect_backward(const envgs_glue_camera* cam, const float* allmap, const float* g_ro, const float* g_rd, float* dL_dallmap,
                                void* stream) {
    int rc = check_glue_cam(cam);
    if (rc != ENVGS_OK) return rc;
    if (!allmap || !g_ro || !g_rd || !dL_dallmap) { set_error("glue pointer is NULL"); return ENVGS_ERR_INVALID; }
    return glue_reflect_backward(*cam, allmap, g_ro, g_rd, dL_dallmap, (cudaStream_t)stream);
}
int envgs_glue_blend_loss_forward(const envgs_glue_camera* cam, const float* image, const float* allmap, const float* env_rgb,
                                  const float* target, float lambda_n, float* rgb_out, float* loss2, void* stream) {
    int rc = check_glue_cam(cam);
    if (rc != ENVGS_OK) return rc;
    if (!image || !allmap || !target || !rgb_out || !loss2) { set_error("glue pointer is NULL"); return ENVGS_ERR_INVALID; }
    return glue_blend_loss_forward(*cam, image, allmap, env_rgb, target, lambda_n, rgb_out, loss2, (cudaStream_t)stream);
}
int envgs_glue_blend_loss_backward(const envgs_glue_camera* cam, const float* image, const float* allmap, const float* env_rgb,
                                   const float* target, const float* rgb, float lambda_n, const float* g_loss, const float* g_rgb,
                                   float* dL_dimage, float* dL_dallmap, float* dL_denv, void* stream) {
    int rc = check_glue_cam(cam);
    if (rc != ENVGS_OK) return rc;
    if (!image || !allmap || !target || !rgb || !g_loss || !dL_dimage || !dL_dallmap || (env_rgb && !dL_denv)) { set_error("glue pointer is NULL"); return ENVGS_ERR_INVALID; }
    return glue_blend_loss_backward(*cam, image, allmap, env_rgb, target, rgb, lambda_n, g_loss, g_rgb, dL_dimage, dL_dallmap, dL_denv, (cudaStream_t)stream);
}

static int check_features(int32_t P, int32_t D, int32_t M, int32_t E) {
    if (P < 0 || E < 0 || D < 0 || D > 3 || (D + 1) * (D + 1) > M || M > 16) { set_error("bad sh_features arguments P=%d D=%d M=%d E=%d", P, D, M, E); return ENVGS_ERR_INVALID; }
    return ENVGS_OK;
}
int envgs_sh_features_forward(int32_t P, int32_t D, int32_t M, int32_t E, const float* means3D, const float* shs, const float* extra,
                              const float* campos, float* out, void* stream) {
    int rc = check_features(P, D, M, E);
    if (rc != ENVGS_OK) return rc;
    if (P > 0 && (!means3D || !shs || !campos || !out || (E > 0 && !extra))) { set_error("sh_features pointer is NULL"); return ENVGS_ERR_INVALID; }
    return sh_features_forward_impl(P, D, M, E, means3D, shs, extra, campos, out, (cudaStream_t)stream);
}
int envgs_sh_features_backward(int32_t P, int32_t D, int32_t M, int32_t E, const float* means3D, const float* shs, const float* campos,
                               const float* dL_dout, float* dL_dmeans3D, float* dL_dshs, float* dL_dextra, void* stream) {
    int rc = check_features(P, D, M, E);
    if (rc != ENVGS_OK) return rc;
    if (P > 0 && (!means3D || !shs || !campos || !dL_dout || !dL_dmeans3D || !dL_dshs || (E > 0 && !dL_dextra))) { set_error("sh_features pointer is NULL"); return ENVGS_ERR_INVALID; }
    return sh_features_backward_impl(P, D, M, E, means3D, shs, campos, dL_dout, dL_dmeans3D, dL_dshs, dL_dextra, nullptr, (cudaStream_t)stream);
}

int envgs_sh_features_backward_factored(int32_t P, int32_t D, int32_t M, int32_t E, const float* means3D, const float* shs, const float* campos,
                                        const float* dL_dout, float* dL_dmeans3D, float* dL_dfactor, float* dL_dextra, void* stream) {
    int rc = check_features(P, D, M, E);
    if (rc != ENVGS_OK) return rc;
    if (P > 0 && (!means3D || !shs || !campos || !dL_dout || !dL_dmeans3D || !dL_dfactor || (E > 0 && !dL_dextra))) { set_error("sh_features pointer is NULL"); return ENVGS_ERR_INVALID; }
    return sh_features_backward_impl(P, D, M, E, means3D, shs, campos, dL_dout, dL_dmeans3D, nullptr, dL_dextra, dL_dfactor, (cudaStream_t)stream);
}

int envgs_sh_adam_factored(int32_t P, int32_t D, int32_t M, int32_t n_views, const float* means3D, const float* campos_all,
                           const float* factors_all, float* shs, float* exp_avg, float* exp_avg_sq, int32_t step, float lr, float beta1,
                           float beta2, float eps, float grad_scale, void* stream) {
    int rc = check_features(P, D, M, 0);
    if (rc != ENVGS_OK) return rc;
    if (n_views < 1 || step < 1) { set_error("bad sh_adam_factored arguments n_views=%d step=%d", n_views, step); return ENVGS_ERR_INVALID; }
    if (P > 0 && (!means3D || !campos_all || !factors_all || !shs || !exp_avg || !exp_avg_sq)) { set_error("sh_adam_factored pointer is NULL"); return ENVGS_ERR_INVALID; }
    return sh_adam_factored_impl(P, D, M, n_views, means3D, campos_all, factors_all, shs, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps,
                                 grad_scale, (cudaStream_t)stream);
}

int envgs_debug_trace_stats(envgs_tracer* h, int32_t enable, uint64_t out[8]) {
    if (!h) { set_error("tracer handle is NULL"); return ENVGS_ERR_INVALID; }
    ENVGS_CUDA_TRY(cudaDeviceSynchronize());
    if (out) ENVGS_CUDA_TRY(cudaMemcpy(out, reinterpret_cast<char*>(h->counter) + 64, 64, cudaMemcpyDeviceToHost));
    ENVGS_CUDA_TRY(cudaMemset(reinterpret_cast<char*>(h->counter) + 64, 0, 64));
    h->stats_on = enable != 0;
    return ENVGS_OK;
}

int envgs_tracer_build(envgs_tracer* h, const float* vertices, int32_t num_vertices, int32_t rebuild, envgs_alloc_fn alloc,
                       void* alloc_user, void* stream) {
    if (!h) { set_error("tracer handle is NULL"); return ENVGS_ERR_INVALID; }
    if (num_vertices < 0 || num_vertices % 4 != 0) { set_error("vertices must hold 4 corners per surfel (got %d)", num_vertices); return ENVGS_ERR_INVALID; }
    const int P = num_vertices / 4;
    if (P > 0 && !vertices) { set_error("vertices is NULL"); return ENVGS_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    if (P > h->bvh.capacity) {
        int cap = P + P / 8;   // head-room so densification does not reallocate every step
        size_t bytes = lbvh_storage_bytes(cap);
        char* mem = nullptr;
        if (alloc) {
            mem = (char*)alloc(alloc_user, bytes);
            if (!mem) { set_error("BVH allocator returned NULL for %zu bytes", bytes); return ENVGS_ERR_ALLOC; }
            if (h->bvh.owns_storage && h->bvh.storage) cudaFree(h->bvh.storage);
            h->bvh.owns_storage = false;
        } else {
            ENVGS_CUDA_TRY(cudaStreamSynchronize(s));
            if (h->bvh.owns_storage && h->bvh.storage) cudaFree(h->bvh.storage);
            ENVGS_CUDA_TRY(cudaMalloc(&mem, bytes));
            h->bvh.owns_storage = true;
        }
        h->bvh.storage = mem;
        h->bvh.capacity = cap;
        h->bvh.P = 0;
        rebuild = 1;
    }
    if (P > 0) lbvh_carve(h->bvh, h->bvh.storage, h->bvh.capacity);
    return lbvh_build(h->bvh, vertices, P, rebuild != 0, s);
}

static int check_trace_args(envgs_tracer* h, const envgs_trace_args* a) {
    if (!h) { set_error("tracer handle is NULL"); return ENVGS_ERR_INVALID; }
    if (!a) { set_error("args is NULL"); return ENVGS_ERR_INVALID; }
    if (a->H <= 0 || a->W <= 0) { set_error("bad ray grid %dx%d", a->H, a->W); return ENVGS_ERR_INVALID; }
    if (h->bvh.P != a->P || a->P <= 0) {
        set_error("acceleration structure holds %d surfels but %d were passed; call build_acceleration_structure first", h->bvh.P, a->P);
        return ENVGS_ERR_STATE;
    }
    if (a->max_trace_depth < 0 || a->max_trace_depth > TR_MAX_TRACE_DEPTH) { set_error("max_trace_depth must be 0..%d", TR_MAX_TRACE_DEPTH); return ENVGS_ERR_INVALID; }
    if (!a->ray_o || !a->ray_d || !a->background || !a->means3D || !a->opacities || !a->scales || !a->rotations || !a->campos ||
        !a->viewmatrix || !a->projmatrix) { set_error("a required tracer pointer is NULL"); return ENVGS_ERR_INVALID; }
    if ((a->shs == nullptr) == (a->colors_precomp == nullptr)) { set_error("provide exactly one of shs / colors_precomp"); return ENVGS_ERR_INVALID; }
    if (a->shs && (a->sh_degree < 0 || a->sh_degree > 3 || (a->sh_degree + 1) * (a->sh_degree + 1) > a->sh_coeffs)) {
        set_error("sh_degree %d does not fit %d coefficients", a->sh_degree, a->sh_coeffs);
        return ENVGS_ERR_INVALID;
    }
    return ENVGS_OK;
}

int envgs_trace_forward(envgs_tracer* h, const envgs_trace_args* args, const envgs_trace_out* out, void* stream) {
    int rc = check_trace_args(h, args);
    if (rc != ENVGS_OK) return rc;
    if (!out || !out->out_rgb || !out->out_dpt || !out->out_acc || !out->out_norm || !out->out_dist || !out->out_aux || !out->mid_val) {
        set_error("tracer output pointer is NULL");
        return ENVGS_ERR_INVALID;
    }
    if (args->training && !out->a_weights) { set_error("a_weights is NULL although training is set"); return ENVGS_ERR_INVALID; }
    if (trace_fast_eligible(*args, out->hit_cache))
        return trace_forward_fast(h->bvh, h->counter, h->stats(), *args, *out, out->hit_cache, (cudaStream_t)stream, h->side);
    return trace_forward_impl(h->bvh, h->counter, h->stats(), *args, *out, (cudaStream_t)stream);
}

int envgs_trace_backward(envgs_tracer* h, const envgs_trace_args* args, const envgs_trace_grads* g, void* stream) {
    int rc = check_trace_args(h, args);
    if (rc != ENVGS_OK) return rc;
    if (!g || !g->out_rgb || !g->out_dpt || !g->out_acc || !g->out_norm || !g->out_aux || !g->mid_val || !g->dL_drgb || !g->dL_ddpt ||
        !g->dL_dacc || !g->dL_dnorm || !g->dL_daux || !g->dL_dray_o || !g->dL_dray_d || !g->dL_dmeans3D || !g->dL_dcolors ||
        !g->dL_dopacities || !g->dL_dscales || !g->dL_drotations) {
        set_error("tracer gradient pointer is NULL");
        return ENVGS_ERR_INVALID;
    }
    if (args->shs && !g->dL_dshs) { set_error("dL_dshs is NULL although shs were given"); return ENVGS_ERR_INVALID; }
    if (args->others_precomp && !g->dL_dothers) { set_error("dL_dothers is NULL although others_precomp was given"); return ENVGS_ERR_INVALID; }
    if (trace_fast_eligible(*args, g->hit_cache))
        return trace_backward_fast(h->bvh, h->counter, h->stats(), *args, *g, g->hit_cache, (cudaStream_t)stream, h->side);
    return trace_backward_impl(h->bvh, h->counter, h->stats(), *args, *g, (cudaStream_t)stream);
}

int envgs_adam_step(uint64_t n, float* params, float* grads, float* exp_avg, float* exp_avg_sq, int32_t step, float lr, float beta1,
                    float beta2, float eps, float grad_scale, int32_t zero_grad, void* stream) {
    if (n > 0 && (!params || !grads || !exp_avg || !exp_avg_sq)) { set_error("adam_step pointer is NULL"); return ENVGS_ERR_INVALID; }
    if (step < 1 || !(beta1 >= 0.0f && beta1 < 1.0f) || !(beta2 >= 0.0f && beta2 < 1.0f)) { set_error("bad adam_step arguments step=%d beta1=%g beta2=%g", step, beta1, beta2); return ENVGS_ERR_INVALID; }
    if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) {
        set_error("adam_step buffers must be 16-byte aligned");
        return ENVGS_ERR_INVALID;
    }
    return adam_step_impl((size_t)n, params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, grad_scale, zero_grad, (cudaStream_t)stream);
}

}  // extern "C"
