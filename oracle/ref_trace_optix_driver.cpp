// ORACLE tooling (checker / timing baseline only, never linked into the product): a host driver that runs the reference
// tracer's UNMODIFIED device programs — forward.ptx / backward.ptx generated from optix_tracer/{forward,backward}.cu the way
// the reference's CMake does — under the REAL OptiX runtime of the driver (libnvoptix.so.1).  It performs the steps of the
// reference host code, with the same options:
//     context + module + program groups + pipeline + SBT    optix_tracer/optix_wrapper.cpp:33-215
//     triangle GAS build (flags, build options)              trace_surfels.cpp:47-121
//     Params block + optixLaunch(H, W, 1) + stream sync      trace_surfels.cpp:233-311, :382-466
// (The reference's own pybind extension builds here too — oracle/build_ref.py:build_tracer_optix — but crashes inside
// OptiXStateWrapper() on this image; this C-ABI driver has no torch / pybind / iostream dependency.)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>

#include <optix.h>
#include <optix_stubs.h>
#include <optix_stack_size.h>
#include <optix_function_table_definition.h>

#include "params.h"            // the reference's Params block (optix_tracer/params.h), unmodified
#include "ref_trace_args.h"

namespace {

char g_err[1024] = "";
#define OX(call) do { OptixResult r_ = (call); if (r_ != OPTIX_SUCCESS) { snprintf(g_err, sizeof(g_err), "%s -> %d (%s:%d)", #call, (int)r_, __FILE__, __LINE__); return -1; } } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(g_err, sizeof(g_err), "%s -> %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return -2; } } while (0)

template <typename T> struct alignas(OPTIX_SBT_RECORD_ALIGNMENT) SbtRecord { char header[OPTIX_SBT_RECORD_HEADER_SIZE]; T data; };
struct Empty {};

struct Pipe { OptixModule module = nullptr; OptixPipeline pipeline = nullptr; OptixShaderBindingTable sbt = {}; };
struct State {
    OptixDeviceContext ctx = nullptr;
    Pipe fwd, bwd;
    OptixTraversableHandle gas = 0;
    void* d_gas = nullptr; size_t gas_bytes = 0;
    void* d_tmp = nullptr; size_t tmp_bytes = 0;
    void* d_params = nullptr;
} S;

void logcb(unsigned int level, const char* tag, const char* msg, void*) { if (level <= 2) fprintf(stderr, "[optix %u %s] %s\n", level, tag ? tag : "", msg ? msg : ""); }

int make_pipe(const char* ptx_path, Pipe& P) {
    std::ifstream f(ptx_path);
    if (!f) { snprintf(g_err, sizeof(g_err), "cannot open %s", ptx_path); return -3; }
    std::stringstream ss; ss << f.rdbuf();
    const std::string ptx = ss.str();
    OptixModuleCompileOptions mo = {};
    mo.maxRegisterCount = OPTIX_COMPILE_DEFAULT_MAX_REGISTER_COUNT;
    mo.optLevel = OPTIX_COMPILE_OPTIMIZATION_DEFAULT;
    mo.debugLevel = OPTIX_COMPILE_DEBUG_LEVEL_DEFAULT;
    OptixPipelineCompileOptions po = {};
    po.usesMotionBlur = false;
    po.traversableGraphFlags = OPTIX_TRAVERSABLE_GRAPH_FLAG_ALLOW_SINGLE_GAS;
    po.numPayloadValues = 2;
    po.numAttributeValues = 2;
    po.exceptionFlags = OPTIX_EXCEPTION_FLAG_NONE;
    po.pipelineLaunchParamsVariableName = "params";
    po.usesPrimitiveTypeFlags = OPTIX_PRIMITIVE_TYPE_FLAGS_TRIANGLE;
    OX(optixModuleCreate(S.ctx, &mo, &po, ptx.c_str(), ptx.size(), nullptr, nullptr, &P.module));
    OptixProgramGroupOptions pgo = {};
    OptixProgramGroup rg = nullptr, ms = nullptr, hg = nullptr;
    OptixProgramGroupDesc d = {};
    d.kind = OPTIX_PROGRAM_GROUP_KIND_RAYGEN; d.raygen.module = P.module; d.raygen.entryFunctionName = "__raygen__ot";
    OX(optixProgramGroupCreate(S.ctx, &d, 1, &pgo, nullptr, nullptr, &rg));
    d = {}; d.kind = OPTIX_PROGRAM_GROUP_KIND_MISS;
    OX(optixProgramGroupCreate(S.ctx, &d, 1, &pgo, nullptr, nullptr, &ms));
    d = {}; d.kind = OPTIX_PROGRAM_GROUP_KIND_HITGROUP; d.hitgroup.moduleAH = P.module; d.hitgroup.entryFunctionNameAH = "__anyhit__ot";
    OX(optixProgramGroupCreate(S.ctx, &d, 1, &pgo, nullptr, nullptr, &hg));
    const uint32_t max_trace_depth = 31;
    OptixProgramGroup groups[] = {rg, ms, hg};
    OptixPipelineLinkOptions lo = {};
    lo.maxTraceDepth = max_trace_depth;
    OX(optixPipelineCreate(S.ctx, &po, &lo, groups, 3, nullptr, nullptr, &P.pipeline));
    OptixStackSizes ssz = {};
    for (auto& g : groups) OX(optixUtilAccumulateStackSizes(g, &ssz, P.pipeline));
    uint32_t a, b, c;
    OX(optixUtilComputeStackSizes(&ssz, max_trace_depth, 0, 0, &a, &b, &c));
    OX(optixPipelineSetStackSize(P.pipeline, a, b, c, 1));
    SbtRecord<Empty> rec;
    void *d_rg, *d_ms, *d_hg;
    CU(cudaMalloc(&d_rg, sizeof(rec))); OX(optixSbtRecordPackHeader(rg, &rec)); CU(cudaMemcpy(d_rg, &rec, sizeof(rec), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&d_ms, sizeof(rec))); OX(optixSbtRecordPackHeader(ms, &rec)); CU(cudaMemcpy(d_ms, &rec, sizeof(rec), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&d_hg, sizeof(rec))); OX(optixSbtRecordPackHeader(hg, &rec)); CU(cudaMemcpy(d_hg, &rec, sizeof(rec), cudaMemcpyHostToDevice));
    P.sbt.raygenRecord = (CUdeviceptr)d_rg;
    P.sbt.missRecordBase = (CUdeviceptr)d_ms; P.sbt.missRecordStrideInBytes = sizeof(rec); P.sbt.missRecordCount = 1;
    P.sbt.hitgroupRecordBase = (CUdeviceptr)d_hg; P.sbt.hitgroupRecordStrideInBytes = sizeof(rec); P.sbt.hitgroupRecordCount = 1;
    return 0;
}

}  // namespace

extern "C" const char* ref_optix_last_error() { return g_err; }

extern "C" int ref_optix_create(const char* fwd_ptx, const char* bwd_ptx, int validation) {
    if (S.ctx) return 0;
    CU(cudaFree(0));
    OX(optixInit());
    OptixDeviceContextOptions opt = {};
    opt.logCallbackFunction = &logcb;
    opt.logCallbackLevel = 3;
    opt.validationMode = validation ? OPTIX_DEVICE_CONTEXT_VALIDATION_MODE_ALL : OPTIX_DEVICE_CONTEXT_VALIDATION_MODE_OFF;
    OX(optixDeviceContextCreate(0, &opt, &S.ctx));
    if (int r = make_pipe(fwd_ptx, S.fwd)) return r;
    if (int r = make_pipe(bwd_ptx, S.bwd)) return r;
    CU(cudaMalloc(&S.d_params, sizeof(Params)));
    return 0;
}

// trace_surfels.cpp:47-121 with rebuild > 0 (EnvGS rebuilds every iteration); buffers are kept and grown instead of
// cudaFree/cudaMalloc per call, which only removes allocator time from the timed build
extern "C" int ref_optix_build(const float* vertices, int num_vertices, const int* triangles, int num_triangles, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    OptixAccelBuildOptions ao = {};
    ao.buildFlags = OPTIX_BUILD_FLAG_ALLOW_COMPACTION | OPTIX_BUILD_FLAG_ALLOW_UPDATE | OPTIX_BUILD_FLAG_ALLOW_RANDOM_VERTEX_ACCESS;
    ao.operation = OPTIX_BUILD_OPERATION_BUILD;
    CUdeviceptr dv = (CUdeviceptr)vertices;
    const uint32_t flags[1] = {OPTIX_GEOMETRY_FLAG_NONE};
    OptixBuildInput in = {};
    in.type = OPTIX_BUILD_INPUT_TYPE_TRIANGLES;
    in.triangleArray.vertexFormat = OPTIX_VERTEX_FORMAT_FLOAT3;
    in.triangleArray.numVertices = (uint32_t)num_vertices;
    in.triangleArray.vertexBuffers = &dv;
    in.triangleArray.flags = flags;
    in.triangleArray.numSbtRecords = 1;
    in.triangleArray.indexFormat = OPTIX_INDICES_FORMAT_UNSIGNED_INT3;
    in.triangleArray.numIndexTriplets = (uint32_t)num_triangles;
    in.triangleArray.indexBuffer = (CUdeviceptr)triangles;
    OptixAccelBufferSizes sz;
    OX(optixAccelComputeMemoryUsage(S.ctx, &ao, &in, 1, &sz));
    if (sz.tempSizeInBytes > S.tmp_bytes) { if (S.d_tmp) cudaFree(S.d_tmp); CU(cudaMalloc(&S.d_tmp, sz.tempSizeInBytes)); S.tmp_bytes = sz.tempSizeInBytes; }
    if (sz.outputSizeInBytes > S.gas_bytes) { if (S.d_gas) cudaFree(S.d_gas); CU(cudaMalloc(&S.d_gas, sz.outputSizeInBytes)); S.gas_bytes = sz.outputSizeInBytes; }
    OX(optixAccelBuild(S.ctx, stream, &ao, &in, 1, (CUdeviceptr)S.d_tmp, sz.tempSizeInBytes, (CUdeviceptr)S.d_gas, sz.outputSizeInBytes, &S.gas,
                       nullptr, 0));
    return 0;
}

extern "C" int ref_optix_launch(const RefTraceArgs* a, int backward, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    Params p;
    memset(&p, 0, sizeof(p));
    p.handle = S.gas;
    p.training = a->training != 0;
    p.max_trace_depth = a->max_trace_depth;
    p.start_from_first = a->start_from_first != 0;
    p.specular_threshold = a->specular_threshold;
    p.P = a->P; p.H = a->H; p.W = a->W; p.D = a->D; p.M = a->M;
    p.ray_o = (float3*)a->ray_o; p.ray_d = (float3*)a->ray_d; p.vertices = (float3*)a->vertices;
    p.background = a->background; p.means3D = (float3*)a->means3D; p.shs = (float3*)a->shs;
    p.colors_precomp = a->colors_precomp; p.others_precomp = a->others_precomp; p.opacities = a->opacities;
    p.scales = (float2*)a->scales; p.scale_modifier = a->scale_modifier; p.rotations = (float4*)a->rotations;
    p.transMat_precomp = nullptr; p.viewmatrix = a->viewmatrix; p.projmatrix = a->projmatrix; p.campos = (float3*)a->campos;
    p.out_rgb = a->out_rgb; p.out_dpt = a->out_dpt; p.out_acc = a->out_acc; p.out_dist = (float3*)a->out_dist;
    p.out_norm = (float3*)a->out_norm; p.out_aux = a->out_aux; p.mid_val = a->mid_val; p.a_weights = a->a_weights;
    p.dL_drgb = a->dL_drgb; p.dL_ddpt = a->dL_ddpt; p.dL_dacc = a->dL_dacc; p.dL_ddist = a->dL_ddist;
    p.dL_dnorm = (float3*)a->dL_dnorm; p.dL_daux = a->dL_daux;
    p.dL_dray_o = (float3*)a->dL_dray_o; p.dL_dray_d = (float3*)a->dL_dray_d; p.dL_dmeans3D = (float3*)a->dL_dmeans3D;
    p.dL_dshs = (float3*)a->dL_dshs; p.dL_dcolors = a->dL_dcolors; p.dL_dothers = a->dL_dothers;
    p.dL_dopacities = a->dL_dopacities; p.dL_dscales = (float2*)a->dL_dscales; p.dL_drotations = (float4*)a->dL_drotations;
    p.dL_dtransMat_precomp = a->dL_dtransMat;
    CU(cudaMemcpyAsync(S.d_params, &p, sizeof(Params), cudaMemcpyHostToDevice, stream));
    Pipe& P = backward ? S.bwd : S.fwd;
    OX(optixLaunch(P.pipeline, stream, (CUdeviceptr)S.d_params, sizeof(Params), &P.sbt, (unsigned)a->H, (unsigned)a->W, 1));
    CU(cudaStreamSynchronize(stream));    // the reference synchronises after every launch (trace_surfels.cpp:308, :463)
    return 0;
}
