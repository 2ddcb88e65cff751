// ORACLE tooling (probe only): smallest possible OptiX host program — does the driver's libnvoptix accept ABI 7.7 and the
// reference's PTX on this GPU?  Prints one line per step to stderr.  Build: see oracle/build_ref.py (build_optix_probe).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <optix.h>
#include <optix_stubs.h>
#include <optix_function_table_definition.h>

static void logcb(unsigned int level, const char* tag, const char* msg, void*) { fprintf(stderr, "[optix %u %s] %s\n", level, tag ? tag : "?", msg ? msg : "?"); }
#define STEP(x) do { fprintf(stderr, "step: %s\n", #x); fflush(stderr); OptixResult r_ = (x); fprintf(stderr, "  -> %d\n", (int)r_); fflush(stderr); if (r_ != OPTIX_SUCCESS) return 10; } while (0)

int main(int argc, char** argv) {
    int validation = argc > 2 ? atoi(argv[2]) : 0;
    cudaFree(0);
    fprintf(stderr, "OPTIX_VERSION %d ABI %d\n", OPTIX_VERSION, OPTIX_ABI_VERSION);
    STEP(optixInit());
    OptixDeviceContextOptions opt = {};
    opt.logCallbackFunction = &logcb;
    opt.logCallbackLevel = 4;
    opt.validationMode = validation ? OPTIX_DEVICE_CONTEXT_VALIDATION_MODE_ALL : OPTIX_DEVICE_CONTEXT_VALIDATION_MODE_OFF;
    OptixDeviceContext ctx = nullptr;
    STEP(optixDeviceContextCreate(0, &opt, &ctx));
    if (argc > 1) {
        std::ifstream f(argv[1]);
        std::stringstream ss; ss << f.rdbuf();
        std::string ptx = ss.str();
        fprintf(stderr, "ptx bytes %zu\n", ptx.size());
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
        char log[4096]; size_t ls = sizeof(log);
        OptixModule mod = nullptr;
        STEP(optixModuleCreate(ctx, &mo, &po, ptx.c_str(), ptx.size(), log, &ls, &mod));
        fprintf(stderr, "module log: %s\n", log);
    }
    fprintf(stderr, "probe ok\n");
    return 0;
}
