// Library-level entry points of include/fldr_b200.h (status strings, error state, device queries).
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace fldr {

static thread_local int g_last_cuda_error = 0;

void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = (int)e; }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace fldr

namespace fldr {
static const char* const kOptionNames[kOptCount] = {"splat_tma", "splat_fused_max", "corr_th", "splat_pf_rows", "corr_bwd_rows", "splat_snake"};
static const char* const kOptionEnv[kOptCount] = {"FLDR_SPLAT_TMA", "FLDR_SPLAT_FUSED_MAX", "FLDR_CORR_TH", "FLDR_SPLAT_PF_ROWS", "FLDR_CORR_BWD_ROWS", "FLDR_SPLAT_SNAKE"};
static const int kOptionDefault[kOptCount] = {1, 40000, 0, 0, 1, 1};
static int g_options[kOptCount];
static std::once_flag g_options_once;
static void init_options() {
    std::call_once(g_options_once, [] {
        for (int i = 0; i < kOptCount; ++i) {
            const char* e = getenv(kOptionEnv[i]);
            g_options[i] = e ? atoi(e) : kOptionDefault[i];
        }
    });
}
int get_option(int opt) {
    init_options();
    return (opt >= 0 && opt < kOptCount) ? g_options[opt] : 0;
}
}  // namespace fldr

extern "C" int fldr_set_option(const char* name, int value) {
    if (!name) return FLDR_ERR_INVALID_ARGUMENT;
    fldr::init_options();
    for (int i = 0; i < fldr::kOptCount; ++i)
        if (strcmp(name, fldr::kOptionNames[i]) == 0) { fldr::g_options[i] = value; return FLDR_OK; }
    return FLDR_ERR_INVALID_ARGUMENT;
}

extern "C" int fldr_get_option(const char* name) {
    if (!name) return 0;
    for (int i = 0; i < fldr::kOptCount; ++i)
        if (strcmp(name, fldr::kOptionNames[i]) == 0) return fldr::get_option(i);
    return 0;
}

extern "C" int fldr_abi_version(void) { return FLDR_B200_ABI_VERSION; }

extern "C" int fldr_last_cuda_error(void) { return fldr::g_last_cuda_error; }

extern "C" const char* fldr_status_string(int status) {
    switch (status) {
        case FLDR_OK: return "FLDR_OK";
        case FLDR_ERR_INVALID_ARGUMENT: return "FLDR_ERR_INVALID_ARGUMENT: null pointer, non-positive size, unknown mode or bad alignment";
        case FLDR_ERR_WORKSPACE_TOO_SMALL: return "FLDR_ERR_WORKSPACE_TOO_SMALL: workspace missing or smaller than *_workspace_bytes()";
        case FLDR_ERR_CUDA: return "FLDR_ERR_CUDA: a CUDA runtime call failed (see fldr_last_cuda_error)";
        case FLDR_ERR_UNSUPPORTED: return "FLDR_ERR_UNSUPPORTED: argument combination not supported";
        case FLDR_ERR_NO_DEVICE: return "FLDR_ERR_NO_DEVICE: no CUDA device";
        default: return "FLDR_ERR_UNKNOWN";
    }
}

// ---- TMA tensor-map encoding through the runtime's driver entry point (no -lcuda at link time) ----
#include "tma.cuh"

namespace fldr {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

bool encode_tensor_map_4d(CUtensorMap* map, const float* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                          const uint32_t box[4]) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
    for (int i = 0; i < 3; ++i)
        if ((strides_bytes[i] & 15) != 0 || strides_bytes[i] >= (1ull << 40)) return false;
    for (int i = 0; i < 4; ++i)
        if (box[i] == 0 || box[i] > 256) return false;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT) {
        // a thread that has only used the runtime API lazily (torch's autograd workers) may have no driver context bound yet:
        // bind the primary context through the runtime and try once more, instead of silently losing the TMA path
        cudaFree(nullptr);
        r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides_bytes, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS;
}

}  // namespace fldr
