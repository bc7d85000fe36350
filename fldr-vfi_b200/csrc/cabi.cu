// Library-level entry points of include/fldr_b200.h (status strings, error state, device queries).
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
