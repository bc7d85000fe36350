// Shared helpers for the sm_100a kernels behind include/fldr_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fldr_b200.h"

namespace fldr {

// [N,C,H,W] view with element strides (the reference specialises strides into the kernel source by
// regex per call, softSplat.py:178-210; here they are runtime arguments of one AOT binary).
struct View4 {
    const float* p;
    long long sn, sc, sh, sw;
};

inline View4 make_view(const float* p, const int64_t* s) {
    View4 v;
    v.p = p;
    v.sn = s ? s[0] : 0; v.sc = s ? s[1] : 0; v.sh = s ? s[2] : 0; v.sw = s ? s[3] : 0;
    return v;
}

void set_last_cuda_error(cudaError_t e);

inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
    return FLDR_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count();

// tuning / diagnostic options (fldr_set_option); defaults may come from FLDR_<NAME> environment variables
enum Option { kOptSplatStream = 0, kOptSplatFusedMax = 1, kOptCorrTh = 2, kOptSplatPfRows = 3, kOptCorrBwdRows = 4, kOptSplatSnake = 5, kOptCount = 6 };
int get_option(int opt);

// red.global.add.v4.f32 (sm_90+): one 16-byte reduction request instead of four scalar REDs.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
}  // namespace fldr
