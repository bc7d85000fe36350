// Declarations shared by the splat kernels (splat.cu: whole-frame and small-frame paths, backward).
#pragma once
#include "common.cuh"

namespace fldr {

struct SplatGeom {
    int N, C, H, W;
    int CA;        // accumulated channels: C (+1 when the mode carries a normaliser)
    int CP;        // CA rounded up to a multiple of 4
    int mode;      // fldr_splat_mode
    int has_metric;
};

__host__ __device__ inline bool mode_has_norm(int mode) {
    return mode == FLDR_SPLAT_AVERAGE || mode == FLDR_SPLAT_LINEAR || mode == FLDR_SPLAT_SOFTMAX;
}

// softSplat.py:343-349 for one accumulated value: hole fix-up, divide, post-scale.  Every path (normalise pass, small-frame
// kernel) goes through these two helpers, so a call gives the same bits apart from the summation order
// whichever path serves it.  The divide is one reciprocal per pixel (rcp.approx.f32: 1 ulp, subnormals handled) and a
// multiply per channel - within 2 ulp of the reference's S / norm, far inside the summation-order noise; an IEEE division
// per channel made the normalise pass instruction-bound.
__device__ __forceinline__ float norm_recip(float nrm) {
    float r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(nrm));
    return (nrm == 0.f) ? 1.f : r;
}
__device__ __forceinline__ float post_scale(float sv, float d, bool raw, bool has_norm) {
    if (raw) return sv;
    if (!has_norm) return (sv - 0.5f) * 2.f;
    return (sv * d - 0.5f) * 2.f;
}

}  // namespace fldr
