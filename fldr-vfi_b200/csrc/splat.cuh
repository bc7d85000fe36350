// Declarations shared by the splat translation units (splat.cu: whole-frame and small-frame paths, backward;
// splat_ring.cu: single-launch streaming path with the L2-resident ring accumulator).
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
// kernel, ring kernel) goes through these two helpers, so a call gives the same bits apart from the summation order
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

// ---- ring path (splat_ring.cu) ----
struct RingGeom {
    int NS;          // strips per sample = ceil(H / 8)
    int NT;          // N * NS absolute strips
    int T;           // column tiles = ceil(W / 128)
    int Q;           // channel quads
    int RS;          // ring strips
    int Ds;          // vertical reach in strips
    int pitch;       // ring row pitch in float4 cells = W + 2 (one guard cell either side: no x test on the reductions)
    int ring_rows;   // RS * 8
    int mask;        // ring_rows - 1 (bounded: a power of two) or all ones (the ring holds the whole batch: no wrap)
    int nZs, nZ;     // ring slots zeroed up front = min(RS, NT), and the zero-fill items that takes (nZs * T)
    int TQ;          // T * Q scatter items per strip (T normalise items per strip)
    int total;       // all work items
    int bounded;     // ring smaller than the batch: reach is bounded, the whole-frame fallback must be armed
    int spin_limit;  // watchdog of the dependency polls
};

struct RingPlan {
    RingGeom rg;
    bool ok;
    size_t ring_bytes, ctrl_bytes;
};

constexpr int kRingCtrlClaim = 0;      // ctrl word indices (unsigned): {N tickets claimed, S tickets claimed, scatter frontier, clean frontier}
constexpr int kRingCtrlFlag = 32;      // bit 0: a source left the ring's reach, bit 1: watchdog
constexpr int kRingCtrlCounters = 64;  // sdone[NT], then ev[nZs + NT]

// Can the ring kernel serve this call?  (layout / alignment requirements of its bulk copies and vector stores)
bool ring_eligible(const SplatGeom& g, const View4& in, const View4& flow, const View4& metric, const float* out, const float* norm);
void plan_ring(const SplatGeom& g, RingPlan& p);
int launch_ring(const RingPlan& p, const SplatGeom& g, const View4& in, const View4& flow, const View4& metric, void* ring,
                unsigned* ctrl, float* out, float* norm, cudaStream_t s);

}  // namespace fldr
