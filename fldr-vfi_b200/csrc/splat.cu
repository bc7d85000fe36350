// Softmax / average / linear / summation forward splat and its backward, sm_100a.
//
// Replaces softSplat.py:12-158 (three CuPy string kernels) AND the ~12 torch elementwise kernels
// FunctionSoftsplat wraps around them (softSplat.py:320-352): pre-scale, exp, cat, zero-init,
// normaliser fix-up, divide, post-scale are all folded into the two passes below.
//
// Data layout in HBM
//   inputs      NCHW fp32 with arbitrary element strides (callers pass views: fLDRnet.py:386,449)
//   accumulator pixel-interleaved [N, H, W, CP] fp32, CP = round_up(C + has_norm, 4): one source pixel's
//               whole payload for one corner is CP/4 red.global.add.v4.f32 requests (16 B each) instead
//               of CP scalar REDs to CP planes (the reference issues 4 scalar REDs per element, 39-50).
//   outputs     NCHW contiguous (what the reference allocates, softSplat.py:234).
#include <math.h>
#include <stdlib.h>

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

// Target coordinate, NW corner and the four bilinear weights exactly as softSplat.py:23-38 forms them
// (integer corner converted back to float, then subtracted).  Returns false when no corner can be in frame
// or the coordinate is not finite (the reference device-asserts there; we skip the pixel).
struct Corners {
    float X, Y;
    int x0, y0;
    float w[4];       // NW, NE, SW, SE
    bool valid[4];
};

__device__ __forceinline__ bool make_corners(int x, int y, float u, float v, int W, int H, Corners& k) {
    k.X = (float)x + u;
    k.Y = (float)y + v;
    if (!(isfinite(k.X) && isfinite(k.Y))) return false;
    const float fx0 = floorf(k.X), fy0 = floorf(k.Y);
    if (fx0 < -1.f || fx0 >= (float)W || fy0 < -1.f || fy0 >= (float)H) return false;
    k.x0 = (int)fx0;
    k.y0 = (int)fy0;
    const float x1f = (float)(k.x0 + 1), y1f = (float)(k.y0 + 1);
    k.w[0] = (x1f - k.X) * (y1f - k.Y);
    k.w[1] = (k.X - fx0) * (y1f - k.Y);
    k.w[2] = (x1f - k.X) * (k.Y - fy0);
    k.w[3] = (k.X - fx0) * (k.Y - fy0);
    const bool xl = k.x0 >= 0, xr = k.x0 + 1 < W, yt = k.y0 >= 0, yb = k.y0 + 1 < H;
    k.valid[0] = xl && yt;
    k.valid[1] = xr && yt;
    k.valid[2] = xl && yb;
    k.valid[3] = xr && yb;
    return true;
}

__device__ __forceinline__ float source_weight(const SplatGeom& g, const View4& metric, int n, int y, int x) {
    if (!g.has_metric) return 1.f;
    const float z = __ldg(metric.p + n * metric.sn + y * metric.sh + x * metric.sw);
    if (g.mode == FLDR_SPLAT_SOFTMAX) return expf(z);   // accurate expf: parity bar is 1e-5 relative
    if (g.mode == FLDR_SPLAT_LINEAR) return z;
    return 1.f;
}

// ------------------------------------------------------------------------------------------------
// Pass 1: scatter with merged reductions.
//
// Measured on B200 (profiles/r1_microbench_design.txt): the L2 reduction unit sustains ~333 G 16-byte REDs/s for
// coalesced targets (113 us for the 4 REDs/pixel of one 4K image), ~81 G/s for random targets, and shared-memory
// atomics are slower still (226 us) - so the lever is issuing FEWER reductions, not privatising them.
// A thread owns one column and walks R source rows; lanes of a warp own adjacent columns.  Wherever the flow is
// locally constant
//   * the NE/SE contribution of column x lands on the pixel that column x+1 hits with NW/SW -> handed to the right
//     lane by warp shuffle, and
//   * the SW/SE contribution of row y lands on the pixel that row y+1 hits with NW/NE       -> carried in registers
// so a source pixel costs ~1 red.global.add.v4.f32 instead of 4 (the reference issues 4 scalar REDs per ELEMENT,
// softSplat.py:39-50).  A target mismatch simply flushes the carried value as its own RED: arbitrary flow stays
// correct, it only merges less.  Inactive pixels (out of frame, non-finite) carry sentinel coordinates that never match
// and never pass the frame test, so no per-corner flags are kept.
// Accumulator layout: [N][Q][H][W][4] fp32, Q = ceil(CA/4): every channel quad is a pixel-interleaved float4 image,
// so lanes (adjacent x) reduce into adjacent 16-byte slots.
// ------------------------------------------------------------------------------------------------
constexpr int kSentinel = -(1 << 28);

template <int PX> __device__ __forceinline__ void vstore(float* p, const float* t);
template <> __device__ __forceinline__ void vstore<1>(float* p, const float* t) { __stcs(p, t[0]); }
template <> __device__ __forceinline__ void vstore<4>(float* p, const float* t) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(t[0], t[1], t[2], t[3]));
}

// WKIND: 0 = weight 1 (no metric / average / summation / raw), 1 = exp(z) (softmax), 2 = z (linear)
template <int R, int WKIND, bool PRE>
__global__ void __launch_bounds__(128) splat_scatter_merged_kernel(View4 in, View4 flow, View4 metric,
                                                                   float* __restrict__ acc, SplatGeom g, int Q) {
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int yb = blockIdx.y * R;
    const int q = blockIdx.z % Q, n = blockIdx.z / Q;
    const int W = g.W, H = g.H;
    const bool inb = x < W;
    float4* accq = reinterpret_cast<float4*>(acc) + (long long)(n * Q + q) * H * W;
    const int rows = min(R, H - yb);

    // per-tensor element offsets of (n, *, yb, x); advanced by the row stride each iteration
    const float* fu = flow.p + n * flow.sn + (long long)yb * flow.sh + (long long)x * flow.sw;
    const float* fv = fu + flow.sc;
    const float* zp = WKIND ? metric.p + n * metric.sn + (long long)yb * metric.sh + (long long)x * metric.sw : nullptr;
    const float* ip = in.p + n * in.sn + (long long)(q * 4) * in.sc + (long long)yb * in.sh + (long long)x * in.sw;
    const int nch = min(4, g.C - q * 4);                 // real channels in this quad (<= 0: only the weight slot)
    const int wslot = (g.CA > g.C) ? g.C - q * 4 : -1;    // slot of this quad that carries the weight, if in [0,4)
    const float xf = (float)x;

    float pw[4] = {0.f, 0.f, 0.f, 0.f}, pe[4] = {0.f, 0.f, 0.f, 0.f};
    int px = kSentinel, pex = kSentinel, py = kSentinel;

    // software pipeline: the loads of row r+1 are issued before row r is processed
    float nu = 0.f, nv = 0.f, nz = 0.f, nx[4] = {0.f, 0.f, 0.f, 0.f};
    auto load = [&](int r) {
        if (inb) {
            nu = __ldg(fu + (long long)r * flow.sh);
            nv = __ldg(fv + (long long)r * flow.sh);
            if (WKIND) nz = __ldg(zp + (long long)r * metric.sh);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < nch) nx[j] = __ldg(ip + (long long)j * in.sc + (long long)r * in.sh);
        }
    };
    if (rows > 0) load(0);

#pragma unroll 2
    for (int r = 0; r < rows; ++r) {
        const float u = nu, v = nv, z = nz;
        const float xv[4] = {nx[0], nx[1], nx[2], nx[3]};
        if (r + 1 < rows) load(r + 1);

        // softSplat.py:23-38
        const float X = xf + u, Y = (float)(yb + r) + v;
        const float fx0 = floorf(X), fy0 = floorf(Y);
        const bool ok = inb && fx0 >= -1.f && fx0 < (float)W && fy0 >= -1.f && fy0 < (float)H;   // false for NaN / inf
        const int x0 = ok ? (int)fx0 : kSentinel;
        const int y0 = ok ? (int)fy0 : kSentinel;
        const float ax = (fx0 + 1.f) - X, bx = X - fx0, ay = (fy0 + 1.f) - Y, by = Y - fy0;
        const float wNW = ax * ay, wNE = bx * ay, wSW = ax * by, wSE = bx * by;
        float m = 1.f;
        if (WKIND == 1) m = expf(z);          // accurate expf: the parity bar is 1e-5 relative
        if (WKIND == 2) m = z;
        float tW[4], tE[4], bW[4], bE[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = xv[j];
            if (PRE) a = (a + 1.f) * 0.5f;                      // softSplat.py:334
            a = (j < nch) ? a * m : (j == wslot ? m : 0.f);     // softSplat.py:328 / 338, weight slot, padding
            if (!ok) a = 0.f;
            tW[j] = a * wNW; tE[j] = a * wNE; bW[j] = a * wSW; bE[j] = a * wSE;
        }

        // vertical: the previous row's bottom contributions join this row's top ones when they hit the same pixels
        if (x0 == px && y0 == py) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { tW[j] += pw[j]; tE[j] += pe[j]; }
        } else {
            if ((unsigned)py < (unsigned)H) {
                if ((unsigned)px < (unsigned)W) { float4* d = accq + (long long)py * W + px; red_add_v4((float*)d, pw[0], pw[1], pw[2], pw[3]); }
                if ((unsigned)pex < (unsigned)W) { float4* d = accq + (long long)py * W + pex; red_add_v4((float*)d, pe[0], pe[1], pe[2], pe[3]); }
            }
        }

        // horizontal: this column's E slots are the next lane's W slots when its NW target is (x0 + 1, y0)
        int ex = ok ? x0 + 1 : kSentinel;
        {
            const int rx = __shfl_up_sync(0xffffffffu, ex, 1);
            const int ry = __shfl_up_sync(0xffffffffu, y0, 1);
            float rt[4], rb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                rt[j] = __shfl_up_sync(0xffffffffu, tE[j], 1);
                rb[j] = __shfl_up_sync(0xffffffffu, bE[j], 1);
            }
            const bool take = (lane > 0) && ok && rx == x0 && ry == y0;
            if (take) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { tW[j] += rt[j]; bW[j] += rb[j]; }
            }
            const bool taken = __shfl_down_sync(0xffffffffu, (int)take, 1) != 0;
            if (lane < 31 && taken) {
                ex = kSentinel;
#pragma unroll
                for (int j = 0; j < 4; ++j) { tE[j] = 0.f; bE[j] = 0.f; }
            }
        }

        // the top row is final for this thread: one reduction in the merged case
        if ((unsigned)y0 < (unsigned)H) {
            if ((unsigned)x0 < (unsigned)W) { float4* d = accq + (long long)y0 * W + x0; red_add_v4((float*)d, tW[0], tW[1], tW[2], tW[3]); }
            if ((unsigned)ex < (unsigned)W) { float4* d = accq + (long long)y0 * W + ex; red_add_v4((float*)d, tE[0], tE[1], tE[2], tE[3]); }
        }
        px = x0; pex = ex; py = ok ? y0 + 1 : kSentinel;
#pragma unroll
        for (int j = 0; j < 4; ++j) { pw[j] = bW[j]; pe[j] = bE[j]; }
    }
    if ((unsigned)py < (unsigned)H) {
        if ((unsigned)px < (unsigned)W) { float4* d = accq + (long long)py * W + px; red_add_v4((float*)d, pw[0], pw[1], pw[2], pw[3]); }
        if ((unsigned)pex < (unsigned)W) { float4* d = accq + (long long)py * W + pex; red_add_v4((float*)d, pe[0], pe[1], pe[2], pe[3]); }
    }
}

// ------------------------------------------------------------------------------------------------
// Pass 2: normalise + post-scale + quad-interleaved -> NCHW.  One thread per (PX consecutive pixels, channel quad):
// PX float4 loads of the accumulator (+ PX of the quad holding the normaliser), 4 channel-plane stores of PX floats.
//   softSplat.py:343-349: norm==0 -> 1, divide, (y - 0.5) * 2 (post-scale in every mode but RAW).
// ------------------------------------------------------------------------------------------------
template <int PX>
__global__ void __launch_bounds__(256) splat_normalise_kernel(const float* __restrict__ acc, float* __restrict__ out,
                                                              float* __restrict__ norm_out, SplatGeom g, int Q) {
    const long long HW = (long long)g.H * g.W;
    const long long pix = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * PX;
    if (pix >= HW) return;      // PX > 1 only when HW % PX == 0
    const int q = blockIdx.y % Q, n = blockIdx.y / Q;
    const bool has_norm = g.CA > g.C;
    const float4* accn = reinterpret_cast<const float4*>(acc) + (long long)n * Q * HW;
    float4 s4[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) s4[k] = __ldcs(accn + q * HW + pix + k);
    float d[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) d[k] = 1.f;
    if (has_norm) {
        const int qn = g.C >> 2, slot = g.C & 3;
        float nrm[PX];
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            const float4 n4 = (qn == q) ? s4[k] : __ldcs(accn + qn * HW + pix + k);
            nrm[k] = slot == 0 ? n4.x : slot == 1 ? n4.y : slot == 2 ? n4.z : n4.w;
            d[k] = (nrm[k] == 0.f) ? 1.f : nrm[k];
        }
        if (norm_out && q == 0) vstore<PX>(norm_out + (long long)n * HW + pix, nrm);
    }
    float* op = out + (long long)n * g.C * HW + pix;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = q * 4 + j;
        if (c < g.C) {
            float yv[PX];
#pragma unroll
            for (int k = 0; k < PX; ++k) {
                const float sv = j == 0 ? s4[k].x : j == 1 ? s4[k].y : j == 2 ? s4[k].z : s4[k].w;
                if (g.mode == FLDR_SPLAT_RAW) yv[k] = sv;
                else if (!has_norm) yv[k] = (sv - 0.5f) * 2.f;
                else yv[k] = (sv / d[k] - 0.5f) * 2.f;
            }
            vstore<PX>(op + (long long)c * HW, yv);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward.  One thread per source pixel; gathers grad_out / forward output / normaliser at its 4 corners,
// forms gS on the fly and emits grad_in, grad_flow, grad_metric in one pass:
//   gS_c = 2 gY_c / norm'            gS_C = -sum_c gS_c * (S_c / norm')   (0 where norm was 0)
//   gA   = sum_corners w * gS        (kernel_Softsplat_updateGradInput, softSplat.py:84-95)
//   gF   = sum_c A_c * sum_corners gS_c * dw   (kernel_Softsplat_updateGradFlow, 130-155)
//   softmax: g_x = gA_c * e^z / 2 ; g_z = e^z (sum_c gA_c x~_c + gA_C)      linear: g_x = gA_c z ; g_z = sum_c gA_c x_c + gA_C
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) splat_bwd_kernel(View4 in, View4 flow, View4 metric, const float* __restrict__ Yf,
                                                        const float* __restrict__ norm, View4 gout,
                                                        float* __restrict__ gin, float* __restrict__ gflow,
                                                        float* __restrict__ gmetric, SplatGeom g) {
    const long long HW = (long long)g.H * g.W;
    const long long total = HW * g.N;
    const bool pre = g.mode == FLDR_SPLAT_SOFTMAX;
    const bool has_norm = g.CA > g.C;
    const float gscale = (g.mode == FLDR_SPLAT_RAW) ? 1.f : 2.f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % g.W);
        const int y = (int)((idx / g.W) % g.H);
        const int n = (int)(idx / HW);
        const long long pix = (long long)y * g.W + x;
        const float* fp = flow.p + n * flow.sn + y * flow.sh + x * flow.sw;
        Corners k;
        const bool live = make_corners(x, y, __ldg(fp), __ldg(fp + flow.sc), g.W, g.H, k);
        float* ginp = gin ? gin + (long long)n * g.C * HW + pix : nullptr;
        if (!live) {
            if (ginp) for (int c = 0; c < g.C; ++c) ginp[(long long)c * HW] = 0.f;
            if (gflow) { gflow[(long long)n * 2 * HW + pix] = 0.f; gflow[(long long)n * 2 * HW + HW + pix] = 0.f; }
            if (gmetric) gmetric[(long long)n * HW + pix] = 0.f;
            continue;
        }
        float m = 1.f, z = 0.f;
        if (g.has_metric) {
            z = __ldg(metric.p + n * metric.sn + y * metric.sh + x * metric.sw);
            m = (g.mode == FLDR_SPLAT_SOFTMAX) ? expf(z) : (g.mode == FLDR_SPLAT_LINEAR ? z : 1.f);
        }
        long long cpix[4];      // corner pixel index inside one plane
        cpix[0] = (long long)k.y0 * g.W + k.x0;
        cpix[1] = cpix[0] + 1;
        cpix[2] = cpix[0] + g.W;
        cpix[3] = cpix[2] + 1;
        float rd[4], gsC[4];
        bool hole[4];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            rd[c4] = gscale; gsC[c4] = 0.f; hole[c4] = false;
            if (has_norm && k.valid[c4]) {
                const float nr = __ldg(norm + (long long)n * HW + cpix[c4]);
                hole[c4] = (nr == 0.f);
                rd[c4] = gscale / (hole[c4] ? 1.f : nr);
            }
        }
        const float wy1 = (float)(k.y0 + 1) - k.Y, wy0 = k.Y - (float)k.y0;
        const float wx1 = (float)(k.x0 + 1) - k.X, wx0 = k.X - (float)k.x0;
        float gfx = 0.f, gfy = 0.f, sum_gx = 0.f;
        const float* ip = in.p + n * in.sn + y * in.sh + x * in.sw;
        const float* gop = gout.p + n * gout.sn;
        const float* yp = Yf ? Yf + (long long)n * g.C * HW : nullptr;
        for (int c = 0; c < g.C; ++c) {
            const float xv = __ldg(ip + c * in.sc);
            const float xt = pre ? (xv + 1.f) * 0.5f : xv;
            const float A = xt * m;
            float gs[4];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                gs[c4] = 0.f;
                if (k.valid[c4]) {
                    const int cy = k.y0 + (c4 >> 1), cx = k.x0 + (c4 & 1);
                    const float go = __ldg(gop + c * gout.sc + cy * gout.sh + cx * gout.sw);
                    gs[c4] = go * rd[c4];
                    if (has_norm && !hole[c4]) {
                        const float q = __ldg(yp + (long long)c * HW + cpix[c4]) * 0.5f + 0.5f;   // S_c / norm'
                        gsC[c4] -= gs[c4] * q;
                    }
                }
            }
            const float gA = gs[0] * k.w[0] + gs[1] * k.w[1] + gs[2] * k.w[2] + gs[3] * k.w[3];
            if (ginp) {
                float gx = gA;
                if (g.mode == FLDR_SPLAT_SOFTMAX) gx = gA * m * 0.5f;
                else if (g.mode == FLDR_SPLAT_LINEAR) gx = gA * m;
                ginp[(long long)c * HW] = gx;
            }
            sum_gx += gA * xt;
            gfx += A * ((gs[1] - gs[0]) * wy1 + (gs[3] - gs[2]) * wy0);
            gfy += A * ((gs[2] - gs[0]) * wx1 + (gs[3] - gs[1]) * wx0);
        }
        float gAC = 0.f;
        if (has_norm) {
            gAC = gsC[0] * k.w[0] + gsC[1] * k.w[1] + gsC[2] * k.w[2] + gsC[3] * k.w[3];
            gfx += m * ((gsC[1] - gsC[0]) * wy1 + (gsC[3] - gsC[2]) * wy0);
            gfy += m * ((gsC[2] - gsC[0]) * wx1 + (gsC[3] - gsC[1]) * wx0);
        }
        if (gflow) {
            gflow[(long long)n * 2 * HW + pix] = gfx;
            gflow[(long long)n * 2 * HW + HW + pix] = gfy;
        }
        if (gmetric) {
            float gz = sum_gx + gAC;
            if (g.mode == FLDR_SPLAT_SOFTMAX) gz *= m;
            gmetric[(long long)n * HW + pix] = gz;
        }
    }
}

static int make_geom(int mode, int N, int C, int H, int W, bool has_metric, SplatGeom& g) {
    if (mode < 0 || mode > FLDR_SPLAT_RAW) return FLDR_ERR_INVALID_ARGUMENT;
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return FLDR_ERR_INVALID_ARGUMENT;
    if ((long long)H * W >= (1ll << 31)) return FLDR_ERR_UNSUPPORTED;
    if (mode == FLDR_SPLAT_LINEAR && !has_metric) return FLDR_ERR_UNSUPPORTED;   // softSplat.py:328 needs tenMetric
    g.N = N; g.C = C; g.H = H; g.W = W;
    g.mode = mode;
    g.CA = C + (mode_has_norm(mode) ? 1 : 0);
    g.CP = (g.CA + 3) / 4 * 4;
    g.has_metric = (has_metric && (mode == FLDR_SPLAT_LINEAR || mode == FLDR_SPLAT_SOFTMAX)) ? 1 : 0;
    return FLDR_OK;
}

static unsigned grid_for(long long total, int block) {
    long long b = (total + block - 1) / block;
    const long long cap = (long long)sm_count() * 64;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace fldr

using namespace fldr;

extern "C" size_t fldr_splat_fwd_workspace_bytes(int mode, int N, int C, int H, int W) {
    SplatGeom g;
    if (make_geom(mode, N, C, H, W, true, g) != FLDR_OK) return 0;
    return align_up((size_t)N * H * W * g.CP * sizeof(float), 256);
}

extern "C" int fldr_splat_fwd(int mode, const float* in, const int64_t* in_strides, const float* flow,
                              const int64_t* flow_strides, const float* metric, const int64_t* metric_strides,
                              float* out, float* norm, int N, int C, int H, int W, void* ws, size_t ws_bytes,
                              fldr_stream_t stream) {
    SplatGeom g;
    int st = make_geom(mode, N, C, H, W, metric != nullptr, g);
    if (st != FLDR_OK) return st;
    if (!in || !in_strides || !flow || !flow_strides || !out || (metric && !metric_strides)) return FLDR_ERR_INVALID_ARGUMENT;
    const size_t need = fldr_splat_fwd_workspace_bytes(mode, N, C, H, W);
    if (!ws || ws_bytes < need) return FLDR_ERR_WORKSPACE_TOO_SMALL;
    if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return FLDR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    float* acc = static_cast<float*>(ws);
    cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)N * H * W * g.CP * sizeof(float), s);
    if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
    const int Q = g.CP / 4;
    if ((long long)N * Q > 65535) return FLDR_ERR_UNSUPPORTED;
    const View4 vin = make_view(in, in_strides), vfl = make_view(flow, flow_strides), vme = make_view(metric, metric_strides);
    {
        // small frames: short row runs and narrow blocks keep enough CTAs in flight; large frames: 16-row runs
        const bool small = (long long)H * W * Q * N < 512 * 1024;
        const int bx = W >= 128 ? 128 : ((W + 31) / 32) * 32;
        dim3 grid((W + bx - 1) / bx, small ? (H + 3) / 4 : (H + 15) / 16, N * Q);
        const int wkind = !g.has_metric ? 0 : (mode == FLDR_SPLAT_SOFTMAX ? 1 : 2);
        const bool pre = mode == FLDR_SPLAT_SOFTMAX;
#define FLDR_LAUNCH_SCATTER(WK_, PRE_)                                                                      \
    do {                                                                                                    \
        if (small) splat_scatter_merged_kernel<4, WK_, PRE_><<<grid, bx, 0, s>>>(vin, vfl, vme, acc, g, Q);  \
        else splat_scatter_merged_kernel<16, WK_, PRE_><<<grid, bx, 0, s>>>(vin, vfl, vme, acc, g, Q);       \
    } while (0)
        if (wkind == 1) FLDR_LAUNCH_SCATTER(1, true);
        else if (wkind == 2) FLDR_LAUNCH_SCATTER(2, false);
        else if (pre) FLDR_LAUNCH_SCATTER(0, true);
        else FLDR_LAUNCH_SCATTER(0, false);
#undef FLDR_LAUNCH_SCATTER
    }
    if ((st = check_launch()) != FLDR_OK) return st;
    {
        const long long HW = (long long)H * W;
        const bool px4 = (HW % 4 == 0) && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                         (!norm || (reinterpret_cast<uintptr_t>(norm) & 15) == 0);
        if (px4) {
            dim3 grid((unsigned)((HW / 4 + 255) / 256), N * Q, 1);
            splat_normalise_kernel<4><<<grid, 256, 0, s>>>(acc, out, norm, g, Q);
        } else {
            dim3 grid((unsigned)((HW + 255) / 256), N * Q, 1);
            splat_normalise_kernel<1><<<grid, 256, 0, s>>>(acc, out, norm, g, Q);
        }
    }
    return check_launch();
}

extern "C" size_t fldr_splat_bwd_workspace_bytes(int mode, int N, int C, int H, int W) {
    (void)mode; (void)N; (void)C; (void)H; (void)W;
    return 0;
}

extern "C" int fldr_splat_bwd(int mode, const float* in, const int64_t* in_strides, const float* flow,
                              const int64_t* flow_strides, const float* metric, const int64_t* metric_strides,
                              const float* out, const float* norm, const float* grad_out,
                              const int64_t* grad_out_strides, float* grad_in, float* grad_flow, float* grad_metric,
                              int N, int C, int H, int W, void* ws, size_t ws_bytes, fldr_stream_t stream) {
    (void)ws; (void)ws_bytes;
    SplatGeom g;
    int st = make_geom(mode, N, C, H, W, metric != nullptr, g);
    if (st != FLDR_OK) return st;
    if (!in || !in_strides || !flow || !flow_strides || !grad_out || !grad_out_strides || (metric && !metric_strides))
        return FLDR_ERR_INVALID_ARGUMENT;
    if (mode_has_norm(mode) && (!out || !norm)) return FLDR_ERR_INVALID_ARGUMENT;
    if (grad_metric && !g.has_metric) return FLDR_ERR_INVALID_ARGUMENT;
    if (!grad_in && !grad_flow && !grad_metric) return FLDR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long total = (long long)N * H * W;
    splat_bwd_kernel<<<grid_for(total, 256), 256, 0, s>>>(make_view(in, in_strides), make_view(flow, flow_strides),
                                                          make_view(metric, metric_strides), out, norm,
                                                          make_view(grad_out, grad_out_strides), grad_in, grad_flow,
                                                          grad_metric, g);
    return check_launch();
}
