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
// Pass 1: scatter.  One thread per source pixel, lanes along x (coalesced plane reads; neighbouring lanes
// hit neighbouring accumulator pixels for smooth flow).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) splat_scatter_kernel(View4 in, View4 flow, View4 metric,
                                                            float* __restrict__ acc, SplatGeom g) {
    const long long HW = (long long)g.H * g.W;
    const long long total = HW * g.N;
    const bool pre = g.mode == FLDR_SPLAT_SOFTMAX;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % g.W);
        const int y = (int)((idx / g.W) % g.H);
        const int n = (int)(idx / HW);
        const float* fp = flow.p + n * flow.sn + y * flow.sh + x * flow.sw;
        Corners k;
        if (!make_corners(x, y, __ldg(fp), __ldg(fp + flow.sc), g.W, g.H, k)) continue;
        const float m = source_weight(g, metric, n, y, x);
        float* accn = acc + (long long)n * HW * g.CP;
        float* dst[4];
        dst[0] = accn + ((long long)k.y0 * g.W + k.x0) * g.CP;
        dst[1] = dst[0] + g.CP;
        dst[2] = dst[0] + (long long)g.W * g.CP;
        dst[3] = dst[2] + g.CP;
        const float* ip = in.p + n * in.sn + y * in.sh + x * in.sw;
        for (int q = 0; q < g.CP; q += 4) {
            float a[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = q + j;
                float val = 0.f;
                if (c < g.C) {
                    float xv = __ldg(ip + c * in.sc);
                    if (pre) xv = (xv + 1.f) * 0.5f;   // softSplat.py:334
                    val = xv * m;                      // softSplat.py:328 / 338
                } else if (c == g.C && g.CA > g.C) {
                    val = m;                           // normaliser channel
                }
                a[j] = val;
            }
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4)
                if (k.valid[c4])
                    red_add_v4(dst[c4] + q, a[0] * k.w[c4], a[1] * k.w[c4], a[2] * k.w[c4], a[3] * k.w[c4]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Pass 2: normalise + post-scale + interleaved -> NCHW.  One thread per target pixel.
//   softSplat.py:343-349: norm==0 -> 1, divide, (y - 0.5) * 2 (post-scale in every mode but RAW).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) splat_normalise_kernel(const float* __restrict__ acc, float* __restrict__ out,
                                                              float* __restrict__ norm_out, SplatGeom g) {
    const long long HW = (long long)g.H * g.W;
    const long long total = HW * g.N;
    const bool has_norm = g.CA > g.C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long pix = idx % HW;
        const int n = (int)(idx / HW);
        const float4* a4 = reinterpret_cast<const float4*>(acc + idx * g.CP);
        float d = 1.f;
        if (has_norm) {
            const float nrm = acc[idx * g.CP + g.C];
            if (norm_out) norm_out[idx] = nrm;
            d = (nrm == 0.f) ? 1.f : nrm;
        }
        float* op = out + (long long)n * g.C * HW + pix;
        for (int q = 0; q < g.CP; q += 4) {
            const float4 s4 = a4[q >> 2];
            const float s[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = q + j;
                if (c < g.C) {
                    float yv;
                    if (g.mode == FLDR_SPLAT_RAW) yv = s[j];
                    else if (!has_norm) yv = (s[j] - 0.5f) * 2.f;
                    else yv = (s[j] / d - 0.5f) * 2.f;
                    op[(long long)c * HW] = yv;
                }
            }
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
    const long long total = (long long)N * H * W;
    splat_scatter_kernel<<<grid_for(total, 256), 256, 0, s>>>(make_view(in, in_strides), make_view(flow, flow_strides),
                                                              make_view(metric, metric_strides), acc, g);
    if ((st = check_launch()) != FLDR_OK) return st;
    splat_normalise_kernel<<<grid_for(total, 256), 256, 0, s>>>(acc, out, norm, g);
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
