// Backward warp and splat metric of fLDRnet, sm_100a.
//
// Replaces DCTVFInet.bwarp (fLDRnet.py:546-581: mesh grid + flow, normalise, grid_sample of the image, grid_sample of
// a ones tensor, two masked_fill_, multiply - about ten torch kernels and three full-size temporaries) by ONE gather
// kernel, and the metric  z = mean_c(z_alpha * |I_ref - bwarp(I_src, flow)|)  (fLDRnet.py:442-446) by the same kernel
// with a fused epilogue, so the warped image never reaches memory.
//
// Coordinates follow the reference operation by operation (the __f*_rn intrinsics keep nvcc from contracting what torch
// evaluates as separate kernels, and fuse exactly what grid_sample fuses), so the sample positions - and with them the
// 0.999 mask decision - are bit-identical to the torch path:
//   X = x + u;  g = 2*X * (1/max(W-1, 1)) - 1        (fLDRnet.py:562-566; "/ scalar" as torch's CUDA kernel does it)
//   ix = ((g + 1) * W - 1) / 2                        (grid_sample, align_corners=False default, :568)
// i.e. ix = X*W/(W-1) - 0.5: the reference normalises for align_corners=True and samples with False; preserved.
// Bilinear, zero padding; mask = (sum of in-frame weights >= 0.999)  (fLDRnet.py:569-574).
#include "common.cuh"

namespace fldr {

struct WarpTaps {
    int off[4];          // element offset inside one channel plane, clamped into the frame (always loadable)
    unsigned keepbits[4];  // all ones when the tap is inside the frame, else 0: ANDed onto the loaded bits, so a tap
                         // outside the frame contributes exactly nothing (also when the clamped pixel holds NaN / inf)
    float w[4];          // bilinear weight, 0 for taps outside the frame
    bool keep;           // mask decision (true when masking is off)
};

// Per-axis constants of the coordinate transform, computed once on the host (IEEE float division there gives the same
// bits as on the device; doing them per thread cost four division sequences per pixel pair).
struct WarpAxis {
    float inv_sm1;      // 1 / max(size-1, 1)          convention 0: "x / scalar" on CUDA is x * (1/scalar)
    float step;         // 2 / (size-1)                convention 1: torch.linspace(-1, 1, size) step (0 when size == 1)
    float inv_half;     // 1 / ((size-1)/2)            convention 1: flow / ((size-1)/2), same CUDA rule
    float half_size;    // size / 2
};
struct WarpConsts { WarpAxis x, y; };

static WarpAxis make_axis(int size) {
    WarpAxis a;
    a.inv_sm1 = 1.0f / (float)(size - 1 > 1 ? size - 1 : 1);
    a.step = size > 1 ? 2.0f / (float)(size - 1) : 0.0f;
    a.inv_half = 1.0f / (((float)size - 1.0f) * 0.5f);       // inf for size == 1, like the reference's division by zero
    a.half_size = (float)size * 0.5f;
    return a;
}

// convention 0 - fLDRnet.bwarp (fLDRnet.py:562-568):   g = 2*(i + d) / max(size-1, 1) - 1
// convention 1 - PWCNet Backward (PWCNet.py:117-137):   g = linspace(-1, 1, size)[i] + d / ((size-1)/2)
//   torch.linspace on the CPU (where the reference builds its grid, :130) is  fma(step, i, -1)  for the first half and
//   fma(-step, size-1-i, 1)  for the second, step = 2/(size-1) in float32 - reproduced bit for bit (checked for sizes
//   17..4096 in tests/test_oracle.py)
// "tensor / python scalar" on CUDA is a multiplication by the float32 reciprocal (torch BinaryDivTrueKernel.cu), one ulp
// away from the true division the CPU performs; the reference runs on the GPU, so that is what is followed.
// both: ix = ((g + 1) * size - 1) / 2   (grid_sample's align_corners=False default), rounded once
__device__ __forceinline__ float warp_source_index(int i, float d, int size, const WarpAxis& ax, int convention) {
    float g;
    if (convention == 0) {
        g = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, __fadd_rn((float)i, d)), ax.inv_sm1), 1.0f);
    } else {
        const float lin = size > 1 ? (i < size / 2 ? __fmaf_rn(ax.step, (float)i, -1.0f) : __fmaf_rn(-ax.step, (float)(size - 1 - i), 1.0f))
                                   : -1.0f;
        g = __fadd_rn(lin, __fmul_rn(d, ax.inv_half));
    }
    // ((g + 1) * size - 1) / 2 with ONE rounding, as torch's grid_sample evaluates it on the CPU vector path and on
    // CUDA (size/2 and 0.5 are exact, so this fma is that expression rounded once)
    return __fmaf_rn(__fadd_rn(g, 1.0f), ax.half_size, -0.5f);
}

// sh, sw: row / pixel stride of the source plane in elements (32-bit: check_warp_args bounds the plane span)
__device__ __forceinline__ WarpTaps warp_taps(float u, float v, int x, int y, int H, int W, int sh, int sw, bool with_mask,
                                              int convention, const WarpConsts& wc) {
    WarpTaps t;
    const float ix = warp_source_index(x, u, W, wc.x, convention);
    const float iy = warp_source_index(y, v, H, wc.y, convention);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float fx1 = fx0 + 1.0f, fy1 = fy0 + 1.0f;
    // compare as floats: NaN / huge coordinates fail every test and sample nothing (torch does the same)
    const bool inx0 = fx0 >= 0.f && fx0 <= (float)(W - 1), inx1 = fx1 >= 0.f && fx1 <= (float)(W - 1);
    const bool iny0 = fy0 >= 0.f && fy0 <= (float)(H - 1), iny1 = fy1 >= 0.f && fy1 <= (float)(H - 1);
    // clamped integer corners: the conversion is defined for any input (NaN -> 0) and every address is inside the plane
    const int x0 = (int)fminf(fmaxf(fx0, 0.f), (float)(W - 1)), x1 = (int)fminf(fmaxf(fx1, 0.f), (float)(W - 1));
    const int y0 = (int)fminf(fmaxf(fy0, 0.f), (float)(H - 1)), y1 = (int)fminf(fmaxf(fy1, 0.f), (float)(H - 1));
    const int r0 = y0 * sh, r1 = y1 * sh, c0 = x0 * sw, c1 = x1 * sw;
    t.off[0] = r0 + c0; t.off[1] = r0 + c1; t.off[2] = r1 + c0; t.off[3] = r1 + c1;
    const bool ok[4] = {inx0 && iny0, inx1 && iny0, inx0 && iny1, inx1 && iny1};
    const float w[4] = {__fmul_rn(fx1 - ix, fy1 - iy), __fmul_rn(ix - fx0, fy1 - iy), __fmul_rn(fx1 - ix, iy - fy0),
                        __fmul_rn(ix - fx0, iy - fy0)};
    float msum = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        t.keepbits[k] = ok[k] ? 0xffffffffu : 0u;
        t.w[k] = ok[k] ? w[k] : 0.f;
        msum = __fadd_rn(msum, t.w[k]);          // + 0 for skipped taps: same value as skipping them
    }
    // fLDRnet: mask < 0.999 -> 0 (fLDRnet.py:573);  PWC-Net: mask > 0.999 -> 1, the rest 0 (PWCNet.py:139-141)
    t.keep = !with_mask || (convention == 0 ? msum >= 0.999f : msum > 0.999f);
    return t;
}

// Two horizontally adjacent pixels per thread, CU channels per pass: all 2 x 4 x CU gathers of a pass are issued,
// unpredicated, before the first is consumed.  The grid is (pairs of a row, row, sample), so no thread divides.
// (History, 4K image warp: one pixel per thread + channel loop 200 us; this shape with per-tap predicates and 64-bit
// index div/mod 150-160 us at 720 instructions per thread, issue-bound.)
// EXACT: C == CU, no channel loop or tail tests.
// METRIC: out is [N,1,H,W] = mean_c(alpha * |ref - warp|), else out is [N,C,H,W] = warp.
template <bool METRIC, int CU, bool EXACT>
__global__ void __launch_bounds__(128) bwarp_kernel(View4 src, View4 ref, View4 flow, float* __restrict__ out, int C_,
                                                    int H, int W, float alpha, int with_mask, int y_base,
                                                    int convention, const __grid_constant__ WarpConsts wc) {
    const int C = EXACT ? CU : C_;
    const int x = (blockIdx.x * 128 + threadIdx.x) * 2, y = y_base + blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const bool two = x + 1 < W;
    const long long HW = (long long)H * W;
    const float* fp = flow.p + n * flow.sn + y * flow.sh + x * flow.sw;
    const float* fq = two ? fp + flow.sw : fp;                 // odd-width tail: re-read pixel 0, result unused
    const float u0 = __ldcs(fp), v0 = __ldcs(fp + flow.sc), u1 = __ldcs(fq), v1 = __ldcs(fq + flow.sc);
    WarpTaps t[2];
    t[0] = warp_taps(u0, v0, x, y, H, W, (int)src.sh, (int)src.sw, with_mask != 0, convention, wc);
    t[1] = warp_taps(u1, v1, two ? x + 1 : x, y, H, W, (int)src.sh, (int)src.sw, with_mask != 0, convention, wc);
    const float* sp = src.p + n * src.sn;
    const float* rp = METRIC ? ref.p + n * ref.sn + y * ref.sh + x * ref.sw : nullptr;
    const int rstep = (METRIC && two) ? (int)ref.sw : 0;
    const long long idx = (long long)y * W + x;
    const bool vec = two && ((W & 1) == 0);          // idx even and rows even-sized: 8-byte aligned pair
    float sum[2] = {0.f, 0.f};
    for (int c0 = 0; c0 < C; c0 += CU) {
        float g[CU][2][4], r[CU][2];
#pragma unroll
        for (int cc = 0; cc < CU; ++cc) {
            const int c = EXACT ? cc : min(c0 + cc, C - 1);       // tail channels re-read the last one, result unused
            const float* plane = sp + (long long)c * src.sc;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    g[cc][p][k] = __uint_as_float(__float_as_uint(__ldg(plane + t[p].off[k])) & t[p].keepbits[k]);
                if (METRIC) r[cc][p] = __ldcs(rp + (long long)c * ref.sc + p * rstep);
            }
        }
#pragma unroll
        for (int cc = 0; cc < CU; ++cc) {
            if (!EXACT && c0 + cc >= C) break;
            float w[2];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) acc = fmaf(g[cc][p][k], t[p].w[k], acc);    // contracted, as torch's CUDA grid_sample is
                w[p] = t[p].keep ? acc : 0.f;
                if (METRIC) sum[p] = __fadd_rn(sum[p], __fmul_rn(alpha, fabsf(__fsub_rn(r[cc][p], w[p]))));
            }
            if (!METRIC) {
                float* op = out + ((long long)n * C + c0 + cc) * HW + idx;
                if (vec) __stcs(reinterpret_cast<float2*>(op), make_float2(w[0], w[1]));
                else { __stcs(op, w[0]); if (two) __stcs(op + 1, w[1]); }
            }
        }
    }
    if (METRIC) {
        float* op = out + (long long)n * HW + idx;
        // torch.mean on CUDA multiplies the sum by the float32 factor 1/C (ReduceMomentKernel.cu)
        const float inv_c = __fdiv_rn(1.0f, (float)C);
        const float z0 = __fmul_rn(sum[0], inv_c), z1 = __fmul_rn(sum[1], inv_c);
        if (vec) __stcs(reinterpret_cast<float2*>(op), make_float2(z0, z1));
        else { __stcs(op, z0); if (two) __stcs(op + 1, z1); }
    }
}

template <bool METRIC>
static int launch_bwarp(const View4& src, const View4& ref, const View4& flow, float* out, int N, int C, int H, int W,
                        float alpha, int with_mask, int convention, cudaStream_t s) {
    const int W2 = (W + 1) / 2;
    const WarpConsts wc = {make_axis(W), make_axis(H)};
    for (int y0 = 0; y0 < H; y0 += 65535) {          // gridDim.y limit; one launch for every frame under 65 536 rows
        const int rows = H - y0 < 65535 ? H - y0 : 65535;
        dim3 grid((unsigned)((W2 + 127) / 128), (unsigned)rows, (unsigned)N);
#define FLDR_BWARP(CU, EXACT) bwarp_kernel<METRIC, CU, EXACT><<<grid, 128, 0, s>>>(src, ref, flow, out, C, H, W, alpha, with_mask, y0, convention, wc)
        if (C == 1) FLDR_BWARP(1, true);
        else if (C == 2) FLDR_BWARP(2, true);
        else if (C == 3) FLDR_BWARP(3, true);
        else if (C == 4) FLDR_BWARP(4, true);
        else FLDR_BWARP(4, false);
#undef FLDR_BWARP
        const int st = check_launch();
        if (st != FLDR_OK) return st;
    }
    return FLDR_OK;
}

// Backward of the warp (autograd of fLDRnet.py:556-578 / PWCNet.py:134-143; the mask is piecewise constant and
// carries no gradient, floor() neither):
//   grad_x[c, tap_k]  += keep * w_k * g[c]                                  (scatter, red.global.add.f32)
//   d out / d ix = keep * sum_c g[c] * ((x_ne - x_nw) * (y1 - iy) + (x_se - x_sw) * (iy - y0))      taps outside = 0
//   d out / d iy = keep * sum_c g[c] * ((x_sw - x_nw) * (x1 - ix) + (x_se - x_ne) * (ix - x0))
//   grad_flow = (d/d ix * W/(W-1), d/d iy * H/(H-1))        both conventions: (2/(W-1)) * (W/2)
// One thread per pixel, channel loop; grad_x is zero-filled by the entry point.
__global__ void __launch_bounds__(128) bwarp_bwd_kernel(View4 src, View4 flow, View4 gout, float* __restrict__ grad_x,
                                                        float* __restrict__ grad_flow, int C, int H, int W, int with_mask,
                                                        int y_base, int convention, const __grid_constant__ WarpConsts wc) {
    const int x = blockIdx.x * 128 + threadIdx.x, y = y_base + blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const long long HW = (long long)H * W, idx = (long long)y * W + x;
    const float* fp = flow.p + n * flow.sn + y * flow.sh + x * flow.sw;
    const float u = __ldg(fp), v = __ldg(fp + flow.sc);
    // taps relative to a CONTIGUOUS plane for grad_x (sh = W, sw = 1) and to the source view for the x reads
    const WarpTaps ts = warp_taps(u, v, x, y, H, W, (int)src.sh, (int)src.sw, with_mask != 0, convention, wc);
    const WarpTaps tg = warp_taps(u, v, x, y, H, W, W, 1, with_mask != 0, convention, wc);
    const float ix = warp_source_index(x, u, W, wc.x, convention), iy = warp_source_index(y, v, H, wc.y, convention);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float ax = ix - fx0, bx = (fx0 + 1.0f) - ix, ay = iy - fy0, by = (fy0 + 1.0f) - iy;
    const float* gp = gout.p + n * gout.sn + y * gout.sh + x * gout.sw;
    const float* sp = src.p + n * src.sn;
    float dix = 0.f, diy = 0.f;
    for (int c = 0; c < C; ++c) {
        const float g = ts.keep ? __ldg(gp + c * gout.sc) : 0.f;
        if (grad_x) {
            float* gx = grad_x + ((long long)n * C + c) * HW;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (tg.keepbits[k] && g != 0.f) atomicAdd(gx + tg.off[k], g * tg.w[k]);
        }
        if (grad_flow) {
            const float* plane = sp + (long long)c * src.sc;
            float t[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] = __uint_as_float(__float_as_uint(__ldg(plane + ts.off[k])) & ts.keepbits[k]);
            dix += g * ((t[1] - t[0]) * by + (t[3] - t[2]) * ay);
            diy += g * ((t[2] - t[0]) * bx + (t[3] - t[1]) * ax);
        }
    }
    if (grad_flow) {
        const bool finite = (ix == ix) && (iy == iy) && fabsf(ix) < 3.0e38f && fabsf(iy) < 3.0e38f;
        float* gf = grad_flow + (long long)n * 2 * HW + idx;
        gf[0] = finite ? dix * ((float)W / (float)max(W - 1, 1)) : 0.f;
        gf[HW] = finite ? diy * ((float)H / (float)max(H - 1, 1)) : 0.f;
    }
}

static int check_warp_args(int N, int C, int H, int W, const View4& src) {
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return FLDR_ERR_INVALID_ARGUMENT;
    if ((long long)H * W >= (1ll << 31) || N > 65535) return FLDR_ERR_UNSUPPORTED;
    // tap offsets inside one plane are 32-bit
    if (src.sh < 0 || src.sw < 0) return FLDR_ERR_UNSUPPORTED;
    if ((long long)(H + 2) * src.sh + (long long)(W + 2) * src.sw >= (1ll << 31)) return FLDR_ERR_UNSUPPORTED;
    return FLDR_OK;
}

}  // namespace fldr

using namespace fldr;

extern "C" int fldr_bwarp_fwd(const float* x, const int64_t* x_strides, const float* flow, const int64_t* flow_strides,
                              float* out, int N, int C, int H, int W, int with_mask, int convention,
                              fldr_stream_t stream) {
    if (!x || !x_strides || !flow || !flow_strides || !out) return FLDR_ERR_INVALID_ARGUMENT;
    if (convention != 0 && convention != 1) return FLDR_ERR_INVALID_ARGUMENT;
    const View4 vx = make_view(x, x_strides), vf = make_view(flow, flow_strides);
    const int st = check_warp_args(N, C, H, W, vx);
    if (st != FLDR_OK) return st;
    return launch_bwarp<false>(vx, vx, vf, out, N, C, H, W, 0.f, with_mask, convention, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int fldr_warp_metric_fwd(const float* ref, const int64_t* ref_strides, const float* src,
                                    const int64_t* src_strides, const float* flow, const int64_t* flow_strides, float alpha,
                                    float* out, int N, int C, int H, int W, int with_mask, fldr_stream_t stream) {
    if (!ref || !ref_strides || !src || !src_strides || !flow || !flow_strides || !out) return FLDR_ERR_INVALID_ARGUMENT;
    const View4 vr = make_view(ref, ref_strides), vs = make_view(src, src_strides), vf = make_view(flow, flow_strides);
    const int st = check_warp_args(N, C, H, W, vs);
    if (st != FLDR_OK) return st;
    return launch_bwarp<true>(vs, vr, vf, out, N, C, H, W, alpha, with_mask, 0, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int fldr_bwarp_bwd(const float* x, const int64_t* x_strides, const float* flow, const int64_t* flow_strides,
                              const float* grad_out, const int64_t* grad_out_strides, float* grad_x, float* grad_flow,
                              int N, int C, int H, int W, int with_mask, int convention, fldr_stream_t stream) {
    if (!x || !x_strides || !flow || !flow_strides || !grad_out || !grad_out_strides) return FLDR_ERR_INVALID_ARGUMENT;
    if (convention != 0 && convention != 1) return FLDR_ERR_INVALID_ARGUMENT;
    const View4 vx = make_view(x, x_strides), vf = make_view(flow, flow_strides), vg = make_view(grad_out, grad_out_strides);
    const int st = check_warp_args(N, C, H, W, vx);
    if (st != FLDR_OK) return st;
    if (!grad_x && !grad_flow) return FLDR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (grad_x) {
        const cudaError_t e = cudaMemsetAsync(grad_x, 0, (size_t)N * C * H * W * sizeof(float), s);
        if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
    }
    const WarpConsts wc = {make_axis(W), make_axis(H)};
    for (int y0 = 0; y0 < H; y0 += 65535) {
        const int rows = H - y0 < 65535 ? H - y0 : 65535;
        dim3 grid((unsigned)((W + 127) / 128), (unsigned)rows, (unsigned)N);
        bwarp_bwd_kernel<<<grid, 128, 0, s>>>(vx, vf, vg, grad_x, grad_flow, C, H, W, with_mask, y0, convention, wc);
        const int st2 = check_launch();
        if (st2 != FLDR_OK) return st2;
    }
    return FLDR_OK;
}
