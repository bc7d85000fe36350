// Input pyramid on the device (second half): main.py:855-856 (test) / 562-563 (train) build level i of
// `input_gpu` as F.interpolate(full-resolution frames, scale_factor = scales[0] / scales[i], mode='bicubic') ON THE CPU and copy
// every level to the GPU.  Here the frames are uploaded once and every level is produced from them in ONE pass:
//   pyramid_pow2_kernel   factors 1/2 .. 1/32, align_corners = False (every shipped preset).  src = 2^k (dst + 0.5) - 0.5 has
//                         fractional part 0.5 on every level, so the four cubic weights are the constants (-3/32, 19/32, 19/32,
//                         -3/32) and the taps of level k are columns / rows 2^k d + 2^(k-1) - 2 ... + 1: inside an aligned
//                         32 x 128 block for k >= 2, one pixel of halo for k = 1.  A warp owns such a block: a lane reads one
//                         float4 per row (512 contiguous bytes per warp), forms the horizontal sums of all levels from its own
//                         four pixels and three shuffled neighbours, and carries the vertical sums in registers while it walks
//                         down 34 rows - the frame is read once, no shared memory, no intermediate.
//   pyramid_generic_kernel  any factor, align_corners True or False: one thread per output pixel, index and weight arithmetic in
//                         float32 exactly as ATen's upsample_bicubic2d (UpSample.h: area_pixel_compute_source_index,
//                         guard_index_and_lambda, get_cubic_upsample_coefficients, A = -0.75), taps clamped to the frame.
// Summation order of both: horizontal first, partial sums left to right, like ATen's Interpolate<2>.
// Algorithmic bytes: 4 * planes * H * W * (1 + sum_k 4^-k).  Bound: HBM.
#include "common.cuh"

namespace fldr {
namespace pyr {
constexpr int MAXL = 5;                   // levels of the single-pass kernel
constexpr int SW = 128, SH = 32;          // block of a warp
constexpr int WARPS = 8;
struct Levels {
    float* p[MAXL];                       // level k+1: [planes][H >> (k+1)][W >> (k+1)] contiguous
};
}  // namespace pyr

#ifndef PYR_D
#define PYR_D 4
#endif
#ifndef PYR_MINB
#define PYR_MINB 4
#endif
#define C0 (-0.09375f)
#define C1 (0.59375f)

__device__ __forceinline__ float cubic_h(float a, float b, float c, float d) {       // ((a w0 + b w1) + c w2) + d w3
    return fmaf(d, C0, fmaf(c, C1, fmaf(b, C1, a * C0)));
}

__global__ void __launch_bounds__(pyr::WARPS * 32, PYR_MINB) pyramid_pow2_kernel(const float* __restrict__ im, long long s_plane, long long s_row,
                                                                       pyr::Levels out, int planes, int H, int W, int n_levels,
                                                                       int strips_x, int strips_y) {
    using namespace pyr;
    const int lane = threadIdx.x & 31;
    const long long strip = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (strip >= (long long)planes * strips_y * strips_x) return;
    const int sx = (int)(strip % strips_x);
    const int sy = (int)((strip / strips_x) % strips_y);
    const int pl = (int)(strip / ((long long)strips_x * strips_y));
    const int x0 = sx * SW, r0 = sy * SH;
    const int col = x0 + 4 * lane;
    const bool active = col < W;
    const bool last_lane = col + 4 >= W;                  // right neighbour is the clamped border column (own .w)
    const float* base = im + pl * s_plane;
    const int rows = min(SH, H - r0);                     // H is a multiple of 2^n_levels, not necessarily of 32

    constexpr int D = PYR_D;                              // rows in flight per warp (512 B each)
    float4 v[D];
    float ev[D];                                          // column x0 - 1 (lane 0) / x0 + 128 (lane 31) of the neighbouring blocks
    const int ecol = lane == 0 ? x0 - 1 : x0 + SW;
    const bool eload = (lane == 0 && x0 > 0) || (lane == 31 && !last_lane);
    auto load_row = [&](int j, float4& q, float& e) {
        const int gy = min(max(r0 - 1 + j, 0), H - 1);
        const float* rp = base + gy * s_row;
        q = active ? __ldcs(reinterpret_cast<const float4*>(rp + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
        e = eload ? __ldg(rp + ecol) : 0.f;
    };
#pragma unroll
    for (int j = 0; j < D; ++j) load_row(j, v[j], ev[j]);

    float a1x = 0.f, a1y = 0.f, b1x = 0.f, b1y = 0.f;     // level 1: current / previous output row, two columns per lane
    float a2 = 0.f, aP = 0.f;                             // level 2; levels 3-5 (their tap rows are disjoint)
    const int W1 = W >> 1, W2 = W >> 2;
    float* o1 = out.p[0] + ((long long)pl * (H >> 1) + (r0 >> 1)) * W1 + (x0 >> 1) + 2 * lane;
    float* o2 = n_levels >= 2 ? out.p[1] + ((long long)pl * (H >> 2) + (r0 >> 2)) * W2 + (x0 >> 2) + lane : nullptr;

#pragma unroll
    for (int j = 0; j < SH + 2; ++j) {
        const float4 q = v[j % D];
        const float qe = ev[j % D];
        if (j + D < SH + 2) load_row(j + D, v[j % D], ev[j % D]);
        if (j - 1 > rows) break;                          // (warp-uniform) nothing below the frame's last strip rows + halo
        float lz = __shfl_up_sync(0xffffffffu, q.z, 1);
        float lw = __shfl_up_sync(0xffffffffu, q.w, 1);
        float rx = __shfl_down_sync(0xffffffffu, q.x, 1);
        if (lane == 0) lw = x0 > 0 ? qe : q.x;            // column -1 clamps to column 0
        if (last_lane) rx = q.w; else if (lane == 31) rx = qe;
        // ---- level 1: rows 2k-1 .. 2k+2 of output k = local j 2k .. 2k+3
        const float h1x = cubic_h(lw, q.x, q.y, q.z), h1y = cubic_h(q.y, q.z, q.w, rx);
        if ((j & 1) == 0) {
            b1x = fmaf(h1x, C1, b1x); b1y = fmaf(h1y, C1, b1y);
            a1x = h1x * C0; a1y = h1y * C0;
        } else {
            b1x = fmaf(h1x, C0, b1x); b1y = fmaf(h1y, C0, b1y);
            const int k = (j >> 1) - 1;
            if (k >= 0 && 2 * k < rows && active) *reinterpret_cast<float2*>(o1 + (long long)k * W1) = make_float2(b1x, b1y);
            b1x = fmaf(h1x, C1, a1x); b1y = fmaf(h1y, C1, a1y);
        }
        const int i = j - 1;                              // local row of the block
        if (i >= 0 && i < SH) {
            // ---- level 2: rows 4m .. 4m+3, the lane's own four columns
            if (n_levels >= 2) {
                const float h2 = cubic_h(q.x, q.y, q.z, q.w);
                const int t = i & 3;
                a2 = t == 0 ? h2 * C0 : fmaf(h2, (t == 3) ? C0 : C1, a2);
                if (t == 3 && i < rows && active) o2[(long long)(i >> 2) * W2] = a2;
            }
            // ---- levels 3 / 4 / 5: columns 2^k d + 2^(k-1) - 2 .. + 1 = (.z .w) of the left lane and (.x .y) of this one
            const int lv = ((i & 7) >= 2 && (i & 7) <= 5) ? 3 : ((i & 15) >= 6 && (i & 15) <= 9) ? 4 : 5;   // i in 14..17 otherwise
            if (n_levels >= lv) {
                const float hp = cubic_h(lz, lw, q.x, q.y);
                const int first = lv == 3 ? 2 : lv == 4 ? 6 : 14;
                const int t = (i & ((1 << lv) - 1)) - first;
                aP = t == 0 ? hp * C0 : fmaf(hp, (t == 3) ? C0 : C1, aP);
                const int half = 1 << (lv - 3);           // owning lane: l = 2^(lv-2) d + 2^(lv-3)
                if (t == 3 && i < rows && active && (lane & (2 * half - 1)) == half) {
                    const int Wk = W >> lv;
                    out.p[lv - 1][((long long)pl * (H >> lv) + ((r0 + i) >> lv)) * Wk + (x0 >> lv) + (lane >> (lv - 2))] = aP;
                }
            }
        }
    }
}

__device__ __forceinline__ float conv1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float conv2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__device__ __forceinline__ void axis_taps(int dst, int in_size, float scale, int align_corners, int idx[4], float w[4]) {
    const float src = align_corners ? scale * (float)dst : scale * ((float)dst + 0.5f) - 0.5f;
    const int i0 = min((int)floorf(src), in_size - 1);
    const float lam = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
    const float om = 1.f - lam;
    w[0] = conv2(lam + 1.f); w[1] = conv1(lam); w[2] = conv1(om); w[3] = conv2(om + 1.f);
#pragma unroll
    for (int t = 0; t < 4; ++t) idx[t] = min(max(i0 - 1 + t, 0), in_size - 1);
}

// grid (ceil(ow / 128), oh, planes)
__global__ void __launch_bounds__(128) pyramid_generic_kernel(const float* __restrict__ im, long long s_plane, long long s_row,
                                                              float* __restrict__ out, int H, int W, int oh, int ow, float scale_y,
                                                              float scale_x, int align_corners) {
    const int ox = blockIdx.x * 128 + threadIdx.x, oy = blockIdx.y;
    if (ox >= ow) return;
    int ix[4], iy[4];
    float wx[4], wy[4];
    axis_taps(ox, W, scale_x, align_corners, ix, wx);
    axis_taps(oy, H, scale_y, align_corners, iy, wy);
    const float* base = im + blockIdx.z * s_plane;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float* rp = base + iy[r] * s_row;
        float h = __ldg(rp + ix[0]) * wx[0];
#pragma unroll
        for (int t = 1; t < 4; ++t) h = fmaf(__ldg(rp + ix[t]), wx[t], h);
        acc = r == 0 ? h * wy[0] : fmaf(h, wy[r], acc);
    }
    out[((long long)blockIdx.z * oh + oy) * ow + ox] = acc;
}
}  // namespace fldr

using namespace fldr;

extern "C" int fldr_bicubic_pyramid_fwd(const float* frames, int64_t plane_stride, int64_t row_stride, int planes, int H, int W,
                                        int n_levels, const double* scale_factors, int align_corners, float* const* out_levels,
                                        fldr_stream_t stream) {
    if (!frames || !scale_factors || !out_levels || planes < 0 || H <= 0 || W <= 0 || n_levels < 0) return FLDR_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < n_levels; ++i)
        if (!out_levels[i] || !(scale_factors[i] > 0.0) || (long long)floor((double)H * scale_factors[i]) < 1 ||
            (long long)floor((double)W * scale_factors[i]) < 1)
            return FLDR_ERR_INVALID_ARGUMENT;
    if (planes == 0 || n_levels == 0) return FLDR_OK;
    if (planes > 65535) return FLDR_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    bool pow2 = !align_corners && n_levels <= pyr::MAXL && W % 4 == 0 && row_stride % 4 == 0 && plane_stride % 4 == 0 &&
                ((uintptr_t)frames & 15) == 0 && H % (1 << n_levels) == 0 && W % (1 << n_levels) == 0;
    for (int i = 0; pow2 && i < n_levels; ++i)
        pow2 = scale_factors[i] == 1.0 / (double)(1 << (i + 1)) && ((uintptr_t)out_levels[i] & 7) == 0;
    if (pow2) {
        pyr::Levels lv;
        for (int i = 0; i < pyr::MAXL; ++i) lv.p[i] = i < n_levels ? out_levels[i] : nullptr;
        const int strips_x = (W + pyr::SW - 1) / pyr::SW, strips_y = (H + pyr::SH - 1) / pyr::SH;
        const long long strips = (long long)planes * strips_x * strips_y;
        const long long blocks = (strips + pyr::WARPS - 1) / pyr::WARPS;
        if (blocks > 0x7fffffffLL) return FLDR_ERR_INVALID_ARGUMENT;
        pyramid_pow2_kernel<<<(unsigned)blocks, pyr::WARPS * 32, 0, st>>>(frames, plane_stride, row_stride, lv, planes, H, W, n_levels,
                                                                           strips_x, strips_y);
        return check_launch();
    }
    for (int i = 0; i < n_levels; ++i) {
        const int oh = (int)floor((double)H * scale_factors[i]), ow = (int)floor((double)W * scale_factors[i]);
        if (oh > 65535) return FLDR_ERR_INVALID_ARGUMENT;
        // UpSample.h area_pixel_compute_scale<float>: (in - 1) / (out - 1) with align_corners, else 1 / scale_factor
        const float sy = align_corners ? (oh > 1 ? (float)(H - 1) / (float)(oh - 1) : 0.f) : (float)(1.0 / scale_factors[i]);
        const float sx = align_corners ? (ow > 1 ? (float)(W - 1) / (float)(ow - 1) : 0.f) : (float)(1.0 / scale_factors[i]);
        dim3 grid((ow + 127) / 128, oh, planes);
        pyramid_generic_kernel<<<grid, 128, 0, st>>>(frames, plane_stride, row_stride, out_levels[i], H, W, oh, ow, sy, sx, align_corners);
        int rc = check_launch();
        if (rc != FLDR_OK) return rc;
    }
    return FLDR_OK;
}
