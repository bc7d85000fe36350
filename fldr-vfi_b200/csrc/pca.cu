// Block-PCA feature extraction: pca_comp.py:473-528 `to_pca_diff`, the first device step of
// every fLDRnet forward (fLDRnet.py:146).  The reference unfolds the two stacked frames into 8x8 blocks (nn.Unfold + four
// reshape / permute copies), subtracts the mean block, multiplies by the eigenvector matrix in float64 (cuBLAS DGEMM
// [blocks, 64] x [64, 16]), optionally divides by mean_vec, permutes to [chan * 16, H/8, W/8], takes a global min / max
// (two reductions) and rescales to [-1, 1]: ~12 kernels and five float64 temporaries of the frame's size.  Here:
//   pass 1  pca_project_kernel   persistent CTAs; reads the float32 frame once (coalesced 8-row stripes, the next stripe's
//                                loads in flight while this one is projected; converted and centred into shared memory),
//                                projects every block on the eigenvectors in float64 with a 4 x 4 register tile
//                                per thread, in the order sum_j (x_j - mean_j) * EV[k][j],
//                                writes the result already permuted, and folds the global min / max into two 64-bit
//                                atomics (order-preserving bit pattern)
//   pass 2  pca_rescale_kernel   ((t - min) / (max - min)) * 2 - 1, float64 like the reference or straight to the float32
//                                the caller converts to (`.float()`, fLDRnet.py:146)
// Algorithmic bytes: 4 * chan * H * W read + chan * 16 * (H/8) * (W/8) * (8 | 4) written.  Bound: HBM.
#include "common.cuh"

namespace fldr {
namespace pca {
constexpr int WS = 8, NV = WS * WS;
constexpr int NBLK = 64;                  // blocks of a stripe per CTA (512 pixels x 8 rows)
#ifndef PCA_NT
#define PCA_NT 128
#endif
constexpr int NT = PCA_NT;                // warps of a CTA share one stripe
constexpr int MT = NBLK / (NT / 32) / 8;  // 8-block mma row tiles per warp
constexpr int PITCH = NBLK + 4;           // doubles between consecutive pixels j of the stripe in shared memory
}  // namespace pca

__device__ __forceinline__ unsigned long long ordered_bits(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);      // monotone in v
}
__host__ __device__ inline double from_ordered_bits(unsigned long long o) {
    const unsigned long long b = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
    double v;
    memcpy(&v, &b, sizeof(v));
    return v;
}

// Persistent CTAs (five per SM, two warps each) loop over 8 x 512 stripes, stripe index = (channel, block row, 64-block chunk).
//   phase 1  the float32 stripe - already in registers: its sixteen 16-byte loads per thread were issued BEFORE the previous
//            stripe's projection, so DRAM latency hides behind the float64 math - is converted to float64 and stored to shared
//            memory (the centring x - mean of pca_comp.py:502 is applied to the projected value: mean . EV[k]) as [pixel of the block j][block b] (pitch 68: the fragment loads
//            below are bank-conflict free);
//   phase 2  [64 blocks, 64] x [64, 16] on the float64 tensor cores: a warp owns 32 blocks = four 8-row tiles, the
//            eigenvectors live in REGISTERS as the B fragments of all sixteen k-steps (loaded once per CTA: no shared memory
//            for them at all), so a k-step is 4 LDS.64 (A fragments) + 8 mma.sync.m8n8k4.f64.
// The accumulation order inside an mma differs from the reference's cuBLAS DGEMM (whose order is not specified either); both
// are float64 sums of 64 products of O(1) values (parity bound 2e-13, tests/test_gpu_pca.py).
// History (6 x 2304 x 4096, profiles/r2_pca_kernel_history.txt): one element per iteration in phase 1 (sixteen dependent DRAM
// round trips per stripe) + DFMA register tile fed by 8 LDS per 16 DFMA: 378 us at 16 % of the float64 pipe; all loads in flight
// 241 us; persistent + prefetch, still DFMA: 219 us at 70 % of the shared-memory pipe; this version: see the file.
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(pca::NT) pca_project_kernel(const float* __restrict__ im, long long s_c, long long s_h,
                                                              const double* __restrict__ mean, const double* __restrict__ ev,
                                                              long long ev_stride, const double* __restrict__ mean_vec,
                                                              double* __restrict__ t, unsigned long long* __restrict__ mm,
                                                              int chan, int by, int bx, int ncomp, int vec, int chunks, int total) {
    using namespace pca;
    extern __shared__ __align__(16) double s_loc[];                       // [2][NV][PITCH]: pixel j of block b at j * PITCH + b, two stripes
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int g = lane >> 2, tg = lane & 3;                               // mma fragment coordinates
    const int W = bx * WS;
    // B fragments of every k-step: element (k = 4 ks + tg, n = 8 nt + g) = EV[n][k]   (EV.permute(1, 0) of line 507)
    double bfr[NV / 4][2];
#pragma unroll
    for (int ks = 0; ks < NV / 4; ++ks)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int n = nt * 8 + g;
            bfr[ks][nt] = (n < ncomp) ? __ldg(ev + (long long)n * ev_stride + ks * 4 + tg) : 0.0;
        }
    // (x - mean) . EV[k] = x . EV[k] - mean . EV[k]: the centring of line 502 becomes one subtraction per output instead of one per
    // pixel (float64 throughout; the two forms differ by rounding only, ~1e-15 on O(1) values)
    double rmv[2][2], bias[2][2];                  // 1 / mean_vec (1 when there is none) and mean . EV[k] of this thread's four components
    {
        // every warp forms all sixteen dot products: lane l multiplies pixels l and l + 32, butterfly sum (all loads in one round trip)
        double part[16];
        const double m0 = __ldg(mean + lane), m1 = __ldg(mean + lane + 32);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            part[k] = (k < ncomp) ? fma(m1, __ldg(ev + (long long)k * ev_stride + lane + 32), m0 * __ldg(ev + (long long)k * ev_stride + lane)) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int k = 0; k < 16; ++k) part[k] += __shfl_xor_sync(0xffffffffu, part[k], o);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int kk = nt * 8 + tg * 2 + e;
                rmv[nt][e] = (mean_vec && kk < ncomp) ? 1.0 / __ldg(mean_vec + kk) : 1.0;
                double bsum = 0.0;
#pragma unroll
                for (int k = 0; k < 16; ++k) bsum = (k == kk) ? part[k] : bsum;
                bias[nt][e] = bsum;
            }
    }
    double vmin = __longlong_as_double(0x7ff0000000000000ll), vmax = -vmin;
    constexpr int V4 = NBLK * WS / 4;              // float4s per stripe row
    constexpr int NQ = WS * V4 / NT;               // float4s per thread and stripe
    float4 q[NQ];
    auto decode = [&](int sidx, int& c, int& yb, int& xb0) {
        xb0 = (sidx % chunks) * NBLK;
        yb = (sidx / chunks) % by;
        c = sidx / (chunks * by);
    };
    auto prefetch = [&](int sidx) {
        int c, yb, xb0;
        decode(sidx, c, yb, xb0);
        const float* src = im + c * s_c + (long long)(yb * WS) * s_h + xb0 * WS;
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            const int e4 = tid + k * NT, r = e4 / V4, x4 = (e4 % V4) * 4;
            q[k] = (xb0 * WS + x4 < W) ? __ldcs(reinterpret_cast<const float4*>(src + r * s_h + x4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    // float4 k of the stripe held in q -> float64, centred, into stripe buffer `buf`
    auto stage = [&](double* buf, int k) {
        const int e4 = tid + k * NT, r = e4 / V4, x4 = (e4 % V4) * 4;
        double* dst = buf + (r * WS + (x4 & 7)) * PITCH + (x4 >> 3);
        dst[0] = (double)q[k].x;
        dst[PITCH] = (double)q[k].y;
        dst[2 * PITCH] = (double)q[k].z;
        dst[3 * PITCH] = (double)q[k].w;
    };
    auto stage_scalar = [&](double* buf, int sx) {                                           // views that are not 16-byte aligned
        int c, yb, xb0;
        decode(sx, c, yb, xb0);
        const float* src = im + c * s_c + (long long)(yb * WS) * s_h + xb0 * WS;
        for (int e = tid; e < WS * NBLK * WS; e += NT) {
            const int r = e / (NBLK * WS), xx = e % (NBLK * WS);
            const float v = (xb0 * WS + xx < W) ? __ldg(src + r * s_h + xx) : 0.f;
            buf[(r * WS + (xx & 7)) * PITCH + (xx >> 3)] = (double)v;
        }
    };
    int sidx = blockIdx.x;
    const int G = gridDim.x;
    if (sidx < total) {                            // prologue: stripe 0 into buffer 0, stripe 1 into the registers
        if (vec) {
            prefetch(sidx);
#pragma unroll
            for (int k = 0; k < NQ; ++k) stage(s_loc, k);
            if (sidx + G < total) prefetch(sidx + G);
        } else {
            stage_scalar(s_loc, sidx);
        }
    }
    for (int it = 0; sidx < total; sidx += G, it ^= 1) {
        int c, yb, xb0;
        decode(sidx, c, yb, xb0);
        const double* cur = s_loc + it * (NV * PITCH);
        double* nxt = s_loc + (it ^ 1) * (NV * PITCH);
        __syncthreads();                           // `cur` is complete; everybody is done reading `nxt` (the stripe before)
        const bool more = sidx + G < total;
        double acc[MT][2][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        const double* ap = cur + tg * PITCH + wrp * (MT * 8) + g;                  // A element (row = block g of the tile, k = tg)
#pragma unroll
        for (int ks = 0; ks < NV / 4; ++ks) {
            double a[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) a[mt] = ap[ks * 4 * PITCH + mt * 8];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) dmma884(acc[mt][nt], a[mt], bfr[ks][nt]);
            // the next stripe's conversion rides in the issue slots between the tensor-core instructions
            if (vec && more) {
#pragma unroll
                for (int k = ks * NQ / (NV / 4); k < (ks + 1) * NQ / (NV / 4); ++k) stage(nxt, k);
            }
        }
        if (!vec && more) stage_scalar(nxt, sidx + G);
        if (vec && sidx + 2 * G < total) prefetch(sidx + 2 * G);                   // lands during the epilogue and the next projection
        // accumulator element e of tile (mt, nt): block = 8 mt + g, component = 8 nt + 2 tg + e
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int kk = nt * 8 + tg * 2 + e;
                if (kk < ncomp) {
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const int xb = xb0 + wrp * (MT * 8) + mt * 8 + g;
                        if (xb < bx) {
                            const double v = (acc[mt][nt][e] - bias[nt][e]) * rmv[nt][e];              // 502, 510-511 (x * (1 / m): <= 1 ulp from x / m)
                            t[((long long)(c * ncomp + kk) * by + yb) * bx + xb] = v;                  // 516-518: [chan*ncomp, by, bx]
                            vmin = fmin(vmin, v);
                            vmax = fmax(vmax, v);
                        }
                    }
                }
            }
    }
    // global min / max (521-522): warp reduce on the order-preserving bit patterns, one atomic pair per warp
    unsigned long long hi = vmax >= vmin ? ordered_bits(vmax) : 0ull, lo = vmax >= vmin ? ~ordered_bits(vmin) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long h2 = __shfl_xor_sync(0xffffffffu, hi, o), l2 = __shfl_xor_sync(0xffffffffu, lo, o);
        hi = h2 > hi ? h2 : hi;
        lo = l2 > lo ? l2 : lo;
    }
    if (lane == 0) { atomicMax(&mm[0], hi); atomicMax(&mm[1], lo); }
}

template <typename OutT>
__global__ void __launch_bounds__(256) pca_rescale_kernel(const double* __restrict__ t, OutT* __restrict__ out,
                                                          const unsigned long long* __restrict__ mm, long long n) {
    const double ma = from_ordered_bits(mm[0]), mi = from_ordered_bits(~mm[1]);
    const double span = ma - mi;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (OutT)(((t[i] - mi) / span) * 2.0 - 1.0);                                    // pca_comp.py:523-526
}

}  // namespace fldr

using namespace fldr;

extern "C" size_t fldr_pca_features_workspace_bytes(int chan, int H, int W, int ncomp, int out_is_f32) {
    if (chan <= 0 || H <= 0 || W <= 0 || ncomp <= 0) return 0;
    size_t b = 256;                                                                // min / max words
    if (out_is_f32) b += align_up((size_t)chan * ncomp * (H / 8) * (W / 8) * sizeof(double), 256);   // float64 intermediate
    return b;
}

extern "C" int fldr_pca_features_fwd(const float* im, const int64_t* im_strides, const double* mean, const double* ev,
                                     int64_t ev_row_stride, const double* mean_vec, void* out, int out_is_f32, int chan, int H,
                                     int W, int ncomp, void* ws, size_t ws_bytes, fldr_stream_t stream) {
    if (!im || !im_strides || !mean || !ev || !out || chan <= 0 || H <= 0 || W <= 0 || ncomp <= 0) return FLDR_ERR_INVALID_ARGUMENT;
    if (H % 8 != 0 || W % 8 != 0) return FLDR_ERR_INVALID_ARGUMENT;                // pca_comp.py:486-487 raises
    if (ncomp > 16 || im_strides[2] != 1) return FLDR_ERR_UNSUPPORTED;
    if (!ws || ws_bytes < fldr_pca_features_workspace_bytes(chan, H, W, ncomp, out_is_f32)) return FLDR_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    unsigned long long* mm = static_cast<unsigned long long*>(ws);
    double* t = out_is_f32 ? reinterpret_cast<double*>(static_cast<char*>(ws) + 256) : static_cast<double*>(out);
    cudaError_t e = cudaMemsetAsync(mm, 0, 16, s);
    if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
    const int by = H / 8, bx = W / 8;
    const int chunks = (bx + pca::NBLK - 1) / pca::NBLK;
    const long long total = (long long)chunks * by * chan;
    if (total > 0x7fffffffLL) return FLDR_ERR_UNSUPPORTED;
    const size_t smem = (size_t)(2 * pca::NV * pca::PITCH) * sizeof(double);     // two stripe buffers, 68 KB: three CTAs per SM
    static bool attr_set[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !__atomic_load_n(&attr_set[dev], __ATOMIC_ACQUIRE)) {
        cudaError_t ea = cudaFuncSetAttribute(pca_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ea != cudaSuccess) { set_last_cuda_error(ea); return FLDR_ERR_CUDA; }
        __atomic_store_n(&attr_set[dev], true, __ATOMIC_RELEASE);
    }
    const int vec = (im_strides[0] % 4 == 0) && (im_strides[1] % 4 == 0) && ((reinterpret_cast<uintptr_t>(im) & 15) == 0);
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pca_project_kernel, pca::NT, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > total) grid = total;
    pca_project_kernel<<<(unsigned)grid, pca::NT, smem, s>>>(im, im_strides[0], im_strides[1], mean, ev, ev_row_stride, mean_vec, t, mm,
                                                             chan, by, bx, ncomp, vec, chunks, (int)total);
    int st = check_launch();
    if (st != FLDR_OK) return st;
    const long long n = (long long)chan * ncomp * by * bx;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    if (out_is_f32) pca_rescale_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(t, static_cast<float*>(out), mm, n);
    else pca_rescale_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(t, static_cast<double*>(out), mm, n);
    return check_launch();
}
