// Block-PCA feature extraction: pca_comp.py:473-528 `to_pca_diff`, the first device step of
// every fLDRnet forward (fLDRnet.py:146).  The reference unfolds the two stacked frames into 8x8 blocks (nn.Unfold + four
// reshape / permute copies), subtracts the mean block, multiplies by the eigenvector matrix in float64 (cuBLAS DGEMM
// [blocks, 64] x [64, 16]), optionally divides by mean_vec, permutes to [chan * 16, H/8, W/8], takes a global min / max
// (two reductions) and rescales to [-1, 1]: ~12 kernels and five float64 temporaries of the frame's size.  Here:
//   pass 1  pca_project_kernel   reads the float32 frame once (coalesced 8-row stripes, converted and centred into shared
//                                memory), projects every block on the eigenvectors in float64 with a 4 x 4 register tile
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
constexpr int BG = 4, KG = 4;             // register blocking: a thread owns 4 adjacent blocks x 4 components
constexpr int NT = (NBLK / BG) * (16 / KG);      // 64 threads
constexpr int EVP = NV + 1;               // eigenvector row pitch in shared memory (doubles): conflict-free across components
constexpr int LP = NBLK * WS + 2 * (NBLK / BG) + 2;   // row pitch of the centred stripe (doubles), 2 doubles of skew per 4 blocks
__host__ __device__ constexpr int lcol(int xx) { return xx + ((xx >> 5) << 1); }   // block groups land 4 banks apart: conflict-free
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

// grid (ceil(bx / 64), by, chan).  Phase 1: the 8 x 512 float32 stripe is read once (coalesced), converted to float64 and
// centred (x - mean, pca_comp.py:502) into shared memory.  Phase 2: a thread accumulates a 4-block x 4-component register
// tile - 8 shared-memory loads feed 16 DFMA per pixel of the block, so the float64 pipe, not the LSU, is the limit.
__global__ void __launch_bounds__(pca::NT) pca_project_kernel(const float* __restrict__ im, long long s_c, long long s_h,
                                                              const double* __restrict__ mean, const double* __restrict__ ev,
                                                              long long ev_stride, const double* __restrict__ mean_vec,
                                                              double* __restrict__ t, unsigned long long* __restrict__ mm,
                                                              int chan, int by, int bx, int ncomp) {
    using namespace pca;
    extern __shared__ __align__(16) double smem_d[];
    double* s_loc = smem_d;                       // [WS][LP]
    double* s_ev = s_loc + WS * LP;               // [16][EVP]
    double* s_mean = s_ev + 16 * EVP;             // [NV]
    const int tid = threadIdx.x;
    const int c = blockIdx.z, yb = blockIdx.y, xb0 = blockIdx.x * NBLK;
    const int W = bx * WS;
    for (int e = tid; e < 16 * NV; e += NT) s_ev[(e / NV) * EVP + (e % NV)] = (e / NV < ncomp) ? ev[(long long)(e / NV) * ev_stride + (e % NV)] : 0.0;
    if (tid < NV) s_mean[tid] = mean[tid];
    __syncthreads();
    const float* src = im + c * s_c + (long long)(yb * WS) * s_h + xb0 * WS;
    for (int e = tid; e < WS * NBLK * WS; e += NT) {
        const int r = e / (NBLK * WS), xx = e % (NBLK * WS);
        const float v = (xb0 * WS + xx < W) ? __ldg(src + r * s_h + xx) : 0.f;
        s_loc[r * LP + lcol(xx)] = (double)v - s_mean[r * WS + (xx & 7)];                           // pca_comp.py:502
    }
    __syncthreads();
    const int bg = tid / (16 / KG), kg = tid % (16 / KG);          // block group, component group
    double acc[BG][KG];
#pragma unroll
    for (int b = 0; b < BG; ++b)
#pragma unroll
        for (int k = 0; k < KG; ++k) acc[b][k] = 0.0;
#pragma unroll 4
    for (int j = 0; j < NV; ++j) {
        double e4[KG], l4[BG];
#pragma unroll
        for (int k = 0; k < KG; ++k) e4[k] = s_ev[(kg * KG + k) * EVP + j];
#pragma unroll
        for (int b = 0; b < BG; ++b) l4[b] = s_loc[(j >> 3) * LP + lcol((bg * BG + b) * WS) + (j & 7)];
#pragma unroll
        for (int b = 0; b < BG; ++b)
#pragma unroll
            for (int k = 0; k < KG; ++k) acc[b][k] = fma(l4[b], e4[k], acc[b][k]);            // 507, j ascending
    }
    unsigned long long hi = 0ull, lo = 0ull;
#pragma unroll
    for (int k = 0; k < KG; ++k) {
        const int kk = kg * KG + k;
        if (kk < ncomp) {
            const double mv = mean_vec ? mean_vec[kk] : 1.0;
#pragma unroll
            for (int b = 0; b < BG; ++b) {
                const int xb = xb0 + bg * BG + b;
                if (xb < bx) {
                    const double v = mean_vec ? acc[b][k] / mv : acc[b][k];                    // 510-511
                    t[((long long)(c * ncomp + kk) * by + yb) * bx + xb] = v;                  // 516-518: [chan*ncomp, by, bx]
                    const unsigned long long o = ordered_bits(v);
                    hi = o > hi ? o : hi;
                    lo = ~o > lo ? ~o : lo;
                }
            }
        }
    }
    // global min / max (521-522): warp reduce on the order-preserving bit patterns, one atomic pair per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long h2 = __shfl_xor_sync(0xffffffffu, hi, o), l2 = __shfl_xor_sync(0xffffffffu, lo, o);
        hi = h2 > hi ? h2 : hi;
        lo = l2 > lo ? l2 : lo;
    }
    if ((tid & 31) == 0) { atomicMax(&mm[0], hi); atomicMax(&mm[1], lo); }
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
    if (chan > 65535 || H / 8 > 65535) return FLDR_ERR_UNSUPPORTED;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    unsigned long long* mm = static_cast<unsigned long long*>(ws);
    double* t = out_is_f32 ? reinterpret_cast<double*>(static_cast<char*>(ws) + 256) : static_cast<double*>(out);
    cudaError_t e = cudaMemsetAsync(mm, 0, 16, s);
    if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
    const int by = H / 8, bx = W / 8;
    dim3 grid((bx + pca::NBLK - 1) / pca::NBLK, by, chan);
    const size_t smem = (size_t)(pca::WS * pca::LP + 16 * pca::EVP + pca::NV) * sizeof(double);     // ~41 KB
    pca_project_kernel<<<grid, pca::NT, smem, s>>>(im, im_strides[0], im_strides[1], mean, ev, ev_row_stride, mean_vec, t, mm, chan,
                                                by, bx, ncomp);
    int st = check_launch();
    if (st != FLDR_OK) return st;
    const long long n = (long long)chan * ncomp * by * bx;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    if (out_is_f32) pca_rescale_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(t, static_cast<float*>(out), mm, n);
    else pca_rescale_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(t, static_cast<double*>(out), mm, n);
    return check_launch();
}
