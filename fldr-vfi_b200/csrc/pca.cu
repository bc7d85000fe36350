// Block-PCA feature extraction: pca_comp.py:473-528 `to_pca_diff`, the first device step of
// every fLDRnet forward (fLDRnet.py:146).  The reference unfolds the two stacked frames into 8x8 blocks (nn.Unfold + four
// reshape / permute copies), subtracts the mean block, multiplies by the eigenvector matrix in float64 (cuBLAS DGEMM
// [blocks, 64] x [64, 16]), optionally divides by mean_vec, permutes to [chan * 16, H/8, W/8], takes a global min / max
// (two reductions) and rescales to [-1, 1]: ~12 kernels and five float64 temporaries of the frame's size.  Here:
//   pass 1  pca_project_kernel   reads the float32 frame once (coalesced 8-row stripes staged in shared memory), projects
//                                every block on the eigenvectors in float64 in the order sum_j (x_j - mean_j) * EV[k][j],
//                                writes the result already permuted, and folds the global min / max into two 64-bit
//                                atomics (order-preserving bit pattern)
//   pass 2  pca_rescale_kernel   ((t - min) / (max - min)) * 2 - 1, float64 like the reference or straight to the float32
//                                the caller converts to (`.float()`, fLDRnet.py:146)
// Algorithmic bytes: 4 * chan * H * W read + chan * 16 * (H/8) * (W/8) * (8 | 4) written.  Bound: HBM.
#include "common.cuh"

namespace fldr {
namespace pca {
constexpr int WS = 8, NV = WS * WS, NBLK = 16, NT = 256;      // 16 blocks (128 pixels of a stripe) x 16 components per CTA
constexpr int EVP = NV + 1;                                   // eigenvector row pitch in shared memory: no bank conflicts
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

// grid (ceil(bx / 16), by, chan); thread = (block of the stripe, component)
__global__ void __launch_bounds__(pca::NT) pca_project_kernel(const float* __restrict__ im, long long s_c, long long s_h,
                                                              const double* __restrict__ mean, const double* __restrict__ ev,
                                                              long long ev_stride, const double* __restrict__ mean_vec,
                                                              double* __restrict__ t, unsigned long long* __restrict__ mm,
                                                              int chan, int by, int bx, int ncomp) {
    using namespace pca;
    __shared__ float tile[WS][NBLK * WS];
    __shared__ double s_ev[16 * EVP];
    __shared__ double s_mean[NV];
    __shared__ unsigned long long s_mm[2];
    const int tid = threadIdx.x;
    const int c = blockIdx.z, yb = blockIdx.y, xb0 = blockIdx.x * NBLK;
    const int W = bx * WS;
    for (int e = tid; e < ncomp * NV; e += NT) s_ev[(e / NV) * EVP + (e % NV)] = ev[(long long)(e / NV) * ev_stride + (e % NV)];
    if (tid < NV) s_mean[tid] = mean[tid];
    if (tid < 2) s_mm[tid] = 0ull;
    const float* src = im + c * s_c + (long long)(yb * WS) * s_h + xb0 * WS;
    for (int e = tid; e < WS * NBLK * WS; e += NT) {
        const int r = e / (NBLK * WS), xx = e % (NBLK * WS);
        tile[r][xx] = (xb0 * WS + xx < W) ? __ldg(src + r * s_h + xx) : 0.f;
    }
    __syncthreads();
    const int blk = tid / 16, k = tid % 16;
    const int xb = xb0 + blk;
    double acc = 0.0;
    bool live = xb < bx && k < ncomp;
    if (live) {
#pragma unroll 8
        for (int j = 0; j < NV; ++j) {
            const double loc = (double)tile[j >> 3][blk * WS + (j & 7)] - s_mean[j];          // pca_comp.py:502
            acc = fma(loc, s_ev[k * EVP + j], acc);                                           // 507
        }
        if (mean_vec) acc = acc / mean_vec[k];                                                // 510-511
        t[((long long)(c * ncomp + k) * by + yb) * bx + xb] = acc;                            // 516-518: [chan*ncomp, by, bx]
    }
    // global min / max (521-522): warp reduce on the order-preserving bit patterns, one atomic pair per CTA
    unsigned long long hi = live ? ordered_bits(acc) : 0ull, lo = live ? ~ordered_bits(acc) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long h2 = __shfl_xor_sync(0xffffffffu, hi, o), l2 = __shfl_xor_sync(0xffffffffu, lo, o);
        hi = h2 > hi ? h2 : hi;
        lo = l2 > lo ? l2 : lo;
    }
    if ((tid & 31) == 0) { atomicMax(&s_mm[0], hi); atomicMax(&s_mm[1], lo); }
    __syncthreads();
    if (tid == 0) { atomicMax(&mm[0], s_mm[0]); atomicMax(&mm[1], s_mm[1]); }
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
    pca_project_kernel<<<grid, pca::NT, 0, s>>>(im, im_strides[0], im_strides[1], mean, ev, ev_row_stride, mean_vec, t, mm, chan,
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
