// FMA-pipe utilisation of candidate correlation inner loops in isolation (operands resident in shared memory, no TMA,
// no stores): separates code-generation / register-file limits from pipeline and memory effects.
//   classic      : the shipped mapping, thread = 4 px x 3 dy x 9 dx, 1 + 9 LDS.128 per 108 FFMA
//   row-sharing  : thread = 3 output rows x one row of `second` (dy_i = s - y_i), 3 + 3 LDS.128 per 108 FFMA,
//                  all scalar / even-dx products as fma.rn.f32x2 / both
// Result on B200 (profiles/r1_corr_fma_ceiling.txt): every variant saturates at ~49 TFLOP/s (66 % of 148 x 128 x 2 x
// 1.965 GHz) whatever the LDS count or the packing, a register-only outer product (tools/ffma_probe.cu) at 53-56.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/corr_inner_probe tools/corr_inner_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

namespace fwdrs { constexpr int TW = 32, TH = 12, CK = 8, F2H = 20, F2W = 40, S1 = CK * TH * TW, S2 = CK * F2H * F2W, STAGE_BYTES = (S1 + S2) * 4; }
constexpr int kD = 9;
struct RowAcc { unsigned long long e[5][2]; float s[4][4]; };
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }

template <int PART>   // 1: packed even-o products only, 2: scalar odd-o only, 3: everything scalar, 4: packed even + scalar odd
__device__ __forceinline__ void probe_row(RowAcc& acc, float (&sc)[5][4], const ulonglong2& a4, const ulonglong2 (&q)[3]) {
    const unsigned long long fe[6] = {q[0].x, q[0].y, q[1].x, q[1].y, q[2].x, q[2].y};
    const unsigned long long a2[2] = {a4.x, a4.y};
    float f[12], a[4];
    for (int k = 0; k < 6; ++k) unpack2(fe[k], f[2 * k], f[2 * k + 1]);
    unpack2(a4.x, a[0], a[1]);
    unpack2(a4.y, a[2], a[3]);
    if (PART == 1 || PART == 4) {
#pragma unroll
        for (int oe = 0; oe < 5; ++oe)
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) ffma2(acc.e[oe][jp], a2[jp], fe[jp + oe]);
    }
    if (PART == 2 || PART == 3 || PART == 4) {
#pragma unroll
        for (int oo = 0; oo < 4; ++oo)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc.s[oo][j] = fmaf(a[j], f[j + 2 * oo + 1], acc.s[oo][j]);
    }
    if (PART == 3) {
#pragma unroll
        for (int oe = 0; oe < 5; ++oe)
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[oe][j] = fmaf(a[j], f[j + 2 * oe], sc[oe][j]);
    }
}

template <int PART>
__global__ void __launch_bounds__(384, 1) inner_part(float* out, int iters, int nwarps_active) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s = reinterpret_cast<float*>(smem_raw);
    for (int i = threadIdx.x; i < fwdrs::S1 + fwdrs::S2; i += blockDim.x) s[i] = (float)((i * 37) & 255) * 1e-3f;
    __syncthreads();
    const int tid = threadIdx.x, pg = tid & 7, a = 3 * ((tid >> 3) & 3), w = tid >> 5;
    if (w >= nwarps_active) return;
    RowAcc acc[3];
    float sc[3][5][4];
    for (int k = 0; k < 3; ++k) {
        for (int oe = 0; oe < 5; ++oe) { acc[k].e[oe][0] = acc[k].e[oe][1] = 0ull; for (int j = 0; j < 4; ++j) sc[k][oe][j] = 0.f; }
        for (int oo = 0; oo < 4; ++oo) for (int j = 0; j < 4; ++j) acc[k].s[oo][j] = 0.f;
    }
    const float* p1 = s + a * fwdrs::TW + pg * 4;
    const float* p2 = s + fwdrs::S1 + (a + 2 + (w % 7)) * fwdrs::F2W + pg * 4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int c = 0; c < 8; ++c) {
            const ulonglong2* rp = reinterpret_cast<const ulonglong2*>(p2 + c * (fwdrs::F2H * fwdrs::F2W));
            const ulonglong2 q[3] = {rp[0], rp[1], rp[2]};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const ulonglong2 a4 = *reinterpret_cast<const ulonglong2*>(p1 + c * (fwdrs::TH * fwdrs::TW) + k * fwdrs::TW);
                probe_row<PART>(acc[k], sc[k], a4, q);
            }
        }
    }
    float r = 0.f;
    for (int k = 0; k < 3; ++k) {
        for (int oe = 0; oe < 5; ++oe) { float x, y; unpack2(acc[k].e[oe][0], x, y); r += x + y; unpack2(acc[k].e[oe][1], x, y); r += x + y; for (int j = 0; j < 4; ++j) r += sc[k][oe][j]; }
        for (int oo = 0; oo < 4; ++oo) for (int j = 0; j < 4; ++j) r += acc[k].s[oo][j];
    }
    out[blockIdx.x * blockDim.x + tid] = r;
}

template <int PART>
void run_part(const char* what, int nwarps, float* out, int fma_per_row) {
    const int iters = 2000, smem = fwdrs::STAGE_BYTES;
    cudaFuncSetAttribute(inner_part<PART>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    inner_part<PART><<<148, 384, smem>>>(out, 10, nwarps);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    inner_part<PART><<<148, 384, smem>>>(out, iters, nwarps);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 148.0 * nwarps * 32 * iters * 8 * 3 * fma_per_row;
    printf("%-34s %2d warps/SM: %7.3f ms  %6.1f TFLOP/s  (%s)\n", what, nwarps, ms, 2 * fma / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

__device__ unsigned long long g_clk[4];
template <int MODE>   // the shipped mapping: 1 row of first x 3 rows of second, scalar FFMA
__global__ void __launch_bounds__(384, 1) inner(float* out, int iters, int nwarps_active) {
    unsigned long long c0 = 0, t0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) { c0 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); }
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s = reinterpret_cast<float*>(smem_raw);
    for (int i = threadIdx.x; i < fwdrs::S1 + fwdrs::S2; i += blockDim.x) s[i] = (float)((i * 37) & 255) * 1e-3f;
    __syncthreads();
    const int tid = threadIdx.x, pg = tid & 7, w = tid >> 5;
    if (w >= nwarps_active) return;
    float r = 0.f;
    {
        float acc[3][kD][4];
        for (int d = 0; d < 3; ++d) for (int o = 0; o < kD; ++o) for (int j = 0; j < 4; ++j) acc[d][o][j] = 0.f;
        const int row = (tid >> 3) & 7, dyg = w % 3;
        const float* p1 = s + row * 32 + pg * 4;
        const float* p2 = s + fwdrs::S1 + (row + dyg * 3) * 40 + pg * 4;
        for (int it = 0; it < iters; ++it) {
#pragma unroll 4
            for (int c = 0; c < 8; ++c) {
                const float4 a4 = *reinterpret_cast<const float4*>(p1 + c * (12 * 32));
                const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float4* rp = reinterpret_cast<const float4*>(p2 + c * (20 * 40) + d * 40);
                    const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2];
                    const float f[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
#pragma unroll
                    for (int o = 0; o < kD; ++o)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[d][o][j] = fmaf(av[j], f[j + o], acc[d][o][j]);
                }
            }
        }
        for (int d = 0; d < 3; ++d) for (int o = 0; o < kD; ++o) for (int j = 0; j < 4; ++j) r += acc[d][o][j];
    }
    out[blockIdx.x * blockDim.x + tid] = r;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        g_clk[0] = clock64() - c0; g_clk[1] = t1 - t0;
    }
}

template <int MODE>
void run(const char* what, int nthreads, int nwarps, float* out) {
    const int iters = 20000, smem = fwdrs::STAGE_BYTES;
    cudaFuncSetAttribute(inner<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    inner<MODE><<<148, nthreads, smem>>>(out, 10, nwarps);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    inner<MODE><<<148, nthreads, smem>>>(out, iters, nwarps);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 148.0 * nwarps * 32 * iters * 8 * 108;          // thread-level FMAs
    unsigned long long clk[4]; cudaMemcpyFromSymbol(clk, g_clk, sizeof(clk));
    printf("%-34s %2d warps/SM: %7.3f ms  %6.1f TFLOP/s  SM clock under load %.0f MHz (%s)\n", what, nwarps, ms, 2 * fma / ms * 1e-9,
           (double)clk[0] / (double)clk[1] * 1e3, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* out; cudaMalloc(&out, 148 * 384 * 4);
    for (int nw : {4, 8, 11, 12}) run_part<4>("row-sharing FFMA2 even + FFMA odd", nw, out, 36);
    for (int nw : {4, 8, 12}) run<1>("classic scalar FFMA, 10 LDS/108", 384, nw, out);
    for (int nw : {4, 8, 12}) run_part<1>("packed even-o only (10 FFMA2/row)", nw, out, 20);
    for (int nw : {4, 8, 12}) run_part<2>("scalar odd-o only (16 FFMA/row)", nw, out, 16);
    for (int nw : {4, 8, 12}) run_part<3>("all scalar row-sharing (36 FFMA/row)", nw, out, 36);
    return 0;
}
