// FP32 FMA rate of a register-only outer product (acc[i][j] += a[i] * b[j]): separates register-file limits from
// shared-memory effects in the correlation inner loop.
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, int NJ>
__global__ void __launch_bounds__(384, 1) outer(float* out, int iters, int nwarps) {
    if ((threadIdx.x >> 5) >= nwarps) return;
    float acc[NI][NJ], a[NI], b[NJ];
    for (int i = 0; i < NI; ++i) { a[i] = threadIdx.x * 1e-3f + i; for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f; }
    for (int j = 0; j < NJ; ++j) b[j] = blockIdx.x * 1e-3f + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        // perturb operands so nothing is hoisted (cheap: NI + NJ adds per NI*NJ fmas)
#pragma unroll
        for (int i = 0; i < NI; ++i) a[i] += 1e-7f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) b[j] -= 1e-7f;
    }
    float r = 0.f;
    for (int i = 0; i < NI; ++i) for (int j = 0; j < NJ; ++j) r += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int NI, int NJ> void run(int nwarps, float* out) {
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    outer<NI, NJ><<<148, 384>>>(out, 10, nwarps);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    outer<NI, NJ><<<148, 384>>>(out, iters, nwarps);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("outer %2dx%2d  %2d warps/SM: %6.1f TFLOP/s\n", NI, NJ, nwarps, 2.0 * 148 * nwarps * 32 * iters * NI * NJ / ms * 1e-9);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 384 * 4);
    for (int nw : {4, 8, 12}) run<8, 8>(nw, out);
    for (int nw : {4, 8, 12}) run<4, 12>(nw, out);
    for (int nw : {4, 8, 12}) run<12, 9>(nw, out);
    for (int nw : {4, 8, 12}) run<4, 27>(nw, out);
    return 0;
}
