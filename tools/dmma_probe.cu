// float64 tensor-core (mma.sync.m8n8k4.f64) and DFMA issue-rate probe on one GPU: TFLOP/s against warps per SM and chains per warp.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_probe tools/dmma_probe.cu && tools/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void dmma_kernel(double* out, int iters) {
    double c[CH][2];
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = blockIdx.x * 1e-3 + 1.0;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double s = 0;
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void dfma_kernel(double* out, int iters) {
    double c[CH];
    for (int i = 0; i < CH; ++i) c[i] = i;
    double a = threadIdx.x * 1e-3, b = blockIdx.x * 1e-3 + 1.0;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) c[i] = fma(a, b, c[i]);
    double s = 0;
    for (int i = 0; i < CH; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    cudaMalloc(&out, 1 << 24);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    for (int warps = 4; warps <= 32; warps *= 2) {
        float ms;
        dmma_kernel<8><<<sms, warps * 32>>>(out, 16);
        cudaEventRecord(e0); dmma_kernel<8><<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA 8 chains, %2d warps/SM: %.2f TFLOP/s\n", warps, 512.0 * 8 * iters * warps * sms / ms / 1e9);
        dmma_kernel<2><<<sms, warps * 32>>>(out, 16);
        cudaEventRecord(e0); dmma_kernel<2><<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA 2 chains, %2d warps/SM: %.2f TFLOP/s\n", warps, 512.0 * 2 * iters * warps * sms / ms / 1e9);
        dfma_kernel<16><<<sms, warps * 32>>>(out, 16);
        cudaEventRecord(e0); dfma_kernel<16><<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA 16 chains, %2d warps/SM: %.2f TFLOP/s\n", warps, 64.0 * 16 * iters * warps * sms / ms / 1e9);
    }
    return 0;
}
