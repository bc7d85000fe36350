// How fast does the L2 apply red.global.add.v4.f32?  Per 16-byte request or per 32-byte SECTOR touched by an instruction?
// A warp issues one v4 RED per lane into an L2-resident buffer (32 MB, re-walked) with three lane -> cell mappings:
//   dense   lane l -> cell base + l        (two requests per sector: 16 sectors per instruction)
//   half    lane l -> cell base + 2 l      (one request per sector: 32 sectors per instruction)
//   sparse  only every 4th lane active, cell base + 2 l   (8 requests, 8 sectors per instruction)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/red_sector_probe tools/red_sector_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void red4(float* a, float v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(a), "f"(v) : "memory");
}
__global__ void k(float4* acc, long long cells, int mode, int iters) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const int span = mode == 0 ? 32 : 64;
    for (int it = 0; it < iters; ++it) {
        const long long base = ((warp + (long long)it * nw) * span) % (cells - 64);
        const long long c = base + (mode == 0 ? lane : 2 * lane);
        if (mode != 2 || (lane & 3) == 0) red4(reinterpret_cast<float*>(acc + c), 1.0f);
    }
}
int main() {
    const long long cells = 2 << 20;   // 32 MB of float4 cells: L2 resident
    float4* acc;
    cudaMalloc(&acc, cells * 16);
    cudaMemset(acc, 0, cells * 16);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048;
    const char* names[3] = {"dense  (2 requests / sector)", "half   (1 request / sector) ", "sparse (8 of 32 lanes)      "};
    for (int mode = 0; mode < 3; ++mode) {
        k<<<sms * 8, 256>>>(acc, cells, mode, 16);
        cudaEventRecord(e0); k<<<sms * 8, 256>>>(acc, cells, mode, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double instr = (double)sms * 8 * 8 * iters;
        const double req = instr * (mode == 2 ? 8 : 32), sect = instr * (mode == 0 ? 16 : mode == 1 ? 32 : 8);
        printf("%s: %7.1f us  %6.1f G requests/s  %6.1f G sector updates/s  %5.2f G warp instructions/s\n", names[mode], ms * 1e3, req / ms / 1e6, sect / ms / 1e6, instr / ms / 1e6);
    }
    return 0;
}
