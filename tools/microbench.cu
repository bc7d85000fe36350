// Micro-benchmarks that drive the kernel design (run on the B200 box via gpurun; results -> profiles/).
//   red_*   : L2 reduction (REDG) throughput for the access patterns a splat produces
//   atoms_* : shared-memory atomic throughput for the privatised-tile alternative
//   ffma*   : FP32 FMA issue rate, scalar FFMA vs packed fma.rn.f32x2 (correlation inner loop)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void red_v4(float* a, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void red_s(float* a, float x) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a), "f"(x) : "memory"); }

__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// pattern 0: identity-like (target = source + const), 1: random within +-64 px, 2: fully random
template <int VEC>
__global__ void red_kernel(float* acc, int H, int W, int pattern, int planes) {
    const long long total = (long long)H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int x = i % W, y = i / W;
        int tx = x, ty = y;
        if (pattern == 1) { unsigned h = hash((unsigned)i); tx = x + (int)(h & 127) - 64; ty = y + (int)((h >> 8) & 127) - 64; }
        if (pattern == 2) { unsigned h = hash((unsigned)i); tx = h % W; ty = (h >> 12) % H; }
        if (tx < 0 || tx + 1 >= W || ty < 0 || ty + 1 >= H) continue;
        float v = 0.25f;
        if (VEC == 4) {
            float* p = acc + ((long long)ty * W + tx) * 4;
            red_v4(p, v, v, v, v); red_v4(p + 4, v, v, v, v); red_v4(p + 4ll * W, v, v, v, v); red_v4(p + 4ll * W + 4, v, v, v, v);
        } else {
            for (int c = 0; c < planes; ++c) {
                float* p = acc + (long long)c * total + (long long)ty * W + tx;
                red_s(p, v); red_s(p + 1, v); red_s(p + W, v); red_s(p + W + 1, v);
            }
        }
    }
}

// smem privatised: each CTA owns a 64x16 tile (+1 halo) with 4 channels interleaved, identity-like scatter, 16 ATOMS per pixel,
// then flushes the tile with plain stores (upper bound for the privatised design without the flush REDs)
__global__ void atoms_kernel(float* out, int H, int W, int reps) {
    __shared__ float tile[17][65][4];
    const int tid = threadIdx.x;
    for (int r = 0; r < reps; ++r) {
        for (int e = tid; e < 17 * 65 * 4; e += blockDim.x) (&tile[0][0][0])[e] = 0.f;
        __syncthreads();
        for (int e = tid; e < 64 * 16; e += blockDim.x) {
            int x = e % 64, y = e / 64;
            for (int c = 0; c < 4; ++c) {
                atomicAdd(&tile[y][x][c], 0.25f); atomicAdd(&tile[y][x + 1][c], 0.25f);
                atomicAdd(&tile[y + 1][x][c], 0.25f); atomicAdd(&tile[y + 1][x + 1][c], 0.25f);
            }
        }
        __syncthreads();
        long long base = ((long long)blockIdx.x * reps + r) * 64 * 16 * 4;
        for (int e = tid; e < 64 * 16 * 4; e += blockDim.x) out[base % ((long long)H * W * 4 - 64 * 16 * 4) + e] = (&tile[0][0][0])[e];
        __syncthreads();
    }
}

__global__ void ffma_kernel(float* out, int iters) {
    float a[16], b = threadIdx.x * 1e-9f + 1.0f, c = 1e-7f + threadIdx.x * 1e-12f;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma2_kernel(float* out, int iters) {
    unsigned long long a[8], b, c;
    float bf = threadIdx.x * 1e-9f + 1.0f, cf = 1e-7f + threadIdx.x * 1e-12f;
    asm("mov.b64 %0, {%1,%1};" : "=l"(b) : "f"(bf));
    asm("mov.b64 %0, {%1,%1};" : "=l"(c) : "f"(cf));
#pragma unroll
    for (int i = 0; i < 8; ++i) { float v = i + threadIdx.x; asm("mov.b64 %0, {%1,%1};" : "=l"(a[i]) : "f"(v)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeit(F f, int warm = 2, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < warm; ++i) f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int H = 2304, W = 4096;
    float* acc; CK(cudaMalloc(&acc, (size_t)H * W * 4 * sizeof(float)));
    float* out; CK(cudaMalloc(&out, (size_t)H * W * 4 * sizeof(float)));
    const long long px = (long long)H * W;
    int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    printf("SMs %d\n", nsm);
    float t0 = timeit([&] { CK(cudaMemsetAsync(acc, 0, (size_t)px * 16)); });
    printf("memset 151MB: %.1f us  (%.0f GB/s)\n", t0 * 1e3, px * 16 / t0 / 1e6);
    for (int pattern = 0; pattern < 3; ++pattern) {
        float t4 = timeit([&] { red_kernel<4><<<nsm * 16, 256>>>(acc, H, W, pattern, 4); });
        float t1 = timeit([&] { red_kernel<1><<<nsm * 16, 256>>>(acc, H, W, pattern, 4); });
        printf("pattern %d (4K, 9.4M px): v4 REDs (4/px) %.1f us = %.2f G v4-RED/s | scalar REDs (16/px) %.1f us = %.2f G RED/s\n",
               pattern, t4 * 1e3, px * 4 / t4 / 1e6, t1 * 1e3, px * 16 / t1 / 1e6);
    }
    // L2-resident accumulator (256 rows band = 16.8 MB)
    for (int pattern = 0; pattern < 2; ++pattern) {
        float t4 = timeit([&] { for (int b = 0; b < 9; ++b) red_kernel<4><<<nsm * 16, 256>>>(acc, 256, W, pattern, 4); });
        printf("pattern %d (9 x 256-row band, L2 resident): v4 REDs %.1f us total\n", pattern, t4 * 1e3);
    }
    {
        int reps = (int)(px / (64 * 16) / (nsm * 4)) + 1;
        float t = timeit([&] { atoms_kernel<<<nsm * 4, 256>>>(out, H, W, reps); });
        printf("smem privatised 64x16 tiles, 16 ATOMS/px + flush stores, %lld px: %.1f us\n", (long long)reps * nsm * 4 * 64 * 16, t * 1e3);
    }
    {
        int iters = 4096; int blocks = nsm * 8, thr = 256;
        float t = timeit([&] { ffma_kernel<<<blocks, thr>>>(out, iters); });
        double fl = 2.0 * 16 * iters * (double)blocks * thr;
        printf("FFMA : %.1f us  %.1f TFLOP/s\n", t * 1e3, fl / t / 1e9);
        float t2 = timeit([&] { ffma2_kernel<<<blocks, thr>>>(out, iters); });
        printf("FFMA2: %.1f us  %.1f TFLOP/s\n", t2 * 1e3, fl / t2 / 1e9);
    }
    return 0;
}
