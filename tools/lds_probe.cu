// Shared-memory probe for the correlation kernel's lane mapping: how many cycles does a warp-wide LDS.128 cost when
// lanes in DIFFERENT quarter-warps read the same 16 bytes?  (Decides whether grouping the threads that share a row of
// `second` into one warp saves shared-memory bandwidth.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/lds_probe tools/lds_probe.cu && tools/lds_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ROWF = 40;   // floats per smem row, as corr.cu's second tile

// pattern: lane -> (pg = lane & 7, k = lane >> 3); row(k) chosen per mode
//  0: rows k          (4 distinct rows: today's mapping)
//  1: rows k >> 1     (2 distinct rows)
//  2: row 0 for all   (1 distinct row: full cross-quarter broadcast)
//  3: rows k, but 3 of 4 quarters share -> {0,0,0,1}
//  4: scalar-equivalent: every lane the same 16 B
template <int MODE>
__global__ void probe(float* out, int iters, long long* cyc) {
    __shared__ __align__(16) float tile[64 * ROWF];
    for (int i = threadIdx.x; i < 64 * ROWF; i += blockDim.x) tile[i] = (float)(i & 255) * 1e-3f;
    __syncthreads();
    const int lane = threadIdx.x & 31, pg = lane & 7, k = lane >> 3;
    int row;
    if (MODE == 0) row = k;
    else if (MODE == 1) row = k >> 1;
    else if (MODE == 2) row = 0;
    else if (MODE == 3) row = (k == 3);
    else row = 0;
    int off = row * ROWF + (MODE == 4 ? 0 : pg * 4);
    float4 acc = make_float4(0, 0, 0, 0);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            float4 v;
            const unsigned a = (unsigned)__cvta_generic_to_shared(&tile[off + (u * 4 * ROWF) % (48 * ROWF)]);
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <int MODE>
void run(const char* what, float* out, long long* cyc) {
    const int iters = 2000, warps = 16;
    probe<MODE><<<1, warps * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    probe<MODE><<<1, warps * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double n = (double)iters * 16 * warps;
    printf("%-52s %.2f cycles per warp-LDS.128 (one SM, %d warps)\n", what, (double)c / n, warps);
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 1 << 12);
    run<0>("4 distinct rows (one per quarter-warp)", out, cyc);
    run<1>("2 distinct rows (quarters 0,1 | 2,3 share)", out, cyc);
    run<2>("1 row (all four quarters share)", out, cyc);
    run<3>("rows {0,0,0,1}", out, cyc);
    run<4>("all 32 lanes the same 16 bytes", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
