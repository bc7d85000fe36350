// PWC-Net 9x9 cost volume (81 displacements, zero padding 4, mean over channels), forward + backward, sm_100a.
//
// Replaces OpticalFlow/correlation.py:17-242 (rearrange x2 + updateOutput; updateGradFirst / updateGradSecond
// launched once per sample).  No NHWC scratch copies: tiles of the NCHW inputs are staged in shared memory by TMA
// with their 4-pixel halo (the TMA unit's out-of-bounds zero fill IS the reference's zero padding), and every thread
// register-blocks 4 pixels x 3 dy x 9 dx (forward) or 8 channels x 4 pixels (backward).
//   corr81_fwd_tma_kernel   persistent CTAs, 3-stage mbarrier ring refilled by the last warp out of a chunk;
//                           channel-split + red.global.add.v4 for levels with fewer tiles than SMs
//   corr81_fwd_kernel       plain-load fallback (W % 4 != 0, unaligned or exotic views)
//   corr81_bwd_tile_kernel  both gradients in one launch (gradSecond gathers the shifted gradOut tile itself)
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "tma.cuh"

namespace fldr {

constexpr int kPad = 4;
constexpr int kD = 9;

// Forward epilogue: mean over channels, then the activation PWC-Net applies to every cost volume
// (leaky_relu(., 0.1), PWCNet.py:146-158; slope 1 = none), stored with an arbitrary sample stride so the volume can
// land directly in the first 81 channels of the decoder's concatenation buffer (PWCNet.py:160).
struct FwdEpilogue {
    float slope;            // negative_slope, 1 = identity
    long long out_sn;       // elements between consecutive samples of the output
};
__device__ __forceinline__ float corr_act(float v, float slope) { return v > 0.f ? v : v * slope; }

// ------------------------------------------------------------------------------------------------
// Forward.  CTA = 32 x 8 output pixels, 192 threads: thread = (4-pixel group, row, dy-group of 3).
// Per channel and thread: 1 LDS.128 (first) + 9 LDS.128 (second, 3 rows x 12 floats) feed 108 FFMA.
// ------------------------------------------------------------------------------------------------
namespace fwd {
constexpr int TW = 32, TH = 8, CK = 8, NT = 192;
constexpr int F2W = TW + 2 * kPad, F2H = TH + 2 * kPad;
}  // namespace fwd

__global__ void __launch_bounds__(fwd::NT) corr81_fwd_kernel(View4 f1, View4 f2, float* __restrict__ out, int C, int H,
                                                             int W, FwdEpilogue ep) {
    using namespace fwd;
    __shared__ __align__(16) float s1[CK][TH][TW];
    __shared__ __align__(16) float s2[CK][F2H][F2W];
    const int tid = threadIdx.x;
    const int pg = tid & 7, row = (tid >> 3) & 7, dyg = tid >> 6;
    const int x0t = blockIdx.x * TW, y0t = blockIdx.y * TH, b = blockIdx.z;
    const float* p1 = f1.p + b * f1.sn;
    const float* p2 = f2.p + b * f2.sn;

    float acc[3][kD][4];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 0; o < kD; ++o)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[d][o][j] = 0.f;

    for (int c0 = 0; c0 < C; c0 += CK) {
        __syncthreads();
        for (int e = tid; e < CK * TH * TW; e += NT) {
            const int c = e / (TH * TW), r = (e / TW) % TH, xx = e % TW;
            const int gy = y0t + r, gx = x0t + xx;
            float v = 0.f;
            if (c0 + c < C && gy < H && gx < W) v = __ldg(p1 + (c0 + c) * f1.sc + gy * f1.sh + gx * f1.sw);
            s1[c][r][xx] = v;
        }
        for (int e = tid; e < CK * F2H * F2W; e += NT) {
            const int c = e / (F2H * F2W), r = (e / F2W) % F2H, xx = e % F2W;
            const int gy = y0t + r - kPad, gx = x0t + xx - kPad;
            float v = 0.f;
            if (c0 + c < C && gy >= 0 && gy < H && gx >= 0 && gx < W)
                v = __ldg(p2 + (c0 + c) * f2.sc + gy * f2.sh + gx * f2.sw);
            s2[c][r][xx] = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int c = 0; c < CK; ++c) {
            const float4 a4 = *reinterpret_cast<const float4*>(&s1[c][row][pg * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float4* rp = reinterpret_cast<const float4*>(&s2[c][row + dyg * 3 + d][pg * 4]);
                const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2];
                const float f[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
#pragma unroll
                for (int o = 0; o < kD; ++o)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[d][o][j] = fmaf(a[j], f[j + o], acc[d][o][j]);
            }
        }
    }

    const int y = y0t + row, x = x0t + pg * 4;
    if (y >= H || x >= W) return;
    const float rc = 1.0f / (float)C;   // sum * (1/C): within 1 ulp of the reference's sum / C (exact for power-of-two C)
    const long long HW = (long long)H * W;
    float* ob = out + (long long)b * ep.out_sn + (long long)y * W + x;
    const bool vec = ((W & 3) == 0) && (x + 3 < W) && ((ep.out_sn & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 0; o < kD; ++o) {
            float* op = ob + (long long)((dyg * 3 + d) * kD + o) * HW;
            if (vec) {
                *reinterpret_cast<float4*>(op) =
                    make_float4(corr_act(acc[d][o][0] * rc, ep.slope), corr_act(acc[d][o][1] * rc, ep.slope),
                                corr_act(acc[d][o][2] * rc, ep.slope), corr_act(acc[d][o][3] * rc, ep.slope));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (x + j < W) op[j] = corr_act(acc[d][o][j] * rc, ep.slope);
            }
        }
}

// ------------------------------------------------------------------------------------------------
// Forward, TMA path (W % 4 == 0, 16-byte aligned rows): persistent CTAs, one per SM, loop over 32 x TH output
// tiles.  Per tile and chunk of CK channels two TMA boxes land in shared memory - first[CK][TH][32] and
// second[CK][TH+8][40] with its 4-pixel halo; out-of-frame rows/columns/channels are zero-filled by the TMA
// unit, which IS the reference's zero padding (correlation.py:297-298 + rearrange) at no cost.  A 3-stage
// full/empty mbarrier ring keeps the loads STAGES-1 chunks ahead of the FFMA loop, across tile boundaries.
// ------------------------------------------------------------------------------------------------
namespace fwdtma {
constexpr int TW = 32, CK = 8;
template <int TH> struct Cfg {
    static constexpr int NT = 8 * TH * 3;
    static constexpr int STAGES = TH == 16 ? 4 : 3;        // 4 x 46 KB (1 CTA/SM) or 3 x 28 KB (2 CTAs/SM)
    static constexpr int F2H = TH + 2 * kPad, F2W = TW + 2 * kPad;
    static constexpr int S1 = CK * TH * TW, S2 = CK * F2H * F2W;       // floats
    static constexpr int STAGE_BYTES = (S1 + S2) * 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8;
};
}  // namespace fwdtma

// SPLIT: the channel range of a tile is divided over `ksplit` work items (small pyramid levels have fewer tiles than
// the GPU has SMs: 20 tiles x 25 chunks at C = 196); partial sums are reduced into the zero-filled output with
// red.global.add.v4.f32.
// ACT: leaky-relu in the epilogue (never together with SPLIT).  STRIDED: output samples ep.out_sn apart instead of
// 81*H*W (a separate instantiation: the extra 64-bit value costs the dense path its last free registers)
template <int TH, bool SPLIT, bool ACT, bool STRIDED>
__global__ void __launch_bounds__(fwdtma::Cfg<TH>::NT, TH == 8 ? 2 : 1)
corr81_fwd_tma_kernel(const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2,
                      float* __restrict__ out, int B, int C, int H, int W, int tilesX, int tilesY, int ksplit,
                      int chunks_per_split, const __grid_constant__ FwdEpilogue ep) {
    using namespace fwdtma;
    using K = Cfg<TH>;
    constexpr int STAGES = K::STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * K::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    const int tid = threadIdx.x;
    const int pg = tid & 7, row = (tid >> 3) % TH, dyg = tid / (8 * TH);
    constexpr int NWARPS = K::NT / 32;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_barrier_init();
        tma_prefetch_desc(&tm1);
        tma_prefetch_desc(&tm2);
    }
    __syncthreads();

    const int nchunks_all = (C + CK - 1) / CK;
    const int nitems = tilesX * tilesY * B * ksplit;
    const int my_items = ((int)blockIdx.x < nitems) ? (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    // chunks of item `it`: [c_begin, c_end) ; every item of this CTA has the same split geometry except the last split
    auto item_chunks = [&](int it, int& tile, int& c_begin, int& c_end) {
        const int ks = SPLIT ? it % ksplit : 0;
        tile = SPLIT ? it / ksplit : it;
        c_begin = ks * chunks_per_split;
        c_end = min(nchunks_all, c_begin + chunks_per_split);
    };

    // Refill protocol: the LAST warp to finish a chunk refills the stage it just freed with the chunk STAGES ahead.
    // No warp ever waits for an "empty" signal and no fixed producer thread sits on the critical path; the cursor of
    // the next chunk to issue lives in shared memory (refills happen in chunk order: the last finisher of chunk g+1
    // cannot precede the last finisher of chunk g), so it costs the consumers no registers.
    __shared__ volatile int ps[8];   // [0] item (local index), [1] tx, [2] ty, [3] b, [4] chunk end, [5] next chunk
    __shared__ int done_cnt[STAGES];
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) done_cnt[s] = 0;
        int tile = 0, cb = 0, ce = 0;
        if (my_items > 0) item_chunks(blockIdx.x, tile, cb, ce);
        ps[0] = 0; ps[1] = tile % tilesX; ps[2] = (tile / tilesX) % tilesY; ps[3] = tile / (tilesX * tilesY);
        ps[4] = ce; ps[5] = cb;
    }
    auto issue_into = [&](int st) {     // one thread at a time
        const int il = ps[0];
        if (il >= my_items) return;
        const int tx = ps[1], ty = ps[2], b = ps[3], ch = ps[5];
        float* s1 = reinterpret_cast<float*>(smem_raw + st * K::STAGE_BYTES);
        float* s2 = s1 + K::S1;
        fence_proxy_async();            // the generic-proxy reads of this stage are ordered before the async-proxy refill
        mbar_arrive_expect_tx(&full[st], K::STAGE_BYTES);
        tma_load_4d(s1, &tm1, &full[st], tx * TW, ty * TH, ch * CK, b);
        tma_load_4d(s2, &tm2, &full[st], tx * TW - kPad, ty * TH - kPad, ch * CK, b);
        if (ch + 1 >= ps[4]) {
            ps[0] = il + 1;
            if (il + 1 < my_items) {
                int ntile, cb, ce;
                item_chunks(blockIdx.x + (il + 1) * gridDim.x, ntile, cb, ce);
                ps[1] = ntile % tilesX; ps[2] = (ntile / tilesX) % tilesY; ps[3] = ntile / (tilesX * tilesY);
                ps[4] = ce; ps[5] = cb;
            }
        } else {
            ps[5] = ch + 1;
        }
    };
    if (tid == 0)
        for (int i = 0; i < STAGES; ++i) issue_into(i);
    __syncthreads();

    const float rc = 1.0f / (float)C;
    const long long HW = (long long)H * W;
    int g = 0;
    for (int il = 0; il < my_items; ++il) {
        int tile, c_begin, c_end;
        item_chunks(blockIdx.x + il * gridDim.x, tile, c_begin, c_end);
        float acc[3][kD][4];
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int o = 0; o < kD; ++o)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[d][o][j] = 0.f;

        for (int ch = c_begin; ch < c_end; ++ch, ++g) {
            const int st = (int)(g % STAGES);
            mbar_wait(&full[st], (uint32_t)((g / STAGES) & 1));
            const float* s1 = reinterpret_cast<const float*>(smem_raw + st * K::STAGE_BYTES);
            const float* s2 = s1 + K::S1;
            const float* p1 = s1 + row * TW + pg * 4;
            const float* p2 = s2 + (row + dyg * 3) * K::F2W + pg * 4;
#pragma unroll 4
            for (int c = 0; c < CK; ++c) {
                const float4 a4 = *reinterpret_cast<const float4*>(p1 + c * (TH * TW));
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float4* rp = reinterpret_cast<const float4*>(p2 + c * (K::F2H * K::F2W) + d * K::F2W);
                    const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2];
                    const float f[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
#pragma unroll
                    for (int o = 0; o < kD; ++o)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[d][o][j] = fmaf(a[j], f[j + o], acc[d][o][j]);
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) {
                if (atomicAdd(&done_cnt[st], 1) == NWARPS - 1) {     // last warp out refills this stage
                    done_cnt[st] = 0;
                    __threadfence_block();
                    issue_into(st);
                }
            }
        }

        const int tx = tile % tilesX, ty = (tile / tilesX) % tilesY, b = tile / (tilesX * tilesY);
        const int y = ty * TH + row, x = tx * TW + pg * 4;
        if (y < H && x < W) {      // W % 4 == 0 on this path: the 4-pixel group is all in or all out
            float* ob = out + (long long)b * (STRIDED ? ep.out_sn : 81 * HW) + (long long)y * W + x;
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int o = 0; o < kD; ++o) {
                    float* op = ob + (long long)((dyg * 3 + d) * kD + o) * HW;
                    // SPLIT: partial sums are reduced in memory; the activation runs afterwards (corr81_act_kernel)
                    if (SPLIT) red_add_v4(op, acc[d][o][0] * rc, acc[d][o][1] * rc, acc[d][o][2] * rc, acc[d][o][3] * rc);
                    else if (ACT) __stcs(reinterpret_cast<float4*>(op),
                                make_float4(corr_act(acc[d][o][0] * rc, ep.slope), corr_act(acc[d][o][1] * rc, ep.slope),
                                            corr_act(acc[d][o][2] * rc, ep.slope), corr_act(acc[d][o][3] * rc, ep.slope)));
                    else __stcs(reinterpret_cast<float4*>(op),
                                make_float4(acc[d][o][0] * rc, acc[d][o][1] * rc, acc[d][o][2] * rc, acc[d][o][3] * rc));
                }
        }
    }
}

// in-place activation of a channel-split result (small pyramid levels only: a few hundred KB)
__global__ void __launch_bounds__(256) corr81_act_kernel(float* __restrict__ out, long long per_sample, long long out_sn,
                                                         float slope) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= per_sample) return;
    float* p = out + blockIdx.y * out_sn + i;
    *p = corr_act(*p, slope);
}

template <int TH>
static int launch_fwd_tma(const View4& f1, const View4& f2, float* out, int B, int C, int H, int W, cudaStream_t s,
                          bool* used, const FwdEpilogue& ep) {
    using namespace fwdtma;
    using K = Cfg<TH>;
    *used = false;
    CUtensorMap tm1, tm2;
    const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)B};
    const uint64_t st1[3] = {(uint64_t)f1.sh * 4, (uint64_t)f1.sc * 4, (uint64_t)f1.sn * 4};
    const uint64_t st2[3] = {(uint64_t)f2.sh * 4, (uint64_t)f2.sc * 4, (uint64_t)f2.sn * 4};
    const uint32_t box1[4] = {TW, TH, CK, 1};
    const uint32_t box2[4] = {(uint32_t)K::F2W, (uint32_t)K::F2H, CK, 1};
    if (!encode_tensor_map_4d(&tm1, f1.p, dims, st1, box1) || !encode_tensor_map_4d(&tm2, f2.p, dims, st2, box2))
        return FLDR_OK;   // not describable: caller falls back to the generic kernel
    const int tilesX = (W + TW - 1) / TW, tilesY = (H + TH - 1) / TH;
    const int ntiles = tilesX * tilesY * B;
    const int per_sm = (TH == 8) ? 2 : 1;
    const int slots = sm_count() * per_sm;
    const int nchunks = (C + CK - 1) / CK;
    // fewer tiles than CTA slots: split the channel range so every SM has work
    int ksplit = 1;
    if (ntiles * 2 <= slots && nchunks >= 2) {
        ksplit = slots / ntiles;
        if (ksplit > nchunks) ksplit = nchunks;
    }
    const int cps = (nchunks + ksplit - 1) / ksplit;
    ksplit = (nchunks + cps - 1) / cps;
    const int nitems = ntiles * ksplit;
    const int grid = nitems < slots ? nitems : slots;
    const bool strided = ep.out_sn != (long long)81 * H * W;
    cudaError_t e;
    if (ksplit > 1) {
        const size_t per_sample = (size_t)81 * H * W;
        e = (size_t)ep.out_sn == per_sample
                ? cudaMemsetAsync(out, 0, (size_t)B * per_sample * sizeof(float), s)
                : cudaMemset2DAsync(out, (size_t)ep.out_sn * sizeof(float), 0, per_sample * sizeof(float), (size_t)B, s);
        if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
        auto kern = strided ? corr81_fwd_tma_kernel<TH, true, false, true> : corr81_fwd_tma_kernel<TH, true, false, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES);
        if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
        kern<<<grid, K::NT, K::SMEM_BYTES, s>>>(tm1, tm2, out, B, C, H, W, tilesX, tilesY, ksplit, cps, ep);
        if (ep.slope != 1.0f) {
            dim3 ag((unsigned)((per_sample + 255) / 256), (unsigned)B);
            corr81_act_kernel<<<ag, 256, 0, s>>>(out, (long long)per_sample, ep.out_sn, ep.slope);
        }
    } else {
        const bool act = ep.slope != 1.0f;
        auto kern = act ? (strided ? corr81_fwd_tma_kernel<TH, false, true, true> : corr81_fwd_tma_kernel<TH, false, true, false>)
                        : (strided ? corr81_fwd_tma_kernel<TH, false, false, true> : corr81_fwd_tma_kernel<TH, false, false, false>);
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES);
        if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
        kern<<<grid, K::NT, K::SMEM_BYTES, s>>>(tm1, tm2, out, B, C, H, W, tilesX, tilesY, 1, nchunks, ep);
    }
    *used = true;
    return check_launch();
}

// ------------------------------------------------------------------------------------------------
// Backward.  Both gradients are the same tile computation
//     G[b,c,y,x] = (1/C) sum_{p,o} A[b,(p,o),y,x] * Fz[b,c,y+p,x+o]          (Fz = F zero-extended)
//   gradFirst : A = gradOut,  F = second                                        (correlation.py:151-169)
//   gradSecond: A = G2,       F = first,  with  G2[(q,r)][y][x] = gradOut[(-q,-r)][y+q][x+r]  (0 out of frame)
//               which is correlation.py:200-236 with the substitution q = -p, r = -o; G2 is plane 80-t of gradOut
//               shifted by its own displacement, built by one elementwise pass into the workspace.
// One batched launch per gradient (the reference: one launch per SAMPLE per gradient, correlation.py:365,385).
// CTA = 32 x 8 pixels x 32 channels per pass, 256 threads = (4-pixel group, row, 8-channel group):
//   the A tile [81][8][32] stays in shared memory for the whole tile, F is staged per 32-channel chunk with its halo;
//   per displacement row p a thread loads 9 LDS.128 of A (its 4 pixels, 9 dx) once and 3 LDS.128 of F per channel,
//   feeding 36 FFMA per channel into acc[8 channels][4 pixels].
// ------------------------------------------------------------------------------------------------
namespace bwd {
#ifndef FLDR_CORR_BWD_CK
#define FLDR_CORR_BWD_CK 16
#endif
constexpr int TW = 32, TH = 4, CK = FLDR_CORR_BWD_CK, NT = 128;      // CK = 16: 72 KB of shared memory, 3 CTAs per SM (measured 636 vs 734 us at CK = 32: the kernel is shared-memory-bandwidth bound, more warps hide more)
constexpr int CPT = CK / 4;                             // channels per thread
constexpr int FH = TH + 2 * kPad, FW = TW + 2 * kPad;
constexpr int A_FLOATS = 81 * TH * TW, F_FLOATS = CK * FH * FW;
constexpr int SMEM_BYTES = (A_FLOATS + F_FLOATS) * 4 + 16;          // 41472 + 61440 + mbarrier
}  // namespace bwd

// acc[8 channels][4 pixels] += sum over the 81 displacements, operands in shared memory
__device__ __forceinline__ void bwd_tile_compute(const float* __restrict__ sA, const float* __restrict__ sF, int pg, int row,
                                                 int cg, float (&acc)[bwd::CPT][4]) {
    using namespace bwd;
#pragma unroll 1
    for (int p = 0; p < kD; ++p) {
        float g[kD][4];
#pragma unroll
        for (int o = 0; o < kD; ++o) {
            const float4 g4 = *reinterpret_cast<const float4*>(sA + ((p * kD + o) * TH + row) * TW + pg * 4);
            g[o][0] = g4.x; g[o][1] = g4.y; g[o][2] = g4.z; g[o][3] = g4.w;
        }
#pragma unroll
        for (int cc = 0; cc < CPT; ++cc) {
            const float4* rp = reinterpret_cast<const float4*>(sF + ((cg * CPT + cc) * FH + row + p) * FW + pg * 4);
            const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2];
            const float f[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
#pragma unroll
            for (int o = 0; o < kD; ++o)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[cc][j] = fmaf(g[o][j], f[j + o], acc[cc][j]);
        }
    }
}

// One launch serves both gradients: blockIdx.z < nfirst * B works on gradFirst (A = gradOut as it is, F = second), the rest on
// gradSecond (F = first, A = gradOut read THROUGH the shift: A[(q,r)][y][x] = gradOut[(-q,-r)][y+q][x+r], zero outside the
// frame - correlation.py:200-236 with q = -p, r = -o).  The shifted tile is gathered by the CTA's threads straight from
// gradOut (81 coalesced row segments per thread position) while the TMA unit fetches the F chunk, so the 81-plane shifted copy
// the first version materialised in a workspace (one write + one read of 81*B*H*W floats, and a launch) is gone.
// TMA: the unshifted A tile and every F chunk arrive as bulk tensor loads (zero fill outside the frame / beyond C).
template <bool TMA>
__global__ void __launch_bounds__(bwd::NT, bwd::CK == 32 ? 2 : 3)
corr81_bwd_tile_kernel(const __grid_constant__ CUtensorMap tmF1, const __grid_constant__ CUtensorMap tmF2,
                       const __grid_constant__ CUtensorMap tmA, View4 F1, View4 F2, View4 A, float* __restrict__ G1,
                       float* __restrict__ G2, int B, int nfirst, int C, int H, int W, float rc) {
    using namespace bwd;
    extern __shared__ __align__(128) float smem_f[];
    float* sA = smem_f;                       // [81][TH][TW]
    float* sF = smem_f + A_FLOATS;            // [CK][FH][FW]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_f + A_FLOATS + F_FLOATS);
    const int tid = threadIdx.x;
    const int pg = tid & 7, row = (tid >> 3) & 3, cg = tid >> 5;
    const int x0t = blockIdx.x * TW, y0t = blockIdx.y * TH;
    const bool second = (int)blockIdx.z >= nfirst * B;            // which gradient this CTA works on
    const int b = second ? blockIdx.z - nfirst * B : blockIdx.z;
    const CUtensorMap* tmF = second ? &tmF2 : &tmF1;
    const View4& F = second ? F2 : F1;
    float* G = second ? G2 : G1;
    uint32_t phase = 0;
    if (TMA) {
        if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(bar, ((second ? 0 : A_FLOATS) + F_FLOATS) * 4);
            if (!second) tma_load_4d(sA, &tmA, bar, x0t, y0t, 0, b);
            tma_load_4d(sF, tmF, bar, x0t - kPad, y0t - kPad, 0, b);
        }
    }
    if (second) {
        // shifted gather: thread (r, xx) of the 4 x 32 tile fills its position of all 81 planes
        const int r = tid >> 5, xx = tid & 31;
        const float* pa = A.p + b * A.sn;
#pragma unroll 9
        for (int t = 0; t < 81; ++t) {
            const int q = t / kD - kPad, rr = t % kD - kPad;
            const int gy = y0t + r + q, gx = x0t + xx + rr;
            float v = 0.f;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(pa + (80 - t) * A.sc + gy * A.sh + gx * A.sw);
            sA[t * (TH * TW) + tid] = v;
        }
    } else if (!TMA) {
        const float* pa = A.p + b * A.sn;
        for (int e = tid; e < A_FLOATS; e += NT) {
            const int t = e / (TH * TW), r = (e / TW) % TH, xx = e % TW;
            const int gy = y0t + r, gx = x0t + xx;
            sA[e] = (gy < H && gx < W) ? __ldg(pa + t * A.sc + gy * A.sh + gx * A.sw) : 0.f;
        }
    }
    const int y = y0t + row, x = x0t + pg * 4;
    const long long HW = (long long)H * W;
    for (int c0 = 0; c0 < C; c0 += CK) {
        if (TMA) {
            __syncthreads();                           // c0 == 0: the gathered A tile is complete; later: the previous chunk was read
            if (c0 > 0 && tid == 0) {
                mbar_arrive_expect_tx(bar, F_FLOATS * 4);
                tma_load_4d(sF, tmF, bar, x0t - kPad, y0t - kPad, c0, b);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
        } else {
            const float* pf = F.p + b * F.sn;
            __syncthreads();
            for (int e = tid; e < F_FLOATS; e += NT) {
                const int c = e / (FH * FW), r = (e / FW) % FH, xx = e % FW;
                const int gy = y0t + r - kPad, gx = x0t + xx - kPad;
                float v = 0.f;
                if (c0 + c < C && gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(pf + (c0 + c) * F.sc + gy * F.sh + gx * F.sw);
                sF[e] = v;
            }
            __syncthreads();
        }
        float acc[CPT][4];
#pragma unroll
        for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[cc][j] = 0.f;
        bwd_tile_compute(sA, sF, pg, row, cg, acc);
        if (y < H && x < W) {
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) {
                const int c = c0 + cg * CPT + cc;
                if (c < C) {
                    float* gp = G + ((long long)b * C + c) * HW + (long long)y * W + x;
                    if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0) {
                        __stcs(reinterpret_cast<float4*>(gp), make_float4(acc[cc][0] * rc, acc[cc][1] * rc, acc[cc][2] * rc, acc[cc][3] * rc));
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (x + j < W) gp[j] = acc[cc][j] * rc;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward, TMA path: three output rows per thread.
// The tile kernel above moves 2.3 bytes of shared memory per FFMA (9 LDS.128 of A + 12 of F per 144 FFMA) and is bound by the
// shared-memory pipe, not by the FMA pipe.  Here a thread owns 3 rows x 4 pixels x 4 channels: an F row segment (12 floats per
// channel) loaded once serves the three output rows that see it under three different displacement rows p, so the same 12 + 27
// LDS.128 feed 432 FFMA (1.5 B / FFMA - the forward kernel's ratio).  The thread walks s = 0..10: F row 3*rg + s, output row i
// uses displacement row p = s - i.  A never sits in shared memory as a whole (81 planes x tile): its nine p-slices [9 o][TH][32]
// stream through a 4-slot TMA ring in the order the walk consumes them (slices s, s-1, s-2 are live at step s).
// gradSecond reads gradOut THROUGH the shift (correlation.py:200-236 with q = -p, r = -o): slice p is nine single-plane boxes,
// plane 80 - (9p + o) at offset (x + o - 4, y + p - 4); the TMA unit's zero fill outside the frame is the frame mask.
// ------------------------------------------------------------------------------------------------
namespace bwr {
constexpr int TW = 32, TH = 6, CK = 32, NT = 128, NRG = TH / 3;
constexpr int FW = TW + 2 * kPad, FROWS = TH + 2 * kPad;            // F rows a tile needs (14)
constexpr int FROW = CK * FW;                       // floats of one F row slot: [CK][40]
constexpr int NF = 7;                               // F-row ring: rows s .. s+3 are live at step s, three more are in flight
constexpr int AW2 = TW + 2 * kPad;                  // gradSecond: slice rows hold the aligned 40-column box around the tile
constexpr int PP2 = 256;                            // gradSecond: plane pitch inside a slice (6 x 40 floats padded to 1024 B: a TMA destination is 128-byte aligned)
constexpr int SLICE = kD * PP2;                     // floats reserved per p-slice (gradFirst uses 9 * TH * 32 of them)
constexpr int RING = 4;                             // A-slice ring: slices s-2 .. s are live at step s
constexpr int STEPS = kD + 2;
constexpr int SMEM_BYTES = (NF * FROW + RING * SLICE) * 4 + 128;   // 35 840 + 36 864 + barriers: three CTAs per SM
constexpr int CTAS_PER_SM = 3;
}  // namespace bwr

// four consecutive floats starting OFF floats after a 16-byte aligned shared-memory address (OFF known at compile time):
// one LDS.128 when aligned, LDS.64 x2 or LDS.32 + LDS.64 + LDS.32 otherwise - 4 wavefronts in every case
template <int OFF>
__device__ __forceinline__ void lds4_at(const float* base, float (&v)[4]) {
    const float* p = base + OFF;
    if (OFF % 4 == 0) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if (OFF % 2 == 0) {
        const float2 a = *reinterpret_cast<const float2*>(p), c = *reinterpret_cast<const float2*>(p + 2);
        v[0] = a.x; v[1] = a.y; v[2] = c.x; v[3] = c.y;
    } else {
        const float2 m = *reinterpret_cast<const float2*>(p + 1);
        v[0] = p[0]; v[1] = m.x; v[2] = m.y; v[3] = p[3];
    }
}

// one step of the walk: F row (3 rg + s) against the displacement rows p = s - i of output rows i = 0, 1, 2
// ROWS: which of the three output rows have a displacement row in range at this step (bit i); known at compile time so that a
// step is ONE basic block - the loads of a row's A values are scheduled above the previous row's FFMAs
template <bool SECOND, int ROWS>
__device__ __forceinline__ void bwd_rows_step(float (&acc)[4][3][4], const float* frow, const float* ring, int slice0, int s, int goff) {
    using namespace bwr;
    constexpr int AW = SECOND ? AW2 : TW;             // row pitch of a slice in shared memory
    constexpr int PP = SECOND ? PP2 : TH * TW;        // plane pitch
    float f[4][12];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        const float4* rp = reinterpret_cast<const float4*>(frow + cc * FW);
        const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2];
        f[cc][0] = r0.x; f[cc][1] = r0.y; f[cc][2] = r0.z; f[cc][3] = r0.w;
        f[cc][4] = r1.x; f[cc][5] = r1.y; f[cc][6] = r1.z; f[cc][7] = r1.w;
        f[cc][8] = r2.x; f[cc][9] = r2.y; f[cc][10] = r2.z; f[cc][11] = r2.w;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int p = s - i;
        if (ROWS >> i & 1) {
            const float* ga = ring + ((slice0 + p) & (RING - 1)) * SLICE + goff + i * AW;
            float g[kD][4];
            if (!SECOND) {
#pragma unroll
                for (int o = 0; o < kD; ++o) lds4_at<0>(ga + o * PP, g[o]);
            } else {
                lds4_at<0>(ga + 0 * PP, g[0]); lds4_at<1>(ga + 1 * PP, g[1]); lds4_at<2>(ga + 2 * PP, g[2]);
                lds4_at<3>(ga + 3 * PP, g[3]); lds4_at<4>(ga + 4 * PP, g[4]); lds4_at<5>(ga + 5 * PP, g[5]);
                lds4_at<6>(ga + 6 * PP, g[6]); lds4_at<7>(ga + 7 * PP, g[7]); lds4_at<8>(ga + 8 * PP, g[8]);
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                for (int o = 0; o < kD; ++o)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[cc][i][k] = fmaf(g[o][k], f[cc][k + o], acc[cc][i][k]);
        }
    }
}

// Persistent CTAs (three per SM) loop over (tile, gradient) items; both operands stream through shared-memory rings that run on
// across chunk and tile boundaries, so a CTA never sits through a load prologue:
//   A: the nine p-slices of gradOut, 4 slots (slices s-2 .. s are live at step s);
//   F: the fourteen halo rows of the CK-channel chunk, [CK][40] each, 7 slots (rows s .. s+3 are live at step s).
// Thread 0 refills after every step whatever slots the step just retired.
template <bool SECOND>
__global__ void __launch_bounds__(bwr::NT, bwr::CTAS_PER_SM)
corr81_bwd_rows_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmA, float* __restrict__ G, int B, int C,
                       int H, int W, int tilesX, int tilesY, float rc) {
    using namespace bwr;
    extern __shared__ __align__(128) float smem_f[];
    float* sF = smem_f;                               // [NF][CK][FW]
    float* ring = smem_f + NF * FROW;                 // [RING][9][TH][32 or 40 (+ pad)]
    uint64_t* fullA = reinterpret_cast<uint64_t*>(smem_f + NF * FROW + RING * SLICE);
    uint64_t* fullF = fullA + RING;
    const int tid = threadIdx.x;
    // a warp holds one row group: the unaligned 4- and 8-byte loads of gradSecond then touch 8 distinct, 16-byte spaced addresses
    // (conflict-free, the four channel groups of the warp broadcast)
    const int pg = tid & 7, w = tid >> 5, rg = w % NRG, cg = (w / NRG) * 4 + ((tid >> 3) & 3);
    const int nchunks = (C + CK - 1) / CK;
    const int nitems = tilesX * tilesY * B;
    const int my_items = ((int)blockIdx.x < nitems) ? (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int units = my_items * nchunks;             // (item, chunk) pairs of this CTA
    // item k of this CTA -> gradient, sample, tile origin
    auto decode = [&](int k, int& b, int& x0t, int& y0t) {
        int id = blockIdx.x + k * gridDim.x;
        b = id / (tilesX * tilesY);
        id -= b * (tilesX * tilesY);
        y0t = (id / tilesX) * TH;
        x0t = (id % tilesX) * TW;
    };
    if (tid == 0) {
        for (int i = 0; i < RING; ++i) mbar_init(&fullA[i], 1);
        for (int i = 0; i < NF; ++i) mbar_init(&fullF[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    // ---- producer state (shared memory, one writer at a time): cursors over the global slice index a = unit * 9 + p and the
    // global row index r = unit * 14 + fr.  There is no producer thread and no CTA barrier: the LAST warp to finish a step
    // refills what that step retired (the last finisher of step t+1 cannot precede the last finisher of step t, who refills
    // before it moves on, so refills happen in step order; a warp can run at most three steps ahead of the slowest - step t+4
    // needs the F row that the refill after step t issues - hence four arrival counters).
    struct Cursor { int idx, k, j, q, b, x0t, y0t; };       // q = p (slices) or fr (rows) inside the unit
    __shared__ Cursor curA, curF;
    __shared__ int cnt[4];
    const int total_a = units * kD, total_f = units * FROWS;
    auto cursor_item = [&](Cursor& c) {
        int b, x0t, y0t;
        decode(c.k, b, x0t, y0t);
        c.b = b; c.x0t = x0t; c.y0t = y0t;
    };
    auto issue_a = [&]() {                             // next slice
        Cursor c = curA;
        const int slot = c.idx & (RING - 1), p = c.q;
        float* dst = ring + slot * SLICE;
        if (!SECOND) {
            mbar_arrive_expect_tx(&fullA[slot], kD * TH * TW * 4);
            tma_load_4d(dst, &tmA, &fullA[slot], c.x0t, c.y0t, p * kD, c.b);
        } else {
            // plane 80 - (9p + o) shifted by (o - 4, p - 4): the box starts at the 16-byte aligned column x0t - 4 (a tensor-map
            // coordinate of the innermost dimension must be a multiple of 16 bytes); the x shift is applied when the slice is read
            mbar_arrive_expect_tx(&fullA[slot], kD * TH * AW2 * 4);
#pragma unroll
            for (int o = 0; o < kD; ++o)
                tma_load_4d(dst + o * PP2, &tmA, &fullA[slot], c.x0t - kPad, c.y0t + p - kPad, 80 - (p * kD + o), c.b);
        }
        ++c.idx;
        if (++c.q == kD) {
            c.q = 0;
            if (++c.j == nchunks) { c.j = 0; ++c.k; if (c.k < my_items) cursor_item(c); }
        }
        curA = c;
    };
    auto issue_f = [&]() {                             // next row
        Cursor c = curF;
        const int slot = c.idx % NF;
        mbar_arrive_expect_tx(&fullF[slot], FROW * 4);
        tma_load_4d(sF + slot * FROW, &tmF, &fullF[slot], c.x0t - kPad, c.y0t - kPad + c.q, c.j * CK, c.b);
        ++c.idx;
        if (++c.q == FROWS) {
            c.q = 0;
            if (++c.j == nchunks) { c.j = 0; ++c.k; if (c.k < my_items) cursor_item(c); }
        }
        curF = c;
    };
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) cnt[i] = 0;
        Cursor c = {0, 0, 0, 0, 0, 0, 0};
        if (my_items > 0) cursor_item(c);
        curA = c; curF = c;
        while (curA.idx < RING && curA.idx < total_a) issue_a();
        while (curF.idx < NF && curF.idx < total_f) issue_f();
    }
    __syncthreads();
    const long long HW = (long long)H * W;
    const int goff = (rg * 3) * (SECOND ? AW2 : TW) + pg * 4;
    int u = 0;
    for (int k = 0; k < my_items; ++k) {
        int b, x0t, y0t;
        decode(k, b, x0t, y0t);
        for (int j = 0; j < nchunks; ++j, ++u) {
            float acc[4][3][4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[cc][i][q] = 0.f;
            // one step: wait for what it consumes, compute, retire
            auto step = [&](int s, auto rows_tag) {
                constexpr int ROWS = decltype(rows_tag)::value;
                if (s < kD) {
                    const int a = u * kD + s;
                    mbar_wait(&fullA[a & (RING - 1)], (uint32_t)((a / RING) & 1));
                }
                {   // F rows up to s + 3 have landed (rows 0..3 before the first step)
                    const int r1 = u * FROWS + min(s + 3, FROWS - 1);
                    for (int r = (s == 0 ? u * FROWS : r1); r <= r1; ++r) mbar_wait(&fullF[r % NF], (uint32_t)((r / NF) & 1));
                }
                const float* frow = sF + ((u * FROWS + rg * 3 + s) % NF) * FROW + (cg * 4) * FW + pg * 4;
                bwd_rows_step<SECOND, ROWS>(acc, frow, ring, u * kD, s, goff);
                // this step's retirements: slice s - 2, F row s (everything after the last step); the last warp out refills
                __syncwarp();
                if ((tid & 31) == 0) {
                    const int t = u * STEPS + s;
                    if (atomicAdd(&cnt[t & 3], 1) == NT / 32 - 1) {
                        cnt[t & 3] = 0;
                        __threadfence_block();
                        const int dead_a = u * kD + min(kD, max(0, s - 1));
                        const int dead_f = u * FROWS + (s < STEPS - 1 ? s + 1 : FROWS);
                        fence_proxy_async();           // generic-proxy reads of the retired slots before the async-proxy refills
                        while (curA.idx < total_a && curA.idx - RING < dead_a) issue_a();
                        while (curF.idx < total_f && curF.idx - NF < dead_f) issue_f();
                    }
                }
                __syncwarp();
            };
            step(0, std::integral_constant<int, 1>());
            step(1, std::integral_constant<int, 3>());
#pragma unroll 1
            for (int s = 2; s < kD; ++s) step(s, std::integral_constant<int, 7>());
            step(kD, std::integral_constant<int, 6>());
            step(kD + 1, std::integral_constant<int, 4>());
            const int x = x0t + pg * 4;
            if (x < W) {                                // W % 4 == 0 on this path
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = j * CK + cg * 4 + cc;
                    if (c < C) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const int y = y0t + rg * 3 + i;
                            if (y < H)
                                __stcs(reinterpret_cast<float4*>(G + ((long long)b * C + c) * HW + (long long)y * W + x),
                                       make_float4(acc[cc][i][0] * rc, acc[cc][i][1] * rc, acc[cc][i][2] * rc, acc[cc][i][3] * rc));
                        }
                    }
                }
            }
        }
    }
}

static int check_corr_args(int B, int C, int H, int W) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return FLDR_ERR_INVALID_ARGUMENT;
    if ((long long)H * W * 81 >= (1ll << 31) || B > 65535) return FLDR_ERR_UNSUPPORTED;
    return FLDR_OK;
}

}  // namespace fldr

using namespace fldr;

extern "C" size_t fldr_corr81_fwd_workspace_bytes(int B, int C, int H, int W) {
    (void)B; (void)C; (void)H; (void)W;
    return 0;
}

static int corr81_fwd_impl(const float* first, const int64_t* first_strides, const float* second,
                           const int64_t* second_strides, float* out, int B, int C, int H, int W, const FwdEpilogue& ep,
                           fldr_stream_t stream) {
    int st = check_corr_args(B, C, H, W);
    if (st != FLDR_OK) return st;
    if (!first || !second || !first_strides || !second_strides || !out) return FLDR_ERR_INVALID_ARGUMENT;
    if (ep.out_sn < (long long)81 * H * W || !(ep.slope == ep.slope)) return FLDR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const View4 v1 = make_view(first, first_strides), v2 = make_view(second, second_strides);
    // TMA path: unit inner stride, rows 16-byte aligned, strides positive (views with stride 0 / negative fall back)
    const bool tma_ok = (W % 4 == 0) && v1.sw == 1 && v2.sw == 1 && v1.sh > 0 && v1.sc > 0 && v2.sh > 0 && v2.sc > 0 &&
                        (B == 1 || (v1.sn > 0 && v2.sn > 0));
    if (tma_ok) {
        bool used = false;
        const long long tiles16 = (long long)((W + 31) / 32) * ((H + 15) / 16) * B;
        const int th_opt = get_option(kOptCorrTh);      // tuning hook: 0 = automatic, 8 or 16 forces the tile height
        // measured (profiles/): two 8-row CTAs per SM overlap each other's epilogue / TMA-wait bubbles slightly better
        // than one 16-row CTA (180 vs 187 us on the C = 32 level), so 8 is the default
        const bool use16 = th_opt == 16;
        (void)tiles16;
        // float4 stores / reductions need 16-byte aligned samples
        if ((ep.out_sn & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {      // float4 stores / reductions: 16-byte aligned base and samples
            st = use16 ? launch_fwd_tma<16>(v1, v2, out, B, C, H, W, s, &used, ep) : launch_fwd_tma<8>(v1, v2, out, B, C, H, W, s, &used, ep);
            if (st != FLDR_OK || used) return st;
        }
    }
    dim3 grid((W + fwd::TW - 1) / fwd::TW, (H + fwd::TH - 1) / fwd::TH, B);
    corr81_fwd_kernel<<<grid, fwd::NT, 0, s>>>(v1, v2, out, C, H, W, ep);
    return check_launch();
}

extern "C" int fldr_corr81_fwd(const float* first, const int64_t* first_strides, const float* second,
                               const int64_t* second_strides, float* out, int B, int C, int H, int W, void* ws,
                               size_t ws_bytes, fldr_stream_t stream) {
    (void)ws; (void)ws_bytes;
    const FwdEpilogue ep = {1.0f, (long long)81 * H * W};
    return corr81_fwd_impl(first, first_strides, second, second_strides, out, B, C, H, W, ep, stream);
}

extern "C" int fldr_corr81_fwd_act(const float* first, const int64_t* first_strides, const float* second,
                                   const int64_t* second_strides, float* out, int64_t out_sample_stride,
                                   float negative_slope, int B, int C, int H, int W, fldr_stream_t stream) {
    const FwdEpilogue ep = {negative_slope, out_sample_stride > 0 ? (long long)out_sample_stride : (long long)81 * H * W};
    return corr81_fwd_impl(first, first_strides, second, second_strides, out, B, C, H, W, ep, stream);
}

extern "C" size_t fldr_corr81_bwd_workspace_bytes(int B, int C, int H, int W) {
    (void)B; (void)C; (void)H; (void)W;
    return 0;          // the shifted gradOut tile is gathered inside the kernel: no scratch
}

extern "C" int fldr_corr81_bwd(const float* first, const int64_t* first_strides, const float* second,
                               const int64_t* second_strides, const float* grad_out, const int64_t* grad_out_strides,
                               float* grad_first, float* grad_second, int B, int C, int H, int W, void* ws,
                               size_t ws_bytes, fldr_stream_t stream) {
    (void)ws; (void)ws_bytes;
    int st = check_corr_args(B, C, H, W);
    if (st != FLDR_OK) return st;
    if (!first || !second || !first_strides || !second_strides || !grad_out || !grad_out_strides)
        return FLDR_ERR_INVALID_ARGUMENT;
    if (!grad_first && !grad_second) return FLDR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const float rc = 1.0f / (float)C;
    const View4 vg = make_view(grad_out, grad_out_strides);
    const View4 v1 = make_view(first, first_strides), v2 = make_view(second, second_strides);
    const int nfirst = grad_first ? 1 : 0, nsecond = grad_second ? 1 : 0;
    if ((long long)B * (nfirst + nsecond) > 65535) return FLDR_ERR_UNSUPPORTED;
    dim3 grid((W + bwd::TW - 1) / bwd::TW, (H + bwd::TH - 1) / bwd::TH, B * (nfirst + nsecond));
    CUtensorMap tmF1, tmF2, tmA;
    const uint64_t dF[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)B};
    const uint64_t dA[4] = {(uint64_t)W, (uint64_t)H, 81, (uint64_t)B};
    const uint32_t bF[4] = {(uint32_t)bwd::FW, (uint32_t)bwd::FH, (uint32_t)bwd::CK, 1};
    const uint32_t bA[4] = {(uint32_t)bwd::TW, (uint32_t)bwd::TH, 81, 1};
    auto view_ok = [&](const View4& v) { return v.sw == 1 && v.sh > 0 && v.sc > 0 && (B == 1 || v.sn > 0); };
    auto enc = [&](CUtensorMap* m, const View4& v, const uint64_t* d, const uint32_t* bx) {
        const uint64_t sb[3] = {(uint64_t)v.sh * 4, (uint64_t)v.sc * 4, (uint64_t)(B == 1 && v.sn <= 0 ? (long long)d[2] * v.sc : v.sn) * 4};
        return encode_tensor_map_4d(m, v.p, d, sb, bx);
    };
    bool tma = (W % 4 == 0) && view_ok(v1) && view_ok(v2) && view_ok(vg);
    // F of gradFirst is `second`, F of gradSecond is `first`
    if (tma) tma = enc(&tmF1, v2, dF, bF) && enc(&tmF2, v1, dF, bF) && enc(&tmA, vg, dA, bA);
    if (!tma) { memset(&tmF1, 0, sizeof(tmF1)); memset(&tmF2, 0, sizeof(tmF2)); memset(&tmA, 0, sizeof(tmA)); }
    const bool g_ok = (!grad_first || (reinterpret_cast<uintptr_t>(grad_first) & 15) == 0) && (!grad_second || (reinterpret_cast<uintptr_t>(grad_second) & 15) == 0);
    // measured (profiles/r2_training_shapes_fwd_bwd.txt): the three-row kernel wins where one chunk holds all channels (C <= 32:
    // 526 vs 626 us at 64x32x128x128); with more chunks it re-streams gradOut per chunk and ties with the tile kernel (316 vs 304 us
    // at 64x64x64x64), which stays in charge there.  "corr_bwd_rows" = 2 forces it for every C.
    const int rows_opt = get_option(kOptCorrBwdRows);
    if (tma && g_ok && (rows_opt == 2 || (rows_opt == 1 && C <= bwr::CK))) {
        CUtensorMap tmR1, tmR2, tmA9, tmA1;
        const uint32_t bFr[4] = {(uint32_t)bwr::FW, 1, (uint32_t)bwr::CK, 1};
        const uint32_t bA9[4] = {(uint32_t)bwr::TW, (uint32_t)bwr::TH, 9, 1};
        const uint32_t bA1[4] = {(uint32_t)bwr::AW2, (uint32_t)bwr::TH, 1, 1};
        if (enc(&tmR1, v2, dF, bFr) && enc(&tmR2, v1, dF, bFr) && enc(&tmA9, vg, dA, bA9) && enc(&tmA1, vg, dA, bA1)) {
            static unsigned char attr_done[64];
            int dev = 0;
            cudaGetDevice(&dev);
            if (dev < 0 || dev >= 64) dev = 0;
            if (!__atomic_load_n(&attr_done[dev], __ATOMIC_ACQUIRE)) {
                cudaError_t e = cudaFuncSetAttribute(corr81_bwd_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwr::SMEM_BYTES);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(corr81_bwd_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwr::SMEM_BYTES);
                if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
                __atomic_store_n(&attr_done[dev], 1, __ATOMIC_RELEASE);
            }
            const int tilesX = (W + bwr::TW - 1) / bwr::TW, tilesY = (H + bwr::TH - 1) / bwr::TH;
            const long long nitems = (long long)tilesX * tilesY * B;
            if (nitems >= (1ll << 30)) return FLDR_ERR_UNSUPPORTED;
            const long long slots = (long long)bwr::CTAS_PER_SM * sm_count();
            const int gridp = (int)(nitems < slots ? nitems : slots);
            // one persistent launch per gradient (F of gradFirst is `second`, F of gradSecond is `first`)
            if (grad_first) corr81_bwd_rows_kernel<false><<<gridp, bwr::NT, bwr::SMEM_BYTES, s>>>(tmR1, tmA9, grad_first, B, C, H, W, tilesX, tilesY, rc);
            if (grad_second) corr81_bwd_rows_kernel<true><<<gridp, bwr::NT, bwr::SMEM_BYTES, s>>>(tmR2, tmA1, grad_second, B, C, H, W, tilesX, tilesY, rc);
            return check_launch();
        }
    }
    auto kern = tma ? corr81_bwd_tile_kernel<true> : corr81_bwd_tile_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::SMEM_BYTES);
    if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
    kern<<<grid, bwd::NT, bwd::SMEM_BYTES, s>>>(tmF1, tmF2, tmA, v2, v1, vg, grad_first, grad_second, B, nfirst, C, H, W, rc);
    return check_launch();
}
