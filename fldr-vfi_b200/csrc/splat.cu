// Softmax / average / linear / summation forward splat and its backward, sm_100a.
//
// Replaces softSplat.py:12-158 (three CuPy string kernels) AND the ~12 torch elementwise kernels
// FunctionSoftsplat wraps around them (softSplat.py:320-352): pre-scale, exp, cat, zero-init,
// normaliser fix-up, divide, post-scale are all folded into the passes below.
//
// Data layout in HBM
//   inputs      NCHW fp32 with arbitrary element strides (callers pass views: fLDRnet.py:386,449)
//   accumulator [N][Q][H][W][4] fp32, Q = ceil((C + has_norm) / 4): every channel quad is a pixel-interleaved
//               float4 image, so one corner of one source pixel is ONE 16-byte red.global.add.v4.f32 and adjacent
//               lanes (adjacent x) reduce into adjacent slots (the reference issues 4 scalar REDs per element,
//               softSplat.py:39-50).  The normaliser rides in slot C % 4 of quad C / 4.
//   outputs     NCHW contiguous (what the reference allocates, softSplat.py:234).
//
// Paths (fldr_splat_fwd picks; DESIGN.md section 4.1 has the measurements behind each choice)
//   default        cudaMemsetAsync + splat_scatter_merged_kernel + splat_normalise_kernel
//   tiny frames    splat_fused_small_kernel: the three phases in one cooperative launch
//   opt-in         splat_scatter_za_kernel ("splat_za"): the scatter zeroes the accumulator ahead of itself
//   opt-in         splat_stream_kernel ("splat_stream"): one launch, L2-resident ring accumulator, dataflow counters
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

#ifndef FLDR_SCATTER_MIN_CTAS
#define FLDR_SCATTER_MIN_CTAS 8
#endif

namespace fldr {

struct SplatGeom {
    int N, C, H, W;
    int CA;        // accumulated channels: C (+1 when the mode carries a normaliser)
    int CP;        // CA rounded up to a multiple of 4
    int mode;      // fldr_splat_mode
    int has_metric;
};

__host__ __device__ inline bool mode_has_norm(int mode) {
    return mode == FLDR_SPLAT_AVERAGE || mode == FLDR_SPLAT_LINEAR || mode == FLDR_SPLAT_SOFTMAX;
}

// Target coordinate, NW corner and the four bilinear weights exactly as softSplat.py:23-38 forms them
// (integer corner converted back to float, then subtracted).  Returns false when no corner can be in frame
// or the coordinate is not finite (the reference device-asserts there; we skip the pixel).
struct Corners {
    float X, Y;
    int x0, y0;
    float w[4];       // NW, NE, SW, SE
    bool valid[4];
};

__device__ __forceinline__ bool make_corners(int x, int y, float u, float v, int W, int H, Corners& k) {
    k.X = (float)x + u;
    k.Y = (float)y + v;
    if (!(isfinite(k.X) && isfinite(k.Y))) return false;
    const float fx0 = floorf(k.X), fy0 = floorf(k.Y);
    if (fx0 < -1.f || fx0 >= (float)W || fy0 < -1.f || fy0 >= (float)H) return false;
    k.x0 = (int)fx0;
    k.y0 = (int)fy0;
    const float x1f = (float)(k.x0 + 1), y1f = (float)(k.y0 + 1);
    k.w[0] = (x1f - k.X) * (y1f - k.Y);
    k.w[1] = (k.X - fx0) * (y1f - k.Y);
    k.w[2] = (x1f - k.X) * (k.Y - fy0);
    k.w[3] = (k.X - fx0) * (k.Y - fy0);
    const bool xl = k.x0 >= 0, xr = k.x0 + 1 < W, yt = k.y0 >= 0, yb = k.y0 + 1 < H;
    k.valid[0] = xl && yt;
    k.valid[1] = xr && yt;
    k.valid[2] = xl && yb;
    k.valid[3] = xr && yb;
    return true;
}

// ------------------------------------------------------------------------------------------------
// Pass 1: scatter with merged reductions.
//
// Measured on B200 (profiles/r1_microbench_design.txt): the L2 reduction unit sustains ~333 G 16-byte REDs/s for
// coalesced targets (113 us for the 4 REDs/pixel of one 4K image), ~81 G/s for random targets, and shared-memory
// atomics are slower still (226 us) - so the lever is issuing FEWER reductions, not privatising them.
// A thread owns one column and walks R source rows; lanes of a warp own adjacent columns.  Wherever the flow is
// locally constant
//   * the NE/SE contribution of column x lands on the pixel that column x+1 hits with NW/SW -> handed to the right
//     lane by warp shuffle, and
//   * the SW/SE contribution of row y lands on the pixel that row y+1 hits with NW/NE       -> carried in registers
// so a source pixel costs ~1 red.global.add.v4.f32 instead of 4 (the reference issues 4 scalar REDs per ELEMENT,
// softSplat.py:39-50).  A target mismatch simply flushes the carried value as its own RED: arbitrary flow stays
// correct, it only merges less.  Inactive pixels (out of frame, non-finite) carry sentinel coordinates that never match
// and never pass the frame test, so no per-corner flags are kept.
// Accumulator layout: [N][Q][H][W][4] fp32, Q = ceil(CA/4): every channel quad is a pixel-interleaved float4 image,
// so lanes (adjacent x) reduce into adjacent 16-byte slots.
// ------------------------------------------------------------------------------------------------
constexpr int kSentinel = -(1 << 28);

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int PX> __device__ __forceinline__ void vstore(float* p, const float* t);
template <> __device__ __forceinline__ void vstore<1>(float* p, const float* t) { __stcs(p, t[0]); }
template <> __device__ __forceinline__ void vstore<4>(float* p, const float* t) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(t[0], t[1], t[2], t[3]));
}

// Accumulator addressing policies for scatter_rows: where target pixel (x, y) of the current (n, q) plane lives, and
// which target rows a source strip is allowed to reach (the streaming kernel's ring has a bounded reach).
struct PlaneAcc {          // whole-frame accumulator [H][W] float4
    float4* base;
    int W;
    __device__ __forceinline__ float4* at(int x, int y) const { return base + (y * W + x); }
    __device__ __forceinline__ bool reachable(int) const { return true; }
    // The whole-frame accumulator lives in DRAM; a reduction into a line that is not in L2 stalls the L2 reduction unit
    // on the fill.  Pulling the target lines of the rows this thread will reach next into L2 ahead of time turns those
    // fills into ordinary, well-pipelined reads.
    int H, pf_rows;     // pf_rows: how many accumulator rows ahead of the row being scattered are pulled into L2
    __device__ __forceinline__ int prefetch_rows() const { return pf_rows; }
    __device__ __forceinline__ void prefetch(int x, int y) const {
        if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (y * W + x)));
    }
};
struct BoundedPlaneAcc {   // whole-frame accumulator, but only rows [ylo, yhi) are known to be zeroed already (zero-ahead)
    float4* base;
    int W, ylo, yhi;
    __device__ __forceinline__ float4* at(int x, int y) const { return base + (y * W + x); }
    __device__ __forceinline__ bool reachable(int y0) const { return y0 >= ylo && y0 < yhi; }
    __device__ __forceinline__ int prefetch_rows() const { return 0; }
    __device__ __forceinline__ void prefetch(int, int) const {}
};
struct RingAcc {           // ring of RR rows (power of two): row y of sample n lives in slot row (row0 + y) & mask
    float4* base;          // plane of this quad: [RR][W] float4
    int W, row0, mask, ylo, yhi;
    __device__ __forceinline__ float4* at(int x, int y) const { return base + (((row0 + y) & mask) * W + x); }
    __device__ __forceinline__ bool reachable(int y0) const { return y0 >= ylo && y0 < yhi; }
    __device__ __forceinline__ int prefetch_rows() const { return 0; }
    __device__ __forceinline__ void prefetch(int, int) const {}       // the ring is L2-resident by construction
};

__device__ __forceinline__ void red4p(float4* d, const float* v, bool p) {
    // single-instruction body: ptxas predicates the REDG instead of branching around it
    if (p) red_add_v4(reinterpret_cast<float*>(d), v[0], v[1], v[2], v[3]);
}

// Walks `rows` source rows of column x starting at row yb for channel quad q of sample n.
// WKIND: 0 = weight 1 (no metric / average / summation / raw), 1 = exp(z) (softmax), 2 = z (linear)
// Per row and lane:
//   * the carried bottom-W slot of the previous row joins this row's top-W slot when both hit the same pixel
//     (vertical merge), else it is flushed as its own reduction;
//   * the E slots (top and bottom) always travel to the lane on the right (rotate shuffle: lane 0 gets lane 31's),
//     which adds them to its own W slots when they hit the same pixels (horizontal merge) and otherwise issues
//     them as reductions on the sender's behalf - so no lane keeps E state and nothing is shuffled back.
// Returns true when some pixel's target row was out of the accumulator's reach (streaming ring only).
template <int WKIND, bool PRE, int QS, class Acc>
__device__ __forceinline__ bool scatter_rows(const View4& in, const View4& flow, const View4& metric, const SplatGeom& g,
                                             int n, int q, int x, int yb, int rows, const Acc& acc) {
    const int lane = threadIdx.x & 31;
    const int src_lane = (lane + 31) & 31;
    const int W = g.W, H = g.H;
    const bool inb = x < W;
    const float* fu = flow.p + n * flow.sn + (long long)yb * flow.sh + (long long)x * flow.sw;
    const float* fv = fu + flow.sc;
    const float* zp = WKIND ? metric.p + n * metric.sn + (long long)yb * metric.sh + (long long)x * metric.sw : nullptr;
    const float* ip = in.p + n * in.sn + (long long)(q * 4) * in.sc + (long long)yb * in.sh + (long long)x * in.sw;
    // quad shape: real channels in this quad (<= 0: only the weight slot) and the slot that carries the weight, if any.
    // QS fixes them at compile time for the two shapes that matter: 1 = image (3 channels + weight), 2 = full quad.
    const int nch = QS == 1 ? 3 : QS == 2 ? 4 : min(4, g.C - q * 4);
    const int wslot = QS == 1 ? 3 : QS == 2 ? -1 : ((g.CA > g.C) ? g.C - q * 4 : -1);
    const float xf = (float)x;
    bool overflow = false;

    float pw[4] = {0.f, 0.f, 0.f, 0.f};
    int px = kSentinel, py = kSentinel;

    // software pipeline, two rows deep: the loads of rows r+1 and r+2 are in flight while row r is processed (the walk
    // is serial per thread; ncu showed ~40 % of all stall samples on the first use of a just-loaded row with one row)
    struct Row { float u, v, z, x[4]; };
    auto load = [&](Row& L) {
        L.u = 0.f; L.v = 0.f; L.z = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) L.x[j] = 0.f;
        if (inb) {
            // streaming (evict-first) loads: every input element is read exactly once and must not push the
            // accumulator lines out of L2
            L.u = __ldcs(fu);
            L.v = __ldcs(fv);
            if (WKIND) L.z = __ldcs(zp);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < nch) L.x[j] = __ldcs(ip + (long long)j * in.sc);
        }
        fu += flow.sh; fv += flow.sh; ip += in.sh;
        if (WKIND) zp += metric.sh;
    };
    Row slot0, slot1;
    if (rows > 0) load(slot0);
    if (rows > 1) load(slot1);

    auto process = [&](const Row& L, int r) {
        const float u = L.u, v = L.v, z = L.z;
        const float xv[4] = {L.x[0], L.x[1], L.x[2], L.x[3]};
        // softSplat.py:23-38
        const float X = xf + u, Y = (float)(yb + r) + v;
        const float fx0 = floorf(X), fy0 = floorf(Y);
        bool ok = inb && fx0 >= -1.f && fx0 < (float)W && fy0 >= -1.f && fy0 < (float)H;   // false for NaN / inf
        if (ok && !acc.reachable((int)fy0)) { ok = false; overflow = true; }
        const int x0 = ok ? (int)fx0 : kSentinel;
        const int y0 = ok ? (int)fy0 : kSentinel;
        if (ok && acc.prefetch_rows() > 0) {
            if (r == 0)
                for (int k = 0; k < acc.prefetch_rows(); ++k) acc.prefetch(x0, y0 + k);
            acc.prefetch(x0, y0 + acc.prefetch_rows());
        }
        const float ax = (fx0 + 1.f) - X, bx = X - fx0, ay = (fy0 + 1.f) - Y, by = Y - fy0;
        float m = 1.f;
        if (WKIND == 1) m = expf(z);          // accurate expf: the parity bar is 1e-5 relative
        if (WKIND == 2) m = z;
        // inactive pixels contribute clean zeros (their weights may be NaN / inf)
        const float wNW = ok ? ax * ay : 0.f, wNE = ok ? bx * ay : 0.f, wSW = ok ? ax * by : 0.f, wSE = ok ? bx * by : 0.f;
        float tW[4], tE[4], bW[4], bE[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = xv[j];
            if (PRE) a = (a + 1.f) * 0.5f;                      // softSplat.py:334
            a = (j < nch) ? a * m : (j == wslot ? m : 0.f);     // softSplat.py:328 / 338, weight slot, padding
            tW[j] = a * wNW; tE[j] = a * wNE; bW[j] = a * wSW; bE[j] = a * wSE;
        }

        // vertical merge of the carried W slot, or flush it
        const bool vmatch = ok && (x0 == px) && (y0 == py);
        red4p(acc.at(px, py), pw, !vmatch && (unsigned)py < (unsigned)H && (unsigned)px < (unsigned)W);
#pragma unroll
        for (int j = 0; j < 4; ++j) tW[j] = vmatch ? tW[j] + pw[j] : tW[j];

        // horizontal: receive the left lane's E slots (targets (rx, ry) and (rx, ry + 1))
        const int ex = ok ? x0 + 1 : kSentinel;
        const int rx = __shfl_sync(0xffffffffu, ex, src_lane);
        const int ry = __shfl_sync(0xffffffffu, y0, src_lane);
        float rt[4], rb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            rt[j] = __shfl_sync(0xffffffffu, tE[j], src_lane);
            rb[j] = __shfl_sync(0xffffffffu, bE[j], src_lane);
        }
        const bool take = (lane > 0) && ok && rx == x0 && ry == y0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tW[j] = take ? tW[j] + rt[j] : tW[j];
            bW[j] = take ? bW[j] + rb[j] : bW[j];
        }
        const bool rxin = !take && (unsigned)rx < (unsigned)W;
        red4p(acc.at(rx, ry), rt, rxin && (unsigned)ry < (unsigned)H);
        red4p(acc.at(rx, ry + 1), rb, rxin && (unsigned)(ry + 1) < (unsigned)H);

        // this row's top-W slot is final; the bottom-W slot is carried to the next row
        red4p(acc.at(x0, y0), tW, (unsigned)y0 < (unsigned)H && (unsigned)x0 < (unsigned)W);
        px = x0; py = ok ? y0 + 1 : kSentinel;
#pragma unroll
        for (int j = 0; j < 4; ++j) pw[j] = bW[j];
    };

    for (int r = 0; r < rows; r += 2) {
        const Row c0 = slot0;
        if (r + 2 < rows) load(slot0);
        process(c0, r);
        if (r + 1 < rows) {
            const Row c1 = slot1;
            if (r + 3 < rows) load(slot1);
            process(c1, r + 1);
        }
    }
    red4p(acc.at(px, py), pw, (unsigned)py < (unsigned)H && (unsigned)px < (unsigned)W);
    return overflow;
}

template <int WKIND, bool PRE, int QS>
__global__ void __launch_bounds__(128, FLDR_SCATTER_MIN_CTAS) splat_scatter_merged_kernel(View4 in, View4 flow, View4 metric,
                                                                   float* __restrict__ acc, SplatGeom g, int Q,
                                                                   const unsigned* __restrict__ guard, int pf_rows, int R) {
    if (guard && *guard == 0) return;      // fallback launch that turned out not to be needed
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int yb = blockIdx.y * R;
    const int q = blockIdx.z % Q, n = blockIdx.z / Q;
    PlaneAcc pa;
    pa.base = reinterpret_cast<float4*>(acc) + (long long)(n * Q + q) * g.H * g.W;
    pa.W = g.W;
    pa.H = g.H;
    pa.pf_rows = pf_rows;
    scatter_rows<WKIND, PRE, QS>(in, flow, metric, g, n, q, x, yb, min(R, g.H - yb), pa);
}

// ------------------------------------------------------------------------------------------------
// Pass 1 with zero-ahead (large frames).  The separate zero fill costs the accumulator two extra DRAM crossings: the
// zeros are written out, and every line is fetched back when its first reduction arrives.  Here the scatter kernel
// zeroes the accumulator itself, D strips AHEAD of the strip it scatters, so the reductions land on lines that are
// still dirty-zero in L2 and the accumulator crosses DRAM once (write-back) instead of three times.
//   * CTAs take tickets from an atomic counter (ticket order = strip order, column block fastest); the first tickets
//     zero strips [0, D), every later ticket (G, c) first zeroes tile c of strip G + D, publishes it
//     (__syncthreads, then thread 0: __threadfence + atomicAdd on zdone[G+D] - the cooperative-groups grid.sync
//     pattern), then scatters tile c of strip G.
//   * before scattering it waits (ld.acquire poll) until strips G-Rs-1 .. G+Rs+1 are completely zeroed; their zeroers
//     hold lower tickets, so they are running or done: no deadlock, no co-residency assumption.
//   * a source whose target row leaves that window sets ctrl[1]; the host-side sequence then re-does the call with
//     the plain whole-frame path (guarded launches that exit at once otherwise).  Rs strips = +-128 rows by default.
// ctrl (unsigned): [0] ticket counter, [1] overflow flag, [2 ..] zdone[global strip]
// ------------------------------------------------------------------------------------------------
template <int R, int WKIND, bool PRE, int QS>
__global__ void __launch_bounds__(128) splat_scatter_za_kernel(View4 in, View4 flow, View4 metric, float* __restrict__ acc,
                                                               SplatGeom g, int Q, unsigned* __restrict__ ctrl, int Tc,
                                                               int NSr, int D, int Rs) {
    __shared__ int s_ticket;
    const int tid = threadIdx.x;
    if (tid == 0) s_ticket = (int)atomicAdd(&ctrl[0], 1u);
    __syncthreads();
    int k = s_ticket;
    unsigned* zdone = ctrl + 2;
    const int W = g.W, H = g.H;
    const int total_strips = g.N * Q * NSr;
    const int nz = min(D, total_strips) * Tc;          // leading zero-only tickets
    float4* acc4 = reinterpret_cast<float4*>(acc);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    auto zero_tile = [&](int Gz, int c) {              // tile c of global strip Gz, then publish
        const int plane = Gz / NSr, sz = Gz % NSr;
        const int x = c * 128 + tid;
        if (x < W) {
            float4* p = acc4 + ((long long)plane * H + (long long)sz * R) * W + x;
            const int rows = min(R, H - sz * R);
#pragma unroll 4
            for (int r = 0; r < rows; ++r) p[(long long)r * W] = zero4;
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicAdd(&zdone[Gz], 1u); }
    };

    if (k < nz) { zero_tile(k / Tc, k % Tc); return; }
    k -= nz;
    const int G = k / Tc, c = k % Tc;
    if (G >= total_strips) return;
    if (G + D < total_strips) zero_tile(G + D, c);
    const int plane = G / NSr, sidx = G % NSr;
    const int q = plane % Q, n = plane / Q;
    const int slo = max(0, sidx - Rs - 1), shi = min(NSr - 1, sidx + Rs + 1);
    if (tid < 32) {
        for (;;) {
            bool ready = true;
            for (int ss = slo + tid; ss <= shi; ss += 32)
                if (ld_acquire_u32(&zdone[plane * NSr + ss]) < (unsigned)Tc) ready = false;
            if (__all_sync(0xffffffffu, ready)) break;
            __nanosleep(100);
        }
    }
    __syncthreads();
    BoundedPlaneAcc pa;
    pa.base = acc4 + (long long)plane * H * W;
    pa.W = W;
    pa.ylo = (slo == 0) ? -1 : slo * R;
    pa.yhi = (shi == NSr - 1) ? H : (shi + 1) * R - 1;       // y0 + 1 must stay inside strip shi
    const int yb = sidx * R;
    const bool ovf = scatter_rows<WKIND, PRE, QS>(in, flow, metric, g, n, q, c * 128 + tid, yb, min(R, H - yb), pa);
    if (__syncthreads_or(ovf) && tid == 0) atomicOr(&ctrl[1], 1u);
}

// ------------------------------------------------------------------------------------------------
// Pass 2: normalise + post-scale + quad-interleaved -> NCHW.  One thread per (PX consecutive pixels, channel quad):
// PX float4 loads of the accumulator (+ PX of the quad holding the normaliser), 4 channel-plane stores of PX floats.
//   softSplat.py:343-349: norm==0 -> 1, divide, (y - 0.5) * 2 (post-scale in every mode but RAW).
// ------------------------------------------------------------------------------------------------
template <int PX>
__global__ void __launch_bounds__(256) splat_normalise_kernel(const float* __restrict__ acc, float* __restrict__ out,
                                                              float* __restrict__ norm_out, SplatGeom g, int Q,
                                                              const unsigned* __restrict__ guard) {
    if (guard && *guard == 0) return;      // fallback launch that turned out not to be needed
    const long long HW = (long long)g.H * g.W;
    const long long pix = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * PX;
    if (pix >= HW) return;      // PX > 1 only when HW % PX == 0
    const int q = blockIdx.y % Q, n = blockIdx.y / Q;
    const bool has_norm = g.CA > g.C;
    const float4* accn = reinterpret_cast<const float4*>(acc) + (long long)n * Q * HW;
    float4 s4[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) s4[k] = __ldcs(accn + q * HW + pix + k);
    float d[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) d[k] = 1.f;
    if (has_norm) {
        const int qn = g.C >> 2, slot = g.C & 3;
        float nrm[PX];
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            const float4 n4 = (qn == q) ? s4[k] : __ldcs(accn + qn * HW + pix + k);
            nrm[k] = slot == 0 ? n4.x : slot == 1 ? n4.y : slot == 2 ? n4.z : n4.w;
            d[k] = (nrm[k] == 0.f) ? 1.f : nrm[k];
            // one IEEE reciprocal per pixel instead of one division per channel (a float division is ~12 instructions
            // and this pass is issue-limited as much as DRAM-limited); S * (1/norm) is within 2 ulp of the reference's
            // S / norm, far inside the summation-order noise of the accumulation itself
#ifndef FLDR_NORMALISE_TRUE_DIV
            d[k] = __frcp_rn(d[k]);
#else
            if (Q > 1) d[k] = __frcp_rn(d[k]);
#endif
        }
        if (norm_out && q == 0) vstore<PX>(norm_out + (long long)n * HW + pix, nrm);
    }
    float* op = out + (long long)n * g.C * HW + pix;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = q * 4 + j;
        if (c < g.C) {
            float yv[PX];
#pragma unroll
            for (int k = 0; k < PX; ++k) {
                const float sv = j == 0 ? s4[k].x : j == 1 ? s4[k].y : j == 2 ? s4[k].z : s4[k].w;
                if (g.mode == FLDR_SPLAT_RAW) yv[k] = sv;
                else if (!has_norm) yv[k] = (sv - 0.5f) * 2.f;
#ifndef FLDR_NORMALISE_TRUE_DIV
                else yv[k] = (sv * d[k] - 0.5f) * 2.f;
#else
                else yv[k] = ((Q > 1 ? sv * d[k] : sv / d[k]) - 0.5f) * 2.f;
#endif
            }
            vstore<PX>(op + (long long)c * HW, yv);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fused streaming forward: ONE launch does zero-init, scatter, normalise, post-scale and the NCHW store, and the
// accumulator never leaves L2.
//
// Why: with a whole-frame accumulator the 4K image splat moves 943 MB through DRAM for 340 MB of algorithmic
// traffic (memset 151 W, scatter 226 R + 151 R/W of accumulator lines, normalise 151 R + 113 W; ncu, profiles/).
// Here the accumulator is a ring of RR rows (a power of two, <= ~32 MB, L2-resident) and the frame streams through it.
//
// Work is cut into items, handed out in a fixed order by an atomic counter to however many CTAs are resident:
//   Z(slot, t, q)  zero ring strip `slot`                                   (all Z first)
//   S(J, t, q)     scatter source strip J (R = 16 rows x 128 columns, quad q) into the ring (scatter_rows above)
//   N(J, t, q)     read ring strip J, normalise + post-scale, store NCHW, zero the ring strip again
// ordered as  S(0) ... S(D2-1), then S(g) N(g-D2) interleaved, then the last N's  (D2 = Ds + lag strips).
// Dependencies are per-strip completion counters in global memory (release: __threadfence + atomicAdd after a CTA
// barrier; acquire: ld.acquire.gpu poll by the waiting CTA):
//   S(J) waits until every ring strip it may touch, J-Ds .. J+Ds, has been cleaned for its epoch
//   N(J) waits until S(J-Ds) .. S(J+Ds) are complete (all t, q)
// An item only ever waits for items EARLIER in the order, which are already owned by running CTAs, so the schedule
// cannot deadlock and needs no co-residency guarantee (no cooperative launch).
// Ds bounds the vertical reach: a source whose target row leaves J-Ds .. J+Ds sets ctrl[1] and the host-side
// sequence re-does the call with the whole-frame path (guarded launches that exit at once otherwise).  When the ring
// holds all N*NS strips (every splat of the pyramid except the two 4K image splats) the reach is unbounded.
// Ring reads use ld.global.cg (L2): L1 is not coherent with the reductions performed at L2.
// ------------------------------------------------------------------------------------------------
namespace stream {
constexpr int R = 8;         // rows per strip: 888 resident CTAs x 8 rows x 128 columns = 222 rows of a 4K frame in flight
constexpr int TWC = 128;     // columns per item = threads per CTA
constexpr int kCtasPerSm = 7;  // resident CTAs per SM the schedule is sized for (72 registers x 128 threads)
}  // namespace stream

struct StreamGeom {
    int NS;        // strips per sample = ceil(H / R)
    int NT;        // N * NS absolute strips
    int T;         // column tiles = ceil(W / 128)
    int Q;         // channel quads
    int RS;        // ring strips (RS * R rows = power of two)
    int Ds;        // reach in strips
    int D2;        // Ds + LAG
    int nZ, nA, nB, total;   // item-order bookkeeping (see decode below)
    int jstart;    // first strip of the trailing N-only region
    int vecN;      // N items may use 4-pixel vector loads / stores (W % 4 == 0, 16-byte aligned outputs)
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_cg4(const float4* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_cg4(float4* p, float4 v) {
    asm volatile("st.global.cg.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ctrl layout (unsigned): [0] next item, [1] overflow flag, [2 .. 2+NT) sdone[J], [2+NT .. +RS) clean[slot],
// then nread[J * T + t] (readers of the normaliser quad per strip tile, Q > 1 only)
template <int WKIND, bool PRE, int QS>
__global__ void __launch_bounds__(stream::TWC) splat_stream_kernel(View4 in, View4 flow, View4 metric,
                                                                   float4* __restrict__ ring, unsigned* __restrict__ ctrl,
                                                                   float* __restrict__ out, float* __restrict__ norm_out,
                                                                   SplatGeom g, StreamGeom sg) {
    using namespace stream;
    __shared__ int s_item;
    __shared__ int s_overflow;
    unsigned* sdone = ctrl + 2;
    unsigned* clean = ctrl + 2 + sg.NT;
    const int tid = threadIdx.x;
    const int TQ = sg.T * sg.Q;
    const int H = g.H, W = g.W;
    const long long HW = (long long)H * W;
    const int ring_rows = sg.RS * R;
    const bool has_norm = g.CA > g.C;

    for (;;) {
        if (tid == 0) { s_item = (int)atomicAdd(&ctrl[0], 1u); s_overflow = 0; }
        __syncthreads();
        int idx = s_item;
        if (idx >= sg.total) break;

        // ---- decode the item
        int type, J, tq;          // type 0 = Z (J = slot), 1 = S, 2 = N
        if (idx < sg.nZ) { type = 0; J = idx / TQ; tq = idx % TQ; }
        else {
            idx -= sg.nZ;
            if (idx < sg.nA) { type = 1; J = idx / TQ; tq = idx % TQ; }
            else {
                idx -= sg.nA;
                if (idx < sg.nB) {
                    const int a = sg.nA / TQ;
                    const int gidx = idx / (2 * TQ), r = idx % (2 * TQ);
                    if (r < TQ) { type = 1; J = a + gidx; tq = r; }
                    else { type = 2; J = a + gidx - sg.D2; tq = r - TQ; }
                } else {
                    idx -= sg.nB;
                    type = 2; J = sg.jstart + idx / TQ; tq = idx % TQ;
                }
            }
        }
        const int t = tq % sg.T, q = tq / sg.T;
        const int x = t * TWC + tid;

        if (type == 0) {
            // ---- Z: zero ring strip J (rows J*R .. J*R+R-1 of quad q), columns of tile t
            float4* rq = ring + (long long)q * ring_rows * W;
            if (x < W)
#pragma unroll 4
                for (int r = 0; r < R; ++r) st_cg4(rq + (long long)(J * R + r) * W + x, make_float4(0.f, 0.f, 0.f, 0.f));
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicAdd(&clean[J], 1u);
            continue;
        }

        const int n = J / sg.NS, j = J % sg.NS;          // sample, strip inside the sample
        const int jlo = max(0, j - sg.Ds), jhi = min(sg.NS - 1, j + sg.Ds);

        if (type == 1) {
            // ---- S: wait for the ring strips this source strip may touch to be clean for their epoch
            if (tid < 32) {
                for (;;) {
                    bool ready = true;
                    for (int jj = jlo + tid; jj <= jhi; jj += 32) {
                        const int JJ = n * sg.NS + jj;
                        const unsigned need = (unsigned)(JJ / sg.RS + 1) * (unsigned)TQ;
                        if (ld_acquire(&clean[JJ % sg.RS]) < need) ready = false;
                    }
                    if (__all_sync(0xffffffffu, ready)) break;
                    __nanosleep(200);
                }
            }
            __syncthreads();
            RingAcc ra;
            ra.base = ring + (long long)q * ring_rows * W;
            ra.W = W;
            ra.row0 = (n * sg.NS * R) & (ring_rows - 1);
            ra.mask = ring_rows - 1;
            ra.ylo = jlo * R;
            ra.yhi = (jhi + 1) * R - 1;      // y0 + 1 must stay inside strip jhi
            if (jhi == sg.NS - 1) ra.yhi = H;   // bottom strip: rows >= H are dropped by the frame test anyway
            if (jlo == 0) ra.ylo = -1;
            const int yb = j * R;
            const bool ovf = scatter_rows<WKIND, PRE, QS>(in, flow, metric, g, n, q, x, yb, min(R, H - yb), ra);
            if (ovf) s_overflow = 1;
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                if (s_overflow) atomicOr(&ctrl[1], 1u);
                atomicAdd(&sdone[J], 1u);
            }
            continue;
        }

        // ---- N: wait for every source strip that can reach strip J
        if (tid < 32) {
            for (;;) {
                bool ready = true;
                for (int jj = jlo + tid; jj <= jhi; jj += 32)
                    if (ld_acquire(&sdone[n * sg.NS + jj]) < (unsigned)TQ) ready = false;
                if (__all_sync(0xffffffffu, ready)) break;
                __nanosleep(200);
            }
        }
        __syncthreads();
        const int qn = g.C >> 2, slot = g.C & 3;          // quad / lane of the normaliser channel
        const int row0 = (n * sg.NS * R) & (ring_rows - 1);
        const int yb = j * R;
        const int rows = min(R, H - yb);
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        {
            float4* rq = ring + (long long)q * ring_rows * W;
            const float4* rn = ring + (long long)qn * ring_rows * W;
            float* op = out + (long long)n * g.C * HW;
            // every quad's item of this (strip, tile) reads the normaliser quad, so that quad is zeroed last (below)
            const bool zero_now = !has_norm || q != qn || sg.Q == 1;
            if (sg.vecN) {
                // 4 pixels per thread: 64 contiguous bytes of ring per thread, float4 stores per channel plane
                const int xg = t * TWC + (tid & 31) * 4;
                if (xg < W) {
#pragma unroll
                    for (int r = tid >> 5; r < rows; r += TWC / 32) {
                        const int y = yb + r;
                        const long long ro = (long long)((row0 + y) & (ring_rows - 1)) * W + xg;
                        float4 s4[4], n4[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) s4[k] = ld_cg4(rq + ro + k);
                        float d[4] = {1.f, 1.f, 1.f, 1.f};
                        if (has_norm) {
                            float nrm[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                n4[k] = (qn == q) ? s4[k] : ld_cg4(rn + ro + k);
                                nrm[k] = slot == 0 ? n4[k].x : slot == 1 ? n4[k].y : slot == 2 ? n4[k].z : n4[k].w;
                                d[k] = (nrm[k] == 0.f) ? 1.f : __frcp_rn(nrm[k]);
                            }
                            if (norm_out && q == 0) vstore<4>(norm_out + (long long)n * HW + (long long)y * W + xg, nrm);
                        }
                        if (zero_now)
#pragma unroll
                            for (int k = 0; k < 4; ++k) st_cg4(rq + ro + k, zero4);
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const int c = q * 4 + c4;
                            if (c < g.C) {
                                float yv[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const float sv = c4 == 0 ? s4[k].x : c4 == 1 ? s4[k].y : c4 == 2 ? s4[k].z : s4[k].w;
                                    if (g.mode == FLDR_SPLAT_RAW) yv[k] = sv;
                                    else if (!has_norm) yv[k] = (sv - 0.5f) * 2.f;
                                    else yv[k] = (sv * d[k] - 0.5f) * 2.f;
                                }
                                vstore<4>(op + (long long)c * HW + (long long)y * W + xg, yv);
                            }
                        }
                    }
                }
            } else if (x < W) {
#pragma unroll 4
                for (int r = 0; r < rows; ++r) {
                    const int y = yb + r;
                    const long long ro = (long long)((row0 + y) & (ring_rows - 1)) * W + x;
                    const float4 s4 = ld_cg4(rq + ro);
                    float d = 1.f;
                    if (has_norm) {
                        const float4 n4 = (qn == q) ? s4 : ld_cg4(rn + ro);
                        const float nrm = slot == 0 ? n4.x : slot == 1 ? n4.y : slot == 2 ? n4.z : n4.w;
                        if (norm_out && q == 0) __stcs(norm_out + (long long)n * HW + (long long)y * W + x, nrm);
                        d = (nrm == 0.f) ? 1.f : nrm;
                    }
                    if (zero_now) st_cg4(rq + ro, zero4);
                    const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const int c = q * 4 + c4;
                        if (c < g.C) {
                            float yv;
                            if (g.mode == FLDR_SPLAT_RAW) yv = sv[c4];
                            else if (!has_norm) yv = (sv[c4] - 0.5f) * 2.f;
                            else yv = (sv[c4] / d - 0.5f) * 2.f;
                            __stcs(op + (long long)c * HW + (long long)y * W + x, yv);
                        }
                    }
                }
            }
        }
        if (has_norm && sg.Q > 1) {
            // last reader of this (strip, tile) clears the normaliser quad's tile
            __shared__ int s_last;
            __syncthreads();
            if (tid == 0) {
                unsigned* nread = ctrl + 2 + sg.NT + sg.RS;
                const unsigned prev = atomicAdd(&nread[J * sg.T + t], 1u);
                s_last = (prev + 1u == (unsigned)sg.Q) ? 1 : 0;
            }
            __syncthreads();
            if (s_last && x < W) {
                float4* rn = ring + (long long)qn * ring_rows * W;
                for (int r = 0; r < rows; ++r) st_cg4(rn + (long long)((row0 + yb + r) & (ring_rows - 1)) * W + x, zero4);
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(&clean[J % sg.RS], 1u);
    }
}

// guarded zero fill for the whole-frame fallback (a cudaMemsetAsync cannot be made conditional on a device flag)
__global__ void __launch_bounds__(256) splat_zero_kernel(float4* __restrict__ p, long long n4, const unsigned* __restrict__ guard) {
    if (guard && *guard == 0) return;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------------
// Small frames: zero + scatter + normalise in ONE cooperative launch (two grid barriers instead of two kernel
// boundaries and a memset).  The pyramid's small splats (C = 48 at 144x256 ... 18x32) are launch-latency bound:
// three dependent launches cost ~20 us on an otherwise idle GPU, this costs ~8.
// Phase B hands (plane, row run, 32-column block) units to warps; scatter_rows only needs warp-level convergence.
// ------------------------------------------------------------------------------------------------
template <int R, int WKIND, bool PRE>
__global__ void __launch_bounds__(256) splat_fused_small_kernel(View4 in, View4 flow, View4 metric, float* __restrict__ acc,
                                                                float* __restrict__ out, float* __restrict__ norm_out,
                                                                SplatGeom g, int Q) {
    cg::grid_group grid = cg::this_grid();
    const long long HW = (long long)g.H * g.W;
    const long long n4 = (long long)g.N * Q * HW;
    const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long gthreads = (long long)gridDim.x * blockDim.x;
    float4* acc4 = reinterpret_cast<float4*>(acc);
    // ---- A: zero the accumulator
    for (long long i = gtid; i < n4; i += gthreads) acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    grid.sync();
    // ---- B: scatter, one unit per warp
    {
        const int runs = (g.H + R - 1) / R, cblocks = (g.W + 31) / 32;
        const long long units = (long long)g.N * Q * runs * cblocks;
        const int lane = threadIdx.x & 31;
        for (long long u = gtid >> 5; u < units; u += gthreads >> 5) {
            const int cb = (int)(u % cblocks);
            const int run = (int)((u / cblocks) % runs);
            const int nq = (int)(u / ((long long)cblocks * runs));
            const int q = nq % Q, n = nq / Q;
            PlaneAcc pa;
            pa.base = acc4 + (long long)nq * HW;
            pa.W = g.W;
            pa.H = g.H;
            pa.pf_rows = 0;      // small frames: the accumulator is L2-resident anyway
            const int yb = run * R;
            scatter_rows<WKIND, PRE, 0>(in, flow, metric, g, n, q, cb * 32 + lane, yb, min(R, g.H - yb), pa);
        }
    }
    grid.sync();
    // ---- C: normalise + post-scale + NCHW store (softSplat.py:343-349), one thread per (plane, pixel)
    {
        const bool has_norm = g.CA > g.C;
        const int qn = g.C >> 2, slot = g.C & 3;
        for (long long i = gtid; i < n4; i += gthreads) {
            const long long pix = i % HW;
            const int nq = (int)(i / HW);
            const int q = nq % Q, n = nq / Q;
            const float4 s4 = __ldcg(acc4 + i);
            float d = 1.f;
            if (has_norm) {
                const float4 n4v = (qn == q) ? s4 : __ldcg(acc4 + ((long long)n * Q + qn) * HW + pix);
                const float nrm = slot == 0 ? n4v.x : slot == 1 ? n4v.y : slot == 2 ? n4v.z : n4v.w;
                if (norm_out && q == 0) norm_out[(long long)n * HW + pix] = nrm;
                d = (nrm == 0.f) ? 1.f : nrm;
            }
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
            float* op = out + (long long)n * g.C * HW + pix;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = q * 4 + j;
                if (c < g.C) {
                    float yv;
                    if (g.mode == FLDR_SPLAT_RAW) yv = sv[j];
                    else if (!has_norm) yv = (sv[j] - 0.5f) * 2.f;
                    else yv = (sv[j] / d - 0.5f) * 2.f;
                    op[(long long)c * HW] = yv;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward.  One thread per source pixel; gathers grad_out / forward output / normaliser at its 4 corners,
// forms gS on the fly and emits grad_in, grad_flow, grad_metric in one pass:
//   gS_c = 2 gY_c / norm'            gS_C = -sum_c gS_c * (S_c / norm')   (0 where norm was 0)
//   gA   = sum_corners w * gS        (kernel_Softsplat_updateGradInput, softSplat.py:84-95)
//   gF   = sum_c A_c * sum_corners gS_c * dw   (kernel_Softsplat_updateGradFlow, 130-155)
//   softmax: g_x = gA_c * e^z / 2 ; g_z = e^z (sum_c gA_c x~_c + gA_C)      linear: g_x = gA_c z ; g_z = sum_c gA_c x_c + gA_C
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) splat_bwd_kernel(View4 in, View4 flow, View4 metric, const float* __restrict__ Yf,
                                                        const float* __restrict__ norm, View4 gout,
                                                        float* __restrict__ gin, float* __restrict__ gflow,
                                                        float* __restrict__ gmetric, SplatGeom g) {
    const long long HW = (long long)g.H * g.W;
    const long long total = HW * g.N;
    const bool pre = g.mode == FLDR_SPLAT_SOFTMAX;
    const bool has_norm = g.CA > g.C;
    const float gscale = (g.mode == FLDR_SPLAT_RAW) ? 1.f : 2.f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % g.W);
        const int y = (int)((idx / g.W) % g.H);
        const int n = (int)(idx / HW);
        const long long pix = (long long)y * g.W + x;
        const float* fp = flow.p + n * flow.sn + y * flow.sh + x * flow.sw;
        Corners k;
        const bool live = make_corners(x, y, __ldg(fp), __ldg(fp + flow.sc), g.W, g.H, k);
        float* ginp = gin ? gin + (long long)n * g.C * HW + pix : nullptr;
        if (!live) {
            if (ginp) for (int c = 0; c < g.C; ++c) ginp[(long long)c * HW] = 0.f;
            if (gflow) { gflow[(long long)n * 2 * HW + pix] = 0.f; gflow[(long long)n * 2 * HW + HW + pix] = 0.f; }
            if (gmetric) gmetric[(long long)n * HW + pix] = 0.f;
            continue;
        }
        float m = 1.f, z = 0.f;
        if (g.has_metric) {
            z = __ldg(metric.p + n * metric.sn + y * metric.sh + x * metric.sw);
            m = (g.mode == FLDR_SPLAT_SOFTMAX) ? expf(z) : (g.mode == FLDR_SPLAT_LINEAR ? z : 1.f);
        }
        long long cpix[4];      // corner pixel index inside one plane
        cpix[0] = (long long)k.y0 * g.W + k.x0;
        cpix[1] = cpix[0] + 1;
        cpix[2] = cpix[0] + g.W;
        cpix[3] = cpix[2] + 1;
        float rd[4], gsC[4];
        bool hole[4];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            rd[c4] = gscale; gsC[c4] = 0.f; hole[c4] = false;
            if (has_norm && k.valid[c4]) {
                const float nr = __ldg(norm + (long long)n * HW + cpix[c4]);
                hole[c4] = (nr == 0.f);
                rd[c4] = gscale / (hole[c4] ? 1.f : nr);
            }
        }
        const float wy1 = (float)(k.y0 + 1) - k.Y, wy0 = k.Y - (float)k.y0;
        const float wx1 = (float)(k.x0 + 1) - k.X, wx0 = k.X - (float)k.x0;
        float gfx = 0.f, gfy = 0.f, sum_gx = 0.f;
        const float* ip = in.p + n * in.sn + y * in.sh + x * in.sw;
        const float* gop = gout.p + n * gout.sn;
        const float* yp = Yf ? Yf + (long long)n * g.C * HW : nullptr;
        for (int c = 0; c < g.C; ++c) {
            const float xv = __ldg(ip + c * in.sc);
            const float xt = pre ? (xv + 1.f) * 0.5f : xv;
            const float A = xt * m;
            float gs[4];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                gs[c4] = 0.f;
                if (k.valid[c4]) {
                    const int cy = k.y0 + (c4 >> 1), cx = k.x0 + (c4 & 1);
                    const float go = __ldg(gop + c * gout.sc + cy * gout.sh + cx * gout.sw);
                    gs[c4] = go * rd[c4];
                    if (has_norm && !hole[c4]) {
                        const float q = __ldg(yp + (long long)c * HW + cpix[c4]) * 0.5f + 0.5f;   // S_c / norm'
                        gsC[c4] -= gs[c4] * q;
                    }
                }
            }
            const float gA = gs[0] * k.w[0] + gs[1] * k.w[1] + gs[2] * k.w[2] + gs[3] * k.w[3];
            if (ginp) {
                float gx = gA;
                if (g.mode == FLDR_SPLAT_SOFTMAX) gx = gA * m * 0.5f;
                else if (g.mode == FLDR_SPLAT_LINEAR) gx = gA * m;
                ginp[(long long)c * HW] = gx;
            }
            sum_gx += gA * xt;
            gfx += A * ((gs[1] - gs[0]) * wy1 + (gs[3] - gs[2]) * wy0);
            gfy += A * ((gs[2] - gs[0]) * wx1 + (gs[3] - gs[1]) * wx0);
        }
        float gAC = 0.f;
        if (has_norm) {
            gAC = gsC[0] * k.w[0] + gsC[1] * k.w[1] + gsC[2] * k.w[2] + gsC[3] * k.w[3];
            gfx += m * ((gsC[1] - gsC[0]) * wy1 + (gsC[3] - gsC[2]) * wy0);
            gfy += m * ((gsC[2] - gsC[0]) * wx1 + (gsC[3] - gsC[1]) * wx0);
        }
        if (gflow) {
            gflow[(long long)n * 2 * HW + pix] = gfx;
            gflow[(long long)n * 2 * HW + HW + pix] = gfy;
        }
        if (gmetric) {
            float gz = sum_gx + gAC;
            if (g.mode == FLDR_SPLAT_SOFTMAX) gz *= m;
            gmetric[(long long)n * HW + pix] = gz;
        }
    }
}

static int make_geom(int mode, int N, int C, int H, int W, bool has_metric, SplatGeom& g) {
    if (mode < 0 || mode > FLDR_SPLAT_RAW) return FLDR_ERR_INVALID_ARGUMENT;
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return FLDR_ERR_INVALID_ARGUMENT;
    if ((long long)H * W >= (1ll << 31)) return FLDR_ERR_UNSUPPORTED;
    if (mode == FLDR_SPLAT_LINEAR && !has_metric) return FLDR_ERR_UNSUPPORTED;   // softSplat.py:328 needs tenMetric
    g.N = N; g.C = C; g.H = H; g.W = W;
    g.mode = mode;
    g.CA = C + (mode_has_norm(mode) ? 1 : 0);
    g.CP = (g.CA + 3) / 4 * 4;
    g.has_metric = (has_metric && (mode == FLDR_SPLAT_LINEAR || mode == FLDR_SPLAT_SOFTMAX)) ? 1 : 0;
    return FLDR_OK;
}

static unsigned grid_for(long long total, int block) {
    long long b = (total + block - 1) / block;
    const long long cap = (long long)sm_count() * 64;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

struct FwdPlan {
    SplatGeom g;
    StreamGeom sg;
    bool stream_ok;       // streaming kernel usable
    bool bounded;         // ring smaller than the frame: reach is bounded, whole-frame fallback must be armed
    size_t ring_bytes, ctrl_bytes, full_bytes, total_bytes;
};

static int plan_fwd(int mode, int N, int C, int H, int W, bool has_metric, FwdPlan& p) {
    int st = make_geom(mode, N, C, H, W, has_metric, p.g);
    if (st != FLDR_OK) return st;
    using namespace stream;
    StreamGeom& sg = p.sg;
    sg.Q = p.g.CP / 4;
    sg.NS = (H + R - 1) / R;
    sg.T = (W + TWC - 1) / TWC;
    const long long NT = (long long)N * sg.NS;
    p.full_bytes = align_up((size_t)N * sg.Q * H * W * 16, 256);
    p.stream_ok = NT * sg.T * sg.Q < (1ll << 28) && NT * sg.T < (1ll << 28);
    // ring: largest power-of-two row count with ring <= 32 MiB (stays L2-resident next to the streaming traffic),
    // no larger than needed to hold every strip; frames whose whole accumulator is <= 64 MiB are held entirely
    long long rows = 64;
    const long long row_bytes = (long long)W * 16 * sg.Q;
    const int ring_mb = get_option(kOptSplatRingMb) > 0 ? get_option(kOptSplatRingMb) : 32;   // tuning hook
    while (rows * 2 * row_bytes <= ((long long)ring_mb << 20)) rows *= 2;
    long long need_rows = 64;
    while (need_rows < NT * R) need_rows *= 2;
    if (rows > need_rows || need_rows * row_bytes <= (64ll << 20)) rows = need_rows;   // whole frame fits: unbounded reach
    sg.RS = (int)(rows / R);
    sg.NT = (int)NT;
    int lag = 2;
    if (sg.RS >= NT) { sg.Ds = sg.NS; p.bounded = false; }
    else {
        // Items are handed out in order, so the ~7 x 148 resident CTAs hold a window of `gif` consecutive groups.
        // N(J) is issued `lag` groups after its last producer S(J+Ds) so that producer has normally finished, and the
        // ring slot of strip J+Ds is not needed again before its previous tenant's N is `lag` groups old as well:
        //   RS >= 2 Ds + 2 lag   ->   Ds = (RS - 2 lag) / 2
        const int lag_env = get_option(kOptSplatLag);   // tuning hook
        const int gif = (kCtasPerSm * 148 + 2 * sg.T * sg.Q - 1) / (2 * sg.T * sg.Q);
        lag = lag_env > 0 ? lag_env : gif + 4;
        sg.Ds = (sg.RS - 2 * lag) / 2;
        p.bounded = true;
        if (sg.Ds < 1) p.stream_ok = false;
    }
    sg.D2 = sg.Ds + lag;
    const int TQ = sg.T * sg.Q;
    const int a = sg.D2 < sg.NT ? sg.D2 : sg.NT;
    sg.nZ = (sg.RS < sg.NT ? sg.RS : sg.NT) * TQ;
    sg.nA = a * TQ;
    sg.nB = sg.NT > a ? (sg.NT - a) * 2 * TQ : 0;
    sg.jstart = sg.NT - sg.D2 > 0 ? sg.NT - sg.D2 : 0;
    sg.total = sg.nZ + sg.nA + sg.nB + (sg.NT - sg.jstart) * TQ;
    sg.vecN = 0;
    p.ring_bytes = align_up((size_t)rows * row_bytes, 256);
    {
        const size_t stream_words = 2 + (size_t)sg.NT + sg.RS + (size_t)sg.NT * sg.T;
        const size_t za_words = 2 + (size_t)N * sg.Q * ((H + 7) / 8);      // zero-ahead scatter: one counter per 8-row strip
        p.ctrl_bytes = align_up((stream_words > za_words ? stream_words : za_words) * 4, 256);
    }
    if (!p.stream_ok) { p.ring_bytes = 0; p.ctrl_bytes = 256; p.bounded = true; }
    p.total_bytes = p.ring_bytes + p.ctrl_bytes + (p.bounded ? p.full_bytes : 0);
    return FLDR_OK;
}

}  // namespace fldr

using namespace fldr;

extern "C" size_t fldr_splat_fwd_workspace_bytes(int mode, int N, int C, int H, int W) {
    FwdPlan p;
    if (plan_fwd(mode, N, C, H, W, true, p) != FLDR_OK) return 0;
    return p.total_bytes;
}

static int launch_normalise(const FwdPlan& p, float* acc, float* out, float* norm, const unsigned* guard, cudaStream_t s) {
    const SplatGeom& g = p.g;
    const int N = g.N, Q = p.sg.Q;
    const long long HW = (long long)g.H * g.W;
    const bool px4 = (HW % 4 == 0) && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                     (!norm || (reinterpret_cast<uintptr_t>(norm) & 15) == 0);
    if (px4) {
        dim3 grid((unsigned)((HW / 4 + 255) / 256), N * Q, 1);
        splat_normalise_kernel<4><<<grid, 256, 0, s>>>(acc, out, norm, g, Q, guard);
    } else {
        dim3 grid((unsigned)((HW + 255) / 256), N * Q, 1);
        splat_normalise_kernel<1><<<grid, 256, 0, s>>>(acc, out, norm, g, Q, guard);
    }
    return check_launch();
}

// whole-frame path: zero + merged scatter + normalise.  `guard` (device flag) makes the three launches no-ops unless set.
static int launch_whole_frame(const FwdPlan& p, const View4& vin, const View4& vfl, const View4& vme, float* acc, float* out,
                              float* norm, const unsigned* guard, cudaStream_t s) {
    const SplatGeom& g = p.g;
    const int N = g.N, H = g.H, W = g.W, Q = p.sg.Q;
    int st;
    if ((long long)N * Q > 65535) return FLDR_ERR_UNSUPPORTED;
    const long long n4 = (long long)N * Q * H * W;
    if (!guard && n4 <= (long long)get_option(kOptSplatFusedMax) && get_option(kOptSplatFusedMax) > 0) {
        // small frame: single cooperative launch (zero / scatter / normalise separated by grid barriers)
        const int wkind = !g.has_metric ? 0 : (g.mode == FLDR_SPLAT_SOFTMAX ? 1 : 2);
        const bool pre = g.mode == FLDR_SPLAT_SOFTMAX;
        const void* fn = wkind == 1 ? (const void*)splat_fused_small_kernel<4, 1, true>
                       : wkind == 2 ? (const void*)splat_fused_small_kernel<4, 2, false>
                       : pre        ? (const void*)splat_fused_small_kernel<4, 0, true>
                                    : (const void*)splat_fused_small_kernel<4, 0, false>;
        static int per_sm[4] = {0, 0, 0, 0};
        const int slot = wkind == 1 ? 0 : wkind == 2 ? 1 : pre ? 2 : 3;
        if (per_sm[slot] == 0) {
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, 256, 0) != cudaSuccess || nb < 1) nb = 1;
            per_sm[slot] = nb;
        }
        const long long units = (long long)N * Q * ((H + 3) / 4) * ((W + 31) / 32);     // warps wanted in phase B
        long long blocks = (units + 7) / 8;
        const long long blocks_c = (n4 + 255) / 256;
        if (blocks < blocks_c) blocks = blocks_c;
        const long long cap = (long long)sm_count() * per_sm[slot];
        if (blocks > cap) blocks = cap;
        SplatGeom gg = g;
        int QQ = Q;
        View4 a0 = vin, a1 = vfl, a2 = vme;
        void* args[] = {&a0, &a1, &a2, &acc, &out, &norm, &gg, &QQ};
        cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(256), args, 0, s);
        if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
        return FLDR_OK;
    }
    if (guard) {
        long long blocks = (n4 + 255) / 256;
        if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
        splat_zero_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<float4*>(acc), n4, guard);
        if ((st = check_launch()) != FLDR_OK) return st;
    } else {
        cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)n4 * 16, s);
        if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
    }
    {
        // rows walked per thread: 16 merges best vertically, but the walk is serial - shorten it until the grid offers
        // at least two waves of CTAs (8 x 128 threads per SM)
        const int bx = W >= 128 ? 128 : ((W + 31) / 32) * 32;
        const long long per_row_ctas = (long long)((W + bx - 1) / bx) * N * Q;
        const long long two_waves = 2ll * sm_count() * 8;
        int R = 16;
        while (R > 4 && per_row_ctas * ((H + R - 1) / R) < two_waves) R >>= 1;
        dim3 grid((W + bx - 1) / bx, (H + R - 1) / R, N * Q);
        const int wkind = !g.has_metric ? 0 : (g.mode == FLDR_SPLAT_SOFTMAX ? 1 : 2);
        const bool pre = g.mode == FLDR_SPLAT_SOFTMAX;
        // quad shape known at compile time for the image splat (C = 3 + weight) and for all-full-quad inputs
        const int qs = (g.C == 3 && g.CA == 4) ? 1 : (g.CA == g.C && g.C % 4 == 0) ? 2 : 0;
        // accumulators larger than ~half the L2 are DRAM-resident when the reductions arrive: prefetch their lines
        const int pf_opt = get_option(kOptSplatPfRows);
        const int pf = ((size_t)n4 * 16 > (48u << 20)) ? (pf_opt > 0 ? pf_opt : (pf_opt < 0 ? 0 : 4)) : 0;
#define FLDR_LAUNCH_SCATTER2(WK_, PRE_, QS_) \
    splat_scatter_merged_kernel<WK_, PRE_, QS_><<<grid, bx, 0, s>>>(vin, vfl, vme, acc, g, Q, guard, pf, R)
#define FLDR_LAUNCH_SCATTER(WK_, PRE_)                                       \
    do {                                                                     \
        if (qs == 1) FLDR_LAUNCH_SCATTER2(WK_, PRE_, 1);                     \
        else if (qs == 2) FLDR_LAUNCH_SCATTER2(WK_, PRE_, 2);                \
        else FLDR_LAUNCH_SCATTER2(WK_, PRE_, 0);                             \
    } while (0)
        if (wkind == 1) FLDR_LAUNCH_SCATTER(1, true);
        else if (wkind == 2) FLDR_LAUNCH_SCATTER(2, false);
        else if (pre) FLDR_LAUNCH_SCATTER(0, true);
        else FLDR_LAUNCH_SCATTER(0, false);
#undef FLDR_LAUNCH_SCATTER2
#undef FLDR_LAUNCH_SCATTER
    }
    if ((st = check_launch()) != FLDR_OK) return st;
    return launch_normalise(p, acc, out, norm, guard, s);
}

extern "C" int fldr_splat_fwd(int mode, const float* in, const int64_t* in_strides, const float* flow,
                              const int64_t* flow_strides, const float* metric, const int64_t* metric_strides,
                              float* out, float* norm, int N, int C, int H, int W, void* ws, size_t ws_bytes,
                              fldr_stream_t stream) {
    FwdPlan p;
    int st = plan_fwd(mode, N, C, H, W, metric != nullptr, p);
    if (st != FLDR_OK) return st;
    if (!in || !in_strides || !flow || !flow_strides || !out || (metric && !metric_strides)) return FLDR_ERR_INVALID_ARGUMENT;
    FwdPlan pmax;
    plan_fwd(mode, N, C, H, W, true, pmax);
    if (!ws || ws_bytes < pmax.total_bytes) return FLDR_ERR_WORKSPACE_TOO_SMALL;
    if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return FLDR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const SplatGeom& g = p.g;
    const View4 vin = make_view(in, in_strides), vfl = make_view(flow, flow_strides), vme = make_view(metric, metric_strides);
    char* base = static_cast<char*>(ws);
    float4* ring = reinterpret_cast<float4*>(base);
    unsigned* ctrl = reinterpret_cast<unsigned*>(base + p.ring_bytes);
    float* full = reinterpret_cast<float*>(base + p.ring_bytes + p.ctrl_bytes);

    // Default: whole-frame path.  The streaming kernel is opt-in (fldr_set_option("splat_stream", 1)): it cuts DRAM traffic
    // of the 4K image splat from 943 MB to 349 MB but, at ~215 us against ~205 us, does not yet beat the three-pass
    // sequence, and its L2-sized ring bounds the vertical flow it can take without the fallback.
    // whole-frame accumulator: behind ring + ctrl when the plan is bounded, else the (frame-sized) ring region itself
    if (!p.stream_ok || get_option(kOptSplatStream) == 0) {
        float* acc = p.bounded ? full : reinterpret_cast<float*>(ring);
        const long long n4 = (long long)N * p.sg.Q * H * W;
        // Zero-ahead scatter (opt-in: "splat_za" = 8 or 16 rows per strip).  It removes the separate zero fill and most
        // of the accumulator re-fetch (DRAM 640 -> 482 MB for the 4K image scatter) but its per-CTA hand-shake
        // (ticket, zero tile + publish, poll) costs as much as it saves: 146 us vs 27 + 111 us (profiles/).
        const int za = get_option(kOptSplatZa);
        if ((za == 8 || za == 16) && (size_t)n4 * 16 > (48u << 20) && p.stream_ok && H >= 256) {
            const int R = (za == 16) ? 16 : 8;
            const int Tc = (W + 127) / 128;
            const int NSr = (H + R - 1) / R;
            const int Rs = 128 / R;                                     // +-128 rows of vertical reach
            const int span = (sm_count() * 8 + Tc - 1) / Tc;            // strips covered by the resident CTAs
            const int D = Rs + 1 + span + 4;
            const long long total_strips = (long long)N * p.sg.Q * NSr;
            const long long tickets = (total_strips < D ? total_strips : D) * Tc + total_strips * Tc;
            if (tickets < (1ll << 31)) {
                cudaError_t e = cudaMemsetAsync(ctrl, 0, p.ctrl_bytes, s);
                if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
                const int wkind = !g.has_metric ? 0 : (g.mode == FLDR_SPLAT_SOFTMAX ? 1 : 2);
                const bool pre = g.mode == FLDR_SPLAT_SOFTMAX;
                const int qs = (g.C == 3 && g.CA == 4) ? 1 : (g.CA == g.C && g.C % 4 == 0) ? 2 : 0;
                const int Q = p.sg.Q;
#define FLDR_LAUNCH_ZA3(R_, WK_, PRE_, QS_) \
    splat_scatter_za_kernel<R_, WK_, PRE_, QS_><<<(unsigned)tickets, 128, 0, s>>>(vin, vfl, vme, acc, g, Q, ctrl, Tc, NSr, D, Rs)
#define FLDR_LAUNCH_ZA2(WK_, PRE_, QS_) do { if (R == 16) FLDR_LAUNCH_ZA3(16, WK_, PRE_, QS_); else FLDR_LAUNCH_ZA3(8, WK_, PRE_, QS_); } while (0)
#define FLDR_LAUNCH_ZA(WK_, PRE_) do { if (qs == 1) FLDR_LAUNCH_ZA2(WK_, PRE_, 1); else if (qs == 2) FLDR_LAUNCH_ZA2(WK_, PRE_, 2); else FLDR_LAUNCH_ZA2(WK_, PRE_, 0); } while (0)
                if (wkind == 1) FLDR_LAUNCH_ZA(1, true);
                else if (wkind == 2) FLDR_LAUNCH_ZA(2, false);
                else if (pre) FLDR_LAUNCH_ZA(0, true);
                else FLDR_LAUNCH_ZA(0, false);
#undef FLDR_LAUNCH_ZA
#undef FLDR_LAUNCH_ZA2
#undef FLDR_LAUNCH_ZA3
                if ((st = check_launch()) != FLDR_OK) return st;
                if ((st = launch_normalise(p, acc, out, norm, nullptr, s)) != FLDR_OK) return st;
                // bounded reach: arm the plain path; its launches exit at once unless the scatter flagged an overflow
                return launch_whole_frame(p, vin, vfl, vme, acc, out, norm, ctrl + 1, s);
            }
        }
        return launch_whole_frame(p, vin, vfl, vme, acc, out, norm, nullptr, s);
    }

    cudaError_t e = cudaMemsetAsync(ctrl, 0, p.ctrl_bytes, s);
    if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
    p.sg.vecN = (W % 4 == 0) && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (!norm || (reinterpret_cast<uintptr_t>(norm) & 15) == 0);
    {
        static int ctas_per_sm = 0;
        if (ctas_per_sm == 0) {
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, splat_stream_kernel<1, true, 0>, stream::TWC, 0) != cudaSuccess || nb < 1) nb = 4;
            ctas_per_sm = nb > stream::kCtasPerSm ? stream::kCtasPerSm : nb;
        }
        long long grid = (long long)sm_count() * ctas_per_sm;
        if (grid > p.sg.total) grid = p.sg.total;
        const int wkind = !g.has_metric ? 0 : (g.mode == FLDR_SPLAT_SOFTMAX ? 1 : 2);
        const bool pre = g.mode == FLDR_SPLAT_SOFTMAX;
        const int qs = (g.C == 3 && g.CA == 4) ? 1 : (g.CA == g.C && g.C % 4 == 0) ? 2 : 0;
        // experiment hook ("splat_l2_persist" = 1): pin the ring in the persisting L2 carve-out for this launch
        const bool persist = get_option(kOptSplatL2Persist) > 0;
        if (persist) {
            static bool limit_set = false;
            if (!limit_set) { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)96 << 20); limit_set = true; }
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof(av));
            av.accessPolicyWindow.base_ptr = ring;
            av.accessPolicyWindow.num_bytes = p.ring_bytes;
            av.accessPolicyWindow.hitRatio = 1.0f;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av);
        }
#define FLDR_LAUNCH_STREAM2(WK_, PRE_, QS_) \
    splat_stream_kernel<WK_, PRE_, QS_><<<(unsigned)grid, stream::TWC, 0, s>>>(vin, vfl, vme, ring, ctrl, out, norm, g, p.sg)
#define FLDR_LAUNCH_STREAM(WK_, PRE_)                        \
    do {                                                     \
        if (qs == 1) FLDR_LAUNCH_STREAM2(WK_, PRE_, 1);      \
        else if (qs == 2) FLDR_LAUNCH_STREAM2(WK_, PRE_, 2); \
        else FLDR_LAUNCH_STREAM2(WK_, PRE_, 0);              \
    } while (0)
        if (wkind == 1) FLDR_LAUNCH_STREAM(1, true);
        else if (wkind == 2) FLDR_LAUNCH_STREAM(2, false);
        else if (pre) FLDR_LAUNCH_STREAM(0, true);
        else FLDR_LAUNCH_STREAM(0, false);
#undef FLDR_LAUNCH_STREAM2
#undef FLDR_LAUNCH_STREAM
        if (persist) {
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof(av));
            cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av);
        }
        if ((st = check_launch()) != FLDR_OK) return st;
    }
    // bounded reach: arm the whole-frame path; its launches exit at once unless the streaming pass flagged an overflow
    if (p.bounded) return launch_whole_frame(p, vin, vfl, vme, full, out, norm, ctrl + 1, s);
    return FLDR_OK;
}

extern "C" size_t fldr_splat_bwd_workspace_bytes(int mode, int N, int C, int H, int W) {
    (void)mode; (void)N; (void)C; (void)H; (void)W;
    return 0;
}

extern "C" int fldr_splat_bwd(int mode, const float* in, const int64_t* in_strides, const float* flow,
                              const int64_t* flow_strides, const float* metric, const int64_t* metric_strides,
                              const float* out, const float* norm, const float* grad_out,
                              const int64_t* grad_out_strides, float* grad_in, float* grad_flow, float* grad_metric,
                              int N, int C, int H, int W, void* ws, size_t ws_bytes, fldr_stream_t stream) {
    (void)ws; (void)ws_bytes;
    SplatGeom g;
    int st = make_geom(mode, N, C, H, W, metric != nullptr, g);
    if (st != FLDR_OK) return st;
    if (!in || !in_strides || !flow || !flow_strides || !grad_out || !grad_out_strides || (metric && !metric_strides))
        return FLDR_ERR_INVALID_ARGUMENT;
    if (mode_has_norm(mode) && (!out || !norm)) return FLDR_ERR_INVALID_ARGUMENT;
    if (grad_metric && !g.has_metric) return FLDR_ERR_INVALID_ARGUMENT;
    if (!grad_in && !grad_flow && !grad_metric) return FLDR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long total = (long long)N * H * W;
    splat_bwd_kernel<<<grid_for(total, 256), 256, 0, s>>>(make_view(in, in_strides), make_view(flow, flow_strides),
                                                          make_view(metric, metric_strides), out, norm,
                                                          make_view(grad_out, grad_out_strides), grad_in, grad_flow,
                                                          grad_metric, g);
    return check_launch();
}
