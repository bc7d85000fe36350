// Softmax / average / linear / summation forward splat and its backward, sm_100a.
//
// Replaces softSplat.py:12-158 (three CuPy string kernels) AND the ~12 torch elementwise kernels
// FunctionSoftsplat wraps around them (softSplat.py:320-352): pre-scale, exp, cat, zero-init,
// normaliser fix-up, divide, post-scale are all folded into the passes below.
//
// Data layout in HBM
//   inputs      NCHW fp32 with arbitrary element strides (callers pass views: fLDRnet.py:386,449)
//   accumulator [N][Q][H][W + 2][4] fp32, Q = ceil((C + has_norm) / 4): every channel quad is a pixel-interleaved
//               float4 image with one guard cell either side of a row (corners one column outside the frame land
//               there and are never read: no x test on the reductions), so one corner of one source pixel is ONE 16-byte red.global.add.v4.f32 and adjacent
//               lanes (adjacent x) reduce into adjacent slots (the reference issues 4 scalar REDs per element,
//               softSplat.py:39-50).  The normaliser rides in slot C % 4 of quad C / 4.
//   outputs     NCHW contiguous (what the reference allocates, softSplat.py:234).
//
// Paths (fldr_splat_fwd picks; DESIGN.md section 4.1 has the measurements behind each choice)
//   tiny frames    splat_fused_small_kernel: zero + scatter + normalise in one cooperative launch
//   default        splat_zero_kernel (front to back, L2 eviction hints) + splat_scatter_tile_kernel (back to front; inputs staged by
//                  TMA, merged reductions) + splat_normalise_kernel (front to back; drops the dead accumulator lines from L2)
//   odd views      the same three passes with splat_scatter_merged_kernel (plain loads): rows that are not 16-byte aligned,
//                  non-unit pixel stride, metric together with >= 4 channels
// (Single-launch variants that keep the accumulator in L2 - streaming ring, tile owner, cluster per sample - were built and
//  measured in round 2 and lost on time: profiles/r2_splat_*_negative_result.txt.)
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <cooperative_groups.h>

#include "splat.cuh"
#include "tma.cuh"

namespace cg = cooperative_groups;

#ifndef FLDR_ZERO_KEEP_MB
#define FLDR_ZERO_KEEP_MB 48
#endif

#ifndef FLDR_SCATTER_MIN_CTAS
#define FLDR_SCATTER_MIN_CTAS 8
#endif

namespace fldr {

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

// Optional debug flag (fldr_splat_set_nonfinite_flag): the reference device-asserts on a non-finite target coordinate
// (softSplat.py:25-26) and kills the context; here the pixel is skipped and, when a flag word is registered, that word is set
// so a caller can find out (the wrapper polls it only in debug mode: reading it costs a synchronisation).
__device__ unsigned* g_nonfinite_flag = nullptr;
__device__ __forceinline__ void note_nonfinite(float X, float Y) {
    if (!(fabsf(X) <= 3.0e38f && fabsf(Y) <= 3.0e38f)) {
        unsigned* f = g_nonfinite_flag;
        if (f) *f = 1u;
    }
}

template <int PX> __device__ __forceinline__ void vstore(float* p, const float* t);
template <> __device__ __forceinline__ void vstore<1>(float* p, const float* t) { __stcs(p, t[0]); }
template <> __device__ __forceinline__ void vstore<4>(float* p, const float* t) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(t[0], t[1], t[2], t[3]));
}

// Whole-frame accumulator plane [H][W + 2] float4 of one (sample, quad); pixel x lives in cell x + 1.
struct PlaneAcc {
    float4* base;
    int W, P;           // frame width, row pitch in cells (W + 2)
    __device__ __forceinline__ float4* at(int x, int y) const { return base + (y * P + x + 1); }
    // The whole-frame accumulator lives in DRAM; a reduction into a line that is not in L2 stalls the L2 reduction unit
    // on the fill.  Pulling the target lines of the rows this thread will reach next into L2 ahead of time turns those
    // fills into ordinary, well-pipelined reads.
    int H, pf_rows;     // pf_rows: how many accumulator rows ahead of the row being scattered are pulled into L2
    __device__ __forceinline__ int prefetch_rows() const { return pf_rows; }
    __device__ __forceinline__ void prefetch(int x, int y) const {
        if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (y * P + x + 1)));
    }
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
template <int WKIND, bool PRE, int QS, class Acc>
__device__ __forceinline__ void scatter_rows(const View4& in, const View4& flow, const View4& metric, const SplatGeom& g,
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
        const bool ok = inb && fx0 >= -1.f && fx0 < (float)W && fy0 >= -1.f && fy0 < (float)H;   // false for NaN / inf
        if (inb && !ok) note_nonfinite(X, Y);
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
}

template <int WKIND, bool PRE, int QS>
__global__ void __launch_bounds__(128, FLDR_SCATTER_MIN_CTAS) splat_scatter_merged_kernel(View4 in, View4 flow, View4 metric,
                                                                   float* __restrict__ acc, SplatGeom g, int Q, int pf_rows, int R, int flip) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int yb = (flip ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y) * R;          // back to front, like the tile kernel
    const int bz = flip ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
    const int q = bz % Q, n = bz / Q;
    PlaneAcc pa;
    pa.P = g.W + 2;
    pa.base = reinterpret_cast<float4*>(acc) + (long long)(n * Q + q) * g.H * pa.P;
    pa.W = g.W;
    pa.H = g.H;
    pa.pf_rows = pf_rows;
    scatter_rows<WKIND, PRE, QS>(in, flow, metric, g, n, q, x, yb, min(R, g.H - yb), pa);
}

// ------------------------------------------------------------------------------------------------
// Pass 1, default: tile kernel.  A CTA owns 8 source rows x 128 columns of one channel quad.  Thread 0 stages the tile's
// inputs in shared memory with two or three TMA box loads ([planes][8][128]; rows / columns / channels outside the
// tensor are zero-filled by the TMA unit), so the walk below has no address arithmetic on the inputs and never waits for
// DRAM row by row (the other resident CTAs cover the one load latency per tile).
// The walk is the merged-reduction scheme of scatter_rows in a leaner form (144 instead of 227 instructions per
// warp-row): a corner is identified by its accumulator CELL OFFSET (negative = not in frame), two corners merge iff they
// are the same memory cell, and the guard cells of the row pitch remove every x test from the reductions.
// ------------------------------------------------------------------------------------------------
namespace tile {
#ifndef FLDR_TILE_R
#define FLDR_TILE_R 8
#endif
constexpr int R = FLDR_TILE_R, TW = 128, PLANES = 6;
constexpr int kSent = -(1 << 30);             // "no cell": stays negative after + 1
}  // namespace tile

// one [planes][8 rows][128 columns] box of an NCHW input, global -> shared; the inputs are read exactly once: evict-first in L2
__device__ __forceinline__ void tma_box_load(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int c, int n, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(c), "r"(n), "l"(pol)
        : "memory");
}
// cell `off` of plane `base` += v, only when off >= 0
__device__ __forceinline__ void red4_at(float4* base, int off, const float* v) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u64 a;\n\t"
        "setp.ge.s32 p, %1, 0;\n\t"
        "mad.wide.s32 a, %1, 16, %0;\n\t"
        "@p red.global.add.v4.f32 [a], {%2, %3, %4, %5};\n\t}"
        ::"l"(base), "r"(off), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]));
}
__device__ __forceinline__ void prefetch_l2_at(float4* base, int off) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u64 a;\n\t"
        "setp.ge.s32 p, %1, 0;\n\t"
        "mad.wide.s32 a, %1, 16, %0;\n\t"
        "@p prefetch.global.L2 [a];\n\t}"
        ::"l"(base), "r"(off));
}
// e^z to ~2 ulp for any z: ex2.approx of the rounded product, corrected by the product's rounding error and the
// representation error of log2(e) (the plain ex2(z * log2e) loses |z| * 6e-8 relative; expf costs twice the instructions)
__device__ __forceinline__ float exp_splat(float z) {
    const float l2e = 1.4426950408889634f;
    const float t = z * l2e;
    const float e = fmaf(z, l2e, -t) + z * 1.92596299e-8f;
    float p;
    asm("ex2.approx.f32 %0, %1;" : "=f"(p) : "f"(t));
    return fmaf(p, e * 0.6931471805599453f, p);
}

template <int WKIND, bool PRE, int QS, int TR>
__global__ void __launch_bounds__(tile::TW, 6) splat_scatter_tile_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                                         const __grid_constant__ CUtensorMap tm_flow,
                                                                         const __grid_constant__ CUtensorMap tm_metric,
                                                                         float* __restrict__ acc, SplatGeom g, int Q, int nbox,
                                                                         int pf_rows, int flip) {
    using namespace tile;
    constexpr int R = TR;                                // rows of a tile (shadows tile::R)
    __shared__ __align__(128) float st[PLANES * R * TW];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, lane = tid & 31;
    // flip: tiles are dispatched bottom-up, so the scatter starts on the accumulator rows the zero fill wrote LAST (still in L2)
    // and ends on the rows the normalise pass reads FIRST
    const int x0 = blockIdx.x * TW, yb = (flip ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y) * R;
    const int bz = flip ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
    const int q = bz % Q, n = bz / Q;
    const int nch = QS == 1 ? 3 : QS == 2 ? 4 : min(4, g.C - q * 4);
    const int wslot = QS == 1 ? 3 : QS == 2 ? -1 : ((g.CA > g.C) ? g.C - q * 4 : -1);
    constexpr int HM = WKIND ? 1 : 0, CH0 = 2 + HM;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        mbar_arrive_expect_tx(&bar, (uint32_t)((2 + HM + (nch > 0 ? nbox : 0)) * R * TW * 4));
        tma_box_load(st, &tm_flow, &bar, x0, yb, 0, n, pol);
        if (HM) tma_box_load(st + 2 * R * TW, &tm_metric, &bar, x0, yb, 0, n, pol);
        if (nch > 0) tma_box_load(st + CH0 * R * TW, &tm_in, &bar, x0, yb, q * 4, n, pol);
    }
    __syncthreads();                                   // barrier initialised before anyone waits on it
    const int rows = min(R, g.H - yb), P = g.W + 2;
    float4* rq = reinterpret_cast<float4*>(acc) + (size_t)(n * Q + q) * g.H * P;
    const bool inb = x0 + tid < g.W;
    const float Wf = (float)g.W, Hf = (float)g.H, Hm1 = (float)(g.H - 1);
    const float xf = (float)(x0 + tid);
    float yf = (float)yb;
    const int src_lane = (lane + 31) & 31;
    const bool lane_gt0 = lane > 0;
    const int pfo = pf_rows * P;
    int prev_t = kSent;
    float pw[4] = {0.f, 0.f, 0.f, 0.f};
    const float* sp = st + tid;
    mbar_wait(&bar, 0);
#pragma unroll 2
    for (int r = 0; r < rows; ++r, yf += 1.f, sp += TW) {
        const float u = sp[0], v = sp[R * TW];
        float m = 1.f;
        if (WKIND == 1) m = exp_splat(sp[2 * R * TW]);
        if (WKIND == 2) m = sp[2 * R * TW];
        float xv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = (j < nch) ? sp[(CH0 + j) * R * TW] : 0.f;
        // softSplat.py:23-38
        const float X = xf + u, Y = yf + v;
        const float fx0 = floorf(X), fy0 = floorf(Y);
        const float x1f = fx0 + 1.f, y1f = fy0 + 1.f;
        const bool pR = inb && fx0 >= -1.f && fx0 < Wf && fy0 >= -1.f && fy0 < Hf;   // false for NaN / inf (the reference asserts, 25-26)
        if (inb && !pR) note_nonfinite(X, Y);
        const bool pT = pR && fy0 >= 0.f;                        // top corners in frame
        const bool pB = pR && fy0 < Hm1;                         // bottom corners in frame
        const int cx = (int)x1f;                                 // x0 + 1 = cell of the W corners (guard cell at 0)
        const int y0 = (int)fy0;
        const int tT = pT ? y0 * P + cx : kSent;
        const int tB = pB ? (y0 + 1) * P + cx : kSent;
        if (pfo) prefetch_l2_at(rq, pR && fy0 + (float)pf_rows < Hm1 ? (y0 + 1) * P + cx + pfo : kSent);
        const float ax = x1f - X, bx = X - fx0, ay = y1f - Y, by = Y - fy0;
        const float wNW = ax * ay, wNE = bx * ay, wSW = ax * by, wSE = bx * by;
        // a corner that is not in frame keeps whatever this arithmetic produces (NaN included): it is never merged into a
        // valid corner (cell offsets differ) and never issued
        const float hm = PRE ? 0.5f * m : m;                     // ((x+1)*0.5)*m == (x+1)*(0.5*m) exactly
        float tW[4], tE[4], bW[4], bE[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a;
            if (j < nch) a = PRE ? (xv[j] + 1.f) * hm : xv[j] * m;
            else a = (j == wslot) ? m : 0.f;
            tW[j] = a * wNW; tE[j] = a * wNE; bW[j] = a * wSW; bE[j] = a * wSE;
        }
        // vertical: the bottom-W corner carried from the previous row joins this row's top-W corner - or its top-E corner when the
        // flow moved one cell to the left between the rows - or is flushed
        const bool vm = tT == prev_t;
        // (extra matches only for the image splat, QS == 1: its accumulator is DRAM-sized and the pass is bound by the number of
        //  32-byte sectors its reductions touch - 16.2 M -> 13.6 M at 4K, 120 -> 111 us; the L2-resident feature splats are
        //  latency-bound and only pay for the extra instructions)
        const bool vme = QS == 1 && prev_t == tT + 1;    // never true for kSent (tT + 1 stays far below any cell)
        if (QS != 1) red4_at(rq, vm ? kSent : prev_t, pw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tW[j] = vm ? tW[j] + pw[j] : tW[j];
            tE[j] = vme ? tE[j] + pw[j] : tE[j];
        }
        // horizontal: the E corners travel to the lane on the right (rotate: lane 0 gets lane 31's and only forwards them)
        const int rT = __shfl_sync(0xffffffffu, tT + 1, src_lane);
        const int rB = __shfl_sync(0xffffffffu, tB + 1, src_lane);
        float rt4[4], rb4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            rt4[j] = __shfl_sync(0xffffffffu, tE[j], src_lane);
            rb4[j] = __shfl_sync(0xffffffffu, bE[j], src_lane);
        }
        if (QS == 1) {
            // a carried corner that found no partner in its own column often IS the cell of a corner just received (under vertical
            // shear the left neighbour's top-E corner lands where this column's previous row put its bottom-W): join them, else flush
            const bool vrT = !(vm || vme) && prev_t == rT, vrB = !(vm || vme) && prev_t == rB;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                rt4[j] = vrT ? rt4[j] + pw[j] : rt4[j];
                rb4[j] = vrB ? rb4[j] + pw[j] : rb4[j];
            }
            red4_at(rq, (vm || vme || vrT || vrB) ? kSent : prev_t, pw);
        }
        // straight matches (same row) and cross matches (the flow's vertical shear moved the neighbour one row up or down: its top-E
        // corner is my bottom-W cell, or its bottom-E corner my top-W cell).  A received corner matches at most one of my two cells.
        const bool takeT = lane_gt0 && rT == tT, takeB = lane_gt0 && rB == tB;
        const bool crossT = QS == 1 && lane_gt0 && rT == tB, crossB = QS == 1 && lane_gt0 && rB == tT;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tW[j] = takeT ? tW[j] + rt4[j] : (crossB ? tW[j] + rb4[j] : tW[j]);
            bW[j] = takeB ? bW[j] + rb4[j] : (crossT ? bW[j] + rt4[j] : bW[j]);
        }
        red4_at(rq, (takeT || crossT) ? kSent : rT, rt4);
        red4_at(rq, (takeB || crossB) ? kSent : rB, rb4);
        red4_at(rq, tT, tW);
        prev_t = tB;
#pragma unroll
        for (int j = 0; j < 4; ++j) pw[j] = bW[j];
    }
    red4_at(rq, prev_t, pw);
}

// Zero fill with L2-allocating stores, front to back: what it wrote last is what the bottom-up scatter touches first.
__global__ void __launch_bounds__(256) splat_zero_kernel(float4* __restrict__ acc, long long n4, long long keep_from) {
    const long long i0 = (long long)blockIdx.x * (256 * 8) + threadIdx.x;
    // L2 eviction priorities: the head of the buffer (touched last by the back-to-front scatter) may leave L2 at once, the tail should
    // stay (170.5 -> 165.4 us on the 4K image splat, 124 -> 118 us on 32 x 3 x 512^2; 32 / 64 / 96 MB of tail measured alike)
    uint64_t pol_first, pol_last;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const long long i = i0 + k * 256;
        if (i < n4)
            asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %1, %1, %1}, %2;" ::"l"(acc + i), "f"(0.f), "l"(i >= keep_from ? pol_last : pol_first) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Pass 2: normalise + post-scale + quad-interleaved -> NCHW.  One thread per (PX consecutive pixels, channel quad):
// PX float4 loads of the accumulator (+ PX of the quad holding the normaliser), 4 channel-plane stores of PX floats.
//   softSplat.py:343-349: norm==0 -> 1, divide, (y - 0.5) * 2 (post-scale in every mode but RAW).
// ------------------------------------------------------------------------------------------------
template <int PX, bool DISCARD = false>
__global__ void __launch_bounds__(128) splat_normalise_kernel(const float* __restrict__ acc, float* __restrict__ out,
                                                              float* __restrict__ norm_out, SplatGeom g, int Q) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * PX;
    if (DISCARD) {
        // The accumulator is dead once it has been read.  Whatever part of it is still dirty in L2 would be written back to DRAM
        // for nothing: after every thread of the CTA has CONSUMED its loads (the barrier below), the 128-byte lines that lie
        // entirely inside this CTA's cells are dropped from L2 without a write-back (discard.global.L2).  Single-plane frames only
        // (Q == 1): nobody else reads these cells.
        const int y = blockIdx.y, n = blockIdx.z;
        const int P = g.W + 2;
        const long long HW = (long long)g.H * g.W;
        const float4* row = reinterpret_cast<const float4*>(acc) + ((long long)n * g.H + y) * P + 1;
        float4 s4[PX];
        float yv[4][PX], nrm[PX];
        const bool live = x < g.W;
        if (live) {
#pragma unroll
            for (int k = 0; k < PX; ++k) s4[k] = __ldcs(row + x + k);
            const int slot = g.C & 3;
#pragma unroll
            for (int k = 0; k < PX; ++k) {
                nrm[k] = slot == 0 ? s4[k].x : slot == 1 ? s4[k].y : slot == 2 ? s4[k].z : s4[k].w;
                const float d = norm_recip(nrm[k]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float sv = j == 0 ? s4[k].x : j == 1 ? s4[k].y : j == 2 ? s4[k].z : s4[k].w;
                    yv[j][k] = post_scale(sv, d, false, true);
                }
            }
        }
        __syncthreads();                     // every load of the CTA has been consumed
        {
            const int c0 = blockIdx.x * blockDim.x * PX;                                   // first pixel of the CTA in this row
            const int nc = min(g.W - c0, (int)blockDim.x * PX);
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(row + c0), a1 = a0 + (uintptr_t)nc * 16;
            const uintptr_t l0 = (a0 + 127) & ~(uintptr_t)127;
            const uintptr_t line = l0 + (uintptr_t)threadIdx.x * 128;
            if (line + 128 <= a1) asm volatile("discard.global.L2 [%0], 128;" ::"l"(line) : "memory");
        }
        if (live) {
            const long long pix = (long long)y * g.W + x;
            if (norm_out) vstore<PX>(norm_out + (long long)n * HW + pix, nrm);
            float* op = out + (long long)n * g.C * HW + pix;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < g.C) vstore<PX>(op + (long long)j * HW, yv[j]);
        }
        return;
    }
    if (x >= g.W) return;       // PX > 1 only when W % PX == 0
    const int y = blockIdx.y;
    const int q = blockIdx.z % Q, n = blockIdx.z / Q;
    const long long HW = (long long)g.H * g.W;
    const long long pix = (long long)y * g.W + x;
    const int P = g.W + 2;
    const bool has_norm = g.CA > g.C;
    const float4* row = reinterpret_cast<const float4*>(acc) + ((long long)n * Q * g.H + y) * P + 1 + x;     // quad 0 of this row
    const long long qstride = (long long)g.H * P;
    float4 s4[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) s4[k] = __ldcs(row + q * qstride + k);
    float d[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) d[k] = 1.f;
    if (has_norm) {
        const int qn = g.C >> 2, slot = g.C & 3;
        float nrm[PX];
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            const float4 n4 = (qn == q) ? s4[k] : __ldcs(row + qn * qstride + k);
            nrm[k] = slot == 0 ? n4.x : slot == 1 ? n4.y : slot == 2 ? n4.z : n4.w;
            d[k] = norm_recip(nrm[k]);     // hole fix-up + one reciprocal per pixel (splat.cuh)
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
                yv[k] = post_scale(sv, d[k], g.mode == FLDR_SPLAT_RAW, has_norm);
            }
            vstore<PX>(op + (long long)c * HW, yv);
        }
    }
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
    const int P = g.W + 2;
    const long long HP = (long long)g.H * P;               // cells per (sample, quad) plane
    const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long gthreads = (long long)gridDim.x * blockDim.x;
    float4* acc4 = reinterpret_cast<float4*>(acc);
    // ---- A: zero the accumulator (guard cells included)
    for (long long i = gtid; i < (long long)g.N * Q * HP; i += gthreads) acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
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
            pa.base = acc4 + (long long)nq * HP;
            pa.W = g.W;
            pa.P = P;
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
            const long long cell = (pix / g.W) * P + 1 + (pix % g.W);      // inside a plane
            const float4 s4 = __ldcg(acc4 + nq * HP + cell);
            float d = 1.f;
            if (has_norm) {
                const float4 n4v = (qn == q) ? s4 : __ldcg(acc4 + ((long long)n * Q + qn) * HP + cell);
                const float nrm = slot == 0 ? n4v.x : slot == 1 ? n4v.y : slot == 2 ? n4v.z : n4v.w;
                if (norm_out && q == 0) norm_out[(long long)n * HW + pix] = nrm;
                d = norm_recip(nrm);
            }
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
            float* op = out + (long long)n * g.C * HW + pix;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = q * 4 + j;
                if (c < g.C) op[(long long)c * HW] = post_scale(sv[j], d, g.mode == FLDR_SPLAT_RAW, has_norm);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward, two passes (SURVEY.md App. A.2 for the formulas):
//   gS_c = 2 gY_c / norm'            gS_C = -sum_c gS_c * (S_c / norm')   (0 where norm was 0)
//   gA   = sum_corners w * gS        (kernel_Softsplat_updateGradInput, softSplat.py:84-95)
//   gF   = sum_c A_c * sum_corners gS_c * dw   (kernel_Softsplat_updateGradFlow, 130-155)
//   softmax: g_x = gA_c * e^z / 2 ; g_z = e^z (sum_c gA_c x~_c + gA_C)      linear: g_x = gA_c z ; g_z = sum_c gA_c x_c + gA_C
// Pass 1 (splat_bwd_prep_kernel) is elementwise over TARGET pixels: it forms gS from grad_out, the forward output and the
//   saved normaliser and stores it pixel-interleaved, [N][QB][H + 2][W + 2] float4 with a zero border - the same quad layout
//   the forward accumulates in (gS_C in slot C % 4 of quad C / 4, only when a flow / metric gradient is requested).
// Pass 2 (splat_bwd_gather_kernel) is one thread per SOURCE pixel: its four corners are four 16-byte loads per channel quad
//   (the zero border stands in for every bounds test), against 4 x (2C + 1) scalar gathers with their own address arithmetic
//   when gS was formed on the fly (round 1: 590 instructions per pixel and 27 % of the HBM roofline on the cfg5 image splat).
// ------------------------------------------------------------------------------------------------
namespace bwdp {
constexpr int NT = 128;
}
// thread = 4 consecutive target pixels of one row (VEC) or one pixel; blockIdx.y = padded row, blockIdx.z = n * QB + q
template <bool VEC, int MINB>
__global__ void __launch_bounds__(bwdp::NT, MINB) splat_bwd_prep_kernel(View4 gout, const float* __restrict__ Yf, const float* __restrict__ norm,
                                                                  float4* __restrict__ G, SplatGeom g, int QB, int need_c) {
    constexpr int PX = VEC ? 4 : 1;
    const int P = g.W + 2;
    // threads are laid over the (padded row, pixel group) pairs of a plane in one flat index: narrow frames (the coarse training
    // levels are 64 / 32 / 16 pixels wide) still fill their CTAs
    const int gpr = (P + PX - 1) / PX;                      // pixel groups per padded row
    const int flat = blockIdx.x * blockDim.x + threadIdx.x;
    const int yp = flat / gpr;                              // padded row: 0 and H + 1 are the zero border
    if (yp > g.H + 1) return;
    const int q = blockIdx.z % QB, n = blockIdx.z / QB;
    float4* grow = G + ((size_t)(n * QB + q) * (g.H + 2) + yp) * P;
    const int x = (flat - yp * gpr) * PX;
    if (yp == 0 || yp == g.H + 1) {
        for (int k = 0; k < PX; ++k)
            if (x + k < P) grow[x + k] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    if (x >= g.W) return;
    const int y = yp - 1;
    if (x == 0) grow[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x + PX >= g.W) grow[g.W + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool has_norm = g.CA > g.C;
    const float gscale = (g.mode == FLDR_SPLAT_RAW) ? 1.f : 2.f;
    const long long HW = (long long)g.H * g.W;
    const long long pix = (long long)y * g.W + x;
    float rd[PX];
    bool hole[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) { rd[k] = gscale; hole[k] = false; }
    if (has_norm) {
        float nr[PX];
        if (VEC) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(norm + (long long)n * HW + pix));
            nr[0] = t.x; if (PX > 1) { nr[1] = t.y; nr[2] = t.z; nr[3] = t.w; }
        } else nr[0] = __ldg(norm + (long long)n * HW + pix);
#pragma unroll
        for (int k = 0; k < PX; ++k) { hole[k] = nr[k] == 0.f; rd[k] = __fdividef(gscale, hole[k] ? 1.f : nr[k]); }
    }
    const float* gop = gout.p + n * gout.sn + (long long)y * gout.sh + (long long)x * gout.sw;
    const float* yp_ = Yf ? Yf + (long long)n * g.C * HW + pix : nullptr;
    auto load_go = [&](int c, float* v) {
        if (VEC) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(gop + (long long)c * gout.sc));
            v[0] = t.x; if (PX > 1) { v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        } else v[0] = __ldcs(gop + (long long)c * gout.sc);
    };
    auto load_y = [&](int c, float* v) {
        if (VEC) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(yp_ + (long long)c * HW));
            v[0] = t.x; if (PX > 1) { v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        } else v[0] = __ldg(yp_ + (long long)c * HW);
    };
    float o[4][PX];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < PX; ++k) o[j][k] = 0.f;
    const int qn = g.C >> 2, slot = g.C & 3;               // where gS_C lives
    float sC[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) sC[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = q * 4 + j;
        if (c < g.C) {
            float go[PX];
            load_go(c, go);
#pragma unroll
            for (int k = 0; k < PX; ++k) o[j][k] = go[k] * rd[k];
            if (need_c && q == qn) {
                float yv[PX];
                load_y(c, yv);
#pragma unroll
                for (int k = 0; k < PX; ++k) sC[k] -= o[j][k] * (yv[k] * 0.5f + 0.5f);     // S_c / norm'
            }
        }
    }
    if (need_c && q == qn) {
        // the channels of the other quads (C > 3: rare - a flow / metric gradient through a feature splat)
        for (int c = 0; c < qn * 4; ++c) {
            float go[PX], yv[PX];
            load_go(c, go);
            load_y(c, yv);
#pragma unroll
            for (int k = 0; k < PX; ++k) sC[k] -= (go[k] * rd[k]) * (yv[k] * 0.5f + 0.5f);
        }
#pragma unroll
        for (int k = 0; k < PX; ++k) o[slot][k] = hole[k] ? 0.f : sC[k];
    }
#pragma unroll
    for (int k = 0; k < PX; ++k) grow[x + 1 + k] = make_float4(o[0][k], o[1][k], o[2][k], o[3][k]);
}

// One thread per source pixel of a row (grid = column blocks x rows x samples: no index division).  CT = 3 (the image splats)
// unrolls the channel loop completely, CT = 0 walks the channel quads.
template <int CT>
__global__ void __launch_bounds__(128) splat_bwd_gather_kernel(View4 in, View4 flow, View4 metric, const float4* __restrict__ G,
                                                               float* __restrict__ gin, float* __restrict__ gflow,
                                                               float* __restrict__ gmetric, SplatGeom g, int QB, int need_c) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.W) return;
    const int y = blockIdx.y, n = blockIdx.z;
    const int C = CT ? CT : g.C;
    const long long HW = (long long)g.H * g.W;
    const long long pix = (long long)y * g.W + x;
    const bool pre = g.mode == FLDR_SPLAT_SOFTMAX;
    const float* fp = flow.p + n * flow.sn + y * flow.sh + x * flow.sw;
    const float u = __ldcs(fp), v = __ldcs(fp + flow.sc);
    float z = 0.f;
    if (g.has_metric) z = __ldcs(metric.p + n * metric.sn + y * metric.sh + x * metric.sw);
    // softSplat.py:68-83 / 114-125 (the same corner arithmetic as the forward)
    const float X = (float)x + u, Y = (float)y + v;
    const float fx0 = floorf(X), fy0 = floorf(Y);
    const bool live = fx0 >= -1.f && fx0 < (float)g.W && fy0 >= -1.f && fy0 < (float)g.H;     // false for NaN / inf
    float* ginp = gin ? gin + (long long)n * g.C * HW + pix : nullptr;
    if (!live) {
        if (ginp) for (int c = 0; c < C; ++c) ginp[(long long)c * HW] = 0.f;
        if (gflow) { gflow[(long long)n * 2 * HW + pix] = 0.f; gflow[(long long)n * 2 * HW + HW + pix] = 0.f; }
        if (gmetric) gmetric[(long long)n * HW + pix] = 0.f;
        return;
    }
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float wx1 = (fx0 + 1.f) - X, wx0 = X - fx0, wy1 = (fy0 + 1.f) - Y, wy0 = Y - fy0;
    const float w[4] = {wx1 * wy1, wx0 * wy1, wx1 * wy0, wx0 * wy0};                          // NW, NE, SW, SE
    const int P = g.W + 2;
    const size_t plane = (size_t)(g.H + 2) * P;
    const float4* gq = G + (size_t)n * QB * plane + (size_t)((y0 + 1) * P + (x0 + 1));        // NW corner in quad 0 (border: zeros)
    float m = 1.f;
    if (g.has_metric) m = (g.mode == FLDR_SPLAT_SOFTMAX) ? exp_splat(z) : (g.mode == FLDR_SPLAT_LINEAR ? z : 1.f);
    const bool want_f = gflow != nullptr, want_m = gmetric != nullptr;
    float gfx = 0.f, gfy = 0.f, sum_gx = 0.f, gAC = 0.f;
    const float* ip = in.p + n * in.sn + y * in.sh + x * in.sw;
    const int nq = (C + 3) >> 2;
    auto quad = [&](int q, const float4& a, const float4& b, const float4& c4, const float4& d) {
        const float gs[4][4] = {{a.x, b.x, c4.x, d.x}, {a.y, b.y, c4.y, d.y}, {a.z, b.z, c4.z, d.z}, {a.w, b.w, c4.w, d.w}};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = q * 4 + j;
            if (c < C) {
                const float gA = gs[j][0] * w[0] + gs[j][1] * w[1] + gs[j][2] * w[2] + gs[j][3] * w[3];
                if (ginp) {
                    float gx = gA;
                    if (g.mode == FLDR_SPLAT_SOFTMAX) gx = gA * m * 0.5f;
                    else if (g.mode == FLDR_SPLAT_LINEAR) gx = gA * m;
                    __stcs(ginp + (long long)c * HW, gx);
                }
                if (want_f || want_m) {
                    const float xv = __ldcs(ip + c * in.sc);
                    const float xt = pre ? (xv + 1.f) * 0.5f : xv;
                    const float A = xt * m;
                    sum_gx += gA * xt;
                    gfx += A * ((gs[j][1] - gs[j][0]) * wy1 + (gs[j][3] - gs[j][2]) * wy0);
                    gfy += A * ((gs[j][2] - gs[j][0]) * wx1 + (gs[j][3] - gs[j][1]) * wx0);
                }
            } else if (need_c && c == C) {
                gAC = gs[j][0] * w[0] + gs[j][1] * w[1] + gs[j][2] * w[2] + gs[j][3] * w[3];
                gfx += m * ((gs[j][1] - gs[j][0]) * wy1 + (gs[j][3] - gs[j][2]) * wy0);
                gfy += m * ((gs[j][2] - gs[j][0]) * wx1 + (gs[j][3] - gs[j][1]) * wx0);
            }
        }
    };
    if (CT) {
        const float4 a = __ldg(gq), b = __ldg(gq + 1), c4 = __ldg(gq + P), d = __ldg(gq + P + 1);
        quad(0, a, b, c4, d);
    } else {
        const int nql = need_c ? QB : nq;          // the quad holding gS_C may be one past the last channel quad
#pragma unroll 2
        for (int q = 0; q < nql; ++q) {
            const float4* p4 = gq + (size_t)q * plane;
            const float4 a = __ldg(p4), b = __ldg(p4 + 1), c4 = __ldg(p4 + P), d = __ldg(p4 + P + 1);
            quad(q, a, b, c4, d);
        }
    }
    if (gflow) {
        __stcs(gflow + (long long)n * 2 * HW + pix, gfx);
        __stcs(gflow + (long long)n * 2 * HW + HW + pix, gfy);
    }
    if (gmetric) {
        float gz = sum_gx + gAC;
        if (g.mode == FLDR_SPLAT_SOFTMAX) gz *= m;
        __stcs(gmetric + (long long)n * HW + pix, gz);
    }
}

// The image splats of a training step (C = 3, softmax, contiguous tensors, W % 4 == 0, input + flow (+ metric) gradients): PX
// source pixels per thread, every streamed tensor read and written with vector accesses, all corner loads of the thread's
// pixels in flight together, no stride arithmetic and no mode branches (150 instead of 340 instructions per pixel).
template <int PX> struct VecF;
template <> struct VecF<4> {
    static __device__ __forceinline__ void ld(const float* p, float* v) { const float4 t = __ldcs(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void st(float* p, const float* v) { __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3])); }
};
template <> struct VecF<2> {
    static __device__ __forceinline__ void ld(const float* p, float* v) { const float2 t = __ldcs(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y; }
    static __device__ __forceinline__ void st(float* p, const float* v) { __stcs(reinterpret_cast<float2*>(p), make_float2(v[0], v[1])); }
};
template <bool METRIC, int PX, int MINB>
__global__ void __launch_bounds__(128, MINB) splat_bwd_gather_rgb_kernel(const float* __restrict__ in, const float* __restrict__ flow,
                                                                         const float* __restrict__ metric, const float4* __restrict__ G,
                                                                         float* __restrict__ gin, float* __restrict__ gflow,
                                                                         float* __restrict__ gmetric, int H, int W) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * PX;
    if (x4 >= W) return;
    const int y = blockIdx.y, n = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)y * W + x4;
    const int P = W + 2;
    const float4* Gn = G + (size_t)n * (H + 2) * P;
    float u[PX], v[PX], z[PX], xin[3][PX];
    VecF<PX>::ld(flow + (size_t)n * 2 * HW + pix, u);
    VecF<PX>::ld(flow + (size_t)n * 2 * HW + HW + pix, v);
    if (METRIC) VecF<PX>::ld(metric + (size_t)n * HW + pix, z);
#pragma unroll
    for (int j = 0; j < 3; ++j) VecF<PX>::ld(in + (size_t)n * 3 * HW + (size_t)j * HW + pix, xin[j]);
    const float Wf = (float)W, Hf = (float)H, yf = (float)y;
    float4 c[PX][4];                // [pixel][NW, NE, SW, SE]
    float wx0[PX], wx1[PX], wy0[PX], wy1[PX];
    bool live[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        // softSplat.py:68-83 / 114-125 (the same corner arithmetic as the forward)
        const float X = (float)(x4 + k) + u[k], Y = yf + v[k];
        const float fx0 = floorf(X), fy0 = floorf(Y);
        live[k] = fx0 >= -1.f && fx0 < Wf && fy0 >= -1.f && fy0 < Hf;                          // false for NaN / inf
        wx1[k] = (fx0 + 1.f) - X; wx0[k] = X - fx0; wy1[k] = (fy0 + 1.f) - Y; wy0[k] = Y - fy0;
        const int cell = live[k] ? ((int)fy0 + 1) * P + ((int)fx0 + 1) : 0;                    // zero border: no bounds tests
        const float4* gq = Gn + cell;
        c[k][0] = __ldg(gq); c[k][1] = __ldg(gq + 1); c[k][2] = __ldg(gq + P); c[k][3] = __ldg(gq + P + 1);
    }
    float gi[3][PX], gfx[PX], gfy[PX], gz[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        const float m = METRIC ? exp_splat(z[k]) : 1.f;
        const float w0 = wx1[k] * wy1[k], w1 = wx0[k] * wy1[k], w2 = wx1[k] * wy0[k], w3 = wx0[k] * wy0[k];   // NW, NE, SW, SE
        const float gs[4][4] = {{c[k][0].x, c[k][1].x, c[k][2].x, c[k][3].x}, {c[k][0].y, c[k][1].y, c[k][2].y, c[k][3].y},
                                {c[k][0].z, c[k][1].z, c[k][2].z, c[k][3].z}, {c[k][0].w, c[k][1].w, c[k][2].w, c[k][3].w}};
        float fx = 0.f, fy = 0.f, sg = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float gA = gs[j][0] * w0 + gs[j][1] * w1 + gs[j][2] * w2 + gs[j][3] * w3;
            const float xt = (xin[j][k] + 1.f) * 0.5f;
            const float A = xt * m;
            gi[j][k] = live[k] ? gA * m * 0.5f : 0.f;
            sg += gA * xt;
            fx += A * ((gs[j][1] - gs[j][0]) * wy1[k] + (gs[j][3] - gs[j][2]) * wy0[k]);
            fy += A * ((gs[j][2] - gs[j][0]) * wx1[k] + (gs[j][3] - gs[j][1]) * wx0[k]);
        }
        const float gAC = gs[3][0] * w0 + gs[3][1] * w1 + gs[3][2] * w2 + gs[3][3] * w3;
        fx += m * ((gs[3][1] - gs[3][0]) * wy1[k] + (gs[3][3] - gs[3][2]) * wy0[k]);
        fy += m * ((gs[3][2] - gs[3][0]) * wx1[k] + (gs[3][3] - gs[3][1]) * wx0[k]);
        gfx[k] = live[k] ? fx : 0.f;
        gfy[k] = live[k] ? fy : 0.f;
        gz[k] = live[k] ? (sg + gAC) * m : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) VecF<PX>::st(gin + (size_t)n * 3 * HW + (size_t)j * HW + pix, gi[j]);
    VecF<PX>::st(gflow + (size_t)n * 2 * HW + pix, gfx);
    VecF<PX>::st(gflow + (size_t)n * 2 * HW + HW + pix, gfy);
    if (METRIC) VecF<PX>::st(gmetric + (size_t)n * HW + pix, gz);
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

struct FwdPlan {
    SplatGeom g;
    int Q;
    size_t acc_bytes;       // accumulator [N][Q][H][W + 2] float4
};

static int plan_fwd(int mode, int N, int C, int H, int W, bool has_metric, FwdPlan& p) {
    int st = make_geom(mode, N, C, H, W, has_metric, p.g);
    if (st != FLDR_OK) return st;
    p.Q = p.g.CP / 4;
    if ((long long)H * (W + 2) >= (1ll << 30)) return FLDR_ERR_UNSUPPORTED;       // 32-bit cell offsets inside a plane
    p.acc_bytes = align_up((size_t)N * p.Q * H * (W + 2) * 16, 256);
    return FLDR_OK;
}

}  // namespace fldr

using namespace fldr;

extern "C" int fldr_splat_set_nonfinite_flag(unsigned int* device_flag) {
    cudaError_t e = cudaMemcpyToSymbol(g_nonfinite_flag, &device_flag, sizeof(device_flag));
    if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
    return FLDR_OK;
}

extern "C" size_t fldr_splat_fwd_workspace_bytes(int mode, int N, int C, int H, int W) {
    FwdPlan p;
    if (plan_fwd(mode, N, C, H, W, true, p) != FLDR_OK) return 0;
    return p.acc_bytes;
}

static int launch_normalise(const FwdPlan& p, float* acc, float* out, float* norm, cudaStream_t s) {
    const SplatGeom& g = p.g;
    const int N = g.N, Q = p.Q;
    const bool px4 = (g.W % 4 == 0) && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                     (!norm || (reinterpret_cast<uintptr_t>(norm) & 15) == 0);
    if (px4 && Q == 1 && g.CA > g.C && g.mode != FLDR_SPLAT_RAW && get_option(kOptSplatSnake) != 0) {
        dim3 grid((unsigned)((g.W / 4 + 127) / 128), g.H, N);
        splat_normalise_kernel<4, true><<<grid, 128, 0, s>>>(acc, out, norm, g, Q);
    } else if (px4) {
        dim3 grid((unsigned)((g.W / 4 + 127) / 128), g.H, N * Q);
        splat_normalise_kernel<4><<<grid, 128, 0, s>>>(acc, out, norm, g, Q);
    } else {
        dim3 grid((unsigned)((g.W + 127) / 128), g.H, N * Q);
        splat_normalise_kernel<1><<<grid, 128, 0, s>>>(acc, out, norm, g, Q);
    }
    return check_launch();
}

// Can the TMA tile kernel take these views?  (16-byte aligned rows / planes, unit pixel stride, at most 6 staged planes)
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool view_tma_ok(const View4& v) {
    return v.sw == 1 && v.sh > 0 && v.sc > 0 && v.sn > 0 && (v.sh % 4) == 0 && (v.sc % 4) == 0 && (v.sn % 4) == 0 && aligned16(v.p);
}
// [N][C][H][W] view -> 4-D tensor map with an [nbox][8][128] box
static bool encode_view(CUtensorMap* map, const View4& v, int N, int C, int H, int W, int nbox, int tile_rows) {
    const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)v.sh * 4, (uint64_t)v.sc * 4, (uint64_t)v.sn * 4};
    const uint32_t box[4] = {(uint32_t)tile::TW, (uint32_t)tile_rows, (uint32_t)nbox, 1u};
    return encode_tensor_map_4d(map, v.p, dims, strides, box);
}

// zero + scatter + normalise
static int launch_forward(const FwdPlan& p, const View4& vin, const View4& vfl, const View4& vme, float* acc, float* out,
                          float* norm, cudaStream_t s) {
    const SplatGeom& g = p.g;
    const int N = g.N, H = g.H, W = g.W, Q = p.Q;
    int st;
    if ((long long)N * Q > 65535 || H > 65535) return FLDR_ERR_UNSUPPORTED;
    const long long n4 = (long long)N * Q * H * W;
    const int wkind = !g.has_metric ? 0 : (g.mode == FLDR_SPLAT_SOFTMAX ? 1 : 2);
    const bool pre = g.mode == FLDR_SPLAT_SOFTMAX;
    if (n4 <= (long long)get_option(kOptSplatFusedMax) && get_option(kOptSplatFusedMax) > 0) {
        // small frame: single cooperative launch (zero / scatter / normalise separated by grid barriers)
        const void* fn = wkind == 1 ? (const void*)splat_fused_small_kernel<4, 1, true>
                       : wkind == 2 ? (const void*)splat_fused_small_kernel<4, 2, false>
                       : pre        ? (const void*)splat_fused_small_kernel<4, 0, true>
                                    : (const void*)splat_fused_small_kernel<4, 0, false>;
        static int per_sm[64][4];                      // occupancy per device and instantiation, computed once
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) dev = 0;
        const int slot = wkind == 1 ? 0 : wkind == 2 ? 1 : pre ? 2 : 3;
        int nb = __atomic_load_n(&per_sm[dev][slot], __ATOMIC_ACQUIRE);
        if (nb == 0) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, 256, 0) != cudaSuccess || nb < 1) nb = 1;
            __atomic_store_n(&per_sm[dev][slot], nb, __ATOMIC_RELEASE);
        }
        const long long units = (long long)N * Q * ((H + 3) / 4) * ((W + 31) / 32);     // warps wanted in phase B
        long long blocks = (units + 7) / 8;
        const long long blocks_c = (n4 + 255) / 256;
        if (blocks < blocks_c) blocks = blocks_c;
        const long long cap = (long long)sm_count() * nb;
        if (blocks > cap) blocks = cap;
        SplatGeom gg = g;
        int QQ = Q;
        View4 a0 = vin, a1 = vfl, a2 = vme;
        void* args[] = {&a0, &a1, &a2, &acc, &out, &norm, &gg, &QQ};
        cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(256), args, 0, s);
        if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
        return FLDR_OK;
    }
    // row order of the three passes: zero fill front to back, scatter back to front, normalise front to back - each pass starts on
    // the accumulator lines its predecessor touched last (845 -> 780 MB of DRAM traffic and 190 -> 182 us on the 4K image splat)
    const int snake = get_option(kOptSplatSnake) != 0;
    if (snake) {
        const long long cells = (long long)N * Q * H * (W + 2);
        splat_zero_kernel<<<(unsigned)((cells + 2047) / 2048), 256, 0, s>>>(reinterpret_cast<float4*>(acc), cells, cells - (long long)FLDR_ZERO_KEEP_MB * (1 << 16));
        if ((st = check_launch()) != FLDR_OK) return st;
    } else {
        cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)N * Q * H * (W + 2) * 16, s);
        if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
    }
    // quad shape known at compile time for the image splat (C = 3 + weight) and for all-full-quad inputs
    const int qs = (g.C == 3 && g.CA == 4) ? 1 : (g.CA == g.C && g.C % 4 == 0) ? 2 : 0;
    // accumulators larger than ~half the L2 are DRAM-resident when the reductions arrive: prefetch their lines
    const int pf_opt = get_option(kOptSplatPfRows);
    // (with the alternating row order the scatter starts on lines the zero fill left in L2 and the prefetch no longer pays: 175.6
    //  -> 173.6 us without it; it stays the default for the front-to-back order, where it was worth 133 -> 115 us)
    const int pf = ((size_t)n4 * 16 > (48u << 20)) ? (pf_opt > 0 ? pf_opt : (pf_opt < 0 || snake ? 0 : 4)) : 0;
    bool tiled = get_option(kOptSplatStream) != 0 && view_tma_ok(vin) && view_tma_ok(vfl) && (!g.has_metric || (view_tma_ok(vme) && g.C < 4));
    CUtensorMap tm_in, tm_fl, tm_me;
    const int nbox = (Q == 1) ? g.C : 4;               // channels per input box
    // tile height: 8 source rows; 4 for a single DRAM-sized image plane (more, shorter CTAs: 163 -> 160 us on the 4K frame; batches
    // and the L2-resident feature splats measured 2 - 4 % slower with 4 and keep 8)
    const int tr = (qs == 1 && N * Q == 1 && (size_t)n4 * 16 > (48u << 20)) ? 4 : tile::R;
    if (tiled) {
        tiled = encode_view(&tm_in, vin, N, g.C, H, W, nbox, tr) && encode_view(&tm_fl, vfl, N, 2, H, W, 2, tr);
        if (tiled && g.has_metric) tiled = encode_view(&tm_me, vme, N, 1, H, W, 1, tr);
        if (tiled && !g.has_metric) tm_me = tm_fl;
    }
    if (tiled) {
        dim3 grid((W + tile::TW - 1) / tile::TW, (H + tr - 1) / tr, N * Q);
#define FLDR_LAUNCH_TILE2(WK_, PRE_, QS_, TR_) \
    splat_scatter_tile_kernel<WK_, PRE_, QS_, TR_><<<grid, tile::TW, 0, s>>>(tm_in, tm_fl, tm_me, acc, g, Q, nbox, pf, snake)
#define FLDR_LAUNCH_TILE(WK_, PRE_)                                       \
    do {                                                                  \
        if (qs == 1 && tr == 4) FLDR_LAUNCH_TILE2(WK_, PRE_, 1, 4);       \
        else if (qs == 1) FLDR_LAUNCH_TILE2(WK_, PRE_, 1, tile::R);       \
        else if (qs == 2) FLDR_LAUNCH_TILE2(WK_, PRE_, 2, tile::R);       \
        else FLDR_LAUNCH_TILE2(WK_, PRE_, 0, tile::R);                    \
    } while (0)
        if (wkind == 1) FLDR_LAUNCH_TILE(1, true);
        else if (wkind == 2) FLDR_LAUNCH_TILE(2, false);
        else if (pre) FLDR_LAUNCH_TILE(0, true);
        else FLDR_LAUNCH_TILE(0, false);
#undef FLDR_LAUNCH_TILE2
#undef FLDR_LAUNCH_TILE
    } else {
        // rows walked per thread: 16 merges best vertically, but the walk is serial - shorten it until the grid offers
        // at least two waves of CTAs (8 x 128 threads per SM)
        const int bx = W >= 128 ? 128 : ((W + 31) / 32) * 32;
        const long long per_row_ctas = (long long)((W + bx - 1) / bx) * N * Q;
        const long long two_waves = 2ll * sm_count() * 8;
        int R = 16;
        while (R > 4 && per_row_ctas * ((H + R - 1) / R) < two_waves) R >>= 1;
        dim3 grid((W + bx - 1) / bx, (H + R - 1) / R, N * Q);
#define FLDR_LAUNCH_SCATTER2(WK_, PRE_, QS_) \
    splat_scatter_merged_kernel<WK_, PRE_, QS_><<<grid, bx, 0, s>>>(vin, vfl, vme, acc, g, Q, pf, R, snake)
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
    return launch_normalise(p, acc, out, norm, s);
}

extern "C" int fldr_splat_fwd(int mode, const float* in, const int64_t* in_strides, const float* flow,
                              const int64_t* flow_strides, const float* metric, const int64_t* metric_strides,
                              float* out, float* norm, int N, int C, int H, int W, void* ws, size_t ws_bytes,
                              fldr_stream_t stream) {
    FwdPlan p;
    int st = plan_fwd(mode, N, C, H, W, metric != nullptr, p);
    if (st != FLDR_OK) return st;
    if (!in || !in_strides || !flow || !flow_strides || !out || (metric && !metric_strides)) return FLDR_ERR_INVALID_ARGUMENT;
    if (!ws || ws_bytes < p.acc_bytes) return FLDR_ERR_WORKSPACE_TOO_SMALL;
    if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return FLDR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const View4 vin = make_view(in, in_strides), vfl = make_view(flow, flow_strides), vme = make_view(metric, metric_strides);
    if (!mode_has_norm(mode)) norm = nullptr;
    return launch_forward(p, vin, vfl, vme, static_cast<float*>(ws), out, norm, s);
}

extern "C" size_t fldr_splat_bwd_workspace_bytes(int mode, int N, int C, int H, int W) {
    // gS in the forward's quad layout with a zero border: [N][ceil((C + 1) / 4)][H + 2][W + 2] float4
    SplatGeom g;
    if (make_geom(mode, N, C, H, W, true, g) != FLDR_OK) return 0;
    if ((long long)(H + 2) * (W + 2) >= (1ll << 30)) return 0;
    return align_up((size_t)N * ((C + 4) / 4) * (H + 2) * (W + 2) * 16, 256);
}

extern "C" int fldr_splat_bwd(int mode, const float* in, const int64_t* in_strides, const float* flow,
                              const int64_t* flow_strides, const float* metric, const int64_t* metric_strides,
                              const float* out, const float* norm, const float* grad_out,
                              const int64_t* grad_out_strides, float* grad_in, float* grad_flow, float* grad_metric,
                              int N, int C, int H, int W, void* ws, size_t ws_bytes, fldr_stream_t stream) {
    SplatGeom g;
    int st = make_geom(mode, N, C, H, W, metric != nullptr, g);
    if (st != FLDR_OK) return st;
    if (!in || !in_strides || !flow || !flow_strides || !grad_out || !grad_out_strides || (metric && !metric_strides))
        return FLDR_ERR_INVALID_ARGUMENT;
    if (mode_has_norm(mode) && (!out || !norm)) return FLDR_ERR_INVALID_ARGUMENT;
    if (grad_metric && !g.has_metric) return FLDR_ERR_INVALID_ARGUMENT;
    if (!grad_in && !grad_flow && !grad_metric) return FLDR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (H + 2 > 65535 || (long long)(H + 2) * (W + 2) >= (1ll << 30)) return FLDR_ERR_UNSUPPORTED;
    // gS_C (the normaliser's gradient) only feeds the flow and metric gradients
    const int need_c = (mode_has_norm(mode) && (grad_flow || grad_metric)) ? 1 : 0;
    const int QB = (C + need_c + 3) / 4;
    if ((long long)N * QB > 65535) return FLDR_ERR_UNSUPPORTED;
    const size_t need_bytes = (size_t)N * QB * (H + 2) * (W + 2) * 16;
    if (!ws || ws_bytes < need_bytes) return FLDR_ERR_WORKSPACE_TOO_SMALL;
    if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return FLDR_ERR_INVALID_ARGUMENT;
    const View4 vin = make_view(in, in_strides), vfl = make_view(flow, flow_strides), vme = make_view(metric, metric_strides),
                vgo = make_view(grad_out, grad_out_strides);
    float4* G = static_cast<float4*>(ws);
    {
        const bool vec = (W % 4 == 0) && vgo.sw == 1 && (vgo.sh % 4) == 0 && (vgo.sc % 4) == 0 && (vgo.sn % 4) == 0 && aligned16(vgo.p) &&
                         (!out || aligned16(out)) && (!norm || aligned16(norm));
        const long long cells = (long long)(vec ? (W + 2 + 3) / 4 : W + 2) * (H + 2);      // (row, pixel group) pairs of a plane
        const int bx = bwdp::NT;
        const dim3 grid((unsigned)((cells + bx - 1) / bx), 1, N * QB);
        if (vec) splat_bwd_prep_kernel<true, 8><<<grid, bx, 0, s>>>(vgo, out, norm, G, g, QB, need_c);
        else splat_bwd_prep_kernel<false, 4><<<grid, bx, 0, s>>>(vgo, out, norm, G, g, QB, need_c);
        if ((st = check_launch()) != FLDR_OK) return st;
    }
    if (N > 65535) return FLDR_ERR_UNSUPPORTED;
    const int bx = W >= 128 ? 128 : ((W + 31) / 32) * 32;          // narrow frames: no idle lanes beyond the row
    const dim3 grid((W + bx - 1) / bx, H, N);
    auto dense = [&](const View4& v, int Cv) { return v.sw == 1 && v.sh == W && v.sc == (long long)H * W && v.sn == (long long)Cv * H * W && aligned16(v.p); };
    const bool rgb_fast = C == 3 && mode == FLDR_SPLAT_SOFTMAX && (W % 4) == 0 && grad_in && grad_flow && (grad_metric != nullptr) == (g.has_metric != 0) &&
                          dense(vin, 3) && dense(vfl, 2) && (!g.has_metric || dense(vme, 1)) && aligned16(grad_in) && aligned16(grad_flow) &&
                          (!grad_metric || aligned16(grad_metric));
    if (rgb_fast) {
        // two pixels per thread, 8 CTAs per SM (64 registers): measured 185 us against 212 us with four pixels per thread at 4 CTAs per SM
        // (the kernel is latency-bound: flow load -> corner loads -> stores; more warps in flight beat wider accesses)
        const int groups = W / 2;
        const int bxf = groups >= 128 ? 128 : ((groups + 31) / 32) * 32;
        const dim3 gridf((groups + bxf - 1) / bxf, H, N);
        if (g.has_metric) splat_bwd_gather_rgb_kernel<true, 2, 8><<<gridf, bxf, 0, s>>>(in, flow, metric, G, grad_in, grad_flow, grad_metric, H, W);
        else splat_bwd_gather_rgb_kernel<false, 2, 8><<<gridf, bxf, 0, s>>>(in, flow, nullptr, G, grad_in, grad_flow, nullptr, H, W);
    } else if (C == 3) splat_bwd_gather_kernel<3><<<grid, bx, 0, s>>>(vin, vfl, vme, G, grad_in, grad_flow, grad_metric, g, QB, need_c);
    else splat_bwd_gather_kernel<0><<<grid, bx, 0, s>>>(vin, vfl, vme, G, grad_in, grad_flow, grad_metric, g, QB, need_c);
    return check_launch();
}
