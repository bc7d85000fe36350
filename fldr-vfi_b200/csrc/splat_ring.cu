// Forward splat, streaming path: ONE launch does zero-init, scatter, normalise, post-scale and the NCHW store, and the
// accumulator never leaves L2.  Replaces softSplat.py:12-52 (kernel_Softsplat_updateOutput) + the torch glue of
// softSplat.py:320-352 for every frame that is too large for the single cooperative small-frame launch.
//
// Why: with a whole-frame accumulator the 4K image splat moves ~850 MB through DRAM for 340 MB of algorithmic traffic
// (zero fill 151 W, scatter 226 R + accumulator lines fetched and written back, normalise 151 R + 113 W; ncu,
// profiles/r1_ncu_full_top_kernels.txt).  Here the accumulator is a ring of rows (a power of two, <= ~32 MiB, L2-resident)
// and the frame streams through it: DRAM sees the inputs once and the output once.
//
// Layout
//   ring        [Q][ring_rows][W + 2] float4 cells; cell 0 and W+1 of a row are guard cells (targets one column outside
//               the frame land there and are never read), so the reductions need no x test.  Row y of sample n lives in
//               ring row (n * NS * 8 + y) & (ring_rows - 1).
//   work items  Z(slot, t)   zero ring strip `slot`, columns of tile t                     (all Z first)
//               S(J, t, q)   scatter source strip J (8 rows x 128 columns, channel quad q) into the ring
//               N(J, t)      read ring strip J (all quads), normalise + post-scale, store NCHW, zero the cells again
//               handed out by one atomic ticket counter in the fixed order  Z..., then for g = 0, 1, ...: all S(g, ., .)
//               followed by all N(g - D2, .)   (D2 = reach + lag strips).
//   dependencies  per-strip completion counters in global memory:
//               S(J) needs the ring slots of strips J-Ds .. J+Ds cleaned for their epoch; N(J) needs S(J-Ds .. J+Ds) complete.
//               An item only waits for items EARLIER in ticket order, and a CTA works through its tickets in order, so the
//               lowest unfinished ticket can always run: no deadlock, no co-residency requirement.
//               Ds bounds the vertical reach: a source whose target row leaves strips J-Ds .. J+Ds raises a device flag and
//               the guarded whole-frame launches that follow re-do the call (no host sync).  When the ring holds every
//               strip of the batch the reach is unbounded.
//
// CTA = 4 consumer warps + 2 helper warps (warp-specialised, mbarrier pipeline):
//   producer    draws tickets, decodes them, and stages the inputs of S items in shared memory with cp.async.bulk
//               (one 512-byte row segment per copy, L2 evict-first) - up to NST items ahead of the consumers, so the
//               serial row walk below never waits for DRAM;
//   gatekeeper  polls the dependencies of the next item (ld.acquire.gpu) while the consumers are still busy with the
//               current one, then opens the item's `full` barrier;
//   consumers   S: a thread owns one column and walks the strip's rows; the bottom-W corner is carried in registers and
//               merged with the next row's top-W corner, E corners travel to the lane on the right by shuffle and merge
//               with its W corners, so a locally smooth flow costs ~1 red.global.add.v4.f32 per source pixel instead of 4
//               (the reference: 4 scalar REDs per ELEMENT, softSplat.py:39-50).  Corners are identified by their ring
//               cell offset (negative = not in frame): two corners merge iff they are the same memory cell.
//               N: 4 pixels per thread, 128-bit ring loads (ld.global.cg: the reductions happen at L2) and stores;
//   signaller   after the consumers are done with an item: one gpu-scope fence, then the item's completion counter
//               (the cooperative-groups grid.sync pattern: CTA-scope barrier, then one thread fences and signals).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "splat.cuh"
#include "tma.cuh"

namespace fldr {
namespace ring {
constexpr int R = 8;                          // rows per strip
constexpr int TW = 128;                       // columns per tile = consumer threads
constexpr int NST = 2;                        // staged S inputs per CTA
#ifndef FLDR_RING_NSLOT
#define FLDR_RING_NSLOT 4
#endif
constexpr int NSLOT = FLDR_RING_NSLOT;        // item descriptors in flight per CTA (>= NST)
constexpr int PLANES = 6;                     // staged planes per S item: flow 2 + metric 0/1 + channels of the quad
constexpr int STAGE_FLOATS = PLANES * R * TW; // 24 KiB
constexpr int kConsumerWarps = 4;
constexpr int kProducerWarp = 4, kSignalWarp = 5;
constexpr int kThreads = (kConsumerWarps + 2) * 32;
#ifndef FLDR_RING_CTAS
#define FLDR_RING_CTAS 4
#endif
constexpr int kCtasPerSm = FLDR_RING_CTAS;    // resident CTAs per SM the kernel is compiled for (register cap)
constexpr int kSent = -(1 << 30);             // "no cell": stays negative after + 1
enum Kind { kExit = 0, kZero = 1, kScatter = 2, kNorm = 3 };

struct ItemDesc {
    int kind;
    int J;              // absolute strip (S / N) or ring slot (Z)
    int n, q;
    int x0, cols;       // first column of the tile, columns in it (multiple of 4)
    int yb, rows;       // first row inside the sample, rows in the strip
    int row0;           // ring row of the sample's row 0
    int dep_lo, dep_hi; // absolute strips whose counters gate this item
    float ylo, yhi;     // S: floor(Y) must lie in [ylo, yhi)
    int pad[3];
};
static_assert(sizeof(ItemDesc) == 64, "descriptor size");

constexpr size_t kSmemBytes = (size_t)NST * STAGE_FLOATS * 4 + NSLOT * sizeof(ItemDesc) + 4 * NSLOT * 8 + NSLOT * 4 + 16;
}  // namespace ring

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const unsigned* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// strips whose dependency window [J - Ds, min(J + Ds, end of J's sample)] lies below the frontier f (a count of leading
// complete strips / clean events): monotone in f, so tickets below it are claimable in order
__device__ __forceinline__ int frontier_limit(int f, const RingGeom& rg) {
    const int a = f - rg.Ds, b = (f / rg.NS) * rg.NS;
    const int m = a > b ? a : b;
    return m < 0 ? 0 : (m > rg.NT ? rg.NT : m);
}
// Advance the two frontiers (leading complete strips of the scatter counters / leading complete clean events) from the
// per-strip counters.  Warp-wide; any warp may call it at any time: a frontier only moves forward (atomicMax).
__device__ __forceinline__ bool advance_frontiers(unsigned* __restrict__ ctrl, const unsigned* __restrict__ sdone,
                                                  const unsigned* __restrict__ ev, const RingGeom& rg, int sf, int cf, int lane) {
    int f = sf;
    while (f < rg.NT) {
        const unsigned v = (f + lane < rg.NT) ? ld_relaxed_u32(&sdone[f + lane]) : 0u;
        const unsigned m = __ballot_sync(0xffffffffu, v >= (unsigned)rg.TQ);
        const int cnt = (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);
        f += cnt;
        if (cnt < 32) break;
    }
    const int nev = rg.nZs + rg.NT;
    int e = cf;
    while (e < nev) {
        const unsigned v = (e + lane < nev) ? ld_relaxed_u32(&ev[e + lane]) : 0u;
        const unsigned m = __ballot_sync(0xffffffffu, v >= (unsigned)rg.T);
        const int cnt = (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);
        e += cnt;
        if (cnt < 32) break;
    }
    if (lane == 0) {
        if (f > sf) atomicMax(&ctrl[kRingCtrlClaim + 2], (unsigned)f);
        if (e > cf) atomicMax(&ctrl[kRingCtrlClaim + 3], (unsigned)e);
    }
    return f > sf || e > cf;
}
__device__ __forceinline__ void red_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_cg4(const float4* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_cg4_zero(float4* p) {
    asm volatile("st.global.cg.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(0.f) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void bulk_load_plain(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// mbarrier wait with a last-resort watchdog: a protocol bug must abort the launch (sticky CUDA error), not hang the device
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
    for (int i = 0; i < (1 << 24); ++i)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// one row segment global -> shared, completion counted on `bar`; the inputs are read exactly once: evict-first in L2
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
// one [planes][8 rows][128 columns] box of an NCHW input, global -> shared; rows / columns / channels outside the tensor are
// zero-filled by the TMA unit.  The inputs are read exactly once: evict-first in L2.
__device__ __forceinline__ void tma_box_load(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int c, int n, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(c), "r"(n), "l"(pol)
        : "memory");
}
// cell `off` of plane `base` += v, only when off >= 0 (one predicated REDG, no branch)
__device__ __forceinline__ void red4_at(float4* base, int off, const float* v) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u64 a;\n\t"
        "setp.ge.s32 p, %1, 0;\n\t"
        "mad.wide.s32 a, %1, 16, %0;\n\t"
        "@p red.global.add.v4.f32 [a], {%2, %3, %4, %5};\n\t}"
        ::"l"(base), "r"(off), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]));
}
// e^z to ~2 ulp for any z: ex2.approx of the rounded product, corrected by the product's rounding error and the
// representation error of log2(e) (the plain ex2(z * log2e) loses |z| * 6e-8 relative)
__device__ __forceinline__ float exp_splat(float z) {
    const float l2e = 1.4426950408889634f;
    const float t = z * l2e;
    const float e = fmaf(z, l2e, -t) + z * 1.92596299e-8f;
    float p;
    asm("ex2.approx.f32 %0, %1;" : "=f"(p) : "f"(t));
    return fmaf(p, e * 0.6931471805599453f, p);
}

// ---- S item: scatter 8 rows x 128 columns of one channel quad from the staged inputs into the ring --------------------
// WKIND: 0 = weight 1, 1 = exp(z) (softmax), 2 = z (linear); PRE: (x+1)/2 pre-scale (softSplat.py:334);
// QS: 1 = image quad (3 channels + weight), 2 = four channels, 0 = run-time shape.
template <int WKIND, bool PRE, int QS, bool BOUNDED>
__device__ __forceinline__ bool ring_scatter_item(const float* __restrict__ st, const ring::ItemDesc& d, const SplatGeom& g,
                                                  const RingGeom& rg, float4* __restrict__ ringbuf, int tid, int lane) {
    using namespace ring;
    const int nch = QS == 1 ? 3 : QS == 2 ? 4 : min(4, g.C - d.q * 4);
    const int wslot = QS == 1 ? 3 : QS == 2 ? -1 : ((g.CA > g.C) ? g.C - d.q * 4 : -1);
    constexpr int CH0 = 2 + (WKIND ? 1 : 0);
    float4* rq = ringbuf + (size_t)d.q * rg.ring_rows * rg.pitch;
    const int mask = rg.mask, pitch = rg.pitch, row0 = d.row0;
    const bool inb = tid < d.cols;
    const float Wf = (float)g.W, Hf = (float)g.H, Hm1 = (float)(g.H - 1);
    const float ylo = BOUNDED ? d.ylo : -1.f, yhi = BOUNDED ? d.yhi : Hf;
    const float xf = (float)(d.x0 + tid);
    float yf = (float)d.yb;
    const int src_lane = (lane + 31) & 31;
    const bool lane_gt0 = lane > 0;
    int prev_t = kSent;
    float pw[4] = {0.f, 0.f, 0.f, 0.f};
    bool ovf = false;
    const float* sp = st + tid;
#pragma unroll 2
    for (int r = 0; r < d.rows; ++r, yf += 1.f, sp += TW) {
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
        const bool pX = inb && fx0 >= -1.f && fx0 < Wf;          // false for NaN / inf (the reference device-asserts, 25-26)
        const bool pR = pX && fy0 >= ylo && fy0 < yhi;           // both target rows inside the frame+1 ring reach
        if (BOUNDED) ovf = ovf || (pX && fy0 >= -1.f && fy0 < Hf && !pR);
        const bool pT = pR && fy0 >= 0.f;                        // top corners in frame
        const bool pB = pR && fy0 < Hm1;                         // bottom corners in frame
        const int cx = (int)x1f;                                 // x0 + 1 = cell index of the W corners (guard cell at 0)
        const int rt = ((int)fy0 + row0) & mask;
        const int rb = (rt + 1) & mask;
        const int tT = pT ? rt * pitch + cx : kSent;
        const int tB = pB ? rb * pitch + cx : kSent;
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
        // vertical: the bottom-W corner carried from the previous row joins this row's top-W corner, or is flushed
        const bool vm = tT == prev_t;
        red4_at(rq, vm ? kSent : prev_t, pw);
#pragma unroll
        for (int j = 0; j < 4; ++j) tW[j] = vm ? tW[j] + pw[j] : tW[j];
        // horizontal: the E corners travel to the lane on the right (rotate: lane 0 gets lane 31's and only forwards them)
        const int rT = __shfl_sync(0xffffffffu, tT + 1, src_lane);
        const int rB = __shfl_sync(0xffffffffu, tB + 1, src_lane);
        float rt4[4], rb4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            rt4[j] = __shfl_sync(0xffffffffu, tE[j], src_lane);
            rb4[j] = __shfl_sync(0xffffffffu, bE[j], src_lane);
        }
        const bool takeT = lane_gt0 && rT == tT, takeB = lane_gt0 && rB == tB;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tW[j] = takeT ? tW[j] + rt4[j] : tW[j];
            bW[j] = takeB ? bW[j] + rb4[j] : bW[j];
        }
        red4_at(rq, takeT ? kSent : rT, rt4);
        red4_at(rq, takeB ? kSent : rB, rb4);
        red4_at(rq, tT, tW);
        prev_t = tB;
#pragma unroll
        for (int j = 0; j < 4; ++j) pw[j] = bW[j];
    }
    red4_at(rq, prev_t, pw);
    return ovf;
}

// ---- N item: ring strip -> normalise, post-scale (softSplat.py:343-349), NCHW store, cells back to zero ----------------
__device__ __forceinline__ void ring_normalise_item(const ring::ItemDesc& d, const SplatGeom& g, const RingGeom& rg,
                                                    float4* __restrict__ ringbuf, float* __restrict__ out,
                                                    float* __restrict__ norm_out, int warp, int lane) {
    using namespace ring;
    const int xl = lane * 4;
    if (xl >= d.cols) return;
    const long long HW = (long long)g.H * g.W;
    const bool has_norm = g.CA > g.C, raw = g.mode == FLDR_SPLAT_RAW;
    const int qn = g.C >> 2, slot = g.C & 3;
    const size_t plane = (size_t)rg.ring_rows * rg.pitch;
    const int Q = rg.Q;
    for (int r = warp; r < d.rows; r += kConsumerWarps) {
        const int y = d.yb + r;
        const size_t ro = (size_t)((d.row0 + y) & rg.mask) * rg.pitch + 1 + d.x0 + xl;
        float dv[4] = {1.f, 1.f, 1.f, 1.f};
        float4 n4[4];
        if (has_norm) {
            float nrm[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                n4[k] = ld_cg4(ringbuf + qn * plane + ro + k);
                nrm[k] = slot == 0 ? n4[k].x : slot == 1 ? n4[k].y : slot == 2 ? n4[k].z : n4[k].w;
                dv[k] = norm_recip(nrm[k]);
            }
            if (norm_out) __stcs(reinterpret_cast<float4*>(norm_out + (long long)d.n * HW + (long long)y * g.W + d.x0 + xl),
                                 make_float4(nrm[0], nrm[1], nrm[2], nrm[3]));
        }
        float* op = out + (long long)d.n * g.C * HW + (long long)y * g.W + d.x0 + xl;
        for (int q = 0; q < Q; ++q) {
            float4 s4[4];
            float4* cp = ringbuf + q * plane + ro;
            if (has_norm && q == qn) {
#pragma unroll
                for (int k = 0; k < 4; ++k) s4[k] = n4[k];
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) s4[k] = ld_cg4(cp + k);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) st_cg4_zero(cp + k);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = q * 4 + j;
                if (c < g.C) {
                    float yv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float sv = j == 0 ? s4[k].x : j == 1 ? s4[k].y : j == 2 ? s4[k].z : s4[k].w;
                        yv[k] = post_scale(sv, dv[k], raw, has_norm);
                    }
                    __stcs(reinterpret_cast<float4*>(op + (long long)c * HW), make_float4(yv[0], yv[1], yv[2], yv[3]));
                }
            }
        }
    }
}

// N item of a single-quad splat (the image splats): the gatekeeper has copied the strip's ring cells into the stage
// ([8 rows][128 cells] float4) once the dependencies were satisfied, so nothing here waits for L2.  A lane owns the cells
// lane, lane+32, lane+64, lane+96 of a row (conflict-free LDS.128, full-line stores).
__device__ __forceinline__ void ring_normalise_item_staged(const float4* __restrict__ st4, const ring::ItemDesc& d,
                                                           const SplatGeom& g, const RingGeom& rg, float4* __restrict__ ringbuf,
                                                           float* __restrict__ out, float* __restrict__ norm_out, int warp, int lane) {
    using namespace ring;
    const long long HW = (long long)g.H * g.W;
    const bool has_norm = g.CA > g.C, raw = g.mode == FLDR_SPLAT_RAW;
    const int slot = g.C & 3;                       // has_norm: the normaliser's lane inside the quad (C <= 3)
    for (int r = warp; r < d.rows; r += kConsumerWarps) {
        const int y = d.yb + r;
        float4* cp = ringbuf + (size_t)((d.row0 + y) & rg.mask) * rg.pitch + 1 + d.x0;
        float* op = out + (long long)d.n * g.C * HW + (long long)y * g.W + d.x0;
        float* np = norm_out ? norm_out + (long long)d.n * HW + (long long)y * g.W + d.x0 : nullptr;
        float4 s4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) s4[k] = st4[r * TW + lane + 32 * k];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xl = lane + 32 * k;
            if (xl < d.cols) {
                st_cg4_zero(cp + xl);
                float dv = 1.f;
                if (has_norm) {
                    const float nrm = slot == 0 ? s4[k].x : slot == 1 ? s4[k].y : slot == 2 ? s4[k].z : s4[k].w;
                    dv = norm_recip(nrm);
                    if (np) __stcs(np + xl, nrm);
                }
                const float sv[4] = {s4[k].x, s4[k].y, s4[k].z, s4[k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < g.C) __stcs(op + (long long)j * HW + xl, post_scale(sv[j], dv, raw, has_norm));
            }
        }
    }
}

#ifdef FLDR_RING_TRACE
// per-item event times (SM clock) of the first 8 CTAs: trace[blk][item][8] behind the counters (plan_ring reserves the space)
#define TRACE(ev_, val_) do { if (blockIdx.x < 8 && k < 128 && lane == 0) trace[((size_t)blockIdx.x * 128 + k) * 8 + (ev_)] = (val_); } while (0)
#define TRACE_T(ev_) TRACE(ev_, (unsigned long long)clock64())
#else
#define TRACE(ev_, val_) do {} while (0)
#define TRACE_T(ev_) do {} while (0)
#endif
#ifdef FLDR_RING_STATS
#define STAT_T0() const long long _t0 = clock64()
#define STAT_ADD(i) do { if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(ctrl + 40) + (i), (unsigned long long)(clock64() - _t0)); } while (0)
#define STAT_INC(i, v) do { if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(ctrl + 40) + (i), (unsigned long long)(v)); } while (0)
#else
#define STAT_T0() do {} while (0)
#define STAT_ADD(i) do {} while (0)
#define STAT_INC(i, v) do {} while (0)
#endif

template <int WKIND, bool PRE, int QS, bool BOUNDED>
__global__ void __launch_bounds__(ring::kThreads, ring::kCtasPerSm) splat_ring_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                                       const __grid_constant__ CUtensorMap tm_flow,
                                                                       const __grid_constant__ CUtensorMap tm_metric,
                                                                       float4* __restrict__ ringbuf, unsigned* __restrict__ ctrl,
                                                                       float* __restrict__ out, float* __restrict__ norm_out,
                                                                       SplatGeom g, RingGeom rg) {
    using namespace ring;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage = reinterpret_cast<float*>(smem_raw);
    ItemDesc* desc = reinterpret_cast<ItemDesc*>(smem_raw + (size_t)NST * STAGE_FLOATS * 4);
    uint64_t* bar_posted = reinterpret_cast<uint64_t*>(desc + NSLOT);   // producer wrote the descriptor
    uint64_t* bar_full = bar_posted + NSLOT;                            // inputs landed + dependencies satisfied
    uint64_t* bar_done = bar_full + NSLOT;                              // consumers finished the item
    uint64_t* bar_free = bar_done + NSLOT;                              // completion signalled: descriptor reusable
    int* s_ovf = reinterpret_cast<int*>(bar_free + NSLOT);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&bar_posted[i], 1);
            mbar_init(&bar_full[i], 1);            // producer, after issuing the copies (+ their bytes)
            mbar_init(&bar_done[i], kConsumerWarps);
            mbar_init(&bar_free[i], 1);
            s_ovf[i] = 0;
        }
        fence_barrier_init();
    }
    __syncthreads();

    unsigned* sdone = ctrl + kRingCtrlCounters;      // [NT] scatter tiles (x quads) done per strip
    unsigned* ev = sdone + rg.NT;                    // [nZs + NT] clean events: Z(slot) then N(strip); strip JJ's ring slot was last
                                                     // cleaned by event JJ (Z(JJ) for the first tenants, N(JJ - RS) afterwards)
    unsigned* flag = ctrl + kRingCtrlFlag;
#ifdef FLDR_RING_TRACE
    unsigned long long* trace = reinterpret_cast<unsigned long long*>(ev + rg.nZs + rg.NT + ((rg.NT + rg.nZs + rg.NT) & 1));
#endif

    if (warp == kProducerWarp) {
        // ------------------------------------------------------------------ producer: tickets, descriptors, input staging
        const uint64_t pol = policy_evict_first();
        const int n_total = rg.nZ + rg.NT * rg.T, s_total = rg.NT * rg.TQ;
        // Work is assigned statically: CTA b owns the scatter tickets b, b + G, ... and the zero / normalise tickets
        // b, b + G, ... (G = grid size), and takes them in order - no ticket atomics.  A ticket is only started once the
        // frontiers (read with one relaxed load, cached) cover its dependencies, so a CTA never holds an item that waits
        // for work another CTA has not started; zero / normalise tickets go first (they free ring rows).
        int next_n = blockIdx.x, next_s = blockIdx.x;
        int sf = 0, cf = 0;
        if (lane == 0) { tma_prefetch_desc(&tm_in); tma_prefetch_desc(&tm_flow); if (g.has_metric) tma_prefetch_desc(&tm_metric); }
        for (int k = 0;; ++k) {
            const int slot = k % NSLOT, use = k / NSLOT;
            { STAT_T0(); if (k >= NSLOT) mbar_wait_wd(&bar_free[slot], (use - 1) & 1); STAT_ADD(6); }
            TRACE_T(0);
            ItemDesc d;
            d.kind = kExit; d.J = 0; d.n = 0; d.q = 0; d.x0 = 0; d.cols = 0; d.yb = 0; d.rows = 0; d.row0 = 0;
            d.dep_lo = 0; d.dep_hi = -1; d.ylo = 0.f; d.yhi = 0.f;
            int t = 0;
            {
                STAT_T0();
                int spins = 0;
                for (;;) {
                    if (next_n < n_total) {
                        const int J = next_n < rg.nZ ? -1 : (next_n - rg.nZ) / rg.T;
                        if (J < frontier_limit(sf, rg)) {
                            if (J < 0) { d.kind = kZero; d.J = next_n / rg.T; t = next_n % rg.T; }
                            else { d.kind = kNorm; d.J = J; t = (next_n - rg.nZ) % rg.T; }
                            next_n += gridDim.x;
                            break;
                        }
                    }
                    if (next_s < s_total) {
                        const int J = next_s / rg.TQ;
                        if (J < frontier_limit(cf, rg)) {
                            d.kind = kScatter; d.J = J;
                            const int r = next_s % rg.TQ;
                            d.q = r / rg.T; t = r % rg.T;
                            next_s += gridDim.x;
                            break;
                        }
                    }
                    if (next_n >= n_total && next_s >= s_total) break;                  // this CTA's tickets are done: exit
                    // not runnable with the cached frontiers: re-read them; if they did not move, try to move them
                    const uint4 c = ld_relaxed_v4(ctrl + kRingCtrlClaim);
                    if ((int)c.z == sf && (int)c.w == cf) {
                        if (!advance_frontiers(ctrl, sdone, ev, rg, sf, cf, lane)) {
                            if (++spins > rg.spin_limit) { if (lane == 0) atomicOr(flag, 2u); }
                            __nanosleep(64);
                        }
                        if (ld_relaxed_u32(flag) & 2u) break;                           // watchdog tripped somewhere: drain
                    } else {
                        sf = (int)c.z; cf = (int)c.w;
                    }
                }
                STAT_ADD(8);
            }
            TRACE_T(1);
            TRACE(6, (unsigned long long)d.kind);
            if (d.kind == kExit) {
                if (lane == 0) { desc[slot] = d; mbar_arrive(&bar_posted[slot]); mbar_arrive(&bar_full[slot]); }
                break;
            }
            d.x0 = t * TW;
            d.cols = min(TW, g.W - d.x0);
            if (d.kind != kZero) {
                d.n = d.J / rg.NS;
                const int j = d.J - d.n * rg.NS;
                d.yb = j * R;
                d.rows = min(R, g.H - d.yb);
                d.row0 = (d.n * rg.NS * R) & rg.mask;
                const int jlo = max(0, j - rg.Ds), jhi = min(rg.NS - 1, j + rg.Ds);
                d.dep_lo = d.n * rg.NS + jlo;
                d.dep_hi = d.n * rg.NS + jhi;
                d.ylo = (jlo == 0) ? -1.f : (float)(jlo * R);
                d.yhi = (jhi == rg.NS - 1) ? (float)g.H : (float)((jhi + 1) * R - 1);     // y0 + 1 must stay inside strip jhi
            }
            // the item is claimed and decoded ahead of time; only now wait for its stage (k % NST: previous tenant was read)
            if (d.kind == kScatter || (d.kind == kNorm && rg.Q == 1)) {
                STAT_T0(); if (k >= NST) mbar_wait_wd(&bar_done[(k - NST) % NSLOT], ((k - NST) / NSLOT) & 1); STAT_ADD(7);
            }
            if (d.kind == kScatter) {
                const int sidx = k % NST;
                const int hm = g.has_metric ? 1 : 0;
                const int nch = max(0, min(4, g.C - d.q * 4));
                const int nbox = (rg.Q == 1) ? g.C : 4;               // channels per input box (fixed when the map was encoded)
                float* sbase = stage + (size_t)sidx * STAGE_FLOATS;
                if (lane == 0) {
                    desc[slot] = d;
                    mbar_expect_tx(&bar_full[slot], (uint32_t)((2 + hm + (nch > 0 ? nbox : 0)) * R * TW * 4));
                    mbar_arrive(&bar_posted[slot]);
                    tma_box_load(sbase, &tm_flow, &bar_full[slot], d.x0, d.yb, 0, d.n, pol);
                    if (hm) tma_box_load(sbase + 2 * R * TW, &tm_metric, &bar_full[slot], d.x0, d.yb, 0, d.n, pol);
                    if (nch > 0) tma_box_load(sbase + (2 + hm) * R * TW, &tm_in, &bar_full[slot], d.x0, d.yb, d.q * 4, d.n, pol);
                    mbar_arrive(&bar_full[slot]);
                }
                __syncwarp();
            } else if (d.kind == kNorm && rg.Q == 1) {
                // single-quad splats: copy the strip's ring cells into the item's stage, so the consumers never wait for L2.
                // The strip's producers are complete (the item holds a credit); their reductions went through the generic
                // proxy, the copy reads through the async proxy.
                fence_proxy_async_all();
                const uint32_t row_bytes = (uint32_t)d.cols * 16u;
                if (lane == 0) {
                    desc[slot] = d;
                    mbar_expect_tx(&bar_full[slot], row_bytes * (uint32_t)d.rows);
                    mbar_arrive(&bar_posted[slot]);
                }
                __syncwarp();
                if (lane < d.rows) {
                    const float4* src = ringbuf + (size_t)((d.row0 + d.yb + lane) & rg.mask) * rg.pitch + 1 + d.x0;
                    bulk_load_plain(stage + (size_t)(k % NST) * STAGE_FLOATS + lane * TW * 4, src, row_bytes, &bar_full[slot]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[slot]);
            } else {
                if (lane == 0) { desc[slot] = d; mbar_arrive(&bar_posted[slot]); mbar_arrive(&bar_full[slot]); }
            }
            TRACE_T(2);
        }
    } else if (warp == kSignalWarp) {
        // ------------------------------------------------------------------ signaller: fence + completion counter
        for (int k = 0;; ++k) {
            const int slot = k % NSLOT, use = k / NSLOT;
            mbar_wait_wd(&bar_posted[slot], use & 1);
            const int kind = desc[slot].kind, J = desc[slot].J;
            if (kind == kExit) break;
            { STAT_T0(); mbar_wait_wd(&bar_done[slot], use & 1); STAT_ADD(10); }
            STAT_T0();
            // release: everything the consumers wrote (reductions, zero stores) is performed at L2 before the counter moves
            // (CTA-scope barrier above, then one gpu-scope fence: the cooperative-groups grid.sync pattern)
            int completed = 0;
            if (lane == 0) {
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                if (kind == kScatter) {
                    if (s_ovf[slot]) { atomicOr(flag, 1u); s_ovf[slot] = 0; }
                    completed = atomicAdd(&sdone[J], 1u) + 1u == (unsigned)rg.TQ;
                } else if (kind == kNorm) {
                    completed = atomicAdd(&ev[rg.nZs + J], 1u) + 1u == (unsigned)rg.T;
                } else {
                    completed = atomicAdd(&ev[J], 1u) + 1u == (unsigned)rg.T;
                }
                mbar_arrive(&bar_free[slot]);
            }
            // the tile that completes a strip moves the frontier (and hands out the credits) right away
            if (__shfl_sync(0xffffffffu, completed, 0)) {
                const uint4 c = ld_relaxed_v4(ctrl + kRingCtrlClaim);
                advance_frontiers(ctrl, sdone, ev, rg, (int)c.z, (int)c.w, lane);
            }
            __syncwarp();
            TRACE_T(5);
            STAT_ADD(9);
        }
    } else {
        // ------------------------------------------------------------------ consumers
        const size_t plane = (size_t)rg.ring_rows * rg.pitch;
        for (int k = 0;; ++k) {
            const int slot = k % NSLOT, use = k / NSLOT;
            { STAT_T0(); mbar_wait_wd(&bar_full[slot], use & 1); STAT_ADD(0); }
            const ItemDesc d = desc[slot];
            if (d.kind == kExit) break;
            if (warp == 0) TRACE_T(3);
            STAT_T0();
            if (d.kind == kScatter) {
                const float* st = stage + (size_t)(k % NST) * STAGE_FLOATS;
                const bool ovf = ring_scatter_item<WKIND, PRE, QS, BOUNDED>(st, d, g, rg, ringbuf, tid, lane);
                if (BOUNDED && __any_sync(0xffffffffu, ovf) && lane == 0) s_ovf[slot] = 1;
            } else if (d.kind == kNorm) {
                if (rg.Q == 1)
                    ring_normalise_item_staged(reinterpret_cast<const float4*>(stage + (size_t)(k % NST) * STAGE_FLOATS), d, g, rg,
                                               ringbuf, out, norm_out, warp, lane);
                else
                    ring_normalise_item(d, g, rg, ringbuf, out, norm_out, warp, lane);
            } else {
                // Z: cells 1 .. W of the slot's 8 rows, every quad
                if (tid < d.cols)
                    for (int q = 0; q < rg.Q; ++q) {
                        float4* p = ringbuf + q * plane + (size_t)(d.J * R) * rg.pitch + 1 + d.x0 + tid;
#pragma unroll
                        for (int r = 0; r < R; ++r) st_cg4_zero(p + (size_t)r * rg.pitch);
                    }
            }
            __syncwarp();
            if (warp == 0) TRACE_T(4);
            STAT_ADD(d.kind == kScatter ? 1 : 2);
            if (lane == 0) mbar_arrive(&bar_done[slot]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------- host side

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool view_ok(const View4& v) {
    return v.sw == 1 && (v.sh % 4) == 0 && (v.sc % 4) == 0 && (v.sn % 4) == 0 && aligned16(v.p);
}

bool ring_eligible(const SplatGeom& g, const View4& in, const View4& flow, const View4& metric, const float* out, const float* norm) {
    if (g.W % 4 != 0 || g.W < 4) return false;
    if (!view_ok(in) || !view_ok(flow)) return false;
    if (g.has_metric && (!view_ok(metric) || g.C >= 4)) return false;     // at most 6 staged planes per item
    if (!aligned16(out) || (norm && !aligned16(norm))) return false;
    if (((long long)g.H * g.W) % 4 != 0) return false;
    return true;
}

void plan_ring(const SplatGeom& g, RingPlan& p) {
    using namespace ring;
    RingGeom& rg = p.rg;
    memset(&rg, 0, sizeof(rg));
    rg.Q = g.CP / 4;
    rg.NS = (g.H + R - 1) / R;
    rg.T = (g.W + TW - 1) / TW;
    rg.pitch = g.W + 2;
    const long long NT = (long long)g.N * rg.NS;
    p.ok = NT < (1ll << 24) && (long long)rg.T * rg.Q < (1ll << 16);
    const long long row_bytes = (long long)rg.pitch * 16 * rg.Q;
    const int ring_mb = get_option(kOptSplatRingMb) > 0 ? get_option(kOptSplatRingMb) : 34;   // 512 rows of 4096 + 2 cells
    long long rows = 64;
    while (rows * 2 * row_bytes <= ((long long)ring_mb << 20)) rows *= 2;
    // batches whose whole accumulator is <= 48 MiB are held entirely: unbounded reach, rows never wrap (mask = all ones)
    const bool whole = NT * R * row_bytes <= (48ll << 20) || NT * R <= rows;
    if (whole) rows = NT * R;
    if (rows * (long long)rg.pitch >= (1ll << 30)) p.ok = false;          // cell offsets are 32-bit, negative = none
    rg.RS = (int)(rows / R);
    rg.NT = (int)NT;
    rg.ring_rows = (int)rows;
    rg.mask = whole ? -1 : (int)rows - 1;
    rg.TQ = rg.T * rg.Q;
    const int lag_opt = get_option(kOptSplatLag);
    if (whole) {
        rg.bounded = 0;
        rg.Ds = rg.NS;
    } else {
        // The ring holds 2 Ds strips of reach plus the strips in flight between the normalise frontier and the scatter
        // head: S(J) is claimable once the slot of strip J + Ds was cleaned (N(J + Ds - RS) done) and N(J) once
        // S(J + Ds) is done, so scatter runs at most RS - 2 Ds = 2 lag strips ahead of what normalise has retired.
        const int lag = lag_opt > 0 ? lag_opt : 12;
        rg.bounded = 1;
        rg.Ds = (rg.RS - 2 * lag) / 2;
        if (rg.Ds < 1) p.ok = false;
    }
    rg.nZs = (int)(rg.RS < NT ? rg.RS : NT);
    rg.nZ = rg.nZs * rg.T;
    const long long total = (long long)rg.nZ + NT * rg.T + NT * rg.TQ;
    if (total >= (1ll << 30)) p.ok = false;
    rg.total = (int)total;
    rg.spin_limit = 1 << 18;
    p.ring_bytes = align_up((size_t)rows * row_bytes, 256);
    p.ctrl_bytes = align_up(((size_t)kRingCtrlCounters + (size_t)NT + rg.nZs + (size_t)NT) * 4, 256);
#ifdef FLDR_RING_TRACE
    p.ctrl_bytes += 8 * 128 * 8 * 8 + 256;
#endif
    if (!p.ok) { p.ring_bytes = 0; p.ctrl_bytes = 256; }
}

struct RingMaps { CUtensorMap in, flow, metric; };

template <int WKIND, bool PRE, int QS, bool BOUNDED>
static int launch_ring_t(const RingPlan& p, const SplatGeom& g, const RingMaps& m, void* ringbuf, unsigned* ctrl, float* out,
                         float* norm, cudaStream_t s) {
    auto fn = splat_ring_kernel<WKIND, PRE, QS, BOUNDED>;
    // per-instantiation, per-device launch state (attribute set + occupancy), computed once
    static int ctas_per_sm[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int per_sm = __atomic_load_n(&ctas_per_sm[dev], __ATOMIC_ACQUIRE);
    if (per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring::kSmemBytes);
        if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, ring::kThreads, ring::kSmemBytes) != cudaSuccess || nb < 1) nb = 1;
        per_sm = nb > ring::kCtasPerSm ? ring::kCtasPerSm : nb;
        __atomic_store_n(&ctas_per_sm[dev], per_sm, __ATOMIC_RELEASE);
    }
    long long grid = (long long)sm_count() * per_sm;
    if (grid > p.rg.total) grid = p.rg.total;
    fn<<<(unsigned)grid, ring::kThreads, ring::kSmemBytes, s>>>(m.in, m.flow, m.metric, reinterpret_cast<float4*>(ringbuf), ctrl, out,
                                                                norm, g, p.rg);
    return check_launch();
}

// [N][C][H][W] view -> 4-D tensor map with an [nbox][8][128] box
static bool encode_view(CUtensorMap* map, const View4& v, int N, int C, int H, int W, int nbox) {
    const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)v.sh * 4, (uint64_t)v.sc * 4, (uint64_t)v.sn * 4};
    const uint32_t box[4] = {(uint32_t)ring::TW, (uint32_t)ring::R, (uint32_t)nbox, 1u};
    if (v.sh <= 0 || v.sc <= 0 || v.sn <= 0) return false;
    return encode_tensor_map_4d(map, v.p, dims, strides, box);
}

// Returns FLDR_ERR_UNSUPPORTED when the views cannot be described by tensor maps (the caller then takes the whole-frame path).
int launch_ring(const RingPlan& p, const SplatGeom& g, const View4& in, const View4& flow, const View4& metric, void* ringbuf,
                unsigned* ctrl, float* out, float* norm, cudaStream_t s) {
    RingMaps m;
    const int nbox = (p.rg.Q == 1) ? g.C : 4;
    View4 vin = in, vfl = flow, vme = metric;
    // single-sample / single-channel dimensions: any positive 16-byte-multiple stride describes them
    if (g.N == 1) { vin.sn = (long long)g.C * g.H * g.W + 4; vfl.sn = 2ll * g.H * g.W + 4; vme.sn = (long long)g.H * g.W + 4; vin.sn -= vin.sn % 4; vfl.sn -= vfl.sn % 4; vme.sn -= vme.sn % 4; }
    if (g.C == 1) vin.sc = ((long long)g.H * g.W + 4) / 4 * 4;
    vme.sc = ((long long)g.H * g.W + 4) / 4 * 4;
    if (!encode_view(&m.in, vin, g.N, g.C, g.H, g.W, nbox) || !encode_view(&m.flow, vfl, g.N, 2, g.H, g.W, 2)) return FLDR_ERR_UNSUPPORTED;
    if (g.has_metric) { if (!encode_view(&m.metric, vme, g.N, 1, g.H, g.W, 1)) return FLDR_ERR_UNSUPPORTED; }
    else m.metric = m.flow;
    cudaError_t e = cudaMemsetAsync(ctrl, 0, p.ctrl_bytes, s);
    if (e != cudaSuccess) { set_last_cuda_error(e); return FLDR_ERR_CUDA; }
    const int wkind = !g.has_metric ? 0 : (g.mode == FLDR_SPLAT_SOFTMAX ? 1 : 2);
    const bool pre = g.mode == FLDR_SPLAT_SOFTMAX;
    const int qs = (g.C == 3 && g.CA == 4) ? 1 : (g.CA == g.C && g.C % 4 == 0) ? 2 : 0;
    const bool b = p.rg.bounded != 0;
#define FLDR_RING3(WK_, PRE_, QS_) \
    (b ? launch_ring_t<WK_, PRE_, QS_, true>(p, g, m, ringbuf, ctrl, out, norm, s) \
       : launch_ring_t<WK_, PRE_, QS_, false>(p, g, m, ringbuf, ctrl, out, norm, s))
#define FLDR_RING2(WK_, PRE_) (qs == 1 ? FLDR_RING3(WK_, PRE_, 1) : qs == 2 ? FLDR_RING3(WK_, PRE_, 2) : FLDR_RING3(WK_, PRE_, 0))
    if (wkind == 1) return FLDR_RING2(1, true);
    if (wkind == 2) return FLDR_RING2(2, false);
    if (pre) return FLDR_RING2(0, true);
    return FLDR_RING2(0, false);
#undef FLDR_RING2
#undef FLDR_RING3
}

}  // namespace fldr
