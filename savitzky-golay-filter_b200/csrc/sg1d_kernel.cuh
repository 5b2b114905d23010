// sg1d_kernel.cuh -- the 1D Savitzky-Golay stencil for sm_100a.
//
// Replaces the reference's scalar hot loop (src/savgolFilter.c:763-766 calling convolve_ilp
// :547-580, the padded edges :785-801 / :442-535, the polynomial edges :769-784, VALID
// :821-850, strided :877-934) and the steady-state of the stream (src/savgol_stream.c:224-226).
//
// Execution plan:
//   * warp-autonomous pipeline: the unit of work is a segment of 1024 outputs of one signal (32 lanes
//     x 32 consecutive outputs) plus halo; every warp owns two private shared-memory buffers, issues
//     the cp.async copies (16 B, L2-only) of its next segment, waits for the current one, computes
//     and stores.  No __syncthreads anywhere.
//   * the first/last segment of a signal synthesises the virtual pad samples (reflect / periodic /
//     constant / explicit halo / carried stream history) while staging -- one pad element per lane
//     -- so the compute loop is boundary-agnostic (identity Q6 of SURVEY.md: padded modes == VALID
//     over the padded signal).
//   * shared layout: one 16-byte pad chunk after every 8 chunks -> a lane's 32-sample stride
//     becomes 9 chunks (odd), every LDS.128 of the sliding window is bank-conflict free and all
//     offsets stay compile-time immediates.
//   * arithmetic: register sliding window, packed fp32 (FFMA2, fma.rn.f32x2) over output pairs with
//     the sample broadcast and the weight pair (w[k], w[k-1]) in uniform registers.
//   * polynomial edges: one lane per edge output from the transposed edge table, patched into the
//     owners' registers before the store.
//   * stores go through the warp's own buffer so that each store instruction writes one contiguous
//     run of global memory.
#pragma once
#include <type_traits>
#include <utility>

#include "sg_common.cuh"

#ifndef SG_MIN_BLOCKS
#define SG_MIN_BLOCKS 5  // resident CTAs per SM the register allocator must allow
#endif

namespace sg {

template <int LEAD>
struct Geo {
    static constexpr int PAD = (LEAD + 3) & ~3;   // shared position 0 <-> x index o0 - PAD (keeps 16 B alignment)
    static constexpr int DELTA = PAD - LEAD;      // thread t, output j, tap k reads shared sample 32t + j + k + DELTA
};

template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>)
{
    (f(std::integral_constant<int, I>{}), ...);
}
template <int COUNT, class F>
__device__ __forceinline__ void static_for(F&& f)
{
    static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, COUNT>{});
}

__device__ __forceinline__ float ld_sample(const char* xrow, long long stride, long long i)
{
    return *reinterpret_cast<const float*>(xrow + i * stride);
}

// Address of the sample that stands for x-index xi of row `row` (xi may lie outside [0,len)):
// the real sample, an explicit halo entry, or the sample the boundary rule maps it to
// (ref: src/savgolFilter.c:442-482, 64-bit clean -- SURVEY.md Q5).  nullptr = "no such sample"
// (polynomial mode pads and alignment slack), staged as 0.
template <int LEAD, int N>
__device__ __forceinline__ const float* sample_address(const Args1D& a, const char* xrow, long long row, long long xi)
{
    const long long len = a.len;
    long long idx = xi;
    if (xi < 0) {
        if (a.lhalo) {
            const long long h = LEAD + xi;
            return h >= 0 ? a.lhalo + row * a.lhalo_pitch + h : nullptr;
        }
        switch (a.mode) {
            case MODE_REFLECT: idx = -xi - 1; if (idx >= len) idx = len - 1; break;
            case MODE_PERIODIC: idx = ((xi % len) + len) % len; break;
            case MODE_CONSTANT: idx = 0; break;
            default: return nullptr;
        }
    } else if (xi >= len) {
        if (a.rhalo) {
            const long long h = xi - len;
            return h < N ? a.rhalo + row * a.rhalo_pitch + h : nullptr;
        }
        switch (a.mode) {
            case MODE_REFLECT: idx = 2 * len - xi - 1; if (idx < 0) idx = 0; break;
            case MODE_PERIODIC: idx = xi % len; break;
            case MODE_CONSTANT: idx = len - 1; break;
            default: return nullptr;
        }
    }
    return reinterpret_cast<const float*>(xrow + idx * a.in_stride);
}

template <int LEAD, int N>
__device__ __forceinline__ float virtual_sample(const Args1D& a, const char* xrow, long long row, long long xi)
{
    const float* p = sample_address<LEAD, N>(a, xrow, row, xi);
    return p ? *p : 0.0f;
}

// Contiguous row that is not 16-byte aligned (odd pitch, offset view): lane-consecutive 4-byte copies,
// i.e. every copy instruction still covers one contiguous 128-byte run.  Element e = lane + 32 i lives
// at float position e + 4 (e >> 5) = lane + 36 i of the padded buffer; it is copied when lo <= 32 i < hi
// (bounds already relative to the lane).  Out of line, like store_scalar below: these paths are fully
// unrolled and must not cost the aligned path registers.
template <int NE>
__device__ __noinline__ void stage_unaligned(float* d, const char* s, int lo, int hi)
{
#pragma unroll
    for (int i = 0; i < NE; ++i)
        if (32 * i >= lo && 32 * i < hi) cp_async4(d + 36 * i, s + 128 * i);
}

// Lane-interleaved scalar stores of a parked segment (misaligned or strided rows): output f = lane + 32 i is
// parked at float position lane + 36 i; it is stored when lo <= 32 i < lim (bounds relative to the lane).
static __device__ __noinline__ void store_scalar(const float* srcf, char* dst, long long stride, int lo, int lim)
{
    if (stride == 4) {
        float* d = reinterpret_cast<float*>(dst);
#pragma unroll
        for (int i = 0; i < kR; ++i)
            if (32 * i >= lo && 32 * i < lim) d[32 * i] = srcf[36 * i];
    } else {
        const long long step = 32 * stride;
#pragma unroll 4
        for (int i = 0; i < kR; ++i) {
            if (32 * i >= lo && 32 * i < lim) *reinterpret_cast<float*>(dst) = srcf[36 * i];
            dst += step;
        }
    }
}

// A parked segment whose chunks are 16-byte aligned in global memory but whose ends are cut (the first segment of
// a phase-shifted row starts before output 0, the last one is ragged): whole chunks as 16-byte stores, the cut
// chunks element by element.  Chunk lane + 32 i holds outputs 128 i .. 128 i + 3 relative to the lane's first;
// [lo, hi) are the outputs that exist, in the same coordinates.
static __device__ __noinline__ void store_cut(const float4* src /* buf + lane + (lane >> 3) */, float* dst, int lo, int hi)
{
#pragma unroll
    for (int i = 0; i < kR / 4; ++i) {
        const int e = 128 * i;
        if (e + 4 <= lo || e >= hi) continue;
        const float4 v = src[36 * i];
        if (e >= lo && e + 4 <= hi) {
            st_cs_f4(dst + e, v);
        } else {
            if (e >= lo && e < hi) dst[e] = v.x;
            if (e + 1 >= lo && e + 1 < hi) dst[e + 1] = v.y;
            if (e + 2 >= lo && e + 2 < hi) dst[e + 2] = v.z;
            if (e + 3 >= lo && e + 3 < hi) dst[e + 3] = v.w;
        }
    }
}

// Stage one segment (kSeg outputs of one row, plus halo) into a warp's buffer: shared position 4c
// <-> x index o0 - PAD + 4c.  Everything is asynchronous (cp.async), so no lane waits on a global
// load here:
//   * chunks whose four samples exist are copied 16 bytes at a time (4 x 4 bytes when the row is
//     misaligned or strided): 8 full passes of the warp plus a tail,
//   * the remaining elements -- virtual pad samples, ragged ends, alignment slack -- form the
//     "edge path": one element per lane, each lane maps its element to the address the boundary
//     rule designates and copies 4 bytes, or stores 0.
template <int LEAD, int N>
__device__ __forceinline__ void stage_segment(float4* dst0 /* buf + lane + (lane >> 3) */, float4* buf, const Args1D& a,
                                              const char* xrow, long long row, long long o0, int left /* min(len-o0, big) */,
                                              int cap /* outputs this segment may produce: kSeg, or kSeg + kTail */, bool rows_aligned, int lane)
{
    constexpr int PAD = Geo<LEAD>::PAD;
    constexpr int DELTA = Geo<LEAD>::DELTA;
    // segment-local 32-bit bookkeeping: chunk c covers x indices o0 - PAD + 4c .. +3
    const int nout = left < cap ? left : cap;
    const int nch = (nout + 2 * N + DELTA + 3) >> 2;          // chunks the compute loop may touch
    const int c_lo = o0 <= 0 ? (PAD - static_cast<int>(o0) + 3) >> 2 : 0;   // first chunk made of four existing samples (o0 < 0: phase-shifted row)
    const int c_all = (left + PAD) >> 2;                       // chunks that end inside the row
    const int c_hi = c_all < nch ? c_all : nch;
    const char* src0 = xrow + (o0 - PAD) * a.in_stride;        // only dereferenced inside [c_lo, c_hi)
    const bool vec_ok = rows_aligned || ((a.in_stride == 4) && ((reinterpret_cast<uintptr_t>(src0) & 15) == 0));
    if (vec_ok) {
        // this lane's first chunk; the pointer is made opaque so that it stays in a register pair and
        // every copy is "base + immediate" (otherwise it is re-derived from the kernel parameters
        // with three extra instructions per copy)
        const char* s = src0 + lane * 16;
        asm volatile("" : "+l"(s));
        if (c_hi >= 256) {
            // full segment: only the first pass (left pad chunks of a row's first segment) and the
            // tail pass are conditional
            if (lane >= c_lo) cp_async16(dst0, s);
#pragma unroll
            for (int it = 1; it < 8; ++it) cp_async16(dst0 + 36 * it, s + 512 * it);  // chunk lane + 32 it lives at dst0 + 36 it
            if (lane + 256 < c_hi) cp_async16(dst0 + 36 * 8, s + 512 * 8);
        } else {
#pragma unroll
            for (int it = 0; it < 9; ++it) {
                const int c = lane + 32 * it;
                if (c >= c_lo && c < c_hi) cp_async16(dst0 + 36 * it, s + 512 * it);
            }
        }
    } else if (a.in_stride == 4) {
        constexpr int NE = (4 * ((kSeg + kTail + 2 * N + DELTA + 3) / 4) + 31) / 32;
        stage_unaligned<NE>(reinterpret_cast<float*>(buf) + lane, src0 + 4 * lane, 4 * c_lo - lane, 4 * c_hi - lane);
    } else {
        const long long st = a.in_stride;
#pragma unroll 1
        for (int it = 0; it < 9; ++it) {
            const int c = lane + 32 * it;
            if (c >= c_lo && c < c_hi) {
                float* d = reinterpret_cast<float*>(dst0 + 36 * it);
                const char* sp = src0 + 4LL * c * st;
                cp_async4(d, sp); cp_async4(d + 1, sp + st); cp_async4(d + 2, sp + 2 * st); cp_async4(d + 3, sp + 3 * st);
            }
        }
    }
    // (elements more than LEAD samples before the row belong to outputs that do not exist: phase-shifted first segment)
    const int el0 = o0 < 0 ? static_cast<int>(-o0) + DELTA : 0;
    const int nl = 4 * c_lo - el0, nrest = nl + 4 * (nch - c_hi);
#pragma unroll 1
    for (int q = lane; q < nrest; q += 32) {
        const int el = q < nl ? el0 + q : 4 * c_hi + (q - nl);  // element index inside the segment buffer
        const int c = el >> 2;
        float* d = reinterpret_cast<float*>(buf + c + (c >> 3)) + (el & 3);
        const float* sp = sample_address<LEAD, N>(a, xrow, row, o0 - PAD + el);
        if (sp) cp_async4(d, sp);
        else *d = 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// FAST arithmetic: packed FFMA2 over output pairs.
// acc[jj] = (out[2jj], out[2jj+1]).  Sample x[i] (window position i of this thread) is tap k = i-2jj-DELTA
// of out[2jj] and tap k-1 of out[2jj+1], so   acc[jj] += (ws[k], ws[k-1]) * (x, x)   is one FFMA2 with
// the sample broadcast and the weight pair taken from uniform registers.  The first / last tap of a
// pair touch only one of its outputs and are scalar FFMAs.  Pipe work per thread: 32 x (2n+1) MACs,
// exactly the stencil's minimum; weights are pre-scaled by 1/dt^d, so there is no epilogue.
template <int N, int DELTA>
__device__ __forceinline__ void compute_fast(const float4* __restrict__ sb, const W1D& W, float (&out)[kR])
{
    constexpr int WS = 2 * N + 1;
    constexpr int NCHT = (kR + 2 * N + DELTA + 3) / 4;
    float2 acc[kR / 2];
#pragma unroll
    for (int i = 0; i < kR / 2; ++i) acc[i] = make_float2(0.f, 0.f);

    // static_for: the window is up to 24 chunks x 68 FMA instructions; "#pragma unroll" silently gives
    // up on bodies that large (half-windows > 20), and a rolled loop would index weights and
    // accumulators dynamically.  Template expansion keeps every index a compile-time constant.
    static_for<NCHT>([&](auto ci) {
        constexpr int c = decltype(ci)::value;
        const float4 v = sb[c + (c >> 3)];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = 4 * c + e;
            const float x = e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w;
#pragma unroll
            for (int jj = 0; jj < kR / 2; ++jj) {
                const int k = i - 2 * jj - DELTA;  // tap of out[2jj]; out[2jj+1] sees tap k-1
                if (k == 0) acc[jj].x = fmaf(W.ws_first, x, acc[jj].x);
                else if (k == WS) acc[jj].y = fmaf(W.ws_last, x, acc[jj].y);
                else if (k > 0 && k < WS) acc[jj] = __ffma2_rn(W.pw[k], make_float2(x, x), acc[jj]);
            }
        }
    });
#pragma unroll
    for (int jj = 0; jj < kR / 2; ++jj) {
        out[2 * jj] = acc[jj].x;
        out[2 * jj + 1] = acc[jj].y;
    }
}

// One dot product in the reference's order.  X(k) yields the sample multiplied by w(k).
template <int WS, int ARITH, class WF, class XF>
__device__ __forceinline__ float dot_ordered(WF w, XF x)
{
    if constexpr (ARITH == ARITH_EXACT4) {
        // ref: src/savgolFilter.c:547-580 -- remainder taps to chains 0..rem-1, then 4 chains
        constexpr int REM = WS & 3;
        float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int k = 0; k < WS; ++k) {
            const int c = k < REM ? k : ((k - REM) & 3);
            s[c] = __fadd_rn(s[c], __fmul_rn(w(k), x(k)));
        }
        return __fadd_rn(__fadd_rn(s[0], s[1]), __fadd_rn(s[2], s[3]));
    } else if constexpr (ARITH == ARITH_EXACTSEQ) {
        // ref: src/savgol_stream.c:31-35
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < WS; ++k) s = __fadd_rn(s, __fmul_rn(w(k), x(k)));
        return s;
    } else {
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < WS; ++k) s = fmaf(w(k), x(k), s);
        return s;
    }
}

// EXACT arithmetic (verification flavour): groups of 8 outputs, reference summation order.
template <int N, int DELTA, int ARITH>
__device__ __forceinline__ void compute_exact(const float4* __restrict__ sb, const W1D& W, float scale, float (&out)[kR])
{
    constexpr int WS = 2 * N + 1;
    constexpr int G = 8;
    constexpr int NCHG = (G + 2 * N + DELTA + 3) / 4;
#pragma unroll 1
    for (int g = 0; g < kR / G; ++g) {
        float xw[NCHG * 4];
#pragma unroll
        for (int c = 0; c < NCHG; ++c) {
            const int cl = 2 * g + c;  // logical chunk relative to the thread base (thread base is 8-chunk aligned)
            const float4 v = sb[cl + (cl >> 3)];
            xw[4 * c] = v.x; xw[4 * c + 1] = v.y; xw[4 * c + 2] = v.z; xw[4 * c + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const float s = dot_ordered<WS, ARITH>([&](int k) { return W.w[k]; },
                                                   [&](int k) { return xw[j + k + DELTA]; });
            out[g * G + j] = __fmul_rn(s, scale);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-autonomous pipeline.  The unit of work is a SEGMENT: kSeg = 1024 consecutive outputs of one
// row (32 lanes x 32 outputs) together with its halo.  Every warp owns two private segment buffers
// and loops over segments gw, gw + W, gw + 2W, ... (W = warps in the grid): it issues the cp.async
// copies of its NEXT segment, waits for the CURRENT one (cp.async.wait_group + __syncwarp), computes,
// stores.  Warps never synchronise with each other -- there is no __syncthreads in the kernel -- so a
// warp that is waiting on HBM never holds up the FMA pipe of its neighbours, and with ~20 resident
// warps per SM about 80 KB of loads are in flight per SM at any time.


template <int N, int DELTA, bool PT>
struct Smem1D {
    static constexpr int kSegChunks = (kSeg + (PT ? kTail : 0) + 2 * N + DELTA + 3) / 4;
    static constexpr int kSegPhys = kSegChunks + (kSegChunks >> 3) + 1;
};

// PT: the instantiation that knows about per-row phases and short tails (misaligned rows, rows ending just behind a
// full segment; FAST flavours only).  Aligned launches keep the lean PT = false kernel: the extra paths cost the
// compute-bound wide windows 5 % (config 3: 1.148 -> 1.21 ms) through instruction-cache pressure alone.
template <int N, bool LEAD2N, int ARITH, bool PT = false>
__global__ void __launch_bounds__(kThreads, SG_MIN_BLOCKS) sg1d_kernel(const __grid_constant__ W1D W, const __grid_constant__ Args1D a)
{
    constexpr int LEAD = LEAD2N ? 2 * N : N;
    constexpr int DELTA = Geo<LEAD>::DELTA;
    constexpr int WS = 2 * N + 1;
    constexpr int kWarps = kThreads / 32;
    using SM = Smem1D<N, DELTA, PT>;

    __shared__ float4 s_buf[kWarps][2][SM::kSegPhys];
    __shared__ float s_edge[kWarps][2 * kMaxN];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned nseg = static_cast<unsigned>(a.ntiles);        // segments in the launch
    const unsigned spr = static_cast<unsigned>(a.tiles_per_row);  // segments per row
    const unsigned stride = gridDim.x * kWarps;                   // segments between two iterations of a warp
    const long long len = a.len;
    // rows whose base and pitch are 16-byte aligned keep every segment start aligned (o0 and PAD are
    // multiples of 4 samples): decide once instead of per segment
    const bool rows_aligned = a.in_stride == 4 && ((reinterpret_cast<uintptr_t>(a.in) | static_cast<uintptr_t>(a.in_row_bytes)) & 15) == 0;
    constexpr int kBig = kSeg + 4 * kMaxWs;  // "plenty left": clamp for the 32-bit per-segment bookkeeping

    // (row, t) = position of the current segment; advancing by `stride` segments is an add with carry,
    // so the loop has no division
    unsigned seg = blockIdx.x * kWarps + warp;
    unsigned row_u = seg / spr, t = seg - row_u * spr;
    const unsigned step_r = stride / spr, step_t = stride - step_r * spr;
    long long row = row_u;
    float4* buf_cur = s_buf[warp][0];  // buffer holding (or receiving) the current segment
    float4* buf_nxt = s_buf[warp][1];  // buffer the next segment is prefetched into
    const int lane_chunk = lane + (lane >> 3);

    // Contiguous rows that are not 16-byte aligned (odd length or pitch, offset views) are cut on a per-row
    // phase instead: segment t starts at output t*kSeg - sh with sh = (address of x[0] / 4) mod 32, so every
    // segment starts on a 128-byte line of global memory: staging (and, when the output row has the same phase,
    // the stores) runs the aligned path and every warp-wide copy covers exactly four lines, as for aligned rows.
    // The first segment then begins before output 0; those outputs are computed from whatever is staged and never
    // stored.  a.phase is set by the launcher, which also counts the segments per row for the worst phase.
    const char* xrow = a.in + row * a.in_row_bytes;
    auto row_phase = [&](const char* r) -> long long { return (PT && a.phase) ? static_cast<long long>((reinterpret_cast<uintptr_t>(r) >> 2) & (kPhase - 1)) : 0; };
    long long o0 = static_cast<long long>(t) * kSeg - row_phase(xrow);
    // a.tail: rows end with up to kTail outputs behind their last segment (launcher: a nearly empty extra segment
    // would cost a full pass of the warp); that segment stages the extra samples and each lane adds one output.
    const int tail_cap = (PT && a.tail) ? kSeg + kTail : kSeg;
    if (seg < nseg && (!PT || o0 < len)) {
        const long long l64 = len - o0;
        stage_segment<LEAD, N>(buf_cur + lane_chunk, buf_cur, a, xrow, row, o0, static_cast<int>(l64 < kBig ? l64 : kBig),
                               t + 1 == spr ? tail_cap : kSeg, rows_aligned, lane);
    }
    cp_async_commit();

    for (; seg < nseg; seg += stride) {
        // position of this warp's next segment; prefetch it into the other buffer
        unsigned nt = t + step_t;
        long long nrow = row + step_r;
        if (nt >= spr) { nt -= spr; ++nrow; }
        const char* nxrow = a.in + nrow * a.in_row_bytes;
        const long long no0 = static_cast<long long>(nt) * kSeg - row_phase(nxrow);
        if (seg + stride < nseg) {
            const long long l64 = len - no0;
            if (!PT || l64 > 0)   // (a phase-shifted row may not reach into its last segment)
                stage_segment<LEAD, N>(buf_nxt + lane_chunk, buf_nxt, a, nxrow, nrow, no0,
                                       static_cast<int>(l64 < kBig ? l64 : kBig), nt + 1 == spr ? tail_cap : kSeg, rows_aligned, lane);
        }
        cp_async_commit();
        const bool has_tail = PT && a.tail && t + 1 == spr;
        const int seg_cap = has_tail ? kSeg + kTail : kSeg;
        if (!PT || o0 < len) {

        // polynomial edge outputs that fall into this segment (global reads, independent of the
        // staged data): one lane per output.  ref: src/savgolFilter.c:769-784
        const bool lead_seg = a.edge_lead && o0 < N;
        const bool trail_seg = a.edge_trail && (o0 + seg_cap > len - N);
        if (lead_seg && lane < N) {
            // out[e] = scale * sum_k E[e][k] * x[2n-k]   (reversed traversal, ref :593-623)
            const float s = dot_ordered<WS, ARITH>([&](int k) { return a.edge_t[k * 32 + lane]; },
                                                   [&](int k) { return ld_sample(xrow, a.in_stride, 2 * N - k); });
            s_edge[warp][lane] = ARITH == ARITH_FAST ? s * a.scale : __fmul_rn(s, a.scale);
        }
        if (trail_seg && lane < N) {
            // out[len-1-e] = scale * sum_k E[e][k] * x[len-ws+k]
            const long long base = len - WS;
            const float s = dot_ordered<WS, ARITH>([&](int k) { return a.edge_t[k * 32 + lane]; },
                                                   [&](int k) { return ld_sample(xrow, a.in_stride, base + k); });
            s_edge[warp][kMaxN + lane] = ARITH == ARITH_FAST ? s * a.scale : __fmul_rn(s, a.scale);
        }

        cp_async_wait<1>();  // this lane's copies of the current segment have landed ...
        __syncwarp();        // ... and so have the other lanes' (and their s_edge entries)

        const float4* sb = buf_cur + 9 * lane;
        float out[kR];
        if constexpr (ARITH == ARITH_FAST) compute_fast<N, DELTA>(sb, W, out);
        else compute_exact<N, DELTA, ARITH>(sb, W, a.scale, out);

        const long long o = o0 + kR * lane;
        if (lead_seg || trail_seg) {
#pragma unroll
            for (int j = 0; j < kR; ++j) {
                const long long oj = o + j;
                if (lead_seg && oj < N) { if (oj >= 0) out[j] = s_edge[warp][oj]; }
                else if (trail_seg && oj >= len - N && oj < len) out[j] = s_edge[warp][kMaxN + (len - 1 - oj)];
            }
        }

        // stream: hand the last state_w samples of [lead pad | x] to the next chunk
        // (read back from the staged segment whenever it holds them: a global re-read would put a
        // full L2 round trip on every segment's critical path)
        if (a.state_out != nullptr && o0 + seg_cap >= len) {
            constexpr int PAD = Geo<LEAD>::PAD;
            const long long first = len - a.state_w;         // x index of the oldest carried sample
            const bool staged = first >= o0 - PAD;
            const float* bf = reinterpret_cast<const float*>(buf_cur);
            for (int i = lane; i < a.state_w; i += 32) {
                float v;
                if (staged) {
                    const int pos = static_cast<int>(first - (o0 - PAD)) + i;  // sample position in the buffer
                    v = bf[4 * ((pos >> 2) + (pos >> 5)) + (pos & 3)];
                } else {
                    v = virtual_sample<LEAD, N>(a, xrow, row, first + i);
                }
                a.state_out[row * a.state_pitch + i] = v;
            }
        }

        // Store through the warp's own buffer: each lane parks its 32 consecutive outputs (8 chunks at
        // 9*lane, conflict free), then the warp writes the segment out in lane-interleaved order so that
        // every store instruction covers one contiguous run of global memory (512 B when the row is
        // 16-byte aligned, 128 B otherwise) instead of 32 scattered 16-byte pieces.
        __syncwarp();  // every lane has finished reading its window (windows overlap between lanes)
        {
            float4* park = buf_cur + 9 * lane;
#pragma unroll
            for (int q = 0; q < kR / 4; ++q) park[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
        }
        __syncwarp();
        char* orow = a.out + row * a.out_row_bytes;
        const long long remain = a.out_len - o0;  // outputs of this segment that may be stored
        const bool out_vec = a.out_stride == 4 && ((reinterpret_cast<uintptr_t>(orow) + static_cast<uintptr_t>(o0 * 4)) & 15) == 0;
        if (out_vec && remain >= kSeg) {
            float* dst = reinterpret_cast<float*>(orow) + o0 + 4 * lane;
            const float4* src = buf_cur + lane + (lane >> 3);
            // (phase-shifted first segment: outputs before 0 do not exist; they lie in the first chunks only, kPhase <= 128)
            const int lo = (PT && o0 < 0) ? static_cast<int>(-o0) - 4 * lane : 0;   // (o0 exceeds 32 bits on long signals)
            if (lo <= 0) {
                st_cs_f4(dst, src[0]);
            } else if (lo < 4) {
                const float4 v = src[0];
                if (lo <= 1) dst[1] = v.y;
                if (lo <= 2) dst[2] = v.z;
                dst[3] = v.w;
            }
#pragma unroll
            for (int i = 1; i < kR / 4; ++i) st_cs_f4(dst + 128 * i, src[36 * i]);  // chunk lane + 32 i
        } else {
            const int lim = remain < kSeg ? static_cast<int>(remain) : kSeg;
            const int lo = (PT && o0 < 0) ? static_cast<int>(-o0) : 0;
            if (PT && out_vec) store_cut(buf_cur + lane + (lane >> 3), reinterpret_cast<float*>(orow) + o0 + 4 * lane, lo - 4 * lane, lim - 4 * lane);
            else store_scalar(reinterpret_cast<const float*>(buf_cur) + lane, orow + (o0 + lane) * a.out_stride, a.out_stride, lo - lane, lim - lane);
        }
        if (has_tail) {
            // one more output per lane, o0 + kSeg + lane, from the samples staged behind the segment (they lie past
            // the parked outputs).  Same operation order as compute_fast / the exact flavours: bit-identical to what
            // a further segment would have produced.
            const long long ot = o0 + kSeg + lane;
            if (ot < a.out_len) {
                const float* bf = reinterpret_cast<const float*>(buf_cur);
                auto xs = [&](int k) { const int pos = kSeg + lane + k + DELTA; return bf[pos + 4 * (pos >> 5)]; };
                float v;
                if constexpr (ARITH == ARITH_FAST) {
                    // fully unrolled: the weights become constant-bank operands and the shared-memory address of tap k is
                    // base + k + 4 * ((r + k) >> 5) (one pad chunk per 32 samples; r = the lane's position inside its group)
                    const int q0 = DELTA + lane, r = q0 & 31;
                    const float* b0 = bf + (kSeg + q0) + 4 * ((kSeg + q0) >> 5);
                    v = 0.0f;
                    static_for<WS>([&](auto kc) {
                        constexpr int k = decltype(kc)::value;
                        v = fmaf(k == 0 ? W.ws_first : W.pw[k].x, b0[k + 4 * ((r + k) >> 5)], v);
                    });
                } else {
                    v = __fmul_rn(dot_ordered<WS, ARITH>([&](int k) { return W.w[k]; }, xs), a.scale);
                }
                if (trail_seg && ot >= len - N) v = s_edge[warp][kMaxN + (len - 1 - ot)];
                *reinterpret_cast<float*>(orow + ot * a.out_stride) = v;
            }
        }
        }  // o0 < len
        __syncwarp();  // all lanes are done with buf_cur and s_edge[warp] before the refill
        row = nrow; o0 = no0; xrow = nxrow; t = nt;
        float4* const tmp = buf_cur; buf_cur = buf_nxt; buf_nxt = tmp;
    }
    cp_async_wait<0>();
}

}  // namespace sg
