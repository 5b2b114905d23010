// sg1d_kernel.cuh -- the 1D Savitzky-Golay stencil for sm_100a.
//
// Replaces the reference's scalar hot loop (src/savgolFilter.c:763-766 calling convolve_ilp
// :547-580, the padded edges :785-801 / :442-535, the polynomial edges :769-784, VALID
// :821-850, strided :877-934) and the steady-state of the stream (src/savgol_stream.c:224-226).
//
// Execution plan (one persistent CTA of 128 threads per resident slot, looping over tiles of
// 4096 outputs of one signal):
//   * cp.async (16 B, L2-only) stages tile+halo into shared memory, double buffered, so the next
//     tile's HBM reads are in flight while the current tile is computed.  The first/last tile
//     of a signal synthesises the virtual pad samples (reflect / periodic / constant / explicit
//     halo / carried stream history) while staging, so the compute loop is boundary-agnostic
//     (identity Q6 of SURVEY.md: padded modes == VALID over the padded signal).
//   * shared layout: one 16-byte pad chunk after every 8 chunks -> a thread's 32-sample stride
//     becomes 9 chunks (odd), every LDS.128 of the sliding window is bank-conflict free and all
//     offsets stay compile-time immediates.
//   * each thread produces 32 consecutive outputs from a register sliding window.  Arithmetic is
//     packed fp32 (FFMA2, fma.rn.f32x2): accumulator pairs (out[j],out[j+1]) need the sample pair
//     (x[j+k],x[j+k+1]), which is an aligned register pair only when j+k is even; so taps are split
//     by parity into two accumulator sets, one holding pairs that start at even outputs and one
//     holding pairs that start at odd outputs, merged with one add per output at the end.  Weights
//     are uniform-register operands broadcast to both halves.
//   * polynomial edges: warp 0 / warp 1 evaluate the n leading / trailing outputs from the
//     transposed edge table (one lane per output) and the owners patch them in before the store.
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

// Stage one row slot: shared position 4c <-> x index o0 - PAD + 4c.  `p` is the thread's index
// inside the slot's group of `tpr` threads.  Everything is asynchronous (cp.async), so no thread
// waits on a global load here:
//   * chunks whose four samples exist are copied 16 bytes at a time (4 x 4 bytes when the row is
//     misaligned or strided),
//   * the remaining elements -- virtual pad samples, ragged ends, alignment slack -- are spread one
//     element per thread over the slot (the "edge path"): each thread maps its element to the
//     address the boundary rule designates and copies 4 bytes, or stores 0.
template <int LEAD, int N>
__device__ __forceinline__ void stage_slot(float4* slot_buf, const Args1D& a, long long row, long long o0, int nch,
                                           int p, int tpr)
{
    constexpr int PAD = Geo<LEAD>::PAD;
    const char* xrow = a.in + row * a.in_row_bytes;
    const long long xi0 = o0 - PAD;
    const int c_lo = o0 >= PAD ? 0 : static_cast<int>((PAD - o0) >> 2);
    const long long chi = (a.len - xi0) >> 2;  // chunks whose four samples all exist end here
    const int c_hi = static_cast<int>(chi < nch ? chi : nch);
    const char* src0 = xrow + xi0 * a.in_stride;  // only dereferenced for chunks inside [c_lo, c_hi)
    const bool vec_ok = (a.in_stride == 4) && ((reinterpret_cast<uintptr_t>(src0) & 15) == 0);
    float4* dst0 = slot_buf + p + (p >> 3);
    const int dstep = tpr + (tpr >> 3);  // (c + tpr) + ((c + tpr) >> 3) - (c + (c >> 3)), tpr % 8 == 0
    if (vec_ok) {
#pragma unroll
        for (int it = 0; it < 9; ++it) {
            const int c = p + it * tpr;
            if (c >= c_lo && c < c_hi) cp_async16(dst0 + it * dstep, src0 + static_cast<long long>(c) * 16);
        }
    } else {
        const long long st = a.in_stride;
#pragma unroll 1
        for (int it = 0; it < 9; ++it) {
            const int c = p + it * tpr;
            if (c >= c_lo && c < c_hi) {
                float* d = reinterpret_cast<float*>(dst0 + it * dstep);
                const char* sp = src0 + 4LL * c * st;
                cp_async4(d, sp); cp_async4(d + 1, sp + st); cp_async4(d + 2, sp + 2 * st); cp_async4(d + 3, sp + 3 * st);
            }
        }
    }
    const int nl = 4 * c_lo, nrest = nl + 4 * (nch - c_hi);
#pragma unroll 1
    for (int q = p; q < nrest; q += tpr) {
        const int el = q < nl ? q : 4 * c_hi + (q - nl);  // element index inside the slot
        const int c = el >> 2;
        float* d = reinterpret_cast<float*>(slot_buf + c + (c >> 3)) + (el & 3);
        const float* sp = sample_address<LEAD, N>(a, xrow, row, xi0 + el);
        if (sp) cp_async4(d, sp);
        else *d = 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// FAST arithmetic: packed FFMA2, parity-split accumulators.
template <int N, int DELTA>
__device__ __forceinline__ void compute_fast(const float4* __restrict__ sb, const W1D& W, float scale, float (&out)[kR])
{
    constexpr int WS = 2 * N + 1;
    constexpr int NCHT = (kR + 2 * N + DELTA + 3) / 4;
    float2 Pe[kR / 2];      // Pe[jj] = partial (out[2jj],   out[2jj+1]) : taps with k+DELTA even
    float2 Po[kR / 2 + 1];  // Po[jj] = partial (out[2jj-1], out[2jj])   : taps with k+DELTA odd
#pragma unroll
    for (int i = 0; i < kR / 2; ++i) Pe[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kR / 2 + 1; ++i) Po[i] = make_float2(0.f, 0.f);

    // static_for: the window is up to 24 chunks x 66 FFMA2; "#pragma unroll" silently gives up on
    // bodies that large (half-windows > 20), and a rolled loop would index weights and accumulators
    // dynamically.  Template expansion keeps every index a compile-time constant.
    static_for<NCHT>([&](auto ci) {
        constexpr int c = decltype(ci)::value;
        const float4 v = sb[c + (c >> 3)];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int s = 4 * c + 2 * h;  // shared sample index of this aligned pair
            const float2 X = h == 0 ? make_float2(v.x, v.y) : make_float2(v.z, v.w);
#pragma unroll
            for (int jj = 0; jj < kR / 2; ++jj) {
                const int k = s - 2 * jj - DELTA;
                if (k >= 0 && k < WS) Pe[jj] = __ffma2_rn(make_float2(W.w[k], W.w[k]), X, Pe[jj]);
            }
#pragma unroll
            for (int jj = 0; jj < kR / 2 + 1; ++jj) {
                const int k = s - (2 * jj - 1) - DELTA;
                if (k >= 0 && k < WS) Po[jj] = __ffma2_rn(make_float2(W.w[k], W.w[k]), X, Po[jj]);
            }
        }
    });
#pragma unroll
    for (int jj = 0; jj < kR / 2; ++jj) {
        out[2 * jj] = (Pe[jj].x + Po[jj].y) * scale;
        out[2 * jj + 1] = (Pe[jj].y + Po[jj + 1].x) * scale;
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
// Tile geometry.  A CTA of 128 threads owns 4096 outputs per iteration, arranged as `rpt` row
// slots of `tpr` threads (tpr = 128, 64 or 32 -> one long-row tile, or 2 / 4 short rows of at
// most 2048 / 1024 samples).  Every slot has its own halo in shared memory.
template <int N, int DELTA>
struct Smem1D {
    static constexpr int nch(int seg) { return (seg + 2 * N + DELTA + 3) / 4; }
    static constexpr int phys(int seg) { return nch(seg) + (nch(seg) >> 3) + 1; }
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int kBufChunks = cmax(phys(4096), cmax(2 * phys(2048), 4 * phys(1024)));
};

template <int N, bool LEAD2N, int ARITH>
__global__ void __launch_bounds__(kThreads, SG_MIN_BLOCKS) sg1d_kernel(const __grid_constant__ W1D W, const __grid_constant__ Args1D a)
{
    constexpr int LEAD = LEAD2N ? 2 * N : N;
    constexpr int DELTA = Geo<LEAD>::DELTA;
    constexpr int WS = 2 * N + 1;
    using SM = Smem1D<N, DELTA>;

    __shared__ float4 s_buf[2][SM::kBufChunks];
    __shared__ float s_edge[4][2 * kMaxN];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int tpr = a.tpr;                 // threads per row slot (32, 64, 128)
    const int slot = tid / tpr;            // row slot of this thread
    const int p = tid - slot * tpr;        // index inside the slot
    const int seg = tpr * kR;              // outputs per slot
    const int slot_chunks = ((seg + 2 * N + DELTA + 3) >> 2);
    const int slot_phys = slot_chunks + (slot_chunks >> 3) + 1;
    const int rpt = kThreads / tpr;
    const unsigned tiles_per_row = static_cast<unsigned>(a.tiles_per_row);

#define SG_LOCATE(tile_, row_, o0_)                                                                  \
    do {                                                                                            \
        if (tiles_per_row == 1) { row_ = static_cast<long long>(tile_) * rpt + slot; o0_ = 0; }    \
        else { const unsigned r_ = (tile_) / tiles_per_row; row_ = r_;                              \
               o0_ = static_cast<long long>((tile_) - r_ * tiles_per_row) * kTile; }                \
    } while (0)
#define SG_CHUNKS_OF(o0_) static_cast<int>(((a.len - (o0_) > seg ? seg : a.len - (o0_)) + 2 * N + DELTA + 3) >> 2)

    const unsigned ntiles = static_cast<unsigned>(a.ntiles);
    unsigned tile = blockIdx.x;
    long long row, o0;
    if (tile < ntiles) {
        SG_LOCATE(tile, row, o0);
        if (row < a.rows) stage_slot<LEAD, N>(s_buf[0] + slot * slot_phys, a, row, o0, SG_CHUNKS_OF(o0), p, tpr);
    }
    cp_async_commit();

    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        SG_LOCATE(tile, row, o0);
        const bool active = row < a.rows;
        const char* xrow = a.in + row * a.in_row_bytes;

        // prefetch the next tile of this CTA into the other buffer
        const unsigned nxt = tile + gridDim.x;
        if (nxt < ntiles) {
            long long nrow, no0;
            SG_LOCATE(nxt, nrow, no0);
            if (nrow < a.rows) stage_slot<LEAD, N>(s_buf[(it + 1) & 1] + slot * slot_phys, a, nrow, no0, SG_CHUNKS_OF(no0), p, tpr);
        }
        cp_async_commit();

        // polynomial edge outputs of this slot's row (global reads, independent of the staged tile):
        // the slot's first warp evaluates the leading edge, its second warp (or the same one when the
        // slot is a single warp) the trailing edge, one lane per output.
        const bool lead_tile = active && a.edge_lead && o0 < N;
        const bool trail_tile = active && a.edge_trail && (o0 + seg > a.len - N);
        const int warp_in_slot = p >> 5;
        if (lead_tile && warp_in_slot == 0 && lane < N) {
            // out[e] = scale * sum_k E[e][k] * x[2n-k]   ref: src/savgolFilter.c:773-777, 593-623
            const float s = dot_ordered<WS, ARITH>([&](int k) { return a.edge_t[k * 32 + lane]; },
                                                   [&](int k) { return ld_sample(xrow, a.in_stride, 2 * N - k); });
            s_edge[slot][lane] = ARITH == ARITH_FAST ? s * a.scale : __fmul_rn(s, a.scale);
        }
        if (trail_tile && warp_in_slot == (tpr > 32 ? 1 : 0) && lane < N) {
            // out[len-1-e] = scale * sum_k E[e][k] * x[len-ws+k]   ref: src/savgolFilter.c:780-784
            const long long base = a.len - WS;
            const float s = dot_ordered<WS, ARITH>([&](int k) { return a.edge_t[k * 32 + lane]; },
                                                   [&](int k) { return ld_sample(xrow, a.in_stride, base + k); });
            s_edge[slot][kMaxN + lane] = ARITH == ARITH_FAST ? s * a.scale : __fmul_rn(s, a.scale);
        }

        cp_async_wait<1>();
        __syncthreads();

        const float4* sb = s_buf[it & 1] + slot * slot_phys + 9 * p;
        float out[kR];
        if constexpr (ARITH == ARITH_FAST) compute_fast<N, DELTA>(sb, W, a.scale, out);
        else compute_exact<N, DELTA, ARITH>(sb, W, a.scale, out);

        const long long o = o0 + static_cast<long long>(kR) * p;
        if (lead_tile || trail_tile) {
#pragma unroll
            for (int j = 0; j < kR; ++j) {
                const long long oj = o + j;
                if (lead_tile && oj < N) out[j] = s_edge[slot][oj];
                else if (trail_tile && oj >= a.len - N && oj < a.len) out[j] = s_edge[slot][kMaxN + (a.len - 1 - oj)];
            }
        }

        if (active) {
            // stream: hand the last state_w samples of [lead pad | x] to the next chunk
            if (a.state_out != nullptr && o0 + seg >= a.len)
                for (int i = p; i < a.state_w; i += tpr)
                    a.state_out[row * a.state_pitch + i] = virtual_sample<LEAD, N>(a, xrow, row, a.len - a.state_w + i);

            char* orow = a.out + row * a.out_row_bytes;
            if (a.out_stride == 4 && o + kR <= a.out_len && ((reinterpret_cast<uintptr_t>(orow) + static_cast<uintptr_t>(o * 4)) & 15) == 0) {
                float* dst = reinterpret_cast<float*>(orow) + o;
#pragma unroll
                for (int q = 0; q < kR / 4; ++q)
                    st_cs_f4(dst + 4 * q, make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]));
            } else {
#pragma unroll
                for (int j = 0; j < kR; ++j)
                    if (o + j < a.out_len) *reinterpret_cast<float*>(orow + (o + j) * a.out_stride) = out[j];
            }
        }
        __syncthreads();  // everyone is done with s_buf[it&1] and s_edge before they are refilled
    }
    cp_async_wait<0>();
#undef SG_LOCATE
#undef SG_CHUNKS_OF
}

}  // namespace sg
