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

// Stage one segment (kSeg outputs of one row, plus halo) into a warp's buffer: shared position 4c
// <-> x index o0 - PAD + 4c.  Everything is asynchronous (cp.async), so no lane waits on a global
// load here:
//   * chunks whose four samples exist are copied 16 bytes at a time (4 x 4 bytes when the row is
//     misaligned or strided): 8 full passes of the warp plus a tail,
//   * the remaining elements -- virtual pad samples, ragged ends, alignment slack -- form the
//     "edge path": one element per lane, each lane maps its element to the address the boundary
//     rule designates and copies 4 bytes, or stores 0.
template <int LEAD, int N>
__device__ __forceinline__ void stage_segment(float4* buf, const Args1D& a, const char* xrow, long long row,
                                              long long o0, int nch, int lane)
{
    constexpr int PAD = Geo<LEAD>::PAD;
    const long long xi0 = o0 - PAD;
    const int c_lo = o0 >= PAD ? 0 : static_cast<int>((PAD - o0) >> 2);
    const long long chi = (a.len - xi0) >> 2;  // chunks whose four samples all exist end here
    const int c_hi = static_cast<int>(chi < nch ? chi : nch);
    const char* src0 = xrow + xi0 * a.in_stride;  // only dereferenced for chunks inside [c_lo, c_hi)
    const bool vec_ok = (a.in_stride == 4) && ((reinterpret_cast<uintptr_t>(src0) & 15) == 0);
    float4* dst0 = buf + lane + (lane >> 3);  // chunk c = lane + 32*it lives at c + (c >> 3) = dst0 + 36*it
    if (vec_ok) {
        const char* s = src0 + lane * 16;
#pragma unroll
        for (int it = 0; it < 9; ++it) {
            const int c = lane + 32 * it;
            if (c >= c_lo && c < c_hi) cp_async16(dst0 + 36 * it, s + 512 * it);
        }
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
    const int nl = 4 * c_lo, nrest = nl + 4 * (nch - c_hi);
#pragma unroll 1
    for (int q = lane; q < nrest; q += 32) {
        const int el = q < nl ? q : 4 * c_hi + (q - nl);  // element index inside the segment buffer
        const int c = el >> 2;
        float* d = reinterpret_cast<float*>(buf + c + (c >> 3)) + (el & 3);
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
// Warp-autonomous pipeline.  The unit of work is a SEGMENT: kSeg = 1024 consecutive outputs of one
// row (32 lanes x 32 outputs) together with its halo.  Every warp owns two private segment buffers
// and loops over segments gw, gw + W, gw + 2W, ... (W = warps in the grid): it issues the cp.async
// copies of its NEXT segment, waits for the CURRENT one (cp.async.wait_group + __syncwarp), computes,
// stores.  Warps never synchronise with each other -- there is no __syncthreads in the kernel -- so a
// warp that is waiting on HBM never holds up the FMA pipe of its neighbours, and with ~20 resident
// warps per SM about 80 KB of loads are in flight per SM at any time.
constexpr int kSeg = 32 * kR;  // 1024 outputs per segment

template <int N, int DELTA>
struct Smem1D {
    static constexpr int kSegChunks = (kSeg + 2 * N + DELTA + 3) / 4;
    static constexpr int kSegPhys = kSegChunks + (kSegChunks >> 3) + 1;
};

template <int N, bool LEAD2N, int ARITH>
__global__ void __launch_bounds__(kThreads, SG_MIN_BLOCKS) sg1d_kernel(const __grid_constant__ W1D W, const __grid_constant__ Args1D a)
{
    constexpr int LEAD = LEAD2N ? 2 * N : N;
    constexpr int DELTA = Geo<LEAD>::DELTA;
    constexpr int WS = 2 * N + 1;
    constexpr int kWarps = kThreads / 32;
    using SM = Smem1D<N, DELTA>;

    __shared__ float4 s_buf[kWarps][2][SM::kSegPhys];
    __shared__ float s_edge[kWarps][2 * kMaxN];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned nseg = static_cast<unsigned>(a.ntiles);           // segments in the launch
    const unsigned spr = static_cast<unsigned>(a.tiles_per_row);     // segments per row
    const unsigned stride = gridDim.x * kWarps;
    const long long len = a.len;

    unsigned seg = blockIdx.x * kWarps + warp;
    long long row = 0, o0 = 0;
    const char* xrow = nullptr;
    int nch = 0;
    // segment index -> row, first output, row base, number of 16-byte chunks to stage
#define SG_LOCATE(seg_, row_, o0_, xrow_, nch_)                                                       \
    do {                                                                                             \
        unsigned r_ = (seg_), t_ = 0;                                                                \
        if (spr != 1) { r_ = (seg_) / spr; t_ = (seg_) - r_ * spr; }                                 \
        row_ = r_;                                                                                   \
        o0_ = static_cast<long long>(t_) * kSeg;                                                     \
        xrow_ = a.in + row_ * a.in_row_bytes;                                                        \
        const long long left_ = len - o0_;                                                           \
        nch_ = (static_cast<int>(left_ < kSeg ? left_ : kSeg) + 2 * N + DELTA + 3) >> 2;             \
    } while (0)

    if (seg < nseg) {
        SG_LOCATE(seg, row, o0, xrow, nch);
        stage_segment<LEAD, N>(s_buf[warp][0], a, xrow, row, o0, nch, lane);
    }
    cp_async_commit();

    for (int it = 0; seg < nseg; ++it, seg += stride) {
        // prefetch this warp's next segment into its other buffer
        const unsigned nxt = seg + stride;
        long long nrow = 0, no0 = 0;
        const char* nxrow = nullptr;
        int nnch = 0;
        if (nxt < nseg) {
            SG_LOCATE(nxt, nrow, no0, nxrow, nnch);
            stage_segment<LEAD, N>(s_buf[warp][(it + 1) & 1], a, nxrow, nrow, no0, nnch, lane);
        }
        cp_async_commit();

        // polynomial edge outputs that fall into this segment (global reads, independent of the
        // staged data): one lane per output.  ref: src/savgolFilter.c:769-784
        const bool lead_seg = a.edge_lead && o0 < N;
        const bool trail_seg = a.edge_trail && (o0 + kSeg > len - N);
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

        const float4* sb = s_buf[warp][it & 1] + 9 * lane;
        float out[kR];
        if constexpr (ARITH == ARITH_FAST) compute_fast<N, DELTA>(sb, W, a.scale, out);
        else compute_exact<N, DELTA, ARITH>(sb, W, a.scale, out);

        const long long o = o0 + kR * lane;
        if (lead_seg || trail_seg) {
#pragma unroll
            for (int j = 0; j < kR; ++j) {
                const long long oj = o + j;
                if (lead_seg && oj < N) out[j] = s_edge[warp][oj];
                else if (trail_seg && oj >= len - N && oj < len) out[j] = s_edge[warp][kMaxN + (len - 1 - oj)];
            }
        }

        // stream: hand the last state_w samples of [lead pad | x] to the next chunk
        if (a.state_out != nullptr && o0 + kSeg >= len)
            for (int i = lane; i < a.state_w; i += 32)
                a.state_out[row * a.state_pitch + i] = virtual_sample<LEAD, N>(a, xrow, row, len - a.state_w + i);

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
        __syncwarp();  // all lanes are done with s_buf[warp][it&1] and s_edge[warp] before the refill
        row = nrow; o0 = no0; xrow = nxrow; nch = nnch;
    }
    cp_async_wait<0>();
#undef SG_LOCATE
}

}  // namespace sg
