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
#include "sg_common.cuh"

namespace sg {

template <int LEAD>
struct Geo {
    static constexpr int PAD = (LEAD + 3) & ~3;   // shared position 0 <-> x index o0 - PAD (keeps 16 B alignment)
    static constexpr int DELTA = PAD - LEAD;      // thread t, output j, tap k reads shared sample 32t + j + k + DELTA
};

__device__ __forceinline__ float ld_sample(const char* xrow, long long stride, long long i)
{
    return *reinterpret_cast<const float*>(xrow + i * stride);
}

// Value of x-index xi (may lie outside [0,len)) of row `row`: real sample, explicit halo, or
// boundary synthesis.  ref: src/savgolFilter.c:442-482 (64-bit clean).
template <int LEAD, int N>
__device__ __forceinline__ float virtual_sample(const Args1D& a, const char* xrow, long long row, long long xi)
{
    const long long len = a.len;
    if (xi >= 0 && xi < len) return ld_sample(xrow, a.in_stride, xi);
    long long idx;
    if (xi < 0) {
        if (a.lhalo) {
            const long long h = LEAD + xi;
            return h >= 0 ? a.lhalo[row * a.lhalo_pitch + h] : 0.0f;
        }
        switch (a.mode) {
            case MODE_REFLECT: idx = -xi - 1; if (idx >= len) idx = len - 1; break;
            case MODE_PERIODIC: idx = ((xi % len) + len) % len; break;
            case MODE_CONSTANT: idx = 0; break;
            default: return 0.0f;
        }
    } else {
        if (a.rhalo) {
            const long long h = xi - len;
            return h < N ? a.rhalo[row * a.rhalo_pitch + h] : 0.0f;
        }
        switch (a.mode) {
            case MODE_REFLECT: idx = 2 * len - xi - 1; if (idx < 0) idx = 0; break;
            case MODE_PERIODIC: idx = xi % len; break;
            case MODE_CONSTANT: idx = len - 1; break;
            default: return 0.0f;
        }
    }
    return ld_sample(xrow, a.in_stride, idx);
}

// Stage `nch` 16-byte chunks of tile (row, o0) into `buf` (padded layout).
template <int LEAD, int N>
__device__ __forceinline__ void stage_tile(float4* buf, const Args1D& a, long long row, long long o0, int nch)
{
    constexpr int PAD = Geo<LEAD>::PAD;
    const char* xrow = a.in + row * a.in_row_bytes;
    const long long xi0 = o0 - PAD;
    const bool contiguous = (a.in_stride == 4);
    const bool vec_ok = contiguous && (((reinterpret_cast<uintptr_t>(xrow) + static_cast<uintptr_t>(xi0 * 4)) & 15) == 0);
    for (int c = threadIdx.x; c < nch; c += kThreads) {
        const long long xi = xi0 + 4LL * c;
        float4* dst = buf + c + (c >> 3);
        if (xi >= 0 && xi + 3 < a.len) {
            if (vec_ok) {
                cp_async16(dst, xrow + xi * 4);
            } else {
                float* d = reinterpret_cast<float*>(dst);
#pragma unroll
                for (int e = 0; e < 4; ++e) cp_async4(d + e, xrow + (xi + e) * a.in_stride);
            }
        } else {
            float4 v;
            v.x = virtual_sample<LEAD, N>(a, xrow, row, xi);
            v.y = virtual_sample<LEAD, N>(a, xrow, row, xi + 1);
            v.z = virtual_sample<LEAD, N>(a, xrow, row, xi + 2);
            v.w = virtual_sample<LEAD, N>(a, xrow, row, xi + 3);
            *dst = v;
        }
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

#pragma unroll
    for (int c = 0; c < NCHT; ++c) {
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
    }
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
template <int N, bool LEAD2N, int ARITH>
__global__ void __launch_bounds__(kThreads) sg1d_kernel(const __grid_constant__ W1D W, const __grid_constant__ Args1D a)
{
    constexpr int LEAD = LEAD2N ? 2 * N : N;
    constexpr int PAD = Geo<LEAD>::PAD;
    constexpr int DELTA = Geo<LEAD>::DELTA;
    constexpr int WS = 2 * N + 1;
    constexpr int NCH_TILE = (kTile + 2 * N + DELTA + 3) / 4;
    constexpr int PHYS = NCH_TILE + (NCH_TILE >> 3) + 1;

    __shared__ float4 s_buf[2][PHYS];
    __shared__ float s_edge[2 * kMaxN];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    auto tile_chunks = [&](long long o0) -> int {
        long long nout = a.len - o0;
        if (nout > kTile) nout = kTile;
        return static_cast<int>((nout + 2 * N + DELTA + 3) >> 2);
    };

    long long tile = blockIdx.x;
    if (tile < a.ntiles) {
        const long long row = tile / a.tiles_per_row, o0 = (tile % a.tiles_per_row) * kTile;
        stage_tile<LEAD, N>(s_buf[0], a, row, o0, tile_chunks(o0));
    }
    cp_async_commit();

    for (int it = 0; tile < a.ntiles; ++it, tile += gridDim.x) {
        const long long row = tile / a.tiles_per_row;
        const long long o0 = (tile % a.tiles_per_row) * kTile;
        const char* xrow = a.in + row * a.in_row_bytes;

        // prefetch the next tile of this CTA into the other buffer
        const long long nxt = tile + gridDim.x;
        if (nxt < a.ntiles) {
            const long long nrow = nxt / a.tiles_per_row, no0 = (nxt % a.tiles_per_row) * kTile;
            stage_tile<LEAD, N>(s_buf[(it + 1) & 1], a, nrow, no0, tile_chunks(no0));
        }
        cp_async_commit();

        // polynomial edge outputs of this tile (global reads, independent of the staged tile)
        const bool lead_tile = a.edge_lead && o0 < N;
        const bool trail_tile = a.edge_trail && (o0 + kTile > a.len - N);
        if (lead_tile && warp == 0 && lane < N) {
            // out[e] = scale * sum_k E[e][k] * x[2n-k]   ref: src/savgolFilter.c:773-777, 593-623
            const float s = dot_ordered<WS, ARITH>([&](int k) { return a.edge_t[k * 32 + lane]; },
                                                   [&](int k) { return ld_sample(xrow, a.in_stride, 2 * N - k); });
            s_edge[lane] = ARITH == ARITH_FAST ? s * a.scale : __fmul_rn(s, a.scale);
        }
        if (trail_tile && warp == 1 && lane < N) {
            // out[len-1-e] = scale * sum_k E[e][k] * x[len-ws+k]   ref: src/savgolFilter.c:780-784
            const long long base = a.len - WS;
            const float s = dot_ordered<WS, ARITH>([&](int k) { return a.edge_t[k * 32 + lane]; },
                                                   [&](int k) { return ld_sample(xrow, a.in_stride, base + k); });
            s_edge[kMaxN + lane] = ARITH == ARITH_FAST ? s * a.scale : __fmul_rn(s, a.scale);
        }

        cp_async_wait<1>();
        __syncthreads();

        const float4* sb = s_buf[it & 1] + 9 * tid;
        float out[kR];
        if constexpr (ARITH == ARITH_FAST) compute_fast<N, DELTA>(sb, W, a.scale, out);
        else compute_exact<N, DELTA, ARITH>(sb, W, a.scale, out);

        const long long o = o0 + static_cast<long long>(kR) * tid;
        if (lead_tile || trail_tile) {
#pragma unroll
            for (int j = 0; j < kR; ++j) {
                const long long oj = o + j;
                if (lead_tile && oj < N) out[j] = s_edge[oj];
                else if (trail_tile && oj >= a.len - N && oj < a.len) out[j] = s_edge[kMaxN + (a.len - 1 - oj)];
            }
        }

        // stream: hand the last state_w samples of [lead pad | x] to the next chunk
        if (a.state_out != nullptr && o0 + kTile >= a.len && tid < a.state_w) {
            const long long xi = a.len - a.state_w + tid;
            a.state_out[row * a.state_pitch + tid] = virtual_sample<LEAD, N>(a, xrow, row, xi);
        }

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
        __syncthreads();  // everyone is done with s_buf[it&1] and s_edge before they are refilled
    }
    cp_async_wait<0>();
}

}  // namespace sg
