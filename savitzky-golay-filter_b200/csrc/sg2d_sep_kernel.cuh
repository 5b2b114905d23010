// sg2d_sep_kernel.cuh -- production path of the 2D filter: streaming separable stencil for sm_100a.
// (kernel template; instantiated by sg2d_sep.cu -- generic rank-R factors -- and sg2d_add.cu -- additive
// surfaces W = u(x) + v(y), see below)
//
// Replaces the reference's 4-deep tap loop (src/savgol2d.c:417-452, 374-393).  The weight table is
// factorised on the host as W[y][x] = sum_{r<R} col_r[y] * row_r[x] (factor2d.cpp, R <= 4, R = 2 for
// the order-2/3 smoothing filters), all row factors even or all odd in x.
//
// Execution plan -- warp-autonomous, like the 1D kernel; no __syncthreads anywhere:
//   * work item = (image, band of <= 512 output rows, strip of 32*RX output columns), handed out by an
//     atomic ticket (edge strips first); a warp walks DOWN its strip two input rows per step.  Rows are
//     staged by cp.async into a private ring of 8 row buffers, 6 rows ahead of the rows being consumed
//     (~3.4 KB in flight per warp); the row index is mapped by the boundary rule (clamp / half-sample
//     reflect), the few pad columns of the first / last strip are copied element by element from the
//     column the rule maps them to, so the compute is boundary agnostic.
//   * ROW PASS: each lane owns RX consecutive columns; from a register window of the staged row it
//     forms the folded sums s_k = x[c+k] +/- x[c-k] once and evaluates the R row factors on them
//     (n adds + R*(n+1) FMAs per pixel instead of R*(2n+1) MACs; the FMAs packed over column pairs).
//   * COLUMN PASS, in registers: the lane keeps the partially accumulated output rows of its columns.
//     The R values just produced are scattered into them with packed FFMA2 (column pairs packed,
//     weight col[wy] broadcast from a uniform register), the two oldest rows are complete and are
//     stored (one 512-byte store per warp and row).  Half-windows <= 8: a static ring of 2n+2 rows
//     whose indices are compile-time constants per ring phase (one copy of the column pass per
//     phase, selected by a switch; no register ever moves).  Larger half-windows: indices static
//     inside blocks of U = 4 rows, a block ends with a register shift.
//   => every input pixel is read from HBM once and from shared memory (RX+2n)/RX times, no
//      intermediate image ever exists, and there is no vertical halo recomputation except the 2n
//      warm-up rows per band.
#pragma once
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "sg2d.h"
#include "sg_common.cuh"
#include "sg1d_kernel.cuh"  // static_for
#include "sg1d_tma.cuh"     // mbarrier / bulk-tensor helpers

namespace sg { extern std::atomic<unsigned long long> g_launches; }

#ifndef SG2D_KU
#define SG2D_KU 4
#endif
#ifndef SG2D_RXW
#define SG2D_RXW 4
#endif
#ifndef SG2D_RX4
#define SG2D_RX4 1
#endif
#ifndef SG2D_WIDE_MINB
#define SG2D_WIDE_MINB 2   // resident CTAs for the widest rank-3/4 kernels (231 registers unconstrained)
#endif
#ifndef SG2D_RING_MAX
#define SG2D_RING_MAX 1500   // FFMA2 in all column-pass copies of the static ring
#endif
#ifndef SG2D_RING
#define SG2D_RING 1
#endif
#ifndef SG2D_MINB2
#define SG2D_MINB2 3
#endif
#ifndef SG2D_MINB
#define SG2D_MINB 4
#endif
#ifndef SG2D_ADD_MINB
#define SG2D_ADD_MINB 4
#endif
#ifndef SG2D_TMA
#define SG2D_TMA 0   // 1: interior work items stage their rows with one bulk-tensor copy per step (experiment, see below)
#endif

namespace sg2d {

namespace {

using sg::cp_async16;
using sg::cp_async4;
using sg::cp_async_commit;
using sg::cp_async_wait;

constexpr int kU = SG2D_KU;         // rows per statically indexed block
#ifndef SG2D_QUAD
#define SG2D_QUAD 1   // 1: rows are staged four at a time (every other step); 0: two at a time (every step)
#endif
constexpr int kStage = SG2D_QUAD ? 4 : 2;       // rows per staging call / cp.async group
constexpr int kRing = SG2D_QUAD ? 16 : 8;       // staged rows per warp
constexpr int kAhead = SG2D_QUAD ? 8 : 6;       // prefetch distance in rows (kRing >= kAhead + kStage)
#ifndef SG2D_BAND
#define SG2D_BAND 512
#endif
constexpr int kBandMax = SG2D_BAND; // output rows per work item (upper bound; the launcher shrinks it for small batches)
constexpr int kWarps = 4;

template <int R>
struct SepW {
    float rc[R];            // row factor, centre tap
    float rk[R][16];        // rk[r][k-1]: weight of s_k = x[c+k] + sx * x[c-k], k = 1..n
    float col[R][33];       // column factor * scale, col[r][wy], wy = 0..2n
    float sx;               // +1 (even in x) / -1 (odd in x)
    float sxo[3];           // multi-output kernels (gradient / Hessian): parity in x of every output's factors
};

// Tensor map of the input images for the bulk-tensor (TMA) staging of interior work items: {cols, rows, images},
// box = two staged rows of one strip (ROWF floats each), no swizzle.
struct alignas(64) Tma2D {
    CUtensorMap pair;
};

// Additive surface W(y,x) = u(x) + v(y) (sg2d_add.cu).  Row weights as pairs for the sample-broadcast FFMA2 form:
// sample x[c+i] is tap i of column c and tap i-1 of column c+1.
struct AddW {
    float2 pu[2 * 16 + 1];  // pu[j] = (u[j], u[j-1]), j = 1..2n  (u[0..2n] = u(-n..n) * scale)
    float u_first, u_last;  // u[0], u[2n]: the two half pairs at the window ends
    float col[2 * 16 + 1];  // v[0..2n] * scale, v[0] = v[2n] = 0
};
template <int R, bool ADD> struct WSel { using type = SepW<R>; };
template <int R> struct WSel<R, true> { using type = AddW; };

__device__ __forceinline__ int map_index(int i, int n, int boundary)
{
    if (boundary == B_REFLECT) {
        if (i < 0) i = -i - 1;
        else if (i >= n) i = 2 * n - i - 1;
    }
    if (i < 0) i = 0;
    else if (i >= n) i = n - 1;
    return i;
}

// Run f(integral_constant<I>) for the I that equals `v` (0 <= v < COUNT).
template <class F, int... I>
__device__ __forceinline__ void static_switch_impl(int v, F&& f, std::integer_sequence<int, I...>)
{
    (void)((v == I && (f(std::integral_constant<int, I>{}), true)) || ...);
}
template <int COUNT, class F>
__device__ __forceinline__ void static_switch(int v, F&& f)
{
    static_switch_impl(v, static_cast<F&&>(f), std::make_integer_sequence<int, COUNT>{});
}

// Stores of a lane whose columns straddle the stored region or whose row is not vector-aligned.
// (values travel by value: a reference parameter of an out-of-line function would force the caller's
// output registers into local memory -- an STL.128 per row on the hot path)
template <int RX>
struct RowVals { float2 v[RX / 2]; };
// Out of line (the hot loop only pays the call): rows that cannot take the 16-byte store -- the ragged first / last
// strip, and every strip of an image whose row pitch is not a multiple of 4 floats (the alignment then rotates from
// row to row).  Lanes that lie inside the stored region still use the widest stores their address allows.
template <int RX>
__device__ __noinline__ void emit_ragged(float* dst_row, RowVals<RX> r, int X, int Xlo, int Xhi)
{
    if (X >= Xlo && X + RX <= Xhi) {
        const bool a8 = (reinterpret_cast<uintptr_t>(dst_row) & 7) == 0;
        if constexpr (RX == 4) {
            if (a8) {
                *reinterpret_cast<float2*>(dst_row) = r.v[0];
                *reinterpret_cast<float2*>(dst_row + 2) = r.v[1];
            } else {
                dst_row[0] = r.v[0].x;
                *reinterpret_cast<float2*>(dst_row + 1) = make_float2(r.v[0].y, r.v[1].x);
                dst_row[3] = r.v[1].y;
            }
            return;
        } else if constexpr (RX == 2) {
            if (a8) *reinterpret_cast<float2*>(dst_row) = r.v[0];
            else { dst_row[0] = r.v[0].x; dst_row[1] = r.v[0].y; }
            return;
        }
    }
#pragma unroll
    for (int j = 0; j < RX; ++j)
        if (X + j >= Xlo && X + j < Xhi) dst_row[j] = (j & 1) ? r.v[j / 2].y : r.v[j / 2].x;
}

__device__ __forceinline__ void cp_async4_s(unsigned smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void cp_async16_s(unsigned smem_dst, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}


// Pad columns of an edge strip, two rows: pad element q of the slot (the nl columns left of the image,
// then those right of it) is copied from the column the boundary rule maps it to.  d = shared address of
// the slot's first float (row t), r0 / r1 = start of the two source rows.  Out of line to keep the main
// loop inside the instruction cache.
__device__ __noinline__ void stage_pads(unsigned d, const float* r0, const float* r1, int row_bytes, int npad, int nl, int xb, int cols,
                                        int boundary, int lane)
{
#pragma unroll 1
    for (int q = lane; q < npad; q += 32) {
        const int x = q < nl ? xb + q : cols + (q - nl);
        const int m = map_index(x, cols, boundary);
        const unsigned dd = d + 4 * (x - xb);
        cp_async4_s(dd, r0 + m);
        cp_async4_s(dd + row_bytes, r1 + m);
    }
}

// Two rows of an EDGE STRIP (first / last strip of the image), or of an image whose rows are not 16-byte
// aligned, into two consecutive ring slots: chunks that lie inside the image and are aligned are copied
// 16 bytes at a time, everything else element by element (pad columns from the column the boundary rule
// maps them to).  r0 / r1 = start of the (already mapped) source rows, d = this
// lane's shared address in the first slot, xb = image column of the slot's first float.  Out of line:
// 1/16 of the items of a 4096^2 image take it, it must not bloat the main loop.
template <int ROWCH, int ROWF>
__device__ __noinline__ void stage_edge_pair(unsigned d, const float* r0, const float* r1, int xb, int cols, int boundary, int lane)
{
#pragma unroll
    for (int c0 = 0; c0 < ROWCH; c0 += 32) {
        const int c = c0 + lane;
        if (c0 + 32 <= ROWCH || c < ROWCH) {
            const int xin = xb + 4 * c;
            const bool inside = xin >= 0 && xin + 3 < cols;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const float* r = rr ? r1 : r0;
                const unsigned dd = d + rr * (ROWF * 4) + 16 * c0;
                if (inside && (reinterpret_cast<uintptr_t>(r + xin) & 15) == 0) {
                    cp_async16_s(dd, r + xin);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) cp_async4_s(dd + 4 * e, r + (inside ? xin + e : map_index(xin + e, cols, boundary)));
                }
            }
        }
    }
}

// Two rows of an INTERIOR strip of an image whose rows are not 16-byte aligned (odd width / stride, offset views): every
// staged column exists, so the rows travel as lane-consecutive 4-byte copies -- each instruction still covers one
// contiguous 128-byte run -- with no per-element boundary logic.  d = shared address of this lane's first float of the
// first row, r0 / r1 = this lane's first source float.  Out of line: the aligned path must not pay for it.
template <int ROWF>
__device__ __noinline__ void stage_rows4(unsigned d, const float* r0, const float* r1, int lane)
{
#pragma unroll
    for (int e0 = 0; e0 < ROWF; e0 += 32)
        if (e0 + 32 <= ROWF || lane < ROWF - e0) {
            cp_async4_s(d + 4 * e0, r0 + e0);
            cp_async4_s(d + ROWF * 4 + 4 * e0, r1 + e0);
        }
}

// NO > 1: the R factors form NO groups of R / NO, one group per OUTPUT image (the components of a gradient or a
// Hessian, sg2d_multi.cu): every input row is staged and windowed once, each group has its own parity in x and its
// own accumulator ring, and a completed row is stored to a.out, a.out1 (, a.out2).
template <int N, int R, int RX, bool ADD, int NO = 1>
__global__ void __launch_bounds__(kWarps * 32, NO > 1 ? ((N <= 2 && R <= 4 && NO == 2) ? 4 : 3) : ADD ? SG2D_ADD_MINB : RX >= 4 ? ((N <= 6 || (N == 7 && (R == 2 || R == 3))) ? SG2D_MINB : 3) : ((N >= 15 && R >= 3) ? SG2D_WIDE_MINB : SG2D_MINB2)) sep_kernel(const __grid_constant__ typename WSel<R, ADD>::type w,
                                                                            const __grid_constant__ Args2D a, const __grid_constant__ Tma2D maps)
{
    static_assert(!ADD || (R == 1 && (RX == 4 || RX == 2) && N <= 16), "additive variant: one weight set, static ring");
    static_assert(NO == 1 || (!ADD && R % NO == 0 && NO <= 3 && N <= 8), "multi-output variant: generic factors, static ring");
    constexpr int RPO = R / NO;                 // factors per output
    constexpr int TW = 32 * RX;                 // output columns per strip
    constexpr int PADX = (N + 3) & ~3;          // staged row starts PADX columns left of the strip (16 B aligned)
    constexpr int DX = PADX - N;
    constexpr int ROWF = TW + 2 * PADX;         // floats per staged row
    constexpr int ROWCH = ROWF / 4;             // 16-byte chunks per staged row
    // static accumulator ring (one column pass per ring phase) vs shifting blocks: the ring needs
    // (n+1) copies of the column pass, which must stay inside the 32 KB instruction cache
    // (17x17 rank-4: 2448 FFMA2, measured 9 % slower than the shifting blocks)
    // (2 columns per lane, half-windows 9-16: the ring measured 9 % slower than the blocks at 25x25)
    constexpr bool RING = ADD || NO > 1 || (SG2D_RING && N <= 8 && (N + 1) * (2 * N + 1) * R * RX <= SG2D_RING_MAX);
    constexpr int NA = RING ? 2 * N + 2 : 2 * N + kU;   // output rows in flight per column
    constexpr int WIN = RX + DX + 2 * N;        // floats of the row window a lane touches
    constexpr int VW = RX >= 4 ? 4 : 2;         // floats per shared load
    constexpr int NV = (WIN + VW - 1) / VW;

    // Bulk-tensor (TMA) staging of interior work items -- compiled in with -DSG2D_TMA=1 only.  Measured on C4
    // (profiles/r2_c4_tma_experiment.txt): 0.647 of the HBM roofline with it, 0.675 without -- the ELECT / R2UR sequence
    // ptxas wraps around each UTMALDG plus the mbarrier wait cost more issue slots than the four LDGSTS per lane they
    // replace, and the kernel is issue bound.  It needs 128-byte aligned destinations: two rows of a step are
    // 2 * ROWF floats apart, so only the kernels with ROWF = 144 (half-windows 5..8) qualify.
    constexpr bool TMA_OK = SG2D_TMA && (2 * ROWF * 4) % 128 == 0;
    __shared__ __align__(128) float s_ring[kWarps][kRing][ROWF];
    __shared__ __align__(8) unsigned long long s_mbar[kWarps][kRing / 2];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float(*ring)[ROWF] = s_ring[warp];
    const unsigned mbar0 = sg::smem_u32(&s_mbar[warp][0]);
    unsigned par_bits = 0;   // per mbarrier: the phase parity its next completion will have
    if (TMA_OK && a.use_tma) {
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < kRing / 2; ++q) sg::mbar_init(mbar0 + 8 * q, 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
            sg::fence_proxy_async();
        }
        __syncwarp();
    }

    // 16-byte copies need aligned rows, whole chunks a width that is a multiple of 4; other images take the
    // out-of-line per-chunk path for every strip
    const bool simple_rows = ((reinterpret_cast<uintptr_t>(a.in) & 15) | (a.in_stride & 3) | (a.in_image_pitch & 3) | (a.cols & 3)) == 0;
    const int Ylo = a.cy, Yhi = a.cy + a.out_rows;      // stored region in full-image coordinates
    const int Xlo = a.cx, Xhi = a.cx + a.out_cols;
    const int strips = (Xhi + TW - 1) / TW;
    const int kBand = a.band_rows;
    const int bands = (a.out_rows + kBand - 1) / kBand;
    const long long per_img = static_cast<long long>(strips) * bands;
    const long long items = per_img * a.n_images;

    // Work items are handed out dynamically (one atomic per item): item times differ (edge strips, first /
    // last bands, L2 hits on shared halo columns); a static round-robin measured 12 % slower on config 4.
    for (;;) {
        unsigned ticket = 0;
        if (lane == 0) ticket = atomicAdd(a.counter, 1u);
        const long long item = __shfl_sync(0xffffffffu, ticket, 0);
        if (item >= items) break;
        // longest first: the items of the two edge strips (extra pad-column copies; much slower on the
        // out-of-line path of unaligned / ragged images)
        // are handed out before everything else, so none of them is left for the tail of the launch;
        // the interior strips follow image by image, band by band, neighbours in x back to back (their
        // halo columns meet in L2)
        const int nedge = strips < 2 ? strips : 2;
        const long long edge_items = static_cast<long long>(nedge) * bands * a.n_images;
        long long img;
        int band, strip;
        if (item < edge_items) {
            const long long q = item / nedge;
            strip = (item - q * nedge) ? strips - 1 : 0;
            img = q / bands;
            band = static_cast<int>(q - img * bands);
        } else {
            const int inner = strips - 2;
            const long long q = (item - edge_items) / inner;
            strip = 1 + static_cast<int>((item - edge_items) - q * inner);
            img = q / bands;
            band = static_cast<int>(q - img * bands);
        }
        const int x0 = strip * TW;
        if (x0 + TW <= Xlo) continue;                   // strip entirely left of the stored region (VALID)
        const int Y0 = Ylo + band * kBand;
        const int nrows = min(kBand, Yhi - Y0);
        // Additive kernel: which two input rows form a step decides the order in which an output row's terms are
        // summed.  Steps are aligned to EVEN absolute image rows (one extra warm-up row when the band starts on an
        // odd one), so a pixel rounds the same way whatever band / call computes it (row bands of one image on
        // several GPUs reassemble bit for bit; a.row0 = image row of the buffer's first row).
        const int shift = ADD ? ((Y0 + a.row0) & 1) : 0;
        const int yin0 = Y0 - N - shift;                // image row (buffer coordinates) staged as step row t = 0
        const int steps = nrows + 2 * N + shift;
        const float* in = a.in + img * a.in_image_pitch;
        // virtual output base: element (Y, X) of the full-size result lives at vout + Y*os + X
        float* vout = a.out + img * a.out_image_pitch - static_cast<long long>(a.cy) * a.out_stride - a.cx;

        // ---- staging of input rows ----
        // Rows t, t+1 (t even) go to ring slots t mod 8 and the next one.  Images with 16-byte aligned
        // rows and a width that is a multiple of 4 (chunks are then entirely inside or outside the image):
        //   * chunks inside the image: one or two 16-byte copies per lane and row, predicate fixed per item
        //     (always true for interior strips); interior bands advance a running source pointer, the
        //     first / last band maps the row index (clamp / reflect);
        //   * pad columns of the first / last strip: one element per lane (4-byte copy from the column
        //     the boundary rule maps it to) -- 2 x 8 elements per step for a 15x15 window.
        // Everything else (unaligned rows, ragged widths) takes the out-of-line per-chunk path.
        const int steps2 = (steps + 1) & ~1;   // rows are consumed two per step, see below
        const bool y_in = (yin0 >= 0) && (yin0 + steps2 <= a.rows);
        const int xb = x0 - PADX;                                        // image column of the slot's first float
        const bool x_in = xb >= 0 && xb + ROWF <= a.cols;                // interior strip: no pad columns
        bool pch[(ROWCH + 31) / 32];                                     // this lane's chunk(s) lie inside the image
#pragma unroll
        for (int c0 = 0; c0 < ROWCH; c0 += 32) {
            const int xin = xb + 4 * (c0 + lane);
            pch[c0 / 32] = (c0 + 32 <= ROWCH || lane < ROWCH - c0) && xin >= 0 && xin + 4 <= a.cols;
        }
        float* const ring_lane = &ring[0][0] + 4 * lane;
        const unsigned ring_lane_s = static_cast<unsigned>(__cvta_generic_to_shared(ring_lane));
        // this lane's first chunk of the NEXT row to stage (interior bands)
        const float* src_next = in + static_cast<long long>(yin0) * a.in_stride + xb + 4 * lane;
        const float* const xbase = in + xb + 4 * lane;
        // interior work item (no pad columns, no mapped rows) of an aligned image: one lane issues ONE bulk-tensor copy
        // for the two rows of a step; it lands on the slot pair's mbarrier
        const bool item_tma = TMA_OK && a.use_tma && simple_rows && x_in && y_in;
        auto stage_pair = [&](int t) {
            const int slot = t & (kRing - 1);
            if (item_tma) {
                if (lane == 0) {
                    const unsigned mb = mbar0 + 8 * (slot >> 1);
                    sg::mbar_expect_tx(mb, 2 * ROWF * 4);
                    sg::tma_load_3d(ring_lane_s + slot * (ROWF * 4), &maps.pair, xb, yin0 + t, static_cast<int>(img), mb);
                }
            } else if (simple_rows) {
                const float *s0, *s1;
                if (y_in) {
                    s0 = src_next;
                    s1 = s0 + a.in_stride;
                    src_next = s1 + a.in_stride;
                } else {
                    s0 = xbase + static_cast<long long>(map_index(yin0 + t, a.rows, a.boundary)) * a.in_stride;
                    s1 = xbase + static_cast<long long>(map_index(yin0 + t + 1, a.rows, a.boundary)) * a.in_stride;
                }
                const unsigned d = ring_lane_s + slot * (ROWF * 4);
                if (x_in) {   // interior strip: nothing to decide per lane
#pragma unroll
                    for (int c0 = 0; c0 < ROWCH; c0 += 32)
                        if (c0 + 32 <= ROWCH || lane < ROWCH - c0) {
                            cp_async16_s(d + 16 * c0, s0 + 4 * c0);
                            cp_async16_s(d + ROWF * 4 + 16 * c0, s1 + 4 * c0);
                        }
                } else {
#pragma unroll
                    for (int c0 = 0; c0 < ROWCH; c0 += 32)
                        if (pch[c0 / 32]) {
                            cp_async16_s(d + 16 * c0, s0 + 4 * c0);
                            cp_async16_s(d + ROWF * 4 + 16 * c0, s1 + 4 * c0);
                        }
                    const int nl = xb < 0 ? -xb : 0;                                 // pad elements left of the image
                    const int nr = xb + ROWF > a.cols ? xb + ROWF - a.cols : 0;      // ... and right of it
                    stage_pads(d - 16 * lane, s0 - xb - 4 * lane, s1 - xb - 4 * lane, ROWF * 4, nl + nr, nl, xb, a.cols, a.boundary, lane);
                }
            } else if (x_in) {
                stage_rows4<ROWF>(ring_lane_s - 12 * lane + slot * (ROWF * 4),   // ring_lane_s = slot 0 + 16 * lane
                                  in + static_cast<long long>(map_index(yin0 + t, a.rows, a.boundary)) * a.in_stride + xb + lane,
                                  in + static_cast<long long>(map_index(yin0 + t + 1, a.rows, a.boundary)) * a.in_stride + xb + lane, lane);
            } else {
                stage_edge_pair<ROWCH, ROWF>(ring_lane_s + slot * (ROWF * 4),
                                             in + static_cast<long long>(map_index(yin0 + t, a.rows, a.boundary)) * a.in_stride,
                                             in + static_cast<long long>(map_index(yin0 + t + 1, a.rows, a.boundary)) * a.in_stride,
                                             xb, a.cols, a.boundary, lane);
            }
        };
        // Staging schedule: at every kStage-th row the group that holds rows t .. t+kStage-1 must have landed (this lane's
        // cp.async copies -- or the bulk-tensor copies --, then everybody else's), and the rows kAhead ahead are issued
        // into the slots of rows that are fully consumed.  With kStage = 4 the wait / barrier / address bookkeeping runs on
        // every other step only.
        auto wait_rows = [&](int t) {
            if (item_tma) {
#pragma unroll
                for (int u = 0; u < kStage; u += 2)
                    if (t + u < steps2) {
                        const int q = ((t + u) & (kRing - 1)) >> 1;
                        sg::mbar_wait(mbar0 + 8 * q, (par_bits >> q) & 1u);
                        par_bits ^= 1u << q;
                    }
            } else {
                cp_async_wait<kAhead / kStage - 1>();
            }
            __syncwarp();   // ... and everybody else's; the rows before t are fully consumed
        };
        auto stage_rows = [&](int t) {
#pragma unroll
            for (int u = 0; u < kStage; u += 2)
                if (t + u < steps2) stage_pair(t + u);
            cp_async_commit();
        };
        auto advance = [&](int t) {
            if ((t & (kStage - 1)) == 0) {
                wait_rows(t);
                stage_rows(t + kAhead);
            }
        };
        // store side, hoisted: this lane's columns, whether they lie inside the stored region and
        // whether a vector store is legal; the row pointer advances by the output pitch per emitted row
        const int X = x0 + RX * lane;
        float* dst_row = vout + static_cast<long long>(Y0) * a.out_stride + X;
        constexpr int SV = RX >= 4 ? 4 : 2;  // floats per store instruction
        const bool st_vec = X >= Xlo && X + RX <= Xhi && (a.out_stride % SV) == 0 &&
                            (reinterpret_cast<uintptr_t>(dst_row) & (4 * SV - 1)) == 0;

        // Rows are consumed two per step: one wait / sync / loop overhead per two rows, and every
        // column weight (a uniform register that has to be re-loaded each step, 46 weights do not fit
        // the uniform register file next to everything else) is used for both rows.  An odd row count
        // is padded with one extra (boundary-mapped) row whose contributions are never emitted.
        __syncwarp();  // the previous item's last reads of the ring are done
#pragma unroll 1
        for (int t = 0; t < kAhead; t += kStage) stage_rows(t);

        // acc[0][jp][i]: partially accumulated output rows of the column pair (2jp, 2jp+1) of this lane.
        //   RING:  output row y (band-local, y = t - wy) lives in slot (y mod NA), NA = 2n+2; the main loop
        //          is unrolled over a full period of the ring (n+1 steps), so every index is static and
        //          no register ever moves.
        //   else:  slot i = output row (block base - 2n + i); blocks of kU rows end with a register shift.
        float2 acc[NO][RX / 2][NA];
#pragma unroll
        for (int o = 0; o < NO; ++o)
#pragma unroll
            for (int jp = 0; jp < RX / 2; ++jp)
#pragma unroll
                for (int i = 0; i < NA; ++i) acc[o][jp][i] = make_float2(0.f, 0.f);

        auto row_pass = [&](int t, float2 (&hp)[R][RX / 2]) {
          if constexpr (!ADD) {
            float xs[NV * VW];
            const float* rowp = ring_lane + (t & (kRing - 1)) * ROWF + (RX - 4) * lane;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                if constexpr (VW == 4) {
                    const float4 q = *reinterpret_cast<const float4*>(rowp + 4 * v);
                    xs[4 * v] = q.x; xs[4 * v + 1] = q.y; xs[4 * v + 2] = q.z; xs[4 * v + 3] = q.w;
                } else {
                    const float2 q = *reinterpret_cast<const float2*>(rowp + 2 * v);
                    xs[2 * v] = q.x; xs[2 * v + 1] = q.y;
                }
            }
#pragma unroll
            for (int jp = 0; jp < RX / 2; ++jp) {
                const int c0 = 2 * jp + DX + N;
                const float2 ctr = make_float2(xs[c0], xs[c0 + 1]);
#pragma unroll
                for (int r = 0; r < R; ++r) hp[r][jp] = __fmul2_rn(make_float2(w.rc[r], w.rc[r]), ctr);
            }
#pragma unroll
            for (int k = 1; k <= N; ++k)
#pragma unroll
                for (int jp = 0; jp < RX / 2; ++jp) {
                    const int c0 = 2 * jp + DX + N;
#pragma unroll
                    for (int o = 0; o < NO; ++o) {
                        const float sxv = NO == 1 ? w.sx : w.sxo[o];
                        const float2 sk = make_float2(fmaf(sxv, xs[c0 - k], xs[c0 + k]),      // +/-1 multiply is exact
                                                      fmaf(sxv, xs[c0 + 1 - k], xs[c0 + 1 + k]));
#pragma unroll
                        for (int r = o * RPO; r < (o + 1) * RPO; ++r)
                            hp[r][jp] = __ffma2_rn(make_float2(w.rk[r][k - 1], w.rk[r][k - 1]), sk, hp[r][jp]);
                    }
                }
          }
        };
        // ADDITIVE surface W(y,x) = u(x) + v(y) (order <= 3 smoothing and the even/even derivatives of it):
        //   out = sum_y [ A(y) + v(y) * B(y) ],   A = sum_x u(x) x[.],  B = sum_x x[.]  (row BOX sum).
        // Row pass: the folded sums feed A (n+1 packed FMAs per column pair) and, for the lane's first column,
        // B (n adds); the other columns' box sums slide: B(c+1) = B(c) + (x[c+1+n] - x[c-n]) -- at most
        // RX-1 = 3 steps from a directly summed value, so no drift can build up.
        auto row_pass_add = [&](int t, float2 (&hA)[RX / 2], float2 (&hB)[RX / 2]) {
          if constexpr (ADD) {
            float xs[NV * VW];
            const float* rowp = ring_lane + (t & (kRing - 1)) * ROWF + (RX - 4) * lane;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                if constexpr (VW == 4) {
                    const float4 q = *reinterpret_cast<const float4*>(rowp + 4 * v);
                    xs[4 * v] = q.x; xs[4 * v + 1] = q.y; xs[4 * v + 2] = q.z; xs[4 * v + 3] = q.w;
                } else {
                    const float2 q = *reinterpret_cast<const float2*>(rowp + 2 * v);
                    xs[2 * v] = q.x; xs[2 * v + 1] = q.y;
                }
            }
            constexpr int C0 = DX + N;   // window index of the lane's first column
            // A: sample-broadcast form (as in the 1D kernel) -- xs[c + i] is tap i of column c and tap i - 1 of
            // column c + 1, one FFMA2 with the weight pair (u[i], u[i-1]); 2n FFMA2 + 2 FFMA per column pair
#pragma unroll
            for (int jp = 0; jp < RX / 2; ++jp) {
                const int c = C0 + 2 * jp;
                float2 acc = make_float2(w.u_first * xs[c - N], 0.0f);
#pragma unroll
                for (int j = 1; j <= 2 * N; ++j) acc = __ffma2_rn(w.pu[j], make_float2(xs[c - N + j], xs[c - N + j]), acc);
                acc.y = fmaf(w.u_last, xs[c + N + 1], acc.y);
                hA[jp] = acc;
            }
            // B: box sum of the first column as a tree over the aligned register pairs of the window, the other
            // columns slide: B(c+1) = B(c) + (x[c+1+n] - x[c-n]) -- at most RX-1 = 3 steps from a directly summed
            // value, so no drift can build up
            constexpr int LO = C0 - N, HI = C0 + N;             // window of column 0: xs[LO..HI]
            constexpr int PL = (LO + 1) & ~1, PH = (HI + 1) & ~1;   // aligned pairs (xs[q], xs[q+1]), q = PL, PL+2, ... < PH
            constexpr int NP = (PH - PL) / 2;
            float2 pr[NP > 0 ? NP : 1];
#pragma unroll
            for (int q = 0; q < NP; ++q) pr[q] = make_float2(xs[PL + 2 * q], xs[PL + 2 * q + 1]);
#pragma unroll
            for (int st = 1; st < NP; st *= 2)
#pragma unroll
                for (int q = 0; q + st < NP; q += 2 * st) pr[q] = __fadd2_rn(pr[q], pr[q + st]);
            float b0 = NP > 0 ? pr[0].x + pr[0].y : 0.0f;
            if (LO < PL) b0 += xs[LO];
            if (PH <= HI) b0 += xs[HI];
            float B[RX];
            B[0] = b0;
#pragma unroll
            for (int j = 1; j < RX; ++j) B[j] = B[j - 1] + (xs[C0 + j + N] - xs[C0 + j - 1 - N]);
#pragma unroll
            for (int jp = 0; jp < RX / 2; ++jp) hB[jp] = make_float2(B[2 * jp], B[2 * jp + 1]);
          }
        };
        auto emit_to = [&](float* drow, const float2 (&v)[RX / 2]) {
            if (st_vec) {
                if constexpr (RX >= 4) {
#pragma unroll
                    for (int q = 0; q < RX / 4; ++q)
                        sg::st_cs_f4(drow + 4 * q, make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y));
                } else {
                    *reinterpret_cast<float2*>(drow) = v[0];
                }
            } else {
                RowVals<RX> rv;
#pragma unroll
                for (int q = 0; q < RX / 2; ++q) rv.v[q] = v[q];
                emit_ragged<RX>(drow, rv, X, Xlo, Xhi);
            }
        };
        auto emit = [&](const float2 (&v)[RX / 2]) {
            emit_to(dst_row, v);
            dst_row += a.out_stride;
        };
        // multi-output: the other images share geometry, pitch and 16-byte phase with a.out (the launcher checks),
        // so they are a constant element offset away
        const long long off1 = NO > 1 ? a.out1 - a.out : 0, off2 = NO > 2 ? a.out2 - a.out : 0;
        auto emit_multi = [&](const float2 (&v)[NO][RX / 2]) {
            emit_to(dst_row, v[0]);
            if constexpr (NO > 1) emit_to(dst_row + off1, v[1]);
            if constexpr (NO > 2) emit_to(dst_row + off2, v[2]);
            dst_row += a.out_stride;
        };

        if constexpr (ADD) {
            // Column pass of the additive form.  With v(0) = v(2n) = 0 (the host shifts the constant into u) rows
            // t, t+1 of a step contribute  P = A(t) + A(t+1)  to the 2n output rows whose window holds both,
            // plus v * B per row: 3 packed operations per output row and step instead of 4 (two factors, two rows),
            // and the first / last window rows cost one operation or none.
            int phase = 0;
#pragma unroll 1
            for (int t = 0; t < steps2; t += 2) {
                advance(t);

                float2 A0[RX / 2], B0[RX / 2], A1[RX / 2], B1[RX / 2];
                row_pass_add(t, A0, B0);
                row_pass_add(t + 1, A1, B1);

                float2 v0[RX / 2], v1[RX / 2];
                static_switch<NA / 2>(phase, [&](auto sc) {
                    constexpr int s2 = 2 * decltype(sc)::value;   // ring slot of output row t
                    constexpr int kOff = 4 * NA;
#pragma unroll
                    for (int jp = 0; jp < RX / 2; ++jp) {
                        const float2 P = __fadd2_rn(A0[jp], A1[jp]);
                        // output row t - 2n: row t is its last window row (v = 0), row t+1 is outside
                        {
                            constexpr int i = (s2 - 2 * N + kOff) % NA;
                            acc[0][jp][i] = __fadd2_rn(acc[0][jp][i], A0[jp]);
                            v0[jp] = acc[0][jp][i];
                        }
                        // output row t + 1 - 2n: window rows 2n-1 (row t) and 2n (row t+1, v = 0)
                        {
                            constexpr int i = (s2 + 1 - 2 * N + kOff) % NA;
                            const float cw = w.col[2 * N - 1];
                            acc[0][jp][i] = __ffma2_rn(make_float2(cw, cw), B0[jp], __fadd2_rn(acc[0][jp][i], P));
                            v1[jp] = acc[0][jp][i];
                        }
                        // output rows t - wy, wy = 1 .. 2n-2: window rows wy (row t) and wy + 1 (row t+1)
#pragma unroll
                        for (int wy = 1; wy <= 2 * N - 2; ++wy) {
                            const int i = (s2 - wy + kOff) % NA;
                            const float c0w = w.col[wy], c1w = w.col[wy + 1];
                            float2 r = __fadd2_rn(acc[0][jp][i], P);
                            r = __ffma2_rn(make_float2(c0w, c0w), B0[jp], r);
                            acc[0][jp][i] = __ffma2_rn(make_float2(c1w, c1w), B1[jp], r);
                        }
                        // output row t opens: window rows 0 (row t, v = 0) and 1 (row t+1)
                        {
                            const float cw = w.col[1];
                            acc[0][jp][s2 % NA] = __ffma2_rn(make_float2(cw, cw), B1[jp], P);
                        }
                        // output row t + 1 opens with its window row 0 (v = 0)
                        acc[0][jp][(s2 + 1) % NA] = A1[jp];
                    }
                });
                phase = phase + 1 == NA / 2 ? 0 : phase + 1;
                // (stores stay OUTSIDE the phases: one copy of the store code and two register moves per value measured
                // 4 % faster than a store inside each of the n+1 phases -- the loop body must stay small)
                const unsigned te = static_cast<unsigned>(t - 2 * N - shift);   // output row completed by row t
                if (te < static_cast<unsigned>(nrows)) emit(v0);
                if (te + 1u < static_cast<unsigned>(nrows)) emit(v1);
            }
        } else if constexpr (RING) {
            // The loop over steps stays rolled (one copy of the staging and row-pass code); only the column
            // pass exists once per phase of the ring, selected by a switch.  A fully unrolled period
            // (n+1 steps) measured 59 KB of loop body: more than the 32 KB instruction cache, the warps
            // of an SM are spread over the whole body and stall on instruction fetch.
            int phase = 0;
#pragma unroll 1
            for (int t = 0; t < steps2; t += 2) {
                advance(t);

                float2 h0[R][RX / 2], h1[R][RX / 2];
                row_pass(t, h0);
                row_pass(t + 1, h1);

                float2 v0[NO][RX / 2], v1[NO][RX / 2];
                static_switch<NA / 2>(phase, [&](auto sc) {
                    constexpr int s2 = 2 * decltype(sc)::value;   // ring position of row t
                    // column pass: row t is window row wy of output row t - wy -> slot (s2 - wy) mod NA, row
                    // t+1 of the slot after it.  wy = 0 opens a new output row (plain product: the slot
                    // still holds the row stored 2n+2 rows ago).
#pragma unroll
                    for (int wy = 0; wy <= 2 * N; ++wy)
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const float cw = w.col[r][wy];
                            constexpr int kOff = 4 * NA;   // keeps the modulo argument positive
                            const int i0 = (s2 - wy + kOff) % NA, i1 = (s2 + 1 - wy + kOff) % NA;
                            const int o = r / RPO;
#pragma unroll
                            for (int jp = 0; jp < RX / 2; ++jp) {
                                if (wy == 0 && r % RPO == 0) {
                                    acc[o][jp][i0] = __fmul2_rn(make_float2(cw, cw), h0[r][jp]);
                                    acc[o][jp][i1] = __fmul2_rn(make_float2(cw, cw), h1[r][jp]);
                                } else {
                                    acc[o][jp][i0] = __ffma2_rn(make_float2(cw, cw), h0[r][jp], acc[o][jp][i0]);
                                    acc[o][jp][i1] = __ffma2_rn(make_float2(cw, cw), h1[r][jp], acc[o][jp][i1]);
                                }
                            }
                        }
                    // output rows t - 2n and t + 1 - 2n are complete (their slots are dead until wy = 0 re-opens
                    // them, so these copies fold into the last FFMA2 of each row)
#pragma unroll
                    for (int o = 0; o < NO; ++o)
#pragma unroll
                        for (int jp = 0; jp < RX / 2; ++jp) {
                            v0[o][jp] = acc[o][jp][(s2 + 2) % NA];
                            v1[o][jp] = acc[o][jp][(s2 + 3) % NA];
                        }
                });
                phase = phase + 1 == NA / 2 ? 0 : phase + 1;
                if (t >= 2 * N && t - 2 * N < nrows) emit_multi(v0);
                if (t + 1 >= 2 * N && t + 1 - 2 * N < nrows) emit_multi(v1);
            }
        } else {
#pragma unroll 1
            for (int tb = 0; tb * kU < steps2; ++tb) {
#pragma unroll
                for (int u = 0; u < kU; u += 2) {
                    const int t = tb * kU + u;
                    if (t < steps2) {
                        advance(t);

                        float2 h0[R][RX / 2], h1[R][RX / 2];
                        row_pass(t, h0);
                        row_pass(t + 1, h1);

                        // ---- column pass: scatter both rows into the output rows in flight ----
                        // row t is window row wy of output i = u + 2n - wy; row t+1 is window row wy of output i+1
#pragma unroll
                        for (int wy = 0; wy <= 2 * N; ++wy)
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const float cw = w.col[r][wy];
                                const int i = u + 2 * N - wy;
#pragma unroll
                                for (int jp = 0; jp < RX / 2; ++jp) {
                                    acc[0][jp][i] = __ffma2_rn(make_float2(cw, cw), h0[r][jp], acc[0][jp][i]);
                                    acc[0][jp][i + 1] = __ffma2_rn(make_float2(cw, cw), h1[r][jp], acc[0][jp][i + 1]);
                                }
                            }

                        // ---- output rows u and u+1 of the block are complete ----
                        if (t >= 2 * N && t - 2 * N < nrows) {
                            float2 v[RX / 2];
#pragma unroll
                            for (int jp = 0; jp < RX / 2; ++jp) v[jp] = acc[0][jp][u];
                            emit(v);
                        }
                        if (t + 1 >= 2 * N && t + 1 - 2 * N < nrows) {
                            float2 v[RX / 2];
#pragma unroll
                            for (int jp = 0; jp < RX / 2; ++jp) v[jp] = acc[0][jp][u + 1];
                            emit(v);
                        }
                    }
                }
                // block done: drop the kU completed rows
#pragma unroll
                for (int jp = 0; jp < RX / 2; ++jp)
#pragma unroll
                    for (int i = 0; i < NA; ++i) acc[0][jp][i] = (i + kU < NA) ? acc[0][jp][i + kU] : make_float2(0.f, 0.f);
            }
        }
        cp_async_wait<0>();
    }
}

// Grid, band height, ticket counter and launch of one instantiation `kern` (weights `w` already filled).
// s_bps / s_sms: the instantiation's cached occupancy (function-local statics of the caller).
template <int N, int RX, class K, class W>
cudaError_t launch_common(K kern, const W& w, const Args2D& a, cudaStream_t stream, std::atomic<int>& s_bps, std::atomic<int>& s_sms)
{
    // resident CTAs per SM / SM count of this instantiation (same on every B200; filled once, any thread)
    int bps = s_bps.load(std::memory_order_acquire), sms = s_sms.load(std::memory_order_acquire);
    if (bps == 0 || sms == 0) {
        int dev = 0, nb = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kWarps * 32, 0);
        if (e != cudaSuccess) return e;
        bps = nb > 0 ? nb : 1;
        s_sms.store(sms, std::memory_order_release);
        s_bps.store(bps, std::memory_order_release);
    }
    constexpr int TW = 32 * RX;
    const long long strips = (a.cx + a.out_cols + TW - 1) / TW;
    // band height: as tall as possible (2n warm-up rows per band are recomputed), but short enough that
    // the launch has >= 4 work items per resident warp
    Args2D aa = a;
    const long long want = 4LL * sms * bps * kWarps;
    long long nb = (want + strips * a.n_images - 1) / (strips * a.n_images);   // bands per image wanted
    long long band = (a.out_rows + nb - 1) / (nb > 0 ? nb : 1);
    band = (band + kU - 1) / kU * kU;
    if (band < 8 * N + 8) band = 8 * N + 8;   // keep the warm-up overhead <= 25 %
    if (band > kBandMax) band = kBandMax;
    aa.band_rows = static_cast<int>(band);
    const long long bands = (a.out_rows + band - 1) / band;
    const long long items = strips * bands * a.n_images;
    if (items <= 0) return cudaSuccess;
    if (items >= (1LL << 31)) return cudaErrorInvalidValue;
    // bulk-tensor staging of interior work items: the images as a {cols, rows, images} tensor, box = the two rows a
    // step consumes.  Only where the kernel's row ring keeps the destinations 128-byte aligned and the image layout
    // meets the tensor-map rules (16-byte aligned base and strides); everything else stages with cp.async.
    Tma2D maps;
    std::memset(&maps, 0, sizeof(maps));
    aa.use_tma = 0;
    if (SG2D_TMA) {
        constexpr int PADX = (N + 3) & ~3;
        constexpr int ROWF = TW + 2 * PADX;
        static const bool off = [] { const char* e = std::getenv("SAVGOL_B200_NO_TMA2D"); return e && e[0] == '1'; }();
        if ((2 * ROWF * 4) % 128 == 0 && !off && ((reinterpret_cast<uintptr_t>(a.in) & 15) | (a.in_stride & 3) | (a.in_image_pitch & 3) | (a.cols & 3)) == 0 &&
            a.cols >= ROWF && a.rows >= 2) {
            if (sg::EncodeTiled enc = sg::encode_tiled()) {
                const cuuint64_t dims[3] = {static_cast<cuuint64_t>(a.cols), static_cast<cuuint64_t>(a.rows), static_cast<cuuint64_t>(a.n_images)};
                const cuuint64_t strides[2] = {static_cast<cuuint64_t>(a.in_stride) * 4,
                                               a.n_images > 1 ? static_cast<cuuint64_t>(a.in_image_pitch) * 4
                                                              : static_cast<cuuint64_t>(a.in_stride) * 4 * static_cast<cuuint64_t>(a.rows)};
                const cuuint32_t box[3] = {static_cast<cuuint32_t>(ROWF), 2, 1};
                const cuuint32_t es[3] = {1, 1, 1};
                if (strides[1] < (1ull << 40) && strides[0] < (1ull << 40) &&
                    enc(&maps.pair, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a.in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                    aa.use_tma = 1;
            }
        }
    }
    cudaError_t ec = acquire_counter(stream, &aa.counter);
    if (ec != cudaSuccess) return ec;
    long long grid = static_cast<long long>(sms) * bps;
    const long long need = (items + kWarps - 1) / kWarps;
    if (grid > need) grid = need;
    kern<<<static_cast<unsigned>(grid), kWarps * 32, 0, stream>>>(w, aa, maps);
    sg::g_launches.fetch_add(1);
    ec = cudaGetLastError();
    const cudaError_t ef = cudaFreeAsync(aa.counter, stream);
    return ec != cudaSuccess ? ec : ef;
}


template <int N, int R, bool ADD>
cudaError_t launch_nr(const Args2D& a, const SepPlan& plan, cudaStream_t stream)
{
    constexpr int RX = (N <= 8 && SG2D_RX4) ? SG2D_RXW : 2;
    typename WSel<R, ADD>::type w;
    const float sc = a.scale;
    if constexpr (ADD) {
        // W = u(x) + v(y): row[0] = u, col[0] = v with v(+-ny) = 0 (factor2d.cpp); both carry the scale
        float u[2 * N + 1];
        for (int k = 0; k <= 2 * N; ++k) u[k] = plan.row[0][k] * sc;
        for (int j = 1; j <= 2 * N; ++j) w.pu[j] = make_float2(u[j], u[j - 1]);
        w.pu[0] = make_float2(0.f, 0.f);
        w.u_first = u[0];
        w.u_last = u[2 * N];
        for (int k = 0; k <= 2 * N; ++k) w.col[k] = plan.col[0][k] * sc;
    } else {
        for (int r = 0; r < R; ++r) {
            // N = max(nx, ny): the shorter factor is centred and zero-padded (the extra taps multiply
            // boundary-mapped, i.e. finite, samples by 0)
            w.rc[r] = plan.row[r][plan.nx];
            for (int k = 1; k <= N; ++k) w.rk[r][k - 1] = k <= plan.nx ? plan.row[r][plan.nx + k] : 0.0f;
            for (int k = 0; k <= 2 * N; ++k) {
                const int j = k - N + plan.ny;
                w.col[r][k] = (j >= 0 && j <= 2 * plan.ny) ? plan.col[r][j] * sc : 0.0f;
            }
        }
        w.sx = plan.parity_x < 0 ? -1.0f : 1.0f;
    }
    static std::atomic<int> s_bps{0}, s_sms{0};
    return launch_common<N, RX>(sep_kernel<N, R, RX, ADD>, w, a, stream, s_bps, s_sms);
}

}  // namespace

}  // namespace sg2d
