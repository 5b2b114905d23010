// sg1d_tma.cuh -- the 1D stencil with bulk-tensor (TMA) staging and stores, sm_100a.
//
// Same arithmetic, same warp-autonomous pipeline and the same edge path as sg1d_kernel.cuh (which it
// replaces for the common case: contiguous fp32 rows, 16-byte aligned base and pitch, rows of >= 1024
// samples, FAST flavour; ref hot loop: src/savgolFilter.c:763-766, stream steady state
// src/savgol_stream.c:224-226).  What changes is how a segment travels:
//
//   * LOADS.  The batch is described to the TMA unit as a 3-D tensor {32 floats, len/32, rows} with
//     SWIZZLE_128B.  A segment (1024 outputs of one row) is ONE box: its 32 body rows (4 KB) plus the
//     128-byte halo rows that exist on either side (four tensor maps that differ only in their box height)
//     -- one `cp.async.bulk.tensor` issued by ONE lane, landing on the warp's private mbarrier -- instead of
//     nine LDGSTS per lane with their address arithmetic.
//     Halo rows that do not exist as tensor rows (first / last segment of a signal) are never touched by
//     the TMA unit; the edge path fills them exactly as before: one element per lane, a 4-byte cp.async
//     from the address the boundary rule designates (reflect / periodic / constant / explicit halo in a
//     neighbour GPU's memory / carried stream history), or a zero.
//   * SHARED LAYOUT.  128-byte rows, the 16-byte chunk index XOR-ed with the row index (bits 7..9 of the
//     shared address) -- the hardware swizzle.  A lane's window starts 128 bytes after its neighbour's,
//     so the eight lanes of a quarter-warp read eight different bank groups: conflict-free LDS.128 with
//     no padding chunks; the address of chunk c is `row_base ^ (c << 4)`, one LOP3.
//   * STORES.  Each lane parks its 32 outputs in row `lane` of the same swizzled layout (conflict-free
//     STS.128), then one lane issues a single bulk-tensor store of the 4 KB segment.
//
// Ragged last segments (len % 1024 != 0) are staged / stored by out-of-line per-lane code into the same
// layout.  Everything the TMA unit cannot describe (strided or misaligned rows, rows shorter than one
// segment, the exact flavours) stays on sg1d_kernel / sg1d_packed_kernel.
#pragma once
#include "sg1d_kernel.cuh"
#include "sg1d_launch.h"

namespace sg {

template <int LEAD>
struct GeoT {
    static constexpr int HL = (LEAD + 31) / 32;     // 128-byte halo rows left of the segment body
    static constexpr int PADT = 32 * HL;            // buffer position p <-> x index o0 - PADT + p
    static constexpr int DELTA = PADT - LEAD;       // thread t, output j, tap k reads position 32t + j + k + DELTA
    static constexpr int ROWS = HL + 33;            // halo rows + 32 body rows + one right halo row
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

// SWIZZLE_128B: bits 4..6 of the (absolute) shared address are XOR-ed with bits 7..9.
__device__ __forceinline__ unsigned sw_addr(unsigned buf, int p /* float index */)
{
    const unsigned a = buf + 4u * static_cast<unsigned>(p);
    return a ^ ((a >> 3) & 0x70u);
}

__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SG_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SG_DONE_%=;\n"
        "bra SG_WAIT_%=;\n"
        "SG_DONE_%=:\n"
        "}\n" ::"r"(mbar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned mbar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, unsigned src)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];\n" ::"l"(
                     reinterpret_cast<unsigned long long>(map)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ float4 lds128(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32(unsigned addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(unsigned addr, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(unsigned addr, float v)
{
    asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cp_async4_u(unsigned dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_u(unsigned dst, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gsrc) : "memory");
}

// FAST arithmetic on the swizzled layout (see compute_fast in sg1d_kernel.cuh for the FFMA2 scheme).
// lane_base = shared address of this lane's row (buffer row `lane`).
template <int N, int DELTA>
__device__ __forceinline__ void compute_fast_sw(unsigned lane_base, const W1D& W, float (&out)[kR])
{
    constexpr int WS = 2 * N + 1;
    constexpr int C_LO = DELTA / 4;
    constexpr int C_HI = (kR + 2 * N + DELTA + 3) / 4;   // chunks [C_LO, C_HI) of the lane's window hold taps
    constexpr int NROW = (C_HI + 7) / 8;                  // 128-byte rows the window touches (<= 4)
    // row bases with the swizzle key already in bits 4..6: chunk c of row r lives at kb[r] ^ (c << 4)
    unsigned kb[NROW];
#pragma unroll
    for (int r = 0; r < NROW; ++r) {
        const unsigned b = lane_base + 128u * r;
        kb[r] = b | ((b >> 3) & 0x70u);
    }
    float2 acc[kR / 2];
#pragma unroll
    for (int i = 0; i < kR / 2; ++i) acc[i] = make_float2(0.f, 0.f);

    static_for<C_HI - C_LO>([&](auto ci) {
        constexpr int c = C_LO + decltype(ci)::value;
        const float4 v = lds128(kb[c >> 3] ^ ((c & 7) << 4));
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

// Ragged segment (the row ends inside it): per-lane staging into the swizzled layout.  Chunks made of four
// existing samples travel as 16-byte copies, every other element the window can touch through the edge
// path.  Out of line and rolled: one segment per row at most.
template <int LEAD, int N>
__device__ __noinline__ void stage_generic_sw(unsigned buf, const Args1D& a, const char* xrow, long long row, long long o0, int lane)
{
    using G = GeoT<LEAD>;
    const long long left = a.len - o0;
    const int nout = left < kSeg ? static_cast<int>(left) : kSeg;
    const int p_lo = G::PADT - LEAD, p_hi = G::PADT + nout + N;   // positions the compute loop may need
    const int c_end = (p_hi + 3) >> 2;
#pragma unroll 1
    for (int c = (p_lo >> 2) + lane; c < c_end; c += 32) {
        const long long x0 = o0 - G::PADT + 4LL * c;
        const unsigned d = sw_addr(buf, 4 * c);
        if (x0 >= 0 && x0 + 4 <= a.len) {
            cp_async16_u(d, xrow + 4 * x0);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float* sp = sample_address<LEAD, N>(a, xrow, row, x0 + e);
                if (sp) cp_async4_u(d + 4 * e, sp);
                else sts32(d + 4 * e, 0.0f);
            }
        }
    }
}

// Lane-interleaved stores of a parked (swizzled) segment: output f = lane + 32 i is stored when f < lim.
static __device__ __noinline__ void store_generic_sw(unsigned buf, float* dst /* &out[o0] */, int lim, int lane)
{
#pragma unroll 4
    for (int i = 0; i < kR; ++i) {
        const int f = lane + 32 * i;
        if (f < lim) dst[f] = lds32(sw_addr(buf, f));
    }
}

template <int N, bool LEAD2N>
__global__ void __launch_bounds__(kThreads, SG_MIN_BLOCKS)
    sg1d_tma_kernel(const __grid_constant__ W1D W, const __grid_constant__ Args1D a, const __grid_constant__ TmaMaps maps)
{
    constexpr int LEAD = LEAD2N ? 2 * N : N;
    using G = GeoT<LEAD>;
    constexpr int HL = G::HL, PADT = G::PADT, DELTA = G::DELTA;
    constexpr int WS = 2 * N + 1;
    constexpr int kWarps = kThreads / 32;
    constexpr int kBufBytes = G::ROWS * 128;

    __shared__ __align__(128) unsigned char s_buf[kWarps][2][kBufBytes];
    __shared__ float s_edge[kWarps][2 * kMaxN];
    __shared__ __align__(8) unsigned long long s_mbar[kWarps][2];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned nseg = static_cast<unsigned>(a.ntiles);
    const unsigned spr = static_cast<unsigned>(a.tiles_per_row);
    const unsigned stride = gridDim.x * kWarps;
    const long long len = a.len;
    const unsigned nrows_in = static_cast<unsigned>(len >> 5);            // 128-byte tensor rows per signal (input map)
    const unsigned nrows_out = static_cast<unsigned>(a.out_len >> 5);     // ... output map

    unsigned buf_cur = smem_u32(s_buf[warp][0]), buf_nxt = smem_u32(s_buf[warp][1]);
    unsigned mb_cur = smem_u32(&s_mbar[warp][0]), mb_nxt = smem_u32(&s_mbar[warp][1]);
    if (lane == 0) {
        mbar_init(mb_cur, 1);
        mbar_init(mb_nxt, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        fence_proxy_async();
    }
    __syncwarp();
    const bool has_poly = (a.edge_lead | a.edge_trail) != 0;   // launch-uniform
    const bool has_state = a.state_out != nullptr;
    unsigned par_cur = 0, par_nxt = 0;    // phase parity each buffer's mbarrier completes next
    bool tma_cur = false, tma_nxt = false;

    // Stage segment (row, t) into `buf`; returns true when its body travels by TMA (the consumer then waits
    // on `mbar`).  Warp-collective.
    auto stage = [&](unsigned buf, unsigned mbar, const char* xrow, long long row, unsigned t, long long o0) -> bool {
        const bool body = (t + 1) * 32 <= nrows_in;
        if (!body) {
            stage_generic_sw<LEAD, N>(buf, a, xrow, row, o0, lane);
            return false;
        }
        const bool lh = t > 0;                       // the halo rows exist as tensor rows of this signal
        const bool rh = (t + 1) * 32 < nrows_in;
        if (lane == 0) {
            // ONE box per segment: body + whichever halo rows exist (four tensor maps that differ in their box height)
            const CUtensorMap* m = lh ? (rh ? &maps.in_full : &maps.in_last) : (rh ? &maps.in_first : &maps.in_body);
            mbar_expect_tx(mbar, 4096u + (lh ? 128u * HL : 0u) + (rh ? 128u : 0u));
            tma_load_3d(buf + (lh ? 0 : 128 * HL), m, 0, static_cast<int>(32 * t) - (lh ? HL : 0), static_cast<int>(row), mbar);
        }
        // edge path: the halo elements the TMA unit did not bring (true ends of the signal)
        if (!lh) {
#pragma unroll 1
            for (int q = lane; q < LEAD; q += 32) {
                const float* sp = sample_address<LEAD, N>(a, xrow, row, o0 - LEAD + q);
                const unsigned d = sw_addr(buf, PADT - LEAD + q);
                if (sp) cp_async4_u(d, sp);
                else sts32(d, 0.0f);
            }
        }
        if (!rh) {
#pragma unroll 1
            for (int q = lane; q < N; q += 32) {
                const float* sp = sample_address<LEAD, N>(a, xrow, row, o0 + kSeg + q);
                const unsigned d = sw_addr(buf, PADT + kSeg + q);
                if (sp) cp_async4_u(d, sp);
                else sts32(d, 0.0f);
            }
        }
        return true;
    };

    unsigned seg = blockIdx.x * kWarps + warp;
    unsigned row_u = seg / spr, t = seg - row_u * spr;
    const unsigned step_r = stride / spr, step_t = stride - step_r * spr;
    long long row = row_u;
    long long o0 = static_cast<long long>(t) * kSeg;
    const char* xrow = a.in + row * a.in_row_bytes;
    if (seg < nseg) tma_cur = stage(buf_cur, mb_cur, xrow, row, t, o0);
    cp_async_commit();

    for (; seg < nseg; seg += stride) {
        unsigned nt = t + step_t;
        long long nrow = row + step_r;
        if (nt >= spr) { nt -= spr; ++nrow; }
        const long long no0 = static_cast<long long>(nt) * kSeg;
        const char* nxrow = a.in + nrow * a.in_row_bytes;
        // the previous iteration's bulk store has finished reading buf_nxt before anything refills it
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        tma_nxt = false;
        if (seg + stride < nseg) tma_nxt = stage(buf_nxt, mb_nxt, nxrow, nrow, nt, no0);
        cp_async_commit();

        // polynomial edge outputs of this segment, one lane per output.  ref: src/savgolFilter.c:769-784
        // (one launch-uniform flag keeps the polynomial-edge / stream-state bookkeeping off the common path)
        const bool lead_seg = has_poly && a.edge_lead && o0 < N;
        const bool trail_seg = has_poly && a.edge_trail && (o0 + kSeg > len - N);
        if (lead_seg && lane < N) {
            const float s = dot_ordered<WS, ARITH_FAST>([&](int k) { return a.edge_t[k * 32 + lane]; },
                                                        [&](int k) { return ld_sample(xrow, 4, 2 * N - k); });
            s_edge[warp][lane] = s * a.scale;
        }
        if (trail_seg && lane < N) {
            const long long base = len - WS;
            const float s = dot_ordered<WS, ARITH_FAST>([&](int k) { return a.edge_t[k * 32 + lane]; },
                                                        [&](int k) { return ld_sample(xrow, 4, base + k); });
            s_edge[warp][kMaxN + lane] = s * a.scale;
        }

        cp_async_wait<1>();                       // this lane's element copies of the current segment ...
        if (tma_cur) { mbar_wait(mb_cur, par_cur); par_cur ^= 1u; }   // ... the TMA boxes ...
        __syncwarp();                             // ... and the other lanes' copies (and s_edge) have landed

        float out[kR];
        compute_fast_sw<N, DELTA>(buf_cur + 128u * lane, W, out);

        const long long o = o0 + kR * lane;
        if (lead_seg || trail_seg) {
#pragma unroll
            for (int j = 0; j < kR; ++j) {
                const long long oj = o + j;
                if (lead_seg && oj < N) out[j] = s_edge[warp][oj];
                else if (trail_seg && oj >= len - N && oj < len) out[j] = s_edge[warp][kMaxN + (len - 1 - oj)];
            }
        }

        // stream: the last state_w samples of [lead pad | x] become the next chunk's history
        if (has_state && o0 + kSeg >= len) {
            const long long first = len - a.state_w;
            const bool staged = first >= o0 - LEAD;
            for (int i = lane; i < a.state_w; i += 32) {
                const float v = staged ? lds32(sw_addr(buf_cur, static_cast<int>(first - o0) + PADT + i))
                                       : virtual_sample<LEAD, N>(a, xrow, row, first + i);
                a.state_out[row * a.state_pitch + i] = v;
            }
        }

        __syncwarp();  // every lane has finished reading its window
        {
            // park: 32 outputs = the 8 chunks of buffer row `lane`
            const unsigned b = buf_cur + 128u * lane;
            const unsigned kbase = b | ((b >> 3) & 0x70u);
#pragma unroll
            for (int q = 0; q < kR / 4; ++q)
                sts128(kbase ^ (q << 4), make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]));
        }
        if (a.out_tma && (t + 1) * 32 <= nrows_out) {
            fence_proxy_async();   // generic-proxy writes above -> visible to the bulk store
            __syncwarp();
            if (lane == 0) {
                tma_store_3d(&maps.out_body, 0, static_cast<int>(32 * t), static_cast<int>(row), buf_cur);
                bulk_commit();
            }
        } else {
            __syncwarp();
            const long long remain = a.out_len - o0;
            const int lim = remain < kSeg ? static_cast<int>(remain) : kSeg;
            store_generic_sw(buf_cur, reinterpret_cast<float*>(a.out + row * a.out_row_bytes) + o0, lim, lane);
            __syncwarp();
        }
        row = nrow; o0 = no0; xrow = nxrow; t = nt;
        { const unsigned x = buf_cur; buf_cur = buf_nxt; buf_nxt = x; }
        { const unsigned x = mb_cur; mb_cur = mb_nxt; mb_nxt = x; }
        { const unsigned x = par_cur; par_cur = par_nxt; par_nxt = x; }
        tma_cur = tma_nxt;
    }
    cp_async_wait<0>();
    if (lane == 0) bulk_wait_read0();   // shared memory stays valid until the last bulk store has read it
}

}  // namespace sg
