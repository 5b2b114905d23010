// sg1d_packed.cuh -- 1D stencil for batches of SHORT signals (len <= 512): several rows per warp.
//
// sg1d_kernel gives every warp one segment of 1024 outputs of ONE row; a 360-sample signal (the
// reference's own demo dataset, test/iterative/test_savgol_main.c:55-92) would use 12 of its 32
// lanes while all 32 execute the full FFMA2 stream.  Here a warp owns a GROUP of 32/g rows, g = 16,
// 8 or 4 lanes per row (rows of <= 512, 256, 128 samples).  Everything else is the same plan: two
// private shared-memory buffers per warp, cp.async staging with the boundary rule applied while
// staging, the same register sliding window and packed FFMA2 arithmetic (compute_fast), stores
// through the warp's buffer in lane-interleaved order.  Each row slot has its own halo; slot sizes
// are padded so that a lane's window loads stay bank-conflict free
// (lane base = slot * SLOT + 9 * p chunks, SLOT == g (mod 8) when g < 8).
#pragma once
#include "sg1d_kernel.cuh"

namespace sg {

// chunks (16 B) of one row slot: logical, physical (one pad chunk per 8), padded for bank spreading
__host__ __device__ constexpr int packed_slot_chunks(int n, int delta, int g)
{
    const int nch = (32 * g + 2 * n + delta + 3) / 4;
    int phys = nch + (nch >> 3) + 1;
    if (g < 8) while ((phys & 7) != g) ++phys;
    return phys;
}

// Where the sample that stands for x-index xi of ANY row comes from (sample_address of sg1d_kernel.cuh without the
// row): kind 0 = sample idx of the row itself, 1 = entry idx of the row's left halo, 2 = of its right halo,
// 3 = zero.  Rows of one launch share length, mode and halo layout, so the pad elements of a row slot are
// described once per CTA (table in shared memory) instead of being re-derived for every row.
constexpr int kPadKindShift = 28;
template <int LEAD, int N>
__device__ __forceinline__ int pad_source(const Args1D& a, int xi)
{
    const int len = static_cast<int>(a.len);
    int idx = xi;
    if (xi < 0) {
        if (a.lhalo) {
            const int h = LEAD + xi;
            return h >= 0 ? ((1 << kPadKindShift) | h) : (3 << kPadKindShift);
        }
        switch (a.mode) {
            case MODE_REFLECT: idx = -xi - 1; if (idx >= len) idx = len - 1; break;
            case MODE_PERIODIC: idx = ((xi % len) + len) % len; break;
            case MODE_CONSTANT: idx = 0; break;
            default: return 3 << kPadKindShift;
        }
    } else if (xi >= len) {
        if (a.rhalo) {
            const int h = xi - len;
            return h < N ? ((2 << kPadKindShift) | h) : (3 << kPadKindShift);
        }
        switch (a.mode) {
            case MODE_REFLECT: idx = 2 * len - xi - 1; if (idx < 0) idx = 0; break;
            case MODE_PERIODIC: idx = xi % len; break;
            case MODE_CONSTANT: idx = len - 1; break;
            default: return 3 << kPadKindShift;
        }
    }
    return idx;
}
constexpr int kPadTabMax = 112;   // pad elements of a row slot: <= PAD + 4 (<= 68) on the left, <= n + 9 on the right

// PHASE: rows that are not 16-byte aligned start their slot on a per-row phase (the launcher picks the instantiation).
template <int N, bool LEAD2N, bool PHASE>
__global__ void __launch_bounds__(kThreads, N <= 18 ? 4 : 3) sg1d_packed_kernel(const __grid_constant__ W1D W, const __grid_constant__ Args1D a)
{
    constexpr int LEAD = LEAD2N ? 2 * N : N;
    constexpr int PAD = Geo<LEAD>::PAD;
    constexpr int DELTA = Geo<LEAD>::DELTA;
    constexpr int WS = 2 * N + 1;
    constexpr int kWarps = kThreads / 32;

    extern __shared__ __align__(16) float4 s_dyn[];
    const int g = a.pack_g;                       // lanes per row: 16, 8, 4, 2 or 1
    const int rpg = 32 / g;                       // rows per warp group
    const int SLOT = packed_slot_chunks(N, DELTA, g);
    const int buf_chunks = rpg * SLOT;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = lane / g, p = lane - slot * g;
    float4* buf_cur = s_dyn + (warp * 2 + 0) * buf_chunks;
    float4* buf_nxt = s_dyn + (warp * 2 + 1) * buf_chunks;
    float* s_edge = reinterpret_cast<float*>(s_dyn + kWarps * 2 * buf_chunks) + warp * (rpg * 2 * N);  // [slot][lead n | trail n]

    const long long len = a.len;
    const int ilen = static_cast<int>(len);
    const unsigned ngroups = static_cast<unsigned>(a.ntiles);
    const unsigned stride = gridDim.x * kWarps;
    const bool out_aligned = a.out_stride == 4 && ((reinterpret_cast<uintptr_t>(a.out) | static_cast<uintptr_t>(a.out_row_bytes)) & 15) == 0;

    // Geometry of a row slot.  Slot position 0 <-> x index -sh - PAD, where sh is the row's PHASE: 0 for rows whose
    // first sample is 16-byte aligned (and for every row of a launch without PHASE), else (address of x[0] / 4)
    // mod 4 -- the slot of a misaligned contiguous row starts up to 3 outputs early so that its 16-byte chunks are
    // aligned in global memory (same idea as the per-row phase of sg1d_kernel.cuh; outputs before 0 are never
    // stored).  Per phase: chunks [c_lo, c_hi) consist of four existing samples and are copied whole; the other
    // nrest elements -- pads, ragged ends -- are described ONCE per CTA by a table:
    // element el of the slot (float position el + 4 * (el >> 5)) <- pad_source(el - PAD - sh).
    __shared__ int2 s_pad[PHASE ? 4 : 1][kPadTabMax];   // (static shared memory counts against the resident CTAs: keep it small)
    __shared__ int s_geo[PHASE ? 4 : 1][4];   // c_lo, c_hi, nrest, (unused)
    const int nphase = PHASE ? 4 : 1;
    for (int sh = 0; sh < nphase; ++sh) {
        const int nch = (ilen + sh + 2 * N + DELTA + 3) >> 2;      // chunks of a row the compute loop may touch
        const int c_lo = (PAD + sh + 3) >> 2;                      // first chunk made of four existing samples
        const int c_all = (ilen + sh + PAD) >> 2;
        const int c_hi = c_all < nch ? c_all : nch;
        const int nl = 4 * c_lo, nrest = nl + 4 * (nch - c_hi);
        if (threadIdx.x == 0) { s_geo[sh][0] = c_lo; s_geo[sh][1] = c_hi; s_geo[sh][2] = nrest; s_geo[sh][3] = 0; }
        for (int q = threadIdx.x; q < nrest; q += kThreads) {
            const int el = q < nl ? q : 4 * c_hi + (q - nl);
            s_pad[sh][q] = make_int2(pad_source<LEAD, N>(a, el - PAD - sh), el + 4 * (el >> 5));
        }
    }
    __syncthreads();   // the only CTA-wide barrier: once, before the warps go their own ways
    // launches without phases keep their (single) geometry in registers
    const int c_lo0 = (PAD + 3) >> 2;
    const int nch0 = (ilen + 2 * N + DELTA + 3) >> 2, c_all0 = (ilen + PAD) >> 2;
    const int c_hi0 = c_all0 < nch0 ? c_all0 : nch0;
    const int nrest0 = 4 * c_lo0 + 4 * (nch0 - c_hi0);
    auto row_phase = [&](const char* xrow) -> int { return PHASE ? static_cast<int>((reinterpret_cast<uintptr_t>(xrow) >> 2) & 3) : 0; };

    // stage this lane's share of row `row` into slot `slot` of `buf`
    auto stage = [&](float4* buf, long long row) {
        if (row >= a.rows) return;
        const char* xrow = a.in + row * a.in_row_bytes;
        float4* sbuf = buf + slot * SLOT;
        const int sh = row_phase(xrow);
        const int c_lo = PHASE ? s_geo[sh][0] : c_lo0, c_hi = PHASE ? s_geo[sh][1] : c_hi0, nrest = PHASE ? s_geo[sh][2] : nrest0;
        const char* src0 = xrow - static_cast<long long>(PAD + sh) * a.in_stride;
        const bool vec_ok = (a.in_stride == 4) && ((reinterpret_cast<uintptr_t>(src0) & 15) == 0);
        if (vec_ok) {
            for (int c = c_lo + p; c < c_hi; c += g) cp_async16(sbuf + c + (c >> 3), src0 + 16LL * c);
        } else {
            for (int e = 4 * c_lo + p; e < 4 * c_hi; e += g) {
                const int c = e >> 2;
                cp_async4(reinterpret_cast<float*>(sbuf + c + (c >> 3)) + (e & 3), src0 + static_cast<long long>(e) * a.in_stride);
            }
        }
        // pad elements: source and destination come from the CTA's table
        float* sf = reinterpret_cast<float*>(sbuf);
        for (int q = p; q < nrest; q += g) {
            const int2 e = s_pad[sh][q];
            const int kind = e.x >> kPadKindShift, idx = e.x & ((1 << kPadKindShift) - 1);
            float* d = sf + e.y;
            if (kind == 0) cp_async4(d, xrow + static_cast<long long>(idx) * a.in_stride);
            else if (kind == 1) cp_async4(d, a.lhalo + row * a.lhalo_pitch + idx);
            else if (kind == 2) cp_async4(d, a.rhalo + row * a.rhalo_pitch + idx);
            else *d = 0.0f;
        }
    };

    unsigned grp = blockIdx.x * kWarps + warp;
    if (grp < ngroups) stage(buf_cur, static_cast<long long>(grp) * rpg + slot);
    cp_async_commit();

    for (; grp < ngroups; grp += stride) {
        const long long row = static_cast<long long>(grp) * rpg + slot;
        const bool active = row < a.rows;
        const char* xrow = a.in + row * a.in_row_bytes;
        if (grp + stride < ngroups) stage(buf_nxt, static_cast<long long>(grp + stride) * rpg + slot);
        cp_async_commit();

        // polynomial edges of this slot's row: its g lanes share the 2n edge outputs
        float* se = s_edge + slot * (2 * N);
        if (active && (a.edge_lead || a.edge_trail)) {
            for (int e = p; e < N; e += g) {
                if (a.edge_lead) {
                    const float s = dot_ordered<WS, ARITH_FAST>([&](int k) { return a.edge_t[k * 32 + e]; },
                                                                [&](int k) { return ld_sample(xrow, a.in_stride, 2 * N - k); });
                    se[e] = s * a.scale;
                }
                if (a.edge_trail) {
                    const long long base = len - WS;
                    const float s = dot_ordered<WS, ARITH_FAST>([&](int k) { return a.edge_t[k * 32 + e]; },
                                                                [&](int k) { return ld_sample(xrow, a.in_stride, base + k); });
                    se[N + e] = s * a.scale;
                }
            }
        }

        cp_async_wait<1>();
        __syncwarp();

        const float4* sb = buf_cur + slot * SLOT + 9 * p;
        float out[kR];
        compute_fast<N, DELTA>(sb, W, out);

        const int sh = active ? row_phase(xrow) : 0;
        const int o = kR * p - sh;  // first output of this lane inside its row
        if (a.edge_lead || a.edge_trail) {
#pragma unroll
            for (int j = 0; j < kR; ++j) {
                const int oj = o + j;
                if (a.edge_lead && oj < N) { if (oj >= 0) out[j] = se[oj]; }
                else if (a.edge_trail && oj >= ilen - N && oj < ilen) out[j] = se[N + (ilen - 1 - oj)];
            }
        }

        // stream: hand the last state_w samples of [lead pad | x] to the next chunk (always staged here)
        if (a.state_out != nullptr && active) {
            const float* bf = reinterpret_cast<const float*>(buf_cur + slot * SLOT);
            const int first = ilen - a.state_w + PAD + sh;    // buffer position of the oldest carried sample
            for (int i = p; i < a.state_w; i += g) {
                float v;
                if (first >= 0) { const int pos = first + i; v = bf[4 * ((pos >> 2) + (pos >> 5)) + (pos & 3)]; }
                else v = virtual_sample<LEAD, N>(a, xrow, row, len - a.state_w + i);
                a.state_out[row * a.state_pitch + i] = v;
            }
        }

        // park the outputs in the slot (chunks 9p .. 9p+7), then write every row out lane-interleaved
        __syncwarp();
        {
            float4* park = buf_cur + slot * SLOT + 9 * p;
#pragma unroll
            for (int q = 0; q < kR / 4; ++q) park[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
        }
        __syncwarp();
        {
            const int lim = static_cast<int>(a.out_len < len ? a.out_len : len);
            const long long row0 = static_cast<long long>(grp) * rpg;
            const int shift = 36 - __clz(g);                  // log2(outputs parked per slot) = 5 + log2(g)
            if (out_aligned && !PHASE) {
                // chunk q = lane + 32 i of the group's parked outputs: 512 contiguous bytes per store
#pragma unroll
                for (int i = 0; i < kR / 4; ++i) {
                    const int q = lane + 32 * i;
                    const int s_ = q >> (shift - 2), c = q & (8 * g - 1);
                    const long long r_ = row0 + s_;
                    if (r_ < a.rows && 4 * c < lim) {
                        const float4 v = buf_cur[s_ * SLOT + c + (c >> 3)];
                        float* dst = reinterpret_cast<float*>(a.out + r_ * a.out_row_bytes) + 4 * c;
                        if (4 * c + 4 <= lim) st_cs_f4(dst, v);
                        else {
                            dst[0] = v.x;
                            if (4 * c + 1 < lim) dst[1] = v.y;
                            if (4 * c + 2 < lim) dst[2] = v.z;
                        }
                    }
                }
            } else if (a.out_stride == 4) {
                // chunk q = lane + 32 i of the group's parked outputs (contiguous rows: 512 contiguous bytes per store).
                // Chunk c of a row holds outputs 4c - sh .. 4c - sh + 3; it is one 16-byte store when it lies inside
                // the row and its address is aligned (always, when the output row has the phase of the input row)
#pragma unroll
                for (int i = 0; i < kR / 4; ++i) {
                    const int q = lane + 32 * i;
                    const int s_ = q >> (shift - 2), c = q & (8 * g - 1);
                    const long long r_ = row0 + s_;
                    if (r_ < a.rows) {
                        const int o4 = 4 * c - row_phase(a.in + r_ * a.in_row_bytes);   // first output of the chunk
                        if (o4 < lim && o4 + 4 > 0) {
                            const float4 v = buf_cur[s_ * SLOT + c + (c >> 3)];
                            float* dst = reinterpret_cast<float*>(a.out + r_ * a.out_row_bytes) + o4;
                            if (o4 >= 0 && o4 + 4 <= lim && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) st_cs_f4(dst, v);
                            else {
                                if (o4 >= 0) dst[0] = v.x;
                                if (o4 + 1 >= 0 && o4 + 1 < lim) dst[1] = v.y;
                                if (o4 + 2 >= 0 && o4 + 2 < lim) dst[2] = v.z;
                                if (o4 + 3 >= 0 && o4 + 3 < lim) dst[3] = v.w;
                            }
                        }
                    }
                }
            } else {
#pragma unroll 4
                for (int i = 0; i < kR; ++i) {
                    const int q = lane + 32 * i;             // flat index over the group's parked outputs
                    const int s_ = q >> shift, f = q & ((1 << shift) - 1);
                    const long long r_ = row0 + s_;
                    if (r_ < a.rows) {
                        const int of = f - row_phase(a.in + r_ * a.in_row_bytes);   // parked position f holds output f - sh
                        if (of >= 0 && of < lim) {
                            const float v = reinterpret_cast<const float*>(buf_cur + s_ * SLOT)[f + 4 * (f >> 5)];
                            *reinterpret_cast<float*>(a.out + r_ * a.out_row_bytes + static_cast<long long>(of) * a.out_stride) = v;
                        }
                    }
                }
            }
        }
        __syncwarp();
        float4* const tmp = buf_cur; buf_cur = buf_nxt; buf_nxt = tmp;
    }
    cp_async_wait<0>();
}

}  // namespace sg
