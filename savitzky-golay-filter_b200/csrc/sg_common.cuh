// sg_common.cuh -- shared device helpers and the host-visible launch descriptors of the
// sm_100a Savitzky-Golay kernels.  Nothing here is a port of reference code: the reference is
// scalar C (src/savgolFilter.c), this is the B200 execution plan for the same arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sg {

constexpr int kThreads = 128;              // 4 warps per CTA
constexpr int kR = 32;                     // consecutive outputs per thread (register sliding window)
constexpr int kTile = 32 * kR;             // 1024 outputs per segment (one warp)
constexpr int kSeg = kTile;
constexpr int kPhase = 32;                  // generic kernel: misaligned rows start their segments on 128-byte lines, up to kPhase-1 outputs early
constexpr int kTail = 32;                   // generic kernel: a row's last segment may carry up to kTail extra outputs, one per lane
constexpr int kMaxN = 32;
constexpr int kMaxWs = 2 * kMaxN + 1;      // 65, ref: include/iterative/savgolFilter.h:42

// Boundary synthesis of the virtual pad samples (ref: src/savgolFilter.c:442-482).
enum : int { MODE_POLY = 0, MODE_REFLECT = 1, MODE_PERIODIC = 2, MODE_CONSTANT = 3 };
// Arithmetic flavour.  FAST: packed-FMA chains (within 1e-6*max|x|/dt^d of the reference).
// EXACT4: the reference's 4-chain order with unfused multiply/add (src/savgolFilter.c:547-580)
//         -> bit-identical to savgol_apply.
// EXACTSEQ: single sequential accumulator, unfused (src/savgol_stream.c:25-38) -> bit-identical
//         to the stream API.
enum : int { ARITH_FAST = 0, ARITH_EXACT4 = 1, ARITH_EXACTSEQ = 2 };

// Centre weights travel as a kernel parameter: they land in constant bank 0 and ptxas feeds them
// to the FMA pipe as uniform-register operands.
//   w[k]    : the reference's centre weights (exact flavours; polynomial edges use their own table)
//   pw[k]   : FAST flavour, pair (ws[k], ws[k-1]) with ws = w * (1/dt^d) pre-scaled on the host, k = 1..2n:
//             one sample x[i] updates two neighbouring outputs (out[j] by tap k, out[j+1] by tap k-1)
//             with a single packed FFMA2, the sample broadcast to both halves.
//   ws_first / ws_last : scaled w[0] and w[2n] for the two half-pairs at the window ends (scalar FFMA,
//             so no sample outside an output's true window is ever multiplied, not even by zero).
struct W1D {
    float w[kMaxWs];
    float ws_first, ws_last, pad_;
    float2 pw[kMaxWs + 1];
};

// One launch = `rows` independent signals of `len` samples, cut into tiles of kTile outputs.
// Virtual signal V per row:  V = [ lead pad | x[0..len) | n pad ],  out[o] = scale * sum_k w[k] V[o+k],
// lead = n (batch: output o is centred on x[o]) or 2n (stream: output o is centred on x[o-n],
// the pad is the carried history).
struct Args1D {
    const char* in;          // sample i of row r at in + r*in_row_bytes + i*in_stride
    char* out;               // same addressing for outputs
    long long rows, len;
    long long out_len;       // outputs [0,out_len) of every row are stored (out_len <= len)
    long long in_row_bytes, out_row_bytes;
    long long in_stride, out_stride;   // bytes between consecutive samples (4 = contiguous)
    const float* lhalo;      // optional explicit left pad: `lead` samples per row, chronological
    const float* rhalo;      // optional explicit right pad: n samples per row
    long long lhalo_pitch, rhalo_pitch;  // elements between rows
    const float* edge_t;     // polynomial edge table, transposed: edge_t[k*32 + e] = E[e][k]
    float* state_out;        // stream: receives the last state_w samples of [lead pad | x] per row
    long long state_pitch;
    int state_w;
    float scale;             // 1/dt^d, applied as a separate multiply like the reference
    int mode;                // MODE_* used where a halo pointer is null
    int edge_lead, edge_trail;  // 1: outputs [0,n) / [len-n,len) come from the polynomial edge table
    long long tiles_per_row, ntiles;  // segments (1024 outputs) per row / in the launch; ntiles < 2^31
    int pack_g;              // short-row kernel (sg1d_packed.cuh): lanes per row (16, 8, 4); ntiles = row groups
    int tail;                // generic kernel: 1 = the last segment of every row also produces the (<= kTail) outputs behind it
    int phase;               // generic kernel: contiguous rows that are not 16-byte aligned are cut on a per-row phase (sg1d_kernel.cuh)
    int out_tma;             // TMA kernel (sg1d_tma.cuh): full segments are stored by one bulk-tensor copy
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc)
{
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(kPending) : "memory");
}

// Streaming store: outputs are written once and never re-read by this kernel.
__device__ __forceinline__ void st_cs_f4(float* p, float4 v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

}  // namespace sg
