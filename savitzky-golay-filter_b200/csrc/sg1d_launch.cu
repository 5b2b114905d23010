// sg1d_launch.cu -- grid sizing and dispatch for the 1D kernels.
#include "sg1d_launch.h"
#include "sg1d_packed.cuh"

#include <atomic>
#include <cstdlib>
#include <mutex>

namespace sg {

#define SG_DECL(g) const Kernel1D* sg1d_group_table_##g(); const Kernel1DTma* sg1d_tma_group_table_##g();
SG_DECL(0) SG_DECL(1) SG_DECL(2) SG_DECL(3) SG_DECL(4) SG_DECL(5) SG_DECL(6) SG_DECL(7)
#undef SG_DECL

std::atomic<unsigned long long> g_launches{0};
std::atomic<unsigned long long> g_tma_launches{0};   // launches of the bulk-tensor (TMA) 1D kernels
const bool g_tail_enabled = [] { const char* e = getenv("SAVGOL_B200_NO_TAIL"); return !(e && e[0] == '1'); }();
const bool g_phase_enabled = [] { const char* e = getenv("SAVGOL_B200_NO_PHASE"); return !(e && e[0] == '1'); }();
std::atomic<int> g_tma_enabled{[] { const char* e = getenv("SAVGOL_B200_NO_TMA"); return (e && e[0] == '1') ? 0 : 1; }()};

const Kernel1D* sg1d_group_table(int group)
{
    switch (group) {
        case 0: return sg1d_group_table_0();
        case 1: return sg1d_group_table_1();
        case 2: return sg1d_group_table_2();
        case 3: return sg1d_group_table_3();
        case 4: return sg1d_group_table_4();
        case 5: return sg1d_group_table_5();
        case 6: return sg1d_group_table_6();
        case 7: return sg1d_group_table_7();
        default: return nullptr;
    }
}

EncodeTiled encode_tiled()
{
    static std::atomic<EncodeTiled> s_fn{nullptr};
    static std::atomic<int> s_state{0};   // 0 unknown, 1 ok, -1 unavailable
    if (s_state.load(std::memory_order_acquire) == 0) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
        if (e == cudaSuccess && f && q == cudaDriverEntryPointSuccess) {
            s_fn.store(reinterpret_cast<EncodeTiled>(f), std::memory_order_release);
            s_state.store(1, std::memory_order_release);
        } else {
            (void)cudaGetLastError();
            s_state.store(-1, std::memory_order_release);
        }
    }
    return s_state.load(std::memory_order_acquire) == 1 ? s_fn.load(std::memory_order_acquire) : nullptr;
}


const Kernel1DTma* sg1d_tma_group_table(int group)
{
    switch (group) {
        case 0: return sg1d_tma_group_table_0();
        case 1: return sg1d_tma_group_table_1();
        case 2: return sg1d_tma_group_table_2();
        case 3: return sg1d_tma_group_table_3();
        case 4: return sg1d_tma_group_table_4();
        case 5: return sg1d_tma_group_table_5();
        case 6: return sg1d_tma_group_table_6();
        case 7: return sg1d_tma_group_table_7();
        default: return nullptr;
    }
}

namespace {
struct GridInfo { int blocks_per_sm = 0; };
GridInfo g_grid_tma[kMaxN + 1][VT_COUNT];

// {32 floats, nrows, rows} view of a batch of contiguous rows, 128-byte swizzle, box of `box_rows` tensor rows.
bool encode_rows(EncodeTiled enc, CUtensorMap* m, const void* base, unsigned long long nrows, unsigned long long rows,
                 unsigned long long row_bytes, unsigned box_rows)
{
    const cuuint64_t dims[3] = {32, nrows, rows};
    const cuuint64_t strides[2] = {128, rows > 1 ? row_bytes : nrows * 128};
    const cuuint32_t box[3] = {32, box_rows, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The TMA kernels take contiguous fp32 rows of at least one segment whose base and pitch are 16-byte aligned.
// Whether they are FASTER than the cp.async kernels was measured on B200 (tools/r2_sweep1d.py, profiles/r2_sweep1d.txt,
// 65,536 x 4,096): yes for half-windows up to 17 (+2 ... +10 %: 0.91 of the HBM roofline at n <= 4, 0.87 at n = 16; the
// staging / store instructions they save are issue slots the FFMA2 stream can use), no for the wide windows that are
// fp32-pipe bound anyway (18 ... 32: -1 ... -9 %, except 26 where the cp.async instantiation is the slow one), and no
// for launches of fewer than ~2048 segments, where the first tensor-map fetch is exposed latency (64 segments: 13 vs
// 8 us).  g_tma_enabled: 0 off, 1 this rule, 2 always (tests).
// Rows that end with 1..kTail outputs behind their last full segment (4097 = 4 x 1024 + 1): the generic kernel folds
// them into that segment (Args1D::tail); counted for the worst per-row phase (misaligned rows start up to kPhase-1
// outputs early).
// every row of the launch starts on a 16-byte boundary (the pitch of a single row does not matter)
bool rows_aligned16(const Args1D& a)
{
    return ((reinterpret_cast<uintptr_t>(a.in) | (a.rows > 1 ? static_cast<uintptr_t>(a.in_row_bytes) : 0)) & 15) == 0;
}

bool short_tail(const Args1D& a)
{
    if (!g_tail_enabled) return false;
    const bool aligned = rows_aligned16(a);
    const long long span = a.len + ((a.in_stride == 4 && !aligned && g_phase_enabled) ? kPhase - 1 : 0);
    if (span <= kTile) return false;   // one segment holds the row
    const long long over = span - (span - 1) / kTile * kTile;   // outputs in the last segment, 1..kTile
    return over <= kTail;
}

bool tma_eligible(int n, int variant, const Args1D& a)
{
    const int how = g_tma_enabled.load(std::memory_order_relaxed);
    if (how == 0 || (variant != V_BATCH_FAST && variant != V_STREAM_FAST)) return false;
    if (a.in_stride != 4 || a.out_stride != 4 || a.len < kTile) return false;
    if ((reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.out)) & 15) return false;
    if (a.rows > 1 && ((a.in_row_bytes | a.out_row_bytes) & 15)) return false;
    if (a.rows >= (1LL << 31) || a.len >= (1LL << 36) || a.in_row_bytes >= (1LL << 40) || a.out_row_bytes >= (1LL << 40)) return false;
    if (how == 1) {
        if (short_tail(a)) return false;   // 4 segments + tail beat 5 bulk-tensor segments
        if (!(n <= 17 || n == 26)) return false;
        if (((a.len + kTile - 1) / kTile) * a.rows < 2048) return false;
    }
    return true;
}
GridInfo g_grid[kMaxN + 1][V_COUNT][10];  // per (n, variant, packing x edge values); filled lazily (same value on every B200)
int g_sms[64];
std::mutex g_mu;
}  // namespace

constexpr size_t kPackedSmemMax = 106 * 1024;  // two CTAs per SM next to the kernel's 4 KB of static tables

// dynamic shared memory of the short-row kernel: 2 buffers per warp of 32/g row slots + edge values
static size_t packed_smem_bytes(int n, bool lead2n, int g, bool edges)
{
    const int lead = lead2n ? 2 * n : n;
    const int delta = ((lead + 3) & ~3) - lead;
    const int rpg = 32 / g, warps = kThreads / 32;
    // the polynomial edge values (2n per row slot) live behind the buffers; launches without polynomial edges do not
    // pay for them (64-sample rows, n = 16: 70 KB instead of 78 KB per CTA -> three resident CTAs instead of two)
    return static_cast<size_t>(warps) * 2 * rpg * packed_slot_chunks(n, delta, g) * 16 + (edges ? static_cast<size_t>(warps) * rpg * 2 * n * 4 : 0);
}

static cudaError_t sg1d_launch_tma(EncodeTiled enc, int n, int vt, const W1D& w, Args1D& a, cudaStream_t stream)
{
    TmaMaps maps;
    const unsigned long long nin = static_cast<unsigned long long>(a.len) >> 5, nout = static_cast<unsigned long long>(a.out_len) >> 5;
    const unsigned hl = static_cast<unsigned>(((vt == VT_STREAM ? 2 * n : n) + 31) / 32);   // GeoT<LEAD>::HL
    if (!encode_rows(enc, &maps.in_body, a.in, nin, a.rows, a.in_row_bytes, 32) ||
        !encode_rows(enc, &maps.in_first, a.in, nin, a.rows, a.in_row_bytes, 33) ||
        !encode_rows(enc, &maps.in_last, a.in, nin, a.rows, a.in_row_bytes, hl + 32) ||
        !encode_rows(enc, &maps.in_full, a.in, nin, a.rows, a.in_row_bytes, hl + 33))
        return cudaErrorNotSupported;
    a.out_tma = nout >= 32 ? 1 : 0;
    if (a.out_tma && !encode_rows(enc, &maps.out_body, a.out, nout, a.rows, a.out_row_bytes, 32)) return cudaErrorNotSupported;
    if (!a.out_tma) maps.out_body = maps.in_body;   // never dereferenced

    const Kernel1DTma& k = sg1d_tma_group_table((n - 1) / 4)[((n - 1) % 4) * VT_COUNT + vt];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int bps, sms;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        GridInfo& gi = g_grid_tma[n][vt];
        if (gi.blocks_per_sm == 0) {
            int nb = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k.kernel, kThreads, 0);
            if (e != cudaSuccess) return e;
            gi.blocks_per_sm = nb > 0 ? nb : 1;
        }
        if (dev < 64 && g_sms[dev] == 0) {
            e = cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev);
            if (e != cudaSuccess) return e;
        }
        bps = gi.blocks_per_sm;
        sms = dev < 64 ? g_sms[dev] : 148;
    }
    a.tiles_per_row = (a.len + kTile - 1) / kTile;
    a.ntiles = a.tiles_per_row * a.rows;
    if (a.ntiles <= 0) return cudaSuccess;
    if (a.ntiles >= (1LL << 31) - (1LL << 20)) return cudaErrorInvalidValue;
    long long grid = static_cast<long long>(sms) * bps;
    const long long need = (a.ntiles + kThreads / 32 - 1) / (kThreads / 32);
    if (grid > need) grid = need;
    k.kernel<<<static_cast<unsigned>(grid), kThreads, 0, stream>>>(w, a, maps);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    g_tma_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// Which kernel family / instantiation serves a launch, and its work decomposition -- pure host logic (no CUDA call), so
// the CPU test-suite can check it (savgol_b200_plan_1d).  Fills the dispatch fields of `a`.
Plan1D sg1d_plan(int n, int variant, Args1D& a, bool allow_tma)
{
    Plan1D p{};
    p.family = PLAN_GENERIC;
    // short rows (<= 512 samples): several rows per warp instead of idle lanes
    a.pack_g = 0;
    a.tail = a.phase = 0;
    a.out_tma = 0;
    // contiguous rows that are not all 16-byte aligned: row slots / segments start on a per-row phase (up to 3 outputs
    // early in the short-row kernel, up to kPhase-1 in the generic one) so that their chunks are aligned in memory
    const bool in_aligned = rows_aligned16(a);
    const bool phase = a.in_stride == 4 && !in_aligned && g_phase_enabled;
    const long long plen = a.len + (phase ? 3 : 0);   // outputs a short row's slot must hold
    if ((variant == V_BATCH_FAST || variant == V_STREAM_FAST) && plen <= 512 && a.rows > 1) {
        // lanes per row; wide windows on very short rows would need more shared memory per row slot than
        // two resident CTAs allow: fewer, wider slots then
        int g = plen <= 32 ? 1 : plen <= 64 ? 2 : plen <= 128 ? 4 : plen <= 256 ? 8 : 16;
        a.phase = phase ? 1 : 0;
        const bool edges = a.edge_lead || a.edge_trail;
        while (g < 16 && packed_smem_bytes(n, variant == V_STREAM_FAST, g, edges) > kPackedSmemMax) g *= 2;
        a.pack_g = g;
        p.family = PLAN_PACKED;
        p.gi_idx = (g == 1 ? 0 : g == 2 ? 1 : g == 4 ? 2 : g == 8 ? 3 : 4) + (edges ? 5 : 0);
        p.smem = packed_smem_bytes(n, variant == V_STREAM_FAST, g, edges);
        p.variant = variant == V_BATCH_FAST ? (phase ? V_PACK_BATCH_FAST_PH : V_PACK_BATCH_FAST) : (phase ? V_PACK_STREAM_FAST_PH : V_PACK_STREAM_FAST);
        a.tiles_per_row = 1;
        return p;
    }
    p.variant = variant;
    if (allow_tma && tma_eligible(n, variant, a)) {
        p.family = PLAN_TMA;
        a.tiles_per_row = (a.len + kTile - 1) / kTile;
        return p;
    }
    // generic kernel: misaligned rows / short tails take the PT instantiation of the FAST flavours (the exact flavours
    // stage misaligned rows with 4-byte copies, as before)
    if (variant == V_BATCH_FAST || variant == V_STREAM_FAST) {
        const long long span = a.len + (phase ? kPhase - 1 : 0);
        a.phase = phase ? 1 : 0;
        a.tiles_per_row = (span + kTile - 1) / kTile;
        // a last segment of <= kTail outputs would cost a whole pass of its warp: the segment before it takes them
        a.tail = short_tail(a) ? 1 : 0;
        if (a.tail) --a.tiles_per_row;
        if (a.phase || a.tail) p.variant = variant == V_BATCH_FAST ? V_BATCH_FAST_PT : V_STREAM_FAST_PT;
    } else {
        a.tiles_per_row = (a.len + kTile - 1) / kTile;
    }
    return p;
}

cudaError_t sg1d_launch(int n, int variant, const W1D& w, Args1D& a, cudaStream_t stream)
{
    if (n < 1 || n > kMaxN || variant < 0 || variant >= V_COUNT) return cudaErrorInvalidValue;
    if (variant >= V_PACK_BATCH_FAST) return cudaErrorInvalidValue;   // callers name the base flavours only
    Plan1D p = sg1d_plan(n, variant, a, true);
    if (p.family == PLAN_TMA) {
        if (EncodeTiled enc = encode_tiled()) {
            const cudaError_t e = sg1d_launch_tma(enc, n, variant == V_STREAM_FAST ? VT_STREAM : VT_BATCH, w, a, stream);
            if (e != cudaErrorNotSupported) return e;   // NotSupported: the driver refused a tensor map -> generic kernel
        }
        p = sg1d_plan(n, variant, a, false);
    }
    const int gi_idx = p.gi_idx;
    const size_t smem = p.smem;
    variant = p.variant;
    const Kernel1D& k = sg1d_group_table((n - 1) / 4)[((n - 1) % 4) * V_COUNT + variant];

    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int bps, sms;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        GridInfo& gi = g_grid[n][variant][gi_idx];
        if (smem + 6 * 1024 > 48 * 1024) {   // (the limit counts the kernel's ~5 KB of static tables too) per device and function; cheap enough to repeat
            e = cudaFuncSetAttribute(k.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kPackedSmemMax));
            if (e != cudaSuccess) return e;
        }
        if (gi.blocks_per_sm == 0) {
            int nb = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k.kernel, kThreads, smem);
            if (e != cudaSuccess) return e;
            gi.blocks_per_sm = nb > 0 ? nb : 1;
        }
        if (dev < 64 && g_sms[dev] == 0) {
            e = cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev);
            if (e != cudaSuccess) return e;
        }
        bps = gi.blocks_per_sm;
        sms = dev < 64 ? g_sms[dev] : 148;
    }

    if (a.pack_g) {
        // work unit = group of 32/g rows, one warp each
        const int rpg = 32 / a.pack_g;
        a.tiles_per_row = 1;
        a.ntiles = (a.rows + rpg - 1) / rpg;
    } else {
        // work unit = segment of kTile (1024) outputs of one row, one warp each (tiles_per_row: see above)
        a.ntiles = a.tiles_per_row * a.rows;
    }
    if (a.ntiles <= 0) return cudaSuccess;
    if (a.ntiles >= (1LL << 31) - (1LL << 20)) return cudaErrorInvalidValue;
    // persistent CTAs: one per resident slot (148 SMs x blocks/SM on a B200), round-robin over tiles
    long long grid = static_cast<long long>(sms) * bps;
    const long long need = (a.ntiles + kThreads / 32 - 1) / (kThreads / 32);
    if (grid > need) grid = need;
    k.kernel<<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(w, a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

}  // namespace sg

// Dispatch of a 1D launch as the library would decide it, without touching the GPU (include/savgol_b200.h).
extern "C" int savgol_b200_plan_1d(int half_window, int stream_variant, int exact, size_t rows, size_t length, size_t row_pitch,
                                   size_t first_sample_offset, int polynomial_edges, int* family, int* lanes_per_row, int* phase,
                                   int* tail, long long* segments_per_row)
{
    using namespace sg;
    if (half_window < 1 || half_window > kMaxN || rows == 0 || length == 0 || row_pitch < length) return -1;
    Args1D a{};
    a.in = reinterpret_cast<const char*>(static_cast<uintptr_t>(0x100000) + 4 * first_sample_offset);   // only its alignment matters
    a.out = reinterpret_cast<char*>(static_cast<uintptr_t>(0x40000000) + 4 * first_sample_offset);
    a.rows = static_cast<long long>(rows); a.len = a.out_len = static_cast<long long>(length);
    a.in_row_bytes = a.out_row_bytes = static_cast<long long>(row_pitch * 4);
    a.in_stride = a.out_stride = 4;
    a.edge_lead = a.edge_trail = polynomial_edges ? 1 : 0;
    const int variant = stream_variant ? (exact ? V_STREAM_EXACTSEQ : V_STREAM_FAST) : (exact ? V_BATCH_EXACT4 : V_BATCH_FAST);
    const Plan1D p = sg1d_plan(half_window, variant, a, true);
    if (family) *family = p.family;
    if (lanes_per_row) *lanes_per_row = a.pack_g;
    if (phase) *phase = a.phase;
    if (tail) *tail = a.tail;
    if (segments_per_row) *segments_per_row = a.tiles_per_row;
    return p.variant;
}

