// sg1d_launch.cu -- grid sizing and dispatch for the 1D kernels.
#include "sg1d_launch.h"
#include "sg1d_packed.cuh"

#include <atomic>
#include <mutex>

namespace sg {

#define SG_DECL(g) const Kernel1D* sg1d_group_table_##g();
SG_DECL(0) SG_DECL(1) SG_DECL(2) SG_DECL(3) SG_DECL(4) SG_DECL(5) SG_DECL(6) SG_DECL(7)
#undef SG_DECL

std::atomic<unsigned long long> g_launches{0};

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

namespace {
struct GridInfo { int blocks_per_sm = 0; };
GridInfo g_grid[kMaxN + 1][V_COUNT][5];  // per (n, variant, packing); filled lazily (same value on every B200)
int g_sms[64];
std::mutex g_mu;
}  // namespace

constexpr size_t kPackedSmemMax = 110 * 1024;  // two CTAs per SM

// dynamic shared memory of the short-row kernel: 2 buffers per warp of 32/g row slots + edge values
static size_t packed_smem_bytes(int n, bool lead2n, int g)
{
    const int lead = lead2n ? 2 * n : n;
    const int delta = ((lead + 3) & ~3) - lead;
    const int rpg = 32 / g, warps = kThreads / 32;
    return static_cast<size_t>(warps) * 2 * rpg * packed_slot_chunks(n, delta, g) * 16 + static_cast<size_t>(warps) * rpg * 2 * n * 4;
}

cudaError_t sg1d_launch(int n, int variant, const W1D& w, Args1D& a, cudaStream_t stream)
{
    if (n < 1 || n > kMaxN || variant < 0 || variant >= V_COUNT) return cudaErrorInvalidValue;
    // short rows (<= 512 samples): several rows per warp instead of idle lanes
    int gi_idx = 0;
    size_t smem = 0;
    a.pack_g = 0;
    if ((variant == V_BATCH_FAST || variant == V_STREAM_FAST) && a.len <= 512 && a.rows > 1) {
        // lanes per row; wide windows on very short rows would need more shared memory per row slot than
        // two resident CTAs allow: fewer, wider slots then
        int g = a.len <= 32 ? 1 : a.len <= 64 ? 2 : a.len <= 128 ? 4 : a.len <= 256 ? 8 : 16;
        while (g < 16 && packed_smem_bytes(n, variant == V_STREAM_FAST, g) > kPackedSmemMax) g *= 2;
        a.pack_g = g;
        gi_idx = g == 1 ? 0 : g == 2 ? 1 : g == 4 ? 2 : g == 8 ? 3 : 4;
        smem = packed_smem_bytes(n, variant == V_STREAM_FAST, g);
        variant = variant == V_BATCH_FAST ? V_PACK_BATCH_FAST : V_PACK_STREAM_FAST;
    } else if (variant >= V_PACK_BATCH_FAST) {
        return cudaErrorInvalidValue;
    }
    const Kernel1D& k = sg1d_group_table((n - 1) / 4)[((n - 1) % 4) * V_COUNT + variant];

    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int bps, sms;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        GridInfo& gi = g_grid[n][variant][gi_idx];
        if (smem > 48 * 1024) {   // per device and function; cheap enough to repeat
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
        // work unit = segment of kTile (1024) outputs of one row, one warp each
        a.tiles_per_row = (a.len + kTile - 1) / kTile;
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
