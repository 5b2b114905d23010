// sg1d_launch.h -- host-side dispatch table of the 1D kernel instantiations.
#pragma once
#include <cuda.h>

#include "sg_common.cuh"

namespace sg {

// Variant index inside a half-window's row of the table.
enum : int { V_BATCH_FAST = 0, V_BATCH_EXACT4 = 1, V_BATCH_EXACTSEQ = 2, V_STREAM_FAST = 3, V_STREAM_EXACTSEQ = 4,
              V_PACK_BATCH_FAST = 5, V_PACK_STREAM_FAST = 6,  // short rows, several per warp (sg1d_packed.cuh)
              V_PACK_BATCH_FAST_PH = 7, V_PACK_STREAM_FAST_PH = 8,  // ... misaligned rows on a per-row phase
              V_BATCH_FAST_PT = 9, V_STREAM_FAST_PT = 10,  // generic kernel with per-row phases / short tails (sg1d_kernel.cuh, PT)
              V_COUNT = 11 };

struct Kernel1D {
    void (*kernel)(const W1D, const Args1D);  // __global__ entry
};

// Tensor maps of the TMA kernels (sg1d_tma.cuh): the batch seen as {32 floats, len / 32, rows}, SWIZZLE_128B.
struct alignas(64) TmaMaps {
    CUtensorMap in_full;   // box {32, HL + 33, 1}: left halo rows + the 1024 samples under a segment's outputs + right halo row
    CUtensorMap in_first;  // box {32, 33, 1}: body + right halo row (first segment of a signal)
    CUtensorMap in_last;   // box {32, HL + 32, 1}: left halo rows + body (last full segment)
    CUtensorMap in_body;   // box {32, 32, 1}: a signal that is one segment
    CUtensorMap out_body;  // box {32, 32, 1}
};
struct Kernel1DTma {
    void (*kernel)(const W1D, const Args1D, const TmaMaps);
};
enum : int { VT_BATCH = 0, VT_STREAM = 1, VT_COUNT = 2 };

// Defined once per instantiation group (sg1d_inst.cu compiled with -DSG_GROUP=g covers
// half-windows 4g+1 .. 4g+4).
const Kernel1D* sg1d_group_table(int group);
const Kernel1DTma* sg1d_tma_group_table(int group);

// cuTensorMapEncodeTiled, resolved at run time through the runtime's driver entry point table (the library has to
// load on machines without libcuda).  nullptr when unavailable.
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiled encode_tiled();

// Dispatch decision of one launch (sg1d_launch.cu: sg1d_plan).
enum : int { PLAN_GENERIC = 0, PLAN_PACKED = 1, PLAN_TMA = 2 };
struct Plan1D {
    int family;      // PLAN_*
    int variant;     // table column actually launched (V_*)
    int gi_idx;      // occupancy cache slot (packed kernel: packing class x edge values)
    size_t smem;     // dynamic shared memory (packed kernel)
};
Plan1D sg1d_plan(int n, int variant, Args1D& args, bool allow_tma);

// Grid sizing + launch.  Returns cudaSuccess or the launch error.
cudaError_t sg1d_launch(int n, int variant, const W1D& w, Args1D& args, cudaStream_t stream);

}  // namespace sg
