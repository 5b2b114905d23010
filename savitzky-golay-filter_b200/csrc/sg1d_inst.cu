// sg1d_inst.cu -- explicit instantiations of sg1d_kernel for four half-windows.
// Compiled eight times (-DSG_GROUP=0..7) so the 352 fully unrolled kernels build in parallel.
#include "sg1d_kernel.cuh"
#include "sg1d_packed.cuh"
#include "sg1d_launch.h"
#include "sg1d_tma.cuh"

#ifndef SG_GROUP
#error "compile with -DSG_GROUP=<0..7>"
#endif

namespace sg {

#define SG_ROW(N)                                                                                   \
    {&sg1d_kernel<N, false, ARITH_FAST>}, {&sg1d_kernel<N, false, ARITH_EXACT4>},                   \
    {&sg1d_kernel<N, false, ARITH_EXACTSEQ>}, {&sg1d_kernel<N, true, ARITH_FAST>},                  \
    {&sg1d_kernel<N, true, ARITH_EXACTSEQ>}, {&sg1d_packed_kernel<N, false, false>}, {&sg1d_packed_kernel<N, true, false>}, \
    {&sg1d_packed_kernel<N, false, true>}, {&sg1d_packed_kernel<N, true, true>},                \
    {&sg1d_kernel<N, false, ARITH_FAST, true>}, {&sg1d_kernel<N, true, ARITH_FAST, true>}

#define SG_CAT_(a, b) a##b
#define SG_CAT(a, b) SG_CAT_(a, b)

static const Kernel1D SG_CAT(kTable, SG_GROUP)[4 * V_COUNT] = {
    SG_ROW(4 * SG_GROUP + 1), SG_ROW(4 * SG_GROUP + 2), SG_ROW(4 * SG_GROUP + 3), SG_ROW(4 * SG_GROUP + 4)};

const Kernel1D* SG_CAT(sg1d_group_table_, SG_GROUP)() { return SG_CAT(kTable, SG_GROUP); }

#define SG_TROW(N) {&sg1d_tma_kernel<N, false>}, {&sg1d_tma_kernel<N, true>}
static const Kernel1DTma SG_CAT(kTmaTable, SG_GROUP)[4 * VT_COUNT] = {
    SG_TROW(4 * SG_GROUP + 1), SG_TROW(4 * SG_GROUP + 2), SG_TROW(4 * SG_GROUP + 3), SG_TROW(4 * SG_GROUP + 4)};
const Kernel1DTma* SG_CAT(sg1d_tma_group_table_, SG_GROUP)() { return SG_CAT(kTmaTable, SG_GROUP); }

}  // namespace sg
