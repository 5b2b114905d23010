// sg2d_add.cu -- instantiations of the streaming 2D kernel for ADDITIVE weight surfaces
//     W(y,x) = u(x) + v(y)
// (square windows up to 33x33; 4 columns per lane up to 17x17, 2 above).  Every order-2/3 smoothing filter is of this form, and so are its even/even
// derivatives and the fused Laplacian: the least-squares weights of a total-degree <= 3 fit evaluated at the
// window centre lie in span{1, x^2, y^2} (ref: src/savgol2d.c:188-265 computes them as a dense table).
// The "1" factors are box sums, which the kernel gets almost for free -- see sg2d_sep_kernel.cuh.
#include "sg2d_sep_kernel.cuh"

namespace sg2d {

cudaError_t launch_additive(const Args2D& a, const SepPlan& plan, cudaStream_t stream)
{
    if (!plan.additive || plan.nx != plan.ny) return cudaErrorInvalidValue;
    switch (plan.nx) {
#define SG2D_CASE(n) case n: return launch_nr<n, 1, true>(a, plan, stream);
        SG2D_CASE(1) SG2D_CASE(2) SG2D_CASE(3) SG2D_CASE(4) SG2D_CASE(5) SG2D_CASE(6) SG2D_CASE(7) SG2D_CASE(8)
        SG2D_CASE(9) SG2D_CASE(10) SG2D_CASE(11) SG2D_CASE(12) SG2D_CASE(13) SG2D_CASE(14) SG2D_CASE(15) SG2D_CASE(16)
#undef SG2D_CASE
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sg2d
