// sg2d_multi.cu -- ONE launch for the components of a gradient (gx, gy) or a Hessian (hxx, hxy, hyy).
//
// The reference's wrappers (src/savgol2d.c:462-558) create one filter per component and run savgol2d_apply once per
// component: the image is read two or three times.  The components are separable sums of the same kind as every other
// filter (factor2d.cpp), only their parities differ, so the streaming kernel of sg2d_sep_kernel.cuh takes them as
// NO groups of factors: a row is staged and windowed once, every group has its own folded sums and accumulator ring,
// and a completed row is stored to NO images.  HBM traffic per pixel: 4 + 4 NO bytes instead of 8 NO.
// Instantiated for half-windows <= 8 (the static accumulator ring; four columns per lane while NO rings fit the
// register file, two beyond) and the ranks the derivative filters of order <= 5 have (1..3 factors per component);
// everything else keeps the per-component launches (capi_2d.cu).  Two translation units (this file, and
// sg2d_multi_hi.cu which includes it with SG2D_MULTI_HI defined) so that the instantiations compile in parallel.
#include "sg2d_sep_kernel.cuh"

namespace sg2d {

namespace {

constexpr int kMultiMaxN = 8;

template <int N, int RPO, int NO>
cudaError_t launch_multi_nr(const Args2D& a, const SepPlan* const* plans, const float* scales, cudaStream_t stream)
{
    constexpr int R = RPO * NO;
    // columns per lane: the NO accumulator rings hold NO * RX * (2n+2) floats, the row-pass results of a step 2 * R * RX
    constexpr int RX = NO * 4 * (2 * N + 2) + 8 * R <= 128 ? 4 : 2;
    SepW<R> w;
    std::memset(&w, 0, sizeof(w));
    for (int o = 0; o < NO; ++o) {
        const SepPlan& plan = *plans[o];
        for (int q = 0; q < RPO; ++q) {
            const int r = o * RPO + q;
            w.rc[r] = plan.row[q][plan.nx];
            for (int k = 1; k <= N; ++k) w.rk[r][k - 1] = k <= plan.nx ? plan.row[q][plan.nx + k] : 0.0f;
            for (int k = 0; k <= 2 * N; ++k) {
                const int j = k - N + plan.ny;
                w.col[r][k] = (j >= 0 && j <= 2 * plan.ny) ? plan.col[q][j] * scales[o] : 0.0f;
            }
        }
        w.sxo[o] = plan.parity_x < 0 ? -1.0f : 1.0f;
    }
    w.sx = w.sxo[0];
    static std::atomic<int> s_bps{0}, s_sms{0};
    return launch_common<N, RX>(sep_kernel<N, R, RX, false, NO>, w, a, stream, s_bps, s_sms);
}

template <int N>
cudaError_t launch_multi_n(const Args2D& a, const SepPlan* const* plans, const float* scales, int n_out, cudaStream_t stream)
{
    const int rpo = plans[0]->rank;
    if (n_out == 2) {
        switch (rpo) {
            case 1: return launch_multi_nr<N, 1, 2>(a, plans, scales, stream);
            case 2: return launch_multi_nr<N, 2, 2>(a, plans, scales, stream);
            case 3: return launch_multi_nr<N, 3, 2>(a, plans, scales, stream);
        }
    } else if (n_out == 3) {
        switch (rpo) {
            case 1: return launch_multi_nr<N, 1, 3>(a, plans, scales, stream);
            case 2: return launch_multi_nr<N, 2, 3>(a, plans, scales, stream);
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_multi_hi(int n, const Args2D& a, const SepPlan* const* plans, const float* scales, int n_out, cudaStream_t stream);

#ifdef SG2D_MULTI_HI
cudaError_t launch_multi_hi(int n, const Args2D& a, const SepPlan* const* plans, const float* scales, int n_out, cudaStream_t stream)
{
    switch (n) {
        case 5: return launch_multi_n<5>(a, plans, scales, n_out, stream);
        case 6: return launch_multi_n<6>(a, plans, scales, n_out, stream);
        case 7: return launch_multi_n<7>(a, plans, scales, n_out, stream);
        case 8: return launch_multi_n<8>(a, plans, scales, n_out, stream);
        default: return cudaErrorInvalidValue;
    }
}
#else
bool multi_supported(const Args2D& a, const SepPlan* const* plans, int n_out)
{
    if (n_out < 2 || n_out > 3 || a.rows < 1 || a.cols < 4) return false;
    const int rpo = plans[0]->rank;
    if (rpo < 1 || rpo > (n_out == 2 ? 3 : 2)) return false;
    for (int o = 0; o < n_out; ++o) {
        const SepPlan& p = *plans[o];
        if (p.rank != rpo || p.additive || p.nx != plans[0]->nx || p.ny != plans[0]->ny) return false;
        if (p.nx < 1 || p.ny < 1 || p.nx > kMultiMaxN || p.ny > kMultiMaxN) return false;
    }
    // the outputs are addressed as a.out + constant: same 16-byte phase (the row pitch is shared anyway)
    const float* outs[3] = {a.out, a.out1, a.out2};
    for (int o = 1; o < n_out; ++o)
        if ((reinterpret_cast<uintptr_t>(outs[o]) ^ reinterpret_cast<uintptr_t>(a.out)) & 15) return false;
    return true;
}

cudaError_t launch_multi(const Args2D& a, const SepPlan* const* plans, const float* scales, int n_out, cudaStream_t stream)
{
    const int n = plans[0]->nx > plans[0]->ny ? plans[0]->nx : plans[0]->ny;
    switch (n) {
        case 1: return launch_multi_n<1>(a, plans, scales, n_out, stream);
        case 2: return launch_multi_n<2>(a, plans, scales, n_out, stream);
        case 3: return launch_multi_n<3>(a, plans, scales, n_out, stream);
        case 4: return launch_multi_n<4>(a, plans, scales, n_out, stream);
        default: return launch_multi_hi(n, a, plans, scales, n_out, stream);
    }
}
#endif

}  // namespace sg2d
