// sg2d_sep.cu -- dispatch of the streaming separable 2D kernel (sg2d_sep_kernel.cuh) for generic rank-R
// factorisations; the additive surfaces W = u(x) + v(y) are instantiated in sg2d_add.cu.
#include "sg2d_sep_kernel.cuh"

namespace sg2d {

namespace {

template <int N>
cudaError_t launch_n(const Args2D& a, const SepPlan& plan, cudaStream_t stream)
{
    switch (plan.rank) {
        case 1: return launch_nr<N, 1, false>(a, plan, stream);
        case 2: return launch_nr<N, 2, false>(a, plan, stream);
        case 3: return launch_nr<N, 3, false>(a, plan, stream);
        case 4: return launch_nr<N, 4, false>(a, plan, stream);
        default: return cudaErrorInvalidValue;
    }
}

// Ticket counter of one launch: 4 bytes from a PRIVATE stream-ordered pool of the device (created on first
// use; the application's default pool and its release threshold are left alone), zeroed before and freed
// after the kernel in stream order -- private to the launch whatever other streams, threads or captured
// graphs are doing.  The pool keeps its memory, so the alloc / free pair costs about a microsecond of host time.
std::mutex g_pool_mu;
cudaMemPool_t g_pool[64] = {};

}  // namespace

cudaError_t acquire_counter(cudaStream_t stream, unsigned** out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_pool[dev]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            e = cudaMemPoolCreate(&g_pool[dev], &props);
            if (e != cudaSuccess) return e;
            unsigned long long keep = ~0ull;
            e = cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep);
            if (e != cudaSuccess) return e;
        }
        pool = g_pool[dev];
    }
    e = cudaMallocFromPoolAsync(reinterpret_cast<void**>(out), sizeof(unsigned), pool, stream);
    if (e != cudaSuccess) return e;
    return cudaMemsetAsync(*out, 0, sizeof(unsigned), stream);
}

// The streaming kernel is instantiated per half-window N = max(nx, ny) (a rectangular window runs with
// its shorter factor zero-padded).  Half-windows above 16 and the exact flavour run through sg2d_direct.cu.
bool separable_supported(const Args2D& a, const SepPlan& plan)
{
    if (plan.rank < 1 || plan.rank > kMaxRank || plan.nx < 1 || plan.ny < 1 || plan.nx > 16 || plan.ny > 16) return false;
    if (a.rows < 1 || a.cols < 4) return false;
    return true;
}

cudaError_t launch_separable(const Args2D& a, const SepPlan& plan, cudaStream_t stream)
{
    if (plan.additive) return launch_additive(a, plan, stream);
    switch (plan.nx > plan.ny ? plan.nx : plan.ny) {
#define SG2D_CASE(n) case n: return launch_n<n>(a, plan, stream);
        SG2D_CASE(1) SG2D_CASE(2) SG2D_CASE(3) SG2D_CASE(4) SG2D_CASE(5) SG2D_CASE(6) SG2D_CASE(7) SG2D_CASE(8)
        SG2D_CASE(9) SG2D_CASE(10) SG2D_CASE(11) SG2D_CASE(12) SG2D_CASE(13) SG2D_CASE(14) SG2D_CASE(15) SG2D_CASE(16)
#undef SG2D_CASE
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sg2d
