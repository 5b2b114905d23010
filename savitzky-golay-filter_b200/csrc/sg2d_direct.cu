// sg2d_direct.cu -- literal (2ny+1)x(2nx+1) window convolution.
//
// This is the general-purpose / verification path of the 2D filter: any window, any boundary,
// and an EXACT flavour that accumulates in the reference's order (row-major taps, single
// accumulator, unfused multiply/add: src/savgol2d.c:380-391, 419-451) and is therefore
// bit-identical to savgol2d_apply.  The production path for separable-representable filters is
// sg2d_sep.cu; this kernel is what it is checked against on the device.
//
// CTA = 32x8 threads, tile = 128x32 outputs (each thread 4 columns x 4 rows, register blocked);
// the tile plus its halo is staged in shared memory with the boundary rule applied while loading
// (clamp / half-sample reflect then clamp), so the tap loops are branch-free.
#include "sg2d.h"

#include <atomic>

namespace sg { extern std::atomic<unsigned long long> g_launches; }

namespace sg2d {

namespace {

constexpr int TX = 32, TY = 8;      // threads
constexpr int OX = 4, OY = 4;       // outputs per thread
constexpr int TW = TX * OX;         // 128 output columns per tile
constexpr int TH = TY * OY;         // 32 output rows per tile

__device__ __forceinline__ int map_index(int i, int n, int boundary)
{
    if (boundary == B_REFLECT) {
        if (i < 0) i = -i - 1;
        else if (i >= n) i = 2 * n - i - 1;
    }
    if (i < 0) i = 0;
    else if (i >= n) i = n - 1;
    return i;
}

template <bool EXACT>
__global__ void __launch_bounds__(TX * TY) direct_kernel(const Args2D a)
{
    extern __shared__ float smem[];
    const int ww = 2 * a.nx + 1, wh = 2 * a.ny + 1;
    float* s_w = smem;                               // ww*wh weights
    const int sw = TW + 2 * a.nx;                    // staged tile width
    const int sh = TH + 2 * a.ny;
    const int sp = sw | 1;                           // odd pitch: conflict-free column walks
    float* s_in = smem + ((ww * wh + 3) & ~3);

    const int tid = threadIdx.y * TX + threadIdx.x;
    for (int i = tid; i < ww * wh; i += TX * TY) s_w[i] = a.weights[i];

    const long long tiles_x = (a.out_cols + TW - 1) / TW;
    const long long tiles_y = (a.out_rows + TH - 1) / TH;
    const long long per_img = tiles_x * tiles_y;
    const long long total = per_img * a.n_images;

    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const long long img = t / per_img;
        const int ty = static_cast<int>((t % per_img) / tiles_x), tx = static_cast<int>(t % tiles_x);
        const int oy0 = ty * TH, ox0 = tx * TW;
        const float* in = a.in + img * a.in_image_pitch;
        float* out = a.out + img * a.out_image_pitch;

        __syncthreads();  // previous tile fully consumed (also orders the weight load on iteration 0)
        // centre of output (oy,ox) is input (oy+a.cy, ox+a.cx); the staged tile starts ny/nx before that
        for (int i = tid; i < sw * sh; i += TX * TY) {
            const int ry = i / sw, rx = i - ry * sw;
            const int iy = map_index(oy0 + a.cy - a.ny + ry, a.rows, a.boundary);
            const int ix = map_index(ox0 + a.cx - a.nx + rx, a.cols, a.boundary);
            s_in[ry * sp + rx] = in[static_cast<long long>(iy) * a.in_stride + ix];
        }
        __syncthreads();

        float acc[OY][OX];
#pragma unroll
        for (int r = 0; r < OY; ++r)
#pragma unroll
            for (int c = 0; c < OX; ++c) acc[r][c] = 0.0f;

        // thread (threadIdx.x, threadIdx.y) owns columns threadIdx.x + 32*c and rows threadIdx.y*OY + r
        const float* base = s_in + (threadIdx.y * OY) * sp + threadIdx.x;
        for (int wy = 0; wy < wh; ++wy) {
            for (int wx = 0; wx < ww; ++wx) {
                const float w = s_w[wy * ww + wx];
#pragma unroll
                for (int r = 0; r < OY; ++r)
#pragma unroll
                    for (int c = 0; c < OX; ++c) {
                        const float v = base[(wy + r) * sp + wx + 32 * c];
                        if (EXACT) acc[r][c] = __fadd_rn(acc[r][c], __fmul_rn(w, v));
                        else acc[r][c] = fmaf(w, v, acc[r][c]);
                    }
            }
        }
#pragma unroll
        for (int r = 0; r < OY; ++r) {
            const int oy = oy0 + threadIdx.y * OY + r;
            if (oy >= a.out_rows) continue;
#pragma unroll
            for (int c = 0; c < OX; ++c) {
                const int ox = ox0 + threadIdx.x + 32 * c;
                if (ox < a.out_cols)
                    out[static_cast<long long>(oy) * a.out_stride + ox] = EXACT ? __fmul_rn(acc[r][c], a.scale) : acc[r][c] * a.scale;
            }
        }
    }
}

}  // namespace

cudaError_t launch_direct(const Args2D& a, bool exact, cudaStream_t stream)
{
    const int ww = 2 * a.nx + 1, wh = 2 * a.ny + 1;
    const int sw = TW + 2 * a.nx, sh = TH + 2 * a.ny, sp = sw | 1;
    const size_t smem = (static_cast<size_t>((ww * wh + 3) & ~3) + static_cast<size_t>(sp) * sh) * sizeof(float);
    auto kern = exact ? direct_kernel<true> : direct_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const long long tiles = ((a.out_cols + TW - 1) / TW) * static_cast<long long>((a.out_rows + TH - 1) / TH) * a.n_images;
    if (tiles <= 0) return cudaSuccess;
    int dev = 0, sms = 148, bps = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, TX * TY, smem);
    if (bps < 1) bps = 1;
    long long grid = static_cast<long long>(sms) * bps;
    if (grid > tiles) grid = tiles;
    kern<<<static_cast<unsigned>(grid), dim3(TX, TY), smem, stream>>>(a);
    sg::g_launches.fetch_add(1);
    return cudaGetLastError();
}

}  // namespace sg2d
