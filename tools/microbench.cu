// Stencil-shaped instruction-throughput microbenchmark for sm_100a.
// Answers, before any kernel design is frozen: how many fp32 FMA lanes per clock per SM does a
// register-sliding-window stencil sustain when the weight operand is (a) a kernel-parameter
// constant-bank operand, (b) a register, (c) packed FFMA2 (fma.rn.f32x2) with register pairs?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int TAPS = 33;
constexpr int R = 16;

struct Weights { float w[TAPS]; };
struct Weights2 { float2 w[TAPS]; };

// (a) scalar FFMA, weights straight from the kernel parameter space (constant bank 0)
__global__ void __launch_bounds__(128) k_const(const __grid_constant__ Weights W, const float* __restrict__ in,
                                               float* __restrict__ out, int iters) {
    float x[R + TAPS - 1];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < R + TAPS - 1; ++i) x[i] = in[(t + i) & 1023];
    float acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < TAPS; ++k)
#pragma unroll
            for (int j = 0; j < R; ++j) acc[j] = fmaf(W.w[k], x[j + k], acc[j]);
#pragma unroll
        for (int j = 0; j < R; ++j) x[j] = acc[j];   // loop-carried so nothing hoists
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < R; ++j) s += acc[j];
    out[t] = s;
}

// (b) scalar FFMA, weights in registers
__global__ void __launch_bounds__(128) k_reg(const float* __restrict__ wg, const float* __restrict__ in,
                                             float* __restrict__ out, int iters) {
    float x[R + TAPS - 1], w[TAPS];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < TAPS; ++i) w[i] = wg[i];
#pragma unroll
    for (int i = 0; i < R + TAPS - 1; ++i) x[i] = in[(t + i) & 1023];
    float acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < TAPS; ++k)
#pragma unroll
            for (int j = 0; j < R; ++j) acc[j] = fmaf(w[k], x[j + k], acc[j]);
#pragma unroll
        for (int j = 0; j < R; ++j) x[j] = acc[j];
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < R; ++j) s += acc[j];
    out[t] = s;
}

// (c) packed FFMA2: two rows per thread, x and acc are (rowA,rowB) pairs, weights (w,w) from params
__global__ void __launch_bounds__(128) k_f2_const(const __grid_constant__ Weights2 W, const float* __restrict__ in,
                                                  float* __restrict__ out, int iters) {
    float2 x[R + TAPS - 1];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < R + TAPS - 1; ++i) x[i] = make_float2(in[(t + i) & 1023], in[(t + 2 * i) & 1023]);
    float2 acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < TAPS; ++k)
#pragma unroll
            for (int j = 0; j < R; ++j) acc[j] = __ffma2_rn(W.w[k], x[j + k], acc[j]);
#pragma unroll
        for (int j = 0; j < R; ++j) x[j] = acc[j];
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < R; ++j) s += acc[j].x + acc[j].y;
    out[t] = s;
}

// (d) packed FFMA2, weight pairs in registers
__global__ void __launch_bounds__(128) k_f2_reg(const float* __restrict__ wg, const float* __restrict__ in,
                                                float* __restrict__ out, int iters) {
    float2 x[R + TAPS - 1], w[TAPS];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < TAPS; ++i) w[i] = make_float2(wg[i], wg[i]);
#pragma unroll
    for (int i = 0; i < R + TAPS - 1; ++i) x[i] = make_float2(in[(t + i) & 1023], in[(t + 2 * i) & 1023]);
    float2 acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < TAPS; ++k)
#pragma unroll
            for (int j = 0; j < R; ++j) acc[j] = __ffma2_rn(w[k], x[j + k], acc[j]);
#pragma unroll
        for (int j = 0; j < R; ++j) x[j] = acc[j];
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < R; ++j) s += acc[j].x + acc[j].y;
    out[t] = s;
}

template <class F>
static float timeit(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, max clock %d MHz\n", p.name, p.multiProcessorCount, clk_khz / 1000);
    float *in, *out, *wg;
    cudaMalloc(&in, 4096); cudaMalloc(&wg, 4096);
    cudaMemset(in, 0, 4096); cudaMemset(wg, 0, 4096);
    const int iters = 2000;
    Weights W; Weights2 W2;
    for (int i = 0; i < TAPS; ++i) { W.w[i] = 1e-3f * i; W2.w[i] = make_float2(W.w[i], W.w[i]); }
    for (int warps_per_sm : {4, 8, 16}) {
        const int blocks = p.multiProcessorCount * warps_per_sm / 4;
        cudaMalloc(&out, sizeof(float) * blocks * 128);
        const double fma_scalar = double(blocks) * 128 * iters * TAPS * R;
        auto rep = [&](const char* name, float ms, double fmas) {
            double per_clk_sm = fmas / (ms * 1e-3) / p.multiProcessorCount / (clk_khz * 1e3);
            printf("  %-10s warps/SM %2d: %8.3f ms  %7.2f TFMA/s  %6.1f FMA/clk/SM (at max clock)\n", name,
                   warps_per_sm, ms, fmas / ms * 1e-9, per_clk_sm);
        };
        rep("ffma_const", timeit([&] { k_const<<<blocks, 128>>>(W, in, out, iters); }), fma_scalar);
        rep("ffma_reg", timeit([&] { k_reg<<<blocks, 128>>>(wg, in, out, iters); }), fma_scalar);
        rep("ffma2_const", timeit([&] { k_f2_const<<<blocks, 128>>>(W2, in, out, iters); }), 2 * fma_scalar);
        rep("ffma2_reg", timeit([&] { k_f2_reg<<<blocks, 128>>>(wg, in, out, iters); }), 2 * fma_scalar);
        cudaFree(out);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
