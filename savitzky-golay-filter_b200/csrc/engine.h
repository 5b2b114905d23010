// engine.h -- internal C++ layer between the C ABI and the kernels: private filter state,
// pointer classification, scratch memory, host-buffer staging pipelines.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <mutex>

#include "../../include/savgol_b200.h"
#include "sg_common.cuh"

namespace sge {

constexpr int kMaxDevices = 16;

// The object savgol_create() really allocates: the public, ABI-visible struct first, private
// device state behind it.  Callers that copy the public struct by value lose the tail; every
// entry point therefore checks the registry (is_live) and falls back to per-call uploads.
struct FilterImpl {
    SavgolFilter pub;
    uint64_t magic;
    float* edge_t[kMaxDevices];  // device copies of the transposed edge table, uploaded on first use
};
constexpr uint64_t kFilterMagic = 0x53474232303046ULL;  // "SGB200F"

void register_filter(FilterImpl* f);
void unregister_filter(FilterImpl* f);
FilterImpl* live_filter(const SavgolFilter* f);  // nullptr when f was not made by savgol_create()

// Memory kind of a caller pointer.
enum class MemKind { Device, Pinned, Pageable };
MemKind classify(const void* p);

cudaStream_t current_stream();
bool device_ready(bool complain);
int exact_mode();

// Logs "savgol_b200: <what>: <cuda error>" and returns false when e != cudaSuccess.
bool cuda_ok(cudaError_t e, const char* what);

// One batched 1D problem with device-resident operands (the kernel-level contract).
struct Problem1D {
    const SavgolFilter* filter;
    const void* in;  void* out;
    size_t rows, len;
    size_t out_len;                       // 0 = len; otherwise only outputs [0,out_len) are stored
    size_t in_row_bytes, out_row_bytes;   // byte pitch between signals
    size_t in_stride, out_stride;         // bytes between samples
    const float* lhalo; const float* rhalo; size_t lhalo_pitch, rhalo_pitch;  // optional, elements
    int mode;                             // sg::MODE_*
    bool edge_lead, edge_trail;           // polynomial edge tables at the true ends
    bool stream_history;                  // lhalo holds 2n carried samples (stream steady state)
    float* state_out; size_t state_pitch; int state_w;
    int arith;                            // sg::ARITH_*
};

// Launches the kernel(s) for `p` on `stream`.  Device pointers only.  Handles aliasing of in/out.
bool run1d_device(const Problem1D& p, cudaStream_t stream);

// Device copy of the transposed polynomial edge table of `f` on the current device.
// *temp is set when the table had to be uploaded into a temporary (caller frees after the launch).
const float* edge_table_device(const SavgolFilter* f, cudaStream_t stream, float** temp);

// Host-buffer staging.  Three device slots form a ring; H2D, kernel and D2H of consecutive
// chunks overlap on three streams.  Pinned host memory is DMA'd directly; pageable memory goes
// through the driver's staging (cudaMemcpyAsync degrades gracefully to a synchronous copy).
struct Pipeline {
    static constexpr int kSlots = 3;
    int dev = -1;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t e_in[kSlots] = {}, e_k[kSlots] = {}, e_out[kSlots] = {};
    float* d_in[kSlots] = {};
    float* d_out[kSlots] = {};
    size_t cap_in = 0, cap_out = 0;  // floats per slot

    bool ensure(size_t need_in, size_t need_out)
    {
        int cur = 0;
        if (!sge::cuda_ok(cudaGetDevice(&cur), "cudaGetDevice")) return false;
        if (dev != cur) { release(); dev = cur; }
        if (!s_in) {
            if (!sge::cuda_ok(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking), "stream")) return false;
            if (!sge::cuda_ok(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking), "stream")) return false;
            if (!sge::cuda_ok(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking), "stream")) return false;
            for (int i = 0; i < kSlots; ++i) {
                cudaEventCreateWithFlags(&e_in[i], cudaEventDisableTiming);
                cudaEventCreateWithFlags(&e_k[i], cudaEventDisableTiming);
                cudaEventCreateWithFlags(&e_out[i], cudaEventDisableTiming);
            }
        }
        if (need_in > cap_in) {
            for (int i = 0; i < kSlots; ++i) { cudaFree(d_in[i]); d_in[i] = nullptr; }
            for (int i = 0; i < kSlots; ++i)
                if (!sge::cuda_ok(cudaMalloc(&d_in[i], need_in * sizeof(float)), "cudaMalloc(staging in)")) return false;
            cap_in = need_in;
        }
        if (need_out > cap_out) {
            for (int i = 0; i < kSlots; ++i) { cudaFree(d_out[i]); d_out[i] = nullptr; }
            for (int i = 0; i < kSlots; ++i)
                if (!sge::cuda_ok(cudaMalloc(&d_out[i], need_out * sizeof(float)), "cudaMalloc(staging out)")) return false;
            cap_out = need_out;
        }
        return true;
    }
    void release()
    {
        for (int i = 0; i < kSlots; ++i) {
            if (d_in[i]) cudaFree(d_in[i]);
            if (d_out[i]) cudaFree(d_out[i]);
            d_in[i] = d_out[i] = nullptr;
            if (e_in[i]) { cudaEventDestroy(e_in[i]); cudaEventDestroy(e_k[i]); cudaEventDestroy(e_out[i]); }
            e_in[i] = e_k[i] = e_out[i] = nullptr;
        }
        if (s_in) { cudaStreamDestroy(s_in); cudaStreamDestroy(s_k); cudaStreamDestroy(s_out); }
        s_in = s_k = s_out = nullptr;
        cap_in = cap_out = 0;
    }
};



extern Pipeline g_pipe;      // one per process, guarded by g_pipe_mu (host-pointer calls serialise)
extern std::mutex g_pipe_mu;

}  // namespace sge
