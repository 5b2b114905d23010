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
bool device_ready(bool complain);   // the CURRENT device is a compute-capability 10.x part
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
// chunks overlap on three streams.  Pinned host memory is DMA'd directly.  Pageable memory (malloc, numpy)
// cannot be: the driver would stage it synchronously at ~7 GB/s, so it is moved through pinned bounce
// buffers by a few host threads (copy_threads()), which overlaps with the DMA of the neighbouring chunks.
// A pipeline belongs to ONE device and is used by ONE host-pointer call at a time (PipeLease).
struct Pipeline {
    static constexpr int kSlots = 3;
    int dev = -1;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t e_in[kSlots] = {}, e_k[kSlots] = {}, e_out[kSlots] = {};
    float* d_in[kSlots] = {};
    float* d_out[kSlots] = {};
    size_t cap_in = 0, cap_out = 0;  // floats per slot
    // pinned bounce buffers, one per slot and direction, allocated when a call brings pageable memory
    float* h_in[kSlots] = {};
    float* h_out[kSlots] = {};
    size_t cap_h_in = 0, cap_h_out = 0;
    bool bounce_in = false, bounce_out = false;   // this call's input / output is pageable
    struct Pending { float* dst; size_t dst_pitch, width, rows; unsigned long long seq; };
    Pending pend[kSlots] = {};                    // D2H landed (or landing) in h_out[slot], still to be handed to the caller
    unsigned long long seq = 0;

    // Start of a host-pointer call: classifies its buffers.  Returns the staging chunk in floats: the configured
    // chunk (64 MiB; 16 MiB through bounce buffers), shrunk for calls of `total` floats that would otherwise fit one or
    // two chunks and leave H2D, kernel and D2H nothing to overlap with (about eight chunks, never below 2 MiB).
    size_t begin(const void* in, const void* out, size_t total = 0);
    // Streams / events on first use, slots (and bounce buffers, when this call needs them) grown on demand.
    // The current device must be `dev`.
    bool ensure(size_t need_in, size_t need_out);
    // Before slot `s` is refilled: hands a pending bounced output to the caller and orders s_in behind the slot's D2H.
    bool reuse(int s);
    // Copies (floats: pitches, width) on s_in / s_out.  The caller records e_in[s] / e_out[s] afterwards.
    bool h2d(int s, float* dev_dst, size_t dev_pitch, const float* src, size_t src_pitch, size_t width, size_t rows);
    bool d2h(int s, float* dst, size_t dst_pitch, const float* dev_src, size_t dev_pitch, size_t width, size_t rows);
    // End of the call: pending outputs handed over (oldest first), all three streams drained.
    bool finish();
    void release();
};

// Multi-threaded host copy of `rows` rows of `width` BYTES (pitches in bytes); used by the bounce path.
void host_copy2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width, size_t rows);
int copy_threads();

// Lease of an idle pipeline of the CURRENT device for the duration of one host-pointer call.  Concurrent
// calls (several host threads, several devices) each get their own pipeline: nothing serialises them
// except the PCIe link itself.  Returned pipelines are kept for reuse (a few per device).
class PipeLease {
  public:
    PipeLease();
    ~PipeLease();
    PipeLease(const PipeLease&) = delete;
    PipeLease& operator=(const PipeLease&) = delete;
    bool ok() const { return p_ != nullptr; }
    Pipeline* operator->() const { return p_; }
    Pipeline& operator*() const { return *p_; }

  private:
    Pipeline* p_ = nullptr;
};

// Makes the device that owns `ptr` current for the lifetime of the guard (no-op for host pointers and when it
// already is current).
class DeviceGuard {
  public:
    explicit DeviceGuard(const void* ptr);
    explicit DeviceGuard(int device);
    ~DeviceGuard();
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;

  private:
    int prev_ = -1;
};

// Host staging of one long signal: outputs [a, b) of x[0..L) into y (y[0] <-> output a), cut into pieces
// with explicit n-sample halos; samples outside [0, L) follow `mode` / the polynomial edges as in savgol_apply.
// Alias safe: y may be x + a (in place).
bool run1d_host_range(Pipeline& P, const SavgolFilter* f, const float* x, size_t L, size_t a, size_t b, float* y,
                      int mode, bool poly_edges, int arith);
// Batch of contiguous-sample rows in HOST memory (chunks of whole rows, or run1d_host_range per row when a row
// is longer than a staging chunk).  Uses a pipeline of the current device.
bool run1d_host(const SavgolFilter* f, const float* in, float* out, size_t rows, size_t len, size_t in_pitch,
                size_t out_pitch, int mode, bool poly_edges, int arith);
size_t chunk_floats(bool bounce = false);
size_t staging_chunk(size_t total, bool bounce);

}  // namespace sge
