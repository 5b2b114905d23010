// capi_multi.cu -- several GPUs from ONE process through the C ABI (include/savgol_b200.h part 2, "multi-GPU").
//
// The reference's only kind of user is a C caller looping over savgol_apply (include/iterative/savgolFilter.h:16-19).
// These entry points give that caller all the GPUs of a box without a process group, NCCL or Python
// (SURVEY.md 8b rule 4, 8e):
//   * savgol_apply_batch_multi  -- HOST batch: independent signals are sharded over the devices with no
//     communication at all; a batch with fewer signals than devices (one very long signal) is partitioned along
//     its length, each slice staged with its n-sample halos straight from the host signal.  One host thread per
//     device, each with its own staging pipeline, so the devices' PCIe links run concurrently.
//   * savgol_apply_slices       -- DEVICE-resident slices of one long signal, slice i on devices[i]: peer access
//     is enabled between ring neighbours and every device's kernel reads its 2n halo samples directly from the
//     neighbour's HBM over NVLink while it stages its first / last segment (the same fused exchange
//     dist.PeerRing sets up across processes with CUDA IPC).  No copy, no collective, one launch per device.
//   * savgol_b200_device_count / _alloc / _free / _copy -- the four calls a plain-C program needs to own
//     device memory without CUDA headers (examples/c_multi_gpu.c).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "engine.h"

using sge::cuda_ok;
using sge::MemKind;

namespace {

int mode_of(const SavgolFilter* f)
{
    switch (f->config.boundary) {
        case SAVGOL_BOUNDARY_REFLECT: return sg::MODE_REFLECT;
        case SAVGOL_BOUNDARY_PERIODIC: return sg::MODE_PERIODIC;
        case SAVGOL_BOUNDARY_CONSTANT: return sg::MODE_CONSTANT;
        default: return sg::MODE_POLY;
    }
}

bool devices_ok(const int* devices, int n_devices, const char* who)
{
    int count = 0;
    if (!devices || n_devices < 1 || cudaGetDeviceCount(&count) != cudaSuccess) {
        (void)cudaGetLastError();
        fprintf(stderr, "%s: no device list / no CUDA device\n", who);
        return false;
    }
    for (int i = 0; i < n_devices; ++i)
        if (devices[i] < 0 || devices[i] >= count || devices[i] >= sge::kMaxDevices) {
            fprintf(stderr, "%s: device %d does not exist (%d visible)\n", who, devices[i], count);
            return false;
        }
    return true;
}

// one non-blocking stream per device for savgol_apply_slices
std::mutex g_mu;
cudaStream_t g_stream[sge::kMaxDevices] = {};
bool g_peer[sge::kMaxDevices][sge::kMaxDevices] = {};

cudaStream_t slice_stream(int dev)   // current device == dev
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_stream[dev] && !cuda_ok(cudaStreamCreateWithFlags(&g_stream[dev], cudaStreamNonBlocking), "stream")) return nullptr;
    return g_stream[dev];
}

// current device == dev; makes `peer`'s memory loadable from kernels running on dev
bool enable_peer(int dev, int peer)
{
    if (dev == peer) return true;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_peer[dev][peer]) return true;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, dev, peer) != cudaSuccess || !can) { (void)cudaGetLastError(); return false; }
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return false; }
    (void)cudaGetLastError();
    g_peer[dev][peer] = true;
    return true;
}

}  // namespace

extern "C" {

int savgol_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

void* savgol_b200_alloc(int device, size_t bytes)
{
    sge::DeviceGuard g(device);
    void* p = nullptr;
    return cuda_ok(cudaMalloc(&p, bytes), "savgol_b200_alloc") ? p : nullptr;
}

void savgol_b200_free(int device, void* ptr)
{
    if (!ptr) return;
    sge::DeviceGuard g(device);
    cudaFree(ptr);
}

int savgol_b200_copy(void* dst, const void* src, size_t bytes)
{
    return cuda_ok(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault), "savgol_b200_copy") ? 0 : -1;
}

int savgol_apply_batch_multi(const SavgolFilter* filter, const float* input, float* output, size_t n_signals, size_t length,
                             size_t in_pitch, size_t out_pitch, const int* devices, int n_devices)
{
    if (filter == nullptr || input == nullptr || output == nullptr) {
        fprintf(stderr, "savgol_apply: NULL pointer\n");
        return -1;
    }
    if (length < static_cast<size_t>(filter->window_size)) {
        fprintf(stderr, "savgol_apply: data length (%lu) < window size (%d)\n", static_cast<unsigned long>(length), filter->window_size);
        return -1;
    }
    if (n_signals == 0) return 0;
    if (in_pitch < length || out_pitch < length) {
        if (n_signals > 1) { fprintf(stderr, "savgol_apply_batch_multi: pitch < length\n"); return -1; }
        in_pitch = out_pitch = length;
    }
    if (!devices_ok(devices, n_devices, "savgol_apply_batch_multi")) return -1;
    if (sge::classify(input) == MemKind::Device || sge::classify(output) == MemKind::Device) {
        fprintf(stderr, "savgol_apply_batch_multi: host buffers only (device buffers belong to one GPU: savgol_apply_batch / savgol_apply_slices)\n");
        return -1;
    }
    const int mode = mode_of(filter);
    const bool poly = mode == sg::MODE_POLY;
    const int arith = sge::exact_mode() ? sg::ARITH_EXACT4 : sg::ARITH_FAST;   // the caller's flavour, handed to the workers
    const size_t ws = static_cast<size_t>(filter->window_size);
    const size_t nd = static_cast<size_t>(n_devices);

    // work split: whole signals per device when there are enough of them, else every signal cut along its length
    const bool by_rows = n_signals >= nd;
    size_t parts = nd;
    if (!by_rows) parts = std::max<size_t>(1, std::min(nd, length / std::max<size_t>(ws, 1024)));   // slices of >= one window
    std::vector<int> rc(nd, 0);
    std::vector<std::thread> th;
    for (size_t i = 0; i < (by_rows ? nd : parts); ++i) {
        th.emplace_back([&, i] {
            if (!cuda_ok(cudaSetDevice(devices[i]), "cudaSetDevice") || !sge::device_ready(true)) { rc[i] = -1; return; }
            if (by_rows) {
                const size_t r0 = n_signals * i / nd, r1 = n_signals * (i + 1) / nd;
                if (r1 > r0 && !sge::run1d_host(filter, input + r0 * in_pitch, output + r0 * out_pitch, r1 - r0, length, in_pitch, out_pitch, mode, poly, arith))
                    rc[i] = -1;
            } else {
                sge::PipeLease lease;
                if (!lease.ok()) { rc[i] = -1; return; }
                const size_t a = length * i / parts, b = length * (i + 1) / parts;
                for (size_t r = 0; r < n_signals && rc[i] == 0; ++r)
                    if (!sge::run1d_host_range(*lease, filter, input + r * in_pitch, length, a, b, output + r * out_pitch + a, mode, poly, arith))
                        rc[i] = -1;
            }
        });
    }
    for (auto& t : th) t.join();
    for (int v : rc)
        if (v != 0) return -1;
    return 0;
}

int savgol2d_apply_batch_multi(const Savgol2DFilter* filter, const float* input, int rows, int cols, int in_stride, size_t in_image_pitch,
                               float* output, int out_stride, size_t out_image_pitch, size_t n_images, Savgol2DBoundary boundary,
                               const int* devices, int n_devices)
{
    if (!filter || !input || !output) return -1;
    if (n_images == 0) return 0;
    if (!devices_ok(devices, n_devices, "savgol2d_apply_batch_multi")) return -1;
    if (sge::classify(input) == MemKind::Device || sge::classify(output) == MemKind::Device) {
        fprintf(stderr, "savgol2d_apply_batch_multi: host buffers only (device buffers belong to one GPU: savgol2d_apply_batch)\n");
        return -1;
    }
    // independent images: contiguous blocks per device, one host thread (and one staging pipeline) per device
    const size_t nd = std::min<size_t>(static_cast<size_t>(n_devices), n_images);
    const int exact = sge::exact_mode();
    std::vector<int> rc(nd, 0);
    std::vector<std::thread> th;
    for (size_t i = 0; i < nd; ++i) {
        th.emplace_back([&, i] {
            if (!cuda_ok(cudaSetDevice(devices[i]), "cudaSetDevice")) { rc[i] = -1; return; }
            savgol_b200_set_exact(exact);   // the caller's flavour
            const size_t i0 = n_images * i / nd, i1 = n_images * (i + 1) / nd;
            if (i1 > i0)
                rc[i] = savgol2d_apply_batch(filter, input + i0 * in_image_pitch, rows, cols, in_stride, in_image_pitch,
                                             output + i0 * out_image_pitch, out_stride, out_image_pitch, i1 - i0, boundary);
        });
    }
    for (auto& t : th) t.join();
    for (int v : rc)
        if (v != 0) return -1;
    return 0;
}

int savgol_apply_slices(const SavgolFilter* filter, const float* const* in_slices, float* const* out_slices, const size_t* lengths,
                        const int* devices, int n_slices)
{
    if (filter == nullptr || in_slices == nullptr || out_slices == nullptr || lengths == nullptr) {
        fprintf(stderr, "savgol_apply_slices: NULL pointer\n");
        return -1;
    }
    if (!devices_ok(devices, n_slices, "savgol_apply_slices")) return -1;
    const size_t n = filter->config.half_window;
    for (int i = 0; i < n_slices; ++i) {
        if (!in_slices[i] || !out_slices[i]) { fprintf(stderr, "savgol_apply_slices: NULL slice\n"); return -1; }
        if (lengths[i] < static_cast<size_t>(filter->window_size)) {
            fprintf(stderr, "savgol_apply_slices: slice length (%lu) < window size (%d)\n", static_cast<unsigned long>(lengths[i]), filter->window_size);
            return -1;
        }
    }
    const int mode = mode_of(filter);
    const bool periodic = mode == sg::MODE_PERIODIC;
    const int arith = sge::exact_mode() ? sg::ARITH_EXACT4 : sg::ARITH_FAST;
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    bool ok = true;
    std::vector<float*> temp;   // halo strips copied by the driver where a pair of devices has no peer access
    std::vector<int> temp_dev;
    for (int i = 0; i < n_slices && ok; ++i) {
        const int dev = devices[i];
        ok = cuda_ok(cudaSetDevice(dev), "cudaSetDevice") && sge::device_ready(true);
        if (!ok) break;
        cudaStream_t st = slice_stream(dev);
        if (!st) { ok = false; break; }
        const int il = i > 0 ? i - 1 : (periodic ? n_slices - 1 : -1);
        const int ir = i + 1 < n_slices ? i + 1 : (periodic ? 0 : -1);
        const float* lh = il >= 0 ? in_slices[il] + (lengths[il] - n) : nullptr;
        const float* rh = ir >= 0 ? in_slices[ir] : nullptr;
        if (n_slices == 1) lh = rh = nullptr;   // one slice: the kernel wraps / synthesises on its own
        for (int side = 0; side < 2; ++side) {
            const float*& h = side ? rh : lh;
            const int other = side ? ir : il;
            if (!h || devices[other] == dev || enable_peer(dev, devices[other])) continue;
            float* strip = nullptr;   // no peer access between the two devices: the driver moves the 2n samples
            ok = cuda_ok(cudaMalloc(&strip, n * sizeof(float)), "cudaMalloc(halo strip)") &&
                 cuda_ok(cudaMemcpyPeerAsync(strip, dev, h, devices[other], n * sizeof(float), st), "cudaMemcpyPeerAsync");
            if (strip) { temp.push_back(strip); temp_dev.push_back(dev); }
            h = strip;
            if (!ok) break;
        }
        if (!ok) break;
        sge::Problem1D p{};
        p.filter = filter; p.in = in_slices[i]; p.out = out_slices[i]; p.rows = 1; p.len = lengths[i];
        p.in_row_bytes = p.out_row_bytes = lengths[i] * sizeof(float);
        p.in_stride = p.out_stride = 4;
        p.lhalo = lh; p.rhalo = rh;
        p.mode = mode;
        p.edge_lead = mode == sg::MODE_POLY && lh == nullptr;
        p.edge_trail = mode == sg::MODE_POLY && rh == nullptr;
        p.arith = arith;
        ok = sge::run1d_device(p, st);
    }
    for (int i = 0; i < n_slices; ++i) {   // the call returns when every slice is done (host-visible completion)
        if (cudaSetDevice(devices[i]) != cudaSuccess) { ok = false; continue; }
        cudaStream_t st = slice_stream(devices[i]);
        if (st && !cuda_ok(cudaStreamSynchronize(st), "sync")) ok = false;
    }
    for (size_t k = 0; k < temp.size(); ++k) {
        cudaSetDevice(temp_dev[k]);
        cudaFree(temp[k]);
    }
    cudaSetDevice(prev_dev);
    return ok ? 0 : -1;
}

}  // extern "C"
