// capi_1d.cu -- C ABI of the 1D batch path (include/savgol_b200.h part 1a + batch/halo extensions).
// Mirrors the reference's argument checks, return codes and stderr messages
// (src/savgolFilter.c:688-934); all arithmetic on signals happens in sg1d_kernel.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "coeffs.h"
#include "engine.h"

using sge::cuda_ok;
using sge::MemKind;

namespace {

int arith_for_batch() { return sge::exact_mode() ? sg::ARITH_EXACT4 : sg::ARITH_FAST; }

int mode_of(const SavgolFilter* f)
{
    switch (f->config.boundary) {
        case SAVGOL_BOUNDARY_REFLECT: return sg::MODE_REFLECT;
        case SAVGOL_BOUNDARY_PERIODIC: return sg::MODE_PERIODIC;
        case SAVGOL_BOUNDARY_CONSTANT: return sg::MODE_CONSTANT;
        default: return sg::MODE_POLY;
    }
}

// Dispatch on where the caller's buffers live.
bool run1d_any(const SavgolFilter* f, const float* in, float* out, size_t rows, size_t len,
               size_t in_pitch, size_t out_pitch, int mode, bool poly_edges)
{
    const MemKind ki = sge::classify(in), ko = sge::classify(out);
    sge::DeviceGuard guard(in);   // device buffers: run on the device that owns them, whatever is current
    if (!sge::device_ready(true)) return false;
    const int arith = arith_for_batch();
    if (ki == MemKind::Device && ko == MemKind::Device) {
        sge::Problem1D p{};
        p.filter = f; p.in = in; p.out = out; p.rows = rows; p.len = len;
        p.in_row_bytes = in_pitch * sizeof(float); p.out_row_bytes = out_pitch * sizeof(float);
        p.in_stride = p.out_stride = 4;
        p.mode = mode; p.edge_lead = p.edge_trail = poly_edges; p.arith = arith;
        return sge::run1d_device(p, sge::current_stream());
    }
    if (ki != MemKind::Device && ko != MemKind::Device) {
        // Experiment knob: pinned host buffers can be read / written by the kernel directly over PCIe
        // (zero copy, no staging ring).  Off by default -- see DESIGN.md section 5 for the measurement.
        static const int zero_copy = [] { const char* e = getenv("SAVGOL_B200_ZEROCOPY"); return (e && e[0] == '1') ? 1 : 0; }();
        if (zero_copy && ki == MemKind::Pinned && ko == MemKind::Pinned && in != out) {
            void *din = nullptr, *dout = nullptr;
            if (cudaHostGetDevicePointer(&din, const_cast<float*>(in), 0) == cudaSuccess &&
                cudaHostGetDevicePointer(&dout, out, 0) == cudaSuccess) {
                sge::Problem1D p{};
                p.filter = f; p.in = din; p.out = dout; p.rows = rows; p.len = len;
                p.in_row_bytes = in_pitch * sizeof(float); p.out_row_bytes = out_pitch * sizeof(float);
                p.in_stride = p.out_stride = 4;
                p.mode = mode; p.edge_lead = p.edge_trail = poly_edges; p.arith = arith;
                cudaStream_t st = sge::current_stream();
                return sge::run1d_device(p, st) && cuda_ok(cudaStreamSynchronize(st), "sync");
            }
            (void)cudaGetLastError();
        }
        return sge::run1d_host(f, in, out, rows, len, in_pitch, out_pitch, mode, poly_edges, arith);
    }
    fprintf(stderr, "savgol_b200: input and output must both be device pointers or both be host pointers\n");
    return false;
}

}  // namespace

// ============================================================================================
extern "C" {

SavgolFilter* savgol_create(const SavgolConfig* config)
{
    if (config == nullptr) return nullptr;  // ref: src/savgolFilter.c:641-643 (silent)
    const char* why = nullptr;
    if (!sgc::config1d_valid(config->half_window, config->poly_order, config->derivative, config->time_step, &why)) {
        fprintf(stderr, "savgol: %s (half_window=%d poly_order=%d derivative=%d time_step=%f)\n", why,
                config->half_window, config->poly_order, config->derivative, static_cast<double>(config->time_step));
        return nullptr;
    }
    sge::FilterImpl* fi = static_cast<sge::FilterImpl*>(calloc(1, sizeof(sge::FilterImpl)));
    if (!fi) {
        fprintf(stderr, "savgol: failed to allocate filter context\n");
        return nullptr;
    }
    fi->pub.config = *config;
    fi->pub.window_size = 2 * config->half_window + 1;
    fi->pub.dt_scale = sgc::dt_scale(config->time_step, config->derivative);
    sgc::weights1d(config->half_window, config->poly_order, config->derivative, fi->pub.center_weights,
                   &fi->pub.edge_weights[0][0]);
    fi->magic = sge::kFilterMagic;
    sge::register_filter(fi);
    return &fi->pub;
}

void savgol_destroy(SavgolFilter* filter)
{
    if (!filter) return;
    sge::FilterImpl* fi = sge::live_filter(filter);
    if (fi) {
        sge::unregister_filter(fi);
        for (int d = 0; d < sge::kMaxDevices; ++d)
            if (fi->edge_t[d]) {
                int cur = 0;
                cudaGetDevice(&cur);
                cudaSetDevice(d);
                cudaFree(fi->edge_t[d]);
                cudaSetDevice(cur);
            }
        fi->magic = 0;
    }
    free(filter);
}

int savgol_apply_batch(const SavgolFilter* filter, const float* input, float* output,
                       size_t n_signals, size_t length, size_t in_pitch, size_t out_pitch)
{
    if (filter == nullptr || input == nullptr || output == nullptr) {
        fprintf(stderr, "savgol_apply: NULL pointer\n");
        return -1;
    }
    if (length < static_cast<size_t>(filter->window_size)) {
        fprintf(stderr, "savgol_apply: data length (%lu) < window size (%d)\n", static_cast<unsigned long>(length),
                filter->window_size);
        return -1;
    }
    if (n_signals == 0) return 0;
    if (in_pitch < length || out_pitch < length) {
        if (n_signals > 1) { fprintf(stderr, "savgol_apply_batch: pitch < length\n"); return -1; }
        in_pitch = out_pitch = length;
    }
    const int mode = mode_of(filter);
    return run1d_any(filter, input, output, n_signals, length, in_pitch, out_pitch, mode, mode == sg::MODE_POLY) ? 0 : -1;
}

int savgol_apply(const SavgolFilter* filter, const float* input, float* output, size_t length)
{
    return savgol_apply_batch(filter, input, output, 1, length, length, length);
}

size_t savgol_apply_valid(const SavgolFilter* filter, const float* input, size_t input_length, float* output)
{
    if (filter == nullptr || input == nullptr || output == nullptr) return 0;
    if (input_length < static_cast<size_t>(filter->window_size)) return 0;
    const size_t n = filter->config.half_window;
    const size_t out_len = input_length - 2 * n;
    const MemKind ki = sge::classify(input), ko = sge::classify(output);
    sge::DeviceGuard guard(input);
    if (!sge::device_ready(true)) return 0;
    const int arith = arith_for_batch();
    if (ki == MemKind::Device && ko == MemKind::Device) {
        // VALID == the batch stencil on x+n with the first/last n samples as explicit halos
        sge::Problem1D p{};
        p.filter = filter; p.in = input + n; p.out = output; p.rows = 1; p.len = out_len;
        p.in_row_bytes = p.out_row_bytes = out_len * sizeof(float);
        p.in_stride = p.out_stride = 4;
        p.lhalo = input; p.rhalo = input + (input_length - n);
        p.mode = sg::MODE_POLY; p.arith = arith;
        return sge::run1d_device(p, sge::current_stream()) ? out_len : 0;
    }
    if (ki == MemKind::Device || ko == MemKind::Device) {
        fprintf(stderr, "savgol_b200: input and output must both be device pointers or both be host pointers\n");
        return 0;
    }
    // host: outputs [n, L-n) of the signal through the staging pipeline (every piece has real halos)
    sge::PipeLease lease;
    if (!lease.ok()) return 0;
    return sge::run1d_host_range(*lease, filter, input, input_length, n, input_length - n, output, sg::MODE_POLY, false, arith) ? out_len : 0;
}

int savgol_apply_strided(const SavgolFilter* filter, const void* input, size_t in_stride, size_t in_offset,
                         void* output, size_t out_stride, size_t out_offset, size_t count)
{
    if (filter == nullptr || input == nullptr || output == nullptr) return -1;  // ref: :882-884 (silent)
    if (count < static_cast<size_t>(filter->window_size)) return -1;
    const MemKind ki = sge::classify(input), ko = sge::classify(output);
    sge::DeviceGuard guard(input);
    if (!sge::device_ready(true)) return -1;
    const int arith = arith_for_batch();
    const char* ib = static_cast<const char*>(input) + in_offset;
    char* ob = static_cast<char*>(output) + out_offset;
    if (ki == MemKind::Device && ko == MemKind::Device) {
        sge::Problem1D p{};
        p.filter = filter; p.in = ib; p.out = ob; p.rows = 1; p.len = count;
        p.in_row_bytes = count * in_stride; p.out_row_bytes = count * out_stride;
        p.in_stride = in_stride; p.out_stride = out_stride;
        p.mode = sg::MODE_POLY; p.edge_lead = p.edge_trail = true;  // strided ignores config.boundary (ref :911-928)
        p.arith = arith;
        return sge::run1d_device(p, sge::current_stream()) ? 0 : -1;
    }
    if (ki == MemKind::Device || ko == MemKind::Device) {
        fprintf(stderr, "savgol_b200: input and output must both be device pointers or both be host pointers\n");
        return -1;
    }
    // Host records: ship the input extent as raw bytes, gather on the device, return a contiguous
    // result and scatter it on the host so that only the float fields the reference would write
    // are touched.
    cudaStream_t st = sge::current_stream();
    const size_t in_bytes = (count - 1) * in_stride + sizeof(float);
    char* din = nullptr;
    float* dout = nullptr;
    std::vector<float> tmp(count);
    bool ok = cuda_ok(cudaMallocAsync(&din, in_bytes, st), "cudaMallocAsync") &&
              cuda_ok(cudaMallocAsync(&dout, count * sizeof(float), st), "cudaMallocAsync");
    if (ok) ok = cuda_ok(cudaMemcpyAsync(din, ib, in_bytes, cudaMemcpyHostToDevice, st), "H2D");
    if (ok) {
        sge::Problem1D p{};
        p.filter = filter; p.in = din; p.out = dout; p.rows = 1; p.len = count;
        p.in_row_bytes = in_bytes; p.out_row_bytes = count * sizeof(float);
        p.in_stride = in_stride; p.out_stride = 4;
        p.mode = sg::MODE_POLY; p.edge_lead = p.edge_trail = true; p.arith = arith;
        ok = sge::run1d_device(p, st);
    }
    if (ok) ok = cuda_ok(cudaMemcpyAsync(tmp.data(), dout, count * sizeof(float), cudaMemcpyDeviceToHost, st), "D2H");
    if (ok) ok = cuda_ok(cudaStreamSynchronize(st), "sync");
    if (din) cudaFreeAsync(din, st);
    if (dout) cudaFreeAsync(dout, st);
    if (!ok) return -1;
    for (size_t i = 0; i < count; ++i) *reinterpret_cast<float*>(ob + i * out_stride) = tmp[i];
    return 0;
}

int savgol_apply_halo(const SavgolFilter* filter, const float* input, float* output, size_t length,
                      const float* left_halo, const float* right_halo)
{
    if (filter == nullptr || input == nullptr || output == nullptr) {
        fprintf(stderr, "savgol_apply_halo: NULL pointer\n");
        return -1;
    }
    if (length < static_cast<size_t>(filter->window_size)) {
        fprintf(stderr, "savgol_apply_halo: slice length (%lu) < window size (%d)\n", static_cast<unsigned long>(length),
                filter->window_size);
        return -1;
    }
    const int mode = mode_of(filter);
    if (mode == sg::MODE_PERIODIC && ((left_halo == nullptr) != (right_halo == nullptr))) {
        fprintf(stderr, "savgol_apply_halo: a periodic slice needs both halos (or none)\n");
        return -1;
    }
    sge::DeviceGuard guard(input);
    if (!sge::device_ready(true)) return -1;
    if (sge::classify(input) != MemKind::Device || sge::classify(output) != MemKind::Device ||
        (left_halo && sge::classify(left_halo) != MemKind::Device) ||
        (right_halo && sge::classify(right_halo) != MemKind::Device)) {
        fprintf(stderr, "savgol_apply_halo: slices and halos must be device pointers\n");
        return -1;
    }
    sge::Problem1D p{};
    p.filter = filter; p.in = input; p.out = output; p.rows = 1; p.len = length;
    p.in_row_bytes = p.out_row_bytes = length * sizeof(float);
    p.in_stride = p.out_stride = 4;
    p.lhalo = left_halo; p.rhalo = right_halo;
    p.mode = mode;
    p.edge_lead = mode == sg::MODE_POLY && left_halo == nullptr;
    p.edge_trail = mode == sg::MODE_POLY && right_halo == nullptr;
    p.arith = arith_for_batch();
    return sge::run1d_device(p, sge::current_stream()) ? 0 : -1;
}

}  // extern "C"
