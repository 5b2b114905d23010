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

// ---------------------------------------------------------------------------------------------
// Reference-compatible IN-PLACE semantics.  savgol_apply(f, x, x, L) of the reference is not alias
// safe: its centre loop overwrites samples that later windows still read, so the "in-place" result is a recursive
// filter whose value depends on the loop order (src/savgolFilter.c:763-801).  By default this library returns the
// out-of-place result for aliased calls.  A caller that depends on the reference's actual in-place numbers can ask
// for them: savgol_b200_set_inplace_compat(1) makes exactly-aliased calls run the reference's loop order -- one
// thread per signal, sequential, the reference's 4-chain unfused arithmetic -- bit-identical to the reference.
// The recurrence has no parallelism along a signal; a batch is parallel over its signals.
struct CompatW { float w[sg::kMaxWs]; };

__device__ __forceinline__ float compat_dot4(const float* w, int ws, const float* d, int dstep)
{
    // ref: src/savgolFilter.c:547-580 (dstep = +1) and :593-623 (reverse traversal, dstep = -1)
    float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const int rem = ws & 3;
    for (int k = 0; k < ws; ++k) {
        const int c = k < rem ? k : ((k - rem) & 3);
        s[c] = __fadd_rn(s[c], __fmul_rn(w[k], d[static_cast<long long>(k) * dstep]));
    }
    return __fadd_rn(__fadd_rn(s[0], s[1]), __fadd_rn(s[2], s[3]));
}

__global__ void compat_inplace_kernel(float* data, size_t rows, size_t len, size_t pitch, const CompatW W, const float* __restrict__ edge_t,
                                      int n, int mode, float scale)
{
    const size_t r = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (r >= rows) return;
    float* x = data + r * pitch;
    const int ws = 2 * n + 1;
    const long long L = static_cast<long long>(len);
    for (long long j = n; j < L - n; ++j) x[j] = __fmul_rn(compat_dot4(W.w, ws, x + (j - n), 1), scale);
    float win[sg::kMaxWs], ew[sg::kMaxWs];
    if (mode == sg::MODE_POLY) {
        for (int i = 0; i < n; ++i) {   // leading edge, reversed traversal from x[2n]
            for (int k = 0; k < ws; ++k) ew[k] = edge_t[k * 32 + i];
            x[i] = __fmul_rn(compat_dot4(ew, ws, x + (ws - 1), -1), scale);
        }
        for (int i = 0; i < n; ++i) {   // trailing edge
            for (int k = 0; k < ws; ++k) ew[k] = edge_t[k * 32 + i];
            x[L - 1 - i] = __fmul_rn(compat_dot4(ew, ws, x + (L - ws), 1), scale);
        }
    } else {
        // ref: convolve_padded / get_padded_sample, src/savgolFilter.c:442-535 (64-bit indices here)
        auto padded = [&](long long idx) -> float {
            if (idx >= 0 && idx < L) return x[idx];
            if (mode == sg::MODE_REFLECT) {
                if (idx < 0) { idx = -idx - 1; if (idx >= L) idx = L - 1; }
                else { idx = 2 * L - idx - 1; if (idx < 0) idx = 0; }
                return x[idx];
            }
            if (mode == sg::MODE_PERIODIC) return x[((idx % L) + L) % L];
            return idx < 0 ? x[0] : x[L - 1];
        };
        for (long long i = 0; i < n; ++i) {
            for (int k = 0; k < ws; ++k) win[k] = padded(i - n + k);
            x[i] = __fmul_rn(compat_dot4(W.w, ws, win, 1), scale);
        }
        for (long long i = L - n; i < L; ++i) {
            for (int k = 0; k < ws; ++k) win[k] = padded(i - n + k);
            x[i] = __fmul_rn(compat_dot4(W.w, ws, win, 1), scale);
        }
    }
}

thread_local int t_inplace_compat = 0;

bool run_compat_inplace(const SavgolFilter* f, float* data, size_t rows, size_t len, size_t pitch, int mode)
{
    const MemKind k = sge::classify(data);
    sge::DeviceGuard guard(data);
    if (!sge::device_ready(true)) return false;
    cudaStream_t st = sge::current_stream();
    float* d = data;
    size_t dpitch = pitch;
    bool ok = true;
    if (k != MemKind::Device) {
        dpitch = len;
        ok = cuda_ok(cudaMallocAsync(&d, rows * len * sizeof(float), st), "cudaMallocAsync") &&
             cuda_ok(cudaMemcpy2DAsync(d, len * sizeof(float), data, pitch * sizeof(float), len * sizeof(float), rows, cudaMemcpyHostToDevice, st), "H2D");
    }
    float* temp = nullptr;
    const float* et = ok ? sge::edge_table_device(f, st, &temp) : nullptr;
    ok = ok && et;
    if (ok) {
        CompatW W;
        std::memcpy(W.w, f->center_weights, sizeof(W.w));
        const float scale = f->dt_scale != 0.0f ? 1.0f / f->dt_scale : 1.0f;
        const unsigned blocks = static_cast<unsigned>((rows + 63) / 64);
        compat_inplace_kernel<<<blocks, 64, 0, st>>>(d, rows, len, dpitch, W, et, f->config.half_window, mode, scale);
        ok = cuda_ok(cudaGetLastError(), "compat in-place launch");
    }
    if (ok && k != MemKind::Device)
        ok = cuda_ok(cudaMemcpy2DAsync(data, pitch * sizeof(float), d, len * sizeof(float), len * sizeof(float), rows, cudaMemcpyDeviceToHost, st), "D2H") &&
             cuda_ok(cudaStreamSynchronize(st), "sync");
    if (temp) cudaFreeAsync(temp, st);
    if (k != MemKind::Device && d) cudaFreeAsync(d, st);
    return ok;
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
    if (t_inplace_compat && input == output && in_pitch == out_pitch)
        return run_compat_inplace(filter, output, n_signals, length, out_pitch, mode) ? 0 : -1;
    return run1d_any(filter, input, output, n_signals, length, in_pitch, out_pitch, mode, mode == sg::MODE_POLY) ? 0 : -1;
}

void savgol_b200_set_inplace_compat(int on) { t_inplace_compat = on ? 1 : 0; }
int savgol_b200_get_inplace_compat(void) { return t_inplace_compat; }

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
