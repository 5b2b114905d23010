// engine.cu -- runtime glue: filter registry, pointer classification, launch of the 1D problem.
#include "engine.h"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_set>

#include "sg1d_launch.h"

namespace sg { extern std::atomic<unsigned long long> g_launches, g_tma_launches; extern std::atomic<int> g_tma_enabled; }

namespace sge {

namespace {
std::mutex g_reg_mu;
std::unordered_set<const void*> g_live;
thread_local cudaStream_t t_stream = nullptr;
std::atomic<int> g_exact{0};              // process default of the arithmetic flavour
thread_local int t_exact = -1;            // the calling thread's own choice (-1: follow the process default)
std::atomic<int> g_dev_state[kMaxDevices + 1];  // per device: 0 unknown, 1 ok, -1 unusable (last entry: "no device at all")
}  // namespace

void register_filter(FilterImpl* f)
{
    std::lock_guard<std::mutex> lk(g_reg_mu);
    g_live.insert(f);
}
void unregister_filter(FilterImpl* f)
{
    std::lock_guard<std::mutex> lk(g_reg_mu);
    g_live.erase(f);
}
FilterImpl* live_filter(const SavgolFilter* f)
{
    std::lock_guard<std::mutex> lk(g_reg_mu);
    if (g_live.count(f) == 0) return nullptr;
    FilterImpl* fi = reinterpret_cast<FilterImpl*>(const_cast<SavgolFilter*>(f));
    return fi->magic == kFilterMagic ? fi : nullptr;
}

bool cuda_ok(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return true;
    fprintf(stderr, "savgol_b200: %s: %s\n", what, cudaGetErrorString(e));
    return false;
}

bool device_ready(bool complain)
{
    int dev = -1;
    int st = -1;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevices) {
        st = g_dev_state[dev].load();
        if (st == 0) {
            int major = 0;
            const cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
            st = (e == cudaSuccess && major == 10) ? 1 : -1;   // the fatbin holds sm_100a code only
            g_dev_state[dev].store(st);
        }
    }
    (void)cudaGetLastError();
    if (st != 1 && complain)
        fprintf(stderr, "savgol_b200: no usable CUDA device (need compute capability 10.x); there is no CPU fallback\n");
    return st == 1;
}

MemKind classify(const void* p)
{
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return MemKind::Pageable; }
    switch (at.type) {
        case cudaMemoryTypeDevice:
        case cudaMemoryTypeManaged: return MemKind::Device;
        case cudaMemoryTypeHost: return MemKind::Pinned;
        default: return MemKind::Pageable;
    }
}

cudaStream_t current_stream() { return t_stream; }
int exact_mode() { return t_exact >= 0 ? t_exact : g_exact.load(std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------------
const float* edge_table_device(const SavgolFilter* f, cudaStream_t stream, float** temp)
{
    *temp = nullptr;
    int dev = 0;
    if (!cuda_ok(cudaGetDevice(&dev), "cudaGetDevice")) return nullptr;
    FilterImpl* fi = live_filter(f);
    if (fi && dev < kMaxDevices && fi->edge_t[dev]) return fi->edge_t[dev];

    // transposed table: t[k*32 + e] = E[e][k]  (lanes = edge positions read consecutive floats)
    float host[sg::kMaxWs * 32];
    for (int k = 0; k < sg::kMaxWs; ++k)
        for (int e = 0; e < 32; ++e) host[k * 32 + e] = f->edge_weights[e][k];
    float* d = nullptr;
    if (fi && dev < kMaxDevices) {
        std::lock_guard<std::mutex> lk(g_reg_mu);
        if (fi->edge_t[dev]) return fi->edge_t[dev];
        if (!cuda_ok(cudaMalloc(&d, sizeof(host)), "cudaMalloc(edge table)")) return nullptr;
        // synchronous copy: the table must be complete before any stream uses it
        if (!cuda_ok(cudaMemcpy(d, host, sizeof(host), cudaMemcpyHostToDevice), "upload edge table")) {
            cudaFree(d);
            return nullptr;
        }
        fi->edge_t[dev] = d;
        return d;
    }
    if (!cuda_ok(cudaMallocAsync(&d, sizeof(host), stream), "cudaMallocAsync(edge table)")) return nullptr;
    if (!cuda_ok(cudaMemcpyAsync(d, host, sizeof(host), cudaMemcpyHostToDevice, stream), "upload edge table")) {
        cudaFreeAsync(d, stream);
        return nullptr;
    }
    cudaStreamSynchronize(stream);  // `host` is a stack array
    *temp = d;
    return d;
}

// contiguous scratch -> strided destination (only used for aliased strided calls)
__global__ void scatter_kernel(const float* __restrict__ src, char* dst, size_t rows, size_t len,
                               size_t row_bytes, size_t stride)
{
    const size_t total = rows * len;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = i / len, c = i - r * len;
        *reinterpret_cast<float*>(dst + r * row_bytes + c * stride) = src[i];
    }
}

static bool ranges_overlap(const void* a, size_t abytes, const void* b, size_t bbytes)
{
    const char* a0 = static_cast<const char*>(a);
    const char* b0 = static_cast<const char*>(b);
    return a0 < b0 + bbytes && b0 < a0 + abytes;
}

bool run1d_device(const Problem1D& p, cudaStream_t stream)
{
    const SavgolFilter* f = p.filter;
    const int n = f->config.half_window;

    sg::W1D w;
    std::memset(&w, 0, sizeof(w));
    std::memcpy(w.w, f->center_weights, sizeof(float) * static_cast<size_t>(2 * n + 1));
    {
        // FAST flavour: weights pre-scaled by 1/dt^d and paired (see sg_common.cuh)
        const float sc = (f->dt_scale != 0.0f) ? (1.0f / f->dt_scale) : 1.0f;
        const int ws = 2 * n + 1;
        w.ws_first = f->center_weights[0] * sc;
        w.ws_last = f->center_weights[ws - 1] * sc;
        for (int k = 1; k < ws; ++k) w.pw[k] = make_float2(f->center_weights[k] * sc, f->center_weights[k - 1] * sc);
    }

    sg::Args1D a;
    std::memset(&a, 0, sizeof(a));
    a.in = static_cast<const char*>(p.in);
    a.out = static_cast<char*>(p.out);
    a.rows = static_cast<long long>(p.rows);
    a.len = static_cast<long long>(p.len);
    a.out_len = static_cast<long long>(p.out_len ? p.out_len : p.len);
    a.in_row_bytes = static_cast<long long>(p.in_row_bytes);
    a.out_row_bytes = static_cast<long long>(p.out_row_bytes);
    a.in_stride = static_cast<long long>(p.in_stride);
    a.out_stride = static_cast<long long>(p.out_stride);
    a.lhalo = p.lhalo; a.rhalo = p.rhalo;
    a.lhalo_pitch = static_cast<long long>(p.lhalo_pitch);
    a.rhalo_pitch = static_cast<long long>(p.rhalo_pitch);
    a.state_out = p.state_out; a.state_pitch = static_cast<long long>(p.state_pitch); a.state_w = p.state_w;
    a.scale = (f->dt_scale != 0.0f) ? (1.0f / f->dt_scale) : 1.0f;  // ref: src/savgolFilter.c:759
    a.mode = p.mode;
    a.edge_lead = p.edge_lead ? 1 : 0;
    a.edge_trail = p.edge_trail ? 1 : 0;

    float* temp_edges = nullptr;
    if (p.edge_lead || p.edge_trail) {
        a.edge_t = edge_table_device(f, stream, &temp_edges);
        if (!a.edge_t) return false;
    }

    int variant;
    if (p.stream_history) variant = (p.arith == sg::ARITH_FAST) ? sg::V_STREAM_FAST : sg::V_STREAM_EXACTSEQ;
    else variant = p.arith == sg::ARITH_FAST ? sg::V_BATCH_FAST
                 : p.arith == sg::ARITH_EXACT4 ? sg::V_BATCH_EXACT4 : sg::V_BATCH_EXACTSEQ;

    // Aliasing.  A signal that fits one tile is loaded completely before its CTA stores anything, so
    // in == out is safe there.  Longer signals would race between CTAs (one tile's halo is another
    // tile's output), so an aliased call goes through a scratch output and is copied back: the
    // result is the out-of-place result (DESIGN.md "in-place").
    const size_t in_bytes = p.rows ? (p.rows - 1) * p.in_row_bytes + p.len * p.in_stride : 0;
    const size_t out_bytes = p.rows ? (p.rows - 1) * p.out_row_bytes + p.len * p.out_stride : 0;
    const bool alias = ranges_overlap(p.in, in_bytes, p.out, out_bytes);
    const bool same_layout = p.in == p.out && p.in_row_bytes == p.out_row_bytes && p.in_stride == p.out_stride;
    bool ok = true;
    if (alias && !(same_layout && p.len <= static_cast<size_t>(sg::kTile))) {
        float* scratch = nullptr;
        const size_t sbytes = p.rows * p.len * sizeof(float);
        if (!cuda_ok(cudaMallocAsync(&scratch, sbytes, stream), "cudaMallocAsync(in-place scratch)")) ok = false;
        if (ok) {
            a.out = reinterpret_cast<char*>(scratch);
            a.out_row_bytes = static_cast<long long>(p.len * sizeof(float));
            a.out_stride = 4;
            ok = cuda_ok(sg::sg1d_launch(n, variant, w, a, stream), "sg1d launch");
            if (ok && p.out_stride == 4) {
                ok = cuda_ok(cudaMemcpy2DAsync(p.out, p.out_row_bytes, scratch, p.len * sizeof(float),
                                               p.len * sizeof(float), p.rows, cudaMemcpyDeviceToDevice, stream),
                             "in-place copy back");
            } else if (ok) {
                const size_t total = p.rows * p.len;
                const unsigned blocks = static_cast<unsigned>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
                scatter_kernel<<<blocks, 256, 0, stream>>>(scratch, static_cast<char*>(p.out), p.rows, p.len,
                                                          p.out_row_bytes, p.out_stride);
                sg::g_launches.fetch_add(1, std::memory_order_relaxed);
                ok = cuda_ok(cudaGetLastError(), "scatter launch");
            }
            cudaFreeAsync(scratch, stream);
        }
    } else {
        ok = cuda_ok(sg::sg1d_launch(n, variant, w, a, stream), "sg1d launch");
    }
    if (temp_edges) cudaFreeAsync(temp_edges, stream);
    return ok;
}

}  // namespace sge

// ---------------------------------------------------------------------------------------------
// library / device control (C ABI, include/savgol_b200.h part 2)
extern "C" {

int savgol_b200_version(void) { return 100; }
int savgol_b200_device_ok(void) { return sge::device_ready(false) ? 1 : 0; }
void savgol_b200_set_stream(void* s) { sge::t_stream = static_cast<cudaStream_t>(s); }
void* savgol_b200_get_stream(void) { return sge::t_stream; }
unsigned long long savgol_b200_launch_count(void) { return sg::g_launches.load(); }
unsigned long long savgol_b200_tma_launch_count(void) { return sg::g_tma_launches.load(); }
void savgol_b200_set_tma(int how) { sg::g_tma_enabled.store(how < 0 ? 0 : how > 2 ? 2 : how); }
void savgol_b200_set_exact(int exact) { sge::t_exact = exact ? 1 : 0; }
void savgol_b200_set_exact_default(int exact) { sge::g_exact.store(exact ? 1 : 0); }
int savgol_b200_get_exact(void) { return sge::exact_mode(); }
size_t savgol_b200_staging_chunk(size_t total_floats, int pageable) { return sge::staging_chunk(total_floats, pageable != 0); }
int savgol_b200_host_copy2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width, size_t rows)
{
    if ((!dst || !src) && width && rows) return -1;
    sge::host_copy2d(dst, dst_pitch, src, src_pitch, width, rows);
    return sge::copy_threads();
}

}  // extern "C"
