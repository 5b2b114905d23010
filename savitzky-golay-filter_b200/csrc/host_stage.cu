// host_stage.cu -- staging of HOST buffers through the GPU: pipeline pool, device guard, and the 1D host paths.
//
// The reference works on host arrays (src/savgolFilter.c:743-850); a drop-in replacement therefore has to take
// host pointers.  They are streamed through a ring of three device slots per call -- H2D copy, kernel and D2H
// copy of consecutive chunks overlap on three streams -- so a large host batch runs at the speed of the PCIe
// link.  Every call leases its own pipeline (per device), so host threads and devices never serialise on
// anything but the link.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "engine.h"

namespace sge {

// ---------------------------------------------------------------------------------------------
bool Pipeline::ensure(size_t need_in, size_t need_out)
{
    if (!s_in) {
        if (!cuda_ok(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking), "stream")) return false;
        if (!cuda_ok(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking), "stream")) return false;
        if (!cuda_ok(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking), "stream")) return false;
        for (int i = 0; i < kSlots; ++i) {
            cudaEventCreateWithFlags(&e_in[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&e_k[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&e_out[i], cudaEventDisableTiming);
        }
    }
    if (need_in > cap_in) {
        for (int i = 0; i < kSlots; ++i) { cudaFree(d_in[i]); d_in[i] = nullptr; }
        cap_in = 0;
        for (int i = 0; i < kSlots; ++i)
            if (!cuda_ok(cudaMalloc(&d_in[i], need_in * sizeof(float)), "cudaMalloc(staging in)")) return false;
        cap_in = need_in;
    }
    if (need_out > cap_out) {
        for (int i = 0; i < kSlots; ++i) { cudaFree(d_out[i]); d_out[i] = nullptr; }
        cap_out = 0;
        for (int i = 0; i < kSlots; ++i)
            if (!cuda_ok(cudaMalloc(&d_out[i], need_out * sizeof(float)), "cudaMalloc(staging out)")) return false;
        cap_out = need_out;
    }
    return true;
}

void Pipeline::release()
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

namespace {
std::mutex g_pool_mu;
std::vector<Pipeline*> g_idle[kMaxDevices];
constexpr size_t kKeepIdle = 4;   // idle pipelines kept per device (each holds up to 6 staging chunks of device memory)
}  // namespace

PipeLease::PipeLease()
{
    int dev = 0;
    if (!cuda_ok(cudaGetDevice(&dev), "cudaGetDevice") || dev < 0 || dev >= kMaxDevices) return;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_idle[dev].empty()) {
            p_ = g_idle[dev].back();
            g_idle[dev].pop_back();
        }
    }
    if (!p_) {
        p_ = new Pipeline();
        p_->dev = dev;
    }
}

PipeLease::~PipeLease()
{
    if (!p_) return;
    bool keep = false;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_idle[p_->dev].size() < kKeepIdle) {
            g_idle[p_->dev].push_back(p_);
            keep = true;
        }
    }
    if (!keep) {
        DeviceGuard g(p_->dev);
        p_->release();
        delete p_;
    }
}

DeviceGuard::DeviceGuard(const void* ptr)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { (void)cudaGetLastError(); return; }
    if (at.type != cudaMemoryTypeDevice) return;   // host and managed memory: whatever device is current
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess || cur == at.device) return;
    if (cudaSetDevice(at.device) == cudaSuccess) prev_ = cur;
}
DeviceGuard::DeviceGuard(int device)
{
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess || cur == device) return;
    if (cudaSetDevice(device) == cudaSuccess) prev_ = cur;
}
DeviceGuard::~DeviceGuard()
{
    if (prev_ >= 0) cudaSetDevice(prev_);
}

size_t chunk_floats()
{
    // samples per staged chunk: small enough that pipeline fill + drain (one chunk each way) is a few
    // percent of a large transfer, large enough to stay near PCIe peak.  Env override for experiments / tests.
    static size_t v = [] {
        const char* e = getenv("SAVGOL_B200_CHUNK_MIB");
        size_t mib = e ? static_cast<size_t>(atoi(e)) : 64;  // measured on B200: 64 MiB 23.0 ms/GiB-step, 16 MiB 23.2, 4 MiB 26.9
        if (mib < 1) mib = 1;
        return mib << 18;
    }();
    return v;
}

static bool overlap(const void* a, size_t abytes, const void* b, size_t bbytes)
{
    const char* a0 = static_cast<const char*>(a);
    const char* b0 = static_cast<const char*>(b);
    return a0 < b0 + bbytes && b0 < a0 + abytes;
}

// ---------------------------------------------------------------------------------------------
bool run1d_host_range(Pipeline& P, const SavgolFilter* f, const float* x, size_t L, size_t a, size_t b, float* y,
                      int mode, bool poly_edges, int arith)
{
    const size_t n = f->config.half_window;
    const size_t ws = 2 * n + 1;
    const size_t padl = (n + 3) & ~static_cast<size_t>(3);
    const size_t piece = chunk_floats();
    if (b <= a) return true;
    if ((a != 0 && a < n) || (b != L && b + n > L) || b > L) {   // a cut must leave n real samples on its far side (or be a true end)
        fprintf(stderr, "savgol_b200: internal: host range [%lu, %lu) of %lu samples is closer than half_window to an end\n",
                static_cast<unsigned long>(a), static_cast<unsigned long>(b), static_cast<unsigned long>(L));
        return false;
    }
    // slot layout: [padl-n slack | n left halo | piece (+ up to one window) | n right halo]
    if (!P.ensure(padl + piece + 2 * sg::kMaxWs + n, piece + sg::kMaxWs)) return false;

    // Aliasing.  The D2H copy of one piece overwrites host samples that the H2D copy of the NEXT piece still
    // needs as its left halo (and, for a periodic signal, the head that the LAST piece needs as its right halo);
    // the copies run on different streams.  Exactly in place (y == x + a): those few samples are snapshotted
    // before anything is written.  Any other overlap: the input is copied aside first.
    const bool inplace = (y == x + a);
    std::vector<float> aside;
    if (!inplace && overlap(x, L * sizeof(float), y, (b - a) * sizeof(float))) {
        aside.assign(x, x + L);
        x = aside.data();
    }
    std::vector<float> snaps;          // per piece: its n-sample left halo (in-place calls only)
    std::vector<float> head;           // x[0..n) for the periodic wrap of the last piece
    std::vector<size_t> cuts;          // piece boundaries
    for (size_t s0 = a; s0 < b;) {
        size_t s1 = std::min(b, s0 + piece);
        if (b - s1 < ws) s1 = b;       // keep the last piece >= one window
        cuts.push_back(s0);
        s0 = s1;
    }
    cuts.push_back(b);
    const size_t npieces = cuts.size() - 1;
    if (inplace) {
        snaps.resize(npieces * n);
        for (size_t k = 0; k < npieces; ++k)
            if (cuts[k] >= n) std::memcpy(&snaps[k * n], x + (cuts[k] - n), n * sizeof(float));
        if (mode == sg::MODE_PERIODIC) head.assign(x, x + n);
    }

    for (size_t k = 0; k < npieces; ++k) {
        const size_t s0 = cuts[k], s1 = cuts[k + 1], plen = s1 - s0;
        const int s = static_cast<int>(k % Pipeline::kSlots);
        if (k >= Pipeline::kSlots && !cuda_ok(cudaStreamWaitEvent(P.s_in, P.e_out[s], 0), "wait")) return false;
        float* dx = P.d_in[s] + padl;
        const bool has_l = s0 >= n && s0 > 0, has_r = s1 + n <= L && s1 < L;   // n real samples exist on that side
        // body plus whatever neighbouring samples exist
        const size_t c0 = (has_l && !inplace) ? s0 - n : s0, c1 = has_r ? s1 + n : s1;
        if (!cuda_ok(cudaMemcpyAsync(dx - (s0 - c0), x + c0, (c1 - c0) * sizeof(float), cudaMemcpyHostToDevice, P.s_in), "H2D")) return false;
        if (has_l && inplace &&
            !cuda_ok(cudaMemcpyAsync(dx - n, &snaps[k * n], n * sizeof(float), cudaMemcpyHostToDevice, P.s_in), "H2D halo")) return false;
        const float* lh = has_l ? dx - n : nullptr;
        const float* rh = has_r ? dx + plen : nullptr;
        if (mode == sg::MODE_PERIODIC && !(s0 == 0 && s1 == L)) {
            // true ends of a periodic signal wrap around: fetch the far end as an explicit halo
            if (s0 == 0) { cudaMemcpyAsync(dx - n, x + (L - n), n * sizeof(float), cudaMemcpyHostToDevice, P.s_in); lh = dx - n; }
            if (s1 == L) { cudaMemcpyAsync(dx + plen, inplace ? head.data() : x, n * sizeof(float), cudaMemcpyHostToDevice, P.s_in); rh = dx + plen; }
        }
        cudaEventRecord(P.e_in[s], P.s_in);
        cudaStreamWaitEvent(P.s_k, P.e_in[s], 0);
        if (k >= Pipeline::kSlots) cudaStreamWaitEvent(P.s_k, P.e_out[s], 0);
        Problem1D p{};
        p.filter = f; p.in = dx; p.out = P.d_out[s];
        p.rows = 1; p.len = plen;
        p.in_row_bytes = p.out_row_bytes = plen * sizeof(float);
        p.in_stride = p.out_stride = 4;
        p.lhalo = lh; p.rhalo = rh;
        p.mode = mode;
        p.edge_lead = poly_edges && s0 == 0; p.edge_trail = poly_edges && s1 == L;
        p.arith = arith;
        if (!run1d_device(p, P.s_k)) return false;
        cudaEventRecord(P.e_k[s], P.s_k);
        cudaStreamWaitEvent(P.s_out, P.e_k[s], 0);
        if (!cuda_ok(cudaMemcpyAsync(y + (s0 - a), P.d_out[s], plen * sizeof(float), cudaMemcpyDeviceToHost, P.s_out), "D2H")) return false;
        cudaEventRecord(P.e_out[s], P.s_out);
    }
    // `snaps`, `head`, `aside` are pageable: their H2D copies were staged before cudaMemcpyAsync returned
    return cuda_ok(cudaStreamSynchronize(P.s_out), "sync") && cuda_ok(cudaStreamSynchronize(P.s_k), "sync") &&
           cuda_ok(cudaStreamSynchronize(P.s_in), "sync");
}

bool run1d_host(const SavgolFilter* f, const float* in, float* out, size_t rows, size_t len, size_t in_pitch, size_t out_pitch,
                int mode, bool poly_edges, int arith)
{
    PipeLease lease;
    if (!lease.ok()) return false;
    Pipeline& P = *lease;
    const size_t chunk = chunk_floats();

    if (len > chunk) {
        for (size_t r = 0; r < rows; ++r)
            if (!run1d_host_range(P, f, in + r * in_pitch, len, 0, len, out + r * out_pitch, mode, poly_edges, arith)) return false;
        return true;
    }

    // ---- chunks of whole rows ----
    // rows are independent, and a chunk's D2H only overwrites rows whose H2D has completed -- in place is safe
    // when every output row coincides with its input row; any other overlap goes through a copy of the input
    std::vector<float> aside;
    const size_t in_span = (rows - 1) * in_pitch + len, out_span = (rows - 1) * out_pitch + len;
    if (overlap(in, in_span * sizeof(float), out, out_span * sizeof(float)) && !(in == out && in_pitch == out_pitch)) {
        aside.assign(in, in + in_span);
        in = aside.data();
    }
    const size_t rows_per = std::max<size_t>(1, std::min(rows, chunk / len));
    if (!P.ensure(rows_per * len, rows_per * len)) return false;
    size_t done = 0;
    for (size_t c = 0; done < rows; ++c, done += rows_per) {
        const int s = static_cast<int>(c % Pipeline::kSlots);
        const size_t nr = std::min(rows_per, rows - done);
        if (c >= Pipeline::kSlots) {
            // slot reuse: its previous D2H must have drained before we overwrite d_out/d_in
            if (!cuda_ok(cudaStreamWaitEvent(P.s_in, P.e_out[s], 0), "wait")) return false;
        }
        if (!cuda_ok(cudaMemcpy2DAsync(P.d_in[s], len * sizeof(float), in + done * in_pitch, in_pitch * sizeof(float),
                                       len * sizeof(float), nr, cudaMemcpyHostToDevice, P.s_in), "H2D")) return false;
        cudaEventRecord(P.e_in[s], P.s_in);
        cudaStreamWaitEvent(P.s_k, P.e_in[s], 0);
        if (c >= Pipeline::kSlots) cudaStreamWaitEvent(P.s_k, P.e_out[s], 0);
        Problem1D p{};
        p.filter = f; p.in = P.d_in[s]; p.out = P.d_out[s];
        p.rows = nr; p.len = len;
        p.in_row_bytes = p.out_row_bytes = len * sizeof(float);
        p.in_stride = p.out_stride = 4;
        p.mode = mode; p.edge_lead = p.edge_trail = poly_edges; p.arith = arith;
        if (!run1d_device(p, P.s_k)) return false;
        cudaEventRecord(P.e_k[s], P.s_k);
        cudaStreamWaitEvent(P.s_out, P.e_k[s], 0);
        if (!cuda_ok(cudaMemcpy2DAsync(out + done * out_pitch, out_pitch * sizeof(float), P.d_out[s], len * sizeof(float),
                                       len * sizeof(float), nr, cudaMemcpyDeviceToHost, P.s_out), "D2H")) return false;
        cudaEventRecord(P.e_out[s], P.s_out);
    }
    return cuda_ok(cudaStreamSynchronize(P.s_out), "sync") && cuda_ok(cudaStreamSynchronize(P.s_k), "sync") &&
           cuda_ok(cudaStreamSynchronize(P.s_in), "sync");
}

}  // namespace sge
