// host_stage.cu -- staging of HOST buffers through the GPU: pipeline pool, device guard, and the 1D host paths.
//
// The reference works on host arrays (src/savgolFilter.c:743-850); a drop-in replacement therefore has to take
// host pointers.  They are streamed through a ring of three device slots per call -- H2D copy, kernel and D2H
// copy of consecutive chunks overlap on three streams -- so a large host batch runs at the speed of the PCIe
// link.  Every call leases its own pipeline (per device), so host threads and devices never serialise on
// anything but the link.  Pageable memory is not DMA-able: it crosses pinned bounce buffers, copied by a small
// pool of host threads while the neighbouring chunks are on the link (measured on the B200 box, 1 GiB each way:
// driver staging 7 GB/s, see DESIGN.md section 7 for the bounce numbers).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "engine.h"

namespace sge {

// ---------------------------------------------------------------------------------------------
// Host copy pool: the caller plus copy_threads()-1 workers pull ~1 MiB tasks off one shared counter.
namespace {
constexpr size_t kCopyTask = size_t(1) << 20;
constexpr size_t kCopyInline = size_t(256) << 10;   // below this a single memcpy is faster than waking anybody

// memcpy whose stores bypass the cache: the destination is either a bounce buffer the DMA engine reads next or
// user memory nobody touches before the call returns, so write-allocate traffic (a read of every destination
// line) would only take host memory bandwidth away from the copy itself.
inline void copy_stream(char* dst, const char* src, size_t n)
{
#if defined(__SSE2__)
    static const bool nt = [] { const char* e = getenv("SAVGOL_B200_COPY_NT"); return !e || atoi(e) != 0; }();
    if (nt && n >= 4096) {
        const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
        if (head) { std::memcpy(dst, src, head); dst += head; src += head; n -= head; }
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
        }
        _mm_sfence();
        if (i < n) std::memcpy(dst + i, src + i, n - i);
        return;
    }
#endif
    std::memcpy(dst, src, n);
}

class CopyPool {
  public:
    static CopyPool& get()
    {
        static CopyPool* p = new CopyPool();   // leaked on purpose: workers may outlive static destruction
        return *p;
    }
    int threads() const { return nworkers_ + 1; }
    void copy2d(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, size_t rows)
    {
        if (width == 0 || rows == 0) return;
        if (dpitch == width && spitch == width) { width *= rows; rows = 1; }
        if (nworkers_ == 0 || width * rows <= kCopyInline) {
            for (size_t r = 0; r < rows; ++r) std::memcpy(dst + r * dpitch, src + r * spitch, width);
            return;
        }
        std::lock_guard<std::mutex> job(job_mu_);
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = dst; src_ = src; dpitch_ = dpitch; spitch_ = spitch; width_ = width; rows_ = rows;
            if (width >= kCopyTask) {
                segs_ = (width + kCopyTask - 1) / kCopyTask; rows_per_ = 1; ntasks_ = rows * segs_;
            } else {
                segs_ = 1; rows_per_ = kCopyTask / width; ntasks_ = (rows + rows_per_ - 1) / rows_per_;
            }
            next_.store(0, std::memory_order_relaxed);
            open_ = true;
            ++gen_;
        }
        cv_.notify_all();
        run_tasks();
        std::unique_lock<std::mutex> lk(mu_);
        open_ = false;                                    // late wakers skip this job
        done_cv_.wait(lk, [&] { return active_ == 0; });  // every claimed task has been completed
    }

  private:
    CopyPool()
    {
        int want = 0;
        if (const char* e = getenv("SAVGOL_B200_COPY_THREADS")) want = atoi(e);
        if (want <= 0) {
            // measured on the 16-vCPU B200 box (1 GiB each way): 4 threads 3.5, 8: 4.3, 12: 5.0, 16: 4.6 Gsamples/s
            const unsigned hw = std::thread::hardware_concurrency();
            want = static_cast<int>(std::min(12u, std::max(1u, hw - hw / 4)));
        }
        nworkers_ = std::min(want, 64) - 1;
        for (int i = 0; i < nworkers_; ++i) std::thread([this] { worker(); }).detach();
    }
    void run_tasks()
    {
        for (;;) {
            const size_t t = next_.fetch_add(1, std::memory_order_relaxed);
            if (t >= ntasks_) return;
            if (segs_ > 1) {
                const size_t r = t / segs_, off = (t % segs_) * kCopyTask;
                copy_stream(dst_ + r * dpitch_ + off, src_ + r * spitch_ + off, std::min(kCopyTask, width_ - off));
            } else {
                const size_t r1 = std::min(rows_, (t + 1) * rows_per_);
                for (size_t r = t * rows_per_; r < r1; ++r) copy_stream(dst_ + r * dpitch_, src_ + r * spitch_, width_);
            }
        }
    }
    void worker()
    {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            if (!open_) continue;
            ++active_;
            lk.unlock();
            run_tasks();
            lk.lock();
            if (--active_ == 0) done_cv_.notify_all();
        }
    }

    std::mutex job_mu_;   // one job at a time; concurrent callers queue here
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    unsigned long long gen_ = 0;
    bool open_ = false;
    int active_ = 0, nworkers_ = 0;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t dpitch_ = 0, spitch_ = 0, width_ = 0, rows_ = 0, segs_ = 1, rows_per_ = 1, ntasks_ = 0;
    std::atomic<size_t> next_{0};
};
}  // namespace

void host_copy2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width, size_t rows)
{
    CopyPool::get().copy2d(static_cast<char*>(dst), dst_pitch, static_cast<const char*>(src), src_pitch, width, rows);
}
int copy_threads() { return CopyPool::get().threads(); }

// ---------------------------------------------------------------------------------------------
size_t Pipeline::begin(const void* in, const void* out, size_t total)
{
    static const bool off = [] { const char* e = getenv("SAVGOL_B200_NO_BOUNCE"); return e && atoi(e) != 0; }();
    bounce_in = !off && in && classify(in) == MemKind::Pageable;
    bounce_out = !off && out && classify(out) == MemKind::Pageable;
    for (Pending& p : pend) p.dst = nullptr;
    return staging_chunk(total, bounce_in || bounce_out);
}

// Staging chunk (floats) of a host-pointer call of `total` floats; pure host logic (savgol_b200_staging_chunk).
size_t staging_chunk(size_t total, bool bounce)
{
    size_t chunk = chunk_floats(bounce);
    static const bool fixed = [] { const char* e = getenv("SAVGOL_B200_FIXED_CHUNK"); return e && e[0] == '1'; }();
    if (total >= (size_t(8) << 18) && !fixed) {
        // measured on B200 (pinned, tools/r2_host_small.py): 64 MiB each way as one chunk 2.44 ms (H2D, kernel and D2H in
        // sequence), as eight chunks 1.68 ms; 16 MiB: 0.65 -> 0.54 ms.  Below ~8 MiB the per-chunk launch and event
        // latencies outweigh the overlap (4 MB: 0.25 ms as one chunk, 0.37 ms as two).
        const size_t floor_ = std::min<size_t>(chunk, size_t(2) << 18);   // 2 MiB: below that the link efficiency drops
        // (through bounce buffers every chunk also costs a wake-up of the host copy pool: four chunks, not eight --
        // 64 MiB pageable: 2.8 ms in 16 MiB chunks, 3.4 ms in 8 MiB chunks)
        chunk = std::min(chunk, std::max(floor_, total / (bounce ? 4 : 8)));
    }
    return chunk;
}

bool Pipeline::reuse(int s)
{
    if (pend[s].dst) {
        if (!cuda_ok(cudaEventSynchronize(e_out[s]), "wait D2H")) return false;
        const Pending& p = pend[s];
        host_copy2d(p.dst, p.dst_pitch * sizeof(float), h_out[s], p.width * sizeof(float), p.width * sizeof(float), p.rows);
        pend[s].dst = nullptr;
    }
    return cuda_ok(cudaStreamWaitEvent(s_in, e_out[s], 0), "wait");   // no-op until e_out[s] has been recorded
}

bool Pipeline::h2d(int s, float* dev_dst, size_t dev_pitch, const float* src, size_t src_pitch, size_t width, size_t rows)
{
    const size_t fl = width * rows;
    if (fl == 0) return true;
    if (!bounce_in || fl * sizeof(float) <= kCopyInline || fl > cap_h_in)   // small: the driver's own staging is fine
        return cuda_ok(cudaMemcpy2DAsync(dev_dst, dev_pitch * sizeof(float), src, src_pitch * sizeof(float), width * sizeof(float),
                                         rows, cudaMemcpyHostToDevice, s_in), "H2D");
    // the slot's previous H2D has long finished (its output has been consumed), but make it explicit
    if (!cuda_ok(cudaEventSynchronize(e_in[s]), "wait H2D")) return false;
    host_copy2d(h_in[s], width * sizeof(float), src, src_pitch * sizeof(float), width * sizeof(float), rows);
    return cuda_ok(cudaMemcpy2DAsync(dev_dst, dev_pitch * sizeof(float), h_in[s], width * sizeof(float), width * sizeof(float), rows,
                                     cudaMemcpyHostToDevice, s_in), "H2D");
}

bool Pipeline::d2h(int s, float* dst, size_t dst_pitch, const float* dev_src, size_t dev_pitch, size_t width, size_t rows)
{
    const size_t fl = width * rows;
    if (fl == 0) return true;
    if (!bounce_out || fl * sizeof(float) <= kCopyInline || fl > cap_h_out || pend[s].dst)
        return cuda_ok(cudaMemcpy2DAsync(dst, dst_pitch * sizeof(float), dev_src, dev_pitch * sizeof(float), width * sizeof(float), rows,
                                         cudaMemcpyDeviceToHost, s_out), "D2H");
    if (!cuda_ok(cudaMemcpy2DAsync(h_out[s], width * sizeof(float), dev_src, dev_pitch * sizeof(float), width * sizeof(float), rows,
                                   cudaMemcpyDeviceToHost, s_out), "D2H")) return false;
    pend[s] = Pending{dst, dst_pitch, width, rows, ++seq};
    return true;
}

bool Pipeline::finish()
{
    bool ok = true;
    for (;;) {   // oldest pending output first: its copy overlaps the D2H of the younger ones
        int best = -1;
        for (int s = 0; s < kSlots; ++s)
            if (pend[s].dst && (best < 0 || pend[s].seq < pend[best].seq)) best = s;
        if (best < 0) break;
        if (!reuse(best)) { ok = false; pend[best].dst = nullptr; }
    }
    ok = cuda_ok(cudaStreamSynchronize(s_out), "sync") && ok;
    ok = cuda_ok(cudaStreamSynchronize(s_k), "sync") && ok;
    ok = cuda_ok(cudaStreamSynchronize(s_in), "sync") && ok;
    return ok;
}

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
    if (bounce_in && need_in > cap_h_in) {
        for (int i = 0; i < kSlots; ++i) { if (h_in[i]) cudaFreeHost(h_in[i]); h_in[i] = nullptr; }
        cap_h_in = 0;
        for (int i = 0; i < kSlots; ++i)
            if (!cuda_ok(cudaHostAlloc(&h_in[i], need_in * sizeof(float), cudaHostAllocPortable), "cudaHostAlloc(bounce in)")) return false;
        cap_h_in = need_in;
    }
    if (bounce_out && need_out > cap_h_out) {
        for (int i = 0; i < kSlots; ++i) { if (h_out[i]) cudaFreeHost(h_out[i]); h_out[i] = nullptr; }
        cap_h_out = 0;
        for (int i = 0; i < kSlots; ++i)
            if (!cuda_ok(cudaHostAlloc(&h_out[i], need_out * sizeof(float), cudaHostAllocPortable), "cudaHostAlloc(bounce out)")) return false;
        cap_h_out = need_out;
    }
    return true;
}

void Pipeline::release()
{
    for (int i = 0; i < kSlots; ++i) {
        if (d_in[i]) cudaFree(d_in[i]);
        if (d_out[i]) cudaFree(d_out[i]);
        d_in[i] = d_out[i] = nullptr;
        if (h_in[i]) cudaFreeHost(h_in[i]);
        if (h_out[i]) cudaFreeHost(h_out[i]);
        h_in[i] = h_out[i] = nullptr;
        if (e_in[i]) { cudaEventDestroy(e_in[i]); cudaEventDestroy(e_k[i]); cudaEventDestroy(e_out[i]); }
        e_in[i] = e_k[i] = e_out[i] = nullptr;
    }
    if (s_in) { cudaStreamDestroy(s_in); cudaStreamDestroy(s_k); cudaStreamDestroy(s_out); }
    s_in = s_k = s_out = nullptr;
    cap_in = cap_out = cap_h_in = cap_h_out = 0;
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

size_t chunk_floats(bool bounce)
{
    // samples per staged chunk: small enough that pipeline fill + drain (one chunk each way) is a few
    // percent of a large transfer, large enough to stay near PCIe peak.  Env override for experiments / tests.
    // Bounced (pageable) calls use smaller chunks: their fill and drain also pay a host copy, and every slot
    // holds two pinned buffers.
    static size_t v[2] = {0, 0};
    static std::once_flag once;
    std::call_once(once, [] {
        const char* e = getenv("SAVGOL_B200_CHUNK_MIB");
        size_t mib = e ? static_cast<size_t>(atoi(e)) : 64;  // measured on B200: 64 MiB 23.0 ms/GiB-step, 16 MiB 23.2, 4 MiB 26.9
        if (mib < 1) mib = 1;
        v[0] = mib << 18;
        const char* b = getenv("SAVGOL_B200_BOUNCE_MIB");
        size_t bm = b ? static_cast<size_t>(atoi(b)) : 16;
        if (bm < 1) bm = 1;
        v[1] = std::min(mib, bm) << 18;
    });
    return v[bounce ? 1 : 0];
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
    const size_t piece = P.begin(x, y, b > a ? b - a : 0);
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
        if (!P.reuse(s)) return false;
        float* dx = P.d_in[s] + padl;
        const bool has_l = s0 >= n && s0 > 0, has_r = s1 + n <= L && s1 < L;   // n real samples exist on that side
        // body plus whatever neighbouring samples exist
        const size_t c0 = (has_l && !inplace) ? s0 - n : s0, c1 = has_r ? s1 + n : s1;
        if (!P.h2d(s, dx - (s0 - c0), c1 - c0, x + c0, c1 - c0, c1 - c0, 1)) return false;
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
        if (!P.d2h(s, y + (s0 - a), plen, P.d_out[s], plen, plen, 1)) return false;
        cudaEventRecord(P.e_out[s], P.s_out);
    }
    // `snaps`, `head` are pageable and small: their H2D copies were staged before cudaMemcpyAsync returned
    return P.finish();
}

bool run1d_host(const SavgolFilter* f, const float* in, float* out, size_t rows, size_t len, size_t in_pitch, size_t out_pitch,
                int mode, bool poly_edges, int arith)
{
    PipeLease lease;
    if (!lease.ok()) return false;
    Pipeline& P = *lease;
    const size_t chunk = P.begin(in, out, rows * len);

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
        // slot reuse: its previous D2H must have drained before we overwrite d_out/d_in
        if (!P.reuse(s)) return false;
        if (!P.h2d(s, P.d_in[s], len, in + done * in_pitch, in_pitch, len, nr)) return false;
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
        if (!P.d2h(s, out + done * out_pitch, out_pitch, P.d_out[s], len, len, nr)) return false;
        cudaEventRecord(P.e_out[s], P.s_out);
    }
    return P.finish();
}

}  // namespace sge
