// capi_mcstream.cu -- multi-channel chunked streaming (include/savgol_b200.h part 2).
//
// The reference streams one channel, one sample per call (src/savgol_stream.c:152-252).  The
// data-parallel form of the same contract: C independent channels advance in lockstep, one chunk
// of K samples per channel per call, with the reference's per-channel semantics preserved --
// fixed latency of half_window samples, polynomial leading edge when the window first fills,
// one centred output per further sample, polynomial trailing edge on flush, boundary mode
// ignored.  Per channel the carry state is the last 2n+1 samples (device memory,
// double buffered); steady-state chunks run the 1D kernel with that history as the left pad.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "coeffs.h"
#include "engine.h"

namespace sg { extern std::atomic<unsigned long long> g_launches; }

using sge::cuda_ok;
using sge::MemKind;

struct SavgolMCStream {
    SavgolFilter* filter;
    size_t channels;
    int n, ws;
    float* state[2];  // [channels][ws], most recent samples right-aligned, chronological
    int cur;
    size_t received, emitted;  // per channel
    int device;
};

namespace {

// state_new[c][i] = i < ws-K ? state_old[c][i+K] : chunk[c][i-(ws-K)]   (K < ws)
__global__ void state_append_kernel(const float* __restrict__ old_state, float* __restrict__ new_state,
                                    const float* __restrict__ chunk, size_t chunk_pitch, size_t channels, int ws, int K)
{
    const size_t total = channels * static_cast<size_t>(ws);
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t c = i / ws;
        const int k = static_cast<int>(i - c * ws);
        new_state[i] = k < ws - K ? old_state[c * ws + k + K] : chunk[c * chunk_pitch + (k - (ws - K))];
    }
}

// Trailing edge from the carried window: out[c][i] = scale * sum_k E[n-1-i][k] * state[c][k]
// (ref: src/savgol_stream.c:43-56, 245-249).  One thread per channel.
template <bool EXACT>
__global__ void flush_kernel(const float* __restrict__ state, const float* __restrict__ edge_t, float* __restrict__ out,
                             size_t out_pitch, size_t channels, int n, int ws, float scale)
{
    const size_t c = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (c >= channels) return;
    float x[sg::kMaxWs];
    for (int k = 0; k < ws; ++k) x[k] = state[c * ws + k];
    for (int i = 0; i < n; ++i) {
        const int e = n - 1 - i;
        float s = 0.0f;
        for (int k = 0; k < ws; ++k) {
            const float w = edge_t[k * 32 + e];
            s = EXACT ? __fadd_rn(s, __fmul_rn(w, x[k])) : fmaf(w, x[k], s);
        }
        out[c * out_pitch + i] = EXACT ? __fmul_rn(s, scale) : s * scale;
    }
}

unsigned grid_for(size_t work, int block)
{
    size_t g = (work + block - 1) / block;
    const size_t cap = 148 * 32;
    return static_cast<unsigned>(g < cap ? (g ? g : 1) : cap);
}

int stream_arith() { return sge::exact_mode() ? sg::ARITH_EXACTSEQ : sg::ARITH_FAST; }

// Device-pointer push of channels [c0, c0 + nc): `in` / `out` point at channel c0.  Does not advance the stream
// (the caller flips the state buffers and the counters once all channel blocks are queued).  Returns outputs
// per channel or -1.
long long push_range(SavgolMCStream* s, size_t c0, size_t nc, const float* in, size_t in_pitch, size_t K, float* out,
                     size_t out_pitch, cudaStream_t st)
{
    const size_t T = s->received;
    const size_t ws = static_cast<size_t>(s->ws);
    const int n = s->n;
    float* cur = s->state[s->cur] + c0 * ws;
    float* nxt = s->state[s->cur ^ 1] + c0 * ws;
    long long produced;

    if (T + K < ws) {
        state_append_kernel<<<grid_for(nc * ws, 256), 256, 0, st>>>(cur, nxt, in, in_pitch, nc, s->ws,
                                                                               static_cast<int>(K));
        sg::g_launches.fetch_add(1);
        if (!cuda_ok(cudaGetLastError(), "state append")) return -1;
        produced = 0;
    } else if (T < ws) {
        // first fill: [carried T samples | chunk] is the head of the signal -> batch kernel with the
        // polynomial leading edge; the last n positions are not outputs yet.
        const float* x = in;
        size_t xp = in_pitch;
        float* tmp = nullptr;
        const size_t L = T + K;
        if (T > 0) {
            if (!cuda_ok(cudaMallocAsync(&tmp, nc * L * sizeof(float), st), "cudaMallocAsync(first fill)")) return -1;
            bool ok = cuda_ok(cudaMemcpy2DAsync(tmp, L * sizeof(float), cur + (ws - T), ws * sizeof(float), T * sizeof(float),
                                                nc, cudaMemcpyDeviceToDevice, st), "gather state") &&
                      cuda_ok(cudaMemcpy2DAsync(tmp + T, L * sizeof(float), in, in_pitch * sizeof(float), K * sizeof(float),
                                                nc, cudaMemcpyDeviceToDevice, st), "gather chunk");
            if (!ok) { cudaFreeAsync(tmp, st); return -1; }
            x = tmp; xp = L;
        }
        sge::Problem1D p{};
        p.filter = s->filter; p.in = x; p.out = out; p.rows = nc; p.len = L;
        p.in_row_bytes = xp * sizeof(float); p.out_row_bytes = out_pitch * sizeof(float);
        p.in_stride = p.out_stride = 4;
        p.mode = sg::MODE_POLY; p.edge_lead = true; p.edge_trail = false;
        p.out_len = L - n;  // the last n positions are not outputs yet
        p.state_out = nxt; p.state_pitch = ws; p.state_w = s->ws;
        p.arith = sge::exact_mode() ? sg::ARITH_EXACTSEQ : sg::ARITH_FAST;
        const bool ok = sge::run1d_device(p, st);
        if (tmp) cudaFreeAsync(tmp, st);
        if (!ok) return -1;
        produced = static_cast<long long>(L) - n;
    } else {
        sge::Problem1D p{};
        p.filter = s->filter; p.in = in; p.out = out; p.rows = nc; p.len = K;
        p.in_row_bytes = in_pitch * sizeof(float); p.out_row_bytes = out_pitch * sizeof(float);
        p.in_stride = p.out_stride = 4;
        p.lhalo = cur + 1; p.lhalo_pitch = ws;  // the 2n most recent samples
        p.mode = sg::MODE_POLY;
        p.stream_history = true;
        p.state_out = nxt; p.state_pitch = ws; p.state_w = s->ws;
        p.arith = stream_arith();
        if (!sge::run1d_device(p, st)) return -1;
        produced = static_cast<long long>(K);
    }
    return produced;
}

void advance(SavgolMCStream* s, size_t K, long long produced)
{
    s->cur ^= 1;
    s->received += K;
    s->emitted += static_cast<size_t>(produced);
}

bool on_stream_device(const SavgolMCStream* s, const char* who)
{
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev == s->device) return true;
    fprintf(stderr, "%s: the stream lives on device %d but device %d is current\n", who, s->device, dev);
    return false;
}

}  // namespace

extern "C" {

SavgolMCStream* savgol_mcstream_create(const SavgolConfig* config, size_t channels)
{
    if (!config || channels == 0) return nullptr;
    if (!sge::device_ready(true)) return nullptr;
    SavgolFilter* f = savgol_create(config);
    if (!f) return nullptr;
    SavgolMCStream* s = static_cast<SavgolMCStream*>(calloc(1, sizeof(SavgolMCStream)));
    if (!s) { savgol_destroy(f); return nullptr; }
    s->filter = f;
    s->channels = channels;
    s->n = config->half_window;
    s->ws = 2 * s->n + 1;
    cudaGetDevice(&s->device);
    const size_t bytes = channels * static_cast<size_t>(s->ws) * sizeof(float);
    if (!cuda_ok(cudaMalloc(&s->state[0], bytes), "cudaMalloc(stream state)") ||
        !cuda_ok(cudaMalloc(&s->state[1], bytes), "cudaMalloc(stream state)")) {
        savgol_mcstream_destroy(s);
        return nullptr;
    }
    savgol_mcstream_reset(s);
    return s;
}

void savgol_mcstream_destroy(SavgolMCStream* s)
{
    if (!s) return;
    if (s->state[0]) cudaFree(s->state[0]);
    if (s->state[1]) cudaFree(s->state[1]);
    savgol_destroy(s->filter);
    free(s);
}

void savgol_mcstream_reset(SavgolMCStream* s)
{
    if (!s) return;
    const size_t bytes = s->channels * static_cast<size_t>(s->ws) * sizeof(float);
    cudaStream_t st = sge::current_stream();
    cudaMemsetAsync(s->state[0], 0, bytes, st);
    cudaMemsetAsync(s->state[1], 0, bytes, st);
    s->cur = 0;
    s->received = s->emitted = 0;
}

long long savgol_mcstream_push(SavgolMCStream* s, const float* input, size_t in_pitch, size_t chunk_len,
                               float* output, size_t out_pitch)
{
    if (!s || !input || !output || chunk_len == 0) return -1;
    // the chunk that first fills the window emits up to chunk_len + half_window outputs per channel
    if (in_pitch < chunk_len || out_pitch < chunk_len + static_cast<size_t>(s->n)) {
        fprintf(stderr, "savgol_mcstream_push: need in_pitch >= chunk_len and out_pitch >= chunk_len + half_window\n");
        return -1;
    }
    if (!on_stream_device(s, "savgol_mcstream_push")) return -1;
    cudaStream_t st = sge::current_stream();
    const MemKind ki = sge::classify(input), ko = sge::classify(output);
    if (ki == MemKind::Device && ko == MemKind::Device) {
        const long long produced = push_range(s, 0, s->channels, input, in_pitch, chunk_len, output, out_pitch, st);
        if (produced >= 0) advance(s, chunk_len, produced);
        return produced;
    }
    if (ki == MemKind::Device || ko == MemKind::Device) {
        fprintf(stderr, "savgol_b200: input and output must both be device pointers or both be host pointers\n");
        return -1;
    }
    // Host chunks: blocks of channels travel through the three-slot staging pipeline (H2D of block i+1 and D2H of
    // block i-1 overlap the kernel of block i).  The carry state stays on the device; the blocks' kernels run in
    // order on one stream and touch disjoint channels of it.  Work queued earlier on the caller's stream (a
    // previous device push, a reset) is ordered before the first block.
    sge::PipeLease lease;
    if (!lease.ok()) return -1;
    sge::Pipeline& P = *lease;
    const size_t K = chunk_len;
    const size_t opitch = (K + static_cast<size_t>(s->ws) + 3) & ~static_cast<size_t>(3);   // 16-byte aligned rows
    const size_t cb = std::max<size_t>(1, std::min(s->channels, P.begin(input, output, s->channels * opitch) / opitch));
    if (!P.ensure(cb * K, cb * opitch)) return -1;
    cudaEvent_t prior = nullptr;
    if (cudaEventCreateWithFlags(&prior, cudaEventDisableTiming) == cudaSuccess) {
        cudaEventRecord(prior, st);
        cudaStreamWaitEvent(P.s_k, prior, 0);
        cudaEventDestroy(prior);
    }
    long long produced = -1;
    size_t c0 = 0;
    for (size_t b = 0; c0 < s->channels; ++b, c0 += cb) {
        const int sl = static_cast<int>(b % sge::Pipeline::kSlots);
        const size_t nc = std::min(cb, s->channels - c0);
        if (!P.reuse(sl)) return -1;
        if (!P.h2d(sl, P.d_in[sl], K, input + c0 * in_pitch, in_pitch, K, nc)) return -1;
        cudaEventRecord(P.e_in[sl], P.s_in);
        cudaStreamWaitEvent(P.s_k, P.e_in[sl], 0);
        if (b >= sge::Pipeline::kSlots) cudaStreamWaitEvent(P.s_k, P.e_out[sl], 0);
        const long long k = push_range(s, c0, nc, P.d_in[sl], K, K, P.d_out[sl], opitch, P.s_k);
        if (k < 0) { cudaStreamSynchronize(P.s_k); return -1; }
        produced = k;
        cudaEventRecord(P.e_k[sl], P.s_k);
        cudaStreamWaitEvent(P.s_out, P.e_k[sl], 0);
        if (k > 0 && !P.d2h(sl, output + c0 * out_pitch, out_pitch, P.d_out[sl], opitch, static_cast<size_t>(k), nc)) return -1;
        cudaEventRecord(P.e_out[sl], P.s_out);
    }
    if (!P.finish()) return -1;
    advance(s, K, produced);
    return produced;
}

long long savgol_mcstream_flush(SavgolMCStream* s, float* output, size_t out_pitch)
{
    if (!s || !output) return -1;
    if (s->received < static_cast<size_t>(s->ws)) return 0;
    if (out_pitch < static_cast<size_t>(s->n) && s->channels > 1) return -1;
    if (!on_stream_device(s, "savgol_mcstream_flush")) return -1;
    cudaStream_t st = sge::current_stream();
    float* temp_edges = nullptr;
    const float* et = sge::edge_table_device(s->filter, st, &temp_edges);
    if (!et) return -1;
    const float scale = s->filter->dt_scale != 0.0f ? 1.0f / s->filter->dt_scale : 1.0f;
    const bool host_out = sge::classify(output) != MemKind::Device;
    float* dout = output;
    size_t dpitch = out_pitch;
    bool ok = true;
    if (host_out) {
        dpitch = static_cast<size_t>(s->n);
        ok = cuda_ok(cudaMallocAsync(&dout, s->channels * dpitch * sizeof(float), st), "cudaMallocAsync");
    }
    if (ok) {
        const unsigned grid = static_cast<unsigned>((s->channels + 127) / 128);
        if (sge::exact_mode())
            flush_kernel<true><<<grid, 128, 0, st>>>(s->state[s->cur], et, dout, dpitch, s->channels, s->n, s->ws, scale);
        else
            flush_kernel<false><<<grid, 128, 0, st>>>(s->state[s->cur], et, dout, dpitch, s->channels, s->n, s->ws, scale);
        sg::g_launches.fetch_add(1);
        ok = cuda_ok(cudaGetLastError(), "flush launch");
    }
    if (ok && host_out) {
        ok = cuda_ok(cudaMemcpy2DAsync(output, out_pitch * sizeof(float), dout, dpitch * sizeof(float),
                                       static_cast<size_t>(s->n) * sizeof(float), s->channels, cudaMemcpyDeviceToHost, st), "D2H") &&
             cuda_ok(cudaStreamSynchronize(st), "sync");
    }
    if (host_out && dout) cudaFreeAsync(dout, st);
    if (temp_edges) cudaFreeAsync(temp_edges, st);
    if (!ok) return -1;
    s->emitted += static_cast<size_t>(s->n);
    return s->n;
}

size_t savgol_mcstream_channels(const SavgolMCStream* s) { return s ? s->channels : 0; }
size_t savgol_mcstream_latency(const SavgolMCStream* s) { return s ? static_cast<size_t>(s->n) : 0; }
size_t savgol_mcstream_samples_received(const SavgolMCStream* s) { return s ? s->received : 0; }
size_t savgol_mcstream_samples_output(const SavgolMCStream* s) { return s ? s->emitted : 0; }
// Checkpoint blob: header + the carry state.  Plain host memory, self-describing, so a stream can be
// resumed in another process / on another GPU (same configuration and channel count).
struct McCheckpoint {
    unsigned magic, version;
    unsigned long long channels, received, emitted;
    int n, ws;
};
constexpr unsigned kCkptMagic = 0x53474d43u;  // "SGMC"

size_t savgol_mcstream_checkpoint_size(const SavgolMCStream* s)
{
    return s ? sizeof(McCheckpoint) + s->channels * static_cast<size_t>(s->ws) * sizeof(float) : 0;
}

long long savgol_mcstream_save(SavgolMCStream* s, void* blob, size_t capacity)
{
    const size_t need = savgol_mcstream_checkpoint_size(s);
    if (!s || !blob || capacity < need) return -1;
    if (!on_stream_device(s, "savgol_mcstream_save")) return -1;
    McCheckpoint h{kCkptMagic, 1u, s->channels, s->received, s->emitted, s->n, s->ws};
    std::memcpy(blob, &h, sizeof h);
    cudaStream_t st = sge::current_stream();
    if (!cuda_ok(cudaMemcpyAsync(static_cast<char*>(blob) + sizeof h, s->state[s->cur], need - sizeof h, cudaMemcpyDeviceToHost, st),
                 "checkpoint D2H") ||
        !cuda_ok(cudaStreamSynchronize(st), "sync"))
        return -1;
    return static_cast<long long>(need);
}

int savgol_mcstream_restore(SavgolMCStream* s, const void* blob, size_t bytes)
{
    if (!s || !blob || bytes < sizeof(McCheckpoint)) return -1;
    McCheckpoint h;
    std::memcpy(&h, blob, sizeof h);
    const size_t need = savgol_mcstream_checkpoint_size(s);
    // a buffer larger than the checkpoint (the caller's capacity, trailing data) is fine: exactly the carry state
    // of THIS stream is read, never more
    if (h.magic != kCkptMagic || h.version != 1u || h.channels != s->channels || h.n != s->n || h.ws != s->ws || bytes < need ||
        h.emitted > h.received) {
        std::fprintf(stderr, "savgol_mcstream_restore: checkpoint does not match this stream\n");
        return -1;
    }
    if (!on_stream_device(s, "savgol_mcstream_restore")) return -1;
    cudaStream_t st = sge::current_stream();
    if (!cuda_ok(cudaMemcpyAsync(s->state[s->cur], static_cast<const char*>(blob) + sizeof h, need - sizeof h, cudaMemcpyHostToDevice, st),
                 "checkpoint H2D") ||
        !cuda_ok(cudaStreamSynchronize(st), "sync"))
        return -1;
    s->received = static_cast<size_t>(h.received);
    s->emitted = static_cast<size_t>(h.emitted);
    return 0;
}

float* savgol_mcstream_state(SavgolMCStream* s, size_t* n_floats)
{
    if (!s) return nullptr;
    if (n_floats) *n_floats = s->channels * static_cast<size_t>(s->ws);
    return s->state[s->cur];
}

}  // extern "C"
