// capi_stream.cpp -- the reference's single-channel, one-sample-at-a-time stream API
// (include/iterative/savgol_stream.h, src/savgol_stream.c:80-315), kept as a host shim so that
// code written against the reference links unchanged.  One sample in, at most n+1 samples out:
// there is nothing to parallelise, and a kernel launch per sample would be absurd, so this is
// plain host arithmetic in the reference's summation order (single accumulator, oldest sample
// first).  The data-parallel streaming path is savgol_mcstream_* (capi_mcstream.cu).
//
// Compiled with -ffp-contract=off so the sums round exactly like the reference's.
#include <cstdlib>
#include <cstring>

#include "../../include/savgol_b200.h"

namespace {

// Window sample `age` positions after the oldest one (0 = oldest, ws-1 = newest).
// write_pos always points at the oldest sample once the ring is full.
inline float ring_at(const SavgolStream* s, int ws, int age) { return s->buffer[(s->write_pos + age) % ws]; }

float dot_oldest_first(const SavgolStream* s, const float* w)
{
    const int ws = s->filter->window_size;
    float acc = 0.0f;
    for (int i = 0; i < ws; ++i) acc += w[i] * ring_at(s, ws, i);
    return acc;
}

float dot_newest_first(const SavgolStream* s, const float* w)
{
    const int ws = s->filter->window_size;
    float acc = 0.0f;
    for (int i = 0; i < ws; ++i) acc += w[i] * ring_at(s, ws, ws - 1 - i);
    return acc;
}

inline float inv_scale(const SavgolFilter* f) { return f->dt_scale != 0.0f ? 1.0f / f->dt_scale : 1.0f; }

inline void take_sample(SavgolStream* s, float x)
{
    const int ws = s->filter->window_size;
    s->buffer[s->write_pos] = x;
    s->write_pos = (s->write_pos + 1) % ws;
    s->samples_received++;
}

}  // namespace

extern "C" {

SavgolStream* savgol_stream_create(const SavgolConfig* config)
{
    if (!config) return nullptr;
    SavgolFilter* f = savgol_create(config);
    if (!f) return nullptr;
    SavgolStream* s = static_cast<SavgolStream*>(malloc(sizeof(SavgolStream)));
    if (!s) { savgol_destroy(f); return nullptr; }
    s->filter = f;
    s->owns_filter = true;
    s->dt_inv = inv_scale(f);
    savgol_stream_reset(s);
    return s;
}

int savgol_stream_init(SavgolStream* stream, const SavgolFilter* filter)
{
    if (!stream || !filter) return -1;
    stream->filter = filter;
    stream->owns_filter = false;
    stream->dt_inv = inv_scale(filter);
    savgol_stream_reset(stream);
    return 0;
}

void savgol_stream_destroy(SavgolStream* stream)
{
    if (!stream) return;
    if (stream->owns_filter && stream->filter) savgol_destroy(const_cast<SavgolFilter*>(stream->filter));
    free(stream);
}

void savgol_stream_reset(SavgolStream* stream)
{
    if (!stream) return;
    stream->write_pos = 0;
    stream->samples_received = 0;
    stream->samples_output = 0;
    memset(stream->buffer, 0, sizeof(stream->buffer));
}

float savgol_stream_push(SavgolStream* stream, float sample, bool* output_valid)
{
    if (!stream || !stream->filter) {
        if (output_valid) *output_valid = false;
        return 0.0f;
    }
    take_sample(stream, sample);
    const bool full = stream->samples_received >= static_cast<size_t>(stream->filter->window_size);
    if (output_valid) *output_valid = full;
    if (!full) return 0.0f;
    stream->samples_output++;
    return dot_oldest_first(stream, stream->filter->center_weights) * stream->dt_inv;
}

int savgol_stream_push_full(SavgolStream* stream, float sample, float* output, int max_outputs)
{
    if (!stream || !stream->filter || !output || max_outputs <= 0) return 0;
    const SavgolFilter* f = stream->filter;
    const size_t ws = static_cast<size_t>(f->window_size);
    const int n = f->config.half_window;
    const bool first_fill = stream->samples_received + 1 == ws;

    take_sample(stream, sample);
    if (stream->samples_received < ws) return 0;

    int produced = 0;
    if (first_fill) {
        // the window just filled: the n leading-edge outputs come first (newest-first traversal
        // of edge row e), then the first centred output
        for (int e = 0; e < n && produced < max_outputs; ++e) {
            output[produced++] = dot_newest_first(stream, f->edge_weights[e]) * stream->dt_inv;
            stream->samples_output++;
        }
        if (produced < max_outputs) {
            output[produced++] = dot_oldest_first(stream, f->center_weights) * stream->dt_inv;
            stream->samples_output++;
        }
        return produced;
    }
    output[0] = dot_oldest_first(stream, f->center_weights) * stream->dt_inv;
    stream->samples_output++;
    return 1;
}

int savgol_stream_flush(SavgolStream* stream, float* output, int max_count)
{
    if (!stream || !output || max_count <= 0) return -1;
    const SavgolFilter* f = stream->filter;
    const int n = f->config.half_window;
    if (stream->samples_received < static_cast<size_t>(f->window_size)) return 0;
    const int count = max_count < n ? max_count : n;
    for (int i = 0; i < count; ++i) {
        output[i] = dot_oldest_first(stream, f->edge_weights[n - 1 - i]) * stream->dt_inv;  // chronological
        stream->samples_output++;
    }
    return count;
}

int savgol_stream_flush_leading(SavgolStream* stream, float* output, int max_count)
{
    if (!stream || !output || max_count <= 0) return 0;
    const SavgolFilter* f = stream->filter;
    const int n = f->config.half_window;
    if (stream->samples_received < static_cast<size_t>(f->window_size)) return 0;
    const int count = max_count < n ? max_count : n;
    for (int i = 0; i < count; ++i) {
        output[i] = dot_newest_first(stream, f->edge_weights[i]) * stream->dt_inv;
        stream->samples_output++;
    }
    return count;
}

bool savgol_stream_ready(const SavgolStream* stream)
{
    return stream && stream->filter && stream->samples_received >= static_cast<size_t>(stream->filter->window_size);
}

size_t savgol_stream_latency(const SavgolStream* stream)
{
    return (stream && stream->filter) ? stream->filter->config.half_window : 0;
}

size_t savgol_stream_buffered(const SavgolStream* stream)
{
    if (!stream || !stream->filter) return 0;
    const size_t ws = static_cast<size_t>(stream->filter->window_size);
    return stream->samples_received < ws ? stream->samples_received : ws;
}

size_t savgol_stream_samples_received(const SavgolStream* stream) { return stream ? stream->samples_received : 0; }
size_t savgol_stream_samples_output(const SavgolStream* stream) { return stream ? stream->samples_output : 0; }

}  // extern "C"
