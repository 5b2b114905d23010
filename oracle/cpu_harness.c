/*
 * cpu_harness.c -- TEST / BENCH INFRASTRUCTURE (not product code).
 *
 * A pthread driver that runs a CPU implementation of the hot path over
 * independent signals / images / channels on all host cores.  The reference
 * itself has no batching or threading: "a batch" is the caller's loop over
 * savgol_apply, and the header states that apply is thread-safe on a shared
 * filter (ref: include/iterative/savgolFilter.h:16-19).  This file is that
 * caller loop, parallelised over rows, so that bench.py can time the
 * reference's CPU path "with all the host threads it can use".
 *
 * The functions take *function pointers with the reference's own signatures*
 * (obtained by ctypes from oracle/_ref/libsavgol_ref.so, i.e. the unmodified
 * reference compiled from /root/reference), so nothing here restates any
 * arithmetic.
 */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdbool.h>

/* ref: include/iterative/savgolFilter.h:152-153 */
typedef int (*sgh_apply_fn)(const void *filter, const float *in, float *out, size_t len);
/* ref: include/iterative/savgol2d.h:171-174 */
typedef int (*sgh_apply2d_fn)(const void *filter, const float *in, int rows, int cols, int is,
                              float *out, int os, int boundary);
/* ref: include/iterative/savgol_stream.h:58,95-96,106 */
typedef int (*sgh_stream_init_fn)(void *stream, const void *filter);
typedef int (*sgh_stream_push_full_fn)(void *stream, float x, float *out, int max_out);
typedef int (*sgh_stream_flush_fn)(void *stream, float *out, int max_out);

/* ref: include/iterative/savgolFilter.h:201-203 */
typedef size_t (*sgh_valid_fn)(const void *filter, const float *in, size_t in_len, float *out);

typedef struct {
    int kind; /* 0 rows, 1 images, 2 stream channels, 3 chunks of one long signal (VALID) */
    void *fn, *fn2, *fn3;
    const void *filter;
    const float *in;
    float *out;
    size_t begin, end;       /* unit range of this worker */
    size_t len, ipitch, opitch;
    int rows, cols, boundary, flush;
    int rc;
} sgh_job;

static void *sgh_worker(void *arg)
{
    sgh_job *j = (sgh_job *)arg;
    j->rc = 0;
    if (j->kind == 0) {
        sgh_apply_fn f = (sgh_apply_fn)j->fn;
        for (size_t r = j->begin; r < j->end; ++r)
            if (f(j->filter, j->in + r * j->ipitch, j->out + r * j->opitch, j->len)) j->rc = -1;
    } else if (j->kind == 1) {
        sgh_apply2d_fn f = (sgh_apply2d_fn)j->fn;
        for (size_t r = j->begin; r < j->end; ++r)
            if (f(j->filter, j->in + r * j->ipitch, j->rows, j->cols, j->cols,
                  j->out + r * j->opitch, j->cols, j->boundary)) j->rc = -1;
    } else if (j->kind == 3) {
        /* chunk c = outputs [c*len, min((c+1)*len, total)) of one long signal; j->in is the signal with n
         * samples of context on either side (j->rows = n, j->ipitch = total outputs) */
        sgh_valid_fn f = (sgh_valid_fn)j->fn;
        const size_t n = (size_t)j->rows, total = j->ipitch;
        for (size_t c = j->begin; c < j->end; ++c) {
            const size_t o = c * j->len;
            const size_t l = o + j->len <= total ? j->len : total - o;
            if (f(j->filter, j->in + o, l + 2 * n, j->out + o) != l) j->rc = -1;
        }
    } else {
        /* one SavgolStream (296 B in the reference ABI) per channel, on the stack */
        sgh_stream_init_fn init = (sgh_stream_init_fn)j->fn;
        sgh_stream_push_full_fn push = (sgh_stream_push_full_fn)j->fn2;
        sgh_stream_flush_fn flush = (sgh_stream_flush_fn)j->fn3;
        for (size_t c = j->begin; c < j->end; ++c) {
            _Alignas(16) unsigned char st[512];
            float tmp[40];
            if (init(st, j->filter)) { j->rc = -1; continue; }
            const float *x = j->in + c * j->ipitch;
            float *y = j->out + c * j->opitch;
            size_t o = 0;
            for (size_t t = 0; t < j->len; ++t) {
                int k = push(st, x[t], tmp, 40);
                for (int q = 0; q < k; ++q) y[o++] = tmp[q];
            }
            if (j->flush) {
                int k = flush(st, tmp, 40);
                for (int q = 0; q < k; ++q) y[o++] = tmp[q];
            }
        }
    }
    return NULL;
}

static int sgh_run(sgh_job proto, size_t units, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > units) nthreads = (int)(units ? units : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    sgh_job *jobs = (sgh_job *)malloc(sizeof(sgh_job) * (size_t)nthreads);
    if (!th || !jobs) { free(th); free(jobs); return -1; }
    int rc = 0;
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = proto;
        jobs[t].begin = units * (size_t)t / (size_t)nthreads;
        jobs[t].end = units * (size_t)(t + 1) / (size_t)nthreads;
        if (pthread_create(&th[t], NULL, sgh_worker, &jobs[t])) { jobs[t].rc = -2; sgh_worker(&jobs[t]); }
    }
    for (int t = 0; t < nthreads; ++t) {
        if (jobs[t].rc != -2) pthread_join(th[t], NULL);
        if (jobs[t].rc == -1) rc = -1;
    }
    free(th); free(jobs);
    return rc;
}

int sgh_apply_rows(void *apply_fn, const void *filter, const float *in, float *out,
                   size_t rows, size_t len, size_t ipitch, size_t opitch, int nthreads)
{
    sgh_job p = {0};
    p.kind = 0; p.fn = apply_fn; p.filter = filter; p.in = in; p.out = out;
    p.len = len; p.ipitch = ipitch; p.opitch = opitch;
    return sgh_run(p, rows, nthreads);
}

int sgh_apply2d_images(void *apply2d_fn, const void *filter, const float *in, float *out,
                       size_t images, int rows, int cols, int boundary, int nthreads)
{
    sgh_job p = {0};
    p.kind = 1; p.fn = apply2d_fn; p.filter = filter; p.in = in; p.out = out;
    p.rows = rows; p.cols = cols; p.boundary = boundary;
    p.ipitch = p.opitch = (size_t)rows * (size_t)cols;
    return sgh_run(p, images, nthreads);
}

int sgh_stream_channels(void *init_fn, void *push_full_fn, void *flush_fn, const void *filter,
                        const float *in, float *out, size_t channels, size_t len,
                        size_t ipitch, size_t opitch, int flush, int nthreads)
{
    sgh_job p = {0};
    p.kind = 2; p.fn = init_fn; p.fn2 = push_full_fn; p.fn3 = flush_fn; p.filter = filter;
    p.in = in; p.out = out; p.len = len; p.ipitch = ipitch; p.opitch = opitch; p.flush = flush;
    return sgh_run(p, channels, nthreads);
}

/* One long signal through the reference's VALID path (size_t-clean, ref: src/savgolFilter.c:821-850), cut into
 * chunks of `chunk` outputs that run on `nthreads` threads.  `in` holds n samples of context before and after
 * the `total` output positions (for a PERIODIC signal: the wrap-padded signal -- SURVEY.md Q6: periodic ==
 * VALID over the wrap-padded signal, bit for bit). */
int sgh_valid_chunks(void *valid_fn, const void *filter, const float *in, float *out, size_t total,
                     int half_window, size_t chunk, int nthreads)
{
    sgh_job p = {0};
    p.kind = 3; p.fn = valid_fn; p.filter = filter; p.in = in; p.out = out;
    p.len = chunk; p.ipitch = total; p.rows = half_window;
    return sgh_run(p, (total + chunk - 1) / chunk, nthreads);
}
