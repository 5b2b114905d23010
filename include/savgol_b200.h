/*
 * savgol_b200.h -- C ABI of libsavgol_b200.so, the B200-native (sm_100a CUDA)
 * Savitzky-Golay filtering engine.
 *
 * Part 1 is the drop-in boundary: the exact symbols, struct layouts, argument
 * meanings and error conventions of the reference library
 * (Tugbars/Savitzky-Golay-Filter, include/iterative/{savgolFilter,savgol_stream,
 * savgol2d}.h).  Code written against the reference links against this library
 * unchanged; every apply runs on the GPU.  Pointers passed to apply functions may be
 * device pointers (used in place, zero copy) or host pointers (staged through the GPU;
 * pinned host memory is streamed in overlapping chunks).
 *
 * Part 2 holds the extensions the BASELINE configs need and the reference lacks:
 * batched 1D / 2D entry points, explicit halos for partitioned signals, the
 * multi-channel chunked stream, stream/device control.
 *
 * Each declaration cites the reference interface it replaces (path:line inside the
 * reference repository).  There is no CPU fallback: without a CUDA device every
 * apply function fails (-1 / 0) and prints a diagnostic.
 */
#ifndef SAVGOL_B200_H
#define SAVGOL_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ======================================================================== */
/* Part 1a -- 1D batch filter   (ref: include/iterative/savgolFilter.h)      */
/* ======================================================================== */

/* ref: savgolFilter.h:39-48 */
#define SAVGOL_MAX_HALF_WINDOW 32
#define SAVGOL_MAX_WINDOW (2 * SAVGOL_MAX_HALF_WINDOW + 1)
#define SAVGOL_MAX_POLY_ORDER 10
#define SAVGOL_MAX_DERIVATIVE 4

/* ref: savgolFilter.h:63-68.  Edge treatment of savgol_apply(). */
typedef enum {
    SAVGOL_BOUNDARY_POLYNOMIAL = 0, /* asymmetric least-squares fits on the first/last 2n+1 samples */
    SAVGOL_BOUNDARY_REFLECT,        /* half-sample symmetric:  d1 d0 | d0 d1 d2 ...  */
    SAVGOL_BOUNDARY_PERIODIC,       /* wrap-around:            dL-1 | d0 d1 ...      */
    SAVGOL_BOUNDARY_CONSTANT        /* edge replication:       d0 d0 | d0 d1 ...     */
} SavgolBoundaryMode;

/* ref: savgolFilter.h:92-98.  sizeof == 12. */
typedef struct {
    uint8_t half_window;  /* n, window = 2n+1, 1..32                       */
    uint8_t poly_order;   /* m < 2n+1, <= 10                               */
    uint8_t derivative;   /* d <= min(m, 4); 0 = smoothing                 */
    float time_step;      /* dt > 0; outputs are scaled by 1 / dt^d        */
    SavgolBoundaryMode boundary;
} SavgolConfig;

/* ref: savgolFilter.h:107-113.  sizeof == 8600; the fields are public and read by
 * callers (tests, the coefficient export tool), so the layout is ABI.  The object
 * returned by savgol_create() carries private device state behind this prefix. */
typedef struct SavgolFilter {
    SavgolConfig config;
    int window_size;                                               /* 2n+1        */
    float dt_scale;                                                /* dt^d        */
    float center_weights[SAVGOL_MAX_WINDOW];                       /* target t=0  */
    float edge_weights[SAVGOL_MAX_HALF_WINDOW][SAVGOL_MAX_WINDOW]; /* row e: t=n-e */
} SavgolFilter;

/* ref: savgolFilter.h:130 / src/savgolFilter.c:688-718.  Validates, computes the fp32
 * GenFact / Gram-polynomial weight tables on the host (bit-identical to the reference).
 * NULL + one line on stderr for an invalid configuration.  Not thread-safe. */
SavgolFilter *savgol_create(const SavgolConfig *config);

/* ref: savgolFilter.h:137.  NULL is a no-op. */
void savgol_destroy(SavgolFilter *filter);

/* ref: savgolFilter.h:152-153 / src/savgolFilter.c:743-804.
 * output[j] = (1/dt^d) * sum_k w[k] * input[j-n+k]; edges per config.boundary.
 * Returns 0, or -1 (+ stderr) for NULL arguments or length < 2n+1.
 * output == input is allowed and yields the out-of-place result (the reference's own
 * in-place result is order-dependent; see DESIGN.md "in-place").  Thread-safe. */
int savgol_apply(const SavgolFilter *filter, const float *input, float *output, size_t length);

/* ref: savgolFilter.h:181-184 / src/savgolFilter.c:877-934.  Element i lives at
 * base + i*stride + offset (bytes).  Always polynomial edges (as the reference).
 * Returns 0, or -1 silently for NULL / count < 2n+1. */
int savgol_apply_strided(const SavgolFilter *filter,
                         const void *input, size_t in_stride, size_t in_offset,
                         void *output, size_t out_stride, size_t out_offset,
                         size_t count);

/* ref: savgolFilter.h:201-203 / src/savgolFilter.c:821-850.  Interior only:
 * output[j] is centred on input[j+n]; returns input_length-2n, or 0 on error. */
size_t savgol_apply_valid(const SavgolFilter *filter,
                          const float *input, size_t input_length, float *output);

/* ref: savgolFilter.h:210-222 */
#define SAVGOL_SMOOTH(half_win, order) \
    (SavgolConfig){ .half_window = (half_win), .poly_order = (order), .derivative = 0, \
                    .time_step = 1.0f, .boundary = SAVGOL_BOUNDARY_POLYNOMIAL }
#define SAVGOL_DERIV1(half_win, order, dt) \
    (SavgolConfig){ .half_window = (half_win), .poly_order = (order), .derivative = 1, \
                    .time_step = (dt), .boundary = SAVGOL_BOUNDARY_POLYNOMIAL }
#define SAVGOL_DERIV2(half_win, order, dt) \
    (SavgolConfig){ .half_window = (half_win), .poly_order = (order), .derivative = 2, \
                    .time_step = (dt), .boundary = SAVGOL_BOUNDARY_POLYNOMIAL }

/* ======================================================================== */
/* Part 1b -- single-channel stream (ref: include/iterative/savgol_stream.h) */
/* ======================================================================== */

/* ref: savgol_stream.h:29-37.  sizeof == 296, caller-allocatable, so layout is ABI.
 * The scalar one-sample-at-a-time API is host arithmetic by nature (one sample in,
 * at most n+1 samples out); the GPU path for streaming is savgol_mcstream_* below. */
typedef struct SavgolStream {
    const SavgolFilter *filter;
    float buffer[SAVGOL_MAX_WINDOW];
    int write_pos;
    size_t samples_received;
    size_t samples_output;
    bool owns_filter;
    float dt_inv;
} SavgolStream;

SavgolStream *savgol_stream_create(const SavgolConfig *config);              /* ref: savgol_stream.h:49  */
int savgol_stream_init(SavgolStream *stream, const SavgolFilter *filter);     /* ref: savgol_stream.h:58  */
void savgol_stream_destroy(SavgolStream *stream);                             /* ref: savgol_stream.h:64  */
void savgol_stream_reset(SavgolStream *stream);                               /* ref: savgol_stream.h:70  */
float savgol_stream_push(SavgolStream *stream, float sample, bool *output_valid);           /* :84     */
int savgol_stream_push_full(SavgolStream *stream, float sample, float *output, int max_outputs); /* :95-96 */
int savgol_stream_flush(SavgolStream *stream, float *output, int max_count);                /* :106    */
int savgol_stream_flush_leading(SavgolStream *stream, float *output, int max_count);        /* :116    */
bool savgol_stream_ready(const SavgolStream *stream);                                       /* :122    */
size_t savgol_stream_latency(const SavgolStream *stream);                                   /* :123    */
size_t savgol_stream_buffered(const SavgolStream *stream);                                  /* :124    */
size_t savgol_stream_samples_received(const SavgolStream *stream);                          /* :125    */
size_t savgol_stream_samples_output(const SavgolStream *stream);                            /* :126    */

/* ======================================================================== */
/* Part 1c -- 2D filter        (ref: include/iterative/savgol2d.h)           */
/* ======================================================================== */

#define SAVGOL2D_MAX_HALF_WINDOW 16 /* ref: savgol2d.h:64 */
#define SAVGOL2D_MAX_POLY_ORDER 6   /* ref: savgol2d.h:67 */
#define SAVGOL2D_MAX_TERMS 28       /* ref: savgol2d.h:70 */
#define SAVGOL2D_MAX_WINDOW_AREA ((2 * SAVGOL2D_MAX_HALF_WINDOW + 1) * (2 * SAVGOL2D_MAX_HALF_WINDOW + 1))

/* ref: savgol2d.h:82-90.  sizeof == 16. */
typedef struct {
    uint8_t half_window_x; /* columns */
    uint8_t half_window_y; /* rows    */
    uint8_t poly_order;    /* total degree: terms x^i y^j with i+j <= order */
    uint8_t deriv_x;
    uint8_t deriv_y;
    float delta_x;
    float delta_y;
} Savgol2DConfig;

/* ref: savgol2d.h:95-103.  sizeof == 48; `weights` is public (tests read it). */
typedef struct Savgol2DFilter {
    Savgol2DConfig config;
    int window_width;
    int window_height;
    int window_area;
    int num_terms;
    float scale;    /* 1 / (delta_x^dx * delta_y^dy) */
    float *weights; /* [window_height][window_width], host memory */
} Savgol2DFilter;

/* ref: savgol2d.h:108-112 */
typedef enum {
    SAVGOL2D_BOUNDARY_VALID = 0, /* interior written at offset (ny,nx); border untouched */
    SAVGOL2D_BOUNDARY_CONSTANT,  /* clamp to edge pixel  */
    SAVGOL2D_BOUNDARY_REFLECT    /* half-sample symmetric */
} Savgol2DBoundary;

Savgol2DFilter *savgol2d_create(const Savgol2DConfig *config); /* ref: savgol2d.h:126 */
void savgol2d_destroy(Savgol2DFilter *filter);                  /* ref: savgol2d.h:132 */

/* ref: savgol2d.h:154-156 / src/savgol2d.c:356-396.  Strides in elements.
 * 0 ok; -1 for NULL or a non-positive valid size. */
int savgol2d_apply_valid(const Savgol2DFilter *filter,
                         const float *input, int rows, int cols, int in_stride,
                         float *output, int out_stride);

/* ref: savgol2d.h:171-174 / src/savgol2d.c:398-456.
 * Non-finite pixels: the default arithmetic evaluates the window as a sum of separable factors (rank-R or
 * additive), so an Inf / NaN pixel contaminates the same outputs as in the reference (every output whose window
 * holds it) -- except for rectangular windows, which run with the shorter factor zero-padded to the longer
 * half-window: an Inf / NaN up to |hx - hy| pixels outside the true window is multiplied by 0 and yields NaN where
 * the reference stays finite.  savgol_b200_set_exact(1) runs the literal window and has the reference's footprint. */
int savgol2d_apply(const Savgol2DFilter *filter,
                   const float *input, int rows, int cols, int in_stride,
                   float *output, int out_stride, Savgol2DBoundary boundary);

/* ref: savgol2d.h:195-241 / src/savgol2d.c:462-618.  Host-pointer or device-pointer images.  Every requested
 * component equals savgol2d_apply with that component's filter (the reference's composition) bit for bit, but the
 * image is read once: device images with half-windows <= 8 run all components in ONE multi-output launch
 * (others: concurrent per-component launches), host images are uploaded once.  Buffers that overlap each other fall
 * back to the reference's component-by-component order.  The filters of recent configurations are cached.
 * The Laplacian runs as one filter whose table is Wxx/dx^2 + Wyy/dy^2 (exact flavour: the reference's two filters
 * and an add). */
int savgol2d_gradient(int half_win_x, int half_win_y, int poly_order,
                      const float *input, int rows, int cols, int stride,
                      float *grad_x, float *grad_y,
                      float delta_x, float delta_y, Savgol2DBoundary boundary);
int savgol2d_hessian(int half_win_x, int half_win_y, int poly_order,
                     const float *input, int rows, int cols, int stride,
                     float *hess_xx, float *hess_xy, float *hess_yy,
                     float delta_x, float delta_y, Savgol2DBoundary boundary);
/* Would the gradient (hessian = 0) / Hessian (1) of this configuration run as ONE multi-output launch (device images,
 * non-overlapping equally aligned outputs, full-size boundary, default arithmetic)?  1 yes, 0 no, -1 invalid.  Host logic. */
int savgol2d_b200_wrapper_plan(int half_win_x, int half_win_y, int poly_order, int hessian);
int savgol2d_laplacian(int half_win_x, int half_win_y, int poly_order,
                       const float *input, int rows, int cols, int stride,
                       float *output,
                       float delta_x, float delta_y, Savgol2DBoundary boundary);

/* ref: savgol2d.h:250-264 */
static inline void savgol2d_valid_size(const Savgol2DFilter *filter, int in_rows, int in_cols,
                                       int *out_rows, int *out_cols)
{
    *out_rows = in_rows - 2 * filter->config.half_window_y;
    *out_cols = in_cols - 2 * filter->config.half_window_x;
}
static inline int savgol2d_num_terms(int poly_order)
{
    return (poly_order + 1) * (poly_order + 2) / 2;
}
bool savgol2d_config_valid(const Savgol2DConfig *config); /* ref: savgol2d.h:269 */

/* ======================================================================== */
/* Part 2 -- extensions (no reference counterpart; same style)               */
/* ======================================================================== */

/* Library / device control ------------------------------------------------ */

/* ABI version of this header (major*100+minor). */
int savgol_b200_version(void);
/* 1 when a CUDA device of compute capability 10.x is usable in this process. */
int savgol_b200_device_ok(void);
/* All subsequent launches of the calling thread go to `cuda_stream` (a cudaStream_t
 * cast to void*; NULL = the legacy default stream).  Device-pointer calls are then
 * asynchronous with respect to the host; host-pointer calls always complete before
 * returning. */
void savgol_b200_set_stream(void *cuda_stream);
void *savgol_b200_get_stream(void);
/* Number of kernel launches issued by this library in this process so far. */
unsigned long long savgol_b200_launch_count(void);
/* ... of which launches of the bulk-tensor (TMA) 1D kernels (contiguous 16-byte aligned rows of >= 1024
 * samples, default arithmetic); SAVGOL_B200_NO_TMA=1 in the environment routes them to the cp.async kernels. */
unsigned long long savgol_b200_tma_launch_count(void);
/* Experiment / test switch for the above (process-wide): 0 never, 1 (default) where they measured faster
 * (half-windows up to 17, launches of >= 2048 segments), 2 wherever the data layout allows. */
void savgol_b200_set_tma(int how);
/* Floats per staging chunk of a host-pointer call that moves `total_floats` each way (pinned or pageable buffers):
 * the configured chunk (64 MiB / 16 MiB), shrunk to ~1/8 (1/4) of calls of 8 MiB and more, never below 2 MiB. */
size_t savgol_b200_staging_chunk(size_t total_floats, int pageable);
/* Host-side building block of the pageable-memory path, exported for the CPU test-suite: copies `rows` rows of
 * `width` BYTES (pitches in bytes) with the library's host copy pool (streaming stores, several threads); returns
 * the number of threads a large copy uses (SAVGOL_B200_COPY_THREADS).  Needs no GPU. */
int savgol_b200_host_copy2d(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width, size_t rows);
/* How the library would dispatch a 1D launch of `rows` contiguous-sample rows of `length` floats with `row_pitch`
 * floats between rows, the first sample `first_sample_offset` floats behind a 16-byte boundary -- pure host logic, no
 * GPU needed (CPU test-suite, tools).  *family: 0 generic cp.async kernel, 1 short-row kernel (several rows per warp),
 * 2 bulk-tensor (TMA) kernel; *lanes_per_row: short-row kernel only; *phase: rows are cut on a per-row alignment
 * phase; *tail: the last full segment also produces the <= 32 outputs behind it; *segments_per_row: 1024-output work
 * items per row.  Returns the internal instantiation index (>= 0) or -1 on bad arguments. */
int savgol_b200_plan_1d(int half_window, int stream_variant, int exact, size_t rows, size_t length, size_t row_pitch,
                        size_t first_sample_offset, int polynomial_edges, int *family, int *lanes_per_row, int *phase,
                        int *tail, long long *segments_per_row);
/* Arithmetic flavour: 0 (default) = FMA chains, within 1e-6*max|x|/dt^d of the
 * reference; 1 = "exact": the reference's own summation order with unfused
 * multiply/add, bit-identical to the reference C code (slower; for verification). */
void savgol_b200_set_exact(int exact);          /* the calling host thread's flavour */
void savgol_b200_set_exact_default(int exact);  /* process default, for threads that never chose */
int savgol_b200_get_exact(void);
/* In-place semantics.  Default 0: savgol_apply(f, x, x, L) returns the out-of-place result (alias safe).  1: exactly
 * aliased savgol_apply / savgol_apply_batch calls of the calling host thread reproduce what the reference really
 * computes in place -- its centre loop reads samples it has already overwritten (src/savgolFilter.c:763-801), a
 * recursive result that depends on the loop order -- bit for bit, with a sequential one-thread-per-signal kernel
 * (slow; for callers whose downstream numbers were fitted to that behaviour). */
void savgol_b200_set_inplace_compat(int on);
int savgol_b200_get_inplace_compat(void);

/* 1D batch ---------------------------------------------------------------- */

/* `n_signals` independent signals of `length` samples; signal r starts at
 * input + r*in_pitch (pitches in elements).  Equivalent to the caller's loop over
 * savgol_apply() (the reference's only notion of a batch), one launch.
 * Returns 0 / -1 like savgol_apply(). */
int savgol_apply_batch(const SavgolFilter *filter, const float *input, float *output,
                       size_t n_signals, size_t length, size_t in_pitch, size_t out_pitch);

/* One contiguous slice [0,length) of a longer signal whose neighbours live elsewhere
 * (another GPU, another buffer).  left_halo holds the n samples preceding input[0]
 * (chronological order), right_halo the n samples following input[length-1]; either
 * may be NULL, in which case that side is a true signal end and is treated per
 * config.boundary (PERIODIC is not allowed with a NULL halo on only one side).
 * This is the per-GPU piece of a partitioned long signal: the caller exchanges the
 * n-sample halos (NVLink P2P / NCCL) and every slice is then independent. */
int savgol_apply_halo(const SavgolFilter *filter, const float *input, float *output, size_t length,
                      const float *left_halo, const float *right_halo);

/* Peer memory for the partitioned signal (one process per GPU) -------------- */

/* Multi-GPU from one process ---------------------------------------------- */

/* SURVEY.md 8b rule (4) / 8e: the reference's user is a C caller looping over savgol_apply
 * (ref: include/iterative/savgolFilter.h:16-19, 130-203); these calls hand that caller every GPU of the
 * box without a process group.
 *
 * savgol_apply_batch_multi: HOST buffers.  `devices[0..n_devices)` are CUDA device ordinals.  With at least
 * as many signals as devices the signals are sharded in contiguous blocks (no communication); otherwise
 * (one very long signal) every signal is partitioned along its length and each slice is staged with its
 * n-sample halos from the host signal.  One staging pipeline per device, all running concurrently.
 * Result and return codes are those of savgol_apply_batch. */
int savgol_apply_batch_multi(const SavgolFilter *filter, const float *input, float *output,
                             size_t n_signals, size_t length, size_t in_pitch, size_t out_pitch,
                             const int *devices, int n_devices);

/* savgol_apply_slices: one signal that already lives on several GPUs as consecutive slices, slice i =
 * in_slices[i][0..lengths[i]) in the memory of devices[i] (the same device may appear more than once).
 * Peer access is enabled between ring neighbours and each device's kernel reads its 2 x half_window halo
 * samples directly from the neighbour's memory over NVLink -- no copy, no collective, one launch per device.
 * PERIODIC wraps from the last slice to the first; the other modes treat the outer ends as signal ends.
 * The inputs must be complete before the call; it returns when every output slice is complete.  0 / -1. */
int savgol_apply_slices(const SavgolFilter *filter, const float *const *in_slices, float *const *out_slices,
                        const size_t *lengths, const int *devices, int n_slices);

/* savgol2d_apply_batch_multi: HOST images sharded over the devices in contiguous blocks (independent images: no
 * communication), one staging pipeline per device.  Arguments and result as savgol2d_apply_batch. */
int savgol2d_apply_batch_multi(const Savgol2DFilter *filter,
                               const float *input, int rows, int cols, int in_stride, size_t in_image_pitch,
                               float *output, int out_stride, size_t out_image_pitch,
                               size_t n_images, Savgol2DBoundary boundary,
                               const int *devices, int n_devices);

/* Device memory for callers without CUDA headers (examples/c_multi_gpu.c). */
int savgol_b200_device_count(void);
void *savgol_b200_alloc(int device, size_t bytes);
void savgol_b200_free(int device, void *ptr);
int savgol_b200_copy(void *dst, const void *src, size_t bytes);   /* any direction, synchronous; 0 / -1 */

/* The halo pointers of savgol_apply_halo() may point into ANOTHER GPU's memory: the kernel
 * reads the 2n halo samples over NVLink while it stages the slice, so the "halo exchange"
 * needs no collective and no copy.  With one process per GPU the neighbour's slice is mapped
 * through CUDA IPC: the owner exports (handle, offset) for its slice pointer, ships the
 * SAVGOL_B200_IPC_HANDLE_BYTES + the offset to the neighbour by any means (torch.distributed,
 * MPI, a pipe), the neighbour opens it once and keeps the mapping for the life of the buffer.
 * The caller orders producer and consumer (the neighbour's samples must be complete before the
 * launch that reads them).  export/close return 0 or -1, open returns NULL on error. */
#define SAVGOL_B200_IPC_HANDLE_BYTES 64
int savgol_b200_ipc_export(const void *dev_ptr, void *handle64, size_t *offset);
void *savgol_b200_ipc_open(const void *handle64, size_t offset);
int savgol_b200_ipc_close(void *mapped, size_t offset);

/* 2D batch ---------------------------------------------------------------- */

/* `n_images` images, image i at input + i*in_image_pitch (elements). */
int savgol2d_apply_batch(const Savgol2DFilter *filter,
                         const float *input, int rows, int cols, int in_stride, size_t in_image_pitch,
                         float *output, int out_stride, size_t out_image_pitch,
                         size_t n_images, Savgol2DBoundary boundary);

/* Diagnostics: the separable factorisation the fast 2D kernel uses for this filter.  *rank = number of
 * (column factor x row factor) terms, 0 if the weight table was not accepted (every apply then runs the
 * literal window kernel); *sum_err = sum |W - sum_r col_r x row_r| over the table.  Host only.  0 / -1. */
int savgol2d_b200_plan(const Savgol2DFilter *filter, int *rank, float *sum_err);
/* Which kernel family the plan selects: 0 literal window (sg2d_direct.cu), 1 rank-R separable factors
 * (sg2d_sep.cu), 2 additive surface W = u(x) + v(y) with box-sum factors (sg2d_add.cu; every order-2/3
 * smoothing filter with a square window up to 17x17).  -1 when `filter` was not made by savgol2d_create. */
int savgol2d_b200_plan_kind(const Savgol2DFilter *filter);

/* One horizontal band of a larger image (row-band sharding of a single huge image over several
 * GPUs; device pointers).  `input` is a buffer of `rows` rows: `top_halo` rows that precede the
 * band in the image, the band itself, `bottom_halo` rows that follow it.  A halo count is either
 * half_window_y (the neighbour's rows, fetched by the caller) or 0, meaning that side of the band
 * is the image border and follows `boundary` (CONSTANT or REFLECT).  The rows - top_halo -
 * bottom_halo filtered band rows go to `output` (row 0 = first band row); they are bit-identical
 * to the same rows of savgol2d_apply() on the whole image.  Returns 0 / -1. */
int savgol2d_apply_band(const Savgol2DFilter *filter,
                        const float *input, int rows, int cols, int in_stride,
                        float *output, int out_stride, Savgol2DBoundary boundary,
                        int top_halo, int bottom_halo);
/* The same with the position of the band in its image: `image_row0` = image row of the first row of `input`
 * (halo rows included).  The additive 2D kernel sums an output row's terms in an order that depends on the
 * parity of its image row; telling it where the band sits makes every pixel round exactly as in a whole-image
 * call, so bands computed by different GPUs reassemble the whole-image result bit for bit. */
int savgol2d_apply_band_at(const Savgol2DFilter *filter,
                           const float *input, int rows, int cols, int in_stride,
                           float *output, int out_stride, Savgol2DBoundary boundary,
                           int top_halo, int bottom_halo, int image_row0);

/* Multi-channel chunked stream ------------------------------------------- */

/* `channels` independent streams advanced in lockstep, one chunk of K samples per
 * channel per call.  Per channel the semantics are exactly those of the reference's
 * savgol_stream_push_full()/savgol_stream_flush() sequence (src/savgol_stream.c:180-252):
 * fixed latency of half_window samples, polynomial leading/trailing edges, boundary
 * mode ignored.  Carry state (the last 2n samples per channel) lives in device memory. */
typedef struct SavgolMCStream SavgolMCStream;

SavgolMCStream *savgol_mcstream_create(const SavgolConfig *config, size_t channels);
void savgol_mcstream_destroy(SavgolMCStream *s);
void savgol_mcstream_reset(SavgolMCStream *s);

/* Pushes chunk_len samples per channel (channel c at input + c*in_pitch) and writes
 * that channel's new outputs, in chronological order, to output + c*out_pitch.
 * Returns the number of outputs produced per channel: 0 while fewer than 2n+1 samples
 * have been received, received-n on the call that crosses 2n+1, chunk_len afterwards;
 * -1 on error.  out_pitch must be >= chunk_len + half_window. */
long long savgol_mcstream_push(SavgolMCStream *s, const float *input, size_t in_pitch,
                               size_t chunk_len, float *output, size_t out_pitch);

/* Writes the final half_window outputs per channel (trailing edge) to
 * output + c*out_pitch; returns half_window, 0 if the window never filled, -1 on error.
 * Does not consume state (as the reference's flush). */
long long savgol_mcstream_flush(SavgolMCStream *s, float *output, size_t out_pitch);

size_t savgol_mcstream_channels(const SavgolMCStream *s);
size_t savgol_mcstream_latency(const SavgolMCStream *s);          /* == half_window */
size_t savgol_mcstream_samples_received(const SavgolMCStream *s); /* per channel    */
size_t savgol_mcstream_samples_output(const SavgolMCStream *s);   /* per channel    */
/* Device pointer / element count of the carry state ([channels][2n] floats) for
 * checkpointing. */
float *savgol_mcstream_state(SavgolMCStream *s, size_t *n_floats);

/* Checkpoint / resume of a long-running stream (host blob: counters + carry state).
 * save() returns the bytes written (== checkpoint_size) or -1; restore() returns 0, or -1
 * when the blob was written by a stream with another configuration / channel count.
 * A restored stream continues bit-identically to the one that was saved. */
size_t savgol_mcstream_checkpoint_size(const SavgolMCStream *s);
long long savgol_mcstream_save(SavgolMCStream *s, void *blob, size_t capacity);
int savgol_mcstream_restore(SavgolMCStream *s, const void *blob, size_t bytes);

#ifdef __cplusplus
}
#endif

#endif /* SAVGOL_B200_H */
