// capi_2d.cu -- C ABI of the 2D filter (include/savgol_b200.h part 1c + batch extension).
// Argument checks / return codes follow src/savgol2d.c:271-618.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_set>
#include <vector>

#include "coeffs.h"
#include "engine.h"
#include "sg2d.h"

using sge::cuda_ok;
using sge::MemKind;

namespace {

struct Filter2DImpl {
    Savgol2DFilter pub;  // public ABI prefix (ref: include/iterative/savgol2d.h:95-103)
    uint64_t magic;
    float* d_weights[sge::kMaxDevices];
    sg2d::SepPlan plan;  // separable / additive factorisation of the weight surface (factor2d.cpp)
};
constexpr uint64_t kMagic2D = 0x5347423230303244ULL;

std::mutex g_mu2d;
std::unordered_set<const void*> g_live2d;

Filter2DImpl* live2d(const Savgol2DFilter* f)
{
    std::lock_guard<std::mutex> lk(g_mu2d);
    if (!g_live2d.count(f)) return nullptr;
    Filter2DImpl* fi = reinterpret_cast<Filter2DImpl*>(const_cast<Savgol2DFilter*>(f));
    return fi->magic == kMagic2D ? fi : nullptr;
}

// device copy of the weights on the current device (*temp set when it is a per-call upload)
const float* weights_device(const Savgol2DFilter* f, cudaStream_t st, float** temp)
{
    *temp = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    Filter2DImpl* fi = live2d(f);
    const size_t bytes = static_cast<size_t>(f->window_area) * sizeof(float);
    if (fi && dev < sge::kMaxDevices) {
        std::lock_guard<std::mutex> lk(g_mu2d);
        if (!fi->d_weights[dev]) {
            float* d = nullptr;
            if (!cuda_ok(cudaMalloc(&d, bytes), "cudaMalloc(2d weights)")) return nullptr;
            if (!cuda_ok(cudaMemcpy(d, f->weights, bytes, cudaMemcpyHostToDevice), "upload 2d weights")) { cudaFree(d); return nullptr; }
            fi->d_weights[dev] = d;
        }
        return fi->d_weights[dev];
    }
    float* d = nullptr;
    if (!cuda_ok(cudaMallocAsync(&d, bytes, st), "cudaMallocAsync(2d weights)")) return nullptr;
    if (!cuda_ok(cudaMemcpyAsync(d, f->weights, bytes, cudaMemcpyHostToDevice, st), "upload 2d weights")) { cudaFreeAsync(d, st); return nullptr; }
    *temp = d;
    return d;
}

// One launch over device-resident images.
bool run2d_device(const Savgol2DFilter* f, const float* in, int rows, int cols, long long is, long long ipitch,
                  float* out, long long os, long long opitch, long long n_images, int boundary, cudaStream_t st,
                  int top_halo = -1, int bottom_halo = -1, int image_row0 = 0)
{
    const int nx = f->config.half_window_x, ny = f->config.half_window_y;
    sg2d::Args2D a{};
    a.in = in; a.out = out;
    a.rows = rows; a.cols = cols;
    a.nx = nx; a.ny = ny;
    a.in_stride = is; a.out_stride = os;
    a.in_image_pitch = ipitch; a.out_image_pitch = opitch;
    a.n_images = n_images;
    a.boundary = boundary;
    a.scale = f->scale;
    if (top_halo >= 0) {
        // band of a larger image: the buffer holds halo rows above / below, only the rows in between are
        // produced; a side without halo is a true image border and follows the boundary rule
        a.out_rows = rows - top_halo - bottom_halo; a.out_cols = cols;
        a.cy = top_halo; a.cx = 0;
        a.row0 = image_row0;
    } else if (boundary == sg2d::B_VALID) {
        a.out_rows = rows - 2 * ny; a.out_cols = cols - 2 * nx;
        a.cy = ny; a.cx = nx;
    } else {
        a.out_rows = rows; a.out_cols = cols;
        a.cy = a.cx = 0;
    }
    const bool exact = sge::exact_mode() != 0;
    Filter2DImpl* fi = live2d(f);
    if (!exact && fi && sg2d::separable_supported(a, fi->plan)) {
        return cuda_ok(sg2d::launch_separable(a, fi->plan, st), "sg2d separable launch");
    }
    float* temp = nullptr;
    a.weights = weights_device(f, st, &temp);
    if (!a.weights) return false;
    const bool ok = cuda_ok(sg2d::launch_direct(a, exact, st), "sg2d direct launch");
    if (temp) cudaFreeAsync(temp, st);
    return ok;
}

bool ranges_overlap2d(const float* a, size_t an, const float* b, size_t bn) { return a < b + bn && b < a + an; }

// Images anywhere (host or device).  `valid_origin`: VALID writes its first output at out[0]
// (savgol2d_apply_valid) instead of out[ny*os+nx] (savgol2d_apply with BOUNDARY_VALID).
int apply2d_any(const Savgol2DFilter* f, const float* in, int rows, int cols, int is, size_t ipitch,
                float* out, int os, size_t opitch, size_t n_images, int boundary)
{
    sge::DeviceGuard guard(in);
    if (!sge::device_ready(true)) return -1;
    const int nx = f->config.half_window_x, ny = f->config.half_window_y;
    const int orows = boundary == sg2d::B_VALID ? rows - 2 * ny : rows;
    const int ocols = boundary == sg2d::B_VALID ? cols - 2 * nx : cols;
    const MemKind ki = sge::classify(in), ko = sge::classify(out);
    cudaStream_t st = sge::current_stream();
    if (ki == MemKind::Device && ko == MemKind::Device) {
        const size_t in_span = (n_images - 1) * ipitch + static_cast<size_t>(rows - 1) * is + cols;
        const size_t out_span = (n_images - 1) * opitch + static_cast<size_t>(orows - 1) * os + ocols;
        if (ranges_overlap2d(in, in_span, out, out_span)) {
            // aliased: compute into scratch, copy back (out-of-place result, like the 1D path)
            float* scratch = nullptr;
            const size_t img = static_cast<size_t>(orows) * ocols;
            if (!cuda_ok(cudaMallocAsync(&scratch, n_images * img * sizeof(float), st), "cudaMallocAsync(2d in-place)")) return -1;
            bool ok = run2d_device(f, in, rows, cols, is, static_cast<long long>(ipitch), scratch, ocols,
                                   static_cast<long long>(img), static_cast<long long>(n_images), boundary, st);
            for (size_t i = 0; ok && i < n_images; ++i)
                ok = cuda_ok(cudaMemcpy2DAsync(out + i * opitch, static_cast<size_t>(os) * sizeof(float), scratch + i * img,
                                               static_cast<size_t>(ocols) * sizeof(float), static_cast<size_t>(ocols) * sizeof(float),
                                               orows, cudaMemcpyDeviceToDevice, st), "2d copy back");
            cudaFreeAsync(scratch, st);
            return ok ? 0 : -1;
        }
        return run2d_device(f, in, rows, cols, is, static_cast<long long>(ipitch), out, os, static_cast<long long>(opitch),
                            static_cast<long long>(n_images), boundary, st) ? 0 : -1;
    }
    if (ki == MemKind::Device || ko == MemKind::Device) {
        fprintf(stderr, "savgol_b200: input and output must both be device pointers or both be host pointers\n");
        return -1;
    }
    // host images: one image per pipeline slot, H2D / kernel / D2H overlapped across images
    sge::PipeLease lease;
    if (!lease.ok()) return -1;
    sge::Pipeline& P = *lease;
    // in place / overlapping host images: image i's D2H may run while image i+1's H2D is still reading -- stage the
    // input of overlapping calls from a copy (images that coincide exactly are safe: one image per slot)
    std::vector<float> aside;
    {
        const size_t in_span = (n_images - 1) * ipitch + static_cast<size_t>(rows - 1) * is + cols;
        const size_t out_span = (n_images - 1) * opitch + static_cast<size_t>(orows - 1) * os + ocols;
        if (ranges_overlap2d(in, in_span, out, out_span) && !(in == out && is == os && ipitch == opitch && boundary != sg2d::B_VALID)) {
            aside.assign(in, in + in_span);
            in = aside.data();
        }
    }
    const size_t img_in = static_cast<size_t>(rows) * cols, img_out = static_cast<size_t>(orows) * ocols;
    P.begin(in, out);
    // A few LARGE images (the reference's call: ONE image) give an image-per-slot pipeline nothing to overlap: upload,
    // kernel and download would run one after the other.  They are cut into row bands with ny halo rows instead --
    // the band entry point of the multi-GPU path, bit-identical to the whole-image result -- and the bands travel
    // through the slots.  Measured on B200 (tools/r2_host_small.py, profiles/r2_host_small.txt): one pinned 4096 x 4096
    // image 2.47 -> 1.94 ms, pageable 4.35 -> 3.29 ms; three images already overlap image by image.
    static const bool no_bands = [] { const char* e = getenv("SAVGOL_B200_NO_HOST_BANDS"); return e && e[0] == '1'; }();
    const size_t band_floats = std::min<size_t>(sge::chunk_floats(P.bounce_in || P.bounce_out), size_t(8) << 18);   // ~8 MiB bands
    if (!no_bands && boundary != sg2d::B_VALID && sge::exact_mode() == 0 && n_images < 3 && img_in >= 2 * band_floats &&
        aside.empty() && !ranges_overlap2d(in, (n_images - 1) * ipitch + static_cast<size_t>(rows - 1) * is + cols, out,
                                           (n_images - 1) * opitch + static_cast<size_t>(rows - 1) * os + cols)) {
        int rows_b = static_cast<int>(std::max<size_t>(band_floats / cols, 1));
        rows_b = std::max(rows_b & ~1, std::max(2 * ny, 16));            // even: the additive kernel pairs even image rows
        const int nb = std::max(1, rows / rows_b);                       // the last band takes the remainder (< 2 rows_b rows)
        if (nb >= 2) {
            const size_t cap_rows = static_cast<size_t>(2 * rows_b + 2 * ny);
            if (!P.ensure(cap_rows * cols, static_cast<size_t>(2 * rows_b) * cols)) return -1;
            size_t u = 0;
            for (size_t i = 0; i < n_images; ++i)
                for (int b = 0; b < nb; ++b, ++u) {
                    const int s = static_cast<int>(u % sge::Pipeline::kSlots);
                    const int r0 = b * rows_b, r1 = b + 1 == nb ? rows : r0 + rows_b;
                    const int top = r0 > 0 ? ny : 0, bot = r1 < rows ? ny : 0, nbuf = r1 - r0 + top + bot;
                    if (!P.reuse(s)) return -1;
                    if (!P.h2d(s, P.d_in[s], cols, in + i * ipitch + static_cast<size_t>(r0 - top) * is, is, cols, nbuf)) return -1;
                    cudaEventRecord(P.e_in[s], P.s_in);
                    cudaStreamWaitEvent(P.s_k, P.e_in[s], 0);
                    if (u >= sge::Pipeline::kSlots) cudaStreamWaitEvent(P.s_k, P.e_out[s], 0);
                    if (!run2d_device(f, P.d_in[s], nbuf, cols, cols, 0, P.d_out[s], cols, 0, 1, boundary, P.s_k, top, bot, r0 - top)) return -1;   // image row of the buffer's first row
                    cudaEventRecord(P.e_k[s], P.s_k);
                    cudaStreamWaitEvent(P.s_out, P.e_k[s], 0);
                    if (!P.d2h(s, out + i * opitch + static_cast<size_t>(r0) * os, os, P.d_out[s], cols, cols, r1 - r0)) return -1;
                    cudaEventRecord(P.e_out[s], P.s_out);
                }
            return P.finish() ? 0 : -1;
        }
    }
    if (!P.ensure(img_in, img_out)) return -1;
    for (size_t i = 0; i < n_images; ++i) {
        const int s = static_cast<int>(i % sge::Pipeline::kSlots);
        if (!P.reuse(s)) return -1;
        if (!P.h2d(s, P.d_in[s], cols, in + i * ipitch, is, cols, rows)) return -1;
        cudaEventRecord(P.e_in[s], P.s_in);
        cudaStreamWaitEvent(P.s_k, P.e_in[s], 0);
        if (i >= sge::Pipeline::kSlots) cudaStreamWaitEvent(P.s_k, P.e_out[s], 0);
        if (!run2d_device(f, P.d_in[s], rows, cols, cols, 0, P.d_out[s], ocols, 0, 1, boundary, P.s_k)) return -1;
        cudaEventRecord(P.e_k[s], P.s_k);
        cudaStreamWaitEvent(P.s_out, P.e_k[s], 0);
        if (!P.d2h(s, out + i * opitch, os, P.d_out[s], ocols, ocols, orows)) return -1;
        cudaEventRecord(P.e_out[s], P.s_out);
    }
    return P.finish() ? 0 : -1;
}

}  // namespace

extern "C" {

bool savgol2d_config_valid(const Savgol2DConfig* c)
{
    if (!c) return false;
    return sgc::config2d_valid(c->half_window_x, c->half_window_y, c->poly_order, c->deriv_x, c->deriv_y, c->delta_x, c->delta_y);
}

Savgol2DFilter* savgol2d_create(const Savgol2DConfig* config)
{
    if (!savgol2d_config_valid(config)) {
        fprintf(stderr, "savgol2d_create: invalid configuration\n");
        return nullptr;
    }
    Filter2DImpl* fi = static_cast<Filter2DImpl*>(calloc(1, sizeof(Filter2DImpl)));
    if (!fi) return nullptr;
    Savgol2DFilter* f = &fi->pub;
    f->config = *config;
    f->window_width = 2 * config->half_window_x + 1;
    f->window_height = 2 * config->half_window_y + 1;
    f->window_area = f->window_width * f->window_height;
    f->num_terms = savgol2d_num_terms(config->poly_order);
    f->scale = sgc::scale2d(config->deriv_x, config->deriv_y, config->delta_x, config->delta_y);
    f->weights = static_cast<float*>(malloc(static_cast<size_t>(f->window_area) * sizeof(float)));
    if (!f->weights) { free(fi); return nullptr; }
    double coef[28];
    if (!sgc::weights2d(config->half_window_x, config->half_window_y, config->poly_order, config->deriv_x, config->deriv_y,
                        f->weights, coef)) {
        fprintf(stderr, "savgol2d_create: weight computation failed\n");
        free(f->weights);
        free(fi);
        return nullptr;
    }
    sg2d::plan_separable(config->half_window_x, config->half_window_y, config->poly_order, coef, f->weights, &fi->plan);
    fi->magic = kMagic2D;
    {
        std::lock_guard<std::mutex> lk(g_mu2d);
        g_live2d.insert(fi);
    }
    return f;
}

void savgol2d_destroy(Savgol2DFilter* filter)
{
    if (!filter) return;
    Filter2DImpl* fi = live2d(filter);
    if (fi) {
        {
            std::lock_guard<std::mutex> lk(g_mu2d);
            g_live2d.erase(fi);
        }
        for (int d = 0; d < sge::kMaxDevices; ++d)
            if (fi->d_weights[d]) {
                int cur = 0;
                cudaGetDevice(&cur);
                cudaSetDevice(d);
                cudaFree(fi->d_weights[d]);
                cudaSetDevice(cur);
            }
        fi->magic = 0;
    }
    free(filter->weights);
    free(filter);
}

int savgol2d_b200_plan(const Savgol2DFilter* filter, int* rank, float* sum_err)
{
    Filter2DImpl* fi = filter ? live2d(filter) : nullptr;
    if (!fi) return -1;
    if (rank) *rank = fi->plan.rank;
    if (sum_err) *sum_err = fi->plan.sum_err;
    return 0;
}

int savgol2d_b200_plan_kind(const Savgol2DFilter* filter)
{
    Filter2DImpl* fi = filter ? live2d(filter) : nullptr;
    if (!fi) return -1;
    return fi->plan.rank < 1 ? 0 : fi->plan.additive ? 2 : 1;
}

int savgol2d_apply_band(const Savgol2DFilter* filter, const float* input, int rows, int cols, int in_stride, float* output,
                        int out_stride, Savgol2DBoundary boundary, int top_halo, int bottom_halo)
{
    return savgol2d_apply_band_at(filter, input, rows, cols, in_stride, output, out_stride, boundary, top_halo, bottom_halo, 0);
}

int savgol2d_apply_band_at(const Savgol2DFilter* filter, const float* input, int rows, int cols, int in_stride, float* output,
                           int out_stride, Savgol2DBoundary boundary, int top_halo, int bottom_halo, int image_row0)
{
    if (!filter || !input || !output) return -1;
    const int ny = filter->config.half_window_y;
    if (boundary != SAVGOL2D_BOUNDARY_CONSTANT && boundary != SAVGOL2D_BOUNDARY_REFLECT) {
        fprintf(stderr, "savgol2d_apply_band: boundary must be CONSTANT or REFLECT\n");
        return -1;
    }
    if ((top_halo != 0 && top_halo != ny) || (bottom_halo != 0 && bottom_halo != ny) || cols <= 0 ||
        rows - top_halo - bottom_halo <= 0) {
        fprintf(stderr, "savgol2d_apply_band: halos must be 0 (image border) or half_window_y rows, band must not be empty\n");
        return -1;
    }
    sge::DeviceGuard guard(input);
    if (!sge::device_ready(true)) return -1;
    if (sge::classify(input) != MemKind::Device || sge::classify(output) != MemKind::Device) {
        fprintf(stderr, "savgol2d_apply_band: band and output must be device pointers\n");
        return -1;
    }
    const size_t in_span = static_cast<size_t>(rows - 1) * in_stride + cols;
    const size_t out_span = static_cast<size_t>(rows - top_halo - bottom_halo - 1) * out_stride + cols;
    if (ranges_overlap2d(input, in_span, output, out_span)) {
        fprintf(stderr, "savgol2d_apply_band: output must not overlap the band buffer\n");
        return -1;
    }
    return run2d_device(filter, input, rows, cols, in_stride, 0, output, out_stride, 0, 1, static_cast<int>(boundary),
                        sge::current_stream(), top_halo, bottom_halo, image_row0) ? 0 : -1;
}

int savgol2d_apply_batch(const Savgol2DFilter* filter, const float* input, int rows, int cols, int in_stride,
                         size_t in_image_pitch, float* output, int out_stride, size_t out_image_pitch, size_t n_images,
                         Savgol2DBoundary boundary)
{
    if (!filter || !input || !output) return -1;
    if (n_images == 0) return 0;
    const int nx = filter->config.half_window_x, ny = filter->config.half_window_y;
    if (rows <= 0 || cols <= 0) return -1;
    if (boundary == SAVGOL2D_BOUNDARY_VALID) {
        if (rows - 2 * ny <= 0 || cols - 2 * nx <= 0) return -1;  // ref: src/savgol2d.c:371
        // the reference writes the valid block at offset (ny, nx) of `output`; border untouched
        return apply2d_any(filter, input, rows, cols, in_stride, in_image_pitch,
                           output + static_cast<ptrdiff_t>(ny) * out_stride + nx, out_stride, out_image_pitch, n_images,
                           sg2d::B_VALID);
    }
    const int b = boundary == SAVGOL2D_BOUNDARY_REFLECT ? sg2d::B_REFLECT : sg2d::B_CONSTANT;  // ref: :428-445 (else = clamp)
    return apply2d_any(filter, input, rows, cols, in_stride, in_image_pitch, output, out_stride, out_image_pitch, n_images, b);
}

int savgol2d_apply(const Savgol2DFilter* filter, const float* input, int rows, int cols, int in_stride,
                   float* output, int out_stride, Savgol2DBoundary boundary)
{
    return savgol2d_apply_batch(filter, input, rows, cols, in_stride, 0, output, out_stride, 0, 1, boundary);
}

int savgol2d_apply_valid(const Savgol2DFilter* filter, const float* input, int rows, int cols, int in_stride,
                         float* output, int out_stride)
{
    if (!filter || !input || !output) return -1;
    const int nx = filter->config.half_window_x, ny = filter->config.half_window_y;
    if (rows - 2 * ny <= 0 || cols - 2 * nx <= 0) return -1;
    return apply2d_any(filter, input, rows, cols, in_stride, 0, output, out_stride, 0, 1, sg2d::B_VALID);
}

// Convenience wrappers: one filter per requested component, as the reference composes them
// (src/savgol2d.c:462-618).  The reference re-creates the filter (design matrix, normal equations, Cholesky) on
// every call; here the filters of the most recent wrapper configurations stay alive in a small cache -- creation
// is host work in the hundreds of microseconds, more than the kernel takes on a megapixel image.
}  // extern "C"
namespace {
struct WrapKey {
    int hx, hy, order, dx, dy;
    float delta_x, delta_y;
    bool operator==(const WrapKey& o) const
    {
        return hx == o.hx && hy == o.hy && order == o.order && dx == o.dx && dy == o.dy && delta_x == o.delta_x && delta_y == o.delta_y;
    }
};
struct WrapEntry { WrapKey key; Savgol2DFilter* f; unsigned long long used; };
std::mutex g_wrap_mu;
std::vector<WrapEntry> g_wrap;   // filters are immutable and apply is thread-safe: entries are shared, never destroyed while cached
unsigned long long g_wrap_clock = 0;
constexpr size_t kWrapCache = 16;

// dx < 0 marks the fused Laplacian table (made by `make`)
template <class Make>
Savgol2DFilter* cached_filter(const WrapKey& key, Make make)
{
    {
        std::lock_guard<std::mutex> lk(g_wrap_mu);
        for (auto& e : g_wrap)
            if (e.key == key) { e.used = ++g_wrap_clock; return e.f; }
    }
    Savgol2DFilter* f = make();
    if (!f) return nullptr;
    std::lock_guard<std::mutex> lk(g_wrap_mu);
    for (auto& e : g_wrap)
        if (e.key == key) { e.used = ++g_wrap_clock; return e.f; }   // another thread was faster: both filters are equivalent, the
                                                                      // spare one stays alive (a few KB) for callers that already hold it
    if (g_wrap.size() >= kWrapCache) {
        size_t lru = 0;
        for (size_t i = 1; i < g_wrap.size(); ++i)
            if (g_wrap[i].used < g_wrap[lru].used) lru = i;
        g_wrap.erase(g_wrap.begin() + static_cast<long>(lru));   // evicted filters are not destroyed: a concurrent call may still use them
    }
    g_wrap.push_back({key, f, ++g_wrap_clock});
    return f;
}
}  // namespace
extern "C" {

static Savgol2DFilter* component_filter(int hx, int hy, int order, int dx, int dy, float delta_x, float delta_y)
{
    return cached_filter(WrapKey{hx, hy, order, dx, dy, delta_x, delta_y}, [&]() -> Savgol2DFilter* {
        Savgol2DConfig cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.half_window_x = static_cast<uint8_t>(hx);
        cfg.half_window_y = static_cast<uint8_t>(hy);
        cfg.poly_order = static_cast<uint8_t>(order);
        cfg.deriv_x = static_cast<uint8_t>(dx);
        cfg.deriv_y = static_cast<uint8_t>(dy);
        cfg.delta_x = delta_x;
        cfg.delta_y = delta_y;
        return savgol2d_create(&cfg);
    });
}

static int component(int hx, int hy, int order, int dx, int dy, const float* in, int rows, int cols, int stride, float* out,
                     float delta_x, float delta_y, Savgol2DBoundary boundary)
{
    Savgol2DFilter* f = component_filter(hx, hy, order, dx, dy, delta_x, delta_y);
    if (!f) return -1;
    return savgol2d_apply(f, in, rows, cols, stride, out, stride, boundary);
}

// The components of a gradient / Hessian share the input image and nothing else (different parities, different
// factors).  For device images, in order of preference:
//   1. ONE launch of the multi-output kernel (sg2d_multi.cu): the image is staged once, every component has its own
//      accumulator ring -- half-windows <= 8, full-size boundaries, default arithmetic;
//   2. CONCURRENT per-component launches: the first on the caller's stream, the others on side streams forked from
//      it and joined back (one 4096^2 image gives a launch only ~1.3 work items per resident warp, so two or three
//      launches together fill the machine, and the later reads of the image are served from L2);
//   3. the sequential composition of the reference (aliased buffers, the exact flavour).
// Host images are uploaded once (below) and then take the device path.
// Measured on B200 (tools/r2_wrappers.py, profiles/r2_wrappers.txt).  SAVGOL_B200_WRAP_SEQ=1 forces 3,
// SAVGOL_B200_WRAP_FUSED=0 skips 1.
}  // extern "C"
namespace {
struct Comp { int dx, dy; float* out; };
std::mutex g_side_mu;
cudaStream_t g_side[sge::kMaxDevices][2] = {};

int run_components(int hx, int hy, int order, const float* in, int rows, int cols, int stride, float delta_x, float delta_y,
                   Savgol2DBoundary boundary, const Comp* comps, int n)
{
    static const bool seq = [] { const char* e = getenv("SAVGOL_B200_WRAP_SEQ"); return e && e[0] == '1'; }();
    bool concurrent = n >= 2 && !seq && in && sge::classify(in) == MemKind::Device;
    const size_t span = rows > 0 && cols > 0 ? static_cast<size_t>(rows - 1) * stride + cols : 0;
    for (int i = 0; i < n && concurrent; ++i) {
        concurrent = comps[i].out && sge::classify(comps[i].out) == MemKind::Device && !ranges_overlap2d(in, span, comps[i].out, span);
        for (int j = 0; j < i && concurrent; ++j) concurrent = !ranges_overlap2d(comps[j].out, span, comps[i].out, span);
    }
    // HOST images (what a caller of the reference passes): upload the image ONCE, run the device path below -- one
    // multi-output launch where it exists -- and fetch every component, instead of one upload per component.
    // Buffers that overlap keep the reference's component-by-component order.
    if (n >= 2 && !seq && in && rows > 0 && cols > 0 && stride >= cols && sge::classify(in) != MemKind::Device) {
        bool plain = true;
        for (int i = 0; i < n && plain; ++i) {
            plain = comps[i].out && sge::classify(comps[i].out) != MemKind::Device && !ranges_overlap2d(in, span, comps[i].out, span);
            for (int j = 0; j < i && plain; ++j) plain = !ranges_overlap2d(comps[j].out, span, comps[i].out, span);
        }
        if (plain && sge::device_ready(true)) {
            cudaStream_t st = sge::current_stream();
            const size_t img = static_cast<size_t>(rows) * cols;
            // the copies travel in row bands through a leased staging pipeline: pinned memory by DMA, pageable memory
            // through the pinned bounce buffers and the host copy pool (host_stage.cu)
            sge::PipeLease lease;
            if (!lease.ok()) return -1;
            sge::Pipeline& P = *lease;
            const float* some_out = comps[0].out;
            for (int i = 1; i < n; ++i)
                if (sge::classify(comps[i].out) == MemKind::Pageable) some_out = comps[i].out;
            P.begin(in, some_out);
            const int rows_b = static_cast<int>(std::max<size_t>(1, std::min<size_t>(rows, (size_t(8) << 18) / cols)));   // ~8 MiB
            if (!P.ensure(static_cast<size_t>(rows_b) * cols, static_cast<size_t>(rows_b) * cols)) return -1;
            float* d = nullptr;
            if (!cuda_ok(cudaMallocAsync(&d, (n + 1) * img * sizeof(float), st), "cudaMallocAsync(wrapper images)")) return -1;
            cudaEvent_t ev = nullptr;
            int rc = cuda_ok(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "event") ? 0 : -1;
            if (rc == 0) {
                cudaEventRecord(ev, st);                       // the allocation (stream ordered) precedes the first copy
                cudaStreamWaitEvent(P.s_in, ev, 0);
            }
            int u = 0;
            for (int r0 = 0; r0 < rows && rc == 0; r0 += rows_b, ++u) {
                const int s = u % sge::Pipeline::kSlots, nr = std::min(rows_b, rows - r0);
                if (!P.h2d(s, d + static_cast<size_t>(r0) * cols, cols, in + static_cast<size_t>(r0) * stride, stride, cols, nr)) rc = -1;
                cudaEventRecord(P.e_in[s], P.s_in);
            }
            Comp dc[3];
            for (int i = 0; i < n; ++i) dc[i] = Comp{comps[i].dx, comps[i].dy, d + (i + 1) * img};
            if (rc == 0) {
                cudaEventRecord(ev, P.s_in);
                cudaStreamWaitEvent(st, ev, 0);
                rc = run_components(hx, hy, order, d, rows, cols, cols, delta_x, delta_y, boundary, dc, n);
                cudaEventRecord(ev, st);
                cudaStreamWaitEvent(P.s_out, ev, 0);
            }
            // VALID defines the interior only, and the reference leaves the border of the output untouched
            const int oy = boundary == SAVGOL2D_BOUNDARY_VALID ? hy : 0, ox = boundary == SAVGOL2D_BOUNDARY_VALID ? hx : 0;
            const int orows = rows - 2 * oy, ocols = cols - 2 * ox;
            for (int i = 0; i < n && rc == 0; ++i)
                for (int r0 = 0; r0 < orows && rc == 0; r0 += rows_b, ++u) {
                    const int s = u % sge::Pipeline::kSlots, nr = std::min(rows_b, orows - r0);
                    if (!P.reuse(s) ||
                        !P.d2h(s, comps[i].out + static_cast<size_t>(oy + r0) * stride + ox, stride,
                               dc[i].out + static_cast<size_t>(oy + r0) * cols + ox, cols, ocols, nr)) rc = -1;
                    cudaEventRecord(P.e_out[s], P.s_out);
                }
            if (!P.finish()) rc = -1;
            if (!cuda_ok(cudaStreamSynchronize(st), "sync")) rc = -1;
            if (ev) cudaEventDestroy(ev);
            cudaFreeAsync(d, st);
            return rc;
        }
    }
    if (!concurrent) {
        for (int i = 0; i < n; ++i) {
            const int rc = component(hx, hy, order, comps[i].dx, comps[i].dy, in, rows, cols, stride, comps[i].out, delta_x, delta_y, boundary);
            if (rc != 0) return rc;
        }
        return 0;
    }
    sge::DeviceGuard guard(in);
    static const bool fused = [] { const char* e = getenv("SAVGOL_B200_WRAP_FUSED"); return !(e && e[0] == '0'); }();
    if (fused && !sge::exact_mode() && boundary != SAVGOL2D_BOUNDARY_VALID && rows > 0 && cols > 0 && sge::device_ready(true)) {
        const Filter2DImpl* fi[3] = {nullptr, nullptr, nullptr};
        const sg2d::SepPlan* plans[3] = {nullptr, nullptr, nullptr};
        float scales[3] = {0.f, 0.f, 0.f};
        bool have = true;
        for (int i = 0; i < n && have; ++i) {
            Savgol2DFilter* f = component_filter(hx, hy, order, comps[i].dx, comps[i].dy, delta_x, delta_y);
            if (!f) return -1;
            fi[i] = live2d(f);
            have = fi[i] != nullptr;
            if (have) { plans[i] = &fi[i]->plan; scales[i] = f->scale; }
        }
        if (have) {
            sg2d::Args2D a{};
            a.in = in;
            a.out = comps[0].out; a.out1 = comps[1].out; a.out2 = n > 2 ? comps[2].out : nullptr;
            a.rows = rows; a.cols = cols;
            a.nx = hx; a.ny = hy;
            a.in_stride = a.out_stride = stride;
            a.n_images = 1;
            a.boundary = boundary == SAVGOL2D_BOUNDARY_REFLECT ? sg2d::B_REFLECT : sg2d::B_CONSTANT;
            a.scale = 1.0f;   // per component, folded into the column factors
            a.out_rows = rows; a.out_cols = cols;
            if (sg2d::multi_supported(a, plans, n))
                return cuda_ok(sg2d::launch_multi(a, plans, scales, n, sge::current_stream()), "sg2d multi-output launch") ? 0 : -1;
        }
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaStream_t user = sge::current_stream();
    cudaStream_t side[2] = {nullptr, nullptr};
    {
        std::lock_guard<std::mutex> lk(g_side_mu);
        for (int k = 0; k < n - 1; ++k) {
            if (!g_side[dev][k] && !cuda_ok(cudaStreamCreateWithFlags(&g_side[dev][k], cudaStreamNonBlocking), "stream")) return -1;
            side[k] = g_side[dev][k];
        }
    }
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
    if (!cuda_ok(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming), "event")) return -1;
    cudaEventRecord(fork, user);
    int rc = 0;
    for (int i = 0; i < n && rc == 0; ++i) {
        cudaStream_t st = i == 0 ? user : side[i - 1];
        if (i > 0) cudaStreamWaitEvent(st, fork, 0);
        savgol_b200_set_stream(st);
        rc = component(hx, hy, order, comps[i].dx, comps[i].dy, in, rows, cols, stride, comps[i].out, delta_x, delta_y, boundary);
        if (i > 0 && cudaEventCreateWithFlags(&join[i - 1], cudaEventDisableTiming) == cudaSuccess) {
            cudaEventRecord(join[i - 1], st);
            cudaStreamWaitEvent(user, join[i - 1], 0);
        } else if (i > 0) {
            cudaStreamSynchronize(st);
        }
    }
    savgol_b200_set_stream(user);
    cudaEventDestroy(fork);
    for (cudaEvent_t e : join)
        if (e) cudaEventDestroy(e);
    return rc;
}
}  // namespace
extern "C" {

// Would savgol2d_gradient (hessian = 0) / savgol2d_hessian (1) run all components of this configuration in ONE multi-output
// launch -- for device images with non-overlapping, equally aligned outputs, a full-size boundary and the default
// arithmetic?  Pure host logic (CPU test-suite).  1 yes, 0 per-component launches, -1 invalid configuration.
int savgol2d_b200_wrapper_plan(int hx, int hy, int order, int hessian)
{
    if (hessian && order < 2) return -1;
    const Comp grad[2] = {{1, 0, nullptr}, {0, 1, nullptr}}, hess[3] = {{2, 0, nullptr}, {1, 1, nullptr}, {0, 2, nullptr}};
    const Comp* comps = hessian ? hess : grad;
    const int n = hessian ? 3 : 2;
    const sg2d::SepPlan* plans[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < n; ++i) {
        Savgol2DFilter* f = component_filter(hx, hy, order, comps[i].dx, comps[i].dy, 1.0f, 1.0f);
        if (!f) return -1;
        const Filter2DImpl* fi = live2d(f);
        if (!fi) return -1;
        plans[i] = &fi->plan;
    }
    sg2d::Args2D a{};
    a.rows = 64; a.cols = 64;
    a.out = reinterpret_cast<float*>(static_cast<uintptr_t>(0x100000));
    a.out1 = a.out + 64 * 64; a.out2 = a.out1 + 64 * 64;
    return sg2d::multi_supported(a, plans, n) ? 1 : 0;
}

int savgol2d_gradient(int hx, int hy, int order, const float* input, int rows, int cols, int stride,
                      float* grad_x, float* grad_y, float delta_x, float delta_y, Savgol2DBoundary boundary)
{
    Comp c[2];
    int n = 0;
    if (grad_x) c[n++] = {1, 0, grad_x};
    if (grad_y) c[n++] = {0, 1, grad_y};
    return run_components(hx, hy, order, input, rows, cols, stride, delta_x, delta_y, boundary, c, n);
}

int savgol2d_hessian(int hx, int hy, int order, const float* input, int rows, int cols, int stride,
                     float* hess_xx, float* hess_xy, float* hess_yy, float delta_x, float delta_y, Savgol2DBoundary boundary)
{
    if (order < 2) {
        fprintf(stderr, "savgol2d_hessian: poly_order must be >= 2\n");
        return -1;
    }
    Comp c[3];
    int n = 0;
    if (hess_xx) c[n++] = {2, 0, hess_xx};
    if (hess_xy) c[n++] = {1, 1, hess_xy};
    if (hess_yy) c[n++] = {0, 2, hess_yy};
    return run_components(hx, hy, order, input, rows, cols, stride, delta_x, delta_y, boundary, c, n);
}

__global__ void add_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, int rows, int cols, long long stride)
{
    const long long total = static_cast<long long>(rows) * cols;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / cols, c = i - r * cols;
        dst[r * stride + c] = __fadd_rn(dst[r * stride + c], src[r * stride + c]);
    }
}

int savgol2d_laplacian(int hx, int hy, int order, const float* input, int rows, int cols, int stride, float* output,
                       float delta_x, float delta_y, Savgol2DBoundary boundary)
{
    if (order < 2) {
        fprintf(stderr, "savgol2d_laplacian: poly_order must be >= 2\n");
        return -1;
    }
    if (!input || !output) return -1;
    // Fused path (default arithmetic): d2/dx2 + d2/dy2 of the fitted polynomial is ONE weight table,
    //   W = Wxx / dx^2 + Wyy / dy^2,
    // again a polynomial surface of the same degree, so it goes through the separable kernel in a
    // single pass (one image read, one write) instead of two filters plus an add
    // (ref composition: src/savgol2d.c:560-618; the exact flavour below keeps that composition).
    if (!sge::exact_mode()) {
        Savgol2DFilter* fl = cached_filter(WrapKey{hx, hy, order, -1, -1, delta_x, delta_y}, [&]() -> Savgol2DFilter* {
            Savgol2DConfig cxx, cyy;
            memset(&cxx, 0, sizeof(cxx));
            cxx.half_window_x = static_cast<uint8_t>(hx); cxx.half_window_y = static_cast<uint8_t>(hy);
            cxx.poly_order = static_cast<uint8_t>(order); cxx.deriv_x = 2; cxx.deriv_y = 0;
            cxx.delta_x = delta_x; cxx.delta_y = delta_y;
            cyy = cxx; cyy.deriv_x = 0; cyy.deriv_y = 2;
            if (!savgol2d_config_valid(&cxx) || !savgol2d_config_valid(&cyy)) {
                fprintf(stderr, "savgol2d_create: invalid configuration\n");
                return nullptr;
            }
            const int area = (2 * hx + 1) * (2 * hy + 1);
            std::vector<float> wxx(area), wyy(area), wl(area);
            double kxx[28], kyy[28], kl[28];
            if (!sgc::weights2d(hx, hy, order, 2, 0, wxx.data(), kxx) || !sgc::weights2d(hx, hy, order, 0, 2, wyy.data(), kyy)) return nullptr;
            const double sxx = static_cast<double>(sgc::scale2d(2, 0, delta_x, delta_y));
            const double syy = static_cast<double>(sgc::scale2d(0, 2, delta_x, delta_y));
            for (int k = 0; k < area; ++k) wl[k] = static_cast<float>(wxx[k] * sxx + wyy[k] * syy);
            for (int k = 0; k < 28; ++k) kl[k] = kxx[k] * sxx + kyy[k] * syy;
            Filter2DImpl* fi = static_cast<Filter2DImpl*>(calloc(1, sizeof(Filter2DImpl)));
            float* wt = static_cast<float*>(malloc(static_cast<size_t>(area) * sizeof(float)));
            if (!fi || !wt) { free(fi); free(wt); return nullptr; }
            memcpy(wt, wl.data(), static_cast<size_t>(area) * sizeof(float));
            fi->pub.config = cxx;
            fi->pub.window_width = 2 * hx + 1; fi->pub.window_height = 2 * hy + 1; fi->pub.window_area = area;
            fi->pub.num_terms = savgol2d_num_terms(order);
            fi->pub.scale = 1.0f;   // both scales are folded into the table
            fi->pub.weights = wt;
            sg2d::plan_separable(hx, hy, order, kl, wt, &fi->plan);
            fi->magic = kMagic2D;
            {
                std::lock_guard<std::mutex> lk(g_mu2d);
                g_live2d.insert(fi);
            }
            return &fi->pub;
        });
        if (fl) return savgol2d_apply(fl, input, rows, cols, stride, output, stride, boundary);
    }
    int rc = component(hx, hy, order, 2, 0, input, rows, cols, stride, output, delta_x, delta_y, boundary);
    if (rc != 0) return rc;
    const size_t span = static_cast<size_t>(rows) * stride;
    if (sge::classify(output) == MemKind::Device) {
        cudaStream_t st = sge::current_stream();
        float* tmp = nullptr;
        if (!cuda_ok(cudaMallocAsync(&tmp, span * sizeof(float), st), "cudaMallocAsync(laplacian)")) return -1;
        cudaMemsetAsync(tmp, 0, span * sizeof(float), st);
        rc = component(hx, hy, order, 0, 2, input, rows, cols, stride, tmp, delta_x, delta_y, boundary);
        if (rc == 0) {
            // output += tmp over the rows x cols region only (ref: src/savgol2d.c:607-614)
            const long long total = static_cast<long long>(rows) * cols;
            const unsigned grid = static_cast<unsigned>(std::min<long long>((total + 255) / 256, 148 * 16));
            add_rows_kernel<<<grid, 256, 0, st>>>(output, tmp, rows, cols, stride);
            if (!cuda_ok(cudaGetLastError(), "laplacian add")) rc = -1;
        }
        cudaFreeAsync(tmp, st);
        return rc;
    }
    std::vector<float> tmp(span, 0.0f);
    rc = component(hx, hy, order, 0, 2, input, rows, cols, stride, tmp.data(), delta_x, delta_y, boundary);
    if (rc == 0)
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) output[static_cast<size_t>(y) * stride + x] += tmp[static_cast<size_t>(y) * stride + x];
    return rc;
}

}  // extern "C"
