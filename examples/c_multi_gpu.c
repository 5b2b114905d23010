/* c_multi_gpu.c -- every GPU of the box from one plain-C99 process (no CUDA headers, no Python, no MPI).
 *
 *   (1) savgol_apply_batch_multi : a host batch sharded over the devices (BASELINE config 2, scaled down)
 *   (2) savgol_apply_batch_multi : ONE long host signal partitioned along its length (config 3 shape)
 *   (3) savgol_apply_slices      : one periodic signal resident on the GPUs as consecutive slices; each
 *                                  device reads its 2 x half_window halo samples from its ring neighbours' memory
 *   (4) savgol2d_apply_batch_multi: a host batch of images sharded over the devices (BASELINE config 4, scaled down)
 * Each result is compared bit for bit with the single-GPU call of the same library in its `exact` flavour
 * (which is itself bit-identical to the reference C code) and within tolerance in the default flavour.
 *
 * usage: c_multi_gpu [n_devices]     (default: all visible; with one GPU the device list repeats it, which
 *                                     exercises the same partitioning logic)
 * build: cc -std=c99 -Iinclude examples/c_multi_gpu.c -Lsavitzky-golay-filter_b200 -lsavgol_b200 -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "savgol_b200.h"

static int failures = 0;
static void check(int ok, const char *what)
{
    printf("%s %s\n", ok ? "[PASS]" : "[FAIL]", what);
    if (!ok) ++failures;
}

static float frand(unsigned *s)
{
    *s = *s * 1664525u + 1013904223u;
    return (float)((*s >> 8) & 0xffff) / 32768.0f - 1.0f;
}

static float max_abs_diff(const float *a, const float *b, size_t n)
{
    float m = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float d = fabsf(a[i] - b[i]);
        if (d > m) m = d;
    }
    return m;
}

int main(int argc, char **argv)
{
    if (!savgol_b200_device_ok()) {
        printf("no sm_100 device; nothing to run\n");
        return 0;
    }
    int visible = savgol_b200_device_count();
    int nd = argc > 1 ? atoi(argv[1]) : (visible > 1 ? visible : 3);
    if (nd < 1) nd = 1;
    if (nd > 16) nd = 16;
    int devices[16];
    for (int i = 0; i < nd; ++i) devices[i] = i % visible;
    printf("%d visible device(s), device list of %d\n", visible, nd);
    unsigned seed = 12345u;

    /* (1) batch of independent signals: 513 x 4096, n16 m3 d1 reflect */
    {
        SavgolConfig cfg = {16, 3, 1, 1.0f, SAVGOL_BOUNDARY_REFLECT};
        SavgolFilter *f = savgol_create(&cfg);
        const size_t rows = 513, len = 4096;
        float *x = malloc(rows * len * sizeof(float)), *y = malloc(rows * len * sizeof(float)), *z = malloc(rows * len * sizeof(float));
        for (size_t i = 0; i < rows * len; ++i) x[i] = frand(&seed);
        int ok = 1;
        for (int exact = 1; exact >= 0; --exact) {
            savgol_b200_set_exact(exact);
            ok = ok && savgol_apply_batch(f, x, z, rows, len, len, len) == 0;
            ok = ok && savgol_apply_batch_multi(f, x, y, rows, len, len, len, devices, nd) == 0;
            ok = ok && memcmp(y, z, rows * len * sizeof(float)) == 0;
        }
        check(ok, "batch sharded over the device list == single-device batch (both flavours, bit for bit)");
        free(x); free(y); free(z);
        savgol_destroy(f);
    }

    /* (2) one long host signal, partitioned along its length: 2^22 + 77 samples, n32 m4 d2 periodic and polynomial */
    for (int mode = 0; mode < 2; ++mode) {
        SavgolConfig cfg = {32, 4, 2, 1.0f, mode ? SAVGOL_BOUNDARY_POLYNOMIAL : SAVGOL_BOUNDARY_PERIODIC};
        SavgolFilter *f = savgol_create(&cfg);
        const size_t len = ((size_t)1 << 22) + 77;
        float *x = malloc(len * sizeof(float)), *y = malloc(len * sizeof(float)), *z = malloc(len * sizeof(float));
        for (size_t i = 0; i < len; ++i) x[i] = frand(&seed);
        savgol_b200_set_exact(1);
        int ok = savgol_apply(f, x, z, len) == 0 && savgol_apply_batch_multi(f, x, y, 1, len, len, len, devices, nd) == 0 &&
                 memcmp(y, z, len * sizeof(float)) == 0;
        savgol_b200_set_exact(0);
        ok = ok && savgol_apply_batch_multi(f, x, y, 1, len, len, len, devices, nd) == 0 && max_abs_diff(y, z, len) <= 1e-6f * 8.0f;
        check(ok, mode ? "long polynomial-edge signal partitioned over the device list == savgol_apply"
                       : "long periodic signal partitioned over the device list == savgol_apply");
        free(x); free(y); free(z);
        savgol_destroy(f);
    }

    /* (3) device-resident slices of one periodic signal; halos come from the neighbours' memory */
    {
        SavgolConfig cfg = {32, 4, 2, 1.0f, SAVGOL_BOUNDARY_PERIODIC};
        SavgolFilter *f = savgol_create(&cfg);
        const size_t slice = ((size_t)1 << 20) + 4 * 13, total = slice * (size_t)nd;
        float *x = malloc(total * sizeof(float)), *y = malloc(total * sizeof(float)), *z = malloc(total * sizeof(float));
        for (size_t i = 0; i < total; ++i) x[i] = frand(&seed);
        const float *in[16];
        float *out[16];
        size_t lens[16];
        int ok = 1;
        for (int i = 0; i < nd; ++i) {
            float *di = savgol_b200_alloc(devices[i], slice * sizeof(float));
            out[i] = savgol_b200_alloc(devices[i], slice * sizeof(float));
            ok = ok && di && out[i] && savgol_b200_copy(di, x + (size_t)i * slice, slice * sizeof(float)) == 0;
            in[i] = di;
            lens[i] = slice;
        }
        for (int exact = 1; exact >= 0 && ok; --exact) {
            savgol_b200_set_exact(exact);
            ok = ok && savgol_apply(f, x, z, total) == 0;                       /* the whole signal, one device */
            ok = ok && savgol_apply_slices(f, in, out, lens, devices, nd) == 0;
            for (int i = 0; i < nd; ++i) ok = ok && savgol_b200_copy(y + (size_t)i * slice, out[i], slice * sizeof(float)) == 0;
            ok = ok && memcmp(y, z, total * sizeof(float)) == 0;
        }
        savgol_b200_set_exact(0);
        check(ok, "device-resident slices with peer-memory halos == savgol_apply of the whole signal (bit for bit)");
        for (int i = 0; i < nd; ++i) {
            savgol_b200_free(devices[i], (void *)in[i]);
            savgol_b200_free(devices[i], out[i]);
        }
        free(x); free(y); free(z);
        savgol_destroy(f);
    }
    /* (4) a host batch of images sharded over the devices */
    {
        Savgol2DConfig cfg = {7, 7, 3, 0, 0, 1.0f, 1.0f};
        Savgol2DFilter *f = savgol2d_create(&cfg);
        const int rows = 300, cols = 520;
        const size_t images = 13, px = (size_t)rows * cols;
        float *x = malloc(images * px * sizeof(float)), *y = malloc(images * px * sizeof(float)), *z = malloc(images * px * sizeof(float));
        for (size_t i = 0; i < images * px; ++i) x[i] = frand(&seed);
        int ok = f != NULL;
        ok = ok && savgol2d_apply_batch(f, x, rows, cols, cols, px, z, cols, px, images, SAVGOL2D_BOUNDARY_CONSTANT) == 0;
        ok = ok && savgol2d_apply_batch_multi(f, x, rows, cols, cols, px, y, cols, px, images, SAVGOL2D_BOUNDARY_CONSTANT, devices, nd) == 0;
        ok = ok && memcmp(y, z, images * px * sizeof(float)) == 0;
        check(ok, "image batch sharded over the device list == single-device batch (bit for bit)");
        free(x); free(y); free(z);
        savgol2d_destroy(f);
    }
    printf("%s\n", failures ? "FAILED" : "all multi-GPU checks passed");
    return failures ? 1 : 0;
}
