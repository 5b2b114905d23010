/* c_extensions.c -- the extension entry points of include/savgol_b200.h used from plain C with host
 * buffers (the library stages them; no CUDA header needed by the caller).
 *
 *   cc -std=c99 -Iinclude examples/c_extensions.c -Lsavitzky-golay-filter_b200 -lsavgol_b200 -lm
 *
 * Checks, each against the reference-shaped single-call API of the same library:
 *   1. savgol_apply_batch == a loop of savgol_apply          (BASELINE config 2 shape, reduced)
 *   2. savgol_mcstream_push/flush == the scalar savgol_stream_push_full/flush of every channel
 *   3. savgol2d_apply_batch == a loop of savgol2d_apply
 *   4. checkpoint save -> restore continues the stream bit-identically
 * Prints one PASS/FAIL line per check, exit code = number of failures. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "savgol_b200.h"

static float frand(unsigned *s)
{
    *s = *s * 1664525u + 1013904223u;
    return (float)((*s >> 8) & 0xffff) / 32768.0f - 1.0f;
}

static int report(const char *what, int ok)
{
    printf("[%s] %s\n", ok ? "PASS" : "FAIL", what);
    return ok ? 0 : 1;
}

int main(void)
{
    int failures = 0;
    unsigned seed = 12345u;
    if (!savgol_b200_device_ok()) {
        printf("no compute-capability 10.x device: nothing to run\n");
        return 77;
    }

    /* 1. batch of signals */
    {
        const size_t rows = 37, len = 4096, pitch = 4100;
        SavgolConfig cfg = {16, 3, 1, 1.0f, SAVGOL_BOUNDARY_REFLECT};
        SavgolFilter *f = savgol_create(&cfg);
        float *x = malloc(rows * pitch * sizeof(float)), *y = malloc(rows * pitch * sizeof(float)), *z = malloc(len * sizeof(float));
        int ok = f != NULL;
        for (size_t i = 0; i < rows * pitch; ++i) x[i] = frand(&seed);
        ok = ok && savgol_apply_batch(f, x, y, rows, len, pitch, pitch) == 0;
        for (size_t r = 0; ok && r < rows; ++r) {
            ok = savgol_apply(f, x + r * pitch, z, len) == 0 && memcmp(z, y + r * pitch, len * sizeof(float)) == 0;
        }
        failures += report("savgol_apply_batch equals a loop of savgol_apply", ok);
        free(x); free(y); free(z);
        savgol_destroy(f);
    }

    /* 2. + 4. multichannel stream vs the scalar stream, with a checkpoint in the middle */
    {
        const size_t C = 19, n = 10, chunks[4] = {7, 300, 1024, 333};
        size_t total = 0, pos = 0, got = 0;
        SavgolConfig cfg = {10, 2, 1, 0.5f, SAVGOL_BOUNDARY_POLYNOMIAL};
        for (int i = 0; i < 4; ++i) total += chunks[i];
        float *sig = malloc(C * total * sizeof(float)), *out = malloc(C * total * sizeof(float));
        float *tmp = malloc(C * (1024 + n) * sizeof(float)), *ref = malloc(total * sizeof(float));
        for (size_t i = 0; i < C * total; ++i) sig[i] = frand(&seed);
        SavgolMCStream *s = savgol_mcstream_create(&cfg, C), *s2 = NULL;
        int ok = s != NULL, ok_ckpt = 1;
        void *blob = NULL;
        for (int i = 0; ok && i < 4; ++i) {
            const size_t K = chunks[i];
            if (i == 2) { /* checkpoint before the third chunk, resume in a second stream */
                const size_t nb = savgol_mcstream_checkpoint_size(s);
                blob = malloc(nb);
                s2 = savgol_mcstream_create(&cfg, C);
                ok_ckpt = savgol_mcstream_save(s, blob, nb) == (long long)nb && s2 && savgol_mcstream_restore(s2, blob, nb) == 0;
            }
            const long long k = savgol_mcstream_push(s, sig + pos, total, K, tmp, 1024 + n);
            ok = k >= 0;
            for (size_t c = 0; ok && c < C; ++c) memcpy(out + c * total + got, tmp + c * (1024 + n), (size_t)k * sizeof(float));
            if (s2 && ok_ckpt) { /* the resumed stream must produce the same bits */
                float *tmp2 = malloc(C * (1024 + n) * sizeof(float));
                const long long k2 = savgol_mcstream_push(s2, sig + pos, total, K, tmp2, 1024 + n);
                ok_ckpt = k2 == k;
                for (size_t c = 0; ok_ckpt && c < C; ++c)
                    ok_ckpt = memcmp(tmp2 + c * (1024 + n), tmp + c * (1024 + n), (size_t)k * sizeof(float)) == 0;
                free(tmp2);
            }
            pos += K;
            got += (size_t)k;
        }
        if (ok) {
            const long long k = savgol_mcstream_flush(s, tmp, 1024 + n);
            ok = k == (long long)n;
            for (size_t c = 0; ok && c < C; ++c) memcpy(out + c * total + got, tmp + c * (1024 + n), n * sizeof(float));
            got += n;
        }
        ok = ok && got == total;
        double worst = 0.0, amp = 0.0;
        for (size_t c = 0; ok && c < C; ++c) { /* scalar stream of the same library == reference arithmetic */
            SavgolStream *ss = savgol_stream_create(&cfg);
            size_t w = 0;
            float buf[SAVGOL_MAX_HALF_WINDOW + 1];
            for (size_t t = 0; t < total; ++t) {
                const int k = savgol_stream_push_full(ss, sig[c * total + t], buf, SAVGOL_MAX_HALF_WINDOW + 1);
                for (int j = 0; j < k; ++j) ref[w++] = buf[j];
                if (fabs(sig[c * total + t]) > amp) amp = fabs(sig[c * total + t]);
            }
            const int kf = savgol_stream_flush(ss, buf, SAVGOL_MAX_HALF_WINDOW + 1);
            for (int j = 0; j < kf; ++j) ref[w++] = buf[j];
            ok = w == total;
            for (size_t t = 0; ok && t < total; ++t) {
                const double e = fabs((double)ref[t] - (double)out[c * total + t]);
                if (e > worst) worst = e;
            }
            savgol_stream_destroy(ss);
        }
        ok = ok && worst <= 1e-6 * amp / 0.5; /* north_star tolerance, d = 1, dt = 0.5 */
        failures += report("savgol_mcstream_push/flush equals the per-channel scalar stream", ok);
        failures += report("savgol_mcstream_save/restore resumes bit-identically", ok_ckpt && s2 != NULL);
        free(sig); free(out); free(tmp); free(ref); free(blob);
        savgol_mcstream_destroy(s);
        savgol_mcstream_destroy(s2);
    }

    /* 3. batch of images */
    {
        const int rows = 150, cols = 260, images = 3;
        const size_t px = (size_t)rows * cols;
        Savgol2DConfig cfg = {7, 7, 3, 0, 0, 1.0f, 1.0f};
        Savgol2DFilter *f = savgol2d_create(&cfg);
        float *x = malloc(images * px * sizeof(float)), *y = malloc(images * px * sizeof(float)), *z = malloc(px * sizeof(float));
        int ok = f != NULL;
        for (size_t i = 0; i < images * px; ++i) x[i] = frand(&seed);
        ok = ok && savgol2d_apply_batch(f, x, rows, cols, cols, px, y, cols, px, images, SAVGOL2D_BOUNDARY_REFLECT) == 0;
        for (int i = 0; ok && i < images; ++i)
            ok = savgol2d_apply(f, x + i * px, rows, cols, cols, z, cols, SAVGOL2D_BOUNDARY_REFLECT) == 0 &&
                 memcmp(z, y + i * px, px * sizeof(float)) == 0;
        failures += report("savgol2d_apply_batch equals a loop of savgol2d_apply", ok);
        free(x); free(y); free(z);
        savgol2d_destroy(f);
    }
    return failures;
}
