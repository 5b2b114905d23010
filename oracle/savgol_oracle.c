/*
 * savgol_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * An independent plain-C restatement of the arithmetic of
 * Tugbars/Savitzky-Golay-Filter for the three hot paths (1D batch, 2D, stream)
 * plus the coefficient generation they depend on.  It exists so that the CUDA
 * path can be compared against something that (a) travels to the GPU box and
 * (b) was itself pinned, bit for bit, against the unmodified reference
 * compiled from /root/reference (oracle/_ref, see oracle/Makefile and
 * tests/test_oracle_vs_ref.py) and against the committed golden fixtures in
 * tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product library
 * (libsavgol_b200.so) never links, loads or calls anything in oracle/.
 *
 * Build: gcc -O2 -ffp-contract=off (FMA contraction changes the reference's
 * results), see oracle/Makefile.
 *
 * All "ref:" citations are path:line inside the reference repository.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SGO_MAX_N 32
#define SGO_MAX_WS 65
#define SGO_MAX_M 64 /* the reference only requires m < 2n+1 and GenFact in range */
#define SGO_MAX_D 4
#define SGO_GF 76 /* ref: src/savgolFilter.c:110 (2*32 + 10 + 2) */

/* ------------------------------------------------------------------ */
/* 1. coefficient generation                                           */
/* ------------------------------------------------------------------ */

/* Falling factorial a*(a-1)*...*(a-b+1), double product rounded once to
 * fp32.  ref: src/savgolFilter.c:151-176 */
static float sgo_gf[SGO_GF][SGO_GF];
static int sgo_gf_ready = 0;

static void sgo_gf_build(void)
{
    if (sgo_gf_ready) return;
    for (int a = 0; a < SGO_GF; ++a)
        for (int b = 0; b < SGO_GF; ++b) {
            if (b == 0) { sgo_gf[a][b] = 1.0f; continue; }
            if (b > a)  { sgo_gf[a][b] = 0.0f; continue; }
            double p = 1.0;
            for (int j = a - b + 1; j <= a; ++j) p *= (double)j;
            sgo_gf[a][b] = (float)p;
        }
    sgo_gf_ready = 1;
}

/* Gram polynomial values F_k^{(d)}(x) for every k = 0..m at once.
 * The reference re-runs the recurrence from k = 0 for every k it needs
 * (ref: src/savgolFilter.c:236-303); the recurrence is deterministic, so one
 * sweep that records every order yields the identical fp32 numbers.
 * out[k] receives F_k^{(d)}(x). */
static void sgo_gram_all(int n, int d, int m, int x, float *out)
{
    float a[SGO_MAX_D + 1] = {0}, b[SGO_MAX_D + 1] = {0}, c[SGO_MAX_D + 1] = {0};
    float *pp = a, *p = b, *cur = c;
    const float nf = (float)n, xf = (float)x;

    for (int q = 0; q <= d; ++q) pp[q] = (q == 0) ? 1.0f : 0.0f;
    out[0] = pp[d];
    if (m == 0) return;

    const float rn = 1.0f / nf;
    p[0] = rn * (xf * pp[0]);
    for (int q = 1; q <= d; ++q) p[q] = rn * (xf * pp[q] + (float)q * pp[q - 1]);
    out[1] = p[d];

    const float n2 = 2.0f * nf;
    for (int k = 2; k <= m; ++k) {
        const float kf = (float)k;
        const float den = kf * (n2 - kf + 1.0f);
        const float al = (4.0f * kf - 2.0f) / den;
        const float ga = ((kf - 1.0f) * (n2 + kf)) / den;
        cur[0] = al * (xf * p[0]) - ga * pp[0];
        for (int q = 1; q <= d; ++q) {
            float t = xf * p[q] + (float)q * p[q - 1];
            cur[q] = al * t - ga * pp[q];
        }
        out[k] = cur[d];
        float *tmp = pp; pp = p; p = cur; cur = tmp;
    }
}

/* w(i,t) = sum_k (2k+1) * GF(2n,k)/GF(2n+k+1,k+1) * F_k(i) * F_k^{(d)}(t)
 * accumulated in fp32, k ascending.  ref: src/savgolFilter.c:336-356 */
static float sgo_weight(int n, int m, int d, int i, int t)
{
    float fi[SGO_MAX_M + 1], ft[SGO_MAX_M + 1];
    sgo_gram_all(n, 0, m, i, fi);
    sgo_gram_all(n, d, m, t, ft);
    float w = 0.0f;
    for (int k = 0; k <= m; ++k) {
        float num = sgo_gf[2 * n][k];
        float den = sgo_gf[2 * n + k + 1][k + 1];
        float fac = (float)(2 * k + 1) * (num / den);
        w += fac * fi[k] * ft[k];
    }
    return w;
}

/* Returns 0 when (n,m,d,dt) is a configuration the reference accepts.
 * ref: src/savgolFilter.c:639-677 */
int sgo_config_ok(int n, int m, int d, float dt)
{
    if (n < 1 || n > SGO_MAX_N) return -1;
    if (m >= 2 * n + 1) return -1;
    if (2 * n + m + 1 >= SGO_GF) return -1; /* reference would index its table out of range */
    if (d > SGO_MAX_D) return -1;
    if (d > m) return -1;
    if (!(dt > 0.0f)) return -1;
    return 0;
}

/* center: 2n+1 floats (target 0); edge: row e (0..n-1) has target n-e and
 * lives at edge[e*65 .. e*65+2n].  Rows/columns beyond stay untouched.
 * ref: src/savgolFilter.c:368-409 */
int sgo_weights_1d(int n, int m, int d, float *center, float *edge)
{
    if (n < 1 || n > SGO_MAX_N || m < 0 || m >= 2 * n + 1 || 2 * n + m + 1 >= SGO_GF || d > SGO_MAX_D || d > m)
        return -1;
    sgo_gf_build();
    const int ws = 2 * n + 1;
    if (center)
        for (int c = 0; c < ws; ++c) center[c] = sgo_weight(n, m, d, c - n, 0);
    if (edge)
        for (int e = 0; e < n; ++e)
            for (int c = 0; c < ws; ++c)
                edge[e * SGO_MAX_WS + c] = sgo_weight(n, m, d, c - n, n - e);
    return 0;
}

/* dt_scale = powf(dt, d); dt_inv = 1/dt_scale unless dt_scale == 0.
 * ref: src/savgolFilter.c:707,759 */
float sgo_dt_inv(float dt, int d)
{
    float s = powf(dt, (float)d);
    return (s != 0.0f) ? (1.0f / s) : 1.0f;
}

/* ------------------------------------------------------------------ */
/* 2. 1D apply                                                          */
/* ------------------------------------------------------------------ */

/* The reference's four-chain dot product: the ws&3 leading taps go to chains
 * 0..rem-1, the rest round-robin over four chains, result (s0+s1)+(s2+s3);
 * every product is rounded before it is added (no FMA).  `step` is +1 for a
 * forward window and -1 for the reversed leading-edge traversal.
 * ref: src/savgolFilter.c:547-580, 593-623 */
static float sgo_dot4(const float *w, const float *x, int ws, ptrdiff_t step)
{
    float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const int rem = ws & 3;
    for (int k = 0; k < rem; ++k) s[k] += w[k] * x[(ptrdiff_t)k * step];
    for (int k = rem; k < ws; ++k) s[(k - rem) & 3] += w[k] * x[(ptrdiff_t)k * step];
    return (s[0] + s[1]) + (s[2] + s[3]);
}

/* Virtual sample for the padded modes.  mode: 1 reflect (half-sample
 * symmetric), 2 periodic, 3 constant.  64-bit indices: for length < 2^31 this
 * equals the reference's int arithmetic (ref: src/savgolFilter.c:442-482);
 * beyond that the reference overflows and this is the
 * documented extension. */
static float sgo_virtual(const float *x, int64_t L, int64_t i, int mode)
{
    if (i >= 0 && i < L) return x[i];
    switch (mode) {
    case 1:
        if (i < 0) { i = -i - 1; if (i >= L) i = L - 1; }
        else       { i = 2 * L - i - 1; if (i < 0) i = 0; }
        return x[i];
    case 2:
        i = ((i % L) + L) % L;
        return x[i];
    case 3:
        return (i < 0) ? x[0] : x[L - 1];
    default:
        return 0.0f;
    }
}

/* savgol_apply.  mode 0 polynomial, 1 reflect, 2 periodic, 3 constant.
 * ref: src/savgolFilter.c:743-804 */
int sgo_apply(int n, const float *center, const float *edge, float dt_inv, int mode,
              const float *in, float *out, size_t L)
{
    const int ws = 2 * n + 1;
    if (!center || !in || !out) return -1;
    if (L < (size_t)ws) return -1;

    for (size_t j = (size_t)n; j < L - (size_t)n; ++j)
        out[j] = sgo_dot4(center, in + (j - n), ws, 1) * dt_inv;

    if (mode == 0) {
        for (int e = 0; e < n; ++e) {
            const float *row = edge + e * SGO_MAX_WS;
            out[e] = sgo_dot4(row, in + (ws - 1), ws, -1) * dt_inv;
            out[L - 1 - (size_t)e] = sgo_dot4(row, in + (L - ws), ws, 1) * dt_inv;
        }
    } else {
        float win[SGO_MAX_WS];
        for (int side = 0; side < 2; ++side)
            for (int e = 0; e < n; ++e) {
                int64_t c = side ? (int64_t)L - n + e : e;
                for (int k = 0; k < ws; ++k) win[k] = sgo_virtual(in, (int64_t)L, c - n + k, mode);
                out[c] = sgo_dot4(center, win, ws, 1) * dt_inv;
            }
    }
    return 0;
}

/* savgol_apply_valid.  ref: src/savgolFilter.c:821-850 */
size_t sgo_apply_valid(int n, const float *center, float dt_inv,
                       const float *in, size_t L, float *out)
{
    const int ws = 2 * n + 1;
    if (!center || !in || !out || L < (size_t)ws) return 0;
    const size_t outL = L - 2 * (size_t)n;
    for (size_t j = 0; j < outL; ++j) out[j] = sgo_dot4(center, in + j, ws, 1) * dt_inv;
    return outL;
}

/* savgol_apply_strided: byte strides / offsets; polynomial edges whatever the
 * configured boundary.  ref: src/savgolFilter.c:877-934 */
int sgo_apply_strided(int n, const float *center, const float *edge, float dt_inv,
                      const void *in, size_t is, size_t io,
                      void *out, size_t os, size_t oo, size_t count)
{
    const int ws = 2 * n + 1;
    if (!center || !edge || !in || !out || count < (size_t)ws) return -1;
    const char *ib = (const char *)in + io;
    char *ob = (char *)out + oo;
    float win[SGO_MAX_WS];
#define SGO_IN(i) (*(const float *)(ib + (size_t)(i) * is))
#define SGO_OUT(i) (*(float *)(ob + (size_t)(i) * os))
    for (size_t j = (size_t)n; j < count - (size_t)n; ++j) {
        for (int k = 0; k < ws; ++k) win[k] = SGO_IN(j - n + k);
        SGO_OUT(j) = sgo_dot4(center, win, ws, 1) * dt_inv;
    }
    for (int e = 0; e < n; ++e) {
        for (int k = 0; k < ws; ++k) win[k] = SGO_IN(k);
        SGO_OUT(e) = sgo_dot4(edge + e * SGO_MAX_WS, win + (ws - 1), ws, -1) * dt_inv;
    }
    for (int e = 0; e < n; ++e) {
        for (int k = 0; k < ws; ++k) win[k] = SGO_IN(count - ws + k);
        SGO_OUT(count - 1 - e) = sgo_dot4(edge + e * SGO_MAX_WS, win, ws, 1) * dt_inv;
    }
#undef SGO_IN
#undef SGO_OUT
    return 0;
}

/* Row loop = what a caller of the reference does for a batch (SURVEY.md
 * section 1: "a caller's for loop over savgol_apply"). */
int sgo_apply_batch(int n, const float *center, const float *edge, float dt_inv, int mode,
                    const float *in, float *out, size_t rows, size_t L,
                    size_t in_pitch, size_t out_pitch)
{
    for (size_t r = 0; r < rows; ++r)
        if (sgo_apply(n, center, edge, dt_inv, mode, in + r * in_pitch, out + r * out_pitch, L))
            return -1;
    return 0;
}

/* ------------------------------------------------------------------ */
/* 3. stream (single channel; the multi-channel chunked product API is  */
/*    defined as "this, per channel")                                   */
/* ------------------------------------------------------------------ */

/* Sequential single-accumulator sums over the last ws samples, oldest
 * first for centre / trailing, newest first for leading.
 * ref: src/savgol_stream.c:25-74.  `hist` points at the oldest of the ws
 * most recent samples, stored linearly (the ring buffer of the reference is
 * only a storage detail: (start+i)%ws enumerates oldest..newest). */
static float sgo_seq_fwd(const float *w, const float *hist, int ws)
{
    float s = 0.0f;
    for (int i = 0; i < ws; ++i) s += w[i] * hist[i];
    return s;
}
static float sgo_seq_rev(const float *w, const float *hist, int ws)
{
    float s = 0.0f;
    for (int i = 0; i < ws; ++i) s += w[i] * hist[ws - 1 - i];
    return s;
}

/* Whole-signal stream run: push_full for every sample, then flush.
 * Produces exactly len outputs when len >= ws (else 0) in chronological
 * order.  ref: src/savgol_stream.c:180-252.  with_leading = 0 reproduces the
 * plain savgol_stream_push sequence (no leading edge, ref :152-178) followed
 * by flush. */
size_t sgo_stream_run(int n, const float *center, const float *edge, float dt_inv,
                      const float *in, size_t len, float *out, int with_leading)
{
    const int ws = 2 * n + 1;
    size_t o = 0;
    if (len < (size_t)ws) return 0;
    for (size_t t = (size_t)ws - 1; t < len; ++t) {
        const float *hist = in + (t + 1 - ws);
        if (t == (size_t)ws - 1 && with_leading)
            for (int e = 0; e < n; ++e)
                out[o++] = sgo_seq_rev(edge + e * SGO_MAX_WS, hist, ws) * dt_inv;
        out[o++] = sgo_seq_fwd(center, hist, ws) * dt_inv;
    }
    const float *hist = in + (len - ws);
    for (int i = 0; i < n; ++i)
        out[o++] = sgo_seq_fwd(edge + (n - 1 - i) * SGO_MAX_WS, hist, ws) * dt_inv;
    return o;
}

/* ------------------------------------------------------------------ */
/* 4. 2D                                                                */
/* ------------------------------------------------------------------ */

static int sgo_mono(int i, int j) { int t = i + j; return t * (t + 1) / 2 + j; }

/* ref: src/savgol2d.c:271-302 */
int sgo2d_config_ok(int nx, int ny, int order, int dx, int dy, float hx, float hy)
{
    if (nx < 1 || nx > 16 || ny < 1 || ny > 16) return -1;
    if (order < 0 || order > 6) return -1;
    if (dx + dy > order) return -1;
    if (!(hx > 0.0f) || !(hy > 0.0f)) return -1;
    if ((2 * nx + 1) * (2 * ny + 1) < (order + 1) * (order + 2) / 2) return -1;
    return 0;
}

/* Least-squares weights: design matrix of monomials x^i y^j (i+j<=order,
 * ordered by total degree then rising j), normal equations, Cholesky, then
 * weights = (float)((A c) * dx! dy!), all in double.
 * ref: src/savgol2d.c:77-265 */
int sgo2d_weights(int nx, int ny, int order, int dx, int dy, float *W)
{
    const int ww = 2 * nx + 1, wh = 2 * ny + 1, area = ww * wh;
    const int nt = (order + 1) * (order + 2) / 2;
    double *A = (double *)malloc(sizeof(double) * (size_t)area * nt);
    double G[28 * 28], y[28], c[28], rhs[28];
    if (!A) return -1;

    int r = 0;
    for (int yi = -ny; yi <= ny; ++yi)
        for (int xi = -nx; xi <= nx; ++xi, ++r)
            for (int tot = 0; tot <= order; ++tot)
                for (int j = 0; j <= tot; ++j) {
                    int i = tot - j;
                    A[r * nt + sgo_mono(i, j)] = pow((double)xi, i) * pow((double)yi, j);
                }

    for (int p = 0; p < nt; ++p)
        for (int q = 0; q < nt; ++q) {
            double s = 0.0;
            for (int k = 0; k < area; ++k) s += A[k * nt + p] * A[k * nt + q];
            G[p * nt + q] = s;
        }

    for (int p = 0; p < nt; ++p) rhs[p] = 0.0;
    rhs[sgo_mono(dx, dy)] = 1.0;

    /* in-place lower Cholesky, row by row */
    for (int i = 0; i < nt; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = G[i * nt + j];
            for (int k = 0; k < j; ++k) s -= G[i * nt + k] * G[j * nt + k];
            if (i == j) {
                if (s <= 0.0) { free(A); return -1; }
                G[i * nt + i] = sqrt(s);
            } else {
                G[i * nt + j] = s / G[j * nt + j];
            }
        }
    for (int i = 0; i < nt; ++i) {
        double s = rhs[i];
        for (int j = 0; j < i; ++j) s -= G[i * nt + j] * y[j];
        y[i] = s / G[i * nt + i];
    }
    for (int i = nt - 1; i >= 0; --i) {
        double s = y[i];
        for (int j = i + 1; j < nt; ++j) s -= G[j * nt + i] * c[j];
        c[i] = s / G[i * nt + i];
    }

    double fx = 1.0, fy = 1.0;
    for (int i = 2; i <= dx; ++i) fx *= i;
    for (int i = 2; i <= dy; ++i) fy *= i;
    const double ds = fx * fy;
    for (int k = 0; k < area; ++k) {
        double s = 0.0;
        for (int p = 0; p < nt; ++p) s += A[k * nt + p] * c[p];
        W[k] = (float)(s * ds);
    }
    free(A);
    return 0;
}

/* ref: src/savgol2d.c:320-322 */
float sgo2d_scale(int dx, int dy, float hx, float hy)
{
    return 1.0f / (powf(hx, (float)dx) * powf(hy, (float)dy));
}

/* Interior only; out[oy*os+ox] is the result centred on (oy+ny, ox+nx).
 * ref: src/savgol2d.c:356-396 */
int sgo2d_apply_valid(int nx, int ny, const float *W, float scale,
                      const float *in, int rows, int cols, int is, float *out, int os)
{
    if (!W || !in || !out) return -1;
    const int ww = 2 * nx + 1, wh = 2 * ny + 1;
    const int orows = rows - 2 * ny, ocols = cols - 2 * nx;
    if (orows <= 0 || ocols <= 0) return -1;
    for (int oy = 0; oy < orows; ++oy)
        for (int ox = 0; ox < ocols; ++ox) {
            float s = 0.0f;
            const float *w = W;
            for (int wy = 0; wy < wh; ++wy) {
                const float *p = in + (ptrdiff_t)(oy + wy) * is + ox;
                for (int wx = 0; wx < ww; ++wx) s += *w++ * p[wx];
            }
            out[(ptrdiff_t)oy * os + ox] = s * scale;
        }
    return 0;
}

/* boundary: 0 valid (written at offset (ny,nx), border untouched),
 * 1 constant (clamp), 2 reflect (half-sample symmetric then clamp).
 * ref: src/savgol2d.c:398-456 */
int sgo2d_apply(int nx, int ny, const float *W, float scale,
                const float *in, int rows, int cols, int is, float *out, int os, int boundary)
{
    if (!W || !in || !out) return -1;
    if (boundary == 0)
        return sgo2d_apply_valid(nx, ny, W, scale, in, rows, cols, is,
                                 out + (ptrdiff_t)ny * os + nx, os);
    for (int oy = 0; oy < rows; ++oy)
        for (int ox = 0; ox < cols; ++ox) {
            float s = 0.0f;
            const float *w = W;
            for (int wy = -ny; wy <= ny; ++wy)
                for (int wx = -nx; wx <= nx; ++wx) {
                    int iy = oy + wy, ix = ox + wx;
                    if (boundary == 2) {
                        if (iy < 0) iy = -iy - 1; else if (iy >= rows) iy = 2 * rows - iy - 1;
                        if (ix < 0) ix = -ix - 1; else if (ix >= cols) ix = 2 * cols - ix - 1;
                    }
                    if (iy < 0) iy = 0; else if (iy >= rows) iy = rows - 1;
                    if (ix < 0) ix = 0; else if (ix >= cols) ix = cols - 1;
                    s += *w++ * in[(ptrdiff_t)iy * is + ix];
                }
            out[(ptrdiff_t)oy * os + ox] = s * scale;
        }
    return 0;
}
