// coeffs.cpp -- host-side coefficient generation (filter creation time).
//
// Produces, in fp32 and with the reference's exact operation order, the centre weights and
// the polynomial edge-weight table of the 1D filter (GenFact falling-factorial table + Gram
// polynomial three-term recurrence; ref: src/savgolFilter.c:151-176, 236-303, 336-409) and,
// in double, the least-squares weights of the 2D filter (ref: src/savgol2d.c:77-265).
// The results are bit-identical to the reference's tables (tests/test_abi_cpu.py checks every
// valid (n,m,d)); they are uploaded once per filter and never recomputed on the device.
//
// MUST be compiled with -ffp-contract=off and without -ffast-math: an FMA contraction changes
// the low bits of the weights.
#include "coeffs.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace sgc {

namespace {

constexpr int kGF = 2 * 32 + 10 + 2;  // ref: src/savgolFilter.c:110
constexpr int kMaxOrders = 65;        // poly_order < 2n+1 <= 65

// Falling factorials a!/(a-b)! as the reference tabulates them: double product, one rounding.
struct FallingFactorials {
    float v[kGF][kGF];
    FallingFactorials()
    {
        for (int a = 0; a < kGF; ++a) {
            for (int b = 0; b < kGF; ++b) {
                if (b == 0) v[a][b] = 1.0f;
                else if (b > a) v[a][b] = 0.0f;
                else {
                    double p = 1.0;
                    for (int j = a - b + 1; j <= a; ++j) p *= static_cast<double>(j);
                    v[a][b] = static_cast<float>(p);
                }
            }
        }
    }
};

const FallingFactorials& gf()
{
    static const FallingFactorials table;  // thread-safe init (the reference's is racy-benign)
    return table;
}

// Gram polynomial values of every order 0..m at abscissa x, derivative order d.
// Same recurrence and rounding sequence as the reference's per-order evaluation.
void gram_orders(int n, int d, int m, int x, float* f_of_k)
{
    float r0[5] = {0, 0, 0, 0, 0}, r1[5] = {0, 0, 0, 0, 0}, r2[5] = {0, 0, 0, 0, 0};
    float *older = r0, *old = r1, *now = r2;
    const float nf = static_cast<float>(n);
    const float xf = static_cast<float>(x);

    for (int q = 0; q <= d; ++q) older[q] = q == 0 ? 1.0f : 0.0f;
    f_of_k[0] = older[d];
    if (m < 1) return;

    const float inv_n = 1.0f / nf;
    old[0] = inv_n * (xf * older[0]);
    for (int q = 1; q <= d; ++q) old[q] = inv_n * (xf * older[q] + static_cast<float>(q) * older[q - 1]);
    f_of_k[1] = old[d];

    const float two_n = 2.0f * nf;
    for (int k = 2; k <= m; ++k) {
        const float kf = static_cast<float>(k);
        const float denom = kf * (two_n - kf + 1.0f);
        const float alpha = (4.0f * kf - 2.0f) / denom;
        const float gamma = ((kf - 1.0f) * (two_n + kf)) / denom;
        now[0] = alpha * (xf * old[0]) - gamma * older[0];
        for (int q = 1; q <= d; ++q) {
            const float term = xf * old[q] + static_cast<float>(q) * old[q - 1];
            now[q] = alpha * term - gamma * older[q];
        }
        f_of_k[k] = now[d];
        float* t = older; older = old; old = now; now = t;
    }
}

}  // namespace

bool config1d_valid(int n, int m, int d, float dt, const char** why)
{
    const char* msg = nullptr;
    if (n < 1 || n > 32) msg = "half_window must be in [1, 32]";
    else if (m >= 2 * n + 1) msg = "poly_order must be < window_size";
    // The reference does not enforce SAVGOL_MAX_POLY_ORDER; orders above 10 work there as long
    // as its 76x76 GenFact table covers GenFact(2n+m+1, m+1) (src/savgolFilter.c:110,185-194).
    // Beyond that it reads out of range and produces NaN weights; this library refuses instead.
    else if (2 * n + m + 1 >= kGF) msg = "poly_order exceeds the GenFact table (2n+m+1 must be < 76)";
    else if (d > 4) msg = "derivative must be <= 4";
    else if (d > m) msg = "derivative cannot exceed poly_order";
    else if (!(dt > 0.0f)) msg = "time_step must be > 0";
    if (why) *why = msg;
    return msg == nullptr;
}

void weights1d(int n, int m, int d, float* center /*[65]*/, float* edge /*[32][65]*/)
{
    const FallingFactorials& G = gf();
    const int ws = 2 * n + 1;

    // per-order normalisation (2k+1) * GF(2n,k) / GF(2n+k+1,k+1)
    float fac[kMaxOrders];
    for (int k = 0; k <= m; ++k) {
        const float num = G.v[2 * n][k];
        const float den = G.v[2 * n + k + 1][k + 1];
        fac[k] = static_cast<float>(2 * k + 1) * (num / den);
    }
    // F_k(i) for every data position, F_k^{(d)}(t) for every target 0..n
    std::vector<float> at_data(static_cast<size_t>(ws) * kMaxOrders), at_target(static_cast<size_t>(n + 1) * kMaxOrders);
    for (int c = 0; c < ws; ++c) gram_orders(n, 0, m, c - n, &at_data[static_cast<size_t>(c) * kMaxOrders]);
    for (int t = 0; t <= n; ++t) gram_orders(n, d, m, t, &at_target[static_cast<size_t>(t) * kMaxOrders]);

    auto weight = [&](int c, int t) {
        const float* fi = &at_data[static_cast<size_t>(c) * kMaxOrders];
        const float* ft = &at_target[static_cast<size_t>(t) * kMaxOrders];
        float w = 0.0f;
        for (int k = 0; k <= m; ++k) w += fac[k] * fi[k] * ft[k];
        return w;
    };
    for (int c = 0; c < ws; ++c) center[c] = weight(c, 0);
    for (int e = 0; e < n; ++e)
        for (int c = 0; c < ws; ++c) edge[e * 65 + c] = weight(c, n - e);
}

float dt_scale(float dt, int d) { return powf(dt, static_cast<float>(d)); }

// ------------------------------------------------------------------------------------------
// 2D

static inline int mono(int i, int j) { const int t = i + j; return t * (t + 1) / 2 + j; }

bool config2d_valid(int nx, int ny, int order, int dx, int dy, float hx, float hy)
{
    if (nx < 1 || nx > 16 || ny < 1 || ny > 16) return false;
    if (order > 6) return false;
    if (dx + dy > order) return false;
    if (!(hx > 0.0f) || !(hy > 0.0f)) return false;
    return (2 * nx + 1) * (2 * ny + 1) >= (order + 1) * (order + 2) / 2;
}

float scale2d(int dx, int dy, float hx, float hy)
{
    return 1.0f / (powf(hx, static_cast<float>(dx)) * powf(hy, static_cast<float>(dy)));
}

// Solves the normal equations for the row of the pseudo-inverse that belongs to x^dx y^dy.
// coef[] (nterms doubles, monomial order of the reference) describes the weight surface
// W(x,y) = dx! dy! * sum coef[mono(i,j)] x^i y^j; weights[] is that surface sampled on the window,
// rounded to fp32 exactly like the reference.
bool weights2d(int nx, int ny, int order, int dx, int dy, float* weights, double* coef_out)
{
    const int ww = 2 * nx + 1, wh = 2 * ny + 1, area = ww * wh;
    const int nt = (order + 1) * (order + 2) / 2;
    std::vector<double> A(static_cast<size_t>(area) * nt);
    int r = 0;
    for (int y = -ny; y <= ny; ++y)
        for (int x = -nx; x <= nx; ++x, ++r)
            for (int tot = 0; tot <= order; ++tot)
                for (int j = 0; j <= tot; ++j)
                    A[static_cast<size_t>(r) * nt + mono(tot - j, j)] =
                        std::pow(static_cast<double>(x), tot - j) * std::pow(static_cast<double>(y), j);

    double M[28 * 28];
    for (int p = 0; p < nt; ++p)
        for (int q = 0; q < nt; ++q) {
            double s = 0.0;
            for (int k = 0; k < area; ++k) s += A[static_cast<size_t>(k) * nt + p] * A[static_cast<size_t>(k) * nt + q];
            M[p * nt + q] = s;
        }
    double rhs[28], fwd[28], c[28];
    for (int p = 0; p < nt; ++p) rhs[p] = 0.0;
    rhs[mono(dx, dy)] = 1.0;

    for (int i = 0; i < nt; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = M[i * nt + j];
            for (int k = 0; k < j; ++k) s -= M[i * nt + k] * M[j * nt + k];
            if (i == j) {
                if (s <= 0.0) return false;
                M[i * nt + i] = std::sqrt(s);
            } else {
                M[i * nt + j] = s / M[j * nt + j];
            }
        }
    for (int i = 0; i < nt; ++i) {
        double s = rhs[i];
        for (int j = 0; j < i; ++j) s -= M[i * nt + j] * fwd[j];
        fwd[i] = s / M[i * nt + i];
    }
    for (int i = nt - 1; i >= 0; --i) {
        double s = fwd[i];
        for (int j = i + 1; j < nt; ++j) s -= M[j * nt + i] * c[j];
        c[i] = s / M[i * nt + i];
    }
    double fx = 1.0, fy = 1.0;
    for (int i = 2; i <= dx; ++i) fx *= i;
    for (int i = 2; i <= dy; ++i) fy *= i;
    const double ds = fx * fy;
    for (int k = 0; k < area; ++k) {
        double s = 0.0;
        for (int p = 0; p < nt; ++p) s += A[static_cast<size_t>(k) * nt + p] * c[p];
        weights[k] = static_cast<float>(s * ds);
    }
    if (coef_out)
        for (int p = 0; p < nt; ++p) coef_out[p] = c[p] * ds;
    return true;
}

}  // namespace sgc
