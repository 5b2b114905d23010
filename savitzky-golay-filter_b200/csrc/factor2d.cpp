// factor2d.cpp -- host-side separable factorisation of the 2D Savitzky-Golay weight surface.
//
// The reference stores the (2ny+1) x (2nx+1) weights as a dense table and applies them tap by tap
// (src/savgol2d.c:380-391): 225 MACs per pixel for the 15x15 window of BASELINE config 4, far above
// the ~43 fp32 operations per pixel a B200 can afford at the HBM roofline.  But the table is a
// polynomial surface of total degree <= 6 sampled on the window,
//     W(y,x) = dx! dy! * sum_{i+j<=order} c_ij x^i y^j          (src/savgol2d.c:216-256),
// so as a matrix it has rank <= 4 (rank 2 for the common order-2/3 smoothing filters), and every
// factor is even or odd in its variable.  This file computes   W = sum_r col_r(y) * row_r(x)   by a
// one-sided Jacobi SVD of the table in double precision, cleans the parity of the factors, balances
// them (sqrt(sigma) on each side) and accepts the factorisation only if it reproduces the fp32
// table the reference would use to well below the parity tolerance.  The device kernel
// (sg2d_sep.cu) then needs (2nx+1)+(2ny+1) MACs per pixel and factor instead of their product.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sg2d.h"

namespace sg2d {

namespace {

// One-sided Jacobi: orthogonalises the columns of A (m x n, row-major, m >= 1) in place and
// accumulates V (n x n).  On return column j of A is sigma_j * u_j.
void jacobi_svd(std::vector<double>& A, int m, int n, std::vector<double>& V)
{
    V.assign(static_cast<size_t>(n) * n, 0.0);
    for (int i = 0; i < n; ++i) V[static_cast<size_t>(i) * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double app = 0, aqq = 0, apq = 0;
                for (int i = 0; i < m; ++i) {
                    const double x = A[static_cast<size_t>(i) * n + p], y = A[static_cast<size_t>(i) * n + q];
                    app += x * x; aqq += y * y; apq += x * y;
                }
                if (std::fabs(apq) <= 1e-300 || std::fabs(apq) <= 1e-17 * std::sqrt(app * aqq)) continue;
                off = std::fmax(off, std::fabs(apq) / std::sqrt(app * aqq + 1e-300));
                const double tau = (aqq - app) / (2.0 * apq);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < m; ++i) {
                    double& x = A[static_cast<size_t>(i) * n + p];
                    double& y = A[static_cast<size_t>(i) * n + q];
                    const double nx = c * x - s * y, ny = s * x + c * y;
                    x = nx; y = ny;
                }
                for (int i = 0; i < n; ++i) {
                    double& x = V[static_cast<size_t>(i) * n + p];
                    double& y = V[static_cast<size_t>(i) * n + q];
                    const double nx = c * x - s * y, ny = s * x + c * y;
                    x = nx; y = ny;
                }
            }
        if (off < 1e-15) break;
    }
}

}  // namespace

void plan_separable(int nx, int ny, int order, const double* coef, const float* weights, SepPlan* plan)
{
    std::memset(plan, 0, sizeof(*plan));
    plan->nx = nx; plan->ny = ny; plan->rank = 0;
    const int ww = 2 * nx + 1, wh = 2 * ny + 1;

    // exact (double) surface from the polynomial coefficients; parity in x / y read off the table
    std::vector<double> W(static_cast<size_t>(wh) * ww);
    double wmax = 0.0, sum_abs = 0.0;
    for (int y = -ny; y <= ny; ++y)
        for (int x = -nx; x <= nx; ++x) {
            double s = 0.0;
            for (int tot = 0; tot <= order; ++tot)
                for (int j = 0; j <= tot; ++j)
                    s += coef[tot * (tot + 1) / 2 + j] * std::pow(static_cast<double>(x), tot - j) * std::pow(static_cast<double>(y), j);
            W[static_cast<size_t>(y + ny) * ww + (x + nx)] = s;
            wmax = std::fmax(wmax, std::fabs(s));
            sum_abs += std::fabs(s);
        }
    if (!(wmax > 0.0)) return;

    auto parity = [&](bool along_x) -> int {  // +1 even, -1 odd, 0 neither
        double even = 0, odd = 0;
        for (int y = 0; y < wh; ++y)
            for (int x = 0; x < ww; ++x) {
                const double a = W[static_cast<size_t>(y) * ww + x];
                const double b = along_x ? W[static_cast<size_t>(y) * ww + (ww - 1 - x)] : W[static_cast<size_t>(wh - 1 - y) * ww + x];
                even = std::fmax(even, std::fabs(a - b));
                odd = std::fmax(odd, std::fabs(a + b));
            }
        if (even <= 1e-12 * wmax) return 1;
        if (odd <= 1e-12 * wmax) return -1;
        return 0;
    };
    const int px = parity(true), py = parity(false);
    if (px == 0 || py == 0) return;

    std::vector<double> A = W, V;
    jacobi_svd(A, wh, ww, V);
    // singular values / order
    std::vector<double> sig(ww);
    std::vector<int> idx(ww);
    for (int j = 0; j < ww; ++j) {
        double s = 0;
        for (int i = 0; i < wh; ++i) s += A[static_cast<size_t>(i) * ww + j] * A[static_cast<size_t>(i) * ww + j];
        sig[j] = std::sqrt(s);
        idx[j] = j;
    }
    for (int a = 0; a < ww; ++a)
        for (int b = a + 1; b < ww; ++b)
            if (sig[idx[b]] > sig[idx[a]]) { const int t = idx[a]; idx[a] = idx[b]; idx[b] = t; }
    int rank = 0;
    while (rank < ww && rank < wh && sig[idx[rank]] > 1e-10 * sig[idx[0]]) ++rank;
    if (rank == 0 || rank > kMaxRank) return;

    for (int r = 0; r < rank; ++r) {
        const int j = idx[r];
        const double s = sig[j], rs = std::sqrt(s);
        double rowv[33], colv[33];
        for (int x = 0; x < ww; ++x) rowv[x] = V[static_cast<size_t>(x) * ww + j] * rs;
        for (int y = 0; y < wh; ++y) colv[y] = A[static_cast<size_t>(y) * ww + j] / s * rs;
        // exact parity: average each factor with its mirrored self
        for (int x = 0; x <= nx; ++x) {
            const double m = 0.5 * (rowv[x] + px * rowv[ww - 1 - x]);
            rowv[x] = m; rowv[ww - 1 - x] = px * m;
        }
        for (int y = 0; y <= ny; ++y) {
            const double m = 0.5 * (colv[y] + py * colv[wh - 1 - y]);
            colv[y] = m; colv[wh - 1 - y] = py * m;
        }
        if (px < 0) rowv[nx] = 0.0;
        if (py < 0) colv[ny] = 0.0;
        for (int x = 0; x < ww; ++x) plan->row[r][x] = static_cast<float>(rowv[x]);
        for (int y = 0; y < wh; ++y) plan->col[r][y] = static_cast<float>(colv[y]);
    }

    // acceptance: the fp32 factors must reproduce the fp32 table of the reference.  The output error
    // this adds is bounded by sum|dW| * max|image|; the parity budget is 1e-6 * max|image| (times the
    // common scale), so require sum|dW| <= 2e-7 (relative to a unit-gain table: sum|W|).
    double err_sum = 0.0, err_max = 0.0;
    for (int y = 0; y < wh; ++y)
        for (int x = 0; x < ww; ++x) {
            double s = 0.0;
            for (int r = 0; r < rank; ++r) s += static_cast<double>(plan->col[r][y]) * static_cast<double>(plan->row[r][x]);
            const double d = std::fabs(s - static_cast<double>(weights[static_cast<size_t>(y) * ww + x]));
            err_sum += d;
            err_max = std::fmax(err_max, d);
        }
    plan->max_err = static_cast<float>(err_max / wmax);
    plan->sum_err = static_cast<float>(err_sum);
    plan->parity_x = px;
    plan->parity_y = py;
    (void)sum_abs;
    if (err_sum <= 4e-7) plan->rank = rank;

    // Additive surfaces: W(y,x) = u(x) + v(y).  The total-degree <= 3 fits evaluated at the window centre
    // (smoothing, d2/dx2, d2/dy2, their sum) only contain 1, x^2, y^2, so the rank-2 table is a sum of two
    // one-variable functions and both "other" factors are constant 1 -- box sums for the kernel
    // (sg2d_add.cu).  The constant is split so that v vanishes at the window ends (two taps less).
    static const bool no_additive = [] { const char* e = std::getenv("SAVGOL_B200_NO_ADDITIVE"); return e && e[0] == '1'; }();
    if (rank == 2 && px == 1 && py == 1 && nx == ny && nx <= 16 && !no_additive) {
        const double wcc = W[static_cast<size_t>(ny) * ww + nx];
        double dev = 0.0;
        for (int y = 0; y < wh; ++y)
            for (int x = 0; x < ww; ++x)
                dev = std::fmax(dev, std::fabs(W[static_cast<size_t>(y) * ww + x] -
                                               (W[static_cast<size_t>(ny) * ww + x] + W[static_cast<size_t>(y) * ww + nx] - wcc)));
        if (dev <= 1e-12 * wmax) {
            float u[33], v[33];
            const double v_end = W[nx];   // W(y = -ny, x = 0)
            for (int x = 0; x < ww; ++x) u[x] = static_cast<float>(W[static_cast<size_t>(ny) * ww + x] - wcc + v_end);
            for (int y = 0; y < wh; ++y) v[y] = static_cast<float>(W[static_cast<size_t>(y) * ww + nx] - v_end);
            v[0] = v[wh - 1] = 0.0f;
            double esum = 0.0, emax = 0.0;
            for (int y = 0; y < wh; ++y)
                for (int x = 0; x < ww; ++x) {
                    const double d = std::fabs(static_cast<double>(u[x]) + static_cast<double>(v[y]) -
                                               static_cast<double>(weights[static_cast<size_t>(y) * ww + x]));
                    esum += d;
                    emax = std::fmax(emax, d);
                }
            if (esum <= 4e-7) {
                std::memset(plan->row, 0, sizeof(plan->row));
                std::memset(plan->col, 0, sizeof(plan->col));
                for (int x = 0; x < ww; ++x) plan->row[0][x] = u[x];
                for (int y = 0; y < wh; ++y) plan->col[0][y] = v[y];
                plan->rank = 2;
                plan->additive = 1;
                plan->max_err = static_cast<float>(emax / wmax);
                plan->sum_err = static_cast<float>(esum);
            }
        }
    }
}

}  // namespace sg2d
