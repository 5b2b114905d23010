// coeffs.h -- host coefficient generation (see coeffs.cpp).
#pragma once

namespace sgc {

// ref: src/savgolFilter.c:639-677.  *why receives a static message when invalid.
bool config1d_valid(int n, int m, int d, float dt, const char** why);
// center[65], edge[32*65] (row e = target n-e); entries beyond 2n+1 / n rows are left untouched.
void weights1d(int n, int m, int d, float* center, float* edge);
// powf(dt, d), ref: src/savgolFilter.c:707
float dt_scale(float dt, int d);

// ref: src/savgol2d.c:271-302
bool config2d_valid(int nx, int ny, int order, int dx, int dy, float hx, float hy);
// ref: src/savgol2d.c:320-322
float scale2d(int dx, int dy, float hx, float hy);
// weights[(2ny+1)*(2nx+1)] row-major; coef_out (optional, 28 doubles) = polynomial coefficients of
// the weight surface including the dx!dy! factor.  false when the normal equations are singular.
bool weights2d(int nx, int ny, int order, int dx, int dy, float* weights, double* coef_out);

}  // namespace sgc
