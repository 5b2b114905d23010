// sg2d.h -- launch descriptors of the 2D kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sg2d {

// ref: include/iterative/savgol2d.h:108-112
enum : int { B_VALID = 0, B_CONSTANT = 1, B_REFLECT = 2 };

// out[img][oy][ox] = scale * sum_{wy,wx} W[wy][wx] * in[img][map(oy+cy-ny+wy)][map(ox+cx-nx+wx)]
// for oy < out_rows, ox < out_cols.  Full-size modes: cy = cx = 0, out = rows x cols, map = clamp /
// reflect.  VALID: cy = ny, cx = nx, out = (rows-2ny) x (cols-2nx), map is the identity.
struct Args2D {
    const float* in;
    float* out;
    float* out1;            // multi-output launches (sg2d_multi.cu): images of the 2nd / 3rd component, same geometry,
    float* out2;            // pitch and 16-byte phase as `out`
    const float* weights;   // device, [2ny+1][2nx+1]
    int rows, cols;         // input image size
    int out_rows, out_cols;
    int cy, cx;
    int nx, ny;
    long long in_stride, out_stride;            // elements between rows
    long long in_image_pitch, out_image_pitch;  // elements between images
    long long n_images;
    int boundary;
    float scale;
    int use_tma;            // separable kernel: the launcher encoded a tensor map of the input (aligned images only)
    int row0;               // image row of the buffer's first row (bands of a larger image; 0 otherwise)
    int band_rows;          // separable kernel: output rows per work item (set by the launcher)
    unsigned* counter;      // separable kernel: work-item ticket counter, zeroed in stream order before the launch
};

cudaError_t launch_direct(const Args2D& a, bool exact, cudaStream_t stream);

// Separable representation of the weight surface: W[y][x] = sum_r col[r][y] * row[r][x], r < rank
// (exact up to fp32 rounding of the factors, because W is a polynomial of total degree <= 6 in
// (x,y); see sg2d_sep.cu).  rank == 0 means "no usable factorisation, use the direct kernel".
constexpr int kMaxRank = 4;
struct SepPlan {
    int rank;
    int nx, ny;
    float row[kMaxRank][33];  // row[r][x + nx]
    float col[kMaxRank][33];  // col[r][y + ny]
    float max_err;            // max |W - sum_r col*row| / max|W|
    float sum_err;            // sum |W - sum_r col*row|  (bounds the extra output error per unit max|image|)
    int parity_x, parity_y;   // +1: factors even in x / y, -1: odd
    int additive;             // 1: W(y,x) = row[0][x] + col[0][y] with col[0][0] = col[0][2ny] = 0 (rank is 2, the
                              // "1" factors are implicit); runs sg2d_add.cu.  Square windows up to 17x17 only.
};
void plan_separable(int nx, int ny, int order, const double* coef, const float* weights, SepPlan* plan);
bool separable_supported(const Args2D& a, const SepPlan& plan);
cudaError_t launch_separable(const Args2D& a, const SepPlan& plan, cudaStream_t stream);
cudaError_t launch_additive(const Args2D& a, const SepPlan& plan, cudaStream_t stream);   // sg2d_add.cu
// One launch for `n_out` (2 or 3) filters over the same images: the components of a gradient / Hessian (sg2d_multi.cu).
// a.out / a.out1 / a.out2 receive the results; each plan carries its own scale in scales[i].
bool multi_supported(const Args2D& a, const SepPlan* const* plans, int n_out);
cudaError_t launch_multi(const Args2D& a, const SepPlan* const* plans, const float* scales, int n_out, cudaStream_t stream);
// Ticket counter of one launch of the streaming kernels (sg2d_sep.cu): 4 zeroed bytes, stream ordered.
cudaError_t acquire_counter(cudaStream_t stream, unsigned** out);

}  // namespace sg2d
