// One-off key-setup kernels for the nearest-plane engine (per key, not per target):
// from the short basis S (columns b_j) and its GSO S~ (columns b~_i) derive
//   d_i   = ||b~_i||^2
//   Mt    = rows b~_i / d_i                 (centre -> GSO coordinates)
//   St    = S^t                              (rows b_j)
//   U     = Mt * S  (mu-coefficients, upper unitriangular)   -- via gemm_f64
// which turn GPV08 SampleD (MatZ::sample_d_precomputed_gso, gpv.rs:160) into
// "one GEMM + triangular back-substitution with randomized rounding".
#include "common.cuh"
#include "kernels.h"

namespace {

__global__ void colnorm2_kernel(const double* __restrict__ G, long ld, int rows, int cols, double* __restrict__ d) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double acc = 0;
    for (int r = 0; r < rows; ++r) {
        double v = G[(long)r * ld + c];
        acc += v * v;
    }
    d[c] = acc;
}

// out[c][r] = in[r][c] * (scale ? 1/scale[c] : 1)
__global__ void transpose_scale_kernel(const double* __restrict__ in, long ldin, double* __restrict__ out, long ldout,
                                       int rows, int cols, const double* __restrict__ scale) {
    __shared__ double tile[32][33];
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(long)r * ldin + c] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < rows) {
            double v = tile[threadIdx.x][i];
            if (scale) v /= scale[c];
            out[(long)c * ldout + r] = v;
        }
    }
}

__global__ void gather_cols_kernel(const double* __restrict__ in, long ldin, const int* __restrict__ cols, int ncols,
                                   double* __restrict__ out, long ldout, int rows) {
    long total = (long)rows * ncols;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r = i / ncols;
        int j = (int)(i - r * ncols);
        out[r * ldout + j] = in[r * ldin + cols[j]];
    }
}

__global__ void make_dg_kernel(const double* __restrict__ d, int count, double s, DGaussParams* __restrict__ dg) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dg[i] = make_dgauss(s / sqrt(d[i]));
}

}  // namespace

cudaError_t qf_launch_colnorm2(const double* G, long ld, int rows, int cols, double* d, cudaStream_t stream) {
    colnorm2_kernel<<<(cols + 127) / 128, 128, 0, stream>>>(G, ld, rows, cols, d);
    return cudaGetLastError();
}
cudaError_t qf_launch_transpose_scale(const double* in, long ldin, double* out, long ldout, int rows, int cols,
                                      const double* scale, cudaStream_t stream) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_scale_kernel<<<grid, block, 0, stream>>>(in, ldin, out, ldout, rows, cols, scale);
    return cudaGetLastError();
}
cudaError_t qf_launch_gather_cols(const double* in, long ldin, const int* cols, int ncols, double* out, long ldout,
                                  int rows, cudaStream_t stream) {
    long long total = (long long)rows * ncols;
    long long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    gather_cols_kernel<<<(int)g, 256, 0, stream>>>(in, ldin, cols, ncols, out, ldout, rows);
    return cudaGetLastError();
}
cudaError_t qf_launch_make_dg(const double* d, int count, double s, DGaussParams* dg, cudaStream_t stream) {
    make_dg_kernel<<<(count + 127) / 128, 128, 0, stream>>>(d, count, s, dg);
    return cudaGetLastError();
}
