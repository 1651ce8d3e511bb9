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

// ---- fixed-point digit planes of the mu-coefficients for the tensor-core nearest-plane updates ----
// For every row i and every `blk`-column block c with i < c*blk (the entries the big off-diagonal
// updates use): U[i][j] * 2^e(i,c) rounded to an integer of L balanced base-256 digits;
// scale[c*D + i] = 2^-e(i,c).  Rows at or below the block keep zero digits.
namespace {

__global__ void ozaki_scale_kernel(const double* __restrict__ U, long ld, int D, int blk, int L, double* __restrict__ scale) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int nblk = (D + blk - 1) / blk;
    const long total = (long)D * nblk;
    for (long w = (long)blockIdx.x * wpb + (threadIdx.x >> 5); w < total; w += (long)gridDim.x * wpb) {
        const int c = (int)(w / D), i = (int)(w - (long)c * D);
        double mx = 0.0;
        if (i < c * blk) {
            const int j1 = min(D, (c + 1) * blk);
            for (int j = c * blk + lane; j < j1; j += 32) mx = fmax(mx, fabs(U[(long)i * ld + j]));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) {
            // |U| 2^e < 2^(8L-2)  (capacity of L balanced digits is ~0.498 * 256^L)
            int e = (mx > 0.0) ? (8 * L - 2) - (ilogb(mx) + 1) : 0;
            scale[(long)c * D + i] = ldexp(1.0, -e);
        }
    }
}

__global__ void ozaki_digits_kernel(const double* __restrict__ U, long ld, int D, int blk, int L,
                                    const double* __restrict__ scale, int8_t* __restrict__ planes, long plane_stride,
                                    long ldk) {
    const long total = (long)D * D;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int i = (int)(t / D), j = (int)(t - (long)i * D);
        const int c = j / blk;
        long long v = 0;
        if (i < c * blk) v = __double2ll_rn(U[(long)i * ld + j] / scale[(long)c * D + i]);
        for (int l = 0; l < L; ++l) {
            long long lo = ((v + 128) & 255) - 128;
            if (l == L - 1) lo = v;
            planes[l * plane_stride + (long)i * ldk + j] = (int8_t)lo;
            v = (v - lo) >> 8;
        }
    }
}

}  // namespace

cudaError_t qf_launch_ozaki_prepare(const double* U, long ld, int D, int blk, int L, double* scale, int8_t* planes,
                                    long plane_stride, long ldk, cudaStream_t stream) {
    const int nblk = (D + blk - 1) / blk;
    long long warps = (long long)D * nblk;
    long long g = (warps + 7) / 8;
    if (g > 148 * 32) g = 148 * 32;
    ozaki_scale_kernel<<<(int)g, 256, 0, stream>>>(U, ld, D, blk, L, scale);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ozaki_digits_kernel<<<148 * 16, 256, 0, stream>>>(U, ld, D, blk, L, scale, planes, plane_stride, ldk);
    return cudaGetLastError();
}
