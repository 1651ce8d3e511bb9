// Conversion, recombination and elementwise sampler kernels.
// All are HBM-streaming: flat grid-stride loops, lanes on the contiguous
// (coordinate) dimension, one warp per target row where a row reduction is needed.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int TPB = 256;

inline int grid_for(long long work, int per_block, int max_blocks = 148 * 16) {
    long long g = (work + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int)g;
}

// ---- conversions -----------------------------------------------------------
__global__ void i32_to_f64_kernel(const int32_t* __restrict__ in, long ldin, double* __restrict__ out, long ldout,
                                  int B, int M, unsigned long long* __restrict__ norm2) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (long b = (long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); b < B;
         b += (long)gridDim.x * warps_per_block) {
        const int32_t* src = in + b * ldin;
        double* dst = out + b * ldout;
        NormAcc acc;
        acc.clear();
        for (int j = lane; j < M; j += 32) {
            int32_t v = src[j];
            dst[j] = (double)v;
            acc.add_sq(v);
        }
        if (norm2) {
            acc.warp_reduce();
            if (lane == 0) norm2[b] = acc.value();
        }
    }
}

__global__ void i64_to_f64_kernel(const int64_t* __restrict__ in, long ldin, double* __restrict__ out, long ldout,
                                  int B, int M, double scale) {
    long total = (long)B * M;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long b = i / M;
        int j = (int)(i - b * M);
        out[b * ldout + j] = scale * (double)in[b * ldin + j];
    }
}

__global__ void f64_to_i32_kernel(const double* __restrict__ in, long ldin, int32_t* __restrict__ out, long ldout,
                                  int B, int M, int* flag) {
    long total = (long)B * M;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long b = i / M;
        int j = (int)(i - b * M);
        double v = in[b * ldin + j];
        double r = rint(v);
        if (r != v || fabs(r) > 2147483647.0) {
            if (flag) atomicOr(flag, 1);
            r = fmax(fmin(r, 2147483647.0), -2147483647.0);
        }
        out[b * ldout + j] = (int32_t)r;
    }
}

__global__ void domain_flags_kernel(const unsigned long long* __restrict__ norm2, unsigned long long bound,
                                    uint8_t* __restrict__ flags, int B) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x)
        flags[b] = norm2[b] <= bound ? 1 : 0;
}

// ---- recombination of exact fp64 partial products --------------------------
template <typename OutT>
__global__ void combine_kernel(CombineArgs a, OutT* __restrict__ out, long ldout, int B, int N) {
    long total = (long)B * N;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long b = i / N;
        int n = (int)(i - b * N);
        __int128 v = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < a.nacc) {
                long long t = __double2ll_rn(a.acc[c][b * a.ldacc + n]);
                v += ((__int128)t) << a.shift[c];
            }
        }
        if (a.acc_sign < 0) v = -v;
        if (a.base) v += (__int128)a.base[b * a.ldbase + n];
        if (a.q) {
            out[b * ldout + n] = (OutT)mod_i128(v, a.q);
        } else {
            out[b * ldout + n] = (OutT)(long long)v;
        }
    }
}

__global__ void combine_i32_kernel(CombineArgs a, int32_t* __restrict__ out, long ldout, int B, int N, int* flag) {
    long total = (long)B * N;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long b = i / N;
        int n = (int)(i - b * N);
        __int128 v = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < a.nacc) {
                long long t = __double2ll_rn(a.acc[c][b * a.ldacc + n]);
                v += ((__int128)t) << a.shift[c];
            }
        }
        if (a.acc_sign < 0) v = -v;
        if (a.base) v += (__int128)a.base[b * a.ldbase + n];
        if (v > 2147483647 || v < -2147483647) {
            if (flag) atomicOr(flag, 1);
            v = 0;
        }
        out[b * ldout + n] = (int32_t)(long long)v;
    }
}

__global__ void finalize_pert_kernel(const double* __restrict__ P, long ldp, const double* __restrict__ Zb, long ldz,
                                     int32_t* __restrict__ out, long ldo, int B, int M, int split, int* flag) {
    long total = (long)B * M;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long b = i / M;
        int j = (int)(i - b * M);
        double v = P[b * ldp + j];
        if (j >= split) v += Zb[b * ldz + (j - split)];
        double r = rint(v);
        if (r != v || fabs(r) > 2147483647.0) {
            if (flag) atomicOr(flag, 1);
            r = 0;
        }
        out[b * ldo + j] = (int32_t)r;
    }
}

__global__ void split_chunks_kernel(const double* __restrict__ in, long ldin, double* o0, double* o1, double* o2,
                                    double* o3, int nchunks, int bits, long ldout, int B, int M) {
    long total = (long)B * M;
    const double scale = ldexp(1.0, bits), inv = ldexp(1.0, -bits);
    double* outs[4] = {o0, o1, o2, o3};
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long b = i / M;
        int j = (int)(i - b * M);
        double rem = in[b * ldin + j];
        for (int c = 0; c < nchunks; ++c) {
            if (c == nchunks - 1) {
                outs[c][b * ldout + j] = rem;
            } else {
                double hi = rint(rem * inv);
                outs[c][b * ldout + j] = rem - hi * scale;
                rem = hi;
            }
        }
    }
}

__global__ void scatter_cols_kernel(const double* __restrict__ src, long ldsrc, const int* __restrict__ cols,
                                    int ncols, double* __restrict__ dst, long lddst, int B, double scale) {
    long total = (long)B * ncols;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long b = i / ncols;
        int j = (int)(i - b * ncols);
        dst[b * lddst + cols[j]] = scale * src[b * ldsrc + j];
    }
}

// ---- samplers --------------------------------------------------------------
__global__ void normal_fill_kernel(double* __restrict__ out, long ld, int B, int M, uint64_t seed,
                                   uint64_t first_target, uint32_t tag) {
    const int half = (M + 1) >> 1;
    long total = (long)B * half;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long b = i / half;
        int p = (int)(i - b * half);
        Philox rng;
        rng.init(seed, (first_target + (uint64_t)b) * (uint64_t)half + (uint64_t)p, tag);
        float n0, n1;
        rng.normal2(n0, n1);
        double* row = out + b * ld;
        row[2 * p] = (double)n0;
        if (2 * p + 1 < M) row[2 * p + 1] = (double)n1;
    }
}

__global__ void dgauss_kernel(const double* __restrict__ center, long ldc, double* __restrict__ out_f64, long ldo,
                              int32_t* __restrict__ out_i32, long ldoi, int B, int M, DGaussParams dg,
                              uint64_t seed, uint64_t first_target, uint32_t tag, int* flag) {
    long total = (long)B * M;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long b = i / M;
        int j = (int)(i - b * M);
        Philox rng;
        rng.init(seed, (first_target + (uint64_t)b) * (uint64_t)M + (uint64_t)j, tag);
        double c = center ? center[b * ldc + j] : 0.0;
        double z = sample_dgauss(dg, c, rng, flag);
        if (out_f64) out_f64[b * ldo + j] = z;
        if (out_i32) out_i32[b * ldoi + j] = (int32_t)z;
    }
}

__global__ void uniform_modq_kernel(int64_t* __restrict__ out, long count, unsigned long long q, uint64_t seed,
                                    uint64_t first_index) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long)gridDim.x * blockDim.x) {
        Philox rng;
        rng.init(seed, first_index + (uint64_t)i, QF_STREAM_UNIFORM);
        uint64_t rh = ((uint64_t)rng.next() << 32) | rng.next();
        uint64_t rl = ((uint64_t)rng.next() << 32) | rng.next();
        // floor((rh*2^64 + rl) * q / 2^128): bias < q / 2^128
        uint64_t hi = __umul64hi(rh, q);
        uint64_t lo = rh * q;
        uint64_t carry = __umul64hi(rl, q);
        uint64_t s = lo + carry;
        if (s < lo) hi += 1;
        out[i] = (int64_t)hi;
    }
}

// R = U{0,1} - U{0,1} entrywise (trapdoor_distribution.rs:82-86): P(-1)=P(1)=1/4, P(0)=1/2.
// 16 entries per thread (one 128-bit Philox block gives 2 bits per entry for 64 entries;
// we use one block per 16 entries to keep the counter <-> entry map trivial).
__global__ void ternary_kernel(int8_t* __restrict__ out, long count, uint64_t seed) {
    long groups = (count + 15) / 16;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long)gridDim.x * blockDim.x) {
        Philox rng;
        rng.init(seed, (uint64_t)g, QF_STREAM_TERNARY);
        uint32_t bits = rng.next();
        long base = g * 16;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            if (base + t < count) {
                int a = (bits >> (2 * t)) & 1, b = (bits >> (2 * t + 1)) & 1;
                out[base + t] = (int8_t)(a - b);
            }
        }
    }
}

}  // namespace

cudaError_t qf_launch_i32_to_f64(const int32_t* in, long ldin, double* out, long ldout, int B, int M,
                                 unsigned long long* norm2, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    i32_to_f64_kernel<<<grid_for(B, TPB / 32), TPB, 0, stream>>>(in, ldin, out, ldout, B, M, norm2);
    return cudaGetLastError();
}
cudaError_t qf_launch_i64_to_f64(const int64_t* in, long ldin, double* out, long ldout, int B, int M, double scale,
                                 cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    i64_to_f64_kernel<<<grid_for((long long)B * M, TPB), TPB, 0, stream>>>(in, ldin, out, ldout, B, M, scale);
    return cudaGetLastError();
}
cudaError_t qf_launch_f64_to_i32(const double* in, long ldin, int32_t* out, long ldout, int B, int M, int* flag,
                                 cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    f64_to_i32_kernel<<<grid_for((long long)B * M, TPB), TPB, 0, stream>>>(in, ldin, out, ldout, B, M, flag);
    return cudaGetLastError();
}
cudaError_t qf_launch_domain_flags(const unsigned long long* norm2, unsigned long long bound, uint8_t* flags, int B,
                                   cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    domain_flags_kernel<<<grid_for(B, TPB), TPB, 0, stream>>>(norm2, bound, flags, B);
    return cudaGetLastError();
}
cudaError_t qf_launch_combine_i64(const CombineArgs& a, int64_t* out, long ldout, int B, int N, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    combine_kernel<int64_t><<<grid_for((long long)B * N, TPB), TPB, 0, stream>>>(a, out, ldout, B, N);
    return cudaGetLastError();
}
cudaError_t qf_launch_combine_f64(const CombineArgs& a, double* out, long ldout, int B, int N, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    combine_kernel<double><<<grid_for((long long)B * N, TPB), TPB, 0, stream>>>(a, out, ldout, B, N);
    return cudaGetLastError();
}
cudaError_t qf_launch_combine_i32(const CombineArgs& a, int32_t* out, long ldout, int B, int N, int* flag,
                                  cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    combine_i32_kernel<<<grid_for((long long)B * N, TPB), TPB, 0, stream>>>(a, out, ldout, B, N, flag);
    return cudaGetLastError();
}
cudaError_t qf_launch_finalize_pert(const double* P, long ldp, const double* Zb, long ldz, int32_t* out, long ldo,
                                    int B, int M, int split, int* flag, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    finalize_pert_kernel<<<grid_for((long long)B * M, TPB), TPB, 0, stream>>>(P, ldp, Zb, ldz, out, ldo, B, M, split,
                                                                            flag);
    return cudaGetLastError();
}
cudaError_t qf_launch_split_chunks(const double* in, long ldin, double* const* out, int nchunks, int bits, long ldout,
                                   int B, int M, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    double* o[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int c = 0; c < nchunks && c < 4; ++c) o[c] = out[c];
    split_chunks_kernel<<<grid_for((long long)B * M, TPB), TPB, 0, stream>>>(in, ldin, o[0], o[1], o[2], o[3], nchunks,
                                                                              bits, ldout, B, M);
    return cudaGetLastError();
}
cudaError_t qf_launch_scatter_cols_f64(const double* src, long ldsrc, const int* cols, int ncols, double* dst,
                                       long lddst, int B, double scale, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    scatter_cols_kernel<<<grid_for((long long)B * ncols, TPB), TPB, 0, stream>>>(src, ldsrc, cols, ncols, dst, lddst, B,
                                                                                scale);
    return cudaGetLastError();
}
cudaError_t qf_launch_normal_fill(double* out, long ld, int B, int M, uint64_t seed, uint64_t first_target,
                                  uint32_t tag, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    normal_fill_kernel<<<grid_for((long long)B * ((M + 1) / 2), TPB), TPB, 0, stream>>>(out, ld, B, M, seed,
                                                                                       first_target, tag);
    return cudaGetLastError();
}
cudaError_t qf_launch_dgauss(const double* center, long ldc, double* out_f64, long ldo, int32_t* out_i32, long ldoi,
                             int B, int M, double s, uint64_t seed, uint64_t first_target, uint32_t tag, int* flag,
                             cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    DGaussParams dg = make_dgauss(s);
    dgauss_kernel<<<grid_for((long long)B * M, TPB), TPB, 0, stream>>>(center, ldc, out_f64, ldo, out_i32, ldoi, B, M,
                                                                      dg, seed, first_target, tag, flag);
    return cudaGetLastError();
}
cudaError_t qf_launch_uniform_modq(int64_t* out, long count, unsigned long long q, uint64_t seed,
                                   uint64_t first_index, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    uniform_modq_kernel<<<grid_for(count, TPB), TPB, 0, stream>>>(out, count, q, seed, first_index);
    return cudaGetLastError();
}
cudaError_t qf_launch_ternary(int8_t* out, long count, uint64_t seed, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    ternary_kernel<<<grid_for((count + 15) / 16, TPB), TPB, 0, stream>>>(out, count, seed);
    return cudaGetLastError();
}

// ---- limb splitting for the tcgen05 int8 contraction ---------------------------------------
// value = sum_l 256^l d_l with balanced digits d_l in [-128,127]; L digits cover
// [-128 S, 127 S], S = (256^L - 1)/255.  planes: L matrices of B x ldk bytes.
namespace {

// nz (optional): zero-tile map, nz[l * nz_plane + nz_idx] = 1 when digit l is non-zero (benign race: all writers store 1)
__device__ __forceinline__ void split_digits(long long v, int L, int8_t* planes, long plane_stride, long off, int* flag,
                                             uint8_t* nz = nullptr, long nz_plane = 0, long nz_idx = 0) {
    for (int l = 0; l < L - 1; ++l) {
        long long lo = ((v + 128) & 255) - 128;
        planes[l * plane_stride + off] = (int8_t)lo;
        if (nz && lo != 0 && !nz[l * nz_plane + nz_idx]) nz[l * nz_plane + nz_idx] = 1;
        v = (v - lo) >> 8;
    }
    if ((v > 127 || v < -128) && flag) atomicOr(flag, 8);
    planes[(L - 1) * plane_stride + off] = (int8_t)v;
    if (nz && v != 0 && !nz[(L - 1) * nz_plane + nz_idx]) nz[(L - 1) * nz_plane + nz_idx] = 1;
}

// One warp per target row, 128 columns (= one zero-tile) per warp iteration, 4 consecutive columns per lane:
// one vector load, one char4 store per digit plane, and one warp vote per digit for the zero-tile map
// (no global reads on the flag path).  col0 and ldk are multiples of 128 / 4 for every caller.
// returns (warp-uniform, when nz is given) the index of the highest digit plane this warp found non-zero, -1 if none
__device__ __forceinline__ int split4(const long long v[4], int L, int8_t* planes, long plane_stride, long off, int* flag,
                                      uint8_t* nz, long nz_plane, long nz_idx, int valid, int lane) {
    long long w[4] = {v[0], v[1], v[2], v[3]};
    int top = -1;
    for (int l = 0; l < L; ++l) {
        int8_t dd[4];
        bool any = false;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            long long lo = ((w[t] + 128) & 255) - 128;
            if (l == L - 1) {
                lo = w[t];
                if ((lo > 127 || lo < -128) && flag && t < valid) atomicOr(flag, 8);
            }
            dd[t] = (int8_t)lo;
            any |= (t < valid) && (dd[t] != 0);
            w[t] = (w[t] - lo) >> 8;
        }
        if (valid == 4) {
            char4 d;
            d.x = dd[0]; d.y = dd[1]; d.z = dd[2]; d.w = dd[3];
            *reinterpret_cast<char4*>(planes + l * plane_stride + off) = d;
        } else {
            for (int t = 0; t < valid; ++t) planes[l * plane_stride + off + t] = dd[t];
        }
        if (nz) {
            const bool warp_any = __any_sync(0xffffffffu, any);
            if (warp_any && lane == 0) nz[l * nz_plane + nz_idx] = 1;
            if (warp_any) top = l;
        }
    }
    return top;
}

// gate0..2 (optional device ints): raised (atomicMax) to the index of the highest non-zero digit plane written by this
// launch -- the contractions that consume these planes pick, on the device, the variant compiled for that many digits
__global__ void split_f64_limbs_kernel(const double* __restrict__ in, long ldin, int8_t* __restrict__ planes,
                                       long plane_stride, long ldk, int B, int M, int L, int* flag, uint8_t* nz,
                                       int nz_m_tiles, int nz_kb_total, int col0, int* gate0, int* gate1, int* gate2, int* gate3) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int iters = (M + 127) >> 7;
    const long total = (long)B * iters;  // (row, 128-column tile) pairs, one per warp iteration
    for (long w = (long)blockIdx.x * wpb + (threadIdx.x >> 5); w < total; w += (long)gridDim.x * wpb) {
        const long b = w / iters;
        const int j = ((int)(w - b * iters) << 7) + lane * 4;
        const int valid = max(0, min(4, M - j));
        long long v[4] = {0, 0, 0, 0};
        const double* src = in + b * ldin + j;
        if (valid == 4 && ((((uintptr_t)src) & 15) == 0)) {
            const double2 a0 = *reinterpret_cast<const double2*>(src), a1 = *reinterpret_cast<const double2*>(src + 2);
            const double x[4] = {a0.x, a0.y, a1.x, a1.y};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                double xx = x[t];
                if (!(fabs(xx) < 9.0e18)) { if (flag) atomicOr(flag, 8); xx = 0; }
                v[t] = __double2ll_rn(xx);
            }
        } else {
            for (int t = 0; t < valid; ++t) {
                double xx = src[t];
                if (!(fabs(xx) < 9.0e18)) { if (flag) atomicOr(flag, 8); xx = 0; }
                v[t] = __double2ll_rn(xx);
            }
        }
        const int top = split4(v, L, planes, plane_stride, b * ldk + j, flag, nz, (long)nz_m_tiles * nz_kb_total,
                               (b >> 7) * nz_kb_total + ((col0 + j) >> 7), valid, lane);
        if (lane == 0 && top > 0) {  // plain read first: almost every warp finds the gate already high enough
            if (gate0 && *(volatile int*)gate0 < top) atomicMax(gate0, top);
            if (gate1 && *(volatile int*)gate1 < top) atomicMax(gate1, top);
            if (gate2 && *(volatile int*)gate2 < top) atomicMax(gate2, top);
            if (gate3 && *(volatile int*)gate3 < top) atomicMax(gate3, top);
        }
    }
}

// 32-bit variant of split4 for int32 inputs (values beyond +-2^30 belong to out-of-domain targets: clamped)
__device__ __forceinline__ void split4_i32(const int v[4], int L, int8_t* planes, long plane_stride, long off,
                                           uint8_t* nz, long nz_plane, long nz_idx, int valid, int lane) {
    int w[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) w[t] = max(-(1 << 30), min(1 << 30, v[t]));
    for (int l = 0; l < L; ++l) {
        int dd[4];
        int any = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            int lo = ((w[t] + 128) & 255) - 128;
            if (l == L - 1) lo = w[t];
            dd[t] = lo;
            any |= (t < valid) ? (lo & 255) : 0;
            w[t] = (w[t] - lo) >> 8;
        }
        if (valid == 4) {
            const unsigned pk = (dd[0] & 255) | ((dd[1] & 255) << 8) | ((dd[2] & 255) << 16) | ((unsigned)(dd[3] & 255) << 24);
            *reinterpret_cast<unsigned*>(planes + l * plane_stride + off) = pk;
        } else {
            for (int t = 0; t < valid; ++t) planes[l * plane_stride + off + t] = (int8_t)dd[t];
        }
        if (nz) {
            const bool warp_any = __any_sync(0xffffffffu, any != 0);
            if (warp_any && lane == 0) nz[l * nz_plane + nz_idx] = 1;
        }
    }
}

__global__ void __launch_bounds__(256)
split_i32_limbs_kernel(const int32_t* __restrict__ in, long ldin, int8_t* __restrict__ planes,
                       long plane_stride, long ldk, int B, int M, int L,
                       unsigned long long* __restrict__ norm2, uint8_t* nz, int nz_m_tiles,
                       int nz_kb_total) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const bool vec = ((ldin & 3) == 0) && ((((uintptr_t)in) & 15) == 0);
    const int iters = (M + 127) >> 7;
    for (long b = (long)blockIdx.x * wpb + (threadIdx.x >> 5); b < B; b += (long)gridDim.x * wpb) {
        const int32_t* src = in + b * ldin;
        NormAcc acc;
        acc.clear();
        const long nz_plane = (long)nz_m_tiles * nz_kb_total, nz_row = (b >> 7) * nz_kb_total;
        int it = 0;
        // two 128-column tiles per trip: both vector loads are issued before either is consumed
        for (; it + 1 < iters && vec && ((it + 2) << 7) <= M; it += 2) {
            const int j0 = (it << 7) + lane * 4, j1 = j0 + 128;
            const int4 q0 = __ldcs(reinterpret_cast<const int4*>(src + j0));
            const int4 q1 = __ldcs(reinterpret_cast<const int4*>(src + j1));
            const int v0[4] = {q0.x, q0.y, q0.z, q0.w}, v1[4] = {q1.x, q1.y, q1.z, q1.w};
            acc.add_sq4(v0[0], v0[1], v0[2], v0[3]);
            acc.add_sq4(v1[0], v1[1], v1[2], v1[3]);
            split4_i32(v0, L, planes, plane_stride, b * ldk + j0, nz, nz_plane, nz_row + (j0 >> 7), 4, lane);
            split4_i32(v1, L, planes, plane_stride, b * ldk + j1, nz, nz_plane, nz_row + (j1 >> 7), 4, lane);
        }
        for (; it < iters; ++it) {
            const int j = (it << 7) + lane * 4;
            const int valid = max(0, min(4, M - j));
            int v[4] = {0, 0, 0, 0};
            for (int t = 0; t < valid; ++t) v[t] = src[j + t];
            acc.add_sq4(v[0], v[1], v[2], v[3]);
            split4_i32(v, L, planes, plane_stride, b * ldk + j, nz, nz_plane, nz_row + (j >> 7), valid, lane);
        }
        if (norm2) {
            acc.warp_reduce();
            if (lane == 0) norm2[b] = acc.value();
        }
    }
}

__global__ void add_cols_i32_kernel(int32_t* __restrict__ e, long lde, const double* __restrict__ sol, long ldsol,
                                    const int* __restrict__ cols, int ncols, int B) {
    long total = (long)B * ncols;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long b = i / ncols;
        int j = (int)(i - b * ncols);
        e[b * lde + cols[j]] += (int32_t)__double2ll_rn(sol[b * ldsol + j]);
    }
}

}  // namespace

namespace {
__global__ void pert_xb_kernel(const double* __restrict__ G, long ldg, double* __restrict__ X2, long ldx,
                               int8_t* __restrict__ planes, long plane_stride, long ldk, int B, int mb, int nk,
                               double sqrt_beta, double fscale, int L, int* flag) {
    const long total = (long)B * nk;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i / nk;
        const int j = (int)(i - b * nk);
        const double xb = sqrt_beta * G[b * ldg + mb + j];
        X2[b * ldx + mb + j] = xb;
        split_digits(__double2ll_rn(xb * fscale), L, planes, plane_stride, b * ldk + j, flag);
    }
}
}  // namespace
namespace {
__global__ void pert_normal_digits_kernel(int B, int M, int split, uint64_t seed, uint64_t first_target,
                                          int8_t* __restrict__ gplanes, long gplane_stride, long ldkg, int LG, double gscale,
                                          double* __restrict__ X2, long ldx, int8_t* __restrict__ bplanes, long bplane_stride,
                                          long ldkb, int LB, double sqrt_beta, double bscale, int* flag) {
    const int half = (M + 1) >> 1;
    const long total = (long)B * half;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i / half;
        const int p = (int)(i - b * half);
        Philox rng;
        rng.init(seed, (first_target + (uint64_t)b) * (uint64_t)half + (uint64_t)p, QF_STREAM_PERT_NORMAL);
        float nn[2];
        rng.normal2(nn[0], nn[1]);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int j = 2 * p + t;
            if (j >= M) break;
            if (j < split) {
                split_digits(__double2ll_rn((double)nn[t] * gscale), LG, gplanes, gplane_stride, b * ldkg + j, flag);
            } else {
                const double xb = sqrt_beta * (double)nn[t];
                X2[b * ldx + j] = xb;
                split_digits(__double2ll_rn(xb * bscale), LB, bplanes, bplane_stride, b * ldkb + (j - split), flag);
            }
        }
    }
}
}  // namespace
cudaError_t qf_launch_pert_normal_digits(int B, int M, int split, uint64_t seed, uint64_t first_target, int8_t* gplanes,
                                         long gplane_stride, long ldkg, int LG, double gscale, double* X2, long ldx,
                                         int8_t* bplanes, long bplane_stride, long ldkb, int LB, double sqrt_beta, double bscale,
                                         int* flag, cudaStream_t stream) {
    if (B <= 0 || M <= 0) return cudaSuccess;
    pert_normal_digits_kernel<<<grid_for((long long)B * ((M + 1) / 2), TPB), TPB, 0, stream>>>(
        B, M, split, seed, first_target, gplanes, gplane_stride, ldkg, LG, gscale, X2, ldx, bplanes, bplane_stride, ldkb, LB,
        sqrt_beta, bscale, flag);
    return cudaGetLastError();
}
namespace {
// I2[b][row] += sum_c S'[row][c] z1[b][c] for the block-diagonal gadget basis S' = (I_n (x) S_k), columns reversed
// when `reversed` (short_basis_classical.rs:80-82): only the k columns of the row's own block contribute.
__global__ void sprime_apply_kernel(const double* __restrict__ Z, long ldz, double* __restrict__ I2, long ldi, int B, int nk,
                                    int k, const double* __restrict__ sk, int reversed) {
    const long total = (long)B * nk;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i / nk;
        const int row = (int)(i - b * nk);
        const int blk = row / k, t = row - blk * k;
        const double* zr = Z + b * ldz;
        // row t of S_k has at most three non-zeros: the diagonal, the sub-diagonal -1 and (q not a power of the base)
        // the digit of q in the last column
        double acc = 0.0;
        const int cand[3] = {t - 1, t, k - 1};
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const int c = cand[e];
            if (c < 0 || (e == 2 && (c == t || c == t - 1))) continue;
            const double sv = sk[t * k + c];
            if (sv == 0.0) continue;
            const int cc = blk * k + c;                      // column of I_n (x) S_k
            const int col = reversed ? nk - 1 - cc : cc;     // where that column sits in S'
            acc = fma(sv, zr[col], acc);
        }
        I2[b * ldi + row] += acc;
    }
}
// e[b][i] += z2[b][i] (i < mb),  e[b][mb + row] = I2[b][row] (+ g3[b][row], two-phase form): the structured form of e = S z
__global__ void gpv_struct_finalize_kernel(int32_t* __restrict__ e, long lde, const double* __restrict__ Z2, long ldz,
                                           const double* __restrict__ I2, long ldi, int B, int mb, int nk, int* flag,
                                           const int8_t* __restrict__ g3, long ldg) {
    const int m = mb + nk;
    const long total = (long)B * m;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i / m;
        const int j = (int)(i - b * m);
        double v = j < mb ? (double)e[b * lde + j] + Z2[b * ldz + j]
                          : I2[b * ldi + (j - mb)] + (g3 ? (double)g3[b * ldg + (j - mb)] : 0.0);
        if (!(fabs(v) < 2147483647.0)) { if (flag) atomicOr(flag, 4); v = 0.0; }
        e[b * lde + j] = (int32_t)__double2ll_rn(v);
    }
}
}  // namespace
namespace {
// e_bot = g3 + S' z1 in one pass (two-phase form, api.cu samp_p_np2_chunk): four consecutive gadget rows per thread;
// writes the int32 values into e[b][mb + row] and their balanced base-256 digit planes (the x operand of R e_bot).
__global__ void gpv_ebot_kernel(const double* __restrict__ Z, long ldz, const int8_t* __restrict__ g3, long ldg,
                                int32_t* __restrict__ e, long lde, int mb, int8_t* __restrict__ planes, long plane_stride,
                                long ldk, int L, int B, int nk, int k, const double* __restrict__ sk, int reversed, int* flag) {
    const int nq = nk >> 2;
    const long total = (long)B * nq;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i / nq;
        const int r0 = (int)(i - b * nq) << 2;
        const double* zr = Z + b * ldz;
        const unsigned gw = *reinterpret_cast<const unsigned*>(g3 + b * ldg + r0);
        long long v[4];
#pragma unroll
        for (int t4 = 0; t4 < 4; ++t4) {
            const int row = r0 + t4, blk = row / k, t = row - blk * k;
            double acc = (double)(int8_t)((gw >> (8 * t4)) & 255);
            const int cand[3] = {t - 1, t, k - 1};
#pragma unroll
            for (int c3 = 0; c3 < 3; ++c3) {
                const int c = cand[c3];
                if (c < 0 || (c3 == 2 && (c == t || c == t - 1))) continue;
                const double sv = sk[t * k + c];
                if (sv == 0.0) continue;
                const int cc = blk * k + c;
                acc = fma(sv, zr[reversed ? nk - 1 - cc : cc], acc);
            }
            if (!(fabs(acc) < 2147483647.0)) { if (flag) atomicOr(flag, 4); acc = 0.0; }
            v[t4] = __double2ll_rn(acc);
        }
        *reinterpret_cast<int4*>(e + b * lde + mb + r0) = make_int4((int)v[0], (int)v[1], (int)v[2], (int)v[3]);
        for (int l = 0; l < L; ++l) {
            unsigned pk = 0;
#pragma unroll
            for (int t4 = 0; t4 < 4; ++t4) {
                long long d = ((v[t4] + 128) & 255) - 128;
                if (l == L - 1) {
                    d = v[t4];
                    if (d > 127 || d < -128) { if (flag) atomicOr(flag, 8); d = 0; }
                }
                v[t4] = (v[t4] - d) >> 8;
                pk |= (unsigned)(d & 255) << (8 * t4);
            }
            *reinterpret_cast<unsigned*>(planes + (long)l * plane_stride + b * ldk + r0) = pk;
        }
    }
}
}  // namespace
cudaError_t qf_launch_gpv_ebot(const double* Z, long ldz, const int8_t* g3, long ldg, int32_t* e, long lde, int mb,
                               int8_t* planes, long plane_stride, long ldk, int L, int B, int nk, int k, const double* sk,
                               int reversed, int* flag, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    // four rows per thread, 16-byte int32 stores, 4-byte digit stores
    if ((nk & 3) || (ldg & 3) || (ldk & 3) || (plane_stride & 3) || (lde & 3) || (mb & 3) || (((uintptr_t)e) & 15) ||
        (((uintptr_t)g3) & 3) || (((uintptr_t)planes) & 3))
        return cudaErrorInvalidValue;
    gpv_ebot_kernel<<<grid_for((long long)B * (nk / 4), TPB), TPB, 0, stream>>>(Z, ldz, g3, ldg, e, lde, mb, planes, plane_stride,
                                                                               ldk, L, B, nk, k, sk, reversed, flag);
    return cudaGetLastError();
}
cudaError_t qf_launch_sprime_apply(const double* Z, long ldz, double* I2, long ldi, int B, int nk, int k, const double* sk,
                                   int reversed, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    sprime_apply_kernel<<<grid_for((long long)B * nk, TPB), TPB, 0, stream>>>(Z, ldz, I2, ldi, B, nk, k, sk, reversed);
    return cudaGetLastError();
}
cudaError_t qf_launch_gpv_struct_finalize(int32_t* e, long lde, const double* Z2, long ldz, const double* I2, long ldi, int B,
                                          int mb, int nk, int* flag, cudaStream_t stream, const int8_t* g3, long ldg) {
    if (B <= 0) return cudaSuccess;
    gpv_struct_finalize_kernel<<<grid_for((long long)B * (mb + nk), TPB), TPB, 0, stream>>>(e, lde, Z2, ldz, I2, ldi, B, mb, nk,
                                                                                           flag, g3, ldg);
    return cudaGetLastError();
}
namespace {
// g3[b][blk * k + t] = digit t (base `base`) of h[b][blk]: the gadget-lattice coset representative with G g3 = h
// (gadget_classical.rs:169-182, find_solution_gadget_vec), one thread per (target, syndrome entry)
__global__ void gadget_digits_kernel(const int64_t* __restrict__ h, long ldh, int8_t* __restrict__ plane, long ldk, int B,
                                     int n, int k, unsigned base, uint8_t* __restrict__ nz, int nz_kb_total,
                                     double* __restrict__ I2, long ldi) {
    // one thread per four consecutive output digits (coalesced 4-byte / 32-byte stores); nk = n k is a multiple of 4 for
    // every caller (the structured path needs nk % 128 == 0), ragged tails take the scalar branch
    const int nk = n * k, nq = (nk + 3) >> 2;
    const long total = (long)B * nq;
    const bool pow2 = (base & (base - 1)) == 0;
    const int sh = 31 - __clz(base);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i / nq;
        const int c0 = (int)(i - b * nq) << 2;
        unsigned d[4] = {0, 0, 0, 0};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = c0 + e;
            if (c >= nk) break;
            const int blk = c / k, t = c - blk * k;
            unsigned long long v = (unsigned long long)h[b * ldh + blk];
            if (pow2) {
                d[e] = (t * sh < 64) ? (unsigned)((v >> (t * sh)) & (base - 1)) : 0u;
            } else {
                for (int r = 0; r < t; ++r) v /= base;
                d[e] = (unsigned)(v % base);
            }
        }
        if (c0 + 3 < nk && ((ldk & 3) == 0)) {
            *reinterpret_cast<unsigned*>(plane + b * ldk + c0) = d[0] | (d[1] << 8) | (d[2] << 16) | (d[3] << 24);
            if (I2) {
                double* o = I2 + b * ldi + c0;
                if (((ldi & 1) == 0) && ((((uintptr_t)I2) & 15) == 0)) {
                    reinterpret_cast<double2*>(o)[0] = make_double2((double)d[0], (double)d[1]);
                    reinterpret_cast<double2*>(o)[1] = make_double2((double)d[2], (double)d[3]);
                } else {
                    for (int e = 0; e < 4; ++e) o[e] = (double)d[e];
                }
            }
        } else {
            for (int e = 0; e < 4 && c0 + e < nk; ++e) {
                plane[b * ldk + c0 + e] = (int8_t)d[e];
                if (I2) I2[b * ldi + c0 + e] = (double)d[e];
            }
        }
    }
    if (nz) {  // plane 0 of the zero-tile map: every (128-target, 128-column) tile of the digit block is live
        const int kbn = (nk + 127) >> 7, mt = (B + 127) >> 7;
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (long)mt * kbn; i += (long)gridDim.x * blockDim.x)
            nz[(i / kbn) * nz_kb_total + (i % kbn)] = 1;
    }
}
}  // namespace
cudaError_t qf_launch_gadget_digits(const int64_t* h, long ldh, int8_t* plane, long ldk, int B, int n, int k, unsigned base,
                                    uint8_t* nz, int nz_kb_total, cudaStream_t stream, double* I2, long ldi) {
    if (B <= 0) return cudaSuccess;
    if (base < 2 || base > 128) return cudaErrorInvalidValue;
    gadget_digits_kernel<<<grid_for((long long)B * ((n * k + 3) / 4), TPB), TPB, 0, stream>>>(h, ldh, plane, ldk, B, n, k, base, nz,
                                                                                          nz_kb_total, I2, ldi);
    return cudaGetLastError();
}
cudaError_t qf_launch_pert_xb(const double* G, long ldg, double* X2, long ldx, int8_t* planes, long plane_stride, long ldk,
                              int B, int mb, int nk, double sqrt_beta, double fscale, int L, int* flag, cudaStream_t stream) {
    if (B <= 0 || nk <= 0) return cudaSuccess;
    pert_xb_kernel<<<grid_for((long long)B * nk, TPB), TPB, 0, stream>>>(G, ldg, X2, ldx, planes, plane_stride, ldk, B, mb, nk,
                                                                        sqrt_beta, fscale, L, flag);
    return cudaGetLastError();
}
cudaError_t qf_launch_split_f64_limbs(const double* in, long ldin, int8_t* planes, long plane_stride, long ldk, int B,
                                      int M, int L, int* flag, uint8_t* nz, int nz_m_tiles, int nz_kb_total, int col0,
                                      cudaStream_t stream, int* gate0, int* gate1, int* gate2, int* gate3) {
    if (B <= 0) return cudaSuccess;
    // char4 stores need 4-byte aligned plane addresses: planes pointer, ldk and col0 multiples of 4 (true for all callers)
    if ((((uintptr_t)planes) & 3) || (ldk & 3) || (col0 & 3)) return cudaErrorMisalignedAddress;
    if (nz && (col0 & 127)) return cudaErrorInvalidValue;
    split_f64_limbs_kernel<<<grid_for((long long)B * ((M + 127) / 128), TPB / 32), TPB, 0, stream>>>(
        in, ldin, planes, plane_stride, ldk, B, M, L, flag, nz, nz_m_tiles, nz_kb_total, col0, gate0, gate1, gate2, gate3);
    return cudaGetLastError();
}
cudaError_t qf_launch_split_i32_limbs(const int32_t* in, long ldin, int8_t* planes, long plane_stride, long ldk, int B,
                                      int M, int L, unsigned long long* norm2, uint8_t* nz, int nz_m_tiles,
                                      int nz_kb_total, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    split_i32_limbs_kernel<<<grid_for(B, TPB / 32), TPB, 0, stream>>>(in, ldin, planes, plane_stride, ldk, B, M, L, norm2,
                                                                      nz, nz_m_tiles, nz_kb_total);
    return cudaGetLastError();
}
cudaError_t qf_launch_add_cols_i32(int32_t* e, long lde, const double* sol, long ldsol, const int* cols, int ncols,
                                   int B, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    add_cols_i32_kernel<<<grid_for((long long)B * ncols, TPB), TPB, 0, stream>>>(e, lde, sol, ldsol, cols, ncols, B);
    return cudaGetLastError();
}

// ---- narrow boundary types: int16 Domain values (|entry| <= 6 s r < 2^15 at C2 / C3) and the device-side range check of
// the targets ------------------------------------------------------------------------------------------------------------
namespace {
// 8 values per thread: two 128-bit loads, one 128-bit store; *flag |= 64 when a value does not fit int16
__global__ void __launch_bounds__(256) narrow_i32_i16_kernel(const int32_t* __restrict__ in, int16_t* __restrict__ out, size_t count,
                                                             int* flag) {
    const size_t groups = count >> 3;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        const int4 a = __ldcs(reinterpret_cast<const int4*>(in) + 2 * g), b = __ldcs(reinterpret_cast<const int4*>(in) + 2 * g + 1);
        const int v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        unsigned pk[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            bad |= (v[2 * t] != (int)(short)v[2 * t]) | (v[2 * t + 1] != (int)(short)v[2 * t + 1]);
            pk[t] = ((unsigned)v[2 * t] & 0xFFFFu) | ((unsigned)v[2 * t + 1] << 16);
        }
        __stcs(reinterpret_cast<uint4*>(out) + g, make_uint4(pk[0], pk[1], pk[2], pk[3]));
    }
    if (blockIdx.x == 0 && threadIdx.x < (count & 7)) {
        const size_t i = (groups << 3) + threadIdx.x;
        const int v = in[i];
        bad |= v != (int)(short)v;
        out[i] = (int16_t)v;
    }
    if (bad && flag) atomicOr(flag, 64);
}
__global__ void __launch_bounds__(256) widen_i16_i32_kernel(const int16_t* __restrict__ in, int32_t* __restrict__ out, size_t count) {
    const size_t groups = count >> 3;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4*>(in) + g);
        const unsigned w[4] = {a.x, a.y, a.z, a.w};
        int v[8];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            v[2 * t] = (int)(short)(w[t] & 0xFFFFu);
            v[2 * t + 1] = (int)(short)(w[t] >> 16);
        }
        __stcs(reinterpret_cast<int4*>(out) + 2 * g, make_int4(v[0], v[1], v[2], v[3]));
        __stcs(reinterpret_cast<int4*>(out) + 2 * g + 1, make_int4(v[4], v[5], v[6], v[7]));
    }
    if (blockIdx.x == 0 && threadIdx.x < (count & 7)) {
        const size_t i = (groups << 3) + threadIdx.x;
        out[i] = (int32_t)in[i];
    }
}
// *flag |= 32 when some residue lies outside [0, q)
__global__ void __launch_bounds__(256) range_check_i64_kernel(const int64_t* __restrict__ v, size_t count, unsigned long long q, int* flag) {
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        bad |= (unsigned long long)v[i] >= q;  // negative values are huge as unsigned
    if (bad) atomicOr(flag, 32);
}
}  // namespace
cudaError_t qf_launch_narrow_i32_i16(const int32_t* in, int16_t* out, size_t count, int* flag, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    if ((((uintptr_t)in) & 15) || (((uintptr_t)out) & 15)) return cudaErrorMisalignedAddress;
    narrow_i32_i16_kernel<<<grid_for((long long)((count >> 3) + 1), 256, 148 * 8), 256, 0, stream>>>(in, out, count, flag);
    return cudaGetLastError();
}
cudaError_t qf_launch_widen_i16_i32(const int16_t* in, int32_t* out, size_t count, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    if ((((uintptr_t)in) & 15) || (((uintptr_t)out) & 15)) return cudaErrorMisalignedAddress;
    widen_i16_i32_kernel<<<grid_for((long long)((count >> 3) + 1), 256, 148 * 8), 256, 0, stream>>>(in, out, count);
    return cudaGetLastError();
}
cudaError_t qf_launch_range_check_i64(const int64_t* v, size_t count, unsigned long long q, int* flag, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    range_check_i64_kernel<<<grid_for((long long)count, 256, 148 * 4), 256, 0, stream>>>(v, count, q, flag);
    return cudaGetLastError();
}
