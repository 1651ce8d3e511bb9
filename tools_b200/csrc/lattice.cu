// Sequential (per-target) parts of the two lattice samplers.
//
//  * gadget_sample : randomized nearest-plane on the block-diagonal gadget basis
//    I_n (x) S_k  (mp_perturbation.rs:173-191; digits gadget_classical.rs:169-182;
//    basis gadget_classical.rs:248-287).  The reference runs a dense nk x nk
//    SampleD; because S and its GSO are block diagonal every inner product with
//    another block is exactly zero, so each (target,row) pair is an independent
//    k-step recursion.  One thread per (target,row).
//  * np_diag : one nb-wide diagonal block of GPV08 SampleD
//    (MatZ::sample_d_precomputed_gso as used at gpv.rs:160) expressed in GSO
//    coordinates; the off-diagonal part is done by gemm_f64 between calls.
//    One thread per target, the nb x nb block of mu-coefficients broadcast from
//    shared memory, target rows staged through shared memory so that global
//    traffic is coalesced.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int GADGET_TPB = 128;
constexpr int KMAX = 64;

__global__ void __launch_bounds__(GADGET_TPB)
gadget_sample_kernel(const int64_t* __restrict__ V, long ldv, double* __restrict__ Z, long ldz, int B, int n, int k,
                     int base, unsigned long long q, const double* __restrict__ sk_g,
                     const double* __restrict__ gso_g, double s_g, uint64_t seed, uint64_t first_target) {
    extern __shared__ __align__(16) double sm[];
    double* sk = sm;             // k*k   sk[t*k+i] = (b_i)_t
    double* gs = sm + k * k;     // k*k   gs[t*k+i] = (b~_i)_t
    double* inv_n2 = gs + k * k; // k
    DGaussParams* dg = reinterpret_cast<DGaussParams*>(inv_n2 + k);
    for (int i = threadIdx.x; i < k * k; i += blockDim.x) {
        sk[i] = sk_g[i];
        gs[i] = gso_g[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        double n2 = 0;
        for (int t = 0; t < k; ++t) n2 += gs[t * k + i] * gs[t * k + i];
        inv_n2[i] = 1.0 / n2;
        dg[i] = make_dgauss(s_g / sqrt(n2));
    }
    __syncthreads();

    long total = (long)B * n;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long b = idx / n;
        int row = (int)(idx - b * n);
        unsigned long long v = (unsigned long long)V[b * ldv + row] % q;
        double c[KMAX];  // running centre  c = -(x0 + sum z_i b_i)
        for (int t = 0; t < k; ++t) {
            unsigned long long d = v % (unsigned long long)base;
            c[t] = -(double)d;
            v = (v - d) / (unsigned long long)base;
        }
        Philox rng;
        rng.init(seed, (first_target + (uint64_t)b) * (uint64_t)n + (uint64_t)row, QF_STREAM_GADGET);
        for (int i = k - 1; i >= 0; --i) {
            double dot = 0;
            for (int t = 0; t < k; ++t) dot += c[t] * gs[t * k + i];
            double z = sample_dgauss(dg[i], dot * inv_n2[i], rng);
            if (z != 0.0)
                for (int t = 0; t < k; ++t) c[t] -= z * sk[t * k + i];
        }
        double* zr = Z + b * ldz + (long)row * k;
        for (int t = 0; t < k; ++t) zr[t] = -c[t];
    }
}

constexpr int NP_TPB = 64;
constexpr int NP_NB_MAX = 64;

__global__ void __launch_bounds__(NP_TPB)
np_diag_kernel(const double* __restrict__ T, long ldt, double* __restrict__ Z, long ldz, const double* __restrict__ U,
               long ldu, const DGaussParams* __restrict__ dg_g, int B, int j0, int nb, int dim, uint64_t seed,
               uint64_t first_target, double zlimit, int* flag) {
    extern __shared__ __align__(16) double np_sm[];
    double* us = np_sm;                         // nb x nb mu-coefficients of the diagonal block
    double* ts = np_sm + nb * nb;               // nb x (NP_TPB+1) centres, then samples
    DGaussParams* dgs = reinterpret_cast<DGaussParams*>(ts + nb * (NP_TPB + 1));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbe = min(nb, dim - j0);  // valid coordinates in this block
    for (int i = tid; i < nb * nb; i += NP_TPB) {
        int r = i / nb, c = i - r * nb;
        us[i] = (r < nbe && c < nbe && c > r) ? U[(long)(j0 + r) * ldu + (j0 + c)] : 0.0;
    }
    for (int i = tid; i < nbe; i += NP_TPB) dgs[i] = dg_g[j0 + i];
    const long b0 = (long)blockIdx.x * NP_TPB;
    // stage T[b0 .. b0+TPB)[j0 .. j0+nb): each warp copies rows, lanes on coordinates (coalesced)
    for (int r = warp; r < NP_TPB; r += NP_TPB / 32) {
        long b = b0 + r;
        if (b < B)
            for (int c = lane; c < nbe; c += 32) ts[c * (NP_TPB + 1) + r] = T[b * ldt + j0 + c];
    }
    __syncthreads();
    const long b = b0 + tid;
    if (b < B) {
        for (int ii = nbe - 1; ii >= 0; --ii) {
            double cp = ts[ii * (NP_TPB + 1) + tid];
            for (int jj = ii + 1; jj < nbe; ++jj) cp -= us[ii * nb + jj] * ts[jj * (NP_TPB + 1) + tid];
            Philox rng;
            rng.init(seed, (first_target + (uint64_t)b) * (uint64_t)dim + (uint64_t)(j0 + ii), QF_STREAM_NP);
            double z = sample_dgauss(dgs[ii], cp, rng);
            if (!(fabs(z) < zlimit) && flag) atomicOr(flag, 2);
            ts[ii * (NP_TPB + 1) + tid] = z;
        }
    }
    __syncthreads();
    for (int r = warp; r < NP_TPB; r += NP_TPB / 32) {
        long bb = b0 + r;
        if (bb < B)
            for (int c = lane; c < nbe; c += 32) Z[bb * ldz + j0 + c] = ts[c * (NP_TPB + 1) + r];
    }
}

}  // namespace

cudaError_t qf_launch_gadget_sample(const int64_t* V, long ldv, double* Z, long ldz, int B, int n, int k, int base,
                                    unsigned long long q, const double* sk, const double* gso, double s_g,
                                    uint64_t seed, uint64_t first_target, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (k > KMAX) return cudaErrorInvalidValue;
    size_t smem = (size_t)(2 * k * k + k) * sizeof(double) + (size_t)k * sizeof(DGaussParams);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(gadget_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    long long total = (long long)B * n;
    long long g = (total + GADGET_TPB - 1) / GADGET_TPB;
    if (g > 148 * 16) g = 148 * 16;
    gadget_sample_kernel<<<(int)g, GADGET_TPB, smem, stream>>>(V, ldv, Z, ldz, B, n, k, base, q, sk, gso, s_g, seed,
                                                               first_target);
    return cudaGetLastError();
}

cudaError_t qf_launch_np_diag(const double* T, long ldt, double* Z, long ldz, const double* U, long ldu,
                              const DGaussParams* dg, int B, int j0, int nb, int dim, uint64_t seed,
                              uint64_t first_target, double zlimit, int* flag, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (nb > NP_NB_MAX || nb < 1) return cudaErrorInvalidValue;
    int grid = (B + NP_TPB - 1) / NP_TPB;
    size_t smem = (size_t)(nb * nb + nb * (NP_TPB + 1)) * sizeof(double) + (size_t)nb * sizeof(DGaussParams);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(np_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    np_diag_kernel<<<grid, NP_TPB, smem, stream>>>(T, ldt, Z, ldz, U, ldu, dg, B, j0, nb, dim, seed, first_target,
                                                   zlimit, flag);
    return cudaGetLastError();
}
