// Batched "targets x key-matrix" contraction in fp64 on the DMMA tensor path.
//
//   C[b][n] = alpha * sum_{k < Klim(n)} X[b][k] * W[n][k]  +  beta * C[b][n]
//
// X : B x K   one row per target (K contiguous)      -- sigma, g, z, -sol ...
// W : N x K   one row per output coordinate          -- A, sqrt(Sigma_2), U, S, R ...
// C : B x N   one row per target
//
// Both operands are K-major, i.e. exactly the fragment layout of
// mma.sync.m8n8k4.row.col.f64.  Used (a) for the floating contractions of the
// samplers (sqrt(Sigma_2) * g, mp_perturbation.rs:315; the nearest-plane
// coefficient updates of sample_d_precomputed_gso, gpv.rs:160) and (b) as an
// EXACT integer contraction whenever |X| * |W| * K < 2^53 (A * sigma, R * z,
// S * z with the operand split into chunks by the host).
//
// Tile 128 (targets) x 128 (coords) x 16, 8 warps (2 x 4), each warp 64 x 32 =
// 8 x 4 DMMA tiles, 3-stage cp.async pipeline, smem rows padded to 20 doubles so
// that the 8x4 fragment loads are bank-conflict free.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 3, LDS = 20;
constexpr int GEMM_THREADS = 256;
constexpr int STAGE_DOUBLES = (BM + BN) * LDS;
constexpr int GEMM_SMEM = STAGES * STAGE_DOUBLES * (int)sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_f64_kernel(const double* __restrict__ X, long ldx, const double* __restrict__ W, long ldw,
                double* __restrict__ C, long ldc, int B, int N, int K, double alpha, double beta,
                int tri_lower) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int n0 = blockIdx.x * BN, b0 = blockIdx.y * BM;
    int klim = K;
    if (tri_lower) klim = min(K, n0 + BN);
    const int KT = (klim + BK - 1) / BK;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int stage, int kt) {
        double* As = smem + stage * STAGE_DOUBLES;
        double* Ws = As + BM * LDS;
        const int k0 = kt * BK;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            int id = tid + it * GEMM_THREADS;
            int row = id >> 3, ch = id & 7;
            int k = k0 + ch * 2;
            int kb = max(0, min(16, (klim - k) * 8));
            {
                int b = b0 + row;
                int bytes = (b < B) ? kb : 0;
                const double* src = X + (long)(b < B ? b : 0) * ldx + (bytes ? k : 0);
                cp_async16(As + row * LDS + ch * 2, src, bytes);
            }
            {
                int n = n0 + row;
                int bytes = (n < N) ? kb : 0;
                const double* src = W + (long)(n < N ? n : 0) * ldw + (bytes ? k : 0);
                cp_async16(Ws + row * LDS + ch * 2, src, bytes);
            }
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }

    const int fr = lane >> 2, fc = lane & 3;
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const double* As = smem + (kt % STAGES) * STAGE_DOUBLES;
        const double* Ws = As + BM * LDS;
        const double* ap = As + (wm * 64 + fr) * LDS + fc;
        const double* bp = Ws + (wn * 32 + fr) * LDS + fc;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ap[i * 8 * LDS + kk * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = bp[j * 8 * LDS + kk * 4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    const bool vec_ok = ((ldc & 1) == 0) && ((((uintptr_t)C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int b = b0 + wm * 64 + i * 8 + fr;
        if (b >= B) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + wn * 32 + j * 8 + fc * 2;
            if (n >= N) continue;
            double* cp = C + (long)b * ldc + n;
            double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
            if (n + 1 < N && vec_ok) {
                if (beta != 0.0) {
                    double2 old = *reinterpret_cast<double2*>(cp);
                    v0 += beta * old.x;
                    v1 += beta * old.y;
                }
                *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
            } else {
                if (beta != 0.0) v0 += beta * cp[0];
                cp[0] = v0;
                if (n + 1 < N) {
                    if (beta != 0.0) v1 += beta * cp[1];
                    cp[1] = v1;
                }
            }
        }
    }
}

}  // namespace

cudaError_t qf_launch_gemm_f64(const double* X, long ldx, const double* W, long ldw, double* C, long ldc,
                               int B, int N, int K, double alpha, double beta, int tri_lower,
                               cudaStream_t stream) {
    if (B <= 0 || N <= 0) return cudaSuccess;
    // cp.async needs 16-byte aligned rows: even leading dimensions, aligned bases
    if ((ldx & 1) || (ldw & 1) || (((uintptr_t)X) & 15) || (((uintptr_t)W) & 15)) return cudaErrorMisalignedAddress;
    static bool configured_dev[QF_MAX_DEVICES] = {};
    bool& configured = configured_dev[qf_device_slot()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((N + BN - 1) / BN, (B + BM - 1) / BM);
    gemm_f64_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(X, ldx, W, ldw, C, ldc, B, N, K, alpha, beta, tri_lower);
    return cudaGetLastError();
}
