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
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int GADGET_TPB = 128;
constexpr int KMAX = 64;

__global__ void __launch_bounds__(GADGET_TPB)
gadget_sample_kernel(const int64_t* __restrict__ V, long ldv, double* __restrict__ Z, long ldz, int B, int n, int k,
                     int base, unsigned long long q, const double* __restrict__ sk_g,
                     const double* __restrict__ gso_g, double s_g, uint64_t seed, uint64_t first_target, int* flag) {
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
            double z = sample_dgauss(dg[i], dot * inv_n2[i], rng, flag);
            if (z != 0.0)
                for (int t = 0; t < k; ++t) c[t] -= z * sk[t * k + i];
        }
        double* zr = Z + b * ldz + (long)row * k;
        for (int t = 0; t < k; ++t) zr[t] = -c[t];
    }
}

// rare path of np_diag (both pre-generated proposals rejected): continue the coordinate's Philox stream
__device__ __noinline__ double np_sample_slow(DGaussParams dgp, double cp, uint64_t seed, uint64_t index, int* flag) {
    Philox ph;
    ph.init(seed, index, QF_STREAM_NP);
    ph.c3 = 1;
    return sample_dgauss(dgp, cp, ph, flag);
}

// First Philox block of every (target, coordinate) stream of the nearest-plane recursion turned into two
// proposals (two normals, logs of two uniforms).  The transcendental work does not depend on the centres, so it
// runs here at full occupancy instead of inside the sequential per-target chain of np_diag.
// Layout: coordinate-major, out[(i - j_lo) * ldo + b] (ldo = target stride): np_diag reads, per step, the proposals of
// consecutive targets for ONE coordinate -- contiguous.
__global__ void __launch_bounds__(256)
np_propose_kernel(float4* __restrict__ out, long ldo, int B, int j_lo, int width, int dim, uint64_t seed,
                  uint64_t first_target) {
    const long total = (long)B * width;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / B);
        const long b = idx - (long)i * B;
        Philox ph;
        ph.init(seed, (first_target + (uint64_t)b) * (uint64_t)dim + (uint64_t)(j_lo + i), QF_STREAM_NP);
        float n0, n1;
        ph.normal2(n0, n1);
        const float u0 = ph.uniform24(), u1 = ph.uniform24();
        out[(long)i * ldo + b] = make_float4(n0, n1, __logf(u0), __logf(u1));
    }
}

constexpr int NP_TPL = 2;                 // targets per lane
constexpr int NP_TARGETS = 32 * NP_TPL;   // targets per CTA: 4 warps x 8 quads x NP_TPL targets
constexpr int NP_TPB = 128;               // 4 lanes (a quad) per group of NP_TPL targets
constexpr int NP_NB_MAX = 64;
constexpr int NP_TS = NP_NB_MAX + 1;      // centre tile is target-major: ts[t * NP_TS + i]
constexpr int NP_PANEL_LD = NP_NB_MAX + 2;  // row stride of the update panel: even, so that column pairs are 16-byte aligned
constexpr int NP_UST_DOUBLES = NP_NB_MAX * NP_PANEL_LD;  // mu block (64 x 65), later the update panel (64 x 66)

// One nb-wide diagonal block.  Three phases per CTA of 64 targets (two CTAs per SM: 296 CTAs = one chunk of 18944):
//  0. stage the mu-block (transposed) and the 128 x nb tile of centres through shared memory (coalesced);
//  1. the proposals of every (target, coordinate) -- the first Philox block of that coordinate's stream turned into
//     two normals and the logs of two uniforms by np_propose_kernel at full occupancy -- are read from global memory
//     one coordinate GROUP (four steps) ahead (coordinate-major layout: a warp reads contiguous segments);
//  2. the recursion i = nb-1 .. 0.  A quad of lanes owns NP_TPL targets; lane qd of the quad keeps the centres of
//     coordinates 4k+qd of all its targets in registers.  Per step: one quad shuffle per target, the accept / reject
//     arithmetic (a handful of flops per target), then the right-looking update c'_j -= mu_ji z_i of the remaining
//     centres: every mu value read from shared memory is used for all targets of the lane (the kernel is bound
//     by exactly those shared-memory reads: one target per lane needs NP_TPL x the bandwidth for the same flops).
// Draw order and Philox counters are those of sample_dgauss(), so the output is identical to the
// one-thread-per-target formulation.
__global__ void __launch_bounds__(NP_TPB, 2)
np_diag_kernel(double* __restrict__ T, long ldt, double* __restrict__ Z, long ldz, const double* __restrict__ U,
               long ldu, const DGaussParams* __restrict__ dg_g, const float4* __restrict__ prop, long ldprop, int B,
               int j_lo, int j_hi, int prop0, int dim, uint64_t seed, uint64_t first_target, double zlimit, int* flag,
               int fuse_update, NpDigitOut dig) {
    extern __shared__ __align__(16) double np_sm[];
    double* ust = np_sm;                                  // ust[c * us_ld + r] = U[j0+r][j0+c], c > r
    double* ts = ust + NP_UST_DOUBLES;                    // ts[t * NP_TS + i]  (16-byte aligned)
    DGaussParams* dgs = reinterpret_cast<DGaussParams*>(ts + ((NP_TARGETS * NP_TS + 1) & ~1));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long b0 = (long)blockIdx.x * NP_TARGETS;
    const float4* const prop_base = prop;
    // The diagonal blocks of [j_lo, j_hi) from the top down: with the fused rank-64 update every block only touches this
    // CTA's own targets, so the whole 256-block is ONE launch (the CTAs never wait for each other between blocks).
    for (int sub_hi = j_hi; sub_hi > j_lo;) {
    const int j0 = max(j_lo, (sub_hi - 1) / NP_NB_MAX * NP_NB_MAX), nb = sub_hi - j0;
    const int up_lo = fuse_update ? j_lo : j0;
    const int us_ld = nb + 1;                             // padded: the transposing store is (almost) conflict-free
    const int nbe = min(nb, dim - j0);
    prop = prop_base + (long)(j0 - prop0) * ldprop;       // row of coordinate j0 in the coordinate-major proposal buffer
#pragma unroll 8
    for (int i = tid; i < nb * nb; i += NP_TPB) {
        const int r = i / nb, c = i - r * nb;  // coalesced read along c
        ust[c * us_ld + r] = (r < nbe && c < nbe && c > r) ? U[(long)(j0 + r) * ldu + (j0 + c)] : 0.0;
    }
    for (int i = tid; i < nbe; i += NP_TPB) dgs[i] = dg_g[j0 + i];
#pragma unroll 4
    for (int r = warp; r < NP_TARGETS; r += NP_TPB / 32) {
        const long b = b0 + r;
        const double v0 = (b < B && lane < nbe) ? T[b * ldt + j0 + lane] : 0.0;
        const double v1 = (b < B && lane + 32 < nbe) ? T[b * ldt + j0 + lane + 32] : 0.0;
        ts[r * NP_TS + lane] = v0;
        ts[r * NP_TS + lane + 32] = v1;
    }
    __syncthreads();
    // phase 2: the recursion, fully register resident; the loop over coordinate groups stays rolled (the register file
    // is rotated by one group after each group, so all register indices are static)
    {
        const int tl = (tid >> 2) * NP_TPL, qd = tid & 3;   // first local target of this lane's quad, lane within the quad
        const long bq = b0 + tl;
        constexpr int NG = NP_NB_MAX / 4;  // 16 coordinate groups; group j = coordinates 4j .. 4j+3
        double c[NP_TPL][NG];
#pragma unroll
        for (int k = 0; k < NP_TPL; ++k)
#pragma unroll
            for (int j = 0; j < NG; ++j) c[k][j] = (4 * j + qd < nbe) ? ts[(tl + k) * NP_TS + 4 * j + qd] : 0.0;
        // proposals (prop points at coordinate j0): row ii holds the proposals of all targets for coordinate j0 + ii
        const float4* pq = prop + bq;
        float4 pcur[4][NP_TPL], pnxt[4][NP_TPL];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int ii = 4 * (NG - 1) + o;
#pragma unroll
            for (int k = 0; k < NP_TPL; ++k) pcur[o][k] = (ii < nbe) ? pq[(long)ii * ldprop + k] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int kidx = 0; kidx < NG; ++kidx) {
            const int g = NG - 1 - kidx;
#pragma unroll
            for (int o = 0; o < 4; ++o) {  // proposals of the next group, in flight during the four steps of this one
                const int in = 4 * (g - 1) + o;
#pragma unroll
                for (int k = 0; k < NP_TPL; ++k)
                    pnxt[o][k] = (g > 0 && in < nbe) ? pq[(long)in * ldprop + k] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int owner = 3; owner >= 0; --owner) {
                const int ii = 4 * g + owner;
                if (ii < nbe) {  // uniform
                    const DGaussParams dgp = dgs[ii];
                    double z[NP_TPL];
#pragma unroll
                    for (int k = 0; k < NP_TPL; ++k) {
                        const double cp = __shfl_sync(0xffffffffu, c[k][NG - 1], (lane & ~3) | owner);
                        const float4 pr = pcur[owner][k];
                        const double c_int = rint(cp);
                        const float c_frac = (float)(cp - c_int);
                        double zz = 0.0;
                        bool done = false;
                        {
                            float x = rintf(fmaf(dgp.sigma_p, pr.x, c_frac));
                            float d = x - c_frac;
                            float e = fmaf(-d * d, dgp.inv2s2, fmaf(0.5f * pr.x, pr.x, -dgp.emax));
                            if (fabsf(d) <= dgp.tail && pr.z < e) { zz = c_int + (double)x; done = true; }
                        }
                        if (!done) {
                            float x = rintf(fmaf(dgp.sigma_p, pr.y, c_frac));
                            float d = x - c_frac;
                            float e = fmaf(-d * d, dgp.inv2s2, fmaf(0.5f * pr.y, pr.y, -dgp.emax));
                            if (fabsf(d) <= dgp.tail && pr.w < e) { zz = c_int + (double)x; done = true; }
                        }
                        const bool live = bq + k < B;
                        if (!done && live)  // both pre-generated proposals rejected: continue the stream from its second block
                            zz = np_sample_slow(dgp, cp, seed, (first_target + (uint64_t)(bq + k)) * (uint64_t)dim + (uint64_t)(j0 + ii), flag);
                        if (live && qd == 0 && !(fabs(zz) < zlimit) && flag) atomicOr(flag, 2);
                        z[k] = zz;
                    }
                    const double* ucol = ust + ii * us_ld + qd;  // ucol[4 j] = U[j0 + 4 j + qd][j0 + ii]
                    {
                        const double u = ucol[4 * g];
#pragma unroll
                        for (int k = 0; k < NP_TPL; ++k) {
                            if (qd == owner) c[k][NG - 1] = z[k];
                            else if (qd < owner) c[k][NG - 1] = fma(-u, z[k], c[k][NG - 1]);
                        }
                    }
#pragma unroll
                    for (int kk = 0; kk < NG - 1; ++kk) {
                        const int grp = kk - kidx;  // group held by c[.][kk]
                        if (grp >= 0) {
                            const double u = ucol[4 * grp];
#pragma unroll
                            for (int k = 0; k < NP_TPL; ++k) c[k][kk] = fma(-u, z[k], c[k][kk]);
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int k = 0; k < NP_TPL; ++k) pcur[o][k] = pnxt[o][k];
#pragma unroll
            for (int k = 0; k < NP_TPL; ++k) {
                if (4 * g + qd < nbe) ts[(tl + k) * NP_TS + 4 * g + qd] = c[k][NG - 1];
#pragma unroll
                for (int kk = NG - 1; kk > 0; --kk) c[k][kk] = c[k][kk - 1];
            }
        }
    }
    __syncthreads();
    for (int r = warp; r < NP_TARGETS; r += NP_TPB / 32) {
        const long bb = b0 + r;
        if (bb < B)
            for (int c = lane; c < nbe; c += 32) Z[bb * ldz + j0 + c] = ts[r * NP_TS + c];
    }
    // phase 2b (optional): balanced base-256 digit planes of this block of z for the tensor-core updates that consume it
    // (same planes, zero-tile map and digit-count gates as split_f64_limbs_kernel, without a second pass over Z):
    // a warp per target row, two columns per lane.
    if (dig.planes != nullptr) {
        int top = -1;
        for (int r = warp; r < NP_TARGETS; r += NP_TPB / 32) {
            const long bb = b0 + r;
            long long v0 = 0, v1 = 0;
            if (bb < B) {
                if (2 * lane < nbe) v0 = __double2ll_rn(ts[r * NP_TS + 2 * lane]);
                if (2 * lane + 1 < nbe) v1 = __double2ll_rn(ts[r * NP_TS + 2 * lane + 1]);
            }
            for (int l = 0; l < dig.L; ++l) {
                long long d0 = ((v0 + 128) & 255) - 128, d1 = ((v1 + 128) & 255) - 128;
                if (l == dig.L - 1) { d0 = v0; d1 = v1; }  // |z| < zlimit <= capacity of L digits (checked in phase 2)
                v0 = (v0 - d0) >> 8;
                v1 = (v1 - d1) >> 8;
                int8_t* dst = dig.planes + (long)l * dig.plane_stride + bb * dig.ldk + j0 + 2 * lane;
                if (bb < B) {
                    if (2 * lane + 1 < nbe) *reinterpret_cast<uint16_t*>(dst) = (uint16_t)((d0 & 255) | ((d1 & 255) << 8));
                    else if (2 * lane < nbe) dst[0] = (int8_t)d0;
                }
                if (__any_sync(0xffffffffu, (d0 | d1) != 0)) top = max(top, l);
            }
        }
        // zero-tile map: this CTA's 64 targets lie in one 128-target tile, its columns in one 128-column block
        if (lane == 0 && top >= 0) {
            for (int l = 0; l <= top; ++l) {
                // (planes below the top one may still be all zero in this warp's rows; marking them non-zero only costs
                // an MMA that multiplies zeros -- but the exact map is cheap: recompute per plane is not worth it)
                dig.nz[((long)l * dig.nz_m_tiles + (b0 >> 7)) * dig.nz_kb_total + (j0 >> 7)] = 1;
            }
            if (top > 0) {
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    if (dig.gate[g] && *(volatile int*)dig.gate[g] < top) atomicMax(dig.gate[g], top);
            }
        }
    }
    // phase 3: the rank-nb update of the columns [up_lo, j0) that remain in the enclosing 256-block, for this CTA's own
    // targets:  T[b][j] -= sum_i z_i U[j][j0 + i].  The z tile is still in shared memory; the panel of U goes through
    // the (now free) mu buffer 64 columns at a time.  Lanes are targets (lane, lane + 32), warps are 16-column groups:
    // the z reads are conflict-free (row stride 65), the U reads are warp-wide broadcasts, 32 accumulators per lane.
    // (This used to be a separate K = 64 fp64 GEMM launch per diagonal block: launch- and latency-bound, ~10 ms per chunk.)
    for (int c0 = up_lo; c0 < j0; c0 += 64) {
        const int ncol = min(64, j0 - c0);
        __syncthreads();
        for (int idx = tid; idx < 64 * 64; idx += NP_TPB) {
            const int j = idx >> 6, i = idx & 63;  // coalesced along i
            ust[i * NP_PANEL_LD + j] = (j < ncol && i < nbe) ? U[(long)(c0 + j) * ldu + j0 + i] : 0.0;
        }
        __syncthreads();
        double acc0[16], acc1[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc0[c] = acc1[c] = 0.0;
        const double* z0 = ts + lane * NP_TS;
        const double* z1 = ts + (lane + 32) * NP_TS;
        const double2* ub = reinterpret_cast<const double2*>(ust + warp * 16);  // 16-byte aligned: even stride, even offset
#pragma unroll 2
        for (int i = 0; i < nbe; ++i) {
            const double a = z0[i], bq2 = z1[i];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const double2 u = ub[i * (NP_PANEL_LD / 2) + c];  // warp-wide broadcast, two columns per load
                acc0[2 * c] = fma(a, u.x, acc0[2 * c]);
                acc0[2 * c + 1] = fma(a, u.y, acc0[2 * c + 1]);
                acc1[2 * c] = fma(bq2, u.x, acc1[2 * c]);
                acc1[2 * c + 1] = fma(bq2, u.y, acc1[2 * c + 1]);
            }
        }
        if (warp * 16 < ncol) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const long bb = b0 + lane + 32 * k;
                if (bb < B) {
                    double2* tr = reinterpret_cast<double2*>(T + bb * ldt + c0 + warp * 16);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        double2 v = tr[c];
                        v.x -= k ? acc1[2 * c] : acc0[2 * c];
                        v.y -= k ? acc1[2 * c + 1] : acc0[2 * c + 1];
                        tr[c] = v;
                    }
                }
            }
        }
    }
    __syncthreads();  // the next block re-uses the mu buffer and the centre tile, and reads T columns updated above
    sub_hi = j0;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// np_diag2: the same diagonal-block walk with ONE TARGET PER THREAD.
// A thread owns one target; the 64 centres of the block live in its column of a shared-memory tile
// (cs[i * 128 + tid]: conflict-free).  The block is walked in groups of 8 coordinates: the 8 centres of the group
// are in registers for the 8 sequential steps (accept / reject arithmetic done once per target -- the quad formulation
// above repeats it in four lanes and serves 16 targets per warp-instruction where this one serves 32), then the rank-8
// update of the lower coordinates streams their centres through registers once (8 FMAs per load / store pair, the
// mu values as warp-wide 16-byte broadcasts from the row-major block in shared memory).  The loop over groups stays
// rolled: ~600 instructions of loop body (a fully unrolled 64-step recursion is 200 KB of code and runs out of the
// instruction cache at 4 warps per SM -- measured slower than the quad kernel).  Draw order, Philox counters and the
// order of every floating-point operation per (target, coordinate) are those of np_diag_kernel: the two kernels
// produce bit-identical output (tested).  One CTA = 128 targets = one 128-target tile of the zero-tile map.
// ---------------------------------------------------------------------------------------------------------------
constexpr int ND2_TPB = 128;
// Row strides (doubles) of the mu block / update panel and of the centre tile in shared memory: even (16-byte rows), and
// 2 * stride = 16 mod 32 banks, so that the 8 x 4 DMMA fragment loads of the rank-64 update (rows fr = lane / 4 of
// consecutive columns, k = lane % 4 of consecutive rows) take the minimal two wavefronts.
constexpr int ND2_LD = 72;
constexpr int ND2_ZLD = 136;
constexpr int ND2_PLD = 68;  // update panel, stored [column][i]: coalesced conflict-free staging stores, two-wavefront B fragments

__device__ __forceinline__ void nd2_dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(ND2_TPB, 2)
np_diag2_kernel(double* __restrict__ T, long ldt, double* __restrict__ Z, long ldz, const double* __restrict__ U,
                long ldu, const DGaussParams* __restrict__ dg_g, const float4* __restrict__ prop, long ldprop, int B,
                int j_lo, int j_hi, int prop0, int dim, uint64_t seed, uint64_t first_target, double zlimit, int* flag,
                int fuse_update, NpDigitOut dig) {
    extern __shared__ __align__(16) double nd2_sm[];
    double* ur = nd2_sm;             // ur[j * ND2_LD + i] = U[j0 + j][j0 + i] (i > j), later the panel pan[i * ND2_LD + c]
    double* cs = ur + 64 * ND2_LD;   // cs[i * ND2_ZLD + tid]: centres, overwritten by z as the recursion proceeds
    DGaussParams* dgs = reinterpret_cast<DGaussParams*>(cs + 64 * ND2_ZLD);
    const int tid = threadIdx.x, lane = tid & 31;
    const long b0 = (long)blockIdx.x * ND2_TPB, b = b0 + tid;
    const bool live = b < B;
    double* const cst = cs + tid;
    bool tile_on_chip = false;  // the centre tile of this block was left in shared memory by the previous block's update
    for (int sub_hi = j_hi; sub_hi > j_lo;) {
        const int j0 = max(j_lo, (sub_hi - 1) / 64 * 64), nb = sub_hi - j0;
        const int up_lo = fuse_update ? j_lo : j0;
        const int nbe = min(nb, dim - j0);
        {
            // the 64 x 64 block of mu: 32 independent loads per thread in flight, then the stores
            const int mr = tid >> 6, mc = tid & 63;  // rows mr + 2 r, column mc (coalesced along the column index)
            double mv[32];
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const int row = mr + 2 * r;
                mv[r] = (row < nbe && mc < nbe && mc > row) ? U[(long)(j0 + row) * ldu + (j0 + mc)] : 0.0;
            }
            __syncthreads();  // the previous block's panel reads are done
#pragma unroll
            for (int r = 0; r < 32; ++r) ur[(mr + 2 * r) * ND2_LD + mc] = mv[r];
        }
        for (int i = tid; i < 64; i += ND2_TPB) dgs[i] = i < nbe ? dg_g[j0 + i] : DGaussParams{0.f, 0.f, 0.f, 0.f};
        // this target's centres -> its column of the tile
        const bool have_tile = tile_on_chip;
        tile_on_chip = false;
        if (!have_tile) {
            const double* tr = T + b * ldt + j0;
            if (live && nbe == 64 && ((((uintptr_t)tr) & 15) == 0)) {
#pragma unroll 8
                for (int i = 0; i < 32; ++i) {
                    const double2 v = reinterpret_cast<const double2*>(tr)[i];
                    cst[(2 * i) * ND2_ZLD] = v.x;
                    cst[(2 * i + 1) * ND2_ZLD] = v.y;
                }
            } else {
                for (int i = 0; i < 64; ++i) cst[i * ND2_ZLD] = (live && i < nbe) ? tr[i] : 0.0;
            }
        }
        const float4* pq = prop + (long)(j0 - prop0) * ldprop + b;
        const uint64_t index0 = (first_target + (uint64_t)b) * (uint64_t)dim + (uint64_t)j0;
        float4 pcur[8], pnxt[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) pcur[t] = (live && 56 + t < nbe) ? pq[(long)(56 + t) * ldprop] : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
#pragma unroll 1
        for (int g = 7; g >= 0; --g) {
            const int ig = 8 * g;
#pragma unroll
            for (int t = 0; t < 8; ++t)  // proposals of the next group, in flight during the eight steps of this one
                pnxt[t] = (live && g > 0 && ig - 8 + t < nbe) ? pq[(long)(ig - 8 + t) * ldprop] : make_float4(0.f, 0.f, 0.f, 0.f);
            if (ig < nbe) {  // uniform
                double c8[8], z8[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) c8[t] = cst[(ig + t) * ND2_ZLD];
#pragma unroll
                for (int t = 7; t >= 0; --t) {
                    double zz = 0.0;
                    if (ig + t < nbe) {  // uniform
                        const DGaussParams dgp = dgs[ig + t];
                        const float4 p4 = pcur[t];
                        const double cp = c8[t];
                        const double c_int = rint(cp);
                        const float c_frac = (float)(cp - c_int);
                        bool done = false;
                        {
                            float x = rintf(fmaf(dgp.sigma_p, p4.x, c_frac));
                            float d = x - c_frac;
                            float e = fmaf(-d * d, dgp.inv2s2, fmaf(0.5f * p4.x, p4.x, -dgp.emax));
                            if (fabsf(d) <= dgp.tail && p4.z < e) { zz = c_int + (double)x; done = true; }
                        }
                        if (!done) {
                            float x = rintf(fmaf(dgp.sigma_p, p4.y, c_frac));
                            float d = x - c_frac;
                            float e = fmaf(-d * d, dgp.inv2s2, fmaf(0.5f * p4.y, p4.y, -dgp.emax));
                            if (fabsf(d) <= dgp.tail && p4.w < e) { zz = c_int + (double)x; done = true; }
                        }
                        if (!done && live)  // both pre-generated proposals rejected: continue the stream from its second block
                            zz = np_sample_slow(dgp, cp, seed, index0 + (uint64_t)(ig + t), flag);
                        if (live && !(fabs(zz) < zlimit) && flag) atomicOr(flag, 2);
                        const double* ucol = ur + ig * ND2_LD + ig + t;  // ucol[t' * ND2_LD] = U[ig + t'][ig + t]
#pragma unroll
                        for (int tp = 0; tp < t; ++tp) c8[tp] = fma(-ucol[tp * ND2_LD], zz, c8[tp]);
                    }
                    z8[t] = zz;
                }
#pragma unroll
                for (int t = 0; t < 8; ++t) cst[(ig + t) * ND2_ZLD] = z8[t];
                // rank-8 update of the lower coordinates, in the order of the one-step-at-a-time recursion (t descending)
#pragma unroll 8
                for (int j = 0; j < ig; ++j) {
                    double cj = cst[j * ND2_ZLD];
                    const double2* up = reinterpret_cast<const double2*>(ur + j * ND2_LD + ig);  // warp-wide broadcasts
                    const double2 u01 = up[0], u23 = up[1], u45 = up[2], u67 = up[3];
                    cj = fma(-u67.y, z8[7], cj);
                    cj = fma(-u67.x, z8[6], cj);
                    cj = fma(-u45.y, z8[5], cj);
                    cj = fma(-u45.x, z8[4], cj);
                    cj = fma(-u23.y, z8[3], cj);
                    cj = fma(-u23.x, z8[2], cj);
                    cj = fma(-u01.y, z8[1], cj);
                    cj = fma(-u01.x, z8[0], cj);
                    cst[j * ND2_ZLD] = cj;
                }
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) pcur[t] = pnxt[t];
        }
        // the tile now holds z: Z, digit planes
        double c[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) c[i] = cst[i * ND2_ZLD];
        if (live) {
            double* zr = Z + b * ldz + j0;
            if (nbe == 64 && ((((uintptr_t)zr) & 15) == 0)) {
#pragma unroll
                for (int i = 0; i < 32; ++i) reinterpret_cast<double2*>(zr)[i] = make_double2(c[2 * i], c[2 * i + 1]);
            } else {
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (i < nbe) zr[i] = c[i];
            }
        }
        // balanced base-256 digit planes, zero-tile map and digit-count gates of this block of z
        if (dig.planes != nullptr) {
            int top = -1;
            // digit l of v: v_l = floor((v + 128 (256^l - 1) / 255) / 256^l) (the balanced carries of the lower digits,
            // nested floors collapsed), d_l = ((v_l + 128) mod 256) - 128, the top plane keeps all of v_{L-1}
            long long bias = 0;
            for (int l = 0; l < dig.L; ++l) {
                int8_t* dst = dig.planes + (long)l * dig.plane_stride + b * dig.ldk + j0;
                unsigned any = 0;
                unsigned w[16];
                const bool last = l == dig.L - 1;
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const long long v = (live && i < nbe) ? __double2ll_rn(c[i]) : 0ll;
                    const long long vl = (v + bias) >> (8 * l);
                    const long long d = last ? vl : (((vl + 128) & 255) - 128);  // |z| < zlimit <= capacity of L digits
                    const unsigned byte = (unsigned)(d & 255);
                    any |= byte;
                    if ((i & 3) == 0) w[i >> 2] = byte; else w[i >> 2] |= byte << (8 * (i & 3));
                }
                if (live) {
                    if (nbe == 64) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            reinterpret_cast<uint4*>(dst)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
                    } else {
                        for (int i = 0; i < nbe; ++i) dst[i] = (int8_t)((w[i >> 2] >> (8 * (i & 3))) & 255);
                    }
                }
                if (any) top = l;
                bias += 128ll << (8 * l);
            }
            top = __reduce_max_sync(0xffffffffu, top);
            if (lane == 0 && top >= 0) {
                for (int l = 0; l <= top; ++l) dig.nz[((long)l * dig.nz_m_tiles + (b0 >> 7)) * dig.nz_kb_total + (j0 >> 7)] = 1;
                if (top > 0) {
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        if (dig.gate[g] && *(volatile int*)dig.gate[g] < top) atomicMax(dig.gate[g], top);
                }
            }
        }
        // rank-nb update of the columns [up_lo, j0) of the enclosing block for this CTA's own targets:
        //   T[b][col] -= sum_i z_i U[col][j0 + i],  the panel of U through shared memory 64 columns at a time
        // On the fp64 tensor path (mma.sync m8n8k4): a warp owns its 32 targets x 64 columns as 4 x 8 DMMA tiles, A = z from the
        // centre tile, B = the panel, both read as fragments from shared memory (12 loads per 32 MMAs; the CUDA-core form needs
        // one broadcast load per two FMAs and was bound by exactly those loads).  The next panel is fetched into registers while
        // the current one is being multiplied.
        const int pj = tid >> 6, pi = tid & 63;  // this thread's panel elements: columns pj + 2 r, row pi (coalesced along i)
        const int wrp = tid >> 5, fr = lane >> 2, fc = lane & 3;
        double pv[32];
        auto load_panel = [&](int c0, int ncol) {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const int j = pj + 2 * r;
                pv[r] = (j < ncol && pi < nbe) ? U[(long)(c0 + j) * ldu + j0 + pi] : 0.0;
            }
        };
        if (up_lo < j0) load_panel(up_lo, min(64, j0 - up_lo));
        for (int c0 = up_lo; c0 < j0; c0 += 64) {
            const int ncol = min(64, j0 - c0);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 32; ++r) ur[(pj + 2 * r) * ND2_PLD + pi] = pv[r];
            __syncthreads();
            if (c0 + 64 < j0) load_panel(c0 + 64, min(64, j0 - c0 - 64));
            // the accumulators START as the old values of T (requested here, 32 loads in flight per lane, and only needed when
            // the first MMA of their tile issues) and A = -z: T - sum_i z_i U is what the MMAs leave behind, stored as is
            double acc[4][8][2];
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                const long row = b0 + wrp * 32 + rt * 8 + fr;
                const double2* tr = reinterpret_cast<const double2*>(T + row * ldt + c0 + 2 * fc);
#pragma unroll
                for (int ct = 0; ct < 8; ++ct) {
                    const double2 v = (row < B && 8 * ct + 2 * fc < ncol) ? tr[4 * ct] : make_double2(0.0, 0.0);
                    acc[rt][ct][0] = v.x;
                    acc[rt][ct][1] = v.y;
                }
            }
            const double* ap = cs + fc * ND2_ZLD + wrp * 32 + fr;   // A[row = target][k = i] = -z_i(target)
            const double* bp = ur + fr * ND2_PLD + fc;              // B[k = i][n = column] = U[c0 + column][j0 + i]
#pragma unroll 2
            for (int kk = 0; kk < 16; ++kk) {
                double a[4], bb[8];
#pragma unroll
                for (int rt = 0; rt < 4; ++rt) a[rt] = -ap[kk * 4 * ND2_ZLD + rt * 8];
#pragma unroll
                for (int ct = 0; ct < 8; ++ct) bb[ct] = bp[ct * 8 * ND2_PLD + kk * 4];
#pragma unroll
                for (int rt = 0; rt < 4; ++rt)
#pragma unroll
                    for (int ct = 0; ct < 8; ++ct) nd2_dmma884(acc[rt][ct][0], acc[rt][ct][1], a[rt], bb[ct]);
            }
            // C fragment: rows 8 rt + fr, columns 8 ct + 2 fc, + 1
            if (c0 + 64 == j0) {
                // the 64 columns just below this block ARE the next diagonal block: their updated centres go straight into
                // the (now dead: every panel has been multiplied) centre tile instead of through HBM and back
                __syncwarp();  // a warp only reads / writes the tile columns of its own 32 targets
#pragma unroll
                for (int rt = 0; rt < 4; ++rt)
#pragma unroll
                    for (int ct = 0; ct < 8; ++ct) {
                        double* dst = cs + (8 * ct + 2 * fc) * ND2_ZLD + wrp * 32 + rt * 8 + fr;
                        dst[0] = acc[rt][ct][0];
                        dst[ND2_ZLD] = acc[rt][ct][1];
                    }
                tile_on_chip = true;
            } else {
#pragma unroll
                for (int rt = 0; rt < 4; ++rt) {
                    const long row = b0 + wrp * 32 + rt * 8 + fr;
                    if (row < B) {
                        double2* tr = reinterpret_cast<double2*>(T + row * ldt + c0 + 2 * fc);
#pragma unroll
                        for (int ct = 0; ct < 8; ++ct)
                            if (8 * ct + 2 * fc < ncol) tr[4 * ct] = make_double2(acc[rt][ct][0], acc[rt][ct][1]);
                    }
                }
            }
        }
        sub_hi = j0;
    }
}

}  // namespace

cudaError_t qf_launch_gadget_sample(const int64_t* V, long ldv, double* Z, long ldz, int B, int n, int k, int base,
                                    unsigned long long q, const double* sk, const double* gso, double s_g,
                                    uint64_t seed, uint64_t first_target, int* flag, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (k > KMAX) return cudaErrorInvalidValue;
    size_t smem = (size_t)(2 * k * k + k) * sizeof(double) + (size_t)k * sizeof(DGaussParams);
    static size_t configured_dev[QF_MAX_DEVICES] = {};
    size_t& configured = configured_dev[qf_device_slot()];
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(gadget_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    long long total = (long long)B * n;
    long long g = (total + GADGET_TPB - 1) / GADGET_TPB;
    if (g > 148 * 16) g = 148 * 16;
    gadget_sample_kernel<<<(int)g, GADGET_TPB, smem, stream>>>(V, ldv, Z, ldz, B, n, k, base, q, sk, gso, s_g, seed,
                                                               first_target, flag);
    return cudaGetLastError();
}

cudaError_t qf_launch_np_propose(float4* out, long ldo, int B, int j_lo, int width, int dim, uint64_t seed,
                                 uint64_t first_target, cudaStream_t stream) {
    if (B <= 0 || width <= 0) return cudaSuccess;
    long long total = (long long)B * width;
    long long g = (total + 255) / 256;
    if (g > 148 * 32) g = 148 * 32;
    np_propose_kernel<<<(int)g, 256, 0, stream>>>(out, ldo, B, j_lo, width, dim, seed, first_target);
    return cudaGetLastError();
}

cudaError_t qf_launch_np_diag(double* T, long ldt, double* Z, long ldz, const double* U, long ldu,
                              const DGaussParams* dg, const float4* prop, long ldprop, int B, int j0, int nb, int dim,
                              uint64_t seed, uint64_t first_target, double zlimit, int* flag, cudaStream_t stream, int up_lo,
                              const NpDigitOut* dig, int prop0, int variant) {
    if (B <= 0) return cudaSuccess;
    // nb <= 64: one diagonal block; nb > 64 (whole 256-block, up_lo == j0 required): its diagonal blocks from the top
    // down with the rank-64 updates fused
    const bool multi = nb > NP_NB_MAX;
    if (nb < 1 || (multi && (up_lo != j0 || (j0 % NP_NB_MAX) != 0))) return cudaErrorInvalidValue;
    if (prop0 < 0) prop0 = j0;
    int grid = (B + NP_TARGETS - 1) / NP_TARGETS;
    if (up_lo < 0 || up_lo > j0) up_lo = j0;  // no fused update
    // the fused update needs 16-byte aligned rows and whole 16-column groups
    if ((up_lo < j0 || multi) && ((ldt & 1) || ((j0 - up_lo) & 15) || (up_lo & 1) || (((uintptr_t)T) & 15))) return cudaErrorInvalidValue;
    // single block with an update range below it: launched as the range [up_lo .. j0 + nb) restricted to its top block
    // is not expressible -- the kernel walks [j_lo, j_hi) completely; a single block with up_lo < j0 therefore runs as
    // j_lo = j0 with the update range passed through fuse_update = 0 ... (not needed any more: the host fuses whole
    // 256-blocks or nothing)
    if (!multi && up_lo < j0) return cudaErrorInvalidValue;
    size_t smem = (size_t)(NP_UST_DOUBLES + ((NP_TARGETS * NP_TS + 1) & ~1)) * sizeof(double) +
                  (size_t)NP_NB_MAX * sizeof(DGaussParams);
    static size_t configured_dev[QF_MAX_DEVICES] = {};
    size_t& configured = configured_dev[qf_device_slot()];
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(np_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    NpDigitOut d{};
    if (dig) {
        d = *dig;
        // two-byte stores / one map cell per CTA: even column offsets, 64-aligned blocks inside one 128-column cell
        if ((j0 & 63) || (d.ldk & 1) || (((uintptr_t)d.planes) & 1) || (d.plane_stride & 1)) return cudaErrorInvalidValue;
    }
    // variant 1 (test switch QF_NP_DIAG_V1, read at context creation): the quad-per-two-targets kernel (bit-identical output)
    const bool v1 = variant == 1;
    // np_diag2 stores digits 16 bytes at a time and T two doubles at a time
    const bool v2_ok = (!dig || (((d.ldk & 15) == 0) && ((((uintptr_t)d.planes) & 15) == 0) && ((d.plane_stride & 15) == 0))) &&
                       (ldt & 1) == 0 && ((((uintptr_t)T) & 15) == 0) && ((up_lo & 1) == 0);
    if (!v1 && v2_ok) {
        const size_t smem2 = (size_t)(64 * ND2_LD + 64 * ND2_ZLD) * sizeof(double) + 64 * sizeof(DGaussParams);
        static size_t configured2_dev[QF_MAX_DEVICES] = {};
        size_t& configured2 = configured2_dev[qf_device_slot()];
        if (smem2 > configured2) {
            cudaError_t e2 = cudaFuncSetAttribute(np_diag2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            if (e2 != cudaSuccess) return e2;
            configured2 = smem2;
        }
        np_diag2_kernel<<<(B + ND2_TPB - 1) / ND2_TPB, ND2_TPB, smem2, stream>>>(T, ldt, Z, ldz, U, ldu, dg, prop, ldprop, B, j0,
                                                                                 j0 + nb, prop0, dim, seed, first_target, zlimit,
                                                                                 flag, multi ? 1 : 0, d);
        return cudaGetLastError();
    }
    np_diag_kernel<<<grid, NP_TPB, smem, stream>>>(T, ldt, Z, ldz, U, ldu, dg, prop, ldprop, B, j0, j0 + nb, prop0, dim, seed,
                                                   first_target, zlimit, flag, multi ? 1 : 0, d);
    return cudaGetLastError();
}
