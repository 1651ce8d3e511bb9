// f_a = A sigma mod q (gpv.rs:190-193, mp_perturbation.rs:366-369) with the digit split of sigma fused into the
// tensor-core contraction: sigma is read from HBM once as int32, converted to balanced base-256 digit planes by
// the CTA's own warps straight into the SWIZZLE_128B shared-memory operand tiles that tcgen05.mma consumes, and
// the squared norms of check_domain (gpv.rs:219-224) are accumulated from the same registers.  No digit planes
// ever touch HBM (the unfused path writes and re-reads them: 2x the algorithmic traffic plus a second launch).
//
// CTA = 18 warps: warps 0..15 convert (one 128-byte digit row per warp instruction: coalesced 512-byte global reads,
// conflict-free 4-byte shared stores; the 16 rows of the NEXT k block are already in flight in registers while the
// current ones are converted, i.e. 64 KB of loads outstanding per SM) and later run the epilogue (TMEM lane group =
// warp % 4), warp 16 feeds the key digits (A, u8 limbs) through TMA, warp 17 owns TMEM and issues the MMAs.  One 128-target x nt-coordinate tile
// per CTA, n tiles adjacent in the grid so that the CTAs sharing a block of sigma run together (L2 reuse).
// Digit planes of sigma that are zero in a whole 128 x 128 block are skipped by the MMA issuer.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tc05.cuh"

namespace {

using namespace tc05;
constexpr int CONV_WARPS = 16;
constexpr int FUSED_THREADS = (CONV_WARPS + 2) * 32;
constexpr int MAX_LX = 4;

struct FusedParams {
    const int32_t* x; long ldx;
    int B, N, K, LX, LW, nt, stages, w_signed, vec;
    int dmax;  // number of digit sums that can contribute to the result (see the kernel)
    unsigned long long q, qmagic;  // qmagic = floor(2^64 / q) for the Barrett reduction of the epilogue
    int64_t* out; long ldout;
    unsigned long long* norm2;
    int n_tiles;
    unsigned long long* mma_units;
    int* overflow;        // optimistic launch (CHECK): set when some value needs one digit more than LXT
    const int* run_if;    // conditional launch: exit at once unless *run_if != 0
    // optional diagnostics (QF_TRACE), cycles summed over CTAs: [0] MMA warp waiting for full stages, [1] MMA warp main
    // loop, [2] converter warp 0 waiting for empty stages, [3] converter warp 0 main loop, [4] its epilogue,
    // [5] whole CTA (warp 0, kernel entry to exit)
    unsigned long long* tim;
};

// LXT digits of sigma are multiplied.  CHECK: the digits are those of the (LXT+1)-digit representation and the kernel
// reports through p.overflow when a value has a non-zero digit LXT (the caller then re-runs with one digit more).
// PAIR: launched as clusters of two CTAs that own the two adjacent coordinate tiles of the same 128 targets.  Each CTA
// converts only HALF of the 128 sigma rows and stores the digits into its own operand tile and, through distributed
// shared memory (st.shared::cluster), into the peer's: every sigma block is read from L2 and converted once per
// cluster instead of once per CTA (the unpaired kernel is bound by exactly that: converter issue slots and bytes in
// flight).  A stage is full when the 16 local converter warps have arrived and the peer's bytes have landed
// (st.async completes the transaction count of the stage's barrier), and free again when
// BOTH tensor cores have retired their MMAs on it (tcgen05.commit multicast onto both CTAs' empty barriers).
template <int LXT, bool CHECK, bool PAIR>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
f_a_fused_kernel(const __grid_constant__ CUtensorMap map_w, FusedParams p) {
    if (p.run_if != nullptr && *p.run_if == 0) return;
    const long long t_entry = p.tim ? clock64() : 0;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int x_tile = TILE_M * BLOCK_K;
    constexpr int NCONV = PAIR ? 2 * CONV_WARPS : CONV_WARPS;  // converter warps that feed one stage
    const int w_tile = p.nt * BLOCK_K;
    const int stage_bytes = LXT * x_tile + p.LW * w_tile;
    uint64_t* bars = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + p.stages;
    uint64_t* tmem_full = bars + 2 * p.stages;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * p.stages + 1);
    volatile uint32_t* nzflag = (volatile uint32_t*)(tmem_slot + 2);  // [stage][converter warp]: non-zero plane mask

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
    // Digit sums D_d that matter: all LX + LW - 1 of them, except that for q = 2^e every D_d with 8 d >= e is a multiple
    // of q (256^d D_d = 0 mod q) -- neither multiplied nor read back (C2: q = 2^24, 5 of the 6 digit pairs remain).
    const int ND = min(LXT + p.LW - 1, p.dmax);
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    const int tile_n = blockIdx.x % p.n_tiles, tile_m = blockIdx.x / p.n_tiles;  // PAIR: n_tiles even, tile_n & 1 == crank
    const int n0 = tile_n * p.nt, m0 = tile_m * TILE_M;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            // TMA expect_tx arrival + one arrival per LOCAL converter warp; the peer's half of a stage (and its plane
            // masks) is accounted in bytes: st.async completes the transaction count that the TMA warp expects
            mbar_init(&full_bar[s], 1 + CONV_WARPS);
            mbar_init(&empty_bar[s], PAIR ? 2 : 1);  // tcgen05.commit (of both CTAs of a pair)
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == CONV_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (warp < 4) {  // zero the accumulators: every MMA accumulates (one MMA spans several adjacent digit accumulators)
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < ND * p.nt; c += 16) tmem_st16_zero(lane_addr + (uint32_t)c);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (PAIR) cluster_sync_all();  // the peer's barriers exist before anything is stored to / arrives at them

    if (warp == CONV_WARPS) {
        // ===== TMA producer: key digit planes =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sw = smem + (size_t)stage * stage_bytes + LXT * x_tile;
                mbar_expect_tx(&full_bar[stage], (uint32_t)(p.LW * w_tile) + (PAIR ? (uint32_t)(LXT * (x_tile / 2) + CONV_WARPS * 4) : 0u));
                for (int i = 0; i < p.LW; ++i) tma_load_3d(&map_w, sw + i * w_tile, &full_bar[stage], kb * BLOCK_K, n0, i);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == CONV_WARPS + 1) {
        // ===== MMA issuer =====
        const uint32_t idesc0 = (2u << 4) | (1u << 7) | ((uint32_t)(p.w_signed ? 1 : 0) << 10) | ((uint32_t)(TILE_M >> 4) << 24);
        const int G = max(1, min(p.LW, 256 / p.nt));
        const uint64_t desc0 = make_desc(smem_u32(smem));
        int stage = 0;
        uint32_t phase = 0;
        unsigned long long units = 0;
        long long tw_full = 0;
        const long long t_loop = p.tim ? clock64() : 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            const long long t0_ = p.tim ? clock64() : 0;
            if (PAIR) mbar_wait_cluster(&full_bar[stage], phase);
            else mbar_wait(&full_bar[stage], phase);
            if (p.tim) tw_full += clock64() - t0_;
            // The converters' generic-proxy stores (acquired through the barrier above) -> the tensor core's async-proxy
            // reads.  The cross-proxy fence sits HERE, on the consumer side of the release/acquire chain, once per k
            // block: in the converter warps it lowers to MEMBAR.ALL.CTA, which also waits for their prefetched global
            // loads of the NEXT k block, i.e. it serialised every k block on a full DRAM round trip.
            fence_proxy_async_smem();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                uint32_t mk = 0;
#pragma unroll
                for (int w = 0; w < NCONV; ++w) mk |= nzflag[stage * NCONV + w];
                const uint32_t sx_off = (uint32_t)(stage * stage_bytes) >> 4;
                const uint32_t sw_off = sx_off + ((uint32_t)(LXT * x_tile) >> 4);
                for (int j = 0; j < LXT; ++j) {
                    if (!((mk >> j) & 1u)) continue;
                    const uint64_t da = desc0 + (uint64_t)(sx_off + ((uint32_t)(j * x_tile) >> 4));
                    for (int i0 = 0; i0 < p.LW; i0 += G) {
                        const int g = min(min(G, p.LW - i0), ND - (i0 + j));  // planes i0 .. i0+g-1 land in D_{i0+j} ..
                        if (g <= 0) break;
                        const uint32_t idesc = idesc0 | ((uint32_t)((g * p.nt) >> 3) << 17);
                        const uint64_t db = desc0 + (uint64_t)(sw_off + ((uint32_t)(i0 * w_tile) >> 4));
                        const uint32_t dcol = tmem_base + (uint32_t)((i0 + j) * p.nt);
#pragma unroll
                        for (int kk = 0; kk < BLOCK_K / 32; ++kk)
                            mma_i8(dcol, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, 1u);
                        units += (unsigned long long)g;
                    }
                }
                if (PAIR) mma_commit_mc(&empty_bar[stage], (uint16_t)3);
                else mma_commit(&empty_bar[stage]);
                if (kb == num_kb - 1) mma_commit(tmem_full);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (lane == 0 && p.mma_units && units)
            atomicAdd(p.mma_units, units * (2ull * TILE_M * BLOCK_K) * (unsigned long long)p.nt);
        if (lane == 0 && p.tim) {
            atomicAdd(p.tim + 0, (unsigned long long)tw_full);
            atomicAdd(p.tim + 1, (unsigned long long)(clock64() - t_loop));
        }
    } else {
        // ===== converters (warps 0..15): int32 sigma -> digit planes in shared memory; then the epilogue =====
        // A register buffer holds 8 row segments of 128 values (4 per lane): 8 rows of one k block, or (PAIR) 4 rows of
        // two consecutive k blocks -- the same 64 KB of loads in flight per SM either way.
        constexpr int R = PAIR ? 4 : 8;  // rows per warp per k block
        constexpr int U = 8 / R;         // k blocks per register buffer
        const int rb = (int)crank * (TILE_M / 2) * (PAIR ? 1 : 0) + warp * R;  // first row (within the tile) of this warp
        const bool do_norm = p.norm2 != nullptr && (PAIR || tile_n == 0);
        NormAcc acc[R];  // saturating: a plain u64 sum of int32 squares wraps (common.cuh)
#pragma unroll
        for (int i = 0; i < R; ++i) acc[i].clear();
        const int32_t* xrow = p.x + ((long)m0 + rb) * p.ldx + lane * 4;
        const int rows_here = max(0, min(R, p.B - (m0 + rb)));
        // fast path: all rows of this warp exist and are 16-byte aligned -> unpredicated 128-bit loads
        const bool fast_rows = p.vec && rows_here == R;
        const long ldx4 = p.ldx >> 2;  // row stride in int4 units (fast path only)
        auto load_buf = [&](int4 (&dst)[8], int kb0) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int kb = kb0 + u;
                if (fast_rows && (kb + 1) * BLOCK_K <= p.K) {
                    const int4* src = reinterpret_cast<const int4*>(xrow) + kb * (BLOCK_K / 4);
#pragma unroll
                    for (int i = 0; i < R; ++i) dst[u * R + i] = src[i * ldx4];
                    continue;
                }
                const int col = kb * BLOCK_K + lane * 4;
#pragma unroll 1
                for (int i = 0; i < R; ++i) {
                    int4 v = make_int4(0, 0, 0, 0);
                    if (i < rows_here && kb < num_kb) {
                        const int32_t* src = xrow + (long)i * p.ldx + kb * BLOCK_K;
                        if (col + 0 < p.K) v.x = src[0];
                        if (col + 1 < p.K) v.y = src[1];
                        if (col + 2 < p.K) v.z = src[2];
                        if (col + 3 < p.K) v.w = src[3];
                    }
#pragma unroll
                    for (int j = 0; j < R; ++j)
                        if (j == i) dst[u * R + j] = v;  // static register indices
                }
            }
        };
        // Shared-memory address of this lane's 4 digits in row r = rb + i: SWIZZLE_128B puts 16-byte chunk c of row r at
        // chunk c ^ (r & 7), and r & 7 = (rb & 7) ^ i (rb is a multiple of R, i < R).  All addresses are 32-bit
        // shared-window addresses and the stores are st.shared (a generic store costs a 64-bit address computation per
        // row): row i of stage s sits at ((stage base + warp_base) ^ (i << 4)) + 128 i, the stage bases being
        // 1024-byte aligned and warp_base carrying the lane's chunk in bits 2..6.
        const uint32_t lane_const = (uint32_t)(((lane >> 2) << 4) | ((lane & 3) << 2));
        const uint32_t smem_base = smem_u32(smem);
        const uint32_t warp_base = smem_base + (uint32_t)(rb * 128) + (lane_const ^ (uint32_t)((rb & 7) << 4));
        const uint32_t nz_base = smem_u32((const void*)nzflag) + 4u * ((uint32_t)(crank * CONV_WARPS) * (PAIR ? 1u : 0u) + (uint32_t)warp);
        // the peer CTA's window: same offsets, shifted by a constant
        const uint32_t peer_delta = PAIR ? mapa_shared(smem_base, crank ^ 1u) - smem_base : 0u;
        int stage = 0;
        uint32_t phase = 0;
        long long tw_empty = 0;
        const bool timed = p.tim != nullptr && warp == 0;
        auto convert_unit = [&](const int4 (&buf)[8], const int u) {
            const long long t0_ = timed ? clock64() : 0;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (timed) tw_empty += clock64() - t0_;
            const uint32_t sx = warp_base + (uint32_t)(stage * stage_bytes);
            const uint32_t rbar = smem_u32(&full_bar[stage]) + peer_delta;  // the peer's full barrier of this stage
            // plane 0 is never skipped (it is zero only for an all-zero block, which costs one MMA group)
            uint32_t nz1 = 0, nz2 = 0, nz3 = 0, ovf = 0;
            if (do_norm) {  // one uniform branch: the n tiles that do not own the norms issue none of these
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const int w0 = buf[u * R + i].x, w1 = buf[u * R + i].y, w2 = buf[u * R + i].z, w3 = buf[u * R + i].w;
                    acc[i].add_sq4(w0, w1, w2, w3);
                }
            }
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int w0 = buf[u * R + i].x, w1 = buf[u * R + i].y, w2 = buf[u * R + i].z, w3 = buf[u * R + i].w;
                const uint32_t dst = (sx ^ (uint32_t)(i << 4)) + (uint32_t)(i * 128);
                const uint32_t rdst = dst + peer_delta;  // same offset in the peer's window
                // Balanced digits d_l of v: v_0 = v, d_l = low byte of v_l (as s8), v_{l+1} = (v_l + 128) >> 8.
                // With t = v + 0x8080 (0x808080 for four digits): d_0 = byte0(t) ^ 0x80, d_1 = byte1(t) ^ 0x80,
                // d_2 = byte2(t) [^ 0x80 when a fourth digit follows], d_3 = byte3(t): one add per value, the 4 x 4
                // byte transpose done with PRMT.
                constexpr int LB = LXT + (CHECK ? 1 : 0);  // digits of the representation
                constexpr int BIAS = LB == 1 ? 0 : LB == 2 ? 0x80 : LB == 3 ? 0x8080 : 0x808080;
                const uint32_t t0 = (uint32_t)(w0 + BIAS), t1 = (uint32_t)(w1 + BIAS), t2 = (uint32_t)(w2 + BIAS),
                               t3 = (uint32_t)(w3 + BIAS);
                if (CHECK) ovf = (ovf | t0 | t1) | (t2 | t3);
                const uint32_t lo01 = __byte_perm(t0, t1, 0x5140), lo23 = __byte_perm(t2, t3, 0x5140);  // b0 b0' b1 b1'
                {
                    const uint32_t pk = __byte_perm(lo01, lo23, 0x5410) ^ (LB > 1 ? 0x80808080u : 0u);
                    sts32(dst, pk);
                    if (PAIR) st_async32(rdst, pk, rbar);
                }
                if (LXT > 1) {
                    const uint32_t pk = __byte_perm(lo01, lo23, 0x7632) ^ (LB > 2 ? 0x80808080u : 0u);
                    sts32(dst + (uint32_t)x_tile, pk);
                    if (PAIR) st_async32(rdst + (uint32_t)x_tile, pk, rbar);
                    nz1 |= pk;
                }
                if (LXT > 2) {
                    const uint32_t hi01 = __byte_perm(t0, t1, 0x7362), hi23 = __byte_perm(t2, t3, 0x7362);  // b2 b2' b3 b3'
                    const uint32_t pk = __byte_perm(hi01, hi23, 0x5410) ^ (LB > 3 ? 0x80808080u : 0u);
                    sts32(dst + (uint32_t)(2 * x_tile), pk);
                    if (PAIR) st_async32(rdst + (uint32_t)(2 * x_tile), pk, rbar);
                    nz2 |= pk;
                    if (LXT > 3) {
                        const uint32_t pk3 = __byte_perm(hi01, hi23, 0x7632);
                        sts32(dst + (uint32_t)(3 * x_tile), pk3);
                        if (PAIR) st_async32(rdst + (uint32_t)(3 * x_tile), pk3, rbar);
                        nz3 |= pk3;
                    }
                }
            }
            uint32_t nzm = 1u | (nz1 ? 2u : 0u) | (nz2 ? 4u : 0u) | (nz3 ? 8u : 0u);
            if (CHECK && (ovf >> (8 * LXT)) != 0u) nzm |= 0x80u;  // a digit beyond LXT is non-zero
            nzm = __reduce_or_sync(0xffffffffu, nzm);
            if (CHECK && (nzm & 0x80u) && lane == 0) atomicOr(p.overflow, 1);
            __syncwarp();  // (the cross-proxy fence is issued by the MMA warp after it has acquired the stage)
            if (lane == 0) {
                const uint32_t nza = nz_base + (uint32_t)(stage * NCONV * 4);
                sts32(nza, nzm);
                if (PAIR) st_async32(nza + peer_delta, nzm, rbar);
                mbar_arrive(&full_bar[stage]);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        };
        // two register buffers, the k loop unrolled by two buffers: the 8 row-segment loads of the next buffer are in
        // flight while the current one is converted, without register moves between the buffers
        int4 buf_a[8], buf_b[8];
        const long long t_conv = timed ? clock64() : 0;
        load_buf(buf_a, 0);
        for (int kb = 0; kb < num_kb; kb += 2 * U) {
            load_buf(buf_b, kb + U);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (kb + u < num_kb) convert_unit(buf_a, u);
            if (kb + U < num_kb) {
                load_buf(buf_a, kb + 2 * U);
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (kb + U + u < num_kb) convert_unit(buf_b, u);
            }
        }
        if (do_norm) {
#pragma unroll
            for (int i = 0; i < R; ++i) {
                NormAcc a = acc[i];
                a.warp_reduce();
                const long grow = (long)m0 + rb + i;
                if (lane == 0 && grow < p.B) p.norm2[grow] = a.value();
            }
        }
        const long long t_epi = timed ? clock64() : 0;
        if (timed && lane == 0) {
            atomicAdd(p.tim + 2, (unsigned long long)tw_empty);
            atomicAdd(p.tim + 3, (unsigned long long)(t_epi - t_conv));
        }
        // ---- epilogue: V = sum_d 256^d D_d, reduce mod q ----
        const int lg = warp & 3;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + lg * 32 + lane;
        const bool pow2 = (p.q & (p.q - 1)) == 0;
        for (int c0 = (warp >> 2) * 16; c0 < p.nt; c0 += 16 * (CONV_WARPS / 4)) {
            if (ND <= 4) {
                // |V| < 2^(31 + 8 (ND - 1) + 1) <= 2^56: 64-bit recombination, Barrett reduction
                long long v[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = 0;
                for (int d = 0; d < ND; ++d) {
                    int32_t t[16];
                    tmem_ld16(lane_addr + (uint32_t)(d * p.nt + c0), t);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 16; ++c) v[c] += (long long)t[c] * (1ll << (8 * d));
                }
                if (row < p.B) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const int n = n0 + c0 + c;
                        if (n >= p.N) continue;
                        long long r;
                        if (!p.q) r = v[c];
                        else if (pow2) r = (long long)((unsigned long long)v[c] & (p.q - 1));
                        else r = (long long)mod_i64_barrett(v[c], p.q, p.qmagic);
                        p.out[(long)row * p.ldout + n] = r;
                    }
                }
                continue;
            }
            __int128 v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = 0;
            for (int d = 0; d < ND; ++d) {
                int32_t t[16];
                tmem_ld16(lane_addr + (uint32_t)(d * p.nt + c0), t);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] += ((__int128)t[c]) << (8 * d);
            }
            if (row < p.B) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int n = n0 + c0 + c;
                    if (n >= p.N) continue;
                    long long r;
                    if (p.q) {
                        if (pow2) r = (long long)((unsigned long long)v[c] & (p.q - 1));
                        else r = (long long)mod_i128(v[c], p.q);
                    } else {
                        r = (long long)v[c];
                    }
                    p.out[(long)row * p.ldout + n] = r;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (timed && lane == 0) {
            const long long t_end = clock64();
            atomicAdd(p.tim + 4, (unsigned long long)(t_end - t_epi));
            atomicAdd(p.tim + 5, (unsigned long long)(t_end - t_entry));
        }
    }
    __syncthreads();
    if (warp == CONV_WARPS + 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
    // no CTA of a pair may exit while the peer can still store to its shared memory or arrive on its barriers
    if (PAIR) cluster_sync_all();
}

}  // namespace

// coordinate tile: the widest the 512 TMEM columns allow, then balanced over the resulting number of tiles
static int fused_tile_n(int LX, int LW, int N) {
    const int nt_max = qf_i8_tile_n(LX, LW, N);
    if (nt_max < 16) return nt_max;
    const int tiles = (N + nt_max - 1) / nt_max;
    return std::min(nt_max, ((N + tiles - 1) / tiles + 15) / 16 * 16);
}

template <bool PAIR>
static void (*pick_kernel(int LX, bool check))(const CUtensorMap, FusedParams) {
    if (check) return LX == 1 ? f_a_fused_kernel<1, true, PAIR> : LX == 2 ? f_a_fused_kernel<2, true, PAIR> : f_a_fused_kernel<3, true, PAIR>;
    return LX == 1 ? f_a_fused_kernel<1, false, PAIR> : LX == 2 ? f_a_fused_kernel<2, false, PAIR>
         : LX == 3 ? f_a_fused_kernel<3, false, PAIR> : f_a_fused_kernel<4, false, PAIR>;
}

static cudaError_t launch_one(const FaFusedArgs& a, int LX, bool check, const int* run_if, cudaStream_t stream) {
    const int nt = fused_tile_n(LX, a.LW, a.N);
    if (nt < 16) return cudaErrorInvalidValue;
    FusedParams p{};
    p.x = a.x; p.ldx = a.ldx; p.B = a.B; p.N = a.N; p.K = a.K; p.LX = LX; p.LW = a.LW; p.nt = nt; p.w_signed = a.w_signed;
    p.vec = ((a.ldx & 3) == 0 && (((uintptr_t)a.x) & 15) == 0) ? 1 : 0;
    p.q = a.q; p.qmagic = qf_barrett_magic(a.q); p.out = a.out; p.ldout = a.ldout; p.norm2 = a.norm2; p.mma_units = a.mma_units;
    p.overflow = a.retry_flag; p.run_if = run_if; p.tim = a.tim;
    p.dmax = 64;
    if (a.q && (a.q & (a.q - 1)) == 0) {
        int e = 0;
        while ((1ull << e) < a.q) ++e;
        p.dmax = e == 0 ? 1 : (e + 7) / 8;
    }
    p.n_tiles = (a.N + nt - 1) / nt;
    const int m_tiles = (a.B + tc05::TILE_M - 1) / tc05::TILE_M;
    const int stage_bytes = LX * tc05::TILE_M * tc05::BLOCK_K + a.LW * nt * tc05::BLOCK_K;
    const int budget = 227 * 1024 - 1024 - 1024;  // alignment slack, barriers + plane masks
    int stages = budget / stage_bytes;
    if (stages < 2) return cudaErrorInvalidValue;
    if (stages > 6) stages = 6;
    p.stages = stages;
    const int smem = stages * stage_bytes + 1024 + 1024;
    CUtensorMap mw;
    if (!tc05::make_map(&mw, a.w, a.K, a.N, a.LW, a.ldw, a.w_plane, nt)) return cudaErrorInvalidValue;
    // CTA pairs (clusters of 2 along the coordinate tiles) whenever the tiles pair up; QF_FA_PAIR=0 keeps single CTAs
    static const bool pair_off = getenv("QF_FA_PAIR") && getenv("QF_FA_PAIR")[0] == '0';  // test-only switch, read once
    const bool pair = !pair_off && (p.n_tiles % 2 == 0);
    void (*kern)(const CUtensorMap, FusedParams) = pair ? pick_kernel<true>(LX, check) : pick_kernel<false>(LX, check);
    static int configured_dev[QF_MAX_DEVICES][2][2][MAX_LX + 1] = {};
    auto& configured = configured_dev[qf_device_slot()];
    if (smem > configured[pair ? 1 : 0][check ? 1 : 0][LX]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured[pair ? 1 : 0][check ? 1 : 0][LX] = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(m_tiles * p.n_tiles));
    cfg.blockDim = dim3(FUSED_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pair ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, mw, p);
}

// LX is the digit count that covers every in-domain value, LX_typical the count that covers what the samplers
// produce (|x| <= 6 s r).  With a retry flag and LX_typical < LX the launch is optimistic: LX_typical digits first
// (a wider coordinate tile, fewer passes over sigma), and the full-width kernel runs only if some value really
// needed more (decided on the device, no host round trip).
cudaError_t qf_launch_f_a_fused(const FaFusedArgs& a, cudaStream_t stream) {
    if (a.B <= 0 || a.N <= 0) return cudaSuccess;
    if (a.LX < 1 || a.LX > MAX_LX || a.LW < 1 || a.LX + a.LW - 1 > 16 || a.K < 1) return cudaErrorInvalidValue;
    if ((a.ldw & 15) || (a.w_plane & 15) || (((uintptr_t)a.w) & 15)) return cudaErrorMisalignedAddress;
    const bool opt = a.retry_flag && a.LX_typical >= 1 && a.LX_typical < a.LX && a.LX_typical <= 3;
    const int t_full = fused_tile_n(a.LX, a.LW, a.N), t_opt = opt ? fused_tile_n(a.LX_typical, a.LW, a.N) : t_full;
    if (opt && (a.N + t_opt - 1) / t_opt < (a.N + t_full - 1) / t_full) {
        cudaError_t e = cudaMemsetAsync(a.retry_flag, 0, sizeof(int), stream);
        if (e != cudaSuccess) return e;
        e = launch_one(a, a.LX_typical, true, nullptr, stream);
        if (e != cudaSuccess) return e;
        return launch_one(a, a.LX, false, a.retry_flag, stream);
    }
    return launch_one(a, a.LX, false, nullptr, stream);
}
