// Exact integer "targets x key-matrix" contraction on the 5th-gen tensor cores:
// limb-split int8 tcgen05.mma (kind::i8, s32 accumulators in TMEM) fed by TMA.
//
//   V[b][n] = sum_k x[b][k] * w[n][k]          x, w multi-limb integers
//   x = sum_j 256^j x_j  (balanced s8 limbs)     w = sum_i 256^i w_i  (u8 limbs of a residue,
//                                                 or balanced s8 limbs of a signed entry)
//   V = sum_d 256^d D_d ,   D_d = sum_{i+j=d} sum_k x_j[b][k] w_i[n][k]
//
// Every D_d lives in its own TMEM accumulator (128 lanes x NT columns of s32); all limb pairs
// with i+j = d accumulate into the same one, so the epilogue reads ND = LX+LW-1 accumulators
// and recombines them exactly (int64 / int128), then reduces mod q or stores the integer.
// Used for f_a = A sigma (gpv.rs:190-193), v = u - A p and e = p + [R;I] z
// (mp_perturbation.rs:318,335), e = sol + S z (gpv.rs:160) and TrapGen's A_bar R
// (gadget_classical.rs:66).
//
// CTA = 10 warps: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer (one elected lane),
// warps 2..9 epilogue (tcgen05.ld 32x32b, one TMEM lane = one target row per thread, two warps per lane group).
// Tile: 128 targets x NT coordinates x 128-byte K blocks, SWIZZLE_128B K-major operands,
// multi-stage mbarrier pipeline.  K is long (m ~ 10^4) and the epilogue is < 3 % of a tile,
// so one tile per CTA (no TMEM double buffering).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tc05.cuh"

namespace {

using namespace tc05;
constexpr int EPI_WARPS = 8;  // two warps per TMEM lane group, each draining half of the tile's columns
constexpr int I8_THREADS = (2 + EPI_WARPS) * 32;

struct I8Params {
    int B, N, K;
    int LX, LW;
    int nt;        // N tile (multiple of 16)
    int stages;
    int bk;        // K block in bytes: 128 (SWIZZLE_128B tiles) or 64 (SWIZZLE_64B, twice the pipeline depth)
    int w_signed;  // b_format: 1 = s8 limbs, 0 = u8 limbs
    // epilogue
    int out_kind;  // 0: int64 (optionally mod q, optional base, sign), 1: int32 store, 2: fp64 accumulate (+=)
    int sign;      // +1 / -1 applied to V
    unsigned long long q;  // 0: no reduction
    unsigned long long qmagic;  // floor(2^64 / q): Barrett reduction of the 64-bit epilogue
    const int64_t* base; long ldbase;
    void* out; long ldout;
    int* flag;
    const double* scale;        // out_kind 3: out[b][n] -= V * scale[n]
    // L2-aware rasterisation of the 1-D grid: groups of group_m target tiles, n fastest across groups
    int m_tiles, n_tiles, group_m;
    // optional zero-tile map of the x digit planes: nz[(j * m_tiles_total + m_tile) * kb_total + kb_off + kb]
    const uint8_t* x_nz; int nz_m_tiles, nz_kb_total, nz_kb_off, nz_m_off;
    unsigned long long* mma_units;  // optional device counter of executed int8 operations (tensor pipe work)
    // digit-sum window: only D_d with d >= d_lo are computed (accumulator t holds D_{d_lo + t}); the dropped low-order
    // pairs lie below the error budget of the fixed-point update (out_kind 3 only, scale_mul = 256^d_lo)
    int d_lo;
    double scale_mul;
    int overwrite;  // out_kind 3: out = -V * scale (no read of the old values) instead of out -= V * scale
    int acc_bufs;   // 2: the accumulators are double-buffered in TMEM (tile i+1's MMAs overlap tile i's epilogue)
    // conditional launch: the grid exits at once unless gate_lo <= *gate <= gate_hi (device-side choice between
    // variants of the same contraction compiled for different digit counts, no host round trip)
    const int* gate; int gate_lo, gate_hi;
    // structured w: rows [n0, n0 + nt) of the key matrix only have non-zero columns k < n0 + nt + tri_slack (tri_mode 1,
    // lower block-triangular), k >= K - (n0 + nt) - tri_slack (2, the same with the columns reversed), k < K - n0 + tri_slack
    // (3, upper block-triangular with the columns reversed) or k >= n0 - tri_slack (4, upper block-triangular): the
    // k blocks outside that range are neither loaded nor multiplied
    int tri_mode, tri_slack;
    // out_kind 3, full 32 / 64-column tiles: the read-modify-write of T goes through a per-warp shared-memory staging tile
    // (coalesced 128-byte row segments on the global side, lane = row on the TMEM side) instead of 16-byte accesses to
    // 32 different rows per instruction; epi_stage = byte offset of the staging area behind the barriers (0: none)
    int epi_stage;
    // optional diagnostics (QF_TRACE): cycles the MMA warp waited for [0] drained accumulators, [1] operand tiles, and the
    // epilogue spent [2] waiting for the accumulators, [3] draining them (summed over CTAs)
    unsigned long long* tim;
};

// s32 -> f64 without the conversion unit: 2^52 + 2^31 + x is exactly representable, built by bit insertion
__device__ __forceinline__ double s32_to_f64(int32_t x) {
    return __hiloint2double(0x43300000, (int)((uint32_t)x ^ 0x80000000u)) - 4503601774854144.0;  // 2^52 + 2^31
}

// Epilogue of one 128 x 16 half of a 32-column tile of the scaled fp64 update out[b][n] = old - V * scale[n],
// V = sum_d 256^d D_d: NDT accumulators of 32 columns each (NDT compile time: straight-line Horner, no predication),
// 4 columns per TMEM trip, the old values `pre` already in registers.  cb = first column of this warp's half.
template <int NDT, int NC = 16>
__device__ __forceinline__ void epi_update16(uint32_t lane_addr, int cb, const double2 (&pre)[16], double* orow,
                                             const double* __restrict__ scale, int nd_rt = NDT, int nt = 32,
                                             double mul = 1.0) {
#pragma unroll
    for (int c0 = 0; c0 < NC; c0 += 4) {
        int32_t t[NDT][4];
#pragma unroll
        for (int d = 0; d < NDT; ++d)
            if (d < nd_rt) tmem_ld4(lane_addr + (uint32_t)(d * nt + cb + c0), t[d]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        double r[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            // Horner in fp64 from the top digit: rounding stays at 2^-53 of the running value
            double dv = 0.0;
#pragma unroll
            for (int d = NDT - 1; d >= 0; --d)
                if (d < nd_rt) dv = fma(dv, 256.0, s32_to_f64(t[d][c]));
            const double old = (c & 1) ? pre[(c0 + c) >> 1].y : pre[(c0 + c) >> 1].x;
            r[c] = fma(-dv, scale[cb + c0 + c] * mul, old);
        }
        // streaming stores / loads for T: every value is touched once per launch; evict-first keeps the digit planes in L2
        __stcs(reinterpret_cast<double2*>(orow + cb + c0), make_double2(r[0], r[1]));
        __stcs(reinterpret_cast<double2*>(orow + cb + c0 + 2), make_double2(r[2], r[3]));
    }
}

// General tile width / ragged edges: COLS columns per TMEM trip, at most NDT accumulators (nd of them live),
// old values fetched before the accumulators are read.
template <int NDT, int COLS>
__device__ __forceinline__ void epi_update_any(uint32_t lane_addr, double* orow, const double* __restrict__ scale, int n0,
                                               int nt, int N, bool rv, int nd, int cbeg, int cend, double mul,
                                               bool overwrite) {
    for (int c0 = cbeg; c0 < cend; c0 += COLS) {
        double told[COLS];
#pragma unroll
        for (int c = 0; c < COLS; ++c) told[c] = (rv && !overwrite && n0 + c0 + c < N) ? orow[n0 + c0 + c] : 0.0;
        int32_t t[NDT][COLS];
#pragma unroll
        for (int d = 0; d < NDT; ++d)
            if (d < nd) {
                if (COLS == 8) tmem_ld8(lane_addr + (uint32_t)(d * nt + c0), t[d]);
                else tmem_ld4(lane_addr + (uint32_t)(d * nt + c0), t[d]);
            }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            double dv = 0.0;
#pragma unroll
            for (int d = NDT - 1; d >= 0; --d)
                if (d < nd) dv = fma(dv, 256.0, s32_to_f64(t[d][c]));
            const int n = n0 + c0 + c;
            if (rv && n < N) orow[n] = told[c] - dv * (scale[n] * mul);
        }
    }
}

// 16 columns of the scaled fp64 update for one row (lane = TMEM lane): v <- v - V * scale[n] * mul, in registers.
template <int NDT>
__device__ __forceinline__ void epi_cols16(uint32_t lane_addr, int cb, double2 (&v)[8], const double* __restrict__ scale,
                                           int nd, int nt, double mul) {
#pragma unroll
    for (int c0 = 0; c0 < 16; c0 += 4) {
        double sc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) sc[c] = __ldg(scale + c0 + c) * mul;
        int32_t t[NDT][4];
#pragma unroll
        for (int d = 0; d < NDT; ++d)
            if (d < nd) tmem_ld4(lane_addr + (uint32_t)(d * nt + cb + c0), t[d]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double dv = 0.0;  // Horner in fp64 from the top digit
#pragma unroll
            for (int d = NDT - 1; d >= 0; --d)
                if (d < nd) dv = fma(dv, 256.0, s32_to_f64(t[d][c]));
            double& o = (c & 1) ? v[(c0 + c) >> 1].y : v[(c0 + c) >> 1].x;
            o = fma(-dv, sc[c], o);
        }
    }
}
constexpr int EPI_STG_LD = 18;  // doubles per staged row: 16 columns + 2 (rows stay 16-byte aligned, lane = row reads are conflict free)
constexpr int EPI_STG_BYTES = EPI_WARPS * 32 * EPI_STG_LD * 8;

// Persistent: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  TMEM and the barriers are set up
// once; the TMA producer runs ahead into the next tile while the epilogue of the current one drains TMEM.
// OK3 = 1: the scaled fp64 update epilogue (out_kind 3) only; 0: the integer epilogues.
// PAIR: launched as clusters of two CTAs that own the two adjacent coordinate tiles (2p, 2p + 1) of the same 128 targets.
// The x (target digit) tile is the same for both: each CTA fetches HALF of its rows and TMA-multicasts them into both
// CTAs' operand rings, so every x block crosses L2 -> SM once per cluster instead of once per CTA (the x planes are the
// larger operand: 16 KB per digit plane against nt x 128 bytes per w plane).  A stage is free again when BOTH tensor
// cores have retired it (tcgen05.commit multicast onto both CTAs' empty barriers).  map_xh = map_x with 64-row boxes.
template <int OK3, bool PAIR, bool STG = false>
__global__ void __launch_bounds__(I8_THREADS, 1)
gemm_i8_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_xh, I8Params p) {
    if (p.gate != nullptr) {  // uniform over the grid (and over a cluster): either every thread leaves here or none does
        const int gv = *p.gate;
        if (gv < p.gate_lo || gv > p.gate_hi) return;
    }
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte aligned operand ring
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int BK = p.bk;
    const int x_tile = TILE_M * BK, w_tile = p.nt * BK;
    const int stage_bytes = p.LX * x_tile + p.LW * w_tile;
    uint64_t* bars = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + p.stages;
    uint64_t* tmem_full = bars + 2 * p.stages;       // [2], one per accumulator buffer
    uint64_t* tmem_empty = bars + 2 * p.stages + 2;  // [2]
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * p.stages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + BK - 1) / BK;
    const int ND = p.LX + p.LW - 1 - p.d_lo;   // accumulators per buffer: digit sums d_lo .. LX + LW - 2
    const int NB = p.acc_bufs;                 // accumulator buffers in TMEM (NB * ND * nt <= 512 columns)
    const uint32_t buf_cols = (uint32_t)(ND * p.nt);
    // PAIR: the walk is over tile PAIRS (n_tiles / 2 columns of pairs); this CTA takes coordinate tile 2 * pair + crank
    const int n_cols = PAIR ? p.n_tiles / 2 : p.n_tiles;
    const int total_tiles = p.m_tiles * n_cols;
    const int walk0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, walk_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    // tile rasterisation: consecutive tiles walk group_m target tiles for one n tile, then the next n tile
    auto tile_coords = [&](int t, int& tile_m, int& tile_n) {
        const int per_group = p.group_m * n_cols;
        const int g = t / per_group, r = t - g * per_group;
        const int gm = min(p.group_m, p.m_tiles - g * p.group_m);  // last group may be short
        tile_m = g * p.group_m + r % gm;
        tile_n = PAIR ? 2 * (r / gm) + (int)crank : r / gm;
    };
    // k blocks [kb0, kb1) of a tile (structured w; a CTA pair walks the union of its two tiles' ranges in lockstep)
    auto kb_range = [&](int tile_n, int& kb0, int& kb1) {
        kb0 = 0; kb1 = num_kb;
        if (p.tri_mode == 0) return;
        const int nlo = (PAIR ? (tile_n & ~1) : tile_n) * p.nt, nhi = min(p.N, nlo + (PAIR ? 2 : 1) * p.nt);
        if (p.tri_mode == 1) kb1 = min(num_kb, (min(p.K, nhi + p.tri_slack) + BK - 1) / BK);
        else if (p.tri_mode == 2) kb0 = max(0, p.K - nhi - p.tri_slack) / BK;
        else if (p.tri_mode == 3) kb1 = min(num_kb, (min(p.K, max(1, p.K - nlo + p.tri_slack)) + BK - 1) / BK);
        else kb0 = max(0, min(p.K - 1, nlo - p.tri_slack)) / BK;
    };
    // which x digit planes are non-zero in this (target tile, k block)? (bit j of the returned mask)
    auto plane_mask = [&](int tile_m, int kb) -> uint32_t {
        if (!p.x_nz) return (1u << p.LX) - 1u;
        uint32_t mk = 0;
        for (int j = 0; j < p.LX; ++j)
            if (p.x_nz[((size_t)j * p.nz_m_tiles + (p.nz_m_off + tile_m)) * p.nz_kb_total + p.nz_kb_off + ((kb * BK) >> 7)]) mk |= 1u << j;
        return mk;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], PAIR ? 2 : 1);  // tcgen05.commit of this CTA (and of its peer)
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], EPI_WARPS);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // allocate all 512 TMEM columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // zero every accumulator so that all MMAs can accumulate: one MMA then covers several digit planes of w
    // (N = g * nt) whose accumulators D_{j+i0} .. D_{j+i0+g-1} are adjacent column blocks.  Later tiles find the
    // accumulators zeroed by the epilogue of the previous tile.
    if (warp >= 2) {
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (int c = ((warp - 2) >> 2) * 16; c < NB * ND * p.nt; c += 16 * (EPI_WARPS / 4)) tmem_st16_zero(lane_addr + (uint32_t)c);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (PAIR) cluster_sync_all();  // the peer's barriers exist before anything is multicast to / arrives at them

    if (warp == 0) {
        // ===== TMA producer =====
        // The zero-plane masks of 32 consecutive k blocks are fetched by the 32 lanes at once (one L2 round trip per 32
        // k blocks instead of one per k block in front of every TMA issue) and handed out by shuffle.
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = walk0; tile < total_tiles; tile += walk_step) {
            int tile_m, tile_n;
            tile_coords(tile, tile_m, tile_n);
            const int n0 = tile_n * p.nt, m0 = tile_m * TILE_M;
            uint32_t mk_cache = 0;
            int kb0, kb1;
            kb_range(tile_n, kb0, kb1);
            for (int kb = kb0; kb < kb1; ++kb) {
                if (kb == kb0 || (kb & 31) == 0) mk_cache = ((kb & ~31) + lane < num_kb) ? plane_mask(tile_m, (kb & ~31) + lane) : 0u;
                const uint32_t mk = __shfl_sync(0xffffffffu, mk_cache, kb & 31);
                if (lane == 0) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sx = smem + (size_t)stage * stage_bytes;
                    uint8_t* sw = sx + p.LX * x_tile;
                    if (mk == 0) {  // nothing to multiply in this k block: just hand the stage over
                        mbar_expect_tx(&full_bar[stage], 0);
                    } else {
                        mbar_expect_tx(&full_bar[stage], (uint32_t)(__popc(mk) * x_tile + p.LW * w_tile));
                        for (int j = 0; j < p.LX; ++j) {
                            if (!((mk >> j) & 1u)) continue;
                            if (PAIR)  // rows [64 crank, 64 crank + 64) of the tile, into both CTAs
                                tma_load_3d_mc(&map_xh, sx + j * x_tile + (int)crank * (x_tile / 2), &full_bar[stage], kb * BK,
                                               m0 + (int)crank * (TILE_M / 2), j, (uint16_t)3);
                            else tma_load_3d(&map_x, sx + j * x_tile, &full_bar[stage], kb * BK, m0, j);
                        }
                        for (int i = 0; i < p.LW; ++i) tma_load_3d(&map_w, sw + i * w_tile, &full_bar[stage], kb * BK, n0, i);
                    }
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor: D = s32, A = s8 (x digits), B = u8/s8 (w digits), K-major both, M = 128;
        // N = g * nt covers g adjacent digit planes of w (their smem tiles and their accumulators are contiguous)
        const uint32_t idesc0 = (2u << 4) | (1u << 7) | ((uint32_t)(p.w_signed ? 1 : 0) << 10) | ((uint32_t)(TILE_M >> 4) << 24);
        const int G = max(1, min(p.LW, 256 / p.nt));
        const uint64_t desc0 = make_desc(smem_u32(smem), BK);
        int stage = 0;
        uint32_t phase = 0;
        unsigned long long units = 0;  // executed (x digit, w digit, k block) products, for the profiler
        long long tw_empty = 0, tw_full = 0;
        int it = 0;
        for (int tile = walk0; tile < total_tiles; tile += walk_step, ++it) {
            int tile_m, tile_n;
            tile_coords(tile, tile_m, tile_n);
            const int buf = NB == 2 ? (it & 1) : 0, use = NB == 2 ? (it >> 1) : it;
            const uint32_t acc_base = tmem_base + (uint32_t)buf * buf_cols;
            if (use > 0) {  // the epilogue must have drained (and re-zeroed) this buffer's previous tile
                const long long t0_ = p.tim ? clock64() : 0;
                mbar_wait(&tmem_empty[buf], (uint32_t)((use - 1) & 1));
                if (p.tim) tw_empty += clock64() - t0_;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            uint32_t mk_cache = 0;
            int kb0, kb1;
            kb_range(tile_n, kb0, kb1);
            for (int kb = kb0; kb < kb1; ++kb) {
                if (kb == kb0 || (kb & 31) == 0)  // as the producer
                    mk_cache = ((kb & ~31) + lane < num_kb) ? plane_mask(tile_m, (kb & ~31) + lane) : 0u;
                const uint32_t mk = __shfl_sync(0xffffffffu, mk_cache, kb & 31);
                {
                    const long long t0_ = p.tim ? clock64() : 0;
                    mbar_wait(&full_bar[stage], phase);
                    if (p.tim) tw_full += clock64() - t0_;
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t sx_off = (uint32_t)(stage * stage_bytes) >> 4;
                    const uint32_t sw_off = sx_off + ((uint32_t)(p.LX * x_tile) >> 4);
                    for (int j = 0; j < p.LX; ++j) {
                        if (!((mk >> j) & 1u)) continue;
                        const uint64_t da = desc0 + (uint64_t)(sx_off + ((uint32_t)(j * x_tile) >> 4));
                        // (groups of G planes, the remainder last: 5 planes at nt = 64 go as N = 256 + 64; the balanced split
                        // N = 192 + 128 measured 12 % slower over the whole C2 step)
                        for (int i0 = max(0, p.d_lo - j); i0 < p.LW;) {  // pairs with i + j < d_lo are dropped
                            const int g = min(G, p.LW - i0);
                            const uint32_t idesc = idesc0 | ((uint32_t)((g * p.nt) >> 3) << 17);
                            const uint64_t db = desc0 + (uint64_t)(sw_off + ((uint32_t)(i0 * w_tile) >> 4));
                            const uint32_t dcol = acc_base + (uint32_t)((i0 + j - p.d_lo) * p.nt);
#pragma unroll
                            for (int kk = 0; kk < BK / 32; ++kk)  // +32 bytes along K = +2 in 16-byte units
                                mma_i8(dcol, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, 1u);
                            units += (unsigned long long)g;
                            i0 += g;
                        }
                    }
                    // frees this smem stage (in both CTAs of a pair) when the MMAs above retire
                    if (PAIR) mma_commit_mc(&empty_bar[stage], (uint16_t)3);
                    else mma_commit(&empty_bar[stage]);
                    if (kb == kb1 - 1) mma_commit(&tmem_full[buf]);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
        if (lane == 0 && p.tim) { atomicAdd(p.tim + 0, (unsigned long long)tw_empty); atomicAdd(p.tim + 1, (unsigned long long)tw_full); }
        if (lane == 0 && p.mma_units && units)
            atomicAdd(p.mma_units, units * (2ull * TILE_M * BK) * (unsigned long long)p.nt);  // int8 operations
    } else {
        // ===== epilogue: warps 2..9, TMEM lane group = warp % 4; the two warps of a lane group split the columns =====
        const int lg = warp & 3, half = (warp - 2) >> 2;
        // column range of this warp for tiles of any width: whole 16-column groups, the first half rounded up
        const int cmid = ((p.nt / 16 + 1) / 2) * 16;
        const int cbeg = half ? cmid : 0, cend = half ? p.nt : cmid;
        const uint32_t lane_addr0 = tmem_base + ((uint32_t)(lg * 32) << 16);
        long long te_wait = 0, te_body = 0;
        int it = 0;
        for (int tile = walk0; tile < total_tiles; tile += walk_step, ++it) {
            int tile_m, tile_n;
            tile_coords(tile, tile_m, tile_n);
            const int n0 = tile_n * p.nt, m0 = tile_m * TILE_M;
            const int row = m0 + lg * 32 + lane;
            const int buf = NB == 2 ? (it & 1) : 0, use = NB == 2 ? (it >> 1) : it;
            const uint32_t lane_addr = lane_addr0 + (uint32_t)buf * buf_cols;
            // out_kind 3, 32-column tiles: this thread's 32 old values (one 256-byte row segment) are fetched NOW, while
            // the tile's MMAs are still running, so the DRAM latency of the read-modify-write is off the serial
            // MMA -> epilogue -> MMA chain (TMEM holds one tile: the next tile's MMAs wait for this epilogue)
            // (warp-uniform: tcgen05.ld is .aligned, every lane of the warp must take the same path)
            const bool pre_ok = OK3 && (p.nt == 32 || p.nt == 64) && m0 + lg * 32 + 31 < p.B && n0 + p.nt <= p.N &&
                                (p.ldout & 1) == 0 && ((((uintptr_t)p.out) & 15) == 0);
            // staged epilogue: the old values of this warp's 32 rows x 16-column halves are requested as coalesced 128-byte row
            // segments (lane -> row lane / 8 + 4 it, 16-byte chunk lane % 8) while the tile's MMAs are still running
            const bool stg_ok = STG && OK3 && pre_ok;
            const int ncw_s = p.nt >> 1, r4 = lane >> 3, ch = lane & 7;
            // (first half: cp.async straight into the staging tile, no registers; second half: registers until the first is done)
            double2 ldh1[8];
            if (stg_ok && !p.overwrite) {
                const double* gb = (const double*)p.out + (long)(m0 + lg * 32) * p.ldout + n0 + half * ncw_s;
                const uint32_t stg_s = smem_u32(smem + (size_t)p.epi_stage) + (uint32_t)((warp - 2) * (32 * EPI_STG_LD * 8));
#pragma unroll
                for (int it = 0; it < 8; ++it)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stg_s + (uint32_t)(((r4 + 4 * it) * EPI_STG_LD + 2 * ch) * 8)),
                                 "l"(gb + (long)(r4 + 4 * it) * p.ldout + 2 * ch) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (ncw_s > 16) {
#pragma unroll
                    for (int it = 0; it < 8; ++it)
                        ldh1[it] = __ldcs(reinterpret_cast<const double2*>(gb + (long)(r4 + 4 * it) * p.ldout + 16 + 2 * ch));
                }
            }
            // per warp: 16 columns of a 32-column tile, 32 columns of a 64-column tile
            double2 pre[16];
            if (STG) {
            } else if (OK3 && pre_ok && p.overwrite) {
#pragma unroll
                for (int i = 0; i < 16; ++i) pre[i] = make_double2(0.0, 0.0);
            } else if (OK3 && pre_ok) {
                const int ncw = p.nt >> 1;
                const double2* src = reinterpret_cast<const double2*>((const double*)p.out + (long)row * p.ldout + n0 + half * ncw);
#pragma unroll
                for (int i = 0; i < 8; ++i) pre[i] = __ldcs(src + i);
                if (p.nt == 64) {
#pragma unroll
                    for (int i = 8; i < 16; ++i) pre[i] = __ldcs(src + i);
                }
            }
            const long long te0_ = p.tim ? clock64() : 0;
            mbar_wait(&tmem_full[buf], (uint32_t)(use & 1));
            const long long te1_ = p.tim ? clock64() : 0;
            te_wait += te1_ - te0_;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (stg_ok) {
                double* stg = reinterpret_cast<double*>(smem + (size_t)p.epi_stage) + (warp - 2) * (32 * EPI_STG_LD);
                double* gb = (double*)p.out + (long)(m0 + lg * 32) * p.ldout + n0 + half * ncw_s;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (16 * h < ncw_s) {  // warp-uniform
                        double2 v[8];
                        if (!p.overwrite) {
                            if (h == 0) {
                                asm volatile("cp.async.wait_group 0;" ::: "memory");
                            } else {
#pragma unroll
                                for (int it = 0; it < 8; ++it)
                                    *reinterpret_cast<double2*>(stg + (r4 + 4 * it) * EPI_STG_LD + 2 * ch) = ldh1[it];
                            }
                            __syncwarp();
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const double2*>(stg + lane * EPI_STG_LD + 2 * i);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = make_double2(0.0, 0.0);
                        }
                        const int cb = half * ncw_s + 16 * h;
                        if (ND <= 6) epi_cols16<6>(lane_addr, cb, v, p.scale + n0 + cb, ND, p.nt, p.scale_mul);
                        else epi_cols16<8>(lane_addr, cb, v, p.scale + n0 + cb, ND, p.nt, p.scale_mul);  // ND <= 8 (launcher)
#pragma unroll
                        for (int i = 0; i < 8; ++i) *reinterpret_cast<double2*>(stg + lane * EPI_STG_LD + 2 * i) = v[i];
                        __syncwarp();
#pragma unroll
                        for (int it = 0; it < 8; ++it)
                            __stcs(reinterpret_cast<double2*>(gb + (long)(r4 + 4 * it) * p.ldout + 16 * h + 2 * ch),
                                   *reinterpret_cast<const double2*>(stg + (r4 + 4 * it) * EPI_STG_LD + 2 * ch));
                        __syncwarp();  // the next half (or tile) overwrites the staged rows
                    }
                }
            } else if (!STG && OK3 && pre_ok && p.nt == 64) {
                double* orow = (double*)p.out + (long)row * p.ldout + n0;
                if (ND == 7) epi_update16<7, 32>(lane_addr, half * 32, pre, orow, p.scale + n0, 7, 64, p.scale_mul);
                else if (ND == 8) epi_update16<8, 32>(lane_addr, half * 32, pre, orow, p.scale + n0, 8, 64, p.scale_mul);
                else epi_update16<8, 32>(lane_addr, half * 32, pre, orow, p.scale + n0, ND, 64, p.scale_mul);  // ND <= 8 at nt = 64
            } else if (!STG && OK3 && pre_ok) {
                double* orow = (double*)p.out + (long)row * p.ldout + n0;
                if (ND == 11) epi_update16<11>(lane_addr, half * 16, pre, orow, p.scale + n0, 11, 32, p.scale_mul);
                else if (ND == 6) epi_update16<6>(lane_addr, half * 16, pre, orow, p.scale + n0, 6, 32, p.scale_mul);
                else if (ND <= 8) epi_update16<8>(lane_addr, half * 16, pre, orow, p.scale + n0, ND, 32, p.scale_mul);
                else epi_update16<16>(lane_addr, half * 16, pre, orow, p.scale + n0, ND, 32, p.scale_mul);
            } else if (OK3) {
                // ragged / wider tiles
                double* orow = (double*)p.out + (long)row * p.ldout;
                if (ND <= 3) epi_update_any<3, 8>(lane_addr, orow, p.scale, n0, p.nt, p.N, row < p.B, ND, cbeg, cend, p.scale_mul, p.overwrite != 0);
                else if (ND <= 6) epi_update_any<6, 8>(lane_addr, orow, p.scale, n0, p.nt, p.N, row < p.B, ND, cbeg, cend, p.scale_mul, p.overwrite != 0);
                else if (STG) epi_update_any<8, 4>(lane_addr, orow, p.scale, n0, p.nt, p.N, row < p.B, ND, cbeg, cend, p.scale_mul, p.overwrite != 0);
                else epi_update_any<16, 4>(lane_addr, orow, p.scale, n0, p.nt, p.N, row < p.B, ND, cbeg, cend, p.scale_mul, p.overwrite != 0);
            } else if (!OK3) {
                for (int c0 = cbeg; c0 < cend; c0 += 16) {
                    if (ND <= 4) {
                        // |V| < 2^56 (+ base < 2^62): 64-bit recombination, Barrett reduction
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
                        // int32 store of a whole 16-column group: four 16-byte stores per row instead of sixteen 4-byte ones
                        // (a thread owns a ROW: every store instruction of the warp touches 32 different lines, so the
                        // instruction count is what the LSU sees)
                        if (row < p.B && p.out_kind == 1 && n0 + c0 + 16 <= p.N && (p.ldout & 3) == 0 &&
                            ((((uintptr_t)p.out) & 15) == 0)) {
                            int o[16];
#pragma unroll
                            for (int c = 0; c < 16; ++c) {
                                long long val = p.sign < 0 ? -v[c] : v[c];
                                if (val > 2147483647 || val < -2147483647) {
                                    if (p.flag) atomicOr(p.flag, 4);
                                    val = 0;
                                }
                                o[c] = (int)val;
                            }
                            int4* dst = reinterpret_cast<int4*>((int32_t*)p.out + (long)row * p.ldout + n0 + c0);
#pragma unroll
                            for (int c = 0; c < 4; ++c) dst[c] = make_int4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                            continue;
                        }
                        if (row < p.B) {
#pragma unroll
                            for (int c = 0; c < 16; ++c) {
                                const int n = n0 + c0 + c;
                                if (n >= p.N) continue;
                                long long val = p.sign < 0 ? -v[c] : v[c];
                                if (p.out_kind == 0) {
                                    if (p.base) val += p.base[(long)row * p.ldbase + n];
                                    long long r;
                                    if (!p.q) r = val;
                                    else if ((p.q & (p.q - 1)) == 0) r = (long long)((unsigned long long)val & (p.q - 1));
                                    else r = (long long)mod_i64_barrett(val, p.q, p.qmagic);
                                    ((int64_t*)p.out)[(long)row * p.ldout + n] = r;
                                } else if (p.out_kind == 1) {
                                    if (val > 2147483647 || val < -2147483647) {
                                        if (p.flag) atomicOr(p.flag, 4);
                                        val = 0;
                                    }
                                    ((int32_t*)p.out)[(long)row * p.ldout + n] = (int32_t)val;
                                } else {
                                    ((double*)p.out)[(long)row * p.ldout + n] += (double)val;
                                }
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
                            __int128 val = p.sign < 0 ? -v[c] : v[c];
                            if (p.out_kind == 0) {
                                if (p.base) val += (__int128)p.base[(long)row * p.ldbase + n];
                                long long r;
                                if (p.q) {
                                    if ((p.q & (p.q - 1)) == 0) r = (long long)((unsigned long long)val & (p.q - 1));
                                    else r = (long long)mod_i128(val, p.q);
                                } else {
                                    r = (long long)val;
                                }
                                ((int64_t*)p.out)[(long)row * p.ldout + n] = r;
                            } else if (p.out_kind == 1) {
                                if (val > 2147483647 || val < -2147483647) {
                                    if (p.flag) atomicOr(p.flag, 4);
                                    val = 0;
                                }
                                ((int32_t*)p.out)[(long)row * p.ldout + n] = (int32_t)(long long)val;
                            } else {
                                ((double*)p.out)[(long)row * p.ldout + n] += (double)(long long)val;
                            }
                        }
                    }
                }
            }
            if (p.tim) te_body += clock64() - te1_;
            // hand TMEM back: re-zero the accumulators for the next tile's accumulate-only MMAs
            if (tile + NB * walk_step < total_tiles) {  // this buffer is used again
                // (each warp zeroes exactly the columns it has just read: the other warp of the lane group may still be
                // reading its own)
                for (int d = 0; d < ND; ++d)
                    for (int c = cbeg; c < cend; c += 16) tmem_st16_zero(lane_addr + (uint32_t)(d * p.nt + c));
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            }
        }
        if (warp == 2 && lane == 0 && p.tim) { atomicAdd(p.tim + 2, (unsigned long long)te_wait); atomicAdd(p.tim + 3, (unsigned long long)te_body); }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
    // no CTA of a pair may exit while the peer can still multicast into its shared memory or arrive on its barriers
    if (PAIR) cluster_sync_all();
}

// Ceiling of the int8 tensor pipe itself: every CTA loads ONE 128 x 128-byte x tile and ONE 256 x 128-byte w tile and then
// issues the MMAs of that k block `iters` times on the resident operands (no operand traffic at all: what is measured is
// tcgen05.mma kind::i8 at M = 128, N = 256, i.e. 100 % "IMMA pipe active").  The accumulator wraps around; nothing is read back.
__global__ void __launch_bounds__(128, 1)
i8_pipe_probe_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, int iters) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int x_tile = TILE_M * BLOCK_K, w_tile = 256 * BLOCK_K;
    uint64_t* bars = (uint64_t*)(smem + x_tile + w_tile);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 1) {
        if (lane == 0) {
            mbar_expect_tx(&bars[0], (uint32_t)(x_tile + w_tile));
            tma_load_3d(&map_x, smem, &bars[0], 0, (int)(blockIdx.x % 8) * TILE_M, 0);
            tma_load_3d(&map_w, smem + x_tile, &bars[0], 0, 0, 0);
        }
        mbar_wait(&bars[0], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
            const uint32_t idesc = (2u << 4) | (1u << 7) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
            const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + x_tile));
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int kk = 0; kk < BLOCK_K / 32; ++kk) mma_i8(tmem_base, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, 1u);
            }
            mma_commit(&bars[1]);
        }
        __syncwarp();
        mbar_wait(&bars[1], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace

int qf_i8_tile_n(int LX, int LW, int N, int d_lo) {
    const int ND = LX + LW - 1 - d_lo;
    int nt = (512 / ND) / 16 * 16;
    if (nt > 256) nt = 256;
    int need = (N + 15) / 16 * 16;
    if (nt > need) nt = need;
    return nt;
}

cudaError_t qf_launch_gemm_i8(const I8GemmArgs& a, cudaStream_t stream) {
    if (a.B <= 0 || a.N <= 0) return cudaSuccess;
    if (a.LX < 1 || a.LW < 1 || a.LX + a.LW - 1 > 16 || a.K < 1) return cudaErrorInvalidValue;
    if (a.d_lo < 0 || (a.d_lo > 0 && a.out_kind != 3) || a.d_lo > a.LX + a.LW - 2) return cudaErrorInvalidValue;
    if ((a.ldx & 15) || (a.ldw & 15) || (a.x_plane & 15) || (a.w_plane & 15) || (((uintptr_t)a.x) & 15) ||
        (((uintptr_t)a.w) & 15))
        return cudaErrorMisalignedAddress;
    int nt = qf_i8_tile_n(a.LX, a.LW, a.N, a.d_lo);
    if (nt < 16) return cudaErrorInvalidValue;
    const int ND = a.LX + a.LW - 1 - a.d_lo;
    // the fast read-modify-write epilogue of the fixed-point update exists for 32- and 64-column tiles
    // (wider tiles, ND <= 4, keep their width and take the generic epilogue)
    if (a.out_kind == 3 && nt > 128) nt = 128;
    // block-triangular key matrix: 64-column tiles follow the triangle more closely, take the staged epilogue and leave room
    // for two accumulator sets (measured +0.6 % on the C2 step against 128-column tiles)
    if (a.out_kind == 3 && a.tri_mode != 0 && nt > 64) nt = 64;
    if (a.out_kind == 3 && nt > 64 && nt < 128) nt = 64;
    if (a.out_kind == 3 && nt > 32 && nt < 64) nt = 32;
    {
        // Experiment switches (profiles/README_r2.md): 32-column tiles leave room for two accumulator sets in TMEM (tile
        // i+1's MMAs overlap tile i's read-modify-write epilogue) -- measured SLOWER than one 64-column set for every
        // contraction length of the C2 step (305 k/s against 311 k/s), so the wide tile is the default.
        static const int nt32_max_kb = getenv("QF_I8_NT32_MAX_KB") ? atoi(getenv("QF_I8_NT32_MAX_KB")) : 0;
        static const int force_nt = getenv("QF_I8_UPDATE_NT") ? atoi(getenv("QF_I8_UPDATE_NT")) : 0;
        if (a.out_kind == 3 && nt == 64 && (a.K + BLOCK_K - 1) / BLOCK_K <= nt32_max_kb && 2 * ND * 32 <= 512) nt = 32;
        if (a.out_kind == 3 && (force_nt == 32 || force_nt == 64) && nt >= force_nt) nt = force_nt;
    }
    // staged (coalesced) read-modify-write epilogue of the fixed-point update: needs EPI_STG_BYTES behind two pipeline stages
    // (QF_I8_EPI_STAGE=0: the register epilogue, for A/B measurements)
    static const bool stage_off = getenv("QF_I8_EPI_STAGE") && getenv("QF_I8_EPI_STAGE")[0] == '0';
    bool epi_stage = false;
    {   // two pipeline stages of LX x-planes + LW w-planes must fit in shared memory
        // test-only switch, read once per process: 64-byte K blocks (SWIZZLE_64B)
        static const int bk_env = (getenv("QF_I8_BLOCK_K") && atoi(getenv("QF_I8_BLOCK_K")) == 64) ? 64 : BLOCK_K;
        const int bk0 = bk_env;
        const int budget0 = 227 * 1024 - 1024 - 256;
        while (nt > 16 && 2 * (a.LX * TILE_M * bk0 + a.LW * nt * bk0) > budget0) nt = (a.out_kind == 3 && nt == 64) ? 32 : nt - 16;
        epi_stage = !stage_off && a.out_kind == 3 && (nt == 32 || nt == 64) && ND <= 8 &&
                    2 * (a.LX * TILE_M * bk0 + a.LW * nt * bk0) <= budget0 - EPI_STG_BYTES;
    }
    I8Params p{};
    p.d_lo = a.d_lo;
    p.overwrite = a.overwrite;
    p.scale_mul = a.scale_mul > 0.0 ? a.scale_mul : 1.0;
    for (int i = 0; i < a.d_lo; ++i) p.scale_mul *= 256.0;
    p.acc_bufs = (2 * ND * nt <= 512) ? 2 : 1;
    p.gate = a.gate; p.gate_lo = a.gate_lo; p.gate_hi = a.gate_hi;
    p.tri_mode = a.tri_mode; p.tri_slack = a.tri_slack;
    p.B = a.B; p.N = a.N; p.K = a.K; p.LX = a.LX; p.LW = a.LW; p.nt = nt; p.w_signed = a.w_signed;
    p.out_kind = a.out_kind; p.sign = a.sign; p.q = a.q; p.qmagic = qf_barrett_magic(a.q); p.base = a.base; p.ldbase = a.ldbase; p.out = a.out;
    p.ldout = a.ldout; p.flag = a.flag; p.scale = a.scale;
    p.m_tiles = (a.B + TILE_M - 1) / TILE_M;
    p.n_tiles = (a.N + nt - 1) / nt;
    {
        // one wave (148 CTAs) should touch group_m x-tiles and 148/group_m w-tiles with minimal bytes:
        // group_m = sqrt(148 * w_tile / x_tile)
        const double xt = (double)a.LX * TILE_M, wt = (double)a.LW * nt;
        int gm = (int)(sqrt(148.0 * wt / xt) + 0.5);
        if (gm < 1) gm = 1;
        if (gm > p.m_tiles) gm = p.m_tiles;
        p.group_m = gm;
    }
    p.x_nz = a.x_nz; p.nz_m_tiles = a.nz_m_tiles; p.nz_kb_total = a.nz_kb_total; p.nz_kb_off = a.nz_kb_off;
    p.nz_m_off = a.nz_m_off;
    // K block: 128-byte rows (SWIZZLE_128B).  QF_I8_BLOCK_K=64 selects 64-byte rows (SWIZZLE_64B): twice the pipeline
    // depth in the same shared memory, but measured 30 % slower on B200 (round-1 profile notes) -- the kernel is bound by
    // L2 -> shared-memory operand traffic, not by pipeline depth, and the 64-byte layout feeds the tensor core worse.
    const int budget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - (epi_stage ? EPI_STG_BYTES : 0);  // barriers + tmem slot
    static const int bk_env2 = (getenv("QF_I8_BLOCK_K") && atoi(getenv("QF_I8_BLOCK_K")) == 64) ? 64 : BLOCK_K;
    const int bk = bk_env2;
    p.bk = bk;
    const int stage_bytes = a.LX * TILE_M * bk + a.LW * nt * bk;
    int stages = budget / stage_bytes;
    if (stages < 2) return cudaErrorInvalidValue;
    if (stages > 8) stages = 8;
    p.stages = stages;
    const int smem = stages * stage_bytes + 1024 + 256 + (epi_stage ? EPI_STG_BYTES : 0);
    p.epi_stage = epi_stage ? stages * stage_bytes + 256 : 0;
    CUtensorMap mx, mw, mxh;
    if (!make_map(&mx, a.x, a.K, a.B, a.LX, a.ldx, a.x_plane, TILE_M, bk)) return cudaErrorInvalidValue;
    if (!make_map(&mw, a.w, a.K, a.N, a.LW, a.ldw, a.w_plane, nt, bk)) return cudaErrorInvalidValue;
    if (!make_map(&mxh, a.x, a.K, a.B, a.LX, a.ldx, a.x_plane, TILE_M / 2, bk)) return cudaErrorInvalidValue;
    const int dslot = qf_device_slot();
    static int configured_dev[QF_MAX_DEVICES] = {};
    int& configured = configured_dev[dslot];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_i8_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_i8_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_i8_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_i8_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_i8_kernel<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_i8_kernel<1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    p.mma_units = a.mma_units;
    p.tim = a.tim;
    static int sm_count_dev[QF_MAX_DEVICES] = {};
    int& sm_count = sm_count_dev[dslot];
    if (!sm_count) {
        if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dslot) != cudaSuccess || sm_count <= 0) sm_count = 148;
    }
    const int total_tiles = p.m_tiles * p.n_tiles;
    // CTA pairs (x tile multicast) when the coordinate tiles pair up and there is enough work for every pair;
    // QF_I8_PAIR=0 keeps single CTAs (experiments / tests)
    static const bool pair_off = getenv("QF_I8_PAIR") && getenv("QF_I8_PAIR")[0] == '0';
    const bool pair = !pair_off && (p.n_tiles % 2 == 0) && total_tiles >= 2 && bk == BLOCK_K;
    if (pair) {
        // group_m was chosen for single tiles; pairs halve the number of w-tile columns a wave touches: keep it
        const int pairs = total_tiles / 2;
        int ctas = 2 * (pairs < sm_count / 2 ? pairs : sm_count / 2);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)ctas);
        cfg.blockDim = dim3(I8_THREADS);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (a.out_kind == 3 && epi_stage) return cudaLaunchKernelEx(&cfg, gemm_i8_kernel<1, true, true>, mx, mw, mxh, p);
        if (a.out_kind == 3) return cudaLaunchKernelEx(&cfg, gemm_i8_kernel<1, true>, mx, mw, mxh, p);
        return cudaLaunchKernelEx(&cfg, gemm_i8_kernel<0, true>, mx, mw, mxh, p);
    }
    dim3 grid((unsigned)(total_tiles < sm_count ? total_tiles : sm_count));  // persistent: one CTA per SM
    if (a.out_kind == 3 && epi_stage) gemm_i8_kernel<1, false, true><<<grid, I8_THREADS, smem, stream>>>(mx, mw, mxh, p);
    else if (a.out_kind == 3) gemm_i8_kernel<1, false><<<grid, I8_THREADS, smem, stream>>>(mx, mw, mxh, p);
    else gemm_i8_kernel<0, false><<<grid, I8_THREADS, smem, stream>>>(mx, mw, mxh, p);
    return cudaGetLastError();
}

// x: >= 1024 rows x 128 bytes, w: 256 rows x 128 bytes of int8 (row stride 128).  Returns the int8 operations issued.
cudaError_t qf_launch_i8_pipe_probe(const int8_t* x, const uint8_t* w, int iters, int grid, double* ops_out, cudaStream_t stream) {
    CUtensorMap mx, mw;
    if (!make_map(&mx, x, BLOCK_K, 8 * TILE_M, 1, BLOCK_K, (long)8 * TILE_M * BLOCK_K, TILE_M)) return cudaErrorInvalidValue;
    if (!make_map(&mw, w, BLOCK_K, 256, 1, BLOCK_K, (long)256 * BLOCK_K, 256)) return cudaErrorInvalidValue;
    const int smem = TILE_M * BLOCK_K + 256 * BLOCK_K + 1024 + 64;
    cudaError_t e = cudaFuncSetAttribute(i8_pipe_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    i8_pipe_probe_kernel<<<grid, 128, smem, stream>>>(mx, mw, iters);
    if (ops_out) *ops_out = 2.0 * grid * (double)iters * TILE_M * 256.0 * BLOCK_K;
    return cudaGetLastError();
}
