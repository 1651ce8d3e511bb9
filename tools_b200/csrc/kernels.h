// Internal launcher declarations (device pointers everywhere).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

// ---- gemm_f64.cu ----------------------------------------------------------
cudaError_t qf_launch_gemm_f64(const double* X, long ldx, const double* W, long ldw, double* C, long ldc,
                               int B, int N, int K, double alpha, double beta, int tri_lower,
                               cudaStream_t stream);

// ---- elementwise.cu -------------------------------------------------------
// out[b][j] = (double) in[b][j]; optionally norm2[b] = sum_j in[b][j]^2 (exact, u64 saturating)
cudaError_t qf_launch_i32_to_f64(const int32_t* in, long ldin, double* out, long ldout, int B, int M,
                                 unsigned long long* norm2, cudaStream_t stream);
cudaError_t qf_launch_i64_to_f64(const int64_t* in, long ldin, double* out, long ldout, int B, int M,
                                 double scale, cudaStream_t stream);
// out[b][j] = (int32) in[b][j]; sets *flag if a value does not fit / is not an integer
cudaError_t qf_launch_f64_to_i32(const double* in, long ldin, int32_t* out, long ldout, int B, int M,
                                 int* flag, cudaStream_t stream);
// in_domain[b] = norm2[b] <= bound
cudaError_t qf_launch_domain_flags(const unsigned long long* norm2, unsigned long long bound, uint8_t* flags,
                                   int B, cudaStream_t stream);

struct CombineArgs {
    const double* acc[4];   // exact-integer fp64 partial products, B x N each, leading dim ldacc
    int shift[4];           // value = sum_c acc_c * 2^shift_c
    int nacc;
    int acc_sign;           // +1 or -1
    long ldacc;
    const int64_t* base;    // optional B x N (ld ldbase): out = base + acc_sign * value
    long ldbase;
    unsigned long long q;   // 0: plain integer result, else reduce into [0,q)
};
cudaError_t qf_launch_combine_i64(const CombineArgs& a, int64_t* out, long ldout, int B, int N, cudaStream_t stream);
cudaError_t qf_launch_combine_f64(const CombineArgs& a, double* out, long ldout, int B, int N, cudaStream_t stream);
// int32 output; sets *flag when a value does not fit
cudaError_t qf_launch_combine_i32(const CombineArgs& a, int32_t* out, long ldout, int B, int N, int* flag,
                                  cudaStream_t stream);
// out[b][j] = (int32)(P[b][j] + (j >= split ? Zb[b][j-split] : 0))   (e = p + [R;I] z, lower part)
cudaError_t qf_launch_finalize_pert(const double* P, long ldp, const double* Zb, long ldz, int32_t* out, long ldo,
                                    int B, int M, int split, int* flag, cudaStream_t stream);

// split an exact-integer fp64 matrix into balanced chunks of `bits` bits: in = sum_c out_c * 2^(c*bits)
cudaError_t qf_launch_split_chunks(const double* in, long ldin, double* const* out, int nchunks, int bits,
                                   long ldout, int B, int M, cudaStream_t stream);
// dst[b][cols[j]] = src[b][j]  (dst pre-zeroed by the caller)
cudaError_t qf_launch_scatter_cols_f64(const double* src, long ldsrc, const int* cols, int ncols, double* dst,
                                       long lddst, int B, double scale, cudaStream_t stream);

// ---- samplers (elementwise.cu) -------------------------------------------
cudaError_t qf_launch_normal_fill(double* out, long ld, int B, int M, uint64_t seed, uint64_t first_target,
                                  uint32_t tag, cudaStream_t stream);
// out[b][j] <- D_{Z, s, center[b][j]} (center == nullptr: centred at 0); flag (optional): sampler bail-out bit
cudaError_t qf_launch_dgauss(const double* center, long ldc, double* out_f64, long ldo, int32_t* out_i32,
                             long ldoi, int B, int M, double s, uint64_t seed, uint64_t first_target,
                             uint32_t tag, int* flag, cudaStream_t stream);
// structured e = S z for the G-trapdoor short basis S = [[R S', I + R W],[S', W]] (api.cu detect_gpv_structure)
cudaError_t qf_launch_sprime_apply(const double* Z, long ldz, double* I2, long ldi, int B, int nk, int k, const double* sk,
                                   int reversed, cudaStream_t stream);
cudaError_t qf_launch_gpv_struct_finalize(int32_t* e, long lde, const double* Z2, long ldz, const double* I2, long ldi, int B,
                                          int mb, int nk, int* flag, cudaStream_t stream, const int8_t* g3 = nullptr,
                                          long ldg = 0);
// two-phase form: e_bot = g3 + S' z1 written as int32 into e[b][mb + row] and as L balanced digit planes (one pass over z1)
cudaError_t qf_launch_gpv_ebot(const double* Z, long ldz, const int8_t* g3, long ldg, int32_t* e, long lde, int mb,
                               int8_t* planes, long plane_stride, long ldk, int L, int B, int nk, int k, const double* sk,
                               int reversed, int* flag, cudaStream_t stream);
// two-phase nearest plane (api.cu samp_p_np2_chunk): base-b digits of the syndromes h (B x n, in [0,q)) in gadget order,
// g3[b][blk*k + t] = digit t of h[b][blk], written as ONE s8 digit plane (B x ldk, columns [0, n*k)) and marked in
// plane 0 of the zero-tile map (k blocks [0, ceil(n*k/128)) of every target tile)
cudaError_t qf_launch_gadget_digits(const int64_t* h, long ldh, int8_t* plane, long ldk, int B, int n, int k, unsigned base,
                                    uint8_t* nz, int nz_kb_total, cudaStream_t stream, double* I2 = nullptr, long ldi = 0);
// M'[i][bc k + t] = sum_t' U[i][col(bc, t')] Skinv[t'][t] over the upper unitriangular part of U (col = the basis column
// of gadget coordinate bc k + t', reversed when `rev`): the map from a gadget-coordinate vector g (centre -[R;I] g) to
// GSO coordinates of the first nk basis vectors (api.cu samp_p_np2_chunk)
cudaError_t qf_launch_gadget_to_gso(const double* U, long ldu, int nk, int k, int rev, const double* skinv, double* out,
                                    long ldo, cudaStream_t stream);
// structured perturbation (api.cu setup_structured_sigma2): X2[b][mb+j] = sqrt_beta * G[b][mb+j] and the balanced
// base-256 digits of rint(that * fscale) into L planes of B x ldk bytes
cudaError_t qf_launch_pert_xb(const double* G, long ldg, double* X2, long ldx, int8_t* planes, long plane_stride, long ldk,
                              int B, int mb, int nk, double sqrt_beta, double fscale, int L, int* flag, cudaStream_t stream);
// N(0,1) draws of the perturbation (same Philox stream as qf_launch_normal_fill) written as fixed-point digits:
// coordinates j < split: LG planes of rint(g * gscale) at gplanes[b * ldkg + j];
// coordinates j >= split (structured square root only): x_b = sqrt_beta * g into X2[b][j] and LB planes of
// rint(x_b * bscale) at bplanes[b * ldkb + (j - split)].
cudaError_t qf_launch_pert_normal_digits(int B, int M, int split, uint64_t seed, uint64_t first_target, int8_t* gplanes,
                                         long gplane_stride, long ldkg, int LG, double gscale, double* X2, long ldx,
                                         int8_t* bplanes, long bplane_stride, long ldkb, int LB, double sqrt_beta, double bscale,
                                         int* flag, cudaStream_t stream);
cudaError_t qf_launch_uniform_modq(int64_t* out, long count, unsigned long long q, uint64_t seed,
                                   uint64_t first_index, cudaStream_t stream);
cudaError_t qf_launch_ternary(int8_t* out, long count, uint64_t seed, cudaStream_t stream);

// ---- lattice.cu -------------------------------------------------------------
// Gadget-lattice preimage (mp_perturbation.rs:173-191) for every (target, row):
// V: B x n residues, Z: B x (n*k) exact-integer doubles.
// sk: k x k block basis (row-major), gso: k x k GSO (columns are b~_i), both in device memory.
cudaError_t qf_launch_gadget_sample(const int64_t* V, long ldv, double* Z, long ldz, int B, int n, int k,
                                    int base, unsigned long long q, const double* sk, const double* gso,
                                    double s_g, uint64_t seed, uint64_t first_target, int* flag, cudaStream_t stream);
// One diagonal block of the randomized nearest-plane recursion in GSO coordinates:
// for i = j0+nb-1 .. j0:  c' = T[b][i] - sum_{j>i in block} U[i][j] z_j ;  z_i <- D_{Z, s/||b~_i||, c'}
// writes Z[b][i].  dg: per-coordinate sampler parameters (length >= j0+nb).
// *flag is set when |z| >= zlimit (exact-integer range check).
// prop: this block's pre-generated proposals (np_propose), coordinate-major: prop[(i - j0) * ldprop + b]
// dig (optional): the kernel also writes the balanced base-256 digit planes of its block of z (L planes of B x ldk bytes,
// planes already offset to column 0), marks the zero-tile map and raises the digit-count gates, like qf_launch_split_f64_limbs
struct NpDigitOut {
    int8_t* planes; long plane_stride, ldk; int L;
    uint8_t* nz; int nz_m_tiles, nz_kb_total;
    int* gate[4];
};
// nb <= 64: one diagonal block [j0, j0 + nb).  nb > 64 (j0 a multiple of 64, up_lo = j0): the whole block range, its
// 64-wide diagonal blocks from the top down, each followed by the rank-64 update of the columns of the range below it,
//   T[b][j] -= sum_i Z[b][j0' + i] U[j][j0' + i]   for this CTA's own targets (no cross-CTA dependence: one launch).
// prop: proposals of coordinate prop0 onwards (prop0 = -1: j0)
cudaError_t qf_launch_np_diag(double* T, long ldt, double* Z, long ldz, const double* U, long ldu,
                              const DGaussParams* dg, const float4* prop, long ldprop, int B, int j0, int nb, int dim,
                              uint64_t seed, uint64_t first_target, double zlimit, int* flag, cudaStream_t stream, int up_lo = -1,
                              const NpDigitOut* dig = nullptr, int prop0 = -1, int variant = 0);
// two proposals per (target, coordinate) for coordinates [j_lo, j_lo + width): out[(i - j_lo) * ldo + b]
cudaError_t qf_launch_np_propose(float4* out, long ldo, int B, int j_lo, int width, int dim, uint64_t seed,
                                 uint64_t first_target, cudaStream_t stream);

// ---- compress.cu -----------------------------------------------------------
cudaError_t qf_launch_compress_u16(const uint16_t* in, uint16_t* out, size_t count, uint32_t q, uint32_t d,
                                   int decompress, cudaStream_t stream);
// Compress_d + ByteEncode_d / ByteDecode_d + Decompress_d, one warp per degree-256 polynomial (compress.cu)
cudaError_t qf_launch_byte_encode(const uint16_t* in, uint8_t* out, size_t npoly, uint32_t q, uint32_t d, int compress,
                                  cudaStream_t stream);
cudaError_t qf_launch_byte_decode(const uint8_t* in, uint16_t* out, size_t npoly, uint32_t q, uint32_t d, int decompress,
                                  cudaStream_t stream);
cudaError_t qf_launch_compress_i64(const int64_t* in, int64_t* out, size_t count, unsigned long long q,
                                   uint32_t d, int decompress, cudaStream_t stream);

// ---- encodings.cu : common_encodings.rs:49-153, batched (digits as bytes, base <= 256; bit-packed base 2) ----------
cudaError_t qf_launch_encode_digits(const uint8_t* digits, void* coeffs, size_t count, unsigned long long q, unsigned base,
                                    int coeff_bytes, cudaStream_t stream);
cudaError_t qf_launch_decode_digits(const void* coeffs, uint8_t* digits, size_t count, unsigned long long q, unsigned base,
                                    int coeff_bytes, cudaStream_t stream);
cudaError_t qf_launch_encode_bits_u16(const uint8_t* msg, uint16_t* coeffs, size_t nbytes, uint32_t q, cudaStream_t stream);
cudaError_t qf_launch_decode_bits_u16(const uint16_t* coeffs, uint8_t* msg, size_t nbytes, uint32_t q, cudaStream_t stream);

// ---- ring_ntt.cu -----------------------------------------------------------
// Negacyclic products over Z_q[X]/(X^n+1) through an exact NTT over the Goldilocks prime.
// a_hat: npoly x n precomputed transforms of the key polynomials (centred lift of a mod q).
// tw: device table of 2n+1 words (psi powers bit-reversed, inverse powers, 1/n) from qf_ring_make_tables.
void qf_ring_make_tables(int n, uint64_t* host_out);
cudaError_t qf_launch_ring_prepare(const int64_t* a, uint64_t* a_hat, int npoly, int n, unsigned long long q,
                                   const uint64_t* tw, cudaStream_t stream);
// out[b] = sum_j a_j * sigma[b][j] mod (X^n+1, q);  sigma: B x npoly x n (int32), out: B x n
cudaError_t qf_launch_ring_f_a(const int32_t* sigma, const uint64_t* a_hat, int64_t* out, unsigned long long* norm2,
                               int B, int npoly, int n, unsigned long long q, const uint64_t* tw,
                               cudaStream_t stream);
// Schoolbook fallback for degrees that are not a power of two (or < 64): a is the raw key, npoly x n in [0,q).
cudaError_t qf_launch_ring_f_a_schoolbook(const int32_t* sigma, const int64_t* a, int64_t* out,
                                          unsigned long long* norm2, int B, int npoly, int n, unsigned long long q,
                                          cudaStream_t stream);

// ---- setup.cu (per-key, not per-target) --------------------------------------
cudaError_t qf_launch_colnorm2(const double* G, long ld, int rows, int cols, double* d, cudaStream_t stream);
cudaError_t qf_launch_transpose_scale(const double* in, long ldin, double* out, long ldout, int rows, int cols,
                                      const double* scale, cudaStream_t stream);
cudaError_t qf_launch_gather_cols(const double* in, long ldin, const int* cols, int ncols, double* out, long ldout,
                                  int rows, cudaStream_t stream);
cudaError_t qf_launch_make_dg(const double* d, int count, double s, DGaussParams* dg, cudaStream_t stream);
// blocked Cholesky building blocks (compute_sqrt_sigma_2, mp_perturbation.rs:111-139)
cudaError_t qf_launch_potrf_diag(double* A, long ld, int nb, double* Linv, int* info, cudaStream_t stream);
cudaError_t qf_launch_fixed_rows_prepare(const double* L, long ld, int rows, int cols, int Ldig, double mult, double* scale,
                                         int8_t* planes, long plane_stride, long ldk, cudaStream_t stream);
cudaError_t qf_launch_gso_rdiag(double* rdiag, const double* L, long ldl, int nb, int first, cudaStream_t stream);
cudaError_t qf_launch_scale_cols(const double* in, long ldin, double* out, long ldout, long rows, long cols,
                                 const double* colscale, cudaStream_t stream);
cudaError_t qf_launch_tril(double* A, long ld, long n, cudaStream_t stream);
cudaError_t qf_launch_copy_block(const double* in, long ldin, double* out, long ldout, long rows, int cols, cudaStream_t stream);
cudaError_t qf_launch_sigma2_assemble(double* C, long ldc, long n, int full, const double* Gin, long ldg, const double* R,
                                      long ldr, long mb, const double* Sigma, long lds, double diag, double gcoef, double coef,
                                      cudaStream_t stream);

// ---- gemm_i8.cu : exact integer contraction on tcgen05 (kind::i8) --------------------------
struct I8GemmArgs {
    const int8_t* x;    // LX planes of B x K balanced s8 limbs, row stride ldx, plane stride x_plane (bytes)
    long ldx, x_plane;
    const void* w;      // LW planes of N x K limbs (u8 residue limbs or balanced s8 limbs)
    long ldw, w_plane;
    int LX, LW, w_signed;
    int B, N, K;
    int out_kind;       // 0: int64 (base, sign, optional mod q)  1: int32 store  2: fp64 accumulate
    int sign;
    unsigned long long q;
    const int64_t* base;
    long ldbase;
    void* out;
    long ldout;
    int* flag;
    const double* scale;   // out_kind 3: out[b][n] -= V * scale[n]  (fp64)
    // optional zero-tile map of x: nz[(j * nz_m_tiles + nz_m_off + m_tile) * nz_kb_total + nz_kb_off + kb] != 0
    // iff digit plane j has a non-zero byte in that 128-target x 128-column tile (K offset must be 128-aligned)
    const uint8_t* x_nz;
    int nz_m_tiles, nz_kb_total, nz_kb_off, nz_m_off;
    // optional device counter: += int8 operations the tensor pipe actually executed (zero digit tiles are skipped)
    unsigned long long* mma_units;
    unsigned long long* tim;  // optional: 4 phase cycle counters (diagnostics)
    // out_kind 3 only: digit sums D_d with d < d_lo are not computed (their pairs lie below the error budget)
    int d_lo;
    int overwrite;  // out_kind 3 only: out = -V * scale[n] (store) instead of out -= V * scale[n]
    // optional conditional launch: the grid runs only if gate_lo <= *gate <= gate_hi (device int)
    const int* gate;
    int gate_lo, gate_hi;
    // out_kind 3 only: extra factor on scale[n] (0 = 1): the caller passes the TOP planes of a fixed-point matrix
    // (w advanced by `dropped` planes, LW reduced) and 256^dropped here
    double scale_mul;
    // structured key matrix: rows [n0, n0 + nt) are zero in the columns  1: k >= n0 + nt + tri_slack;
    // 2: k < K - (n0 + nt) - tri_slack;  3: k >= K - n0 + tri_slack;  4: k < n0 - tri_slack.  Those k blocks are skipped.
    int tri_mode, tri_slack;
};
int qf_i8_tile_n(int LX, int LW, int N, int d_lo = 0);
// gemm_i8_fused.cu: out = X W^t mod q with X read as int32 (digit split fused into the contraction) and the
// squared row norms of X accumulated on the way (norm2 optional)
struct FaFusedArgs {
    const int32_t* x; long ldx;   // B x K
    const void* w; long ldw, w_plane;  // LW planes of N x K limbs
    int LX, LW, w_signed;
    int LX_typical;   // digits covering the sampler's tail cut (0: unknown)
    int B, N, K;
    unsigned long long q;
    int64_t* out; long ldout;
    unsigned long long* norm2;
    unsigned long long* mma_units;
    int* retry_flag;   // optional device int: enables the optimistic (LX - 1 digits first) launch pair
    unsigned long long* tim;  // optional diagnostics (QF_TRACE): 6 cycle counters summed over CTAs, see gemm_i8_fused.cu
};
cudaError_t qf_launch_f_a_fused(const FaFusedArgs& a, cudaStream_t stream);
cudaError_t qf_launch_gemm_i8(const I8GemmArgs& a, cudaStream_t stream);
// tensor-pipe ceiling probe: MMAs on shared-memory-resident operands (gemm_i8.cu)
cudaError_t qf_launch_i8_pipe_probe(const int8_t* x, const uint8_t* w, int iters, int grid, double* ops_out, cudaStream_t stream);

// ---- limb splitting (elementwise.cu) ---------------------------------------------------------
// balanced s8 digits; *flag |= 8 when a value does not fit L digits
// nz (optional, pre-zeroed): zero-tile map [L][nz_m_tiles][nz_kb_total] over 128-target x 128-column tiles;
// col0 = global column of in[.][0] (the planes pointer is already offset by col0)
// gate0..2 (optional, need nz): device ints raised to the index of the highest non-zero digit plane written
cudaError_t qf_launch_split_f64_limbs(const double* in, long ldin, int8_t* planes, long plane_stride, long ldk, int B,
                                      int M, int L, int* flag, uint8_t* nz, int nz_m_tiles, int nz_kb_total, int col0,
                                      cudaStream_t stream, int* gate0 = nullptr, int* gate1 = nullptr, int* gate2 = nullptr,
                                      int* gate3 = nullptr);
cudaError_t qf_launch_split_i32_limbs(const int32_t* in, long ldin, int8_t* planes, long plane_stride, long ldk, int B,
                                      int M, int L, unsigned long long* norm2, uint8_t* nz, int nz_m_tiles,
                                      int nz_kb_total, cudaStream_t stream);
// e[b][cols[j]] += (int32) sol[b][j]
cudaError_t qf_launch_add_cols_i32(int32_t* e, long lde, const double* sol, long ldsol, const int* cols, int ncols,
                                   int B, cudaStream_t stream);
// fixed-point digit planes of U for the tensor-core nearest-plane updates (see setup.cu)
// split > 0: the entries U[i][j] with i < split <= j are left out (two-phase recursion: that block is never multiplied)
cudaError_t qf_launch_ozaki_prepare(const double* U, long ld, int D, int blk, int sblk, int fs_last, int ss_last, int L,
                                    double* scale, int8_t* planes, long plane_stride, long ldk, cudaStream_t stream,
                                    int split = 0);

// narrow boundary types (elementwise.cu): int16 Domain values, device-side range check of residues
// narrow: *flag |= 64 when a value does not fit int16; range check: *flag |= 32 when a residue is outside [0, q)
cudaError_t qf_launch_narrow_i32_i16(const int32_t* in, int16_t* out, size_t count, int* flag, cudaStream_t stream);
cudaError_t qf_launch_widen_i16_i32(const int16_t* in, int32_t* out, size_t count, cudaStream_t stream);
cudaError_t qf_launch_range_check_i64(const int64_t* v, size_t count, unsigned long long q, int* flag, cudaStream_t stream);

// polynomial arithmetic of the ring short basis (setup.cu): P[2][2][n], Q[k][2][n]
cudaError_t qf_launch_ring_basis_polys(const int32_t* e, const int32_t* r, const int32_t* w, const int64_t* sk, int n, int k,
                                       int64_t* P, int64_t* Q, cudaStream_t stream);

// ---- ring_small.cu : register/shuffle NTT mod q for NTT-friendly primes q < 2^16 ----------------
#ifdef __cplusplus
#include <vector>
int qf_ring_small_plan(unsigned long long q, int n, int* d_out, std::vector<uint32_t>* tables, uint32_t* np_inv);
#endif
void qf_ring_small_key(const int64_t* a, int npoly, int n, int d, unsigned long long q, const uint32_t* tables,
                       uint32_t* a_hat);
cudaError_t qf_launch_ring_small(const int32_t* sigma, const uint32_t* a_hat, int64_t* out, unsigned long long* norm2,
                                 int B, int npoly, int n, int d, unsigned long long q, uint32_t np_inv,
                                 const uint32_t* tw, cudaStream_t stream);
