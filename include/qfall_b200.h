/* qfall_b200.h -- C ABI of the B200 (sm_100a) backend for the batched PSF /
 * FIPS 203 hot path of qfall/tools.
 *
 * This is the drop-in boundary: every entry point replaces one method of the
 * reference's `PSF` trait (src/primitive/psf.rs:39-81) or of its
 * `LossyCompressionFIPS203` trait (src/compression/lossy_compression_fips203.rs:20-59)
 * for a BATCH of B targets.  A Rust shim converts qfall-math values to the
 * fixed-width buffers below (see INTEGRATION.md for the extern "C" block).
 *
 * Conventions
 *  - plain pointers and sizes only; all matrices row-major;
 *  - Range values (u, A, a_bar, tags ...) are residues in [0,q), q < 2^62, stored
 *    as int64 (the FLINT small-word form qfall-math holds them in);
 *  - Domain values (sigma, e, samp_d output) are int32 (|entry| <= s*r*sqrt(m));
 *  - a batch is "one row per target": sigma/e are B x m, u is B x n;
 *    ring values are B x (k+2) x n coefficients (coefficient embedding order,
 *    gpv_ring.rs:172-178);
 *  - every function returns a qf_status, never aborts; qf_last_error() gives text;
 *  - functions without a suffix take HOST pointers and synchronise before
 *    returning; `_dev` variants take DEVICE pointers, enqueue on the context's
 *    stream and do not synchronise;
 *  - one context per thread (or external locking): the reference types are
 *    neither Send nor Sync (gadget_parameters.rs:51).
 *  - samplers are keyed by (seed, first_index + row): a batch split across calls,
 *    chunks or GPUs yields the same preimages as one call.
 */
#ifndef QFALL_B200_H
#define QFALL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    QF_OK = 0,
    QF_ERR_INVALID = 1,       /* bad argument / shape (reference: panic or MathError) */
    QF_ERR_CUDA = 2,          /* CUDA runtime failure */
    QF_ERR_NOT_IN_DOMAIN = 3, /* f_a: some sigma fails check_domain (reference: assert!, gpv.rs:191) */
    QF_ERR_NO_KEY = 4,        /* key / trapdoor not installed */
    QF_ERR_UNSUPPORTED = 5,   /* parameter range outside this backend */
    QF_ERR_NUMERIC = 6        /* internal range check tripped (value did not fit) */
} qf_status;

typedef enum {
    QF_PSF_GPV = 0,          /* src/primitive/psf/gpv.rs */
    QF_PSF_PERTURBATION = 1, /* src/primitive/psf/mp_perturbation.rs */
    QF_PSF_GPV_RING = 2      /* src/primitive/psf/gpv_ring.rs */
} qf_psf_kind;

/* GadgetParameters / GadgetParametersRing (gadget_parameters.rs:44-52, 73-81) plus the
 * Gaussian parameters of the PSF struct (gpv.rs:54-57, mp_perturbation.rs:58-62,
 * gpv_ring.rs:63-67). */
typedef struct {
    int32_t kind;       /* qf_psf_kind */
    int64_t n;          /* security parameter / ring degree */
    int64_t k;          /* gadget length ceil(log_base q) */
    int64_t m_bar;      /* classical: columns of A_bar; ring: k + 2 */
    int64_t base;       /* gadget base */
    uint64_t q;         /* modulus, < 2^62 */
    double s;           /* Gaussian parameter s */
    double r;           /* rounding parameter r (perturbation only; 1 otherwise) */
    uint64_t norm_bound; /* floor of the check_domain bound (s^2 m [r^2]); 0 = derive from s, r */
} qf_params;

typedef struct qf_ctx qf_ctx;

/* ---- context ----------------------------------------------------------- */
qf_status qf_ctx_create(const qf_params* params, int device, qf_ctx** out);
void qf_ctx_destroy(qf_ctx* ctx);
const char* qf_last_error(const qf_ctx* ctx);
/* run on a caller-owned CUDA stream (cudaStream_t); NULL = the context's own stream */
qf_status qf_set_stream(qf_ctx* ctx, void* cuda_stream);
/* targets processed per internal chunk (bounds the workspace); 0 = default.  The default keeps the work matrices of
 * samp_p within ~32 GB of device memory (PSFGPV n = 256, q = 2^24: 37 888 targets = two waves of 148 SMs x 128 targets);
 * smaller chunks trade throughput for memory (one wave: -9 % at that size). */
qf_status qf_set_chunk(qf_ctx* ctx, int64_t targets_per_chunk);
qf_status qf_synchronize(qf_ctx* ctx);
/* number of kernels this context has launched so far */
uint64_t qf_launch_count(const qf_ctx* ctx);

/* Per-launch CUDA-event timing of the two contraction kernels on the context's stream: the fp64
 * DMMA contraction (gemm_*) and the tcgen05 int8 limb contraction (i8_*).  qf_profile_read
 * synchronises, returns the summed kernel time, the algorithmic operations (2*B*N*K per launch;
 * i8_issued_ops additionally counts every digit-pair product the tensor pipe executed) and the
 * launch counts, and resets the counters.  Any output pointer may be NULL. */
qf_status qf_profile(qf_ctx* ctx, int enable);
qf_status qf_profile_read(qf_ctx* ctx, double* gemm_ms, double* gemm_flops, uint64_t* gemm_launches, double* i8_ms,
                          double* i8_ops, double* i8_issued_ops, uint64_t* i8_launches);

/* ---- key material -------------------------------------------------------- */
/* A: n x m residues (PSFGPV / PSFPerturbation `A`, gpv.rs:60, mp_perturbation.rs:194) */
qf_status qf_set_a(qf_ctx* ctx, const int64_t* a);
/* PSFPerturbation trapdoor (mp_perturbation.rs:195): R m_bar x nk in {-1,0,1} (any small ints),
 * sqrt_sigma_2 m x m (any square root of Sigma_2; a lower-triangular one such as the Cholesky factor
 * compute_sqrt_sigma_2 returns costs half the work), and ONE k x k diagonal block of the gadget short basis
 * I_n (x) S_k with its GSO (gadget_classical.rs:248-287; the shim checks block-diagonality).
 * sqrt_sigma_2 == NULL: the backend derives its own square root of the DEFAULT covariance
 * Sigma = s^2 I (what PSFPerturbation::trap_gen builds, mp_perturbation.rs:227-231) from R, s, r on the
 * device, in block form (only an m_bar x m_bar Cholesky factor is dense) -- same law, ~4x less work per target. */
qf_status qf_set_trapdoor_perturbation(qf_ctx* ctx, const int8_t* r, const double* sqrt_sigma_2,
                                       const int64_t* s_block, const double* s_block_gso);
/* gen_short_basis_for_trapdoor with tag = I (short_basis_classical.rs:54-110): the short basis
 * S_A = [[I,R],[0,I]] [[0,I],[S',W]] = [[R S', I + R W],[S', W]] of Lambda^perp(A) for the installed A and the
 * trapdoor R (m_bar x nk); W = digits of -A[:, :m_bar], S' column-reversed iff base^k = q.  The m_bar x nk x m_bar
 * product R W runs on the tensor cores; s_out is m x m (host), columns are the basis vectors.  Exact. */
qf_status qf_gen_short_basis(qf_ctx* ctx, const int8_t* r, int64_t* s_out);
/* compute_sqrt_sigma_2 (mp_perturbation.rs:111-139) on the device: lower Cholesky factor (m x m, row-major, host)
 * of Sigma_2 = r^2/(2 pi) (Sigma - (b^2+1) [R;I][R;I]^t - I); sigma == NULL means Sigma = s^2 I.
 * QF_ERR_INVALID when Sigma_2 is not positive definite (the reference panics in the Cholesky). */
qf_status qf_compute_sqrt_sigma_2(qf_ctx* ctx, const int8_t* r, const double* sigma, double* sqrt_sigma_2_out);
/* PSFGPV trapdoor (gpv.rs:61): short basis S (dim x dim, columns are basis vectors) and its
 * GSO.  dim = m for QF_PSF_GPV, n*(k+2) (coefficient embedding) for QF_PSF_GPV_RING.
 * s_gso == NULL: the GSO is computed on the device (as qf_gso) and never leaves it. */
qf_status qf_set_trapdoor_gpv(qf_ctx* ctx, const int64_t* s, const double* s_gso);
/* MatQ::gso as used at gpv.rs:91: unnormalised Gram-Schmidt of the columns of S (dim x dim) in fp64 on the
 * device (blocked Gram-Schmidt with re-orthogonalisation).  gso_out: dim x dim, host. */
qf_status qf_gso(qf_ctx* ctx, const int64_t* s, double* gso_out);
/* gen_short_basis_for_trapdoor_ring (short_basis_ring.rs:64-166) for the installed ring key and the trapdoor (r, e),
 * k x n small coefficients each, in the coefficient embedding: s_out is dim x dim (host), dim = n (k + 2), row =
 * polynomial row * n + coefficient, columns are the basis vectors.  Exact; feed it to qf_set_trapdoor_gpv. */
qf_status qf_ring_gen_short_basis(qf_ctx* ctx, const int32_t* r, const int32_t* e, int64_t* s_out);
/* ring key: (k+2) polynomials of n coefficients (gpv_ring.rs:70) */
qf_status qf_ring_set_a(qf_ctx* ctx, const int64_t* a);

/* ---- TrapGen (gadget_classical.rs:56-68, gadget_ring.rs:62-81) ------------ */
/* A = [A_bar | tag*G - A_bar*R] mod q from supplied A_bar (n x m_bar), R (m_bar x nk) and
 * tag (n x n, NULL = identity).  Bit-exact against the reference given the same inputs.
 * Writes A (n x m) and installs it as the context key. */
qf_status qf_trap_gen_from(qf_ctx* ctx, const int64_t* a_bar, const int8_t* r, const int64_t* tag, int64_t* a_out);
/* Samples A_bar uniform (gpv.rs:84) and R from PlusMinusOneZero (trapdoor_distribution.rs:82-86)
 * on the device (Philox, `seed`), then as qf_trap_gen_from with tag = I. */
qf_status qf_trap_gen(qf_ctx* ctx, uint64_t seed, int64_t* a_out, int8_t* r_out);
/* Ring: A = [1 | a_bar | g^t - (a_bar*r + e)] mod (X^n+1, q); r, e: k x n small coefficients. */
qf_status qf_ring_trap_gen_from(qf_ctx* ctx, const int64_t* a_bar, const int32_t* r, const int32_t* e,
                                int64_t* a_out);

/* ---- PSF::f_a / check_domain ------------------------------------------------ */
/* u[b] = A * sigma[b] mod q and in_domain[b] = check_domain(sigma[b])
 * (gpv.rs:190-193,219-224; mp_perturbation.rs:366-369,396-402; gpv_ring.rs:243-247,274-283).
 * qf_f_a (host pointers, synchronous) returns QF_ERR_NOT_IN_DOMAIN if any flag is 0 (u rows of such targets are
 * unspecified); the shim turns that into the reference's panic.  Its in_domain may be NULL.
 * qf_f_a_dev is asynchronous and therefore cannot return the verdict: it reports domain failures ONLY through
 * in_domain (device pointer, B bytes), which is mandatory -- NULL is QF_ERR_INVALID, so the assertion of gpv.rs:191
 * can never be dropped silently.  u_out may be NULL (flags only).
 * The squared norm is exact and saturating: entries of any int32 magnitude are handled (no 64-bit wrap-around). */
qf_status qf_f_a(qf_ctx* ctx, const int32_t* sigma, int64_t batch, int64_t* u_out, uint8_t* in_domain);
qf_status qf_f_a_dev(qf_ctx* ctx, const int32_t* sigma, int64_t batch, int64_t* u_out, uint8_t* in_domain);
qf_status qf_check_domain(qf_ctx* ctx, const int32_t* sigma, int64_t batch, uint8_t* in_domain);

/* ---- PSF::samp_d (gpv.rs:113-116, mp_perturbation.rs:264-267, gpv_ring.rs:118-122) --- */
qf_status qf_samp_d(qf_ctx* ctx, int64_t batch, uint64_t seed, uint64_t first_index, int32_t* out);
qf_status qf_samp_d_dev(qf_ctx* ctx, int64_t batch, uint64_t seed, uint64_t first_index, int32_t* out);

/* ---- PSF::samp_p (gpv.rs:152-161, mp_perturbation.rs:304-336, gpv_ring.rs:160-212) --- */
/* e[b] with A e[b] = u[b] mod q, distributed as the reference's sampler. */
qf_status qf_samp_p(qf_ctx* ctx, const int64_t* u, int64_t batch, uint64_t seed, uint64_t first_index,
                    int32_t* e_out);
qf_status qf_samp_p_dev(qf_ctx* ctx, const int64_t* u, int64_t batch, uint64_t seed, uint64_t first_index,
                        int32_t* e_out);
/* Narrow Domain form: the same preimages as int16 -- |e_i| <= 6 s r is far below 2^15 for the parameter sets of the
 * reference (C2: 6 s = 8568; C3: 3135), so the device->host copy, which bounds the host-buffer path, is halved.  An entry
 * that does not fit is reported as QF_ERR_NUMERIC (then use qf_samp_p).  Targets outside [0,q) are detected on the device
 * (both variants) and reported as QF_ERR_INVALID when the call returns. */
qf_status qf_samp_p_i16(qf_ctx* ctx, const int64_t* u, int64_t batch, uint64_t seed, uint64_t first_index,
                        int16_t* e_out);
/* f_a / check_domain on int16 Domain values (host pointers; otherwise as qf_f_a). */
qf_status qf_f_a_i16(qf_ctx* ctx, const int16_t* sigma, int64_t batch, int64_t* u_out, uint8_t* in_domain);
/* int32 -> int16 on the device (e.g. before the final gather of results over NVLink); *overflow (device int, zeroed by
 * the caller) gets bit 64 set when a value does not fit.  in / out 16-byte aligned. */
qf_status qf_narrow_i32_i16_dev(const int32_t* in, int16_t* out, size_t count, int* overflow, void* cuda_stream);

/* ---- PSFPerturbation::randomized_nearest_plane_gadget (mp_perturbation.rs:173-191), the public helper samp_p calls:
 * z[b] = x0 + SampleD(S, S~, -x0, r sqrt(base^2 + 1)) with x0 = find_solution_gadget_mat(v[b]) -- a preimage of v[b]
 * under the gadget matrix G, Gaussian over the coset Lambda_v^perp(G).  v: B x n residues, z_out: B x (n k).  The gadget
 * short basis (and its GSO) is the one installed with qf_set_trapdoor_perturbation.  Host pointers. */
qf_status qf_randomized_nearest_plane_gadget(qf_ctx* ctx, const int64_t* v, int64_t batch, uint64_t seed,
                                             uint64_t first_index, int32_t* z_out);

/* ---- LossyCompressionFIPS203 (lossy_compression_fips203.rs:89-114, 143-172) ----------- */
/* Flat coefficient streams (a polynomial / matrix of polynomials is just `count` coefficients).
 * u16 variants: q < 2^16, 1 <= d <= 16; device_ptrs != 0 -> in/out are device pointers and the
 * call is asynchronous on `cuda_stream` (may be NULL = default stream). */
qf_status qf_compress_u16(const uint16_t* in, uint16_t* out, size_t count, uint32_t q, uint32_t d,
                          int device_ptrs, void* cuda_stream);
qf_status qf_decompress_u16(const uint16_t* in, uint16_t* out, size_t count, uint32_t q, uint32_t d,
                            int device_ptrs, void* cuda_stream);
/* FLINT-word variants, any q < 2^62, 1 <= d <= 62. */
qf_status qf_compress_i64(const int64_t* in, int64_t* out, size_t count, uint64_t q, uint32_t d,
                          int device_ptrs, void* cuda_stream);
qf_status qf_decompress_i64(const int64_t* in, int64_t* out, size_t count, uint64_t q, uint32_t d,
                            int device_ptrs, void* cuda_stream);

/* ---- FIPS 203 ByteEncode_d / ByteDecode_d fused with Compress_d / Decompress_d (the step after
 * `lossy_compress` in ML-KEM; the reference stops at the unpacked polynomial, lossy_compression_fips203.rs:62,99-113).
 * Polynomials of degree 256: `in` holds npoly x 256 u16 coefficients, the packed form npoly x 32 d bytes (FIPS 203
 * Algorithm 5: coefficient i supplies stream bits [i d, (i+1) d), bit k of the stream = bit k % 8 of byte k / 8).
 * 1 <= d <= 12, q < 2^16.  compress != 0: ByteEncode_d(Compress_d(x)), else ByteEncode_d(x) (d = 12: x < q).
 * decompress != 0: Decompress_d(ByteDecode_d(b)), else ByteDecode_d(b) (d = 12: reduced mod q, Algorithm 6). */
qf_status qf_compress_encode_u16(const uint16_t* in, uint8_t* out, size_t npoly, uint32_t q, uint32_t d, int compress,
                                 int device_ptrs, void* cuda_stream);
qf_status qf_decode_decompress_u16(const uint8_t* in, uint16_t* out, size_t npoly, uint32_t q, uint32_t d, int decompress,
                                   int device_ptrs, void* cuda_stream);

/* ---- message encodings (src/utils/common_encodings.rs), batched ------------------------------------
 * encode_value_in_polynomialringzq (:49-92): digit d of the value w.r.t. `base` -> coefficient d * floor(q / base);
 * decode_value_from_polynomialringzq (:125-153): coefficient c (any representative) -> digit
 * floor((c * base + floor(q / (2 base))) / q) mod base.  A batch is a flat stream of `count` digits, one byte each
 * (2 <= base <= 256); coefficients are coeff_bytes wide: 2 (u16, q < 2^16) or 8 (FLINT words, q < 2^62).  The big-integer
 * <-> digit-string conversion of the reference's `value: Z` stays with the caller (one value = up to n digits, padded
 * with zeros).  base < 2 is QF_ERR_INVALID (the reference returns MathError::InvalidIntegerInput).  device_ptrs != 0:
 * device pointers, asynchronous on cuda_stream. */
qf_status qf_encode_digits(const uint8_t* digits, void* coeffs, size_t count, uint64_t q, uint32_t base, int coeff_bytes,
                           int device_ptrs, void* cuda_stream);
qf_status qf_decode_digits(const void* coeffs, uint8_t* digits, size_t count, uint64_t q, uint32_t base, int coeff_bytes,
                           int device_ptrs, void* cuda_stream);
/* base 2 with bit-packed messages (bit j of byte i = coefficient 8 i + j; a 32-byte message <-> 256 u16 coefficients, the
 * mu -> floor(q/2) mu map of the lib.rs example :28-36): nbytes message bytes <-> 8 nbytes coefficients, q < 2^16. */
qf_status qf_encode_bits_u16(const uint8_t* msg, uint16_t* coeffs, size_t nbytes, uint32_t q, int device_ptrs, void* cuda_stream);
qf_status qf_decode_bits_u16(const uint16_t* coeffs, uint8_t* msg, size_t nbytes, uint32_t q, int device_ptrs, void* cuda_stream);

/* ---- Z::sample_discrete_gauss for a batch of (centre) values (qfall-math SampleZ as used at
 * gpv.rs:115 and inside sample_d_precomputed_gso): out[i] <- D_{Z, s, centers[i]}. Host pointers, any count.
 * Value i draws from its own Philox stream (seed, i) under a stream id that no PSF method uses, so equal seeds
 * here and in qf_samp_d give independent outputs.  s < 2e6 (QF_ERR_UNSUPPORTED above); a NaN / infinite centre
 * gives QF_ERR_NUMERIC. */
qf_status qf_sample_z(const double* centers, size_t count, double s, uint64_t seed, int64_t* out);

/* ---- self-test of the tensor-core integer contraction: out = X W^t (mod q if q != 0), exact.
 * x: B x K signed with |x| < 2^(8 LX - 1); w: N x K, residues < 2^(8 LW) (w_signed = 0) or signed
 * |w| < 2^(8 LW - 1) (w_signed = 1).  Host pointers. */
qf_status qf_debug_gemm_i8(const int64_t* x, const int64_t* w, int w_signed, int LX, int LW, int64_t B, int64_t N,
                           int64_t K, uint64_t q, int64_t* out);

/* ---- measured ceiling of the int8 tensor pipe (roofline denominator of the limb contractions; SURVEY 8d: "measure a plain
 * int8 tcgen05 GEMM probe once and use it"): a one-digit-pair B x N x K contraction on random bytes (128 x 256 tiles, K a
 * multiple of 128, <= 65536), `iters` launches timed one by one -> best_tops (burst), then back to back for sustain_ms ->
 * sustained_tops.  TOP/s = 2 B N K / time.  pipe_tops (optional): the tensor pipe on its own -- every SM repeats the MMAs of
 * one shared-memory-resident 128 x 256 x 128 block (no operand traffic), i.e. the rate at 100 % pipe activity; the GEMM probe
 * sits below it because a cta_group::1 tiling is bound by the L2 -> shared-memory operand feed.  Allocates its own buffers. */
qf_status qf_probe_i8_peak(int device, int64_t B, int64_t N, int64_t K, int iters, double sustain_ms, double* best_tops,
                           double* sustained_tops, double* pipe_tops);

/* ---- synthetic inputs for benchmarks (Philox, on device) -------------------------------- */
qf_status qf_fill_uniform_modq_dev(int64_t* out, size_t count, uint64_t q, uint64_t seed, void* cuda_stream);

/* library / build identification */
const char* qf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QFALL_B200_H */
