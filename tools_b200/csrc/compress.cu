// FIPS 203 Compress_d / Decompress_d over a flat coefficient stream
// (lossy_compression_fips203.rs:101-111 and :159-169).
//
//   compress  : y = floor((x * 2^d + floor(q/2)) / q) mod 2^d        x in [0,q)
//   decompress: x = floor((y * q + 2^(d-1)) / 2^d)                    (written unreduced)
//
// HBM-streaming kernel: u16 in, u16 out (4 algorithmic bytes per coefficient),
// 128-bit loads/stores (8 coefficients per access, 4 accesses in flight per
// thread), grid = 148 SMs x 8 resident CTAs, no shared memory.  The division by
// q is an exact multiply-high by ceil(2^64 / q) (numerator < 2^33, q < 2^16).
#include "common.cuh"
#include "kernels.h"

namespace {

struct CParams {
    uint32_t q, d, half_q, mask, round;
    unsigned long long magic;  // ceil(2^64 / q)
    uint32_t magic32;          // ceil(2^(32+sh) / q) for the 32-bit path
    uint32_t sh;
    uint32_t narrow;           // 1: numerator < 2^28, the 32-bit multiply-high is exact
};

__device__ __forceinline__ uint32_t comp1(uint32_t x, const CParams& p) {
    if (p.narrow) {
        // num < 2^28, magic32 = ceil(2^(32+sh)/q) with 2^sh <= q: floor(num * magic32 / 2^(32+sh)) is exact
        // because the excess num * (magic32 q - 2^(32+sh)) / (q 2^(32+sh)) < 2^28 / 2^32 / ... < 1/q
        const uint32_t num = (x << p.d) + p.half_q;
        return (__umulhi(num, p.magic32) >> p.sh) & p.mask;
    }
    unsigned long long num = ((unsigned long long)x << p.d) + p.half_q;
    return (uint32_t)__umul64hi(num, p.magic) & p.mask;
}
__device__ __forceinline__ uint32_t decomp1(uint32_t y, const CParams& p) {
    return (y * p.q + p.round) >> p.d;  // y, q < 2^16: fits 32 bits
}

template <bool DEC>
__device__ __forceinline__ uint32_t map2(uint32_t w, const CParams& p) {
    uint32_t lo = w & 0xffffu, hi = w >> 16;
    if (DEC) {
        lo = decomp1(lo, p);
        hi = decomp1(hi, p);
    } else {
        lo = comp1(lo, p);
        hi = comp1(hi, p);
    }
    return (lo & 0xffffu) | (hi << 16);
}

template <bool DEC>
__global__ void __launch_bounds__(256)
compress_u16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, size_t count, CParams p) {
    const size_t nvec = count / 8;
    const uint4* vin = reinterpret_cast<const uint4*>(in);
    uint4* vout = reinterpret_cast<uint4*>(out);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // four independent 128-bit loads in flight per thread (128 KB per SM at full occupancy) before any is consumed
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcs(vin + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u].x = map2<DEC>(v[u].x, p); v[u].y = map2<DEC>(v[u].y, p);
            v[u].z = map2<DEC>(v[u].z, p); v[u].w = map2<DEC>(v[u].w, p);
            __stcs(vout + i + u * stride, v[u]);
        }
    }
    for (; i < nvec; i += stride) {
        uint4 a = __ldcs(vin + i);
        a.x = map2<DEC>(a.x, p); a.y = map2<DEC>(a.y, p); a.z = map2<DEC>(a.z, p); a.w = map2<DEC>(a.w, p);
        __stcs(vout + i, a);
    }
    // ragged tail (count not a multiple of 8)
    size_t t = nvec * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) out[t] = (uint16_t)(DEC ? decomp1(in[t], p) : comp1(in[t], p));
}

// General path on FLINT words (int64 in/out), any q < 2^62, 1 <= d <= 62.
__global__ void compress_i64_kernel(const int64_t* __restrict__ in, int64_t* __restrict__ out, size_t count,
                                    unsigned long long q, uint32_t d, int dec) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        unsigned __int128 x = (unsigned __int128)(unsigned long long)in[i];
        if (dec) {
            unsigned __int128 v = x * q + ((unsigned __int128)1 << (d - 1));
            out[i] = (int64_t)(unsigned long long)(v >> d);
        } else {
            unsigned __int128 v = (x << d) + (q >> 1);
            unsigned __int128 y = v / q;
            out[i] = (int64_t)((unsigned long long)y & ((d >= 64) ? ~0ull : ((1ull << d) - 1)));
        }
    }
}

// ---- Compress_d + ByteEncode_d and ByteDecode_d + Decompress_d in one pass (FIPS 203 Algorithms 5 and 6; SURVEY 8f rank
// 4: the step that follows compression in ML-KEM).  A polynomial is 256 coefficients -> 32 d bytes: coefficient i
// contributes its d low bits to stream bits [i d, (i + 1) d), bit k of the stream is bit k % 8 of byte k / 8.
// One warp per polynomial: a lane owns 8 consecutive coefficients (one 128-bit load) = exactly d bytes of the stream,
// the warp's 32 d bytes are assembled in shared memory and leave as 128-bit stores, so both HBM streams are coalesced
// and the packed form is the only thing written: 512 + 32 d algorithmic bytes per polynomial instead of 1024.
constexpr int PK_WARPS = 8;    // warps per CTA
constexpr int PK_MAXD = 12;

// store / load the lane's D bytes at byte offset lane * D of the warp's staging row with the widest accesses the
// alignment of lane * D allows (D % 4 == 0: words, D even: half-words, else bytes); r[] = the 8 D bits, little-endian
template <int D>
__device__ __forceinline__ void put_bytes(uint8_t* dst, const uint32_t (&r)[3]) {
    if (D % 4 == 0) {
#pragma unroll
        for (int w = 0; w < D / 4; ++w) reinterpret_cast<uint32_t*>(dst)[w] = r[w];
    } else if (D % 2 == 0) {
#pragma unroll
        for (int h = 0; h < D / 2; ++h) reinterpret_cast<uint16_t*>(dst)[h] = (uint16_t)(r[h >> 1] >> ((h & 1) * 16));
    } else {
#pragma unroll
        for (int b = 0; b < D; ++b) dst[b] = (uint8_t)(r[b >> 2] >> ((b & 3) * 8));
    }
}
template <int D>
__device__ __forceinline__ void get_bytes(const uint8_t* src, uint32_t (&r)[3]) {
    r[0] = r[1] = r[2] = 0;
    if (D % 4 == 0) {
#pragma unroll
        for (int w = 0; w < D / 4; ++w) r[w] = reinterpret_cast<const uint32_t*>(src)[w];
    } else if (D % 2 == 0) {
#pragma unroll
        for (int h = 0; h < D / 2; ++h) r[h >> 1] |= (uint32_t)reinterpret_cast<const uint16_t*>(src)[h] << ((h & 1) * 16);
    } else {
#pragma unroll
        for (int b = 0; b < D; ++b) r[b >> 2] |= (uint32_t)src[b] << ((b & 3) * 8);
    }
}

// D compile time: every bit position is a constant, the 8 values of a lane are packed with 32-bit shifts only.
template <int D, bool COMPRESS>
__global__ void __launch_bounds__(PK_WARPS * 32)
encode_kernel(const uint16_t* __restrict__ in, uint8_t* __restrict__ out, size_t npoly, CParams p) {
    __shared__ __align__(16) uint8_t stage[PK_WARPS][32 * D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t wstride = (size_t)gridDim.x * PK_WARPS;
    const uint4* vin = reinterpret_cast<const uint4*>(in);
    size_t poly = (size_t)blockIdx.x * PK_WARPS + warp;
    uint4 v = make_uint4(0, 0, 0, 0), vnext = v;
    if (poly < npoly) v = __ldcs(vin + poly * 32 + lane);
    for (; poly < npoly; poly += wstride) {
        const size_t nxt = poly + wstride;
        if (nxt < npoly) vnext = __ldcs(vin + nxt * 32 + lane);  // next polynomial in flight while this one is packed
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t r[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t x = (k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffffu);
            x = COMPRESS ? comp1(x, p) : (x & ((1u << D) - 1));
            const int pos = k * D, wi = pos >> 5, sh = pos & 31;
            r[wi] |= x << sh;
            if (sh + D > 32) r[wi + 1] |= x >> (32 - sh);
        }
        put_bytes<D>(&stage[warp][lane * D], r);
        __syncwarp();
        if (lane < 2 * D)
            __stcs(reinterpret_cast<uint4*>(out + poly * 32 * D) + lane, reinterpret_cast<const uint4*>(stage[warp])[lane]);
        __syncwarp();
        v = vnext;
    }
}

template <int D, bool DECOMPRESS>
__global__ void __launch_bounds__(PK_WARPS * 32)
decode_kernel(const uint8_t* __restrict__ in, uint16_t* __restrict__ out, size_t npoly, CParams p, uint32_t modq) {
    __shared__ __align__(16) uint8_t stage[PK_WARPS][32 * D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t wstride = (size_t)gridDim.x * PK_WARPS;
    uint4* vout = reinterpret_cast<uint4*>(out);
    size_t poly = (size_t)blockIdx.x * PK_WARPS + warp;
    uint4 raw = make_uint4(0, 0, 0, 0), rnext = raw;
    if (poly < npoly && lane < 2 * D) raw = __ldcs(reinterpret_cast<const uint4*>(in + poly * 32 * D) + lane);
    for (; poly < npoly; poly += wstride) {
        const size_t nxt = poly + wstride;
        if (nxt < npoly && lane < 2 * D) rnext = __ldcs(reinterpret_cast<const uint4*>(in + nxt * 32 * D) + lane);
        if (lane < 2 * D) reinterpret_cast<uint4*>(stage[warp])[lane] = raw;
        __syncwarp();
        uint32_t r[3];
        get_bytes<D>(&stage[warp][lane * D], r);
        __syncwarp();
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int pos = k * D, wi = pos >> 5, sh = pos & 31;
            uint32_t y = r[wi] >> sh;
            if (sh + D > 32) y |= r[wi + 1] << (32 - sh);
            y &= (1u << D) - 1;
            if (DECOMPRESS) y = decomp1(y, p);
            else if (D == 12 && modq) y = y >= modq ? y % modq : y;  // ByteDecode_12: integers mod q (Algorithm 6, m = q)
            w[k >> 1] |= (y & 0xffffu) << ((k & 1) * 16);
        }
        __stcs(vout + poly * 32 + lane, make_uint4(w[0], w[1], w[2], w[3]));
        raw = rnext;
    }
}

template <int D>
cudaError_t launch_encode_d(const uint16_t* in, uint8_t* out, size_t npoly, const CParams& p, int compress, int grid,
                            cudaStream_t stream) {
    if (compress) encode_kernel<D, true><<<grid, PK_WARPS * 32, 0, stream>>>(in, out, npoly, p);
    else encode_kernel<D, false><<<grid, PK_WARPS * 32, 0, stream>>>(in, out, npoly, p);
    return cudaGetLastError();
}
template <int D>
cudaError_t launch_decode_d(const uint8_t* in, uint16_t* out, size_t npoly, const CParams& p, int decompress, uint32_t q,
                            int grid, cudaStream_t stream) {
    if (decompress) decode_kernel<D, true><<<grid, PK_WARPS * 32, 0, stream>>>(in, out, npoly, p, 0u);
    else decode_kernel<D, false><<<grid, PK_WARPS * 32, 0, stream>>>(in, out, npoly, p, D == 12 ? q : 0u);
    return cudaGetLastError();
}

}  // namespace

cudaError_t qf_launch_compress_u16(const uint16_t* in, uint16_t* out, size_t count, uint32_t q, uint32_t d,
                                   int decompress, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    if (q < 2 || q > 65535 || d < 1 || d > 16) return cudaErrorInvalidValue;
    if ((((uintptr_t)in) & 15) || (((uintptr_t)out) & 15)) return cudaErrorMisalignedAddress;
    CParams p;
    p.q = q; p.d = d; p.half_q = q / 2; p.mask = (d >= 32) ? 0xffffffffu : ((1u << d) - 1);
    p.round = 1u << (d - 1);
    p.magic = ~0ull / q + 1;  // ceil(2^64/q) for q not a power of two; exact quotient otherwise as well
    {
        // 32-bit path: exact when num = x 2^d + q/2 < 2^n_bits and e * num < 2^(32+sh) with e = magic32 q - 2^(32+sh) < q
        uint32_t sh = 0;
        while ((2u << sh) <= q) ++sh;  // 2^sh <= q < 2^(sh+1)
        const unsigned long long pw = 1ull << (32 + sh);
        const unsigned long long m32 = (pw + q - 1) / q;  // < 2^33 / ... fits 32 bits since q >= 2^sh
        const unsigned long long e = m32 * q - pw;
        const unsigned long long num_max = ((unsigned long long)(q - 1) << d) + q / 2;
        p.sh = sh;
        p.magic32 = (uint32_t)m32;
        p.narrow = (m32 <= 0xffffffffull && num_max <= 0xffffffffull && e * num_max < pw) ? 1u : 0u;
    }
    size_t nvec = count / 8;
    size_t want = (nvec + 4 * 256 - 1) / (4 * 256);
    int grid = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
    if (decompress)
        compress_u16_kernel<true><<<grid, 256, 0, stream>>>(in, out, count, p);
    else
        compress_u16_kernel<false><<<grid, 256, 0, stream>>>(in, out, count, p);
    return cudaGetLastError();
}

cudaError_t qf_launch_compress_i64(const int64_t* in, int64_t* out, size_t count, unsigned long long q, uint32_t d,
                                   int decompress, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    if (q < 2 || q >= (1ull << 62) || d < 1 || d > 62) return cudaErrorInvalidValue;
    size_t want = (count + 255) / 256;
    int grid = (int)(want > 148 * 16 ? 148 * 16 : want);
    compress_i64_kernel<<<grid, 256, 0, stream>>>(in, out, count, q, d, decompress);
    return cudaGetLastError();
}

static void fill_cparams(CParams& p, uint32_t q, uint32_t d) {
    p.q = q; p.d = d; p.half_q = q / 2; p.mask = (d >= 32) ? 0xffffffffu : ((1u << d) - 1);
    p.round = 1u << (d - 1);
    p.magic = ~0ull / q + 1;
    uint32_t sh = 0;
    while ((2u << sh) <= q) ++sh;
    const unsigned long long pw = 1ull << (32 + sh);
    const unsigned long long m32 = (pw + q - 1) / q;
    const unsigned long long e = m32 * q - pw;
    const unsigned long long num_max = ((unsigned long long)(q - 1) << d) + q / 2;
    p.sh = sh;
    p.magic32 = (uint32_t)m32;
    p.narrow = (m32 <= 0xffffffffull && num_max <= 0xffffffffull && e * num_max < pw) ? 1u : 0u;
}

// in: npoly x 256 coefficients (u16), out: npoly x 32 d bytes.  compress != 0: ByteEncode_d(Compress_d(x)); else
// ByteEncode_d(x mod 2^d) (d = 12: the caller's coefficients are already < q < 2^12).
cudaError_t qf_launch_byte_encode(const uint16_t* in, uint8_t* out, size_t npoly, uint32_t q, uint32_t d, int compress,
                                  cudaStream_t stream) {
    if (npoly == 0) return cudaSuccess;
    if (q < 2 || q > 65535 || d < 1 || d > PK_MAXD) return cudaErrorInvalidValue;
    if ((((uintptr_t)in) & 15) || (((uintptr_t)out) & 15)) return cudaErrorMisalignedAddress;
    CParams p;
    fill_cparams(p, q, d);
    size_t want = (npoly + PK_WARPS - 1) / PK_WARPS;
    int grid = (int)(want > 148 * 8 ? 148 * 8 : want);
    switch (d) {
#define QF_CASE(D) case D: return launch_encode_d<D>(in, out, npoly, p, compress, grid, stream);
        QF_CASE(1) QF_CASE(2) QF_CASE(3) QF_CASE(4) QF_CASE(5) QF_CASE(6) QF_CASE(7) QF_CASE(8) QF_CASE(9) QF_CASE(10)
        QF_CASE(11) QF_CASE(12)
#undef QF_CASE
    }
    return cudaErrorInvalidValue;
}

// in: npoly x 32 d bytes, out: npoly x 256 coefficients.  decompress != 0: Decompress_d(ByteDecode_d(b)); else
// ByteDecode_d(b) (values mod 2^d; d = 12: mod q, FIPS 203 Algorithm 6).
cudaError_t qf_launch_byte_decode(const uint8_t* in, uint16_t* out, size_t npoly, uint32_t q, uint32_t d, int decompress,
                                  cudaStream_t stream) {
    if (npoly == 0) return cudaSuccess;
    if (q < 2 || q > 65535 || d < 1 || d > PK_MAXD) return cudaErrorInvalidValue;
    if ((((uintptr_t)in) & 15) || (((uintptr_t)out) & 15)) return cudaErrorMisalignedAddress;
    CParams p;
    fill_cparams(p, q, d);
    size_t want = (npoly + PK_WARPS - 1) / PK_WARPS;
    int grid = (int)(want > 148 * 8 ? 148 * 8 : want);
    switch (d) {
#define QF_CASE(D) case D: return launch_decode_d<D>(in, out, npoly, p, decompress, q, grid, stream);
        QF_CASE(1) QF_CASE(2) QF_CASE(3) QF_CASE(4) QF_CASE(5) QF_CASE(6) QF_CASE(7) QF_CASE(8) QF_CASE(9) QF_CASE(10)
        QF_CASE(11) QF_CASE(12)
#undef QF_CASE
    }
    return cudaErrorInvalidValue;
}
