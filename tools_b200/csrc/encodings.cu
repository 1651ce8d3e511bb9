// Message encodings of lattice PKE schemes (src/utils/common_encodings.rs), batched:
//   encode  (common_encodings.rs:49-92):   digit d of the message w.r.t. `base` -> coefficient d * floor(q / base)
//   decode  (common_encodings.rs:125-153): coefficient c -> digit floor((c * base + floor(q / (2 base))) / q) mod base
// The reference walks one big integer per call; a batch is a flat stream of digits (one byte each, base <= 256) or, for
// base 2, of message BYTES (8 coefficients per byte: the ML-KEM style 32-byte message <-> 256 coefficients).  The
// conversion between big integers and digit strings stays with the caller (tools_b200/encodings.py, the Rust shim).
// HBM-streaming kernels: 128-bit accesses on the coefficient side.
#include "common.cuh"
#include "kernels.h"

namespace {

inline int enc_grid(size_t work, int per_block) {
    size_t g = (work + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > 148 * 8) g = 148 * 8;
    return (int)g;
}

template <typename CT>
__global__ void __launch_bounds__(256) encode_digits_kernel(const uint8_t* __restrict__ digits, CT* __restrict__ out, size_t count,
                                                            unsigned long long mul) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (CT)((unsigned long long)digits[i] * mul);
}

template <typename CT>
__global__ void __launch_bounds__(256) decode_digits_kernel(const CT* __restrict__ in, uint8_t* __restrict__ digits, size_t count,
                                                            unsigned long long q, unsigned long long base, unsigned long long q2b) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        // least non-negative representative (common_encodings.rs:139), then (c * base + floor(q / 2 base)) div q mod base
        long long c = (long long)in[i];
        unsigned long long r = c < 0 ? (unsigned long long)(q - (unsigned long long)(-c) % q) % q : (unsigned long long)c % q;
        const unsigned __int128 t = (unsigned __int128)r * base + q2b;
        digits[i] = (uint8_t)((unsigned long long)(t / q) % base);
    }
}

// base 2, q < 2^16: one message byte <-> 8 u16 coefficients (bit j of the byte = coefficient 8 i + j)
__global__ void __launch_bounds__(256) encode_bits_u16_kernel(const uint8_t* __restrict__ msg, uint16_t* __restrict__ out, size_t nbytes,
                                                              uint32_t half) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t b = msg[i];
        uint32_t w[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) w[t] = (((b >> (2 * t)) & 1u) * half) | ((((b >> (2 * t + 1)) & 1u) * half) << 16);
        __stcs(reinterpret_cast<uint4*>(out) + i, make_uint4(w[0], w[1], w[2], w[3]));
    }
}
__global__ void __launch_bounds__(256) decode_bits_u16_kernel(const uint16_t* __restrict__ in, uint8_t* __restrict__ msg, size_t nbytes,
                                                              uint32_t q, uint32_t q4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4*>(in) + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t b = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t c0 = (w[t] & 0xFFFFu) % q, c1 = (w[t] >> 16) % q;
            b |= (((2u * c0 + q4) / q) & 1u) << (2 * t);
            b |= (((2u * c1 + q4) / q) & 1u) << (2 * t + 1);
        }
        msg[i] = (uint8_t)b;
    }
}

}  // namespace

cudaError_t qf_launch_encode_digits(const uint8_t* digits, void* coeffs, size_t count, unsigned long long q, unsigned base,
                                    int coeff_bytes, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    const unsigned long long mul = q / base;  // floor(q / base), common_encodings.rs:87
    if (coeff_bytes == 2) encode_digits_kernel<uint16_t><<<enc_grid(count, 256), 256, 0, stream>>>(digits, (uint16_t*)coeffs, count, mul);
    else encode_digits_kernel<int64_t><<<enc_grid(count, 256), 256, 0, stream>>>(digits, (int64_t*)coeffs, count, mul);
    return cudaGetLastError();
}
cudaError_t qf_launch_decode_digits(const void* coeffs, uint8_t* digits, size_t count, unsigned long long q, unsigned base,
                                    int coeff_bytes, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    const unsigned long long q2b = q / (2ull * base);  // floor(q / (2 base)), common_encodings.rs:131
    if (coeff_bytes == 2)
        decode_digits_kernel<uint16_t><<<enc_grid(count, 256), 256, 0, stream>>>((const uint16_t*)coeffs, digits, count, q, base, q2b);
    else
        decode_digits_kernel<int64_t><<<enc_grid(count, 256), 256, 0, stream>>>((const int64_t*)coeffs, digits, count, q, base, q2b);
    return cudaGetLastError();
}
cudaError_t qf_launch_encode_bits_u16(const uint8_t* msg, uint16_t* coeffs, size_t nbytes, uint32_t q, cudaStream_t stream) {
    if (nbytes == 0) return cudaSuccess;
    if (((uintptr_t)coeffs) & 15) return cudaErrorMisalignedAddress;
    encode_bits_u16_kernel<<<enc_grid(nbytes, 256), 256, 0, stream>>>(msg, coeffs, nbytes, q / 2);
    return cudaGetLastError();
}
cudaError_t qf_launch_decode_bits_u16(const uint16_t* coeffs, uint8_t* msg, size_t nbytes, uint32_t q, cudaStream_t stream) {
    if (nbytes == 0) return cudaSuccess;
    if (((uintptr_t)coeffs) & 15) return cudaErrorMisalignedAddress;
    decode_bits_u16_kernel<<<enc_grid(nbytes, 256), 256, 0, stream>>>(coeffs, msg, nbytes, q, q / 4);
    return cudaGetLastError();
}
