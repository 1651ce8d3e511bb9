"""LossyCompressionFIPS203 (src/compression/lossy_compression_fips203.rs:20-59) over the C ABI.

A PolynomialRingZq / MatPolynomialRingZq value is passed as its coefficient array (any shape,
entries in [0,q)); the compressed value has the same shape (the reference's CompressedType is an
unpacked PolyOverZ / MatPolyOverZ, :62,176)."""
import numpy as np

from . import _ffi


def _run(name_u16, name_i64, arr, d, q):
    d = int(d)
    if d < 1:
        # lossy_compression_fips203.rs:91-94 / :149-152 assert!(d >= 1)
        raise AssertionError("Performing this function with d < 1 implies reducing mod 1")
    a = np.ascontiguousarray(arr)
    lib = _ffi.lib()
    if q < 65536 and d <= 16 and a.dtype == np.uint16:
        out = np.empty_like(a)
        st = getattr(lib, name_u16)(_ffi.ptr(a), _ffi.ptr(out), a.size, int(q), d, 0, None)
    else:
        a = np.ascontiguousarray(a, dtype=np.int64)
        out = np.empty_like(a)
        st = getattr(lib, name_i64)(_ffi.ptr(a), _ffi.ptr(out), a.size, int(q), d, 0, None)
    if st != _ffi.QF_OK:
        raise _ffi.QfError(st, f"{name_u16} failed")
    return out


def lossy_compress(coeffs, d, q):
    """PolynomialRingZq::lossy_compress / MatPolynomialRingZq::lossy_compress (:89-114, :203-217)."""
    return _run("qf_compress_u16", "qf_compress_i64", coeffs, d, q)


def lossy_decompress(compressed, d, q):
    """lossy_decompress (:143-172, :246-268); coefficients are written unreduced like the reference."""
    return _run("qf_decompress_u16", "qf_decompress_i64", compressed, d, q)


def compress_dev(in_ptr, out_ptr, count, q, d, stream=None, decompress=False):
    """Device-resident u16 stream (pointers are integers, e.g. torch.Tensor.data_ptr())."""
    if int(d) < 1:
        raise AssertionError("d < 1")
    lib = _ffi.lib()
    fn = lib.qf_decompress_u16 if decompress else lib.qf_compress_u16
    st = fn(_ffi.ptr(int(in_ptr)), _ffi.ptr(int(out_ptr)), int(count), int(q), int(d), 1,
            _ffi.ptr(int(stream)) if stream else None)
    if st != _ffi.QF_OK:
        raise _ffi.QfError(st, "compress_dev failed")


def _byte_code(fn_name, arr, npoly, out, q, d, flag):
    d = int(d)
    if d < 1:
        raise AssertionError("Performing this function with d < 1 implies reducing mod 1")
    st = getattr(_ffi.lib(), fn_name)(_ffi.ptr(arr), _ffi.ptr(out), int(npoly), int(q), d, int(flag), 0, None)
    if st != _ffi.QF_OK:
        raise _ffi.QfError(st, f"{fn_name} failed")
    return out


def compress_encode(coeffs, d, q, compress=True):
    """ByteEncode_d(Compress_d(f)) for degree-256 polynomials (FIPS 203 Algorithms 5 and 4.7; what ML-KEM does with the
    output of `lossy_compress`).  coeffs: (..., 256) u16 in [0, q) -> (..., 32 d) bytes."""
    a = np.ascontiguousarray(coeffs, dtype=np.uint16)
    if a.shape[-1] != 256:
        raise AssertionError("ByteEncode_d is defined on 256 coefficients")
    out = np.empty(a.shape[:-1] + (32 * int(d),), dtype=np.uint8)
    return _byte_code("qf_compress_encode_u16", a, a.size // 256, out, q, d, compress)


def decode_decompress(packed, d, q, decompress=True):
    """Decompress_d(ByteDecode_d(b)) (FIPS 203 Algorithm 6 then 4.8): (..., 32 d) bytes -> (..., 256) u16."""
    b = np.ascontiguousarray(packed, dtype=np.uint8)
    if b.shape[-1] != 32 * int(d):
        raise AssertionError("ByteDecode_d expects 32 d bytes per polynomial")
    out = np.empty(b.shape[:-1] + (256,), dtype=np.uint16)
    return _byte_code("qf_decode_decompress_u16", b, b.size // (32 * int(d)), out, q, d, decompress)


def byte_code_dev(in_ptr, out_ptr, npoly, q, d, stream=None, decode=False, flag=True):
    """Device-resident streams (pointers are integers): encode (u16 coefficients -> bytes) or decode."""
    lib = _ffi.lib()
    fn = lib.qf_decode_decompress_u16 if decode else lib.qf_compress_encode_u16
    st = fn(_ffi.ptr(int(in_ptr)), _ffi.ptr(int(out_ptr)), int(npoly), int(q), int(d), int(flag), 1,
            _ffi.ptr(int(stream)) if stream else None)
    if st != _ffi.QF_OK:
        raise _ffi.QfError(st, "byte_code_dev failed")
