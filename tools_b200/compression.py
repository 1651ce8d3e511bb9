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
