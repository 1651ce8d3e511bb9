"""Message encodings of lattice PKE schemes (src/utils/common_encodings.rs) over the C ABI.

  encode_value_in_polynomialringzq (:49-92):    value -> digits w.r.t. `base` -> coefficient i = digit_i * floor(q / base)
  decode_value_from_polynomialringzq (:125-153): coefficient c -> digit floor((c base + floor(q / (2 base))) / q) mod base,
                                                 value = sum digit_i base^i

A PolynomialRingZq over X^n + 1 mod q is passed as its coefficient array of length n.  The reference takes one
arbitrary-size integer per call; the batch extension takes B of them.  The big-integer <-> digit-string conversion is host
work (Python integers); the coefficient map runs on the device (qf_encode_digits / qf_decode_digits, base <= 256)."""
import numpy as np

from . import _ffi


class MathError(ValueError):
    """The reference returns MathError::InvalidIntegerInput in these cases (common_encodings.rs:58-70, 133-137)."""


def _digits(value: int, base: int, n: int) -> np.ndarray:
    if base < 2:
        raise MathError(f"The given base {base} is smaller than 2")
    if value < 0:
        raise MathError(f"The given value {value} needs to be non-negative.")
    out = np.zeros(n, dtype=np.uint8 if base <= 256 else np.int64)
    i = 0
    while value > 0:
        if i >= n:
            raise MathError(f"The given value requires more than {n} digits represented w.r.t. base {base}.")
        value, d = divmod(value, base)
        out[i] = d
        i += 1
    return out


def encode_values_batch(values, base: int, n: int, q: int) -> np.ndarray:
    """B values -> (B, n) coefficients (uint16 when q < 2^16, else int64)."""
    base, n, q = int(base), int(n), int(q)
    digs = np.stack([_digits(int(v), base, n) for v in values]) if len(values) else np.zeros((0, n), dtype=np.uint8)
    wide = q > 65535
    out = np.empty(digs.shape, dtype=np.int64 if wide else np.uint16)
    if base > 256:  # outside the device kernel's digit width: same map on the host (one multiplication per digit)
        return (digs.astype(object) * (q // base)).astype(np.int64 if wide else np.uint16)
    digs = np.ascontiguousarray(digs, dtype=np.uint8)
    st = _ffi.lib().qf_encode_digits(_ffi.ptr(digs), _ffi.ptr(out), digs.size, q, base, 8 if wide else 2, 0, None)
    if st != _ffi.QF_OK:
        raise _ffi.QfError(st, "qf_encode_digits failed")
    return out


def encode_value_in_polynomialringzq(value, base, n, q) -> np.ndarray:
    """common_encodings.rs:49-92 for the ring Z_q[X]/(X^n + 1): the n coefficients of floor(q / base) * mu."""
    return encode_values_batch([value], base, n, q)[0]


def decode_values_batch(coeffs, base: int, q: int):
    """(B, n) coefficients (any representatives) -> list of B Python integers."""
    base, q = int(base), int(q)
    if base < 2:
        raise MathError(f"The given base {base} is smaller than 2, which does not allow the encoding of any information.")
    c = np.asarray(coeffs)
    assert c.ndim == 2
    if base > 256:
        digs = ((c.astype(object) % q) * base + q // (2 * base)) // q % base
    else:
        wide = not (c.dtype == np.uint16 and q <= 65535)
        cc = np.ascontiguousarray(c, dtype=np.int64 if wide else np.uint16)
        digs = np.empty(cc.shape, dtype=np.uint8)
        st = _ffi.lib().qf_decode_digits(_ffi.ptr(cc), _ffi.ptr(digs), cc.size, q, base, 8 if wide else 2, 0, None)
        if st != _ffi.QF_OK:
            raise _ffi.QfError(st, "qf_decode_digits failed")
    out = []
    for row in digs:
        v = 0
        for d in row[::-1]:  # Horner from the top coefficient, like the reference's loop (:143-150)
            v = v * base + int(d)
        out.append(v)
    return out


def decode_value_from_polynomialringzq(coeffs, base, q) -> int:
    """common_encodings.rs:125-153."""
    return decode_values_batch(np.asarray(coeffs).reshape(1, -1), base, q)[0]


def encode_bits(msg: np.ndarray, q: int) -> np.ndarray:
    """Base 2, bit-packed messages: (..., nbytes) uint8 -> (..., 8 nbytes) uint16 coefficients in {0, floor(q/2)}."""
    m = np.ascontiguousarray(msg, dtype=np.uint8)
    out = np.empty(m.shape[:-1] + (8 * m.shape[-1],), dtype=np.uint16)
    st = _ffi.lib().qf_encode_bits_u16(_ffi.ptr(m), _ffi.ptr(out), m.size, int(q), 0, None)
    if st != _ffi.QF_OK:
        raise _ffi.QfError(st, "qf_encode_bits_u16 failed")
    return out


def decode_bits(coeffs: np.ndarray, q: int) -> np.ndarray:
    c = np.ascontiguousarray(coeffs, dtype=np.uint16)
    assert c.shape[-1] % 8 == 0
    out = np.empty(c.shape[:-1] + (c.shape[-1] // 8,), dtype=np.uint8)
    st = _ffi.lib().qf_decode_bits_u16(_ffi.ptr(c), _ffi.ptr(out), out.size, int(q), 0, None)
    if st != _ffi.QF_OK:
        raise _ffi.QfError(st, "qf_decode_bits_u16 failed")
    return out
