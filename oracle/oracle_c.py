"""ctypes wrapper of oracle/liboracle_c.so (the timed CPU arm).  TEST/BENCH INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle_c.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _lib = C.CDLL(_LIB)
        _lib.orc_threads.restype = C.c_int
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def threads():
    return int(lib().orc_threads())


def sample_z(s, c, seed, count):
    out = np.empty(count, dtype=np.int64)
    lib().orc_sample_z(C.c_double(s), C.c_double(c), C.c_uint64(seed), C.c_long(count), _p(out))
    return out


def f_a(a, sigma, q, nthreads=1):
    a = np.ascontiguousarray(a, dtype=np.int64)
    sigma = np.ascontiguousarray(sigma, dtype=np.int32)
    n, m = a.shape
    b = sigma.shape[0]
    u = np.empty((b, n), dtype=np.int64)
    lib().orc_f_a(_p(a), _p(sigma), _p(u), C.c_long(b), C.c_long(n), C.c_long(m), C.c_uint64(q), C.c_int(nthreads))
    return u


def compress_u16(x, q, d, dec=False, nthreads=1):
    x = np.ascontiguousarray(x, dtype=np.uint16)
    out = np.empty_like(x)
    lib().orc_compress_u16(_p(x), _p(out), C.c_size_t(x.size), C.c_uint32(q), C.c_uint32(d), C.c_int(int(dec)), C.c_int(nthreads))
    return out


def samp_p_gpv(basis, gso, piv, ainv, u, q, s, seed, nthreads=1):
    """basis/gso: dim x dim with COLUMNS b_i / b~_i (reference orientation); transposed here so
    that rows are contiguous."""
    bt = np.ascontiguousarray(np.asarray(basis, dtype=np.float64).T)
    gt = np.ascontiguousarray(np.asarray(gso, dtype=np.float64).T)
    piv = np.ascontiguousarray(piv, dtype=np.int32)
    ainv = np.ascontiguousarray(ainv, dtype=np.int64)
    u = np.ascontiguousarray(u, dtype=np.int64)
    b, n = u.shape
    dim = bt.shape[0]
    e = np.empty((b, dim), dtype=np.int32)
    lib().orc_samp_p_gpv(_p(bt), _p(gt), _p(piv), _p(ainv), C.c_long(len(piv)), _p(u), _p(e), C.c_long(b), C.c_long(n),
                         C.c_long(dim), C.c_uint64(q), C.c_double(s), C.c_uint64(seed), C.c_int(nthreads))
    return e


def samp_p_pert(l, a, r, s_basis, s_gso, u, n, k, m_bar, base, q, r_par, seed, nthreads=1):
    l = np.ascontiguousarray(l, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.int64)
    r = np.ascontiguousarray(r, dtype=np.int8)
    sgt = np.ascontiguousarray(np.asarray(s_basis, dtype=np.float64).T)
    sggt = np.ascontiguousarray(np.asarray(s_gso, dtype=np.float64).T)
    u = np.ascontiguousarray(u, dtype=np.int64)
    b = u.shape[0]
    m = m_bar + n * k
    e = np.empty((b, m), dtype=np.int32)
    lib().orc_samp_p_pert(_p(l), _p(a), _p(r), _p(sgt), _p(sggt), _p(u), _p(e), C.c_long(b), C.c_long(n), C.c_long(k),
                          C.c_long(m_bar), C.c_long(base), C.c_uint64(q), C.c_double(r_par), C.c_uint64(seed),
                          C.c_int(nthreads))
    return e


def unit_pivots(a, q):
    """Pivot columns and A_P^{-1} rows (hoisted form of MatZq::solve_gaussian_elimination, gpv.rs:153-156):
    sol[piv[k]] = sum_j ainv[k][j] u_j mod q.  Pure Python/numpy-object; setup only."""
    import math

    a = np.asarray(a)
    n, m = a.shape
    t = np.eye(n, dtype=object)
    used = [False] * n
    piv, prow = [], []
    for c in range(m):
        if len(piv) == n:
            break
        w = t.dot(a[:, c].astype(object)) % q
        r = next((i for i in range(n) if not used[i] and w[i] and math.gcd(int(w[i]), q) == 1), None)
        if r is None:
            continue
        inv = pow(int(w[r]), -1, q)
        t[r] = (t[r] * inv) % q
        for i in range(n):
            if i != r and w[i]:
                t[i] = (t[i] - w[i] * t[r]) % q
        used[r] = True
        piv.append(c)
        prow.append(r)
    assert len(piv) == n, "A not surjective by unit pivoting"
    return np.array(piv, dtype=np.int32), np.array([t[r] for r in prow], dtype=object).astype(np.int64)
