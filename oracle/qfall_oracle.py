"""CPU oracle for the batched PSF / FIPS 203 hot path of qfall/tools.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tools_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs use it, and only as the checker or the timed CPU arm.

It restates, in plain Python big-integer arithmetic (exact) and numpy float64
(where the reference uses exact rationals for Gaussian parameters), the
algorithms of the reference crate.  Every function cites the reference
``file:line`` (relative to /root/reference) it follows.

Parity status
-------------
* Deterministic functions are pinned by the reference's own inline golden
  matrices (tests/golden/reference_goldens.json, extracted by
  tests/golden/extract_reference_goldens.py) -- see tests/test_oracle_golden.py.
* The arithmetic itself lives in the un-vendored dependency ``qfall-math = "0"``
  (Cargo.toml:18, no Cargo.lock => no pinned version) over ``flint-sys = "0.7"``.
  Its behaviour at the call sites is restated from the published algorithm
  (GPV08 SampleZ / SampleD, Peikert10 Alg. 1, MP12 Alg. 3).
* PARITY UNPINNED for: every sampler output (the reference has no seed API and
  no known-answer test), exact compress/decompress values (the reference only
  tests the round-trip bound; the formula at lossy_compression_fips203.rs:104-106,
  162-164 is the sole authority) and gso()/cholesky numerics.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass
from fractions import Fraction

import numpy as np

# --------------------------------------------------------------------------
# FLINT string formats used by the reference's golden vectors
# --------------------------------------------------------------------------


def parse_matz(s: str):
    """Parse a FLINT/qfall ``MatZ`` string ``"[[1, 2],[3, 4]]"`` (optionally
    followed by ``" mod q"``) -> (rows, q or None)."""
    s = s.strip()
    q = None
    m = re.match(r"^(.*\])\s*mod\s*(-?\d+)\s*$", s, re.S)
    if m:
        s, q = m.group(1), int(m.group(2))
    body = s.strip()
    assert body.startswith("[") and body.endswith("]")
    rows = re.findall(r"\[([^\[\]]*)\]", body)
    mat = [[int(x) for x in r.split(",") if x.strip() != ""] for r in rows]
    return mat, q


def parse_poly(s: str):
    """Parse a FLINT polynomial string ``"4  2 8 8 12"`` (length, two blanks,
    coefficients low->high); ``"0"`` is the zero polynomial."""
    toks = s.split()
    if not toks:
        return []
    n = int(toks[0])
    coeffs = [int(t) for t in toks[1:]]
    assert len(coeffs) == n, (s, n, coeffs)
    return coeffs


def parse_matpoly(s: str):
    """Parse a ``MatPolyOverZ`` string ``"[[1  1, 4  2 8 8 12],[0, 1  5]]"``."""
    rows = re.findall(r"\[([^\[\]]*)\]", s.strip())
    return [[parse_poly(e) for e in r.split(",")] for r in rows]


# --------------------------------------------------------------------------
# small exact-matrix helpers (lists of lists of Python ints)
# --------------------------------------------------------------------------


def mat_zeros(r, c):
    return [[0] * c for _ in range(r)]


def mat_identity(r, c=None):
    c = r if c is None else c
    return [[1 if i == j else 0 for j in range(c)] for i in range(r)]


def mat_mul(a, b, q=None):
    a = np.array(a, dtype=object)
    b = np.array(b, dtype=object)
    c = a.dot(b)
    if q is not None:
        c = c % q
    return c.tolist()


def mat_transpose(a):
    return [list(r) for r in zip(*a)]


def ceil_log(x: int, base: int) -> int:
    """qfall-math ``Z::log_ceil``: smallest e with base**e >= x."""
    e, p = 0, 1
    while p < x:
        p *= base
        e += 1
    return e


# --------------------------------------------------------------------------
# Gadget parameters  (gadget_parameters.rs)
# --------------------------------------------------------------------------


@dataclass
class GadgetParameters:
    """gadget_parameters.rs:44-52."""

    n: int
    k: int
    m_bar: int
    base: int
    q: int

    @staticmethod
    def init_default(n: int, q: int) -> "GadgetParameters":
        """gadget_parameters.rs:113-133."""
        assert n >= 1
        base = 2
        log_q = ceil_log(q, base)
        log_n = ceil_log(n, base)
        return GadgetParameters(n=n, k=log_q, m_bar=n * log_q + log_n**2, base=base, q=q)

    @property
    def m(self):
        return self.m_bar + self.n * self.k


@dataclass
class GadgetParametersRing:
    """gadget_parameters.rs:73-81; modulus is X^n + 1 mod q (common_moduli.rs:41-48)."""

    n: int
    k: int
    m_bar: int
    base: int
    q: int

    @staticmethod
    def init_default(n: int, q: int) -> "GadgetParametersRing":
        """gadget_parameters.rs:165-185."""
        assert n >= 1
        base = 2
        log_q = ceil_log(q, base)
        return GadgetParametersRing(n=n, k=log_q, m_bar=log_q + 2, base=base, q=q)


# --------------------------------------------------------------------------
# Classical gadget  (gadget_classical.rs)
# --------------------------------------------------------------------------


def gen_gadget_vec(k: int, base: int):
    """gadget_classical.rs:128-136 -> k x 1 column (1, b, ..., b^{k-1})."""
    return [[base**i] for i in range(k)]


def gen_gadget_mat(n: int, k: int, base: int):
    """gadget_classical.rs:91-107 -> n x nk, I_n (x) g^t."""
    out = mat_zeros(n, n * k)
    for j in range(n):
        for i in range(k):
            out[j][j * k + i] = base**i
    return out


def find_solution_gadget_vec(value: int, q: int, k: int, base: int):
    """gadget_classical.rs:169-182: base-b digits of the least non-negative residue."""
    if base**k < q:
        raise ValueError("The modulus is too large, the value is potentially not representable.")
    v = value % q
    out = []
    for _ in range(k):
        d = v % base
        out.append(d)
        v = (v - d) // base
    return out


def find_solution_gadget_mat(value, q: int, k: int, base: int):
    """gadget_classical.rs:219-229: digits of row j at rows k*j .. k*j+k-1."""
    rows, cols = len(value), len(value[0])
    out = mat_zeros(k * rows, cols)
    for i in range(cols):
        for j in range(rows):
            sol = find_solution_gadget_vec(value[j][i], q, k, base)
            for t in range(k):
                out[k * j + t][i] = sol[t]
    return out


def short_basis_gadget_block(k: int, base: int, q: int):
    """The k x k block S_k of gadget_classical.rs:248-272."""
    sk = mat_zeros(k, k)
    for j in range(k):
        sk[j][j] = base
    for i in range(k - 1):
        sk[i + 1][i] = -1
    if base**k != q:
        qq = q
        for i in range(k):
            qi = qq % base
            sk[i][k - 1] = qi
            qq = (qq - qi) // base
    return sk


def short_basis_gadget(p: GadgetParameters):
    """gadget_classical.rs:248-287 -> nk x nk, I_n (x) S_k."""
    sk = short_basis_gadget_block(p.k, p.base, p.q)
    out = mat_zeros(p.n * p.k, p.n * p.k)
    for j in range(p.n):
        for a in range(p.k):
            for b in range(p.k):
                out[j * p.k + a][j * p.k + b] = sk[a][b]
    return out


def sample_pm_one_zero(rng: np.random.Generator, m_bar: int, w: int):
    """trapdoor_distribution.rs:82-86: U{0,1} - U{0,1} entrywise."""
    a = rng.integers(0, 2, size=(m_bar, w))
    b = rng.integers(0, 2, size=(m_bar, w))
    return (a - b).tolist()


def gen_trapdoor(p: GadgetParameters, a_bar, tag, r):
    """gadget_classical.rs:56-68 with R supplied: A = [A_bar | tag*G - A_bar*R]."""
    g = gen_gadget_mat(p.n, p.k, p.base)
    hg = mat_mul(tag, g, p.q)
    ar = mat_mul(a_bar, r, p.q)
    right = [[(hg[i][j] - ar[i][j]) % p.q for j in range(p.n * p.k)] for i in range(p.n)]
    a = [[x % p.q for x in a_bar[i]] + right[i] for i in range(p.n)]
    return a


# --------------------------------------------------------------------------
# Classical short basis  (short_basis_classical.rs)
# --------------------------------------------------------------------------


def gen_sa_l(r):
    """short_basis_classical.rs:66-74: [[I, R],[0, I]]."""
    rr, rc = len(r), len(r[0])
    out = mat_identity(rr + rc)
    for i in range(rr):
        for j in range(rc):
            out[i][rr + j] = r[i][j]
    return out


def mat_inverse_mod(a, q):
    """Inverse of a square matrix over Z_q (q need not be prime): Gauss-Jordan
    with unit pivots.  Stands in for MatZq::inverse (short_basis_classical.rs:106)."""
    n = len(a)
    m = [[x % q for x in row] + [1 if i == j else 0 for j in range(n)] for i, row in enumerate(a)]
    for c in range(n):
        piv = None
        for r in range(c, n):
            if math.gcd(m[r][c], q) == 1:
                piv = r
                break
        if piv is None:
            raise ValueError("matrix not invertible by unit pivoting")
        m[c], m[piv] = m[piv], m[c]
        inv = pow(m[c][c], -1, q)
        m[c] = [(x * inv) % q for x in m[c]]
        for r in range(n):
            if r != c and m[r][c]:
                f = m[r][c]
                m[r] = [(x - f * y) % q for x, y in zip(m[r], m[c])]
    return [row[n:] for row in m]


def compute_w(p: GadgetParameters, tag, a):
    """short_basis_classical.rs:105-110: G*W = -H^{-1} * A[I|0]^t mod q."""
    tag_inv = mat_inverse_mod(tag, p.q)
    a_left = [row[: p.m_bar] for row in a]
    rhs = mat_mul(tag_inv, a_left, p.q)
    rhs = [[(-x) % p.q for x in row] for row in rhs]
    return find_solution_gadget_mat(rhs, p.q, p.k, p.base)


def gen_sa_r(p: GadgetParameters, tag, a):
    """short_basis_classical.rs:77-102: [[0, I],[S', W]] (S' column-reversed iff b^k = q)."""
    s = short_basis_gadget(p)
    if p.base**p.k == p.q:
        s = [list(reversed(row)) for row in s]
    w = compute_w(p, tag, a)
    s_rows, s_cols, w_cols = len(s), len(s[0]), len(w[0])
    out = mat_zeros(s_rows + w_cols, s_cols + w_cols)
    for d in range(w_cols):
        out[d][d + s_cols] = 1
    for i in range(s_rows):
        for j in range(s_cols):
            out[w_cols + i][j] = s[i][j]
        for j in range(w_cols):
            out[w_cols + i][s_cols + j] = w[i][j]
    return out


def gen_short_basis_for_trapdoor(p: GadgetParameters, tag, a, r):
    """short_basis_classical.rs:54-63."""
    return mat_mul(gen_sa_l(r), gen_sa_r(p, tag, a))


# --------------------------------------------------------------------------
# Rotation matrices  (utils/rotation_matrix.rs)
# --------------------------------------------------------------------------


def rot_minus(vec):
    """rotation_matrix.rs:41-63: column j = coefficients of a * X^j mod X^n + 1."""
    if len(vec) >= 1 and isinstance(vec[0], list):
        if len(vec[0]) == 1:
            v = [r[0] for r in vec]
        elif len(vec) == 1:
            v = list(vec[0])
        else:
            raise ValueError("The input must be a vector.")
    else:
        v = list(vec)
    n = len(v)
    out = mat_zeros(n, n)
    for i in range(n):
        for j in range(n):
            t = i + j
            if t >= n:
                out[t % n][j] = -v[i]
            else:
                out[t][j] = v[i]
    return out


def rot_minus_matrix(matrix):
    """rotation_matrix.rs:85-96: blocks rot^-(column i) concatenated horizontally."""
    rows, cols = len(matrix), len(matrix[0])
    blocks = [rot_minus([matrix[r][c] for r in range(rows)]) for c in range(cols)]
    return [sum((b[r] for b in blocks), []) for r in range(rows)]


# --------------------------------------------------------------------------
# Ring arithmetic over Z[X]/(X^n+1)
# --------------------------------------------------------------------------


def poly_trim(p):
    p = list(p)
    while p and p[-1] == 0:
        p.pop()
    return p


def poly_add(a, b):
    n = max(len(a), len(b))
    return poly_trim([(a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0) for i in range(n)])


def poly_neg(a):
    return [-x for x in a]


def poly_mul(a, b):
    if not a or not b:
        return []
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] += x * y
    return poly_trim(out)


def poly_reduce_anticyclic(p, n, q=None):
    """Reduce modulo X^n + 1 (and optionally into [0, q))."""
    out = [0] * n
    for i, c in enumerate(p):
        blk, pos = divmod(i, n)
        out[pos] += c if blk % 2 == 0 else -c
    if q is not None:
        out = [c % q for c in out]
    return out


def ring_mul(a, b, n, q):
    """Product in Z_q[X]/(X^n+1): the semantics rot_minus pins (rotation_matrix.rs:41-63)."""
    return poly_reduce_anticyclic(poly_mul(a, b), n, q)


def ring_mul_np(a, b, n, q):
    """Same product, numpy object arithmetic (exact), for n up to a few hundred."""
    a = np.array(list(a) + [0] * (n - len(a)), dtype=object)
    b = np.array(list(b) + [0] * (n - len(b)), dtype=object)
    full = np.convolve(a, b)
    out = full[:n].copy()
    out[: len(full) - n] -= full[n:]
    return [int(x) % q for x in out]


def coeff_embed(matpoly, n):
    """qfall-math ``into_coefficient_embedding(n)``: r x c polys -> (r*n) x c ints
    (consistent with gpv_ring.rs:172-178 and the golden at short_basis_ring.rs:415-443)."""
    r, c = len(matpoly), len(matpoly[0])
    out = mat_zeros(r * n, c)
    for i in range(r):
        for j in range(c):
            for t, x in enumerate(matpoly[i][j]):
                assert t < n
                out[i * n + t][j] = x
    return out


# --------------------------------------------------------------------------
# Ring gadget  (gadget_ring.rs)
# --------------------------------------------------------------------------


def gen_gadget_ring(k, base):
    """gadget_ring.rs:103-109: k x 1 constant polynomials base^j."""
    return [[[base**j]] for j in range(k)]


def gen_trapdoor_ring_lwe(p: GadgetParametersRing, a_bar, r, e):
    """gadget_ring.rs:62-81 with r, e supplied (each a list of k polys):
    A = [1 | a | g^t - (a*r + e)] reduced mod (X^n+1, q) -> list of k+2 polys."""
    out = [poly_reduce_anticyclic([1], p.n, p.q), poly_reduce_anticyclic(a_bar, p.n, p.q)]
    for j in range(p.k):
        t = poly_add(poly_mul(a_bar, r[j]), e[j])
        t = poly_add([p.base**j], poly_neg(t))
        out.append(poly_reduce_anticyclic(t, p.n, p.q))
    return out


def find_solution_gadget_ring(u, p: GadgetParametersRing):
    """gadget_ring.rs:145-166: poly i, coefficient j = digit i of coefficient j of u."""
    u = poly_reduce_anticyclic(u, p.n, p.q)
    digits = [find_solution_gadget_vec(c, p.q, p.k, p.base) for c in u]
    return [poly_trim([digits[j][i] for j in range(p.n)]) for i in range(p.k)]


# --------------------------------------------------------------------------
# Ring short basis  (short_basis_ring.rs)
# --------------------------------------------------------------------------


def ring_compute_s(p: GadgetParametersRing):
    """short_basis_ring.rs:142-166 -> k x k matrix of constant polynomials."""
    sk = short_basis_gadget_block(p.k, p.base, p.q)
    return [[poly_trim([x]) for x in row] for row in sk]


def ring_gen_sa_l(e, r):
    """short_basis_ring.rs:82-91: [[I_2, [e; r]],[0, I_k]] (entries are polynomials).
    NB the reference test calls gen_sa_l(&r, &e) with swapped names
    (short_basis_ring.rs:386); the first argument is the top row."""
    k = len(e)
    out = [[[] for _ in range(k + 2)] for _ in range(k + 2)]
    for i in range(k + 2):
        out[i][i] = [1]
    for j in range(k):
        out[0][2 + j] = poly_trim(e[j])
        out[1][2 + j] = poly_trim(r[j])
    return out


def ring_compute_w(p: GadgetParametersRing, a):
    """short_basis_ring.rs:128-139: digits of -a_0 and -a_1 as a k x 2 matrix."""
    w0 = find_solution_gadget_ring(poly_neg(a[0]), p)
    w1 = find_solution_gadget_ring(poly_neg(a[1]), p)
    return [[w0[i], w1[i]] for i in range(p.k)]


def _x_pow_times(i, poly):
    return poly_trim([0] * i + list(poly)) if poly else []


def ring_gen_sa_r(p: GadgetParametersRing, a):
    """short_basis_ring.rs:96-124: [[0],[X^i (x) S']] | X^i (x) [[I_2],[W]]."""
    n, k = p.n, p.k
    s = ring_compute_s(p)
    if p.base**p.k == p.q:
        s = [list(reversed(row)) for row in s]
    w = ring_compute_w(p, a)
    rows = k + 2
    out = [[[] for _ in range(n * k + 2 * n)] for _ in range(rows)]
    for i in range(n):
        for r in range(k):
            for c in range(k):
                out[2 + r][i * k + c] = _x_pow_times(i, s[r][c])
    off = n * k
    for i in range(n):
        for c in range(2):
            out[c][off + i * 2 + c] = _x_pow_times(i, [1])
            for r in range(k):
                out[2 + r][off + i * 2 + c] = _x_pow_times(i, w[r][c])
    return out


def matpoly_mul(a, b):
    r, inner, c = len(a), len(b), len(b[0])
    out = [[[] for _ in range(c)] for _ in range(r)]
    for i in range(r):
        for j in range(c):
            acc = []
            for t in range(inner):
                if a[i][t] and b[t][j]:
                    acc = poly_add(acc, poly_mul(a[i][t], b[t][j]))
            out[i][j] = acc
    return out


def gen_short_basis_for_trapdoor_ring(p: GadgetParametersRing, a, r, e):
    """short_basis_ring.rs:64-79: sa_l * sa_r reduced by X^n + 1 (over Z, no mod q)."""
    basis = matpoly_mul(ring_gen_sa_l(e, r), ring_gen_sa_r(p, a))
    return [[poly_trim(poly_reduce_anticyclic(x, p.n)) for x in row] for row in basis]


# --------------------------------------------------------------------------
# FIPS 203 lossy compression  (lossy_compression_fips203.rs)
# --------------------------------------------------------------------------


def compress_coeff(x: int, d: int, q: int) -> int:
    """lossy_compression_fips203.rs:104-106."""
    return ((x * (1 << d) + q // 2) // q) % (1 << d)


def decompress_coeff(y: int, d: int, q: int) -> int:
    """lossy_compression_fips203.rs:162-164 (written unreduced)."""
    return (y * q + (1 << (d - 1))) // (1 << d)


def lossy_compress(coeffs, d: int, q: int):
    """lossy_compression_fips203.rs:89-114 for one polynomial (coefficients in [0,q))."""
    if d < 1:
        raise ValueError("d < 1")
    return [compress_coeff(int(x), d, q) for x in coeffs]


def lossy_decompress(coeffs, d: int, q: int):
    """lossy_compression_fips203.rs:143-172."""
    if d < 1:
        raise ValueError("d < 1")
    return [decompress_coeff(int(y), d, q) for y in coeffs]


def lossy_compress_np(x: np.ndarray, d: int, q: int) -> np.ndarray:
    if d < 1:
        raise ValueError("d < 1")
    x = x.astype(np.uint64)
    return (((x << np.uint64(d)) + np.uint64(q // 2)) // np.uint64(q)) % np.uint64(1 << d)


def lossy_decompress_np(y: np.ndarray, d: int, q: int) -> np.ndarray:
    if d < 1:
        raise ValueError("d < 1")
    y = y.astype(np.uint64)
    return (y * np.uint64(q) + np.uint64(1 << (d - 1))) >> np.uint64(d)


def byte_encode(f, d: int):
    """FIPS 203 Algorithm 5 ByteEncode_d, restated literally (the step that follows
    ``lossy_compress`` in ML-KEM; NOT in the reference crate, whose compressed type stays an unpacked PolyOverZ,
    lossy_compression_fips203.rs:62 -- so PARITY UNPINNED by the reference; pinned here by hand-computed vectors in
    tests/test_oracle_golden.py).  f: 256 integers mod m (m = 2^d for d < 12, q for d = 12) -> 32 d bytes."""
    assert len(f) == 256 and 1 <= d <= 12
    bits = [0] * (256 * d)
    for i in range(256):
        a = int(f[i])
        for j in range(d):  # b[i d + j] <- a mod 2; a <- (a - b[i d + j]) / 2
            bits[i * d + j] = a & 1
            a >>= 1
    out = bytearray(32 * d)  # BitsToBytes (Algorithm 3): B[floor(i / 8)] += b[i] 2^(i mod 8)
    for i, b in enumerate(bits):
        out[i // 8] |= b << (i % 8)
    return bytes(out)


def byte_decode(b, d: int, q: int):
    """FIPS 203 Algorithm 6 ByteDecode_d: 32 d bytes -> 256 integers mod m (m = 2^d if d < 12 else q)."""
    assert len(b) == 32 * d and 1 <= d <= 12
    bits = [(b[i // 8] >> (i % 8)) & 1 for i in range(256 * d)]  # BytesToBits (Algorithm 4)
    m = (1 << d) if d < 12 else q
    return [sum(bits[i * d + j] << j for j in range(d)) % m for i in range(256)]


def byte_encode_np(x: np.ndarray, d: int) -> np.ndarray:
    """Vectorised ByteEncode_d for an (npoly, 256) array (same bit order as byte_encode)."""
    x = np.asarray(x, dtype=np.uint16)
    bits = ((x[..., None] >> np.arange(d, dtype=np.uint16)) & 1).astype(np.uint8)  # (..., 256, d): bit j of coeff i
    return np.packbits(bits.reshape(x.shape[:-1] + (256 * d,)), axis=-1, bitorder="little")


def byte_decode_np(b: np.ndarray, d: int, q: int) -> np.ndarray:
    b = np.asarray(b, dtype=np.uint8)
    bits = np.unpackbits(b, axis=-1, bitorder="little").reshape(b.shape[:-1] + (256, d)).astype(np.uint32)
    v = (bits << np.arange(d, dtype=np.uint32)).sum(-1)
    return (v % ((1 << d) if d < 12 else q)).astype(np.uint16)


# --------------------------------------------------------------------------
# f_a / check_domain  (gpv.rs, mp_perturbation.rs, gpv_ring.rs)
# --------------------------------------------------------------------------


def f_a_classical(a, sigma, q):
    """gpv.rs:190-193 / mp_perturbation.rs:366-369 without the assert: A * sigma mod q.
    a: n x m (list or ndarray), sigma: length-m vector -> length-n list in [0,q)."""
    a = np.array(a, dtype=object)
    s = np.array(sigma, dtype=object)
    return [int(x) % q for x in a.dot(s)]


def f_a_classical_batch(a, sigmas, q):
    a = np.array(a, dtype=object)
    s = np.array(sigmas, dtype=object)  # B x m
    return (s.dot(a.T) % q).astype(np.int64)


def check_domain_gpv(sigma, m: int, s: float) -> bool:
    """gpv.rs:219-224 (column-vector and length checks are the caller's shape checks)."""
    if len(sigma) != m:
        return False
    return sum(int(x) * int(x) for x in sigma) <= Fraction(s) ** 2 * m


def check_domain_perturbation(sigma, m: int, s: float, r: float) -> bool:
    """mp_perturbation.rs:396-402."""
    if len(sigma) != m:
        return False
    return sum(int(x) * int(x) for x in sigma) <= Fraction(s) ** 2 * m * Fraction(r) ** 2


def check_domain_ring(sigma_polys, n: int, k: int, s: float) -> bool:
    """gpv_ring.rs:274-283: coefficient embedding has n*(k+2) rows."""
    if len(sigma_polys) != k + 2:
        return False
    nrm = sum(int(c) * int(c) for p in sigma_polys for c in p)
    return nrm <= Fraction(s) ** 2 * (n * (k + 2))


def f_a_ring(a_polys, sigma_polys, n, q):
    """gpv_ring.rs:243-247: sum_j a_j * sigma_j mod (X^n+1, q)."""
    acc = [0] * n
    for aj, sj in zip(a_polys, sigma_polys):
        prod = ring_mul_np(aj, [c % q for c in sj], n, q)
        acc = [(x + y) % q for x, y in zip(acc, prod)]
    return acc


# --------------------------------------------------------------------------
# GSO / Cholesky (float64 stand-ins for MatQ::gso / cholesky_decomposition_flint)
# --------------------------------------------------------------------------


def gso_exact(basis):
    """Unnormalised Gram-Schmidt of the COLUMNS in exact rationals (MatQ::gso)."""
    cols = [[Fraction(x) for x in c] for c in mat_transpose(basis)]
    out = []
    for v in cols:
        w = list(v)
        for u in out:
            nu = sum(x * x for x in u)
            if nu == 0:
                continue
            mu = sum(x * y for x, y in zip(v, u)) / nu
            w = [a - mu * b for a, b in zip(w, u)]
        out.append(w)
    return mat_transpose(out)


def gso_f64(basis: np.ndarray) -> np.ndarray:
    """Unnormalised Gram-Schmidt of the columns in float64 via QR
    (b~_i = Q[:, i] * R[i, i])."""
    b = np.asarray(basis, dtype=np.float64)
    qm, rm = np.linalg.qr(b)
    return qm * np.diag(rm)[None, :]


def compute_sqrt_sigma_2(r_mat, s: float, r: float, base: int) -> np.ndarray:
    """mp_perturbation.rs:111-139 for Sigma = s^2 I:
    Sigma_2 = r^2/(2 pi) * (s^2 I - (b^2+1) T T^t - I), T = [R; I]; lower Cholesky."""
    rm = np.asarray(r_mat, dtype=np.float64)
    m_bar, nk = rm.shape
    t = np.vstack([rm, np.eye(nk)])
    m = m_bar + nk
    sigma_p = (s * s) * np.eye(m) - (base * base + 1) * (t @ t.T)
    sigma_2 = (r * r) / (2.0 * math.pi) * (sigma_p - np.eye(m))
    return np.linalg.cholesky(sigma_2)


# --------------------------------------------------------------------------
# Message encodings (src/utils/common_encodings.rs)
# --------------------------------------------------------------------------


def encode_value_in_polynomialringzq(value: int, base: int, n: int, q: int):
    """common_encodings.rs:49-92: digits of `value` w.r.t. `base` spread by floor(q / base) over the coefficients of a
    polynomial in Z_q[X]/(X^n + 1).  ValueError where the reference returns MathError::InvalidIntegerInput
    (negative value :58-62; more digits than coefficients :64-70; base < 2 through log_ceil :64).  Returns n coefficients."""
    if value < 0:
        raise ValueError("The given value needs to be non-negative.")
    if base < 2:
        raise ValueError("base < 2")
    min_req_degree = ceil_log(value + 1, base)  # (&value + Z::ONE).log_ceil(&base)
    if min_req_degree > n:
        raise ValueError("not enough coefficients")
    base_repr = []
    while value > 0:  # :73-78
        base_repr.append(value % base)
        value //= base
    q_div_base = q // base  # :87
    coeffs = [0] * n
    for i, digit in enumerate(base_repr):
        coeffs[i] = (digit * q_div_base) % q
    return coeffs


def decode_value_from_polynomialringzq(coeffs, base: int, q: int) -> int:
    """common_encodings.rs:125-153: mu_i = floor((c_i * base + floor(q / (2 base))) / q) mod base on the least
    non-negative representatives, value = sum mu_i base^i (Horner from the top coefficient, :143-150)."""
    if base <= 1:
        raise ValueError("base < 2")
    q_div_2base = q // (2 * base)
    out = 0
    for c in reversed([int(x) % q for x in coeffs]):
        res = ((c * base + q_div_2base) // q) % base
        out = out * base + res
    return out


# --------------------------------------------------------------------------
# Discrete Gaussians (qfall-math behaviour restated; CONTRIBUTING.md:35-45)
# --------------------------------------------------------------------------


def dgauss_pmf(s: float, c: float, lo: int = None, hi: int = None):
    """Exact-law pmf of D_{Z,s,c}, rho(x)=exp(-pi (x-c)^2/s^2), cut to
    [c - ceil(6s), c + floor(6s)] as the reference does.  Returns (support, p)."""
    lo = math.floor(c - math.ceil(6 * s)) if lo is None else lo
    hi = math.ceil(c + math.floor(6 * s)) if hi is None else hi
    xs = np.arange(lo, hi + 1)
    w = np.exp(-math.pi * (xs - c) ** 2 / (s * s))
    return xs, w / w.sum()


def sample_z(rng: np.random.Generator, s: float, c: float) -> int:
    """GPV08 SampleZ as qfall-math implements it: uniform proposal on the cut
    interval, accept with probability rho_{s,c}(x)."""
    lo = math.ceil(c - math.ceil(6 * s))
    hi = math.floor(c + math.floor(6 * s))
    while True:
        x = int(rng.integers(lo, hi + 1))
        if rng.random() < math.exp(-math.pi * (x - c) ** 2 / (s * s)):
            return x


def sample_d_precomputed_gso(rng, basis: np.ndarray, gso: np.ndarray, center: np.ndarray, s: float):
    """GPV08 SampleD (MatZ::sample_d_precomputed_gso): i = cols-1 .. 0,
    c' = <c, b~_i>/<b~_i, b~_i>, s' = s/||b~_i||, z <- D_{Z,s',c'},
    c -= z b_i, v += z b_i; returns v (a lattice vector close to `center`)."""
    b = np.asarray(basis, dtype=np.float64)
    g = np.asarray(gso, dtype=np.float64)
    c = np.asarray(center, dtype=np.float64).copy()
    v = np.zeros(b.shape[0], dtype=object)
    bi = np.array(basis, dtype=object)
    for i in range(b.shape[1] - 1, -1, -1):
        nrm2 = float(g[:, i] @ g[:, i])
        cp = float(c @ g[:, i]) / nrm2
        sp = s / math.sqrt(nrm2)
        z = sample_z(rng, sp, cp)
        c -= z * b[:, i]
        v = v + z * bi[:, i]
    return [int(x) for x in v]


def sample_d_common_non_spherical(rng, sqrt_sigma_2: np.ndarray, r: float):
    """Peikert10 Alg. 1 with B_1 = I (MatZ::sample_d_common_non_spherical):
    x2 = sqrt(Sigma_2) * N(0,1)^m, then p_i <- D_{Z, r, x2[i]}."""
    m = sqrt_sigma_2.shape[0]
    x2 = sqrt_sigma_2 @ rng.standard_normal(m)
    return [sample_z(rng, r, float(c)) for c in x2]


def solve_unit_pivots(a, u, q):
    """Some solution of A x = u over Z_q with free variables 0
    (MatZq::solve_gaussian_elimination, gpv.rs:153-156).  Returns x or None."""
    n, m = len(a), len(a[0])
    aug = [[x % q for x in a[i]] + [u[i] % q] for i in range(n)]
    piv_cols = []
    row = 0
    for c in range(m):
        if row == n:
            break
        pr = None
        for r_ in range(row, n):
            if math.gcd(aug[r_][c], q) == 1:
                pr = r_
                break
        if pr is None:
            continue
        aug[row], aug[pr] = aug[pr], aug[row]
        inv = pow(aug[row][c], -1, q)
        aug[row] = [(x * inv) % q for x in aug[row]]
        for r_ in range(n):
            if r_ != row and aug[r_][c]:
                f = aug[r_][c]
                aug[r_] = [(x - f * y) % q for x, y in zip(aug[r_], aug[row])]
        piv_cols.append(c)
        row += 1
    for r_ in range(row, n):
        if aug[r_][m] % q:
            return None
    x = [0] * m
    for i, c in enumerate(piv_cols):
        x[c] = aug[i][m]
    return x


# --------------------------------------------------------------------------
# samp_p restatements (reference loop structure, float64 for the rationals)
# --------------------------------------------------------------------------


def randomized_nearest_plane_gadget(rng, v, p: GadgetParameters, r: float, s_basis, s_gso):
    """mp_perturbation.rs:173-191."""
    s_g = r * math.sqrt(p.base**2 + 1)
    x0 = [row[0] for row in find_solution_gadget_mat([[x] for x in v], p.q, p.k, p.base)]
    center = -np.array(x0, dtype=np.float64)
    d = sample_d_precomputed_gso(rng, np.array(s_basis, dtype=np.float64), s_gso, center, s_g)
    return [a + b for a, b in zip(x0, d)]


def samp_p_perturbation(rng, p: GadgetParameters, a, r_mat, sqrt_sigma_2, s_basis, s_gso, u, r: float):
    """mp_perturbation.rs:304-336."""
    vec_p = sample_d_common_non_spherical(rng, sqrt_sigma_2, r)
    ap = f_a_classical(a, vec_p, p.q)
    v = [(ui - x) % p.q for ui, x in zip(u, ap)]
    z = randomized_nearest_plane_gadget(rng, v, p, r, s_basis, s_gso)
    rz = np.array(r_mat, dtype=object).dot(np.array(z, dtype=object))
    top = [int(x) + int(y) for x, y in zip(vec_p[: p.m_bar], rz)]
    bot = [int(x) + int(y) for x, y in zip(vec_p[p.m_bar :], z)]
    return top + bot


def samp_p_gpv(rng, a, q, short_base, short_base_gso, u, s: float):
    """gpv.rs:152-161."""
    sol = solve_unit_pivots(a, u, q)
    assert sol is not None
    center = -np.array(sol, dtype=np.float64)
    d = sample_d_precomputed_gso(rng, np.array(short_base, dtype=np.float64), short_base_gso, center, s)
    return [x + y for x, y in zip(sol, d)]


def samp_p_gpv_ring(rng, p: GadgetParametersRing, a, r, e, u, s: float):
    """gpv_ring.rs:160-212: rebuild the basis, embed, solve rot^-(a) x = u, SampleD.
    Returns k+2 polynomials (length-n coefficient lists)."""
    n = p.n
    basis = gen_short_basis_for_trapdoor_ring(p, a, r, e)
    emb = coeff_embed(basis, n)
    gso = gso_f64(np.array(emb, dtype=np.float64))
    a_emb = coeff_embed([[poly_reduce_anticyclic(x, n, p.q) for x in a]], n)
    rot_a = rot_minus_matrix(a_emb)
    sol = solve_unit_pivots(rot_a, poly_reduce_anticyclic(u, n, p.q), p.q)
    assert sol is not None
    center = -np.array(sol, dtype=np.float64)
    d = sample_d_precomputed_gso(rng, np.array(emb, dtype=np.float64), gso, center, s)
    full = [x + y for x, y in zip(sol, d)]
    return [full[i * n : (i + 1) * n] for i in range(p.k + 2)]
