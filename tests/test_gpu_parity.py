"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bit-exact for every deterministic function; for the samplers: A e = u exactly,
check_domain, and chi-square / moment tests against the exact law D_{Z,s,c}
(rho(x) = exp(-pi (x-c)^2 / s^2), CONTRIBUTING.md:35-45) and the oracle's reference sampler."""
import ctypes

import numpy as np
import pytest

from oracle import qfall_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import tools_b200

    return tools_b200


def _chi2_ok(counts, probs, n):
    keep = probs * n >= 8
    exp = probs[keep] * n
    chi = ((counts[keep] - exp) ** 2 / exp).sum()
    # lump the rest
    rest_e = n - exp.sum()
    rest_o = n - counts[keep].sum()
    dof = keep.sum() - 1
    if rest_e >= 8:
        chi += (rest_o - rest_e) ** 2 / rest_e
        dof += 1
    return chi < dof + 5.0 * np.sqrt(2.0 * dof), chi, dof


# ----------------------------------------------------------------------------------
# FIPS 203 compression: bit-exact (lossy_compression_fips203.rs:101-111, 159-169)
# ----------------------------------------------------------------------------------


@pytest.mark.parametrize("q", [257, 3329])
def test_compress_exhaustive(T, q):
    x = np.arange(q, dtype=np.uint16)
    for d in range(1, 13):
        c = T.lossy_compress(x, d, q)
        assert c.dtype == np.uint16
        assert np.array_equal(c.astype(np.uint64), O.lossy_compress_np(x, d, q))
        back = T.lossy_decompress(c, d, q)
        assert np.array_equal(back.astype(np.uint64), O.lossy_decompress_np(c, d, q))
        # int64 (FLINT word) path
        c64 = T.lossy_compress(x.astype(np.int64), d, q)
        assert np.array_equal(c64.astype(np.uint64), O.lossy_compress_np(x, d, q))
        assert np.array_equal(T.lossy_decompress(c64, d, q).astype(np.uint64), O.lossy_decompress_np(c, d, q))
        # round-trip bound (lossy_compression_fips203.rs:281-326)
        dist = np.abs(back.astype(np.int64) - x.astype(np.int64))
        dist = np.minimum(dist, q - dist)
        assert dist.max() <= 2 ** max(O.ceil_log(q, 2) - d - 1, 0)


@pytest.mark.parametrize("count", [0, 1, 7, 8, 9, 255, 256 * 16 + 3, 1_000_003])
def test_compress_ragged_sizes(T, count):
    rng = np.random.default_rng(count)
    x = rng.integers(0, 3329, count).astype(np.uint16)
    for d in (1, 4, 10, 11):
        c = T.lossy_compress(x, d, 3329)
        assert np.array_equal(c.astype(np.uint64), O.lossy_compress_np(x, d, 3329))
        assert np.array_equal(T.lossy_decompress(c, d, 3329).astype(np.uint64), O.lossy_decompress_np(c, d, 3329))


def test_compress_matrix_shape_and_d0(T):
    rng = np.random.default_rng(2)
    mat = rng.integers(0, 3329, (2, 3, 16)).astype(np.uint16)  # MatPolynomialRingZq 2x3, degree 16
    c = T.lossy_compress(mat, 11, 3329)
    assert c.shape == mat.shape
    with pytest.raises(AssertionError):
        T.lossy_compress(mat, 0, 3329)  # lossy_compression_fips203.rs:329-338 should_panic
    with pytest.raises(AssertionError):
        T.lossy_decompress(mat, 0, 3329)


def test_compress_wide_modulus(T):
    rng = np.random.default_rng(9)
    q = 2**61 - 1
    x = rng.integers(0, q, 5000, dtype=np.int64)
    for d in (1, 13, 40, 61):
        want = np.array([O.compress_coeff(int(v), d, q) for v in x], dtype=np.int64)
        got = T.lossy_compress(x, d, q)
        assert np.array_equal(got, want)
        wantd = np.array([O.decompress_coeff(int(v), d, q) for v in want], dtype=np.int64)
        assert np.array_equal(T.lossy_decompress(got, d, q), wantd)


@pytest.mark.parametrize("npoly", [1, 2, 7, 8, 9, 1185, 4099])
def test_byte_encode_decode_bit_exact(T, npoly):
    """Compress_d + ByteEncode_d and ByteDecode_d + Decompress_d (FIPS 203 Algorithms 5 / 6) against
    the oracle's literal restatement, every d, ragged polynomial counts (not a multiple of the warps per CTA)."""
    from tools_b200.compression import compress_encode, decode_decompress

    q = 3329
    rng = np.random.default_rng(npoly)
    x = rng.integers(0, q, (npoly, 256)).astype(np.uint16)
    x[0, :4] = [0, q - 1, q // 2, q // 2 + 1]
    for d in range(1, 13):
        comp = O.lossy_compress_np(x, d, q).astype(np.uint16)
        want = O.byte_encode_np(comp, d)
        got = compress_encode(x, d, q)
        assert got.shape == (npoly, 32 * d) and got.dtype == np.uint8
        assert np.array_equal(got, want), d
        back = decode_decompress(got, d, q)
        assert np.array_equal(back.astype(np.uint64), O.lossy_decompress_np(comp, d, q)), d
        # plain ByteEncode_d / ByteDecode_d (no compression): values mod 2^d, d = 12 -> values < q
        raw = x if d == 12 else (x & ((1 << d) - 1)).astype(np.uint16)
        enc = compress_encode(raw, d, q, compress=False)
        assert np.array_equal(enc, O.byte_encode_np(raw, d))
        assert np.array_equal(decode_decompress(enc, d, q, decompress=False), O.byte_decode_np(enc, d, q))
    # literal (bit-by-bit) oracle on the first polynomial
    for d in (1, 4, 5, 10, 11, 12):
        comp0 = O.lossy_compress(x[0].tolist(), d, q) if d < 12 else x[0].tolist()
        assert bytes(compress_encode(x[:1], d, q, compress=d < 12)[0]) == O.byte_encode(comp0, d)
    # ByteDecode_12 reduces mod q (Algorithm 6, m = q)
    allones = np.full((1, 384), 0xFF, dtype=np.uint8)
    assert np.array_equal(decode_decompress(allones, 12, q, decompress=False), np.full((1, 256), 4095 % q, dtype=np.uint16))
    with pytest.raises(AssertionError):
        compress_encode(x, 0, q)


# ----------------------------------------------------------------------------------
# TrapGen: bit-exact (gadget_classical.rs:56-68)
# ----------------------------------------------------------------------------------


@pytest.mark.parametrize("n,q,with_tag", [(5, 32, False), (8, 64, False), (42, 32, True), (16, 2**24 - 3, False),
                                          (12, 2**31 - 1, True), (6, 2**40 + 15, False)])
def test_trap_gen_from_bit_exact(T, n, q, with_tag):
    from tools_b200.psf import gen_trapdoor

    rng = np.random.default_rng(n)
    gp, po = T.GadgetParameters.init_default(n, q), O.GadgetParameters.init_default(n, q)
    a_bar = rng.integers(0, q, (n, gp.m_bar), dtype=np.int64)
    r = np.array(O.sample_pm_one_zero(rng, gp.m_bar, n * gp.k), dtype=np.int8)
    tag = O.mat_identity(n)
    if with_tag:
        for i in range(n):
            for j in range(i + 1, n):
                tag[i][j] = int(rng.integers(0, q))
    want = O.gen_trapdoor(po, a_bar.tolist(), tag, r.tolist())
    got = gen_trapdoor(gp, a_bar, r, np.array(tag, dtype=np.int64) if with_tag else None)
    assert got.tolist() == want
    # A [R; I] = H G   (gadget_classical.rs:362-414)
    td = np.vstack([r.astype(object), np.eye(n * gp.k, dtype=object)])
    lhs = (got.astype(object).dot(td)) % q
    rhs = np.array(O.mat_mul(tag, O.gen_gadget_mat(n, gp.k, 2), q), dtype=object)
    assert np.array_equal(lhs, rhs)


# ----------------------------------------------------------------------------------
# f_a / check_domain: bit-exact
# ----------------------------------------------------------------------------------


@pytest.mark.parametrize("n,q,s,r", [(8, 64, 25.0, 3.0), (5, 256, 25.0, 2.3219280948873622), (24, 2**24 - 3, 300.0, 4.0),
                                     (16, 2**31 - 1, 4000.0, 2.0), (3, 2**45 + 59, 50.0, 1.5)])
def test_f_a_classical_bit_exact(T, n, q, s, r):
    rng = np.random.default_rng(n + 1)
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFPerturbation(gp, r, s)
    a = rng.integers(0, q, (n, gp.m), dtype=np.int64)
    B = 300
    sig = psf.samp_d_batch(B, seed=5)
    assert sig.shape == (B, gp.m) and sig.dtype == np.int32
    u, flags = psf.f_a_batch(a, sig)
    assert flags.all()
    want = O.f_a_classical_batch(a, sig, q)
    assert np.array_equal(u, want)
    # single-target trait call == A * sigma (mp_perturbation.rs:451-464)
    assert psf.f_a(a, sig[0]).tolist() == O.f_a_classical(a, sig[0], q)
    # extreme in-domain vector: one huge entry (norm exactly at the bound)
    big = np.zeros((2, gp.m), dtype=np.int32)
    big[0, 3] = int(np.floor(s * r * np.sqrt(gp.m)))
    big[1, gp.m - 1] = -int(np.floor(s * r * np.sqrt(gp.m)))
    u2, f2 = psf.f_a_batch(a, big)
    assert f2.all() and np.array_equal(u2, O.f_a_classical_batch(a, big, q))
    # the same huge entry inside an ordinary batch (the fused kernel's optimistic narrow-digit launch must notice it
    # and re-run at full width for every row)
    mixed = sig.copy()
    mixed[17] = 0
    mixed[17, 5] = big[0, 3]
    mixed[B - 1] = 0
    mixed[B - 1, gp.m - 2] = -big[0, 3]
    u3, f3 = psf.f_a_batch(a, mixed)
    assert f3.all() and np.array_equal(u3, O.f_a_classical_batch(a, mixed, q))


def _exact_f_a_int64(a, sig, q):
    """A sigma mod q in exact int64 arithmetic: A split into 16-bit halves so that no partial sum leaves int64
    (|sigma| < 2^16, m < 2^15).  Tied to the oracle on the first rows by the caller."""
    a = np.asarray(a, dtype=np.int64)
    sg = np.asarray(sig, dtype=np.int64)
    lo, hi = a & 0xFFFF, a >> 16
    out = ((sg @ lo.T) % q).astype(object)
    sh = 1
    for _ in range(3):  # hi may itself have up to 46 bits: peel 16 bits at a time
        sh = (sh << 16) % q
        part, hi = hi & 0xFFFF, hi >> 16
        out = (out + ((sg @ part.T) % q).astype(object) * sh) % q  # B x n Python integers: no overflow
    return out.astype(np.int64)


@pytest.mark.parametrize("n,q,s,r,B", [(160, 2**24 - 3, 300.0, 4.0, 300), (300, 2**32 - 5, 467.0, 9.0, 150),
                                         (256, 3329, 40.0, 2.0, 515), (192, 2**16, 200.0, 3.0, 260)])
def test_f_a_fused_cta_pairs(T, n, q, s, r, B, monkeypatch):
    """Shapes with an even number of coordinate tiles: the fused f_a kernel runs as clusters of two CTAs that share the
    converted sigma tiles through distributed shared memory.  Bit-exact against the oracle, ragged batch (B not a
    multiple of 128) and ragged contraction length (m not a multiple of 128), identical to the unpaired kernel,
    norms (check_domain) from both halves of the pair."""
    rng = np.random.default_rng(n)
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFPerturbation(gp, r, s)
    a = rng.integers(0, q, (n, gp.m), dtype=np.int64)
    sig = psf.samp_d_batch(B, seed=11)
    sig[B // 2] = 0  # one out-of-domain row in the second half of a tile, one in the first
    sig[B // 2, 1] = int(s * r * gp.m)
    sig[3] = 0
    sig[3, gp.m - 1] = -int(s * r * gp.m)
    want = _exact_f_a_int64(a, sig, q)
    assert np.array_equal(want[:3], O.f_a_classical_batch(a, sig[:3], q))
    want_flags = np.array([O.check_domain_perturbation(row.tolist(), gp.m, s, r) for row in sig])
    assert not want_flags[3] and not want_flags[B // 2] and want_flags.sum() == B - 2
    monkeypatch.delenv("QF_FA_PAIR", raising=False)
    u, flags = psf.f_a_batch(a, sig, strict=False)
    assert np.array_equal(flags, want_flags)
    assert np.array_equal(u[want_flags], want[want_flags])  # (the value of an out-of-domain row is unspecified)
    monkeypatch.setenv("QF_FA_PAIR", "0")
    u0, flags0 = psf.f_a_batch(a, sig, strict=False)
    assert np.array_equal(u0[want_flags], want[want_flags]) and np.array_equal(flags0, want_flags)


def test_f_a_domain_errors(T):
    # gpv.rs:287-368 / mp_perturbation.rs:466-554
    gp = T.GadgetParameters.init_default(8, 128)
    psf = T.PSFGPV(gp, 10.0)
    rng = np.random.default_rng(0)
    a = rng.integers(0, 128, (8, gp.m), dtype=np.int64)
    m = gp.m
    with pytest.raises(AssertionError):
        psf.f_a(a, np.zeros((m, 2), dtype=np.int64))  # matrix
    with pytest.raises(AssertionError):
        psf.f_a(a, np.zeros(m - 1, dtype=np.int64))  # wrong length
    too_long = np.zeros(m, dtype=np.int64)
    too_long[0] = round(10.0) * m
    with pytest.raises(AssertionError):
        psf.f_a(a, too_long)
    value = round(10.0)
    assert psf.check_domain(np.zeros(m, dtype=np.int64))
    assert psf.check_domain(np.full(m, value, dtype=np.int64))
    assert psf.check_domain(np.full((m, 1), value, dtype=np.int64))
    assert not psf.check_domain(np.zeros((m, 2), dtype=np.int64))
    assert not psf.check_domain(np.zeros(m + 1, dtype=np.int64))
    assert not psf.check_domain(np.zeros(m - 1, dtype=np.int64))
    assert not psf.check_domain(too_long)
    # oracle agrees on random vectors around the boundary
    for scale in (5, 9, 10, 11, 14):
        v = rng.integers(-scale, scale + 1, m)
        assert psf.check_domain(v) == O.check_domain_gpv(v.tolist(), m, 10.0)
    # batch flags: mixed
    sig = np.zeros((3, m), dtype=np.int32)
    sig[1, 0] = value * m
    u, flags = psf.f_a_batch(a, sig, strict=False)
    assert flags.tolist() == [True, False, True]
    with pytest.raises(AssertionError):
        psf.f_a_batch(a, sig)


@pytest.mark.parametrize("path", ["fused", "unfused_i8", "fp64"])
def test_check_domain_norm_does_not_wrap(T, path, monkeypatch):
    """||sigma||^2 is exact (gpv.rs:219-224): a 64-bit sum of int32 squares wraps (16 entries of 2^30 sum to 2^64 = 0,
    four entries of INT32_MIN likewise), which would report a vector of norm 2^32 as in-domain.  Every norm kernel
    saturates instead."""
    if path == "unfused_i8":
        monkeypatch.setenv("QF_DISABLE_FUSED_FA", "1")
    elif path == "fp64":
        monkeypatch.setenv("QF_DISABLE_I8", "1")
    n, q = 8, 64
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFPerturbation(gp, 3.0, 25.0)
    a, _ = psf.trap_gen(seed=1)
    sig = psf.samp_d_batch(6, seed=2)
    sig[1, :16] = 2**30
    sig[2, 3:7] = -(2**31)
    sig[3, :64] = 2**29           # 64 * 2^58 = 2^64
    sig[4, 5] = 2**31 - 1         # a single huge entry (no wrap, plainly out of domain)
    flags = psf.check_domain_batch(sig)
    assert flags.tolist() == [True, False, False, False, False, True]
    u, fl = psf.f_a_batch(a, sig, strict=False)
    assert fl.tolist() == [True, False, False, False, False, True]
    assert np.array_equal(u[[0, 5]], O.f_a_classical_batch(a, sig[[0, 5]], q))
    with pytest.raises(T.NotInDomain):
        psf.f_a_batch(a, sig)
    for row in (1, 2, 3, 4):
        assert not psf.check_domain(sig[row])
        with pytest.raises(T.NotInDomain):
            psf.f_a(a, sig[row])
    # wider-than-int32 inputs must not wrap in the host cast either (2^32 -> 0)
    big = sig.astype(np.int64)
    big[0, 0] = 2**32
    assert psf.check_domain_batch(big).tolist() == [False, False, False, False, False, True]
    _, fl = psf.f_a_batch(a, big, strict=False)
    assert fl.tolist() == [False, False, False, False, False, True]
    obj = sig[:1].astype(object)
    obj[0, 2] = 2**70
    assert not psf.check_domain_batch(obj)[0] and not psf.check_domain(obj[0])


@pytest.mark.parametrize("path", ["dense", "small_ntt", "goldilocks", "schoolbook"])
def test_ring_check_domain_norm_does_not_wrap(T, path, monkeypatch):
    n, q = (64, 3329) if path != "schoolbook" else (6, 128)
    if path in ("small_ntt", "goldilocks"):
        monkeypatch.setenv("QF_DISABLE_RING_DENSE", "1")
    if path == "goldilocks":
        monkeypatch.setenv("QF_DISABLE_SMALL_NTT", "1")
    gr = T.GadgetParametersRing.init_default(n, q)
    pr = T.PSFGPVRing(gr, 300.0, 1.005)
    a, _ = pr.trap_gen(seed=6)
    sig = pr.samp_d_batch(4, seed=7)
    flat = sig.reshape(4, -1)
    flat[1, :16] = 2**30
    flat[2, 1:5] = -(2**31)
    u, fl = pr.f_a_batch(a, sig, strict=False)
    assert fl.tolist() == [True, False, False, True]
    assert u[0].tolist() == O.f_a_ring(a.tolist(), sig[0].tolist(), n, q)
    assert pr.check_domain_batch(sig).tolist() == [True, False, False, True]


def test_key_change_invalidates_trapdoor(T):
    """C ABI contract: installing a new key drops the trapdoor installed for the old one -- samp_p answers QF_ERR_NO_KEY
    instead of combining the new A with the old A^-1 / pivots / S / R (which would return QF_OK and A e != u)."""
    from tools_b200 import _ffi

    n, q = 8, 64
    gp = T.GadgetParameters.init_default(n, q)
    for make in (lambda: T.PSFGPV(gp, 90.0), lambda: T.PSFPerturbation(gp, 3.0, 25.0)):
        psf = make()
        a, td = psf.trap_gen(seed=4)
        u = np.random.default_rng(1).integers(0, q, (4, n), dtype=np.int64)
        e = psf.samp_p_batch(a, td, u, seed=5)
        assert np.array_equal(O.f_a_classical_batch(a, e, q), u)
        a2 = np.ascontiguousarray(np.roll(a, 1, axis=0))
        psf.ctx.call("qf_set_a", _ffi.ptr(a2))  # behind the wrapper's back
        out = np.empty_like(e)
        st = psf.ctx.status("qf_samp_p", _ffi.ptr(u), 4, 5, 0, _ffi.ptr(out))
        assert st == _ffi.QF_ERR_NO_KEY
    gr = T.GadgetParametersRing.init_default(8, 1024)
    pr = T.PSFGPVRing(gr, 300.0, 1.005)
    ar, tdr = pr.trap_gen(seed=1)
    ur = np.random.default_rng(2).integers(0, 1024, (3, 8), dtype=np.int64)
    er = pr.samp_p_batch(ar, tdr, ur, seed=3)
    pr.ctx.call("qf_ring_set_a", _ffi.ptr(np.ascontiguousarray(ar)))
    st = pr.ctx.status("qf_samp_p", _ffi.ptr(ur), 3, 3, 0, _ffi.ptr(np.empty_like(er)))
    assert st == _ffi.QF_ERR_NO_KEY


def test_int16_domain_form_and_device_range_check(T):
    """Narrow boundary types: qf_samp_p_i16 returns the same preimages as qf_samp_p (same seed) in 16 bits, qf_f_a_i16
    accepts them; a target outside [0, q) is caught by the device-side range check (QF_ERR_INVALID); an entry that does
    not fit int16 is reported, not truncated."""
    from tools_b200 import _ffi

    n, q = 8, 64
    gp = T.GadgetParameters.init_default(n, q)
    rng = np.random.default_rng(3)
    u = rng.integers(0, q, (777, n), dtype=np.int64)
    for psf in (T.PSFGPV(gp, 90.0), T.PSFPerturbation(gp, 3.0, 25.0)):
        a, td = psf.trap_gen(seed=4)
        psf.ctx.call("qf_set_chunk", 256)  # several chunks: the overlapped copy path with its two int16 buffers
        e32 = psf.samp_p_batch(a, td, u, seed=5)
        e16 = psf.samp_p_batch(a, td, u, seed=5, dtype=np.int16)
        assert e16.dtype == np.int16 and np.array_equal(e16.astype(np.int32), e32)
        u16, fl = psf.f_a_batch(a, e16)
        assert np.array_equal(u16, u) and fl.all()
        bad = u.copy()
        bad[500, 3] = q
        with pytest.raises(T.QfError) as ei:
            psf.samp_p_batch(a, td, bad, seed=5)
        assert ei.value.status == _ffi.QF_ERR_INVALID
        bad[500, 3] = -1
        with pytest.raises(T.QfError) as ei:
            psf.samp_p_batch(a, td, bad, seed=5, dtype=np.int16)
        assert ei.value.status == _ffi.QF_ERR_INVALID
        assert np.array_equal(psf.samp_p_batch(a, td, u, seed=5), e32)  # the context is usable again
    # s r so large that entries exceed int16
    wide = T.PSFPerturbation(gp, 300.0, 60.0)
    a, td = wide.trap_gen(seed=1)
    with pytest.raises(T.QfError) as ei:
        wide.samp_p_batch(a, td, u, seed=2, dtype=np.int16)
    assert ei.value.status == _ffi.QF_ERR_NUMERIC
    e = wide.samp_p_batch(a, td, u, seed=2)
    assert np.abs(e).max() > 32767 and np.array_equal(O.f_a_classical_batch(a, e, q), u)
    # ring
    gr = T.GadgetParametersRing.init_default(8, 1024)
    pr = T.PSFGPVRing(gr, 300.0, 1.005)
    ar, tdr = pr.trap_gen(seed=1)
    ur = rng.integers(0, 1024, (50, 8), dtype=np.int64)
    er = pr.samp_p_batch(ar, tdr, ur, seed=3)
    er16 = pr.samp_p_batch(ar, tdr, ur, seed=3, dtype=np.int16)
    assert np.array_equal(er16.astype(np.int32), er)
    uu, fl = pr.f_a_batch(ar, er16)
    assert np.array_equal(uu, ur) and fl.all()


# ----------------------------------------------------------------------------------
# samplers: exact law
# ----------------------------------------------------------------------------------


@pytest.mark.parametrize("s,c", [(3.0, 0.0), (3.0, 0.37), (7.5, -12.5), (1.8, 1e6 + 0.25), (20.0, 3.999),
                                 (4.0 * np.sqrt(5), -0.5), (1000.0, 7.3), (60000.0, 123456.789)])
def test_sample_z_law(T, s, c):
    from tools_b200 import _ffi

    n = 400_000
    centers = np.full(n, c, dtype=np.float64)
    out = np.empty(n, dtype=np.int64)
    st = _ffi.lib().qf_sample_z(_ffi.ptr(centers), n, s, 1234, _ffi.ptr(out))
    assert st == 0
    if s < 100:
        xs, pm = O.dgauss_pmf(s, c)
        cnt = np.array([(out == x).sum() for x in xs])
        assert cnt.sum() == n, "sample outside the 6 s tail cut"
        ok, chi, dof = _chi2_ok(cnt, pm, n)
        assert ok, (chi, dof)
    sigma = s / np.sqrt(2 * np.pi)
    xs, pm = O.dgauss_pmf(s, c)
    mean = (xs * pm).sum()
    var = ((xs - mean) ** 2 * pm).sum()
    assert abs(out.mean() - mean) < 5 * sigma / np.sqrt(n)
    assert abs(out.var() - var) < 6 * var * np.sqrt(2.0 / n)
    # fourth moment (kurtosis of a Gaussian is 3)
    m4 = ((out - out.mean()) ** 4).mean() / out.var() ** 2
    assert abs(m4 - ((xs - mean) ** 4 * pm).sum() / var**2) < 0.06


def test_sample_z_matches_reference_sampler(T):
    """Two-sample chi-square between the CUDA sampler and the oracle's restatement of the
    reference SampleZ (uniform proposal + rejection) at the same (s, c)."""
    from tools_b200 import _ffi

    s, c, n = 3.0 * np.sqrt(5), 0.3, 200_000
    out = np.empty(n, dtype=np.int64)
    centers = np.full(n, c, dtype=np.float64)
    assert _ffi.lib().qf_sample_z(_ffi.ptr(centers), n, s, 99, _ffi.ptr(out)) == 0
    rng = np.random.default_rng(5)
    # vectorised restatement of O.sample_z (same proposal / acceptance rule)
    lo, hi = int(np.ceil(c - np.ceil(6 * s))), int(np.floor(c + np.floor(6 * s)))
    ref = []
    while len(ref) < n:
        x = rng.integers(lo, hi + 1, 4 * n)
        acc = rng.random(4 * n) < np.exp(-np.pi * (x - c) ** 2 / (s * s))
        ref.extend(x[acc].tolist())
    ref = np.array(ref[:n])
    xs = np.arange(lo, hi + 1)
    a = np.array([(out == x).sum() for x in xs], dtype=np.float64)
    b = np.array([(ref == x).sum() for x in xs], dtype=np.float64)
    keep = (a + b) >= 20
    chi = (((a - b) ** 2)[keep] / (a + b)[keep]).sum()
    dof = keep.sum() - 1
    assert chi < dof + 5 * np.sqrt(2 * dof), (chi, dof)


def test_samp_d(T):
    # gpv.rs:238-249, mp_perturbation.rs:416-428, gpv_ring.rs:302-313
    for n, q in [(5, 256), (10, 128), (15, 157)]:
        gp = T.GadgetParameters.init_default(n, q)
        psf = T.PSFGPV(gp, 10.0)
        d = psf.samp_d_batch(50, seed=n)
        assert psf.check_domain_batch(d).all()
        assert psf.check_domain(psf.samp_d())
        pp = T.PSFPerturbation(gp, float(np.log2(n)), 25.0)
        assert pp.check_domain_batch(pp.samp_d_batch(50, seed=n)).all()
    gr = T.GadgetParametersRing.init_default(5, 123456789)
    pr = T.PSFGPVRing(gr, 1000.0, 1.005)
    d = pr.samp_d_batch(20, seed=3)
    assert d.shape == (20, gr.k + 2, 5)
    assert pr.check_domain_batch(d).all() and pr.check_domain(pr.samp_d())
    # law of the coordinates: D_{Z, s} with s = 10
    gp = T.GadgetParameters.init_default(8, 64)
    psf = T.PSFGPV(gp, 10.0)
    flat = psf.samp_d_batch(4000, seed=1).reshape(-1)
    xs, pm = O.dgauss_pmf(10.0, 0.0)
    cnt = np.array([(flat == x).sum() for x in xs])
    ok, chi, dof = _chi2_ok(cnt, pm, flat.size)
    assert ok, (chi, dof)
    # determinism and batch splitting
    a = psf.samp_d_batch(64, seed=7)
    b = np.concatenate([psf.samp_d_batch(40, seed=7), psf.samp_d_batch(24, seed=7, first_index=40)])
    assert np.array_equal(a, b)
    assert not np.array_equal(a, psf.samp_d_batch(64, seed=8))


# ----------------------------------------------------------------------------------
# PSFPerturbation::samp_p (mp_perturbation.rs:304-336)
# ----------------------------------------------------------------------------------


def _pert_setup(T, n, q, r, s, seed=1):
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFPerturbation(gp, r, s)
    a, td = psf.trap_gen(seed=seed)
    return gp, psf, a, td


@pytest.mark.parametrize("n,q,r,s", [(8, 64, 3.0, 25.0), (5, 256, float(np.log2(5)), 25.0), (6, 128, float(np.log2(6)), 25.0),
                                     (8, 127, 3.0, 30.0), (16, 2**20 - 3, 4.0, 80.0)])
def test_samp_p_perturbation_preimage_and_domain(T, n, q, r, s):
    gp, psf, a, td = _pert_setup(T, n, q, r, s)
    # the generated key really is a G-trapdoor
    rmat = td[0]
    tdm = np.vstack([rmat.astype(object), np.eye(n * gp.k, dtype=object)])
    assert np.array_equal(a.astype(object).dot(tdm) % q, np.array(O.gen_gadget_mat(n, gp.k, 2), dtype=object) % q)
    rng = np.random.default_rng(3)
    B = 2000
    u = rng.integers(0, q, (B, n), dtype=np.int64)
    e = psf.samp_p_batch(a, td, u, seed=11)
    assert e.shape == (B, gp.m)
    assert np.array_equal(O.f_a_classical_batch(a, e, q), u)  # A e = u exactly, every target
    assert psf.check_domain_batch(e).all()
    u2, flags = psf.f_a_batch(a, e)
    assert np.array_equal(u2, u) and flags.all()
    # README flow (mp_perturbation.rs:432-448): trait-shaped single calls
    ds = psf.samp_d(seed=4)
    rng_fa = psf.f_a(a, ds)
    pre = psf.samp_p(a, td, rng_fa, seed=5)
    assert psf.check_domain(pre) and psf.f_a(a, pre).tolist() == rng_fa.tolist()
    # determinism / batch splitting / seed sensitivity
    e_again = psf.samp_p_batch(a, td, u, seed=11)
    assert np.array_equal(e, e_again)
    e_split = np.concatenate([psf.samp_p_batch(a, td, u[:700], seed=11),
                              psf.samp_p_batch(a, td, u[700:], seed=11, first_index=700)])
    assert np.array_equal(e, e_split)
    assert not np.array_equal(e, psf.samp_p_batch(a, td, u, seed=12))


def test_samp_p_perturbation_distribution(T):
    """Marginals against the reference algorithm (oracle restatement) at the same s, and against
    the spherical law the construction targets: each coordinate ~ D_{Z, s r}-like with
    variance (s r)^2 / (2 pi)."""
    n, q, r, s = 8, 64, 3.0, 25.0
    gp, psf, a, td = _pert_setup(T, n, q, r, s)
    rmat, l, (sb, sg) = td
    rng = np.random.default_rng(8)
    B = 20000
    u = np.tile(rng.integers(0, q, (1, n), dtype=np.int64), (B, 1))  # one fixed syndrome
    e = psf.samp_p_batch(a, td, u, seed=21).astype(np.float64)
    sigma2 = (s * r) ** 2 / (2 * np.pi)
    var = e.var(axis=0)
    assert np.all(np.abs(var / sigma2 - 1) < 0.08), (var.min() / sigma2, var.max() / sigma2)
    assert np.all(np.abs(e.mean(axis=0)) < 5 * np.sqrt(sigma2 / B) + 0.6)
    # covariance is spherical: off-diagonal correlations vanish
    corr = np.corrcoef(e.T)
    off = corr - np.eye(gp.m)
    assert np.abs(off).max() < 0.05
    # reference sampler on the same key: compare ||e||^2 and a few coordinate variances
    po = O.GadgetParameters.init_default(n, q)
    rr = np.random.default_rng(9)
    ref = np.array([O.samp_p_perturbation(rr, po, a, rmat.tolist(), l, sb.tolist(), sg, u[0].tolist(), r)
                    for _ in range(300)], dtype=np.float64)
    n_ref, n_gpu = (ref**2).sum(1), (e**2).sum(1)
    se = np.sqrt(n_ref.var() / len(n_ref) + n_gpu.var() / len(n_gpu))
    assert abs(n_ref.mean() - n_gpu.mean()) < 5 * se
    assert abs(ref.var() / e.var() - 1) < 0.06


@pytest.mark.parametrize("n,q,r,s", [(8, 64, 3.0, 25.0), (16, 2**10, 4.0, 60.0), (40, 2**12, 5.0, 300.0)])
def test_compute_sqrt_sigma_2_matches_oracle(T, n, q, r, s):
    """compute_sqrt_sigma_2 (mp_perturbation.rs:111-139) through the library's blocked Cholesky against the
    oracle's float64 restatement; fp64 tolerance 1e-9 relative to the entries (|L| <= s r)."""
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFPerturbation(gp, r, s)
    rng = np.random.default_rng(n)
    rm = (rng.integers(0, 2, (gp.m_bar, n * gp.k)) - rng.integers(0, 2, (gp.m_bar, n * gp.k))).astype(np.int8)
    want = O.compute_sqrt_sigma_2(rm, s, r, 2)
    got = psf.compute_sqrt_sigma_2(rm)
    assert got.shape == (gp.m, gp.m) and not np.triu(got, 1).any()
    assert np.allclose(got, want, rtol=0, atol=1e-9 * s * r)
    # a supplied covariance (mp_perturbation.rs:94-104): Sigma = s^2 I + a small symmetric perturbation
    d = rng.normal(size=(gp.m, 3))
    sigma = (s * s) * np.eye(gp.m) + d @ d.T
    t = np.vstack([rm.astype(np.float64), np.eye(n * gp.k)])
    ref = np.linalg.cholesky((r * r) / (2 * np.pi) * (sigma - 5.0 * (t @ t.T) - np.eye(gp.m)))
    assert np.allclose(psf.compute_sqrt_sigma_2(rm, sigma), ref, rtol=0, atol=1e-9 * s * r)
    # Sigma_2 not positive definite: the reference panics in the Cholesky (mp_perturbation.rs:109-110)
    bad = T.PSFPerturbation(gp, r, 3.0)
    with pytest.raises(T.QfError):
        bad.compute_sqrt_sigma_2(rm)


def test_samp_p_perturbation_structured_sqrt(T):
    """sqrt_sigma_2 = None: the backend's block-structured square root of the default Sigma_2.  Same law as the
    dense Cholesky factor: A e = u, check_domain, spherical covariance (s r)^2 / (2 pi), and the same ||e||^2
    distribution as the dense path."""
    n, q, r, s = 8, 64, 3.0, 25.0
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFPerturbation(gp, r, s)
    a, td_dense = psf.trap_gen(seed=1)
    rmat, l, basis = td_dense
    td = (rmat, None, basis)
    rng = np.random.default_rng(8)
    B = 20000
    u1 = rng.integers(0, q, (3000, n), dtype=np.int64)
    e1 = psf.samp_p_batch(a, td, u1, seed=5)
    assert np.array_equal(O.f_a_classical_batch(a, e1, q), u1) and psf.check_domain_batch(e1).all()
    assert np.array_equal(e1, np.concatenate([psf.samp_p_batch(a, td, u1[:1000], seed=5),
                                              psf.samp_p_batch(a, td, u1[1000:], seed=5, first_index=1000)]))
    u = np.tile(rng.integers(0, q, (1, n), dtype=np.int64), (B, 1))
    e = psf.samp_p_batch(a, td, u, seed=21).astype(np.float64)
    sigma2 = (s * r) ** 2 / (2 * np.pi)
    var = e.var(axis=0)
    assert np.all(np.abs(var / sigma2 - 1) < 0.08), (var.min() / sigma2, var.max() / sigma2)
    corr = np.corrcoef(e.T)
    assert np.abs(corr - np.eye(gp.m)).max() < 0.05
    ed = psf.samp_p_batch(a, td_dense, u, seed=22).astype(np.float64)
    n_s, n_d = (e**2).sum(1), (ed**2).sum(1)
    se = np.sqrt(n_s.var() / B + n_d.var() / B)
    assert abs(n_s.mean() - n_d.mean()) < 5 * se
    assert abs(n_s.var() / n_d.var() - 1) < 0.1
    z = (e - e.mean(0)) / e.std(0)
    assert np.abs((z**4).mean(0) - 3).max() < 0.35  # Gaussian marginals


def test_samp_p_perturbation_structured_midsize(T):
    """n = 40, q = 2^12 (m = 996, several Cholesky blocks and tensor-core tiles), structured square root."""
    n, q, r = 40, 2**12, float(np.log2(40))
    gp = T.GadgetParameters.init_default(n, q)
    s = 1.3 * np.sqrt(5 * ((np.sqrt(gp.m_bar) + np.sqrt(n * gp.k)) ** 2 / 2 + 1) + 1)
    psf = T.PSFPerturbation(gp, r, s)
    a, td = psf.trap_gen(seed=3, dense_sqrt_sigma_2=False)
    assert td[1] is None
    rng = np.random.default_rng(2)
    u = rng.integers(0, q, (4096, n), dtype=np.int64)
    e = psf.samp_p_batch(a, td, u, seed=9)
    assert np.array_equal(O.f_a_classical_batch(a, e, q), u) and psf.check_domain_batch(e).all()
    ef = e.astype(np.float64)
    ratio = (ef**2).sum(1).mean() / (gp.m * (s * r) ** 2 / (2 * np.pi))
    assert abs(ratio - 1) < 0.01, ratio
    var = ef.var(axis=0) / ((s * r) ** 2 / (2 * np.pi))
    assert np.all(np.abs(var - 1) < 0.15), (var.min(), var.max())
    c = np.corrcoef(ef[:, ::37].T)
    assert np.abs(c - np.eye(c.shape[0])).max() < 0.1
    # and the dense path on the same key agrees in law
    td_d = (td[0], psf.compute_sqrt_sigma_2(td[0]), td[2])
    ed = psf.samp_p_batch(a, td_d, u, seed=9).astype(np.float64)
    assert np.array_equal(O.f_a_classical_batch(a, ed.astype(np.int64), q), u)
    assert abs((ed**2).sum(1).mean() / (ef**2).sum(1).mean() - 1) < 0.01


def test_gadget_sampler_distribution(T):
    """The gadget part alone: z = e[m_bar:] - p[m_bar:] is not observable, but G z' = v structure
    is: with R = 0 the lower block of e is p_low + z; check A e = u and the conditional law of the
    lower block's parity: for q = 2^k the first digit of each block satisfies z_0 = v_0 mod 2."""
    # covered structurally by the preimage tests; here: many different syndromes, tiny modulus
    n, q, r, s = 4, 16, 3.0, 30.0
    gp, psf, a, td = _pert_setup(T, n, q, r, s, seed=5)
    rng = np.random.default_rng(1)
    u = rng.integers(0, q, (5000, n), dtype=np.int64)
    e = psf.samp_p_batch(a, td, u, seed=2)
    assert np.array_equal(O.f_a_classical_batch(a, e, q), u)
    assert psf.check_domain_batch(e).all()


# ----------------------------------------------------------------------------------
# PSFGPV::samp_p (gpv.rs:152-161)
# ----------------------------------------------------------------------------------


@pytest.mark.parametrize("n,q", [(5, 32), (8, 128), (24, 2**16), (40, 2**12)])
def test_gso_device_matches_oracle(T, n, q):
    """qf_gso (MatQ::gso, gpv.rs:91): blocked Gram-Schmidt on the device against the oracle's exact-rational GSO
    (small) and float64 QR-free restatement; tolerance 1e-9 absolute on entries of size <= ||b_j|| ~ 10^2."""
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFGPV(gp, 10.0 * n)
    a, (sb, sg) = psf.trap_gen(seed=3, dense_gso=True)
    assert sg.shape == sb.shape
    ref = O.gso_f64(sb.astype(np.float64))
    assert np.allclose(sg, ref, rtol=0, atol=1e-8)
    if gp.m <= 130:
        ex = np.array(O.gso_exact(sb.tolist()), dtype=np.float64)
        assert np.allclose(sg, ex, rtol=0, atol=1e-9)
    # orthogonality and the triangular relation S = G U
    gram = sg.T @ sg
    off = gram - np.diag(np.diag(gram))
    assert np.abs(off).max() < 1e-7 * np.diag(gram).max()
    # a trapdoor installed without its GSO (None) samples the same law as with it
    rng = np.random.default_rng(1)
    u = rng.integers(0, q, (500, n), dtype=np.int64)
    e1 = psf.samp_p_batch(a, (sb, sg), u, seed=4)
    e2 = psf.samp_p_batch(a, (sb, None), u, seed=4)
    assert np.array_equal(O.f_a_classical_batch(a, e2, q), u) and psf.check_domain_batch(e2).all()
    assert np.array_equal(e1, e2)  # same GSO (bit-identical path) => same draws


@pytest.mark.parametrize("n,q", [(2, 8), (5, 256), (6, 127), (10, 3329), (24, 2**16), (40, 2**20 - 3), (64, 2**24)])
def test_gen_short_basis_device_bit_exact(T, n, q):
    """gen_short_basis_for_trapdoor (short_basis_classical.rs:54-110) with the R W product on the tensor cores against
    the oracle's big-integer restatement (itself pinned on the reference's golden sa_l / sa_r, test_oracle_golden)."""
    from tools_b200 import gadget

    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFGPV(gp, 10.0 * n)
    a, r = T.psf._trap_gen_classical(psf, seed=n + 7)
    got = psf.gen_short_basis_for_trapdoor(r)
    host = gadget.gen_short_basis_for_trapdoor(gp, a, r)
    assert np.array_equal(got, host)
    if n <= 24:
        po = O.GadgetParameters.init_default(n, q)
        tag = [[int(i == j) for j in range(n)] for i in range(n)]
        want = O.gen_short_basis_for_trapdoor(po, tag, a.tolist(), r.tolist())
        assert got.tolist() == [list(map(int, row)) for row in want]
    assert not (a.astype(object).dot(got.astype(object)) % q).any()  # every column lies in Lambda^perp(A)


@pytest.mark.parametrize("n,q,s", [(5, 256, 10.0), (6, 128, 10.0), (8, 128, 90.0), (10, 127, 40.0), (24, 2**16, 60.0)])
def test_samp_p_gpv_preimage_and_domain(T, n, q, s):
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFGPV(gp, s)
    a, td = psf.trap_gen(seed=n)
    sb, sg = td
    assert not (a.astype(object).dot(sb.astype(object)) % q).any()  # basis of the kernel lattice
    rng = np.random.default_rng(4)
    B = 1000
    u = rng.integers(0, q, (B, n), dtype=np.int64)
    e = psf.samp_p_batch(a, td, u, seed=3)
    assert np.array_equal(O.f_a_classical_batch(a, e, q), u)
    assert psf.check_domain_batch(e).all()
    ds = psf.samp_d(seed=1)
    fa = psf.f_a(a, ds)
    pre = psf.samp_p(a, td, fa, seed=2)
    assert psf.f_a(a, pre).tolist() == fa.tolist() and psf.check_domain(pre)
    assert np.array_equal(e, psf.samp_p_batch(a, td, u, seed=3))
    e_split = np.concatenate([psf.samp_p_batch(a, td, u[:300], seed=3), psf.samp_p_batch(a, td, u[300:], seed=3, first_index=300)])
    assert np.array_equal(e, e_split)


@pytest.mark.parametrize("n,q,s", [(8, 2**16, 120.0), (16, 65521, 200.0), (24, 2**16, 300.0), (32, 2**12 - 3, 300.0)])
def test_samp_p_gpv_structured_basis_matches_dense(T, n, q, s, monkeypatch):
    """A short basis built by gen_short_basis_for_trapdoor has the form [[R S', I + R W],[S', W]]; the backend
    recovers and verifies (R, W) exactly and evaluates e = sol + S z through the two thin factors.  Integer
    arithmetic on the same z: the preimages must be bit-identical to the dense S z path."""
    gp = T.GadgetParameters.init_default(n, q)
    assert (n * gp.k) % 128 == 0
    rng = np.random.default_rng(n)
    u = rng.integers(0, q, (700, n), dtype=np.int64)
    psf = T.PSFGPV(gp, s)
    a, td = psf.trap_gen(seed=11)
    e_struct = psf.samp_p_batch(a, td, u, seed=6)
    assert np.array_equal(O.f_a_classical_batch(a, e_struct, q), u) and psf.check_domain_batch(e_struct).all()
    monkeypatch.setenv("QF_DISABLE_GPV_STRUCT", "1")
    psf2 = T.PSFGPV(gp, s)
    psf2._install_a(a)
    e_dense = psf2.samp_p_batch(a, td, u, seed=6)
    assert np.array_equal(e_struct, e_dense)
    # a basis that is NOT of that form (two columns swapped -- still a basis of the same lattice) takes the dense path
    monkeypatch.delenv("QF_DISABLE_GPV_STRUCT")
    sb = td[0].copy()
    sb[:, [0, 1]] = sb[:, [1, 0]]
    psf3 = T.PSFGPV(gp, s)
    psf3._install_a(a)
    e3 = psf3.samp_p_batch(a, (sb, None), u, seed=6)
    assert np.array_equal(O.f_a_classical_batch(a, e3, q), u) and psf3.check_domain_batch(e3).all()


def test_samp_p_gpv_distribution(T):
    """GPV08 SampleD outputs D_{Lambda_u^perp(A), s}: spherical, variance s^2/(2 pi) per coordinate;
    compared with the oracle's restatement of the reference loop on the same key."""
    n, q, s = 5, 32, 10.0
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFGPV(gp, s)
    a, td = psf.trap_gen(seed=2)
    sb, sg = td
    rng = np.random.default_rng(6)
    B = 20000
    u = np.tile(rng.integers(0, q, (1, n), dtype=np.int64), (B, 1))
    e = psf.samp_p_batch(a, td, u, seed=4).astype(np.float64)
    sigma2 = s * s / (2 * np.pi)
    var = e.var(axis=0)
    assert np.all(np.abs(var / sigma2 - 1) < 0.15), (var.min() / sigma2, var.max() / sigma2)
    rr = np.random.default_rng(1)
    ref = np.array([O.samp_p_gpv(rr, a.tolist(), q, sb.tolist(), sg, u[0].tolist(), s) for _ in range(400)], dtype=np.float64)
    n_ref, n_gpu = (ref**2).sum(1), (e**2).sum(1)
    se = np.sqrt(n_ref.var() / len(n_ref) + n_gpu.var() / len(n_gpu))
    assert abs(n_ref.mean() - n_gpu.mean()) < 5 * se


# ----------------------------------------------------------------------------------
# PSFGPVRing (gpv_ring.rs)
# ----------------------------------------------------------------------------------


def _ring_s(n):
    return ((2 * 2 * 1.005 * np.sqrt(n) + 1) * 2) * 4  # gpv_ring.rs:296-298


@pytest.mark.parametrize("n,q", [(64, 3329), (256, 3329), (128, 7681), (256, 7681), (512, 12289), (1024, 12289),
                                 (64, 257), (64, 2**31 - 1), (256, 2**24), (8, 1024), (5, 256), (6, 128)])
@pytest.mark.parametrize("dense", [True, False])
def test_ring_f_a_bit_exact(T, n, q, dense, monkeypatch):
    """dense=True: rot^-(a) as one tensor-core contraction (n <= 512); dense=False: the NTT / schoolbook kernels."""
    if not dense:
        monkeypatch.setenv("QF_DISABLE_RING_DENSE", "1")
    rng = np.random.default_rng(n)
    gp = T.GadgetParametersRing.init_default(n, q)
    psf = T.PSFGPVRing(gp, float(_ring_s(n)), 1.005)
    a = rng.integers(0, q, (gp.k + 2, n), dtype=np.int64)
    B = 40
    sig = psf.samp_d_batch(B, seed=1)
    u, flags = psf.f_a_batch(a, sig)
    assert flags.all()
    for b in range(0, B, 7):
        assert u[b].tolist() == O.f_a_ring(a.tolist(), sig[b].tolist(), n, q)
    # == rot^-(a) sigma (rotation_matrix.rs): first polynomial only
    rot = np.array(O.rot_minus(a[1].tolist()), dtype=object)
    one = np.zeros((1, gp.k + 2, n), dtype=np.int32)
    one[0, 1] = sig[0, 1]
    u1, _ = psf.f_a_batch(a, one)
    assert u1[0].tolist() == [int(x) % q for x in rot.dot(sig[0, 1].astype(object))]
    # domain edge cases (gpv_ring.rs:355-445)
    assert psf.check_domain(np.zeros((gp.k + 2, n), dtype=np.int64))
    assert not psf.check_domain(np.zeros((gp.k + 1, n), dtype=np.int64))
    assert not psf.check_domain(np.zeros((gp.k + 3, n), dtype=np.int64))
    big = np.zeros((gp.k + 2, n), dtype=np.int64)
    big[0, 0] = round(psf.s) * (gp.k + 2) * 8 * max(1, n // 8)  # the reference uses n = 8 (gpv_ring.rs:436-439)
    assert not psf.check_domain(big)
    with pytest.raises(AssertionError):
        psf.f_a(a, big)


@pytest.mark.parametrize("n,q", [(6, 32), (64, 3329), (8, 2**31 - 1)])
def test_ring_trap_gen_bit_exact(T, n, q):
    rng = np.random.default_rng(q % 1000)
    gp, po = T.GadgetParametersRing.init_default(n, q), O.GadgetParametersRing.init_default(n, q)
    psf = T.PSFGPVRing(gp, float(_ring_s(n)), 1.005)
    a_bar = rng.integers(0, q, n, dtype=np.int64)
    r = rng.integers(-10, 11, (gp.k, n)).astype(np.int32)
    e = rng.integers(-10, 11, (gp.k, n)).astype(np.int32)
    got = psf.gen_trapdoor_ring_lwe(a_bar, r, e)
    want = O.gen_trapdoor_ring_lwe(po, a_bar.tolist(), r.tolist(), e.tolist())
    assert got.tolist() == want



@pytest.mark.parametrize("n,q", [(4, 16), (16, 257), (64, 3329), (32, 2**10)])
def test_ring_short_basis_device_bit_exact(T, n, q):
    """qf_ring_gen_short_basis (short_basis_ring.rs:64-166, coefficient embedding) against the oracle's restatement and
    the host assembly, for a power-of-base modulus (reversed S_k) and general ones; every column is in Lambda^perp(a)."""
    from tools_b200 import gadget

    rng = np.random.default_rng(n)
    gp, po = T.GadgetParametersRing.init_default(n, q), O.GadgetParametersRing.init_default(n, q)
    psf = T.PSFGPVRing(gp, 100.0, 1.005)
    a_bar = rng.integers(0, q, n, dtype=np.int64)
    r = rng.integers(-3, 4, (gp.k, n)).astype(np.int32)
    e = rng.integers(-3, 4, (gp.k, n)).astype(np.int32)
    a = psf.gen_trapdoor_ring_lwe(a_bar, r, e)
    got = psf.gen_short_basis_for_trapdoor_ring(a, r, e)
    assert np.array_equal(got, gadget.ring_short_basis_embedded(gp, a, r, e))
    want = O.coeff_embed(O.gen_short_basis_for_trapdoor_ring(po, a.tolist(), r.tolist(), e.tolist()), n)
    assert got.tolist() == want
    # rot^-(a) * basis = 0 mod q: the columns lie in the kernel lattice (short_basis_ring.rs:183-341)
    rot_a = np.hstack([gadget.rot_minus(a[c]) for c in range(gp.k + 2)]).astype(object)
    assert not ((rot_a.dot(got.astype(object))) % q).any()

@pytest.mark.parametrize("n,q", [(5, 2**31 - 1 - 57), (6, 2**31 - 1), (8, 1024), (64, 3329)])
def test_ring_samp_p_preimage_and_domain(T, n, q):
    gp = T.GadgetParametersRing.init_default(n, q)
    psf = T.PSFGPVRing(gp, float(_ring_s(n)), 1.005)
    a, td = psf.trap_gen(seed=n)
    rng = np.random.default_rng(1)
    B = 200
    u = rng.integers(0, q, (B, n), dtype=np.int64)
    e = psf.samp_p_batch(a, td, u, seed=9)
    assert e.shape == (B, gp.k + 2, n)
    for b in range(0, B, 9):
        assert O.f_a_ring(a.tolist(), e[b].tolist(), n, q) == u[b].tolist()
    u2, flags = psf.f_a_batch(a, e)
    assert np.array_equal(u2, u) and flags.all()
    ds = psf.samp_d(seed=2)
    fa = psf.f_a(a, ds)
    pre = psf.samp_p(a, td, fa, seed=3)
    assert psf.f_a(a, pre).tolist() == fa.tolist() and psf.check_domain(pre)


# ----------------------------------------------------------------------------------
# device-resident entry points and a mid-size configuration
# ----------------------------------------------------------------------------------


def test_empty_and_single_target_batches(T):
    """Edge cases of the batch extension: an empty batch is valid (nothing is launched, empty results), a batch of one
    is what the trait-shaped single-target calls use."""
    n, q = 8, 64
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFPerturbation(gp, 3.0, 25.0)
    a, td = psf.trap_gen(seed=1)
    assert psf.samp_d_batch(0, seed=1).shape == (0, gp.m)
    u0, f0 = psf.f_a_batch(a, np.zeros((0, gp.m), dtype=np.int32))
    assert u0.shape == (0, n) and f0.shape == (0,)
    assert psf.check_domain_batch(np.zeros((0, gp.m), dtype=np.int32)).shape == (0,)
    assert psf.samp_p_batch(a, td, np.zeros((0, n), dtype=np.int64), seed=2).shape == (0, gp.m)
    rng = np.random.default_rng(0)
    u = rng.integers(0, q, (1, n), dtype=np.int64)
    e1 = psf.samp_p_batch(a, td, u, seed=3)
    assert e1.shape == (1, gp.m) and np.array_equal(O.f_a_classical_batch(a, e1, q), u)
    assert np.array_equal(psf.samp_p(a, td, u[0], seed=3), e1[0])  # trait call == batch of one
    g = T.PSFGPV(gp, 90.0)
    a2, td2 = g.trap_gen(seed=4)
    assert g.samp_p_batch(a2, td2, np.zeros((0, n), dtype=np.int64), seed=5).shape == (0, gp.m)
    gr = T.GadgetParametersRing.init_default(16, 257)
    pr = T.PSFGPVRing(gr, 300.0, 1.005)
    ar, tdr = pr.trap_gen(seed=6)
    assert pr.samp_d_batch(0, seed=7).shape == (0, gr.k + 2, 16)
    ur, fr = pr.f_a_batch(ar, np.zeros((0, gr.k + 2, 16), dtype=np.int32))
    assert ur.shape == (0, 16) and fr.shape == (0,)


def test_device_entry_points_and_midsize(T):
    import torch
    from tools_b200 import _ffi

    n, q, r, s = 32, 2**24, 5.0, 120.0
    gp, psf, a, td = _pert_setup(T, n, q, r, s, seed=3)
    B = 3000
    dev = torch.device("cuda:0")
    u = torch.empty((B, n), dtype=torch.int64, device=dev)
    assert _ffi.lib().qf_fill_uniform_modq_dev(_ffi.ptr(u.data_ptr()), u.numel(), q, 77, None) == 0
    torch.cuda.synchronize()
    psf._install_a(a)
    psf._install_td(a, td)
    e = torch.empty((B, gp.m), dtype=torch.int32, device=dev)
    psf.ctx.call("qf_set_stream", _ffi.ptr(torch.cuda.current_stream().cuda_stream))
    psf.ctx.call("qf_samp_p_dev", _ffi.ptr(u.data_ptr()), B, 5, 0, _ffi.ptr(e.data_ptr()))
    uo = torch.empty((B, n), dtype=torch.int64, device=dev)
    fl = torch.empty(B, dtype=torch.uint8, device=dev)
    psf.ctx.call("qf_f_a_dev", _ffi.ptr(e.data_ptr()), B, _ffi.ptr(uo.data_ptr()), _ffi.ptr(fl.data_ptr()))
    psf.ctx.call("qf_synchronize")
    assert torch.equal(uo, u) and bool(fl.all())
    assert np.array_equal(O.f_a_classical_batch(a, e.cpu().numpy()[:50], q), u.cpu().numpy()[:50])
    assert (u.cpu().numpy() < q).all() and u.cpu().numpy().std() > q / 8
    # host path gives the same preimages as the device path (same seed)
    e_host = psf.samp_p_batch(a, td, u.cpu().numpy(), seed=5)
    assert np.array_equal(e_host, e.cpu().numpy())
    assert psf.ctx.launch_count() > 0


@pytest.mark.parametrize("n,q", [(40, 2**16), (64, 2**16), (128, 2**12), (64, 4093)])
def test_gpv_tensor_core_updates_match_fp64_path(T, monkeypatch, n, q):
    """The fixed-point (int8 tcgen05) nearest-plane updates against the fp64 DMMA path on the same key and
    seed: both must give exact preimages with the same second moment; with identical Philox streams almost
    every preimage is identical (centres agree to ~2^-40).  m = 1316: digits and S z only; m = 2084 = 2 * 1024 + 36:
    one tensor-core update whose block has absorbed the thin ragged top (K = 1060); m = 3121 = 3 * 1024 + 49: two
    updates (K = 1073 with the merged top, then K = 1024); q = 4093 (not a power of the base): the gadget block S_k carries
    the digits of q, S' is not column-reversed and the gadget centre map M' is triangular the other way round."""
    import math

    gp = T.GadgetParameters.init_default(n, q)
    assert gp.m > 1024
    s = float(math.ceil((math.sqrt(gp.m_bar) + 1.0) * math.sqrt(5.0) * math.log2(n)))
    rng = np.random.default_rng(12)
    u = rng.integers(0, q, (2027, n), dtype=np.int64)  # ragged: the last 128-target tile and its last warp are partial
    outs = {}
    key = None
    # "two_phase": the default for a G-trapdoor basis at tensor-core dimensions (samp_p_np2_chunk; nk % 128 == 0);
    # "two_phase_wide": the same with every fixed-point matrix at full width and no dropped digit sums;
    # "ozaki": the one-pass recursion with tensor-core updates; "fp64": the one-pass recursion on DMMA
    for mode in ("two_phase", "two_phase_wide", "ozaki", "fp64"):
        monkeypatch.delenv("QF_DISABLE_OZAKI", raising=False)
        monkeypatch.delenv("QF_DISABLE_TWO_PHASE", raising=False)
        monkeypatch.delenv("QF_NP2_CFG", raising=False)
        monkeypatch.setenv("QF_OZAKI_MIN_DIM", "1024")
        if mode == "two_phase_wide":
            monkeypatch.setenv("QF_NP2_CFG", "7,0,7,0,7,0,7")
        elif mode == "ozaki":
            monkeypatch.setenv("QF_DISABLE_TWO_PHASE", "1")
        elif mode == "fp64":
            monkeypatch.setenv("QF_DISABLE_OZAKI", "1")
        psf = T.PSFGPV(gp, s)
        if key is None:
            key = psf.trap_gen(seed=31)
        a, td = key
        psf._a_id = None
        e = psf.samp_p_batch(a, td, u, seed=77)
        assert np.array_equal(O.f_a_classical_batch(a, e, q), u), mode
        assert psf.check_domain_batch(e).all(), mode
        ratio = (e.astype(np.float64) ** 2).sum(1).mean() / (gp.m * s * s / (2 * math.pi))
        assert abs(ratio - 1) < 0.02, (mode, ratio)
        outs[mode] = e
    same = (outs["ozaki"] == outs["fp64"]).all(axis=1).mean()
    assert same > 0.9, same
    if (n * gp.k) % 128 == 0:  # the two-phase path ran: its digit budget changes (almost) no preimage
        assert not np.array_equal(outs["two_phase"], outs["ozaki"])
        same2 = (outs["two_phase"] == outs["two_phase_wide"]).all(axis=1).mean()
        assert same2 > 0.9, same2
    else:
        assert np.array_equal(outs["two_phase"], outs["ozaki"])


@pytest.mark.parametrize("n,q,s,B", [(64, 2**16, None, 1500), (24, 2**16, 300.0, 700), (8, 127, 70.0, 333), (5, 32, 10.0, 130)])
def test_np_diag_kernels_agree(T, monkeypatch, n, q, s, B):
    """The diagonal-block kernel (one target per thread, groups of 8 steps, lattice.cu np_diag2) against the
    quad-per-two-targets kernel it replaced: same Philox counters and the same sequence of floating-point operations in
    the recursion itself => the same preimages.  The rank-64 update of np_diag2 runs on the fp64 tensor path (DMMA), whose
    summation order differs in the last bit of a centre: a rounding decision may flip for a handful of targets (each flip
    changes the rest of that target's draws), so the requirement is >= 99 % identical preimages, all of them exact.
    Shapes: tensor-core recursion with fused digit planes (two-phase), fp64 recursion with whole 256-blocks, dimensions
    that are not multiples of 64, a ragged last CTA."""
    import math

    gp = T.GadgetParameters.init_default(n, q)
    if s is None:
        s = float(math.ceil((math.sqrt(gp.m_bar) + 1.0) * math.sqrt(5.0) * math.log2(n)))
    rng = np.random.default_rng(n)
    u = rng.integers(0, q, (B, n), dtype=np.int64)
    monkeypatch.setenv("QF_OZAKI_MIN_DIM", "1024")
    outs, key = [], None
    for v1 in ("0", "1"):
        monkeypatch.setenv("QF_NP_DIAG_V1", v1)
        psf = T.PSFGPV(gp, s)
        if key is None:
            key = psf.trap_gen(seed=17)
        a, td = key
        psf._a_id = None
        outs.append(psf.samp_p_batch(a, td, u, seed=21))
        assert np.array_equal(O.f_a_classical_batch(a, outs[-1], q), u)
    same = (outs[0] == outs[1]).all(axis=1).mean()
    assert same >= 0.99, same


def test_full_size_c2_properties(T):
    """BASELINE.json configs[1] at full key size (PSFGPV n = 256, q = 2^24, m = 12352): the size-independent
    properties -- A e = u for every target (checked here in exact numpy integer arithmetic, not by the
    device), check_domain, the second moment of the spherical law, batch splitting, and f_a linearity."""
    import math

    n, q = 256, 2**24
    gp = T.GadgetParameters.init_default(n, q)
    assert (gp.k, gp.m_bar, gp.m) == (24, 6208, 12352)
    s = float(math.ceil((math.sqrt(gp.m_bar) + 1.0) * math.sqrt(5.0) * math.log2(n)))
    psf = T.PSFGPV(gp, s)
    a, td = psf.trap_gen(seed=2)
    rng = np.random.default_rng(3)
    B = 1536
    u = rng.integers(0, q, (B, n), dtype=np.int64)
    e = psf.samp_p_batch(a, td, u, seed=9)
    # exact integer check on the host: entries of A < 2^24, |e| < 2^15, m < 2^14  =>  fits int64
    assert np.array_equal((e.astype(np.int64) @ a.T) % q, u)
    assert psf.check_domain_batch(e).all()
    ratio = (e.astype(np.float64) ** 2).sum(1).mean() / (gp.m * s * s / (2 * math.pi))
    assert abs(ratio - 1) < 0.01, ratio
    # splitting the batch across calls (= across GPUs) reproduces the same preimages
    e2 = np.concatenate([psf.samp_p_batch(a, td, u[:500], seed=9), psf.samp_p_batch(a, td, u[500:], seed=9, first_index=500)])
    assert np.array_equal(e, e2)
    # f_a on the device agrees with the host product, and is additive mod q where the sum stays in the domain
    uu, flags = psf.f_a_batch(a, e)
    assert flags.all() and np.array_equal(uu, u)
    d = psf.samp_d_batch(64, seed=5)
    half = (d // 2).astype(np.int32)
    rest = (d - half).astype(np.int32)
    u1, _ = psf.f_a_batch(a, half)
    u2, _ = psf.f_a_batch(a, rest)
    u12, _ = psf.f_a_batch(a, d)
    assert np.array_equal((u1 + u2) % q, u12)


def test_full_size_c5_compression_properties(T):
    """C5 shape (degree-256 polynomials mod 3329), 1 Mi polynomials: decompress(compress(x)) stays within the
    FIPS 203 bound, compress is idempotent on decompressed values, and a checksum of the output agrees with
    the oracle on a strided sample."""
    rng = np.random.default_rng(5)
    x = rng.integers(0, 3329, (1 << 20, 256)).astype(np.uint16)
    for d in (1, 4, 10, 11):
        c = T.lossy_compress(x, d, 3329)
        back = T.lossy_decompress(c, d, 3329)
        dist = np.abs(back.astype(np.int32) - x.astype(np.int32))
        dist = np.minimum(dist, 3329 - dist)
        assert dist.max() <= 2 ** (12 - d - 1)
        assert np.array_equal(T.lossy_compress(back % 3329, d, 3329), c)  # Compress(Decompress(y)) = y
        sample = x[::4099]
        assert np.array_equal(c[::4099].astype(np.uint64), O.lossy_compress_np(sample, d, 3329))


def test_full_size_c3_ring_properties(T):
    """BASELINE.json configs[2] at full key size (PSFGPVRing over X^256 + 1, q = 3329, 14 polynomials, embedded
    dimension 3584): ring f_a against the exact negacyclic product on the host, additivity, A e = u for every
    preimage (exact host arithmetic), check_domain, the spherical second moment, and batch splitting."""
    n, q = 256, 3329
    gp = T.GadgetParametersRing.init_default(n, q)
    s = float(_ring_s(n))
    psf = T.PSFGPVRing(gp, s, 1.005)
    a, td = psf.trap_gen(seed=3)
    # the exact ring product via the negacyclic matrix rot^-(a_j) in int64 (entries < 2^12, |sigma| < 2^13, 3584 terms)
    rot = np.concatenate([T.gadget._rot_i64(a[j]) for j in range(gp.k + 2)], axis=1)  # n x n(k+2)
    sig = psf.samp_d_batch(4096, seed=5)
    u, flags = psf.f_a_batch(a, sig)
    assert flags.all()
    assert np.array_equal((sig.reshape(4096, -1).astype(np.int64) @ rot.T) % q, u)
    half = (sig // 2).astype(np.int32)
    u1, _ = psf.f_a_batch(a, half)
    u2, _ = psf.f_a_batch(a, (sig - half).astype(np.int32))
    assert np.array_equal((u1 + u2) % q, u)
    rng = np.random.default_rng(4)
    B = 4096
    tg = rng.integers(0, q, (B, n), dtype=np.int64)
    e = psf.samp_p_batch(a, td, tg, seed=7)
    assert e.shape == (B, gp.k + 2, n)
    assert np.array_equal((e.reshape(B, -1).astype(np.int64) @ rot.T) % q, tg)
    assert psf.check_domain_batch(e).all()
    D = n * (gp.k + 2)
    ratio = (e.astype(np.float64) ** 2).sum((1, 2)).mean() / (D * s * s / (2 * np.pi))
    assert abs(ratio - 1) < 0.01, ratio
    e2 = np.concatenate([psf.samp_p_batch(a, td, tg[:1000], seed=7), psf.samp_p_batch(a, td, tg[1000:], seed=7, first_index=1000)])
    assert np.array_equal(e, e2)


def test_full_size_c4_shard_properties(T):
    """BASELINE.json configs[3] at full key size (PSFPerturbation n = 512, q = 2^32 - 5, k = 32, m = 32849, r = 9),
    one GPU's shard: TrapGen A [R; I] = G exactly, A e = u for every preimage in exact host arithmetic (object
    integers on a sample, int64 with a 16-bit split of A for all), check_domain, the second moment, batch splitting."""
    import math

    n, q, r = 512, 2**32 - 5, 9.0
    gp = T.GadgetParameters.init_default(n, q)
    assert (gp.k, gp.m_bar, gp.m) == (32, 16465, 32849)
    s1 = (math.sqrt(gp.m_bar) + math.sqrt(n * gp.k)) / math.sqrt(2)
    s = float(math.ceil(1.15 * math.sqrt(5 * (s1 * s1 + 1) + 1)))
    psf = T.PSFPerturbation(gp, r, s)
    a, td = psf.trap_gen(seed=4)
    assert td[1] is None  # block-structured square root of the default Sigma_2, built on the device
    rmat = td[0]
    # A [R; I] = G mod q: exact via two 16-bit halves of A in float64 BLAS (|sum| < 2^16 * 16465 < 2^53)
    a_bar, a_g = a[:, :gp.m_bar], a[:, gp.m_bar:]
    rf = rmat.astype(np.float64)
    lo = np.rint((a_bar & 0xFFFF).astype(np.float64) @ rf).astype(np.int64)
    hi = np.rint((a_bar >> 16).astype(np.float64) @ rf).astype(np.int64)
    prod = (lo % q + ((hi % q) << 16) % q + a_g) % q
    g = np.zeros((n, n * gp.k), dtype=np.int64)
    for j in range(n):
        g[j, j * gp.k:(j + 1) * gp.k] = [(1 << t) % q for t in range(gp.k)]
    assert np.array_equal(prod, g)
    rng = np.random.default_rng(6)
    B = 2048
    u = rng.integers(0, q, (B, n), dtype=np.int64)
    e = psf.samp_p_batch(a, td, u, seed=9)
    ef = e.astype(np.float64)
    lo = np.rint(ef @ (a & 0xFFFF).astype(np.float64).T).astype(np.int64)      # |e| < 2^16, 32849 terms: < 2^47
    hi = np.rint(ef @ (a >> 16).astype(np.float64).T).astype(np.int64)
    assert np.array_equal((lo % q + ((hi % q) << 16) % q) % q, u)
    assert [int(x) for x in (e[:2].astype(object) @ a.T.astype(object) % q).ravel()] == u[:2].ravel().tolist()
    assert psf.check_domain_batch(e).all()
    ratio = (ef**2).sum(1).mean() / (gp.m * (s * r) ** 2 / (2 * math.pi))
    assert abs(ratio - 1) < 0.01, ratio
    e2 = np.concatenate([psf.samp_p_batch(a, td, u[:700], seed=9), psf.samp_p_batch(a, td, u[700:], seed=9, first_index=700)])
    assert np.array_equal(e, e2)
    uu, flags = psf.f_a_batch(a, e)
    assert flags.all() and np.array_equal(uu, u)
