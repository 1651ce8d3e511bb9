"""The C restatement (timed CPU arm) agrees with the Python oracle.  CPU only."""
import numpy as np

from oracle import oracle_c as C
from oracle import qfall_oracle as O


def test_f_a_and_compress():
    rng = np.random.default_rng(0)
    q = 2**24 - 3
    a = rng.integers(0, q, (6, 40), dtype=np.int64)
    sig = rng.integers(-500, 500, (9, 40)).astype(np.int32)
    assert np.array_equal(C.f_a(a, sig, q, 2), O.f_a_classical_batch(a, sig, q))
    x = rng.integers(0, 3329, 10000).astype(np.uint16)
    for d in (1, 4, 10, 11):
        c = C.compress_u16(x, 3329, d, False, 2)
        assert np.array_equal(c.astype(np.uint64), O.lossy_compress_np(x, d, 3329))
        assert np.array_equal(C.compress_u16(c, 3329, d, True).astype(np.uint64), O.lossy_decompress_np(c, d, 3329))


def test_sample_z_law():
    s, c, n = 3.0, 0.3, 100000
    out = C.sample_z(s, c, 7, n)
    xs, pm = O.dgauss_pmf(s, c)
    cnt = np.array([(out == x).sum() for x in xs])
    keep = pm * n >= 8
    chi = ((cnt[keep] - pm[keep] * n) ** 2 / (pm[keep] * n)).sum()
    dof = keep.sum() - 1
    assert chi < dof + 5 * np.sqrt(2 * dof)


def test_samp_p_restatements_are_preimages():
    rng = np.random.default_rng(1)
    # PSFPerturbation README parameters
    n, q, r, s = 8, 64, 3.0, 25.0
    p = O.GadgetParameters.init_default(n, q)
    a_bar = rng.integers(0, q, (n, p.m_bar)).tolist()
    rm = O.sample_pm_one_zero(rng, p.m_bar, n * p.k)
    a = np.array(O.gen_trapdoor(p, a_bar, O.mat_identity(n), rm), dtype=np.int64)
    l = O.compute_sqrt_sigma_2(rm, s, r, 2)
    sb = np.array(O.short_basis_gadget(p), dtype=np.float64)
    sg = O.gso_f64(sb)
    u = rng.integers(0, q, (64, n), dtype=np.int64)
    e = C.samp_p_pert(l, a, np.array(rm), sb, sg, u, n, p.k, p.m_bar, 2, q, r, 5, 2)
    assert np.array_equal(O.f_a_classical_batch(a, e, q), u)
    assert all(O.check_domain_perturbation(row.tolist(), p.m, s, r) for row in e)
    # PSFGPV
    n, q, s = 5, 32, 10.0
    p = O.GadgetParameters.init_default(n, q)
    a_bar = rng.integers(0, q, (n, p.m_bar)).tolist()
    rm = O.sample_pm_one_zero(rng, p.m_bar, n * p.k)
    tag = O.mat_identity(n)
    a = np.array(O.gen_trapdoor(p, a_bar, tag, rm), dtype=np.int64)
    sb = np.array(O.gen_short_basis_for_trapdoor(p, tag, a.tolist(), rm), dtype=np.float64)
    sg = O.gso_f64(sb)
    piv, ainv = C.unit_pivots(a, q)
    u = rng.integers(0, q, (64, n), dtype=np.int64)
    e = C.samp_p_gpv(sb, sg, piv, ainv, u, q, s, 3, 2)
    assert np.array_equal(O.f_a_classical_batch(a, e, q), u)
    assert all(O.check_domain_gpv(row.tolist(), p.m, s) for row in e)
    # same law as the Python restatement: compare ||e||^2 means on one syndrome
    uu = np.tile(u[:1], (400, 1))
    ec = C.samp_p_gpv(sb, sg, piv, ainv, uu, q, s, 9, 2).astype(np.float64)
    rr = np.random.default_rng(2)
    ep = np.array([O.samp_p_gpv(rr, a.tolist(), q, sb.tolist(), sg, u[0].tolist(), s) for _ in range(200)], dtype=np.float64)
    n1, n2 = (ec**2).sum(1), (ep**2).sum(1)
    assert abs(n1.mean() - n2.mean()) < 5 * np.sqrt(n1.var() / len(n1) + n2.var() / len(n2))
