"""Statistical parity of the samplers against the reference's own algorithms at the same parameters
(north star: "the discrete-Gaussian marginals must pass chi-square and moment tests against the reference sampler at
the same s").  The reference has no seed API, so sampler outputs cannot be compared value by value; here

  * the 1-D sampler is tested at 10^7 draws per (s, frac c) against the EXACT law D_{Z,s,c}
    (rho(x) = exp(-pi (x-c)^2 / s^2), support [c - ceil 6s, c + floor 6s], CONTRIBUTING.md:35-45) -- the widths
    include those the C2 / C4 pipelines really draw from (s, s r, r, r sqrt(b^2+1) / ||b~_i||, s / ||b~_i||);
  * whole `samp_p` outputs are tested coordinate by coordinate with a two-sample chi-square against >= 2 * 10^4 draws
    of the reference loop (oracle/oracle_c.c: uniform-proposal SampleZ, dense SampleD, Peikert Alg. 1), same key, same
    syndrome;
  * `randomized_nearest_plane_gadget` (mp_perturbation.rs:173-191) is tested on its own: G z = v, the exact law of
    the first sampled coordinate, two-sample chi-square of every coordinate against the oracle's restatement;
  * TrapGen's randomness: R = U{0,1} - U{0,1} (trapdoor_distribution.rs:82-86) and the uniform A_bar (gpv.rs:84).

Acceptance: chi^2 < dof + Z * sqrt(2 dof), Z = 4 (seeds are fixed, so the tests are deterministic; a correct sampler
fails a 4-sigma band with probability ~ 3e-5 per statistic)."""
import math

import numpy as np
import pytest

from oracle import oracle_c as OC
from oracle import qfall_oracle as O

pytestmark = pytest.mark.gpu
Z_BAND = 4.0


@pytest.fixture(scope="module")
def T():
    import tools_b200

    return tools_b200


def chi2_gof(samples, xs, pm, min_exp=10.0):
    """Pearson goodness of fit of integer `samples` against the pmf (xs, pm); cells with expectation below min_exp are
    pooled into one.  Returns (ok, chi2, dof, n_outside_support)."""
    n = samples.size
    lo = int(xs[0])
    idx = samples - lo
    outside = int(((idx < 0) | (idx >= xs.size)).sum())
    cnt = np.bincount(idx[(idx >= 0) & (idx < xs.size)], minlength=xs.size).astype(np.float64)
    exp = pm * n
    keep = exp >= min_exp
    chi = ((cnt[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    dof = int(keep.sum()) - 1
    rest_e, rest_o = exp[~keep].sum(), cnt[~keep].sum() + outside
    if rest_e >= min_exp:
        chi += (rest_o - rest_e) ** 2 / rest_e
        dof += 1
    return chi < dof + Z_BAND * math.sqrt(2.0 * dof), chi, dof, outside


def chi2_two_sample(a, b, min_pool=25):
    """Two-sample chi-square of integer samples a (N1) and b (N2) with unequal sizes (Numerical Recipes 14.3.3):
    chi2 = sum (sqrt(N2/N1) a_i - sqrt(N1/N2) b_i)^2 / (a_i + b_i).  Values are grouped into cells of width
    ~ std / 8 so that the smaller sample has enough counts; sparse cells are pooled."""
    a = np.asarray(a, dtype=np.int64)
    b = np.asarray(b, dtype=np.int64)
    w = max(1, int(min(a.std(), b.std()) / 8.0))
    lo = min(a.min(), b.min())
    ca = np.bincount((a - lo) // w)
    cb = np.bincount((b - lo) // w)
    size = max(ca.size, cb.size)
    ca = np.pad(ca, (0, size - ca.size)).astype(np.float64)
    cb = np.pad(cb, (0, size - cb.size)).astype(np.float64)
    keep = (ca + cb) >= min_pool
    pa, pb = ca[~keep].sum(), cb[~keep].sum()
    ca, cb = ca[keep], cb[keep]
    if pa + pb >= min_pool:
        ca, cb = np.append(ca, pa), np.append(cb, pb)
    k1, k2 = math.sqrt(b.size / a.size), math.sqrt(a.size / b.size)
    chi = ((k1 * ca - k2 * cb) ** 2 / (ca + cb)).sum()
    dof = ca.size - 1
    return chi < dof + Z_BAND * math.sqrt(2.0 * dof), chi, dof


def moments_match(a, b, tol_sigma=5.0):
    """Means and variances of two samples agree within tol_sigma standard errors."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    se_m = math.sqrt(a.var() / a.size + b.var() / b.size)
    se_v = math.sqrt(2.0 * a.var() ** 2 / a.size + 2.0 * b.var() ** 2 / b.size) * 1.2  # discrete: slight excess kurtosis
    return abs(a.mean() - b.mean()) < tol_sigma * se_m and abs(a.var() - b.var()) < tol_sigma * se_v


# ---------------------------------------------------------------------------------------------------------------
# (i) the 1-D sampler at 10^7 draws against the exact law
# ---------------------------------------------------------------------------------------------------------------
_SQ5 = math.sqrt(5.0)
SAMPLE_Z_GRID = [
    # (s, c)  -- what draws from this width
    (1.0, 0.0), (1.0, 0.5),                  # narrowest the library supports sensibly (sigma = 0.4)
    (1.8, 0.25), (1.8, 1e6 + 0.5),           # far-off centre: the integer part is carried in fp64
    (3.0, 0.0), (3.0, 0.37), (3.0, -0.5),    # C1 rounding parameter r = 3 (mp_perturbation.rs:315)
    (3.0 * _SQ5 / 2.0, 0.123),               # gadget sampler, C1: r sqrt(b^2+1) / ||b~_i||, ||b~|| = 2
    (9.0, 0.75), (9.0 * _SQ5 / 2.0, -0.31),  # C4: r = 9 and its gadget width
    (5.7, 0.5), (12.8, 0.049),               # C2: s / ||b~_i|| for the longest Gram-Schmidt vectors (||b~|| ~ 250, 110)
    (25.0, 0.0), (75.0, 0.5),                # C1: s and s r (samp_d)
    (1428.0, 0.0), (1428.0, 0.3),            # C2: s (samp_d) and a mid-range nearest-plane width
    (4203.0, 0.5),                           # C4: s r = 467 * 9
    (158000.0, 0.77),                        # C2: s / ||b~_i|| at the shortest Gram-Schmidt vector (||b~|| ~ 1/111)
]


@pytest.mark.parametrize("s,c", SAMPLE_Z_GRID)
def test_sample_z_exact_law_10m(T, s, c):
    from tools_b200 import _ffi

    n = 10_000_000
    centers = np.full(n, c, dtype=np.float64)
    out = np.empty(n, dtype=np.int64)
    seed = int(s * 1000) * 7919 + int((c % 1.0) * 1000) + 5
    assert _ffi.lib().qf_sample_z(_ffi.ptr(centers), n, float(s), seed, _ffi.ptr(out)) == 0
    xs, pm = O.dgauss_pmf(s, c)
    if s > 2000:
        # 10^7 draws over ~10^6 support points: test the law in cells of width sigma / 16 (about 100 cells in +-3 sigma,
        # 4 * 10^4 draws each near the centre -- 0.5 % resolution) instead of per integer
        w = max(1, int(s / math.sqrt(2 * math.pi) / 16))
        lo = int(xs[0])
        cell = (np.arange(xs.size) // w)
        pm_c = np.bincount(cell, weights=pm)
        ok, chi, dof, outside = chi2_gof((out - lo) // w, np.arange(pm_c.size), pm_c)
    else:
        ok, chi, dof, outside = chi2_gof(out, xs, pm)
    assert outside == 0, "sample outside the reference's 6 s support"
    assert ok, (s, c, chi, dof)
    # moments against the exact ones
    mean = float((xs * pm).sum())
    var = float(((xs - mean) ** 2 * pm).sum())
    m4 = float(((xs - mean) ** 4 * pm).sum())
    of = out.astype(np.float64)
    assert abs(of.mean() - mean) < 5.0 * math.sqrt(var / n)
    assert abs(of.var() - var) < 5.0 * math.sqrt((m4 - var * var) / n)
    assert abs(((of - mean) ** 4).mean() / var**2 - m4 / var**2) < 0.02


def test_sample_z_vs_reference_samplez_10m(T):
    """Two-sample chi-square, 10^7 draws each, CUDA sampler against the C restatement of the reference's SampleZ
    (uniform proposal on the cut interval, accept with rho_{s,c}(x))."""
    from tools_b200 import _ffi

    n = 10_000_000
    for s, c in [(3.0 * _SQ5, 0.3), (25.4, -0.45)]:
        out = np.empty(n, dtype=np.int64)
        centers = np.full(n, c, dtype=np.float64)
        assert _ffi.lib().qf_sample_z(_ffi.ptr(centers), n, s, 424242, _ffi.ptr(out)) == 0
        ref = OC.sample_z(s, c, 99, n)
        ok, chi, dof = chi2_two_sample(out, ref)
        assert ok, (s, c, chi, dof)
        assert moments_match(out, ref)


def test_sample_z_bad_centre_is_reported(T):
    """A NaN centre makes every proposal fail: after the bail-out the call reports QF_ERR_NUMERIC instead of silently
    returning the centre."""
    from tools_b200 import _ffi

    centers = np.array([0.0, float("nan"), 1.5], dtype=np.float64)
    out = np.empty(3, dtype=np.int64)
    assert _ffi.lib().qf_sample_z(_ffi.ptr(centers), 3, 4.0, 1, _ffi.ptr(out)) == _ffi.QF_ERR_NUMERIC
    assert _ffi.lib().qf_sample_z(_ffi.ptr(centers), 3, 3.0e6, 1, _ffi.ptr(out)) == _ffi.QF_ERR_UNSUPPORTED


def test_sample_z_and_samp_d_streams_are_independent(T):
    """Equal seeds in qf_sample_z and qf_samp_d must not give correlated outputs (separate Philox stream ids)."""
    from tools_b200 import _ffi

    gp = T.GadgetParameters.init_default(8, 64)
    psf = T.PSFGPV(gp, 10.0)
    d = psf.samp_d_batch(2000, seed=77).reshape(-1).astype(np.float64)
    out = np.empty(d.size, dtype=np.int64)
    centers = np.zeros(d.size, dtype=np.float64)
    assert _ffi.lib().qf_sample_z(_ffi.ptr(centers), d.size, 10.0, 77, _ffi.ptr(out)) == 0
    assert abs(np.corrcoef(d, out.astype(np.float64))[0, 1]) < 5.0 / math.sqrt(d.size)


# ---------------------------------------------------------------------------------------------------------------
# (ii) samp_p marginals against the reference loop (oracle_c.c), same key and syndrome
# ---------------------------------------------------------------------------------------------------------------
def _coords(m, split, count=12):
    """coordinates spread over both blocks of the Domain vector (top m_bar, bottom nk), ends included"""
    c = sorted({0, 1, split // 2, split - 1, split, split + 1, (split + m) // 2, m - 2, m - 1})
    rng = np.random.default_rng(m)
    extra = [int(x) for x in rng.choice(m, count, replace=False)]
    return sorted(set(c + extra))


@pytest.mark.parametrize("n,q,s", [(8, 64, 60.0), (8, 127, 70.0), (24, 2**16, 300.0)])
def test_samp_p_gpv_marginals_vs_reference(T, n, q, s):
    """PSFGPV::samp_p (gpv.rs:152-161): every tested coordinate of e has the law the reference loop produces."""
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFGPV(gp, s)
    a, (sb, _) = psf.trap_gen(seed=n)
    td = (sb, None)
    rng = np.random.default_rng(n + 1)
    u1 = rng.integers(0, q, (1, n), dtype=np.int64)
    n_gpu, n_ref = 200_000, 24_000 if gp.m < 300 else 20_000
    e = psf.samp_p_batch(a, td, np.tile(u1, (n_gpu, 1)), seed=5)
    assert np.array_equal(O.f_a_classical_batch(a, e[:64], q), np.tile(u1, (64, 1)))
    gso = O.gso_f64(sb.astype(np.float64))
    piv, ainv = OC.unit_pivots(a, q)
    ref = OC.samp_p_gpv(sb, gso, piv, ainv, np.tile(u1, (n_ref, 1)), q, s, 12345, OC.threads())
    assert np.array_equal(O.f_a_classical_batch(a, ref[:16], q), np.tile(u1, (16, 1)))
    bad = []
    for j in _coords(gp.m, gp.m_bar):
        ok, chi, dof = chi2_two_sample(e[:, j], ref[:, j])
        if not ok or not moments_match(e[:, j], ref[:, j]):
            bad.append((j, chi, dof))
    assert not bad, bad
    # joint statistic: squared norm
    assert moments_match((e.astype(np.float64) ** 2).sum(1), (ref.astype(np.float64) ** 2).sum(1))


def test_samp_p_gpv_two_phase_marginals_vs_reference(T, monkeypatch):
    """The two-phase form of the recursion (gadget preimage as centre, residual reduced modulo the [R;I]S' sub-lattice,
    DESIGN 4.3) against the reference loop run with the pivot-column particular solution of gpv.rs:153-156, on the same
    key and syndrome: every tested coordinate of e must have the same law.  n = 32, q = 2^16: m = 1049 > 1024, so the
    tensor-core path and with it the two-phase form are active (checked: the one-pass form gives other preimages)."""
    n, q, s = 32, 2**16, 330.0
    monkeypatch.setenv("QF_OZAKI_MIN_DIM", "1024")
    gp = T.GadgetParameters.init_default(n, q)
    assert gp.m > 1024 and (n * gp.k) % 128 == 0
    psf = T.PSFGPV(gp, s)
    a, (sb, _) = psf.trap_gen(seed=n)
    rng = np.random.default_rng(n + 1)
    u1 = rng.integers(0, q, (1, n), dtype=np.int64)
    n_gpu, n_ref = 120_000, 20_000
    e = psf.samp_p_batch(a, (sb, None), np.tile(u1, (n_gpu, 1)), seed=5)
    assert np.array_equal(O.f_a_classical_batch(a, e[:64], q), np.tile(u1, (64, 1)))
    monkeypatch.setenv("QF_DISABLE_TWO_PHASE", "1")
    psf1 = T.PSFGPV(gp, s)
    psf1._install_a(a)
    e1 = psf1.samp_p_batch(a, (sb, None), np.tile(u1, (256, 1)), seed=5)
    assert not np.array_equal(e1, e[:256])
    gso = O.gso_f64(sb.astype(np.float64))
    piv, ainv = OC.unit_pivots(a, q)
    ref = OC.samp_p_gpv(sb, gso, piv, ainv, np.tile(u1, (n_ref, 1)), q, s, 12345, OC.threads())
    assert np.array_equal(O.f_a_classical_batch(a, ref[:16], q), np.tile(u1, (16, 1)))
    bad = []
    for j in _coords(gp.m, gp.m_bar, 20):
        ok, chi, dof = chi2_two_sample(e[:, j], ref[:, j])
        if not ok or not moments_match(e[:, j], ref[:, j]):
            bad.append((j, chi, dof))
    assert not bad, bad
    assert moments_match((e.astype(np.float64) ** 2).sum(1), (ref.astype(np.float64) ** 2).sum(1))


@pytest.mark.parametrize("n,q,r,s,structured", [(8, 64, 3.0, 25.0, False), (8, 64, 3.0, 25.0, True),
                                                 (16, 2**10, 4.0, 60.0, False), (16, 2**10, 4.0, 60.0, True)])
def test_samp_p_perturbation_marginals_vs_reference(T, n, q, r, s, structured):
    """PSFPerturbation::samp_p (mp_perturbation.rs:304-336), C1 (README example) and a second shape: coordinates of both
    blocks against the reference loop with the dense sqrt(Sigma_2); `structured` = the backend's own block square root
    of the same Sigma_2 (sqrt_sigma_2 = None), which must give the same law."""
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFPerturbation(gp, r, s)
    a, (rmat, l, (sb, sg)) = psf.trap_gen(seed=3, dense_sqrt_sigma_2=True, full_gadget_basis=True)
    td = (rmat, None if structured else l, (sb, sg))
    rng = np.random.default_rng(7)
    u1 = rng.integers(0, q, (1, n), dtype=np.int64)
    n_gpu, n_ref = 200_000, 24_000
    e = psf.samp_p_batch(a, td, np.tile(u1, (n_gpu, 1)), seed=6)
    assert np.array_equal(O.f_a_classical_batch(a, e[:64], q), np.tile(u1, (64, 1)))
    l_ref = O.compute_sqrt_sigma_2(rmat, s, r, 2)
    sb_ref = np.array(O.short_basis_gadget(O.GadgetParameters.init_default(n, q)), dtype=np.float64)
    ref = OC.samp_p_pert(l_ref, a, rmat, sb_ref, O.gso_f64(sb_ref), np.tile(u1, (n_ref, 1)), n, gp.k, gp.m_bar, 2, q, r,
                         4321, OC.threads())
    assert np.array_equal(O.f_a_classical_batch(a, ref[:16], q), np.tile(u1, (16, 1)))
    bad = []
    for j in _coords(gp.m, gp.m_bar):
        ok, chi, dof = chi2_two_sample(e[:, j], ref[:, j])
        if not ok or not moments_match(e[:, j], ref[:, j]):
            bad.append((j, chi, dof))
    assert not bad, bad
    assert moments_match((e.astype(np.float64) ** 2).sum(1), (ref.astype(np.float64) ** 2).sum(1))


def test_samp_p_ring_marginals_vs_reference(T):
    """PSFGPVRing::samp_p (gpv_ring.rs:160-212), n = 8, q = 1024: the reference embeds the basis, solves
    rot^-(a) x = u (a_0 = 1: x = (u, 0, ..)) and runs SampleD -- restated with the oracle's basis / GSO and the C loop."""
    n, q = 8, 1024
    gr = T.GadgetParametersRing.init_default(n, q)
    s = float(((2 * 2 * 1.005 * math.sqrt(n) + 1) * 2) * 4)  # gpv_ring.rs:296-298
    psf = T.PSFGPVRing(gr, s, 1.005)
    a, (r, e_td) = psf.trap_gen(seed=4)
    rng = np.random.default_rng(2)
    u1 = rng.integers(0, q, (1, n), dtype=np.int64)
    n_gpu, n_ref = 200_000, 24_000
    e = psf.samp_p_batch(a, (r, e_td), np.tile(u1, (n_gpu, 1)), seed=8).reshape(n_gpu, -1)
    po = O.GadgetParametersRing.init_default(n, q)
    emb = np.array(O.coeff_embed(O.gen_short_basis_for_trapdoor_ring(po, a.tolist(), r.tolist(), e_td.tolist()), n),
                   dtype=np.int64)
    gso = O.gso_f64(emb.astype(np.float64))
    piv = np.arange(n, dtype=np.int32)
    ref = OC.samp_p_gpv(emb, gso, piv, np.eye(n, dtype=np.int64), np.tile(u1, (n_ref, 1)), q, s, 777, OC.threads())
    polys = ref[0].reshape(gr.k + 2, n)
    assert O.f_a_ring(a.tolist(), polys.tolist(), n, q) == u1[0].tolist()
    d = n * (gr.k + 2)
    bad = []
    for j in _coords(d, 2 * n):
        ok, chi, dof = chi2_two_sample(e[:, j], ref[:, j])
        if not ok or not moments_match(e[:, j], ref[:, j]):
            bad.append((j, chi, dof))
    assert not bad, bad
    assert moments_match((e.astype(np.float64) ** 2).sum(1), (ref.astype(np.float64) ** 2).sum(1))


@pytest.mark.parametrize("n,q", [(64, 2**16), (128, 2**12)])
def test_samp_p_gpv_tensor_core_path_same_law_as_fp64(T, monkeypatch, n, q):
    """The fixed-point (int8 tcgen05) nearest-plane updates against the fp64 DMMA path, DIFFERENT seeds: two-sample
    chi-square of coordinates of both blocks (dimensions 2084 / 3121: one and two tensor-core update levels)."""
    gp = T.GadgetParameters.init_default(n, q)
    s = float(math.ceil((math.sqrt(gp.m_bar) + 1.0) * math.sqrt(5.0) * math.log2(n)))
    rng = np.random.default_rng(12)
    u1 = rng.integers(0, q, (1, n), dtype=np.int64)
    B = 40_000
    outs, key = [], None
    for mode, seed in (("ozaki", 1), ("fp64", 2)):
        if mode == "ozaki":
            monkeypatch.setenv("QF_OZAKI_MIN_DIM", "1024")
            monkeypatch.delenv("QF_DISABLE_OZAKI", raising=False)
        else:
            monkeypatch.setenv("QF_DISABLE_OZAKI", "1")
        psf = T.PSFGPV(gp, s)
        if key is None:
            key = psf.trap_gen(seed=31)
        a, td = key
        psf._a_id = None
        outs.append(psf.samp_p_batch(a, td, np.tile(u1, (B, 1)), seed=seed))
        assert np.array_equal(O.f_a_classical_batch(a, outs[-1][:32], q), np.tile(u1, (32, 1)))
    bad = []
    for j in _coords(gp.m, gp.m_bar, 16):
        ok, chi, dof = chi2_two_sample(outs[0][:, j], outs[1][:, j])
        if not ok or not moments_match(outs[0][:, j], outs[1][:, j]):
            bad.append((j, chi, dof))
    assert not bad, bad


# ---------------------------------------------------------------------------------------------------------------
# (iii) the gadget sampler on its own (mp_perturbation.rs:173-191)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,q,r", [(4, 16, 3.0), (3, 125, 2.5), (2, 2**10, 4.0)])
def test_randomized_nearest_plane_gadget_law(T, n, q, r):
    """z = x0 + SampleD(S, S~, -x0, r sqrt(b^2+1)) for one fixed syndrome v: (a) G z = v exactly; (b) the LAST
    coordinate of every block is the first one SampleD draws, its law is exactly D_{Z, s_G / ||b~_{k-1}||, c'} with
    c' = <-x0, b~_{k-1}> / ||b~_{k-1}||^2 -- chi-square against that pmf; (c) every coordinate: two-sample chi-square
    against the oracle's literal restatement (dense nk x nk SampleD like the reference)."""
    gp = T.GadgetParameters.init_default(n, q)
    s = 1.5 * math.sqrt(5 * ((math.sqrt(gp.m_bar) + math.sqrt(n * gp.k)) ** 2 / 2 + 1) + 1)
    psf = T.PSFPerturbation(gp, r, s)
    a, td = psf.trap_gen(seed=9, full_gadget_basis=True)
    rng = np.random.default_rng(n)
    v1 = rng.integers(0, q, (1, n), dtype=np.int64)
    B = 200_000
    z = psf.randomized_nearest_plane_gadget_batch(a, td, np.tile(v1, (B, 1)), seed=3)
    k, nk = gp.k, n * gp.k
    assert z.shape == (B, nk)
    g = np.array(O.gen_gadget_mat(n, k, gp.base), dtype=object)
    assert np.array_equal(z[:512].astype(object).dot(g.T) % q, np.tile(v1.astype(object), (512, 1)))
    # split invariance (Philox keyed by the global target index)
    z2 = np.concatenate([psf.randomized_nearest_plane_gadget_batch(a, td, np.tile(v1, (100, 1)), seed=3),
                         psf.randomized_nearest_plane_gadget_batch(a, td, np.tile(v1, (60, 1)), seed=3, first_index=100)])
    assert np.array_equal(z[:160], z2)
    # (b) exact law of the first drawn coordinate of each block
    po = O.GadgetParameters.init_default(n, q)
    blk = np.array(O.short_basis_gadget_block(k, gp.base, q), dtype=np.float64)
    gs = O.gso_f64(blk)
    s_g = r * math.sqrt(gp.base**2 + 1)
    for row in range(n):
        x0 = np.array(O.find_solution_gadget_vec(int(v1[0, row]), q, k, gp.base), dtype=np.float64)
        bt = gs[:, k - 1]
        cprime = float(-x0 @ bt) / float(bt @ bt)
        xs, pm = O.dgauss_pmf(s_g / math.sqrt(float(bt @ bt)), cprime)
        # z_{k-1} = x0_{k-1} + (integer drawn) * S[k-1][k-1] + ...: only column k-1 of S_k touches row k-1 when
        # q = b^k (S_k lower bidiagonal); for other q the last column holds q's digits and row k-1 of S_k is
        # (0, .., -1, q_{k-1}): recover the drawn integer from the relation instead
        col = z[:, row * k + k - 1].astype(np.int64)
        skk = blk[k - 1, k - 1]
        if all(blk[k - 1, j] == 0 for j in range(k - 1)):
            drawn = (col - int(x0[k - 1])) / skk
            assert np.all(drawn == np.rint(drawn))
            ok, chi, dof, outside = chi2_gof(drawn.astype(np.int64), xs, pm)
            assert outside == 0 and ok, (row, chi, dof)
    # (c) every coordinate against the oracle's literal restatement
    rr = np.random.default_rng(5)
    sb_full = np.array(O.short_basis_gadget(po), dtype=np.float64)
    sg_full = O.gso_f64(sb_full)
    n_ref = 20_000
    ref = np.array([O.randomized_nearest_plane_gadget(rr, v1[0].tolist(), po, r, sb_full, sg_full) for _ in range(n_ref)],
                   dtype=np.int64)
    assert np.array_equal(ref[:64].astype(object).dot(g.T) % q, np.tile(v1.astype(object), (64, 1)))
    bad = []
    for j in range(nk):
        ok, chi, dof = chi2_two_sample(z[:, j], ref[:, j])
        if not ok or not moments_match(z[:, j], ref[:, j]):
            bad.append((j, chi, dof))
    assert not bad, bad


def test_samp_p_perturbation_with_zero_trapdoor_isolates_gadget(T):
    """R = 0 (A = [A_bar | G]): e = p + [0; z], so the lower block of e is carried by the gadget sampler alone when
    Sigma_2's lower block is small (s^2 just above b^2 + 2).  Lower-block coordinates against the reference loop."""
    n, q, r = 4, 16, 3.0
    gp = T.GadgetParameters.init_default(n, q)
    s = 2.6  # s^2 - 1 - (b^2+1) = 0.76 > 0: positive definite, p_low has variance 0.76 r^2/2pi + rounding
    psf = T.PSFPerturbation(gp, r, s)
    rng = np.random.default_rng(3)
    a_bar = rng.integers(0, q, (n, gp.m_bar), dtype=np.int64)
    rmat = np.zeros((gp.m_bar, n * gp.k), dtype=np.int8)
    from tools_b200.psf import gen_trapdoor

    a = gen_trapdoor(gp, a_bar, rmat)
    assert np.array_equal(a[:, gp.m_bar:], np.array(O.gen_gadget_mat(n, gp.k, 2), dtype=np.int64) % q)
    l = psf.compute_sqrt_sigma_2(rmat)
    po = O.GadgetParameters.init_default(n, q)
    sb = np.array(O.short_basis_gadget(po), dtype=np.int64)
    sg = O.gso_f64(sb.astype(np.float64))
    td = (rmat, l, (sb, sg))
    u1 = rng.integers(0, q, (1, n), dtype=np.int64)
    n_gpu, n_ref = 200_000, 24_000
    e = psf.samp_p_batch(a, td, np.tile(u1, (n_gpu, 1)), seed=2)
    assert np.array_equal(O.f_a_classical_batch(a, e[:64], q), np.tile(u1, (64, 1)))
    ref = OC.samp_p_pert(O.compute_sqrt_sigma_2(rmat, s, r, 2), a, rmat, sb.astype(np.float64), sg, np.tile(u1, (n_ref, 1)),
                         n, gp.k, gp.m_bar, 2, q, r, 99, OC.threads())
    bad = []
    for j in range(gp.m_bar, gp.m):
        ok, chi, dof = chi2_two_sample(e[:, j], ref[:, j])
        if not ok or not moments_match(e[:, j], ref[:, j]):
            bad.append((j, chi, dof))
    assert not bad, bad


# ---------------------------------------------------------------------------------------------------------------
# (iv) TrapGen randomness
# ---------------------------------------------------------------------------------------------------------------
def test_trap_gen_r_and_a_bar_frequencies(T):
    """R = U{0,1} - U{0,1}: P(-1) = P(1) = 1/4, P(0) = 1/2 (trapdoor_distribution.rs:82-86), entries independent of
    their neighbours; A_bar uniform over [0, q) (gpv.rs:84): chi-square over 256 buckets, for a power-of-two and a
    prime modulus."""
    from tools_b200.psf import _trap_gen_classical

    for n, q, seed in [(64, 2**24, 11), (48, 2**20 - 3, 12)]:
        gp = T.GadgetParameters.init_default(n, q)
        psf = T.PSFGPV(gp, 1000.0)
        a, r = _trap_gen_classical(psf, seed)
        cnt = np.array([(r == v).sum() for v in (-1, 0, 1)], dtype=np.float64)
        tot = r.size
        assert set(np.unique(r).tolist()) <= {-1, 0, 1}
        for c, p in zip(cnt, (0.25, 0.5, 0.25)):
            assert abs(c - p * tot) < Z_BAND * math.sqrt(tot * p * (1 - p)), (cnt / tot)
        # neighbouring entries (same 16-entry Philox group) and rows are uncorrelated
        rf = r.astype(np.float64)
        assert abs(np.corrcoef(rf[:, :-1].ravel(), rf[:, 1:].ravel())[0, 1]) < Z_BAND / math.sqrt(tot)
        assert abs(np.corrcoef(rf[:-1].ravel(), rf[1:].ravel())[0, 1]) < Z_BAND / math.sqrt(tot)
        # pair table of horizontally adjacent entries: 9 cells with product probabilities
        pair = (r[:, :-1].astype(np.int64) + 1) * 3 + (r[:, 1:].astype(np.int64) + 1)
        pc = np.bincount(pair.ravel(), minlength=9).astype(np.float64)
        pp = np.outer([0.25, 0.5, 0.25], [0.25, 0.5, 0.25]).ravel()
        chi = ((pc - pp * pair.size) ** 2 / (pp * pair.size)).sum()
        assert chi < 8 + Z_BAND * math.sqrt(16.0), chi
        a_bar = a[:, : gp.m_bar]
        assert a_bar.min() >= 0 and a_bar.max() < q
        buckets = (a_bar.astype(np.float64) * 256.0 / q).astype(np.int64)
        bc = np.bincount(buckets.ravel(), minlength=256).astype(np.float64)
        # bucket widths differ by at most one residue when 256 does not divide q
        edges = np.ceil(np.arange(257) * q / 256.0)
        pb = np.diff(edges) / q
        chi = ((bc - pb * a_bar.size) ** 2 / (pb * a_bar.size)).sum()
        assert chi < 255 + Z_BAND * math.sqrt(510.0), chi
        # low bits too (a multiply-shift map must not bias them)
        lc = np.bincount((a_bar & 255).ravel(), minlength=256).astype(np.float64)
        pl = np.array([((q - 1 - v) // 256 + 1) / q for v in range(256)])
        chi = ((lc - pl * a_bar.size) ** 2 / (pl * a_bar.size)).sum()
        assert chi < 255 + Z_BAND * math.sqrt(510.0), chi


def test_ring_trapdoor_coefficients_law(T):
    """SampleZ::sample for the ring trapdoor (trapdoor_distribution.rs:112-122): r, e coefficients ~ D_{Z, s_td}."""
    gr = T.GadgetParametersRing.init_default(256, 3329)
    s_td = 1.005
    vals = []
    for seed in range(40):
        psf = T.PSFGPVRing(gr, 500.0, s_td)
        _, (r, e) = psf.trap_gen(seed=seed)
        vals.append(np.concatenate([r.ravel(), e.ravel()]))
    v = np.concatenate(vals).astype(np.int64)
    xs, pm = O.dgauss_pmf(s_td, 0.0)
    ok, chi, dof, outside = chi2_gof(v, xs, pm)
    assert outside == 0 and ok, (chi, dof)
