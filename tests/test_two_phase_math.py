"""CPU checks (numpy + the oracle) of the identities behind the two-phase form of PSFGPV::samp_p (DESIGN.md 4.3,
tools_b200/csrc/api.cu samp_p_np2_chunk).  With the randomized rounding replaced by z_i = floor(c'_i + alpha_i) for one
fixed offset alpha_i in (0, 1) per coordinate (common random numbers: like D_{Z,s',c'} this rule commutes with integer
shifts of the centre, and it has no ties) the recursion is a deterministic function of the coset of u, so the one-pass
loop of the reference (gpv.rs:152-161: centre -sol, all m coordinates) and the two-phase form (gadget preimage as
centre, residual reduced modulo the [R;I]S' sub-lattice between the phases) must return the SAME preimage."""
import numpy as np
import pytest

from oracle import qfall_oracle as O


def _key(n, q, seed):
    p = O.GadgetParameters.init_default(n, q)
    rng = np.random.default_rng(seed)
    a_bar = rng.integers(0, q, (n, p.m_bar)).tolist()
    r = O.sample_pm_one_zero(rng, p.m_bar, p.n * p.k)
    tag = O.mat_identity(n)
    a = O.gen_trapdoor(p, a_bar, tag, r)
    s = O.gen_short_basis_for_trapdoor(p, tag, a, r)
    return p, np.array(a, dtype=np.int64), np.array(r, dtype=np.int64), np.array(s, dtype=np.int64), rng


def _gso(s):
    qm, rm = np.linalg.qr(s.astype(np.float64))
    d = np.diag(rm)
    u = rm / d[:, None]            # U_ij = <b_j, b~_i> / ||b~_i||^2, upper unitriangular
    mt = (qm / d[None, :]).T       # row i = b~_i / ||b~_i||^2
    return u, mt


def _digits(h, k, base):
    g = np.zeros((len(h), k), dtype=np.int64)
    v = h.copy()
    for t in range(k):
        g[:, t] = v % base
        v //= base
    return g.reshape(-1)


def _sprime(p):
    sp = np.array(O.short_basis_gadget(p), dtype=np.int64)
    return sp[:, ::-1] if p.base**p.k == p.q else sp   # short_basis_classical.rs:80-82


@pytest.mark.parametrize("n,q", [(3, 2**6), (4, 2**5), (3, 53), (5, 97)])
def test_two_phase_identities_and_equivalence_under_common_random_numbers(n, q):
    p, a, r, s, rng = _key(n, q, seed=n * 1000 + q)
    nk, mb, m, k = p.n * p.k, p.m_bar, p.m, p.k
    assert not ((a @ s) % q).any()
    sp = _sprime(p)
    # the basis has the G-trapdoor form and the key the form [A_bar | G - A_bar R]
    assert np.array_equal(s[:mb, :nk], r @ sp) and np.array_equal(s[mb:, :nk], sp)
    gmat = np.array(O.gen_gadget_mat(p.n, p.k, p.base), dtype=np.int64)
    assert np.array_equal((a[:, :mb] @ r + a[:, mb:]) % q, gmat % q)
    u_mat, mt = _gso(s)
    u11 = np.triu(u_mat[:nk, :nk])
    # M' = Mt_1 [R; I] = U_11 S'^-1, block triangular
    mprime = mt[:nk, :mb] @ r + mt[:nk, mb:]
    assert np.allclose(mprime, u11 @ np.linalg.inv(sp.astype(np.float64)), atol=1e-9)
    for trial in range(6):
        u = rng.integers(0, q, n)
        alpha = rng.random(m)
        # ---- reference form: some solution of A x = u (free variables 0), all m coordinates
        sol = np.array(O.solve_unit_pivots(a.tolist(), u.tolist(), q), dtype=np.int64)
        c = -(mt @ sol)
        z = np.zeros(m)
        for i in range(m - 1, -1, -1):
            z[i] = np.floor(c[i] + alpha[i])
            c[:i] -= u_mat[:i, i] * z[i]
        e_ref = sol + s @ z.astype(np.int64)
        # ---- two-phase form
        g = _digits(u.copy(), k, p.base)
        x = np.concatenate([r @ g, g])
        assert np.array_equal((a @ x) % q, u)
        assert np.allclose((mt @ x)[nk:], 0.0, atol=1e-9)      # the gadget preimage has no GSO coordinates above nk
        z2 = np.zeros(mb)
        c2 = np.zeros(mb)
        for i in range(mb - 1, -1, -1):
            z2[i] = np.floor(c2[i] + alpha[nk + i])
            c2[:i] -= u_mat[nk:nk + i, nk + i] * z2[i]
        z2 = z2.astype(np.int64)
        h = (u - a[:, :mb] @ z2) % q
        g3 = _digits(h.copy(), k, p.base)
        t1 = -(mprime @ g3) - mt[:nk, :mb] @ z2
        z1 = np.zeros(nk)
        for i in range(nk - 1, -1, -1):
            z1[i] = np.floor(t1[i] + alpha[i])
            t1[:i] -= u11[:i, i] * z1[i]
        e_bot = g3 + sp @ z1.astype(np.int64)
        e_top = z2 + r @ e_bot
        e = np.concatenate([e_top, e_bot])
        assert np.array_equal((a @ e) % q, u)
        # the same preimage
        assert np.array_equal(e, e_ref), (trial, np.abs(e - e_ref).max())


def test_balanced_digit_bias_identity():
    """np_diag2 extracts digit plane l of z directly: v_l = floor((v + 128 (256^l - 1) / 255) / 256^l), d_l = ((v_l + 128) mod
    256) - 128, the top plane keeping all of v_{L-1} (lattice.cu).  Must equal the sequential balanced base-256 split
    (carry from plane to plane) used everywhere else, and recombine to v."""
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.integers(-2**40, 2**40, 4000), np.arange(-70000, 70000, 37), [127, 128, -128, -129, 32639, -32640]])
    for L in (1, 2, 3, 4, 5):
        for v in vals.tolist():
            if abs(v) > sum(127 * 256**i for i in range(L)):
                continue
            seq, w = [], v
            for l in range(L):
                d = ((w + 128) & 255) - 128 if l < L - 1 else w
                seq.append(d)
                w = (w - d) >> 8
            direct, bias = [], 0
            for l in range(L):
                vl = (v + bias) >> (8 * l)
                direct.append(vl if l == L - 1 else ((vl + 128) & 255) - 128)
                bias += 128 << (8 * l)
            assert direct == seq and all(-128 <= d <= 127 for d in seq), (v, L, seq, direct)
            assert sum(d * 256**i for i, d in enumerate(seq)) == v
