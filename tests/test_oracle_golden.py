"""Pin the CPU oracle against every golden vector the reference's own unit tests
hold for the hot path.  CPU only."""
import numpy as np
import pytest

from oracle import qfall_oracle as O


def lit(goldens, key, i=0):
    return goldens[key]["literals"][i]


def test_gen_gadget_vec(goldens):
    # gadget_classical.rs:296-311
    assert O.gen_gadget_vec(5, 2) == O.parse_matz(lit(goldens, "gadget_classical::correctness_base_2"))[0]
    assert O.gen_gadget_vec(4, 5) == O.parse_matz(lit(goldens, "gadget_classical::correctness_base_5"))[0]


def test_gen_gadget_mat(goldens):
    # gadget_classical.rs:323-345
    assert O.gen_gadget_mat(3, 3, 2) == O.parse_matz(lit(goldens, "gadget_classical::correctness_base_2_3x3"))[0]
    assert O.gen_gadget_mat(2, 5, 3) == O.parse_matz(lit(goldens, "gadget_classical::correctness_base_3_2x5"))[0]


def test_short_basis_gadget(goldens):
    # gadget_classical.rs:491-572
    p = O.GadgetParameters.init_default(2, 16)
    assert O.short_basis_gadget(p) == O.parse_matz(lit(goldens, "gadget_classical::base_2_power_two"))[0]
    p = O.GadgetParameters.init_default(1, 0b1100110)
    assert O.short_basis_gadget(p) == O.parse_matz(lit(goldens, "gadget_classical::base_2_arbitrary"))[0]
    p = O.GadgetParameters.init_default(1, 625)
    p.k, p.base = 4, 5
    assert O.short_basis_gadget(p) == O.parse_matz(lit(goldens, "gadget_classical::base_5_power_5"))[0]
    q = int(lit(goldens, "gadget_classical::base_5_arbitrary", 0), 5)
    p = O.GadgetParameters.init_default(1, q)
    p.k, p.base = 4, 5
    assert O.short_basis_gadget(p) == O.parse_matz(lit(goldens, "gadget_classical::base_5_arbitrary", 1))[0]


def test_find_solution_gadget(goldens):
    # gadget_classical.rs:449-479
    g = [x[0] for x in O.gen_gadget_vec(5, 3)]
    for i in range(124):
        sol = O.find_solution_gadget_vec(i, 125, 5, 3)
        assert sum(a * b for a, b in zip(g, sol)) == i
        assert all(0 <= d < 3 for d in sol)
    value, q = O.parse_matz(lit(goldens, "gadget_classical::returns_correct_solution_mat"))
    sol = O.find_solution_gadget_mat(value, q, 5, 3)
    assert O.mat_mul(O.gen_gadget_mat(3, 5, 3), sol) == [[x % q for x in r] for r in value]
    with pytest.raises(ValueError):
        O.find_solution_gadget_vec(3, 126, 2, 3)  # 3^2 < 126


def test_default_parameters():
    # gadget_parameters.rs:194-212, gadget_default.rs:114-132
    for n in [5, 10, 50, 100]:
        for k in [5, 10, 25]:
            q = 2**k
            gp = O.GadgetParameters.init_default(n, q)
            assert (gp.base, gp.k, gp.n, gp.q) == (2, k, n, q)
            assert gp.m_bar == n * k + O.ceil_log(n, 2) ** 2
            assert gp.m == gp.m_bar + n * k
    gp = O.GadgetParameters.init_default(8, 64)
    assert (gp.k, gp.m_bar, gp.m) == (6, 57, 105)
    gp = O.GadgetParameters.init_default(256, 2**24)
    assert (gp.k, gp.m_bar, gp.m) == (24, 6208, 12352)
    gr = O.GadgetParametersRing.init_default(256, 3329)
    assert (gr.k, gr.m_bar) == (12, 14)


def test_classical_sa_l_sa_r(goldens):
    # short_basis_classical.rs:273-349
    a, q = O.parse_matz(lit(goldens, "short_basis_classical::get_fixed_trapdoor_for_tag_identity", 0))
    r, _ = O.parse_matz(lit(goldens, "short_basis_classical::get_fixed_trapdoor_for_tag_identity", 1))
    assert q == 8
    p = O.GadgetParameters.init_default(2, 8)
    assert O.gen_sa_l(r) == O.parse_matz(lit(goldens, "short_basis_classical::working_sa_l"))[0]
    tag = O.mat_identity(2)
    assert O.gen_sa_r(p, tag, a) == O.parse_matz(lit(goldens, "short_basis_classical::working_sa_r_identity"))[0]
    # short_basis_classical.rs:366-385: G W = -A [I|0]^t
    w = O.compute_w(p, tag, a)
    gw = O.mat_mul(O.gen_gadget_mat(2, p.k, 2), w, q)
    assert gw == [[(-x) % q for x in row[: p.m_bar]] for row in a]


def test_rot_minus(goldens):
    # rotation_matrix.rs:106-134
    col, _ = O.parse_matz(lit(goldens, "rotation_matrix::correct_rotation_matrix_vec", 0))
    row, _ = O.parse_matz(lit(goldens, "rotation_matrix::correct_rotation_matrix_vec", 1))
    cmp_, _ = O.parse_matz(lit(goldens, "rotation_matrix::correct_rotation_matrix_vec", 2))
    assert O.rot_minus(col) == cmp_ and O.rot_minus(row) == cmp_
    big = 2**64 - 1
    mat, _ = O.parse_matz(lit(goldens, "rotation_matrix::correct_rotation_matrix_mat", 0).replace("{}", str(big)))
    cmp_, _ = O.parse_matz(lit(goldens, "rotation_matrix::correct_rotation_matrix_mat", 1).replace("{}", str(big)))
    assert O.rot_minus_matrix(mat) == cmp_
    with pytest.raises(ValueError):
        O.rot_minus([[1, 5, -1, 9], [1, 2, 3, 4]])


def test_rot_minus_is_ring_product():
    rng = np.random.default_rng(1)
    n, q = 16, 3329
    a = [int(x) for x in rng.integers(0, q, n)]
    b = [int(x) for x in rng.integers(-50, 50, n)]
    via_rot = [sum(r[j] * b[j] for j in range(n)) % q for r in O.rot_minus(a)]
    assert via_rot == O.ring_mul(a, b, n, q) == O.ring_mul_np(a, b, n, q)


def test_ring_compute_s(goldens):
    # short_basis_ring.rs:455-535
    p = O.GadgetParametersRing.init_default(8, 16)
    assert O.ring_compute_s(p) == O.parse_matpoly(lit(goldens, "short_basis_ring::base_2_power_two"))
    p = O.GadgetParametersRing.init_default(1, 0b1100110)
    assert O.ring_compute_s(p) == O.parse_matpoly(lit(goldens, "short_basis_ring::base_2_arbitrary"))
    p = O.GadgetParametersRing.init_default(1, 625)
    p.k, p.base = 4, 5
    assert O.ring_compute_s(p) == O.parse_matpoly(lit(goldens, "short_basis_ring::base_5_power_5"))
    q = int(lit(goldens, "short_basis_ring::base_5_arbitrary", 0), 5)
    p = O.GadgetParametersRing.init_default(1, q)
    p.k, p.base = 4, 5
    assert O.ring_compute_s(p) == O.parse_matpoly(lit(goldens, "short_basis_ring::base_5_arbitrary", 1))


def test_ring_sa_l_sa_r(goldens):
    # short_basis_ring.rs:357-444
    key = "short_basis_ring::get_fixed_trapdoor"
    a = O.parse_matpoly(lit(goldens, key, 0))[0]
    r = O.parse_matpoly(lit(goldens, key, 1))[0]
    e = O.parse_matpoly(lit(goldens, key, 2))[0]
    p = O.GadgetParametersRing.init_default(4, 16)
    # the reference test passes (r, e) into gen_sa_l(e, r): first argument is the top row
    assert O.ring_gen_sa_l(r, e) == O.parse_matpoly(lit(goldens, "short_basis_ring::working_sa_l"))
    sa_r = O.ring_gen_sa_r(p, a)
    sa_r = [[O.poly_trim(O.poly_reduce_anticyclic(x, 4)) for x in row] for row in sa_r]
    assert O.coeff_embed(sa_r, 4) == O.parse_matz(lit(goldens, "short_basis_ring::working_sa_r"))[0]


def test_find_solution_gadget_ring(goldens):
    # gadget_ring.rs:225-239
    p = O.GadgetParametersRing.init_default(3, 32)
    u = O.parse_poly(lit(goldens, "gadget_ring::is_correct_solution"))
    sol = O.find_solution_gadget_ring(u, p)
    acc = [0] * p.n
    for j in range(p.k):
        acc = [(x + (2**j) * (sol[j][t] if t < len(sol[j]) else 0)) % p.q for t, x in enumerate(acc)]
    assert acc == O.poly_reduce_anticyclic(u, p.n, p.q)


# ----- known-answer properties on random instances ---------------


@pytest.mark.parametrize("n,q", [(5, 32), (7, 127), (4, 100)])
def test_trapdoor_identity(n, q):
    # gadget_classical.rs:362-414: A [R; I] = H G
    rng = np.random.default_rng(n * q)
    p = O.GadgetParameters.init_default(n, q)
    a_bar = rng.integers(0, q, (n, p.m_bar)).tolist()
    r = O.sample_pm_one_zero(rng, p.m_bar, n * p.k)
    tag = O.mat_identity(n)
    for i in range(n):
        for j in range(i + 1, n):
            tag[i][j] = int(rng.integers(0, q))
    a = O.gen_trapdoor(p, a_bar, tag, r)
    td = r + O.mat_identity(n * p.k)
    assert O.mat_mul(a, td, q) == O.mat_mul(tag, O.gen_gadget_mat(n, p.k, 2), q)
    # short_basis_classical.rs:128-188: basis columns in the kernel lattice
    sb = O.gen_short_basis_for_trapdoor(p, tag, a, r)
    assert all(x == 0 for row in O.mat_mul(a, sb, q) for x in row)
    # short_basis_classical.rs:193-242: GSO length bound
    gso = O.gso_f64(np.array(sb, dtype=np.float64))
    bound = (np.sqrt(p.m_bar) + 1) * (2 if 2**p.k == q else np.sqrt(5))
    assert np.all(np.linalg.norm(gso, axis=0) <= bound + 1e-9)


@pytest.mark.parametrize("n,q", [(5, 16), (6, 32), (4, 42)])
def test_ring_trapdoor_and_basis(n, q):
    # gadget_ring.rs:190-211, short_basis_ring.rs:183-219, :553-571
    rng = np.random.default_rng(n + q)
    p = O.GadgetParametersRing.init_default(n, q)
    a_bar = rng.integers(0, q, n).tolist()
    r = [rng.integers(-5, 6, n).tolist() for _ in range(p.k)]
    e = [rng.integers(-5, 6, n).tolist() for _ in range(p.k)]
    a = O.gen_trapdoor_ring_lwe(p, a_bar, r, e)
    for j in range(p.k):  # A [e; r; I] = g
        acc = O.ring_mul(a[0], e[j], n, q)
        acc = [(x + y) % q for x, y in zip(acc, O.ring_mul(a[1], r[j], n, q))]
        acc = [(x + y) % q for x, y in zip(acc, a[2 + j])]
        assert acc == [(2**j) % q] + [0] * (n - 1)
    basis = O.gen_short_basis_for_trapdoor_ring(p, a, r, e)
    assert len(basis) == p.k + 2 and len(basis[0]) == n * (p.k + 2)
    for c in range(len(basis[0])):
        col = [basis[i][c] for i in range(p.k + 2)]
        assert all(len(x) <= n for x in col)
        assert O.f_a_ring(a, col, n, q) == [0] * n
    w = O.ring_compute_w(p, a)
    for c in range(2):
        acc = [0] * n
        for j in range(p.k):
            acc = [(x + 2**j * (w[j][c][t] if t < len(w[j][c]) else 0)) % q for t, x in enumerate(acc)]
        assert acc == [(-x) % q for x in a[c]]


def test_compression_round_trip_and_formula():
    # lossy_compression_fips203.rs:281-338: round-trip bound, d = 0 rejected
    for q, d in [(257, 4), (3329, 11), (3329, 1), (3329, 4), (3329, 10)]:
        xs = list(range(q))
        c = O.lossy_compress(xs, d, q)
        assert all(0 <= y < 2**d for y in c)
        back = O.lossy_decompress(c, d, q)
        bound = 2 ** (O.ceil_log(q, 2) - d - 1)
        for x, y in zip(xs, back):
            dist = abs(x - y)
            dist = min(dist, q - dist)
            assert dist <= bound
        xs_np = np.arange(q)
        assert O.lossy_compress_np(xs_np, d, q).tolist() == c
        assert O.lossy_decompress_np(np.array(c), d, q).tolist() == back
    # FIPS 203 Compress_d is round-half-up(2^d/q * x) mod 2^d for odd q
    for x in range(3329):
        assert O.compress_coeff(x, 10, 3329) == int((x * 1024 * 2 + 3329) // (2 * 3329)) % 1024
    with pytest.raises(ValueError):
        O.lossy_compress([1], 0, 3329)
    with pytest.raises(ValueError):
        O.lossy_decompress([1], 0, 3329)


def test_byte_encode_decode_known_vectors():
    """FIPS 203 Algorithms 5 / 6 (not part of the reference crate): hand-computed byte layouts,
    inverse property for every d, vectorised form == literal form."""
    f = [0] * 256
    f[0], f[1] = 1, 2
    assert O.byte_encode(f, 4)[:2] == bytes([0x21, 0x00])          # two 4-bit values per byte, low nibble first
    f[0], f[1] = 0xABC, 0x123
    assert O.byte_encode(f, 12)[:3] == bytes([0xBC, 0x3A, 0x12])   # the 12-bit layout of ML-KEM's ByteEncode_12
    f = [0] * 256
    f[0], f[1], f[2] = 0b101, 0b011, 0b111
    assert O.byte_encode(f, 3)[:2] == bytes([0xDD, 0x01])
    rng = np.random.default_rng(5)
    for d in range(1, 13):
        m = (1 << d) if d < 12 else 3329
        g = rng.integers(0, m, (3, 256))
        for row in g:
            b = O.byte_encode(row.tolist(), d)
            assert len(b) == 32 * d and O.byte_decode(b, d, 3329) == row.tolist()
        enc = O.byte_encode_np(g, d)
        assert [bytes(r) for r in enc] == [O.byte_encode(row.tolist(), d) for row in g]
        assert np.array_equal(O.byte_decode_np(enc, d, 3329), g)
    # ByteDecode_12 reduces mod q
    assert O.byte_decode(bytes([0xFF] * 384), 12, 3329) == [4095 % 3329] * 256


def test_reference_samp_p_restatements():
    """gpv.rs:253-268, mp_perturbation.rs:432-448, gpv_ring.rs:317-334 on the oracle."""
    rng = np.random.default_rng(7)
    # PSFPerturbation (README example: n=8, q=64, r=3, s=25)
    n, q, r_par, s = 8, 64, 3.0, 25.0
    p = O.GadgetParameters.init_default(n, q)
    a_bar = rng.integers(0, q, (n, p.m_bar)).tolist()
    rm = O.sample_pm_one_zero(rng, p.m_bar, n * p.k)
    a = O.gen_trapdoor(p, a_bar, O.mat_identity(n), rm)
    l = O.compute_sqrt_sigma_2(rm, s, r_par, 2)
    sb = O.short_basis_gadget(p)
    sgso = O.gso_f64(np.array(sb, dtype=np.float64))
    for _ in range(3):
        u = rng.integers(0, q, n).tolist()
        e = O.samp_p_perturbation(rng, p, a, rm, l, sb, sgso, u, r_par)
        assert O.f_a_classical(a, e, q) == u
        assert O.check_domain_perturbation(e, p.m, s, r_par)
    # PSFGPV n=5, q=256... use q=32 to keep it quick
    n, q, s = 5, 32, 10.0
    p = O.GadgetParameters.init_default(n, q)
    a_bar = rng.integers(0, q, (n, p.m_bar)).tolist()
    rm = O.sample_pm_one_zero(rng, p.m_bar, n * p.k)
    tag = O.mat_identity(n)
    a = O.gen_trapdoor(p, a_bar, tag, rm)
    sb = O.gen_short_basis_for_trapdoor(p, tag, a, rm)
    sgso = O.gso_f64(np.array(sb, dtype=np.float64))
    for _ in range(3):
        u = rng.integers(0, q, n).tolist()
        e = O.samp_p_gpv(rng, a, q, sb, sgso, u, s)
        assert O.f_a_classical(a, e, q) == u
        assert O.check_domain_gpv(e, p.m, s)
    # PSFGPVRing n=4, q=2^31-1
    n, q = 4, 2**31 - 1
    pr = O.GadgetParametersRing.init_default(n, q)
    s = ((2 * 2 * 1.005 * np.sqrt(n) + 1) * 2) * 4
    a_bar = rng.integers(0, q, n).tolist()
    rr = [[O.sample_z(rng, 1.005, 0) for _ in range(n)] for _ in range(pr.k)]
    ee = [[O.sample_z(rng, 1.005, 0) for _ in range(n)] for _ in range(pr.k)]
    a = O.gen_trapdoor_ring_lwe(pr, a_bar, rr, ee)
    u = rng.integers(0, q, n).tolist()
    e = O.samp_p_gpv_ring(rng, pr, a, rr, ee, u, float(s))
    assert O.f_a_ring(a, e, n, q) == u
    assert O.check_domain_ring(e, n, pr.k, float(s))


def test_sample_z_matches_pmf():
    rng = np.random.default_rng(11)
    s, c = 3.0, 0.3
    xs, pm = O.dgauss_pmf(s, c)
    draws = np.array([O.sample_z(rng, s, c) for _ in range(20000)])
    cnt = np.array([(draws == x).sum() for x in xs])
    keep = pm * len(draws) >= 5
    chi = (((cnt - pm * len(draws)) ** 2) / (pm * len(draws)))[keep].sum()
    dof = keep.sum() - 1
    assert chi < dof + 5 * np.sqrt(2 * dof)
