"""Key / parameter import: the serde-JSON + typetag layout of the reference's structs and the FLINT text
forms of qfall-math values, on hand-written fixtures (tests/golden/serde_fixtures.json).  CPU part: parsing, writers,
round trips; GPU part: a key loaded from its serialised form drives samp_p / f_a."""
import json
import os
from fractions import Fraction

import numpy as np
import pytest

from oracle import qfall_oracle as O
from tools_b200 import gadget, serde_io as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fx():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "serde_fixtures.json")))


def test_flint_text_forms(fx):
    a, q = S.parse_mat_zq(fx["mat_zq_2x13"])
    assert q == 8 and a.shape == (2, 13) and a.dtype == np.int64 and a[1, 11] == 7 and a.min() >= 0
    # the same strings through the oracle's independent parser
    assert a.tolist() == [[x % 8 for x in row] for row in O.parse_matz(fx["mat_zq_2x13"]["matrix"])[0]]
    r = S.parse_mat_z(fx["mat_z_7x6"])
    assert r.shape == (7, 6) and r[0].tolist() == [0, 1, 0, 1, 1, 0] and r[1, 0] == -1
    mq = S.parse_mat_q(fx["mat_q_2x2"])
    assert mq.tolist() == [[0.5, -3.0], [7.0, 22 / 7]]
    assert S.parse_poly_over_z(fx["poly_over_z"]) == [2, 8, 8, 12]
    assert S.parse_poly_over_z("0") == [] and S.parse_poly_over_z("1  -7") == [-7]
    mp = S.parse_mat_poly_over_z(fx["mat_poly_over_z"], 4)
    assert mp.shape == (2, 3, 4) and mp[0, 0].tolist() == [1, 2, 0, 0] and mp[0, 2].tolist() == [0, 0, 0, 0]
    assert mp[1, 0].tolist() == [0, 0, 0, -3] and mp[1, 2].tolist() == [1, 0, 1, 0]
    mat, n, q = S.parse_mat_polynomial_ring_zq(fx["mat_polynomial_ring_zq"])
    assert (n, q) == (4, 16) and mat.shape == (1, 3, 4) and mat[0, 1].tolist() == [3, 0, 15, 2]
    coeffs, q = S.parse_modulus_polynomial_ring_zq({"poly": "5  1 0 0 0 1 mod 16"})
    assert coeffs == [1, 0, 0, 0, 1] and q == 16
    assert S.parse_z({"value": "-12"}) == -12 and S.parse_z("7") == 7 and S.parse_z(9) == 9
    assert S.parse_q({"value": "201/200"}) == Fraction(201, 200) and S.parse_q("3") == 3
    for bad in ("[1, 2]", "[[1, 2],[3]] mod", "3  1 2"):
        with pytest.raises((S.SerdeError, ValueError)):
            S.parse_mat_z(bad) if bad.startswith("[") else S.parse_poly_over_z(bad)
    with pytest.raises(S.SerdeError):
        S.parse_mat_zq("[[1, 2]]")  # no modulus
    with pytest.raises(S.SerdeError):
        S.parse_mat_zq("[[1, 2]] mod 4611686018427387904")  # 2^62


def test_writers_round_trip():
    rng = np.random.default_rng(1)
    a = rng.integers(0, 97, (3, 5))
    assert S.parse_mat_zq(S.fmt_mat(a, 97))[0].tolist() == a.tolist()
    z = rng.integers(-50, 50, (4, 4))
    assert S.parse_mat_z(S.fmt_mat(z)).tolist() == z.tolist()
    qm = rng.normal(size=(3, 3))
    assert np.array_equal(S.parse_mat_q(S.fmt_mat_q(qm)), qm)  # doubles are dyadic rationals: exact
    pm = rng.integers(-3, 4, (2, 3, 8))
    pm[0, 1] = 0
    assert np.array_equal(S.parse_mat_poly_over_z(S.fmt_mat_poly(pm), 8), pm)
    assert S.fmt_poly([0, 0, 0]) == "0" and S.fmt_poly([2, 8, 8, 12, 0]) == "4  2 8 8 12"


def test_parameter_structs(fx):
    gp = S.load_gadget_parameters(fx["gadget_parameters_2_8"])
    assert gp == gadget.GadgetParameters.init_default(2, 8)
    spec = S.load_psf_spec(fx["psf_gpv_2_8"])
    assert spec.kind == "gpv" and spec.gp == gp and spec.s == 10
    spec = S.load_psf_spec(json.dumps(fx["psf_perturbation_readme"]))
    assert spec.kind == "perturbation" and spec.gp == gadget.GadgetParameters.init_default(8, 64) and (spec.r, spec.s) == (3, 25)
    spec = S.load_psf_spec(fx["psf_gpv_ring_4_16"])
    assert spec.kind == "ring" and spec.gp == gadget.GadgetParametersRing.init_default(4, 16) and spec.s_td == Fraction(201, 200)
    spec = S.load_psf_spec(fx["bare_forms"])  # bare scalars / strings instead of one-field objects
    assert spec.gp == gp and spec.s == Fraction(21, 2)
    with pytest.raises(S.SerdeError):
        S.load_gadget_parameters(fx["unsupported_distribution"])
    with pytest.raises(S.SerdeError):  # a classical struct with the ring distribution tag
        S.load_gadget_parameters(dict(fx["gadget_parameters_2_8"], distribution={"SampleZ": None}))
    # writers produce what the loaders read
    for name in ("psf_gpv_2_8", "psf_perturbation_readme", "psf_gpv_ring_4_16"):
        spec = S.load_psf_spec(fx[name])
        again = S.load_psf_spec(json.loads(json.dumps(S.dump_psf(spec))))
        assert (again.kind, again.gp, again.s, again.r, again.s_td) == (spec.kind, spec.gp, spec.s, spec.r, spec.s_td)
        assert S.dump_psf(spec) == fx[name]


def test_key_forms_round_trip_without_device(fx):
    """load_key / dump_key on the fixed (n = 2, q = 8) fixture of short_basis_classical.rs:273-349 and on synthetic
    trapdoors of all three PSF kinds (no device: PsfSpec)."""
    spec = S.load_psf_spec(fx["psf_gpv_2_8"])
    a, _ = S.load_key(spec, fx["mat_zq_2x13"])
    r = S.parse_mat_z(fx["mat_z_7x6"])
    po = O.GadgetParameters.init_default(2, 8)
    basis = np.array(O.gen_short_basis_for_trapdoor(po, O.mat_identity(2), a.tolist(), r.tolist()), dtype=np.int64)
    gso = O.gso_f64(basis.astype(np.float64))
    a_obj, td_obj = S.dump_key(spec, a, (basis, gso))
    a2, (b2, g2) = S.load_key(spec, json.dumps(a_obj), json.dumps(td_obj))
    assert np.array_equal(a2, a) and np.array_equal(b2, basis) and np.array_equal(g2, gso)
    assert not ((a2.astype(object) @ b2.astype(object)) % 8).any()  # the loaded basis lies in Lambda^perp(A)
    with pytest.raises(S.SerdeError):
        S.load_key(spec, {"matrix": S.fmt_mat(a, 16)})  # wrong modulus
    # perturbation
    ps = S.load_psf_spec(fx["psf_perturbation_readme"])
    rng = np.random.default_rng(2)
    gp = ps.gp
    ap = rng.integers(0, 64, (gp.n, gp.m))
    rp = rng.integers(-1, 2, (gp.m_bar, gp.n * gp.k)).astype(np.int8)
    lp = np.tril(rng.normal(size=(gp.m, gp.m)))
    sb = np.array(O.short_basis_gadget(O.GadgetParameters.init_default(8, 64)), dtype=np.int64)
    sg = O.gso_f64(sb.astype(np.float64))
    ao, to = S.dump_key(ps, ap, (rp, lp, (sb, sg)))
    a3, (r3, l3, (sb3, sg3)) = S.load_key(ps, json.dumps(ao), json.dumps(to))
    assert np.array_equal(a3, ap) and np.array_equal(r3, rp) and r3.dtype == np.int8 and np.array_equal(l3, lp)
    assert np.array_equal(sb3, sb) and np.array_equal(sg3, sg)
    # ring
    rs = S.load_psf_spec(fx["psf_gpv_ring_4_16"])
    ar = rng.integers(0, 16, (rs.gp.k + 2, 4))
    rr, er = rng.integers(-2, 3, (rs.gp.k, 4)), rng.integers(-2, 3, (rs.gp.k, 4))
    ao, to = S.dump_key(rs, ar, (rr, er))
    a4, (r4, e4) = S.load_key(rs, json.dumps(ao), json.dumps(to))
    assert np.array_equal(a4, ar) and np.array_equal(r4, rr) and np.array_equal(e4, er) and r4.dtype == np.int32


@pytest.mark.gpu
def test_loaded_keys_drive_the_backend(fx):
    """A PSF and its key, serialised in the reference's layout and loaded back, give exact preimages on the device."""
    import tools_b200 as T

    rng = np.random.default_rng(3)
    for name, n, q in (("psf_perturbation_readme", 8, 64), ("psf_gpv_2_8", 2, 8), ("psf_gpv_ring_4_16", 4, 16)):
        psf0 = S.load_psf(fx[name])
        a, td = psf0.trap_gen(seed=5)
        if name == "psf_gpv_2_8" and td[1] is None:
            td = (td[0], psf0.gso(td[0]))
        text_psf = json.dumps(S.dump_psf(psf0))
        a_obj, td_obj = S.dump_key(psf0, a, td)
        text_a, text_td = json.dumps(a_obj), json.dumps(td_obj)
        psf = S.load_psf(text_psf)  # a fresh context from the serialised struct
        a2, td2 = S.load_key(psf, text_a, text_td)
        u = rng.integers(0, q, (64, n), dtype=np.int64)
        e = psf.samp_p_batch(a2, td2, u, seed=7)
        u2, fl = psf.f_a_batch(a2, e)
        assert np.array_equal(u2, u) and fl.all()
        if name == "psf_gpv_ring_4_16":
            assert O.f_a_ring(np.asarray(a).tolist(), e[0].tolist(), n, q) == u[0].tolist()
        else:
            assert np.array_equal(O.f_a_classical_batch(np.asarray(a), e, q), u)
