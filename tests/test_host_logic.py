"""CPU tests: host-side mirror (tools_b200.gadget) against the oracle, and the C-ABI
library loads and exports every symbol include/qfall_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import qfall_oracle as O
from tools_b200 import gadget as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parameters_match_oracle():
    for n, q in [(5, 32), (8, 64), (42, 42), (256, 2**24), (512, 2**32 - 5)]:
        a, b = G.GadgetParameters.init_default(n, q), O.GadgetParameters.init_default(n, q)
        assert (a.n, a.k, a.m_bar, a.base, a.q, a.m) == (b.n, b.k, b.m_bar, b.base, b.q, b.m)
        ar, br = G.GadgetParametersRing.init_default(n, q), O.GadgetParametersRing.init_default(n, q)
        assert (ar.n, ar.k, ar.m_bar, ar.base, ar.q) == (br.n, br.k, br.m_bar, br.base, br.q)


def test_gadget_functions_match_oracle(goldens):
    assert G.gen_gadget_vec(5, 2).tolist() == O.gen_gadget_vec(5, 2)
    assert G.gen_gadget_mat(3, 3, 2).tolist() == O.gen_gadget_mat(3, 3, 2)
    for n, q in [(2, 16), (1, 0b1100110), (3, 127)]:
        p, po = G.GadgetParameters.init_default(n, q), O.GadgetParameters.init_default(n, q)
        assert G.short_basis_gadget(p).tolist() == O.short_basis_gadget(po)
    rng = np.random.default_rng(0)
    val = rng.integers(0, 125, (3, 4))
    assert G.find_solution_gadget_mat(val, 125, 5, 3).tolist() == O.find_solution_gadget_mat(val.tolist(), 125, 5, 3)
    with pytest.raises(ValueError):
        G.find_solution_gadget_mat(val, 126, 2, 3)
    v = [[1], [5], [-1], [9]]
    assert G.rot_minus(v).tolist() == O.rot_minus(v)
    mat = [[1, 5, -1, 9], [2**64 - 1, 1, 2, 3]]
    assert G.rot_minus_matrix(mat).tolist() == O.rot_minus_matrix(mat)
    with pytest.raises(ValueError):
        G.rot_minus([[1, 2], [3, 4]])


@pytest.mark.parametrize("n,q,with_tag", [(2, 8, False), (5, 32, False), (4, 100, True), (6, 127, False)])
def test_classical_short_basis_matches_oracle(n, q, with_tag, goldens):
    rng = np.random.default_rng(n * q)
    p, po = G.GadgetParameters.init_default(n, q), O.GadgetParameters.init_default(n, q)
    a_bar = rng.integers(0, q, (n, p.m_bar))
    r = np.array(O.sample_pm_one_zero(rng, p.m_bar, n * p.k))
    tag = O.mat_identity(n)
    if with_tag:
        for i in range(n):
            for j in range(i + 1, n):
                tag[i][j] = int(rng.integers(0, q))
    a = O.gen_trapdoor(po, a_bar.tolist(), tag, r.tolist())
    want = O.gen_short_basis_for_trapdoor(po, tag, a, r.tolist())
    got = G.gen_short_basis_for_trapdoor(p, np.array(a), r, np.array(tag, dtype=object) if with_tag else None)
    assert got.tolist() == want


def test_classical_short_basis_golden(goldens):
    a, q = O.parse_matz(goldens["short_basis_classical::get_fixed_trapdoor_for_tag_identity"]["literals"][0])
    r, _ = O.parse_matz(goldens["short_basis_classical::get_fixed_trapdoor_for_tag_identity"]["literals"][1])
    sa_l, _ = O.parse_matz(goldens["short_basis_classical::working_sa_l"]["literals"][0])
    sa_r, _ = O.parse_matz(goldens["short_basis_classical::working_sa_r_identity"]["literals"][0])
    p = G.GadgetParameters.init_default(2, 8)
    got = G.gen_short_basis_for_trapdoor(p, np.array(a), np.array(r))
    assert got.tolist() == O.mat_mul(sa_l, sa_r)


@pytest.mark.parametrize("n,q", [(4, 16), (5, 16), (6, 42), (8, 3329)])
def test_ring_short_basis_matches_oracle(n, q, goldens):
    rng = np.random.default_rng(n + q)
    p, po = G.GadgetParametersRing.init_default(n, q), O.GadgetParametersRing.init_default(n, q)
    if (n, q) == (4, 16):  # the reference's fixed fixture, short_basis_ring.rs:358-379
        key = "short_basis_ring::get_fixed_trapdoor"
        pad = lambda v: v + [0] * (n - len(v))
        a = [pad(x) for x in O.parse_matpoly(goldens[key]["literals"][0])[0]]
        r = [pad(x) for x in O.parse_matpoly(goldens[key]["literals"][1])[0]]
        e = [pad(x) for x in O.parse_matpoly(goldens[key]["literals"][2])[0]]
    else:
        a_bar = rng.integers(0, q, n).tolist()
        r = rng.integers(-5, 6, (p.k, n)).tolist()
        e = rng.integers(-5, 6, (p.k, n)).tolist()
        a = O.gen_trapdoor_ring_lwe(po, a_bar, r, e)
    want = O.coeff_embed(O.gen_short_basis_for_trapdoor_ring(po, a, r, e), n)
    got = G.ring_short_basis_embedded(p, np.array(a), np.array(r), np.array(e))
    assert got.tolist() == want


def test_gso_small_matches_oracle():
    """The k x k gadget block's GSO (exact rationals on the host) against the oracle's exact and float64 GSO."""
    for k, base, q in [(6, 2, 64), (7, 2, 100), (5, 3, 200), (24, 2, 2**24 - 3)]:
        blk = G.short_basis_gadget_block(k, base, q)
        got = G.gso_small(blk)
        ex = np.array(O.gso_exact(blk.tolist()), dtype=np.float64)
        assert np.allclose(got, ex, rtol=1e-15, atol=0)
        assert np.allclose(got, O.gso_f64(blk.astype(np.float64)), atol=1e-9)


def test_library_exports_every_declared_symbol():
    from tools_b200 import _ffi

    hdr = open(os.path.join(ROOT, "include", "qfall_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(qf_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert os.path.exists(_ffi.LIB_PATH), "CUDA library not built (python -m tools_b200.build)"
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in qfall_b200.h but not exported"
    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    assert b"sm_100a" in _ffi.lib().qf_version()


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "tools_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), f"{f} mentions the oracle"


def test_rust_shim_matches_header():
    """rust/qfall-tools-b200/src/ffi.rs (the host side in the reference's own language, shipped as source: no Rust
    toolchain in this image) declares exactly the functions of include/qfall_b200.h, with the same parameter counts."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "qfall_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    c_funcs = {}
    for m in re.finditer(r"\b(qf_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        c_funcs[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    rs = open(os.path.join(root, "rust", "qfall-tools-b200", "src", "ffi.rs")).read()
    rs_funcs = {}
    for m in re.finditer(r"pub fn (qf_[a-z0-9_]+)\s*\(([^)]*)\)", rs, flags=re.S):
        args = m.group(2).strip()
        rs_funcs[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    assert len(c_funcs) >= 30
    assert c_funcs == rs_funcs, (sorted(set(c_funcs) ^ set(rs_funcs)), {k: (c_funcs[k], rs_funcs.get(k)) for k in c_funcs if c_funcs[k] != rs_funcs.get(k)})
    # struct layout: same field order as qf_params
    fields_c = re.findall(r"^\s*(?:int32_t|int64_t|uint64_t|double)\s+(\w+);", re.search(r"typedef struct \{(.*?)\} qf_params;", hdr, flags=re.S).group(1), flags=re.M)
    fields_rs = re.findall(r"pub (\w+):", re.search(r"pub struct qf_params \{(.*?)\}", rs, flags=re.S).group(1))
    assert fields_c == fields_rs
