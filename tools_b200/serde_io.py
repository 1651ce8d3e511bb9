"""Import (and export) of the reference's serialised parameters and keys.

The reference derives `Serialize / Deserialize` for its parameter and PSF structs
(gadget_parameters.rs:44-52, 73-81; gpv.rs:53-57; mp_perturbation.rs:57-62; gpv_ring.rs:62-67) and tags the boxed
trapdoor distribution with `typetag` (trapdoor_distribution.rs:21, 35, 61-70, 89, 112): externally tagged,
`{"PlusMinusOneZero": null}` / `{"SampleZ": null}`.  The leaf values are qfall-math types, which serialise as ONE string
field holding the FLINT text form of the value:

    Z, Q, Zq            {"value": "42"} / {"value": "3/4"}
    Modulus             {"modulus": "17"}
    MatZ / MatQ         {"matrix": "[[1, 2],[3, 4]]"}            rows in brackets, entries separated by ", "
    MatZq               {"matrix": "[[1, 2],[3, 4]] mod 7"}
    PolyOverZ           {"poly": "3  1 2 3"}                      "<length>  <coeff_0> <coeff_1> ..." (two spaces)
    MatPolyOverZ        {"matrix": "[[2  1 2, 1  5]]"}
    ModulusPolynomialRingZq   {"poly": "5  1 0 0 0 1 mod 17"}
    PolynomialRingZq / MatPolynomialRingZq   {"poly" | "matrix": ..., "modulus": <ModulusPolynomialRingZq>}

qfall-math's source is not in the build image, so the field NAMES above are restated from the crate's public
behaviour and the parsers are tolerant: a leaf may be the bare string / number, or an object with any single field name.
The string grammars are the ones the reference's own tests use (`MatZ::from_str("[[1, 2],[3, 4]]")`,
`"4  2 8 8 12"`, `"... mod 8"`: gadget_classical.rs:296-345, short_basis_ring.rs:357-444).

Everything here is host-side format conversion into the fixed-width arrays of the C ABI; no arithmetic of the hot path."""
from __future__ import annotations

import json
import re
from fractions import Fraction

import numpy as np

from . import gadget
from .psf import PSFGPV, PSFGPVRing, PSFPerturbation


class SerdeError(ValueError):
    pass


class PsfSpec:
    """What a serialised PSF struct says, without a device context: kind in {"gpv", "perturbation", "ring"}, the gadget
    parameters and the Gaussian parameters as exact Fractions."""

    def __init__(self, kind, gp, s, r=None, s_td=None):
        self.kind, self.gp, self.s, self.r, self.s_td = kind, gp, s, r, s_td

    def instantiate(self, device: int = 0):
        if self.kind == "ring":
            return PSFGPVRing(self.gp, float(self.s), float(self.s_td), device=device)
        if self.kind == "perturbation":
            return PSFPerturbation(self.gp, float(self.r), float(self.s), device=device)
        return PSFGPV(self.gp, float(self.s), device=device)


def _kind_of(psf) -> str:
    if isinstance(psf, PsfSpec):
        return psf.kind
    return "ring" if isinstance(psf, PSFGPVRing) else "perturbation" if isinstance(psf, PSFPerturbation) else "gpv"


# ---------------------------------------------------------------------------------------------------------------------
# FLINT text forms
# ---------------------------------------------------------------------------------------------------------------------
def _leaf(v, *names):
    """serde leaf: bare scalar / string, or {name: string}."""
    if isinstance(v, dict):
        for nm in names:
            if nm in v:
                return v[nm]
        if len(v) == 1:
            return next(iter(v.values()))
        raise SerdeError(f"expected one of the fields {names}, found {sorted(v)}")
    return v


def _split_mod(s: str):
    s = s.strip()
    if " mod " in s:
        body, mod = s.rsplit(" mod ", 1)
        return body.strip(), int(mod.strip())
    return s, None


def parse_z(v) -> int:
    v = _leaf(v, "value", "modulus")
    return int(v) if not isinstance(v, str) else int(v.strip())


def parse_q(v) -> Fraction:
    v = _leaf(v, "value")
    if isinstance(v, (int, float)):
        return Fraction(v)
    return Fraction(v.strip())


def _matrix_rows(body: str):
    body = body.strip()
    if not (body.startswith("[[") and body.endswith("]]")):
        raise SerdeError(f"not a FLINT matrix string: {body[:40]!r}")
    return [row.split(",") for row in body[2:-2].split("],[")]


def parse_mat_z(v) -> np.ndarray:
    """MatZ -> int64 array (object array if an entry needs more than 63 bits)."""
    body, mod = _split_mod(_leaf(v, "matrix"))
    if mod is not None:
        raise SerdeError("MatZ string carries a modulus: use parse_mat_zq")
    rows = [[int(x) for x in row] for row in _matrix_rows(body)]
    a = np.array(rows, dtype=object)
    return a.astype(np.int64) if all(abs(x) < 2**63 for r in rows for x in r) else a


def parse_mat_zq(v):
    """MatZq -> (int64 residues in [0, q), q)."""
    body, mod = _split_mod(_leaf(v, "matrix"))
    if mod is None:
        raise SerdeError("MatZq string without ' mod q'")
    if mod >= 2**62:
        raise SerdeError("modulus must be below 2^62 (FLINT small words)")
    rows = [[int(x) % mod for x in row] for row in _matrix_rows(body)]
    return np.array(rows, dtype=np.int64), mod


def parse_mat_q(v) -> np.ndarray:
    """MatQ -> float64 (entries "a/b" rounded to nearest double: the device works in fp64)."""
    body, _ = _split_mod(_leaf(v, "matrix"))
    return np.array([[float(Fraction(x.strip())) for x in row] for row in _matrix_rows(body)], dtype=np.float64)


def parse_poly_over_z(v):
    """PolyOverZ "len  c0 c1 ..." -> list of coefficients (length `len`; "0" is the zero polynomial)."""
    s, mod = _split_mod(_leaf(v, "poly"))
    parts = s.split()
    if not parts:
        raise SerdeError("empty polynomial string")
    ln = int(parts[0])
    coeffs = [int(x) for x in parts[1:]]
    if len(coeffs) != ln:
        raise SerdeError(f"polynomial string announces {ln} coefficients, has {len(coeffs)}")
    return coeffs if mod is None else [c % mod for c in coeffs]


def parse_modulus_polynomial_ring_zq(v):
    """ModulusPolynomialRingZq "len  c0 .. c_n mod q" -> (coefficients, q); X^n + 1 is [1, 0, .., 0, 1]."""
    s = _leaf(v, "poly", "modulus")
    body, mod = _split_mod(s)
    if mod is None:
        raise SerdeError("ModulusPolynomialRingZq string without ' mod q'")
    return parse_poly_over_z(body), mod


def parse_mat_poly_over_z(v, n: int) -> np.ndarray:
    """MatPolyOverZ -> int64 array rows x cols x n (coefficients padded with zeros)."""
    body, _ = _split_mod(_leaf(v, "matrix"))
    rows = _matrix_rows(body)
    out = np.zeros((len(rows), len(rows[0]), n), dtype=np.int64)
    for i, row in enumerate(rows):
        for j, ent in enumerate(row):
            c = parse_poly_over_z(ent.strip())
            if len(c) > n:
                raise SerdeError("polynomial of degree >= n in a matrix over X^n + 1")
            out[i, j, : len(c)] = c
    return out


def parse_mat_polynomial_ring_zq(v):
    """MatPolynomialRingZq {"matrix": MatPolyOverZ, "modulus": ModulusPolynomialRingZq} -> (rows x cols x n residues, n, q)."""
    if not isinstance(v, dict) or "modulus" not in v:
        raise SerdeError("MatPolynomialRingZq needs the fields matrix and modulus")
    mod_coeffs, q = parse_modulus_polynomial_ring_zq(v["modulus"])
    n = len(mod_coeffs) - 1
    if mod_coeffs != [1] + [0] * (n - 1) + [1]:
        raise SerdeError("only the anticyclic modulus X^n + 1 (common_moduli.rs:41-48) is supported")
    mat = parse_mat_poly_over_z(v.get("matrix", v.get("poly")), n) % q
    return mat, n, q


# writers (the inverse maps: what qfall-math's Display prints)
def fmt_mat(a, q=None) -> str:
    a = np.asarray(a)
    body = "[" + ",".join("[" + ", ".join(str(int(x)) for x in row) + "]" for row in a) + "]"
    return body if q is None else f"{body} mod {int(q)}"


def fmt_mat_q(a) -> str:
    def one(x):
        f = Fraction(float(x))
        return str(f.numerator) if f.denominator == 1 else f"{f.numerator}/{f.denominator}"

    return "[" + ",".join("[" + ", ".join(one(x) for x in row) + "]" for row in np.asarray(a)) + "]"


def fmt_poly(coeffs) -> str:
    c = [int(x) for x in coeffs]
    while c and c[-1] == 0:
        c.pop()
    return "0" if not c else f"{len(c)}  " + " ".join(str(x) for x in c)


def fmt_mat_poly(a) -> str:
    a = np.asarray(a)
    return "[" + ",".join("[" + ", ".join(fmt_poly(p) for p in row) + "]" for row in a) + "]"


# ---------------------------------------------------------------------------------------------------------------------
# parameter structs
# ---------------------------------------------------------------------------------------------------------------------
def _distribution_tag(v, allowed):
    """typetag external tagging: {"PlusMinusOneZero": null}; a bare string is accepted too."""
    tag = v if isinstance(v, str) else (next(iter(v)) if isinstance(v, dict) and len(v) == 1 else None)
    if tag not in allowed:
        raise SerdeError(f"trapdoor distribution {v!r}: this backend implements {sorted(allowed)} "
                         "(trapdoor_distribution.rs:61-70)")
    return tag


def load_gadget_parameters(obj) -> gadget.GadgetParameters:
    """GadgetParameters {n, k, m_bar, base, q, distribution} (gadget_parameters.rs:44-52)."""
    _distribution_tag(obj["distribution"], {"PlusMinusOneZero"})
    return gadget.GadgetParameters(n=parse_z(obj["n"]), k=parse_z(obj["k"]), m_bar=parse_z(obj["m_bar"]),
                                   base=parse_z(obj["base"]), q=parse_z(obj["q"]))


def load_gadget_parameters_ring(obj) -> gadget.GadgetParametersRing:
    """GadgetParametersRing {n, k, m_bar, base, modulus, distribution} (gadget_parameters.rs:73-81)."""
    _distribution_tag(obj["distribution"], {"SampleZ"})
    mod_coeffs, q = parse_modulus_polynomial_ring_zq(obj["modulus"])
    n = parse_z(obj["n"])
    if mod_coeffs != [1] + [0] * (n - 1) + [1]:
        raise SerdeError("only the anticyclic modulus X^n + 1 is supported")
    return gadget.GadgetParametersRing(n=n, k=parse_z(obj["k"]), m_bar=parse_z(obj["m_bar"]), base=parse_z(obj["base"]), q=q)


def dump_gadget_parameters(gp) -> dict:
    if isinstance(gp, gadget.GadgetParametersRing):
        return {"n": {"value": str(gp.n)}, "k": {"value": str(gp.k)}, "m_bar": {"value": str(gp.m_bar)},
                "base": {"value": str(gp.base)},
                "modulus": {"poly": fmt_poly([1] + [0] * (gp.n - 1) + [1]) + f" mod {gp.q}"},
                "distribution": {"SampleZ": None}}
    return {"n": {"value": str(gp.n)}, "k": {"value": str(gp.k)}, "m_bar": {"value": str(gp.m_bar)},
            "base": {"value": str(gp.base)}, "q": {"modulus": str(gp.q)}, "distribution": {"PlusMinusOneZero": None}}


def load_psf(obj, device: int = 0):
    """PSFGPV {gp, s} (gpv.rs:53-57), PSFPerturbation {gp, r, s} (mp_perturbation.rs:57-62) or PSFGPVRing {gp, s, s_td}
    (gpv_ring.rs:62-67) from its serde form (dict or JSON text); the struct is recognised by its fields."""
    return load_psf_spec(obj).instantiate(device)


def load_psf_spec(obj) -> PsfSpec:
    """As load_psf, without creating a device context."""
    if isinstance(obj, (str, bytes)):
        obj = json.loads(obj)
    s = parse_q(obj["s"])
    if "s_td" in obj:
        return PsfSpec("ring", load_gadget_parameters_ring(obj["gp"]), s, s_td=parse_q(obj["s_td"]))
    if "r" in obj:
        return PsfSpec("perturbation", load_gadget_parameters(obj["gp"]), s, r=parse_q(obj["r"]))
    return PsfSpec("gpv", load_gadget_parameters(obj["gp"]), s)


def dump_psf(psf) -> dict:
    def q(x):
        f = Fraction(x)
        return {"value": str(f.numerator) if f.denominator == 1 else f"{f.numerator}/{f.denominator}"}

    out = {"gp": dump_gadget_parameters(psf.gp)}
    if _kind_of(psf) == "perturbation":
        out["r"] = q(psf.r)
    out["s"] = q(psf.s)
    if _kind_of(psf) == "ring":
        out["s_td"] = q(psf.s_td)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# keys and trapdoors (the associated types of the three PSF impls)
# ---------------------------------------------------------------------------------------------------------------------
def load_key(psf, a_obj, td_obj=None):
    """(A, trapdoor) in the array form the `tools_b200` PSF classes take, from the serde form of the reference's values:
      PSFGPV          A: MatZq              trapdoor: [MatZ S, MatQ S~]                          (gpv.rs:60-61)
      PSFPerturbation A: MatZq              trapdoor: [MatZ R, MatQ sqrt(Sigma_2), [MatZ S, MatQ S~]]  (mp_perturbation.rs:194-195)
      PSFGPVRing      A: MatPolynomialRingZq   trapdoor: [MatPolyOverZ r, MatPolyOverZ e]        (gpv_ring.rs:70-71)
    serde writes tuples as JSON arrays.  The modulus of A must be the PSF's."""
    if isinstance(a_obj, (str, bytes)):
        a_obj = json.loads(a_obj)
    if isinstance(td_obj, (str, bytes)):
        td_obj = json.loads(td_obj)
    gp = psf.gp
    kind = _kind_of(psf)
    if kind == "ring":
        mat, n, q = parse_mat_polynomial_ring_zq(a_obj)
        if (n, q) != (gp.n, gp.q) or mat.shape[:2] != (1, gp.k + 2):
            raise SerdeError("ring key does not match the PSF's parameters")
        a = np.ascontiguousarray(mat[0])
        td = None
        if td_obj is not None:
            r = parse_mat_poly_over_z(td_obj[0], gp.n)
            e = parse_mat_poly_over_z(td_obj[1], gp.n)
            if r.shape[:2] != (1, gp.k) or e.shape[:2] != (1, gp.k):
                raise SerdeError("ring trapdoor: r and e are 1 x k vectors of polynomials")
            td = (np.ascontiguousarray(r[0]).astype(np.int32), np.ascontiguousarray(e[0]).astype(np.int32))
        return a, td
    a, q = parse_mat_zq(a_obj)
    if q != gp.q or a.shape != (gp.n, gp.m):
        raise SerdeError("key does not match the PSF's parameters")
    if td_obj is None:
        return a, None
    if kind == "perturbation":
        r = parse_mat_z(td_obj[0])
        l = parse_mat_q(td_obj[1])
        sb, sg = parse_mat_z(td_obj[2][0]), parse_mat_q(td_obj[2][1])
        if r.shape != (gp.m_bar, gp.n * gp.k) or l.shape != (gp.m, gp.m) or sb.shape != sg.shape:
            raise SerdeError("PSFPerturbation trapdoor has the wrong shape")
        return a, (r.astype(np.int8), l, (sb, sg))
    s_basis, s_gso = parse_mat_z(td_obj[0]), parse_mat_q(td_obj[1])
    if s_basis.shape != (gp.m, gp.m) or s_gso.shape != (gp.m, gp.m):
        raise SerdeError("PSFGPV trapdoor has the wrong shape")
    return a, (s_basis, s_gso)


def dump_key(psf, a, td=None):
    """Inverse of load_key: (A, trapdoor) arrays -> the serde form of the reference's values."""
    gp = psf.gp
    kind = _kind_of(psf)
    if kind == "ring":
        a_obj = {"matrix": {"matrix": fmt_mat_poly(np.asarray(a)[None])},
                 "modulus": {"poly": fmt_poly([1] + [0] * (gp.n - 1) + [1]) + f" mod {gp.q}"}}
        td_obj = None if td is None else [{"matrix": fmt_mat_poly(np.asarray(td[0])[None])},
                                          {"matrix": fmt_mat_poly(np.asarray(td[1])[None])}]
        return a_obj, td_obj
    a_obj = {"matrix": fmt_mat(a, gp.q)}
    if td is None:
        return a_obj, None
    if kind == "perturbation":
        r, l, (sb, sg) = td
        return a_obj, [{"matrix": fmt_mat(r)}, {"matrix": fmt_mat_q(l)}, [{"matrix": fmt_mat(sb)}, {"matrix": fmt_mat_q(sg)}]]
    return a_obj, [{"matrix": fmt_mat(td[0])}, {"matrix": fmt_mat_q(td[1])}]


_FLINT_MAT = re.compile(r"^\[\[.*\]\](?: mod \d+)?$", re.S)


def looks_like_flint_matrix(s: str) -> bool:
    return bool(_FLINT_MAT.match(s.strip()))
