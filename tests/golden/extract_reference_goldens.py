#!/usr/bin/env python
"""Extract the inline golden vectors of the reference's own unit tests.

Run in the build container only (reads /root/reference, which does not exist on
the GPU box):

    python tests/golden/extract_reference_goldens.py

It pulls the string literals (FLINT matrix / polynomial strings) out of the
named ``#[test]`` functions and helper fns and writes them verbatim to
``tests/golden/reference_goldens.json`` together with the file:line they came
from.  No reference code is copied -- only the literal test data.
"""
import json
import os
import re
import sys

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")

# (file, function name) -> key
TARGETS = [
    ("sample/g_trapdoor/gadget_classical.rs", "correctness_base_2"),
    ("sample/g_trapdoor/gadget_classical.rs", "correctness_base_5"),
    ("sample/g_trapdoor/gadget_classical.rs", "correctness_base_2_3x3"),
    ("sample/g_trapdoor/gadget_classical.rs", "correctness_base_3_2x5"),
    ("sample/g_trapdoor/gadget_classical.rs", "returns_correct_solution_mat"),
    ("sample/g_trapdoor/gadget_classical.rs", "base_2_power_two"),
    ("sample/g_trapdoor/gadget_classical.rs", "base_2_arbitrary"),
    ("sample/g_trapdoor/gadget_classical.rs", "base_5_power_5"),
    ("sample/g_trapdoor/gadget_classical.rs", "base_5_arbitrary"),
    ("sample/g_trapdoor/short_basis_classical.rs", "get_fixed_trapdoor_for_tag_identity"),
    ("sample/g_trapdoor/short_basis_classical.rs", "working_sa_l"),
    ("sample/g_trapdoor/short_basis_classical.rs", "working_sa_r_identity"),
    ("sample/g_trapdoor/short_basis_classical.rs", "working_example_tag_identity"),
    ("sample/g_trapdoor/short_basis_ring.rs", "get_fixed_trapdoor"),
    ("sample/g_trapdoor/short_basis_ring.rs", "working_sa_l"),
    ("sample/g_trapdoor/short_basis_ring.rs", "working_sa_r"),
    ("sample/g_trapdoor/short_basis_ring.rs", "base_2_power_two"),
    ("sample/g_trapdoor/short_basis_ring.rs", "base_2_arbitrary"),
    ("sample/g_trapdoor/short_basis_ring.rs", "base_5_power_5"),
    ("sample/g_trapdoor/short_basis_ring.rs", "base_5_arbitrary"),
    ("sample/g_trapdoor/gadget_ring.rs", "is_correct_solution"),
    ("utils/rotation_matrix.rs", "correct_rotation_matrix_vec"),
    ("utils/rotation_matrix.rs", "correct_rotation_matrix_mat"),
]


def fn_body(src: str, name: str):
    m = re.search(r"fn\s+" + re.escape(name) + r"\s*\([^)]*\)[^{]*\{", src)
    if not m:
        raise KeyError(name)
    start = m.end()
    depth, i = 1, start
    in_str = False
    while depth and i < len(src):
        ch = src[i]
        if in_str:
            if ch == "\\":
                i += 1
            elif ch == '"':
                in_str = False
        else:
            if ch == '"':
                in_str = True
            elif ch == "{":
                depth += 1
            elif ch == "}":
                depth -= 1
        i += 1
    line = src[: m.start()].count("\n") + 1
    return src[start : i - 1], line


def literals(body: str):
    out = []
    for m in re.finditer(r'"((?:[^"\\]|\\.)*)"', body, re.S):
        s = m.group(1)
        s = re.sub(r"\\\n\s*", "", s)  # Rust line continuation
        out.append(s)
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; goldens are already committed")
    res = {}
    for rel, name in TARGETS:
        src = open(os.path.join(REF, rel)).read()
        body, line = fn_body(src, name)
        key = f"{os.path.basename(rel)[:-3]}::{name}"
        res[key] = {"source": f"src/{rel}:{line}", "literals": literals(body)}
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1)
    print(f"wrote {OUT}: {len(res)} entries")


if __name__ == "__main__":
    main()
