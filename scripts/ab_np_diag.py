#!/usr/bin/env python
"""A/B of the two diagonal-block kernels (np_diag2 against the quad kernel, QF_NP_DIAG_V1=1) on a few shapes: prints the
number of targets whose preimages differ (same key, same seed): python scripts/ab_np_diag.py"""
import os, sys, math
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools_b200 as T
os.environ["QF_OZAKI_MIN_DIM"] = "1024"
for (n, q, s, B) in [(5, 32, 10.0, 130), (8, 127, 70.0, 333), (24, 2**16, 300.0, 700), (64, 2**16, None, 1500)]:
    gp = T.GadgetParameters.init_default(n, q)
    if s is None:
        s = float(math.ceil((math.sqrt(gp.m_bar) + 1.0) * math.sqrt(5.0) * math.log2(n)))
    rng = np.random.default_rng(n)
    u = rng.integers(0, q, (B, n), dtype=np.int64)
    outs, key = [], None
    for v1 in ("0", "1"):
        os.environ["QF_NP_DIAG_V1"] = v1
        psf = T.PSFGPV(gp, s)
        if key is None:
            key = psf.trap_gen(seed=17)
        a, td = key
        psf._a_id = None
        outs.append(psf.samp_p_batch(a, td, u, seed=21))
    d = outs[0] != outs[1]
    rows = d.any(axis=1)
    print(f"n={n} q={q} m={gp.m}: differing targets {rows.sum()} of {B}; first differing targets {np.where(rows)[0][:10]}", flush=True)
