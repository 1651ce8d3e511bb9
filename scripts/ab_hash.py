#!/usr/bin/env python
"""Print a hash of samp_p outputs for a few shapes under the current environment: run twice with different test
switches (QF_I8_EPI_STAGE=0, QF_NP_DIAG_V1=1, ...) and compare the lines -- switches that only change HOW a step is
computed must not change a single preimage."""
import hashlib
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools_b200 as T  # noqa: E402

os.environ.setdefault("QF_OZAKI_MIN_DIM", "1024")
for (n, q, B) in [(64, 2**16, 1500), (128, 2**12, 1100), (64, 2**24, 700)]:
    gp = T.GadgetParameters.init_default(n, q)
    s = float(math.ceil((math.sqrt(gp.m_bar) + 1.0) * math.sqrt(5.0) * math.log2(n)))
    rng = np.random.default_rng(n)
    u = rng.integers(0, q, (B, n), dtype=np.int64)
    psf = T.PSFGPV(gp, s)
    a, td = psf.trap_gen(seed=17)
    e = psf.samp_p_batch(a, td, u, seed=21)
    ok = np.array_equal((e.astype(np.int64) @ a.T.astype(np.int64)) % q, u)
    print(f"n={n} q={q} m={gp.m} B={B} A_e_eq_u={ok} sha={hashlib.sha256(e.tobytes()).hexdigest()[:16]}", flush=True)
