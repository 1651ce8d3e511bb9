#!/usr/bin/env python
"""C2-size check of a fixed-point digit configuration of the two-phase recursion: same key, same seed, `cfg` (QF_NP2_CFG
syntax, or 'default') against every matrix at full width; prints the fraction of bit-identical preimages.
Usage: python scripts/ab_c2_precision.py cfg [cfg ...]"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools_b200 as T  # noqa: E402
from bench import WORKLOADS, gpv_s  # noqa: E402

n, q = WORKLOADS["c2"]["n"], WORKLOADS["c2"]["q"]
gp = T.GadgetParameters.init_default(n, q)
B = 4096
u = np.random.default_rng(1).integers(0, q, (B, n), dtype=np.int64)
key, outs = None, {}
for cfg in ["7,0,7,0,7,0,7"] + sys.argv[1:]:
    if cfg == "default":
        os.environ.pop("QF_NP2_CFG", None)
    else:
        os.environ["QF_NP2_CFG"] = cfg
    psf = T.PSFGPV(gp, gpv_s(gp))
    if key is None:
        key = psf.trap_gen(seed=2)
    a, td = key
    psf._a_id = None
    outs[cfg] = psf.samp_p_batch(a, td, u, seed=5)
    same = (outs[cfg] == outs["7,0,7,0,7,0,7"]).all(axis=1).mean()
    print(f"cfg {cfg}: identical preimages vs full width: {same:.4f}", flush=True)
    del psf
