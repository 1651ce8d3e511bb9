"""ncu target: ring f_a at C3 (X^256 + 1, q = 3329), 262144 targets, three launches."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tools_b200 as T  # noqa: E402
from tools_b200 import _ffi  # noqa: E402

n, q = 256, 3329
gp = T.GadgetParametersRing.init_default(n, q)
psf = T.PSFGPVRing(gp, ((2 * 2 * 1.005 * math.sqrt(n) + 1) * 2) * 4, 1.005)
a, td = psf.trap_gen(seed=3)
psf._install_a(a)
B = 262144
dev = torch.device("cuda:0")
sig = torch.empty((B, gp.k + 2, n), dtype=torch.int32, device=dev)
u = torch.empty((B, n), dtype=torch.int64, device=dev)
fl = torch.empty(B, dtype=torch.uint8, device=dev)
psf.ctx.call("qf_samp_d_dev", B, 5, 0, _ffi.ptr(sig.data_ptr()))
for _ in range(3):
    psf.ctx.call("qf_f_a_dev", _ffi.ptr(sig.data_ptr()), B, _ffi.ptr(u.data_ptr()), _ffi.ptr(fl.data_ptr()))
psf.ctx.call("qf_synchronize")
