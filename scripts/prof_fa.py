import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tools_b200 as T
from tools_b200 import _ffi
n, q = 256, 2**24
gp = T.GadgetParameters.init_default(n, q)
s = float(math.ceil((math.sqrt(gp.m_bar) + 1.0) * math.sqrt(5.0) * math.log2(n)))
psf = T.PSFGPV(gp, s)
a = np.random.default_rng(1).integers(0, q, (n, gp.m), dtype=np.int64)
psf._install_a(a)
B = 15616
dev = torch.device("cuda:0")
sig = torch.empty((B, gp.m), dtype=torch.int32, device=dev)
u = torch.empty((B, n), dtype=torch.int64, device=dev)
fl = torch.empty(B, dtype=torch.uint8, device=dev)
psf.ctx.call("qf_samp_d_dev", B, 5, 0, _ffi.ptr(sig.data_ptr()))
for _ in range(3):
    psf.ctx.call("qf_f_a_dev", _ffi.ptr(sig.data_ptr()), B, _ffi.ptr(u.data_ptr()), _ffi.ptr(fl.data_ptr()))
psf.ctx.call("qf_synchronize")
