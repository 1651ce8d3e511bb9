#!/usr/bin/env python
"""One warm-up + one timed-size step of the bench workload, for ncu captures:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/prof_step.py
Never report a number measured under the profiler."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tools_b200 as T  # noqa: E402
from tools_b200 import _ffi  # noqa: E402
from bench import WORKLOADS, gpv_s  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 15616
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
n, q = WORKLOADS[wl]["n"], WORKLOADS[wl]["q"]
gp = T.GadgetParameters.init_default(n, q)
psf = T.PSFGPV(gp, gpv_s(gp))
a, td = psf.trap_gen(seed=2)
psf._install_a(a)
psf._install_td(a, td)
dev = torch.device("cuda:0")
u = torch.empty((batch, n), dtype=torch.int64, device=dev)
e = torch.empty((batch, gp.m), dtype=torch.int32, device=dev)
assert _ffi.lib().qf_fill_uniform_modq_dev(_ffi.ptr(u.data_ptr()), u.numel(), q, 7, None) == 0
torch.cuda.synchronize()
print("PROFILE_REGION_BEGIN launches so far:", psf.ctx.launch_count(), flush=True)
for i in range(steps):
    psf.ctx.call("qf_samp_p_dev", _ffi.ptr(u.data_ptr()), batch, 2, i * batch, _ffi.ptr(e.data_ptr()))
    psf.ctx.call("qf_synchronize")
uo = torch.empty_like(u)
fl = torch.empty(batch, dtype=torch.uint8, device=dev)
psf.ctx.call("qf_f_a_dev", _ffi.ptr(e.data_ptr()), batch, _ffi.ptr(uo.data_ptr()), _ffi.ptr(fl.data_ptr()))
psf.ctx.call("qf_synchronize")
assert torch.equal(uo, u) and bool(fl.all())
print("launches total:", psf.ctx.launch_count(), flush=True)
