#!/usr/bin/env python
"""Experiment: C2 samp_p throughput versus the internal chunk size (qf_set_chunk).  usage: exp_chunk.py chunk..."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tools_b200 as T  # noqa: E402
from tools_b200 import _ffi  # noqa: E402
from bench import WORKLOADS, gpv_s  # noqa: E402

n, q, _ = WORKLOADS["c2"]
gp = T.GadgetParameters.init_default(n, q)
psf = T.PSFGPV(gp, gpv_s(gp))
a, td = psf.trap_gen(seed=2)
psf._install_a(a)
psf._install_td(a, td)
dev = torch.device("cuda:0")
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
psf.ctx.call("qf_set_stream", _ffi.ptr(st.cuda_stream))
for chunk in [int(x) for x in sys.argv[1:]] or [15616, 18944, 37888]:
    psf.ctx.call("qf_set_chunk", chunk)
    batch = chunk
    u = torch.empty((batch, n), dtype=torch.int64, device=dev)
    e = torch.empty((batch, gp.m), dtype=torch.int32, device=dev)
    assert _ffi.lib().qf_fill_uniform_modq_dev(_ffi.ptr(u.data_ptr()), u.numel(), q, 7, _ffi.ptr(st.cuda_stream)) == 0
    for i in range(2):
        psf.ctx.call("qf_samp_p_dev", _ffi.ptr(u.data_ptr()), batch, 2, i * batch, _ffi.ptr(e.data_ptr()))
    psf.ctx.call("qf_synchronize")
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for i in range(3):
        psf.ctx.call("qf_samp_p_dev", _ffi.ptr(u.data_ptr()), batch, 2, (2 + i) * batch, _ffi.ptr(e.data_ptr()))
    a1.record()
    psf.ctx.call("qf_synchronize")
    ms = a0.elapsed_time(a1) / 3
    print(json.dumps({"chunk": chunk, "ms": ms, "preimages_per_s": batch / ms * 1e3}), flush=True)
    del u, e
    torch.cuda.empty_cache()
