#!/usr/bin/env python
"""Device-resident timings of every SURVEY 8 row that is not the headline bench (CUDA events, 3 warm-ups):
compression C5, ring f_a / samp_p C3, classical f_a C2, PSFPerturbation samp_p (C1 and n=256), TrapGen.
Writes one JSON object per row to stdout.  usage: python scripts/bench_rows.py [rows...]"""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tools_b200 as T  # noqa: E402
from tools_b200 import _ffi  # noqa: E402
from tools_b200.compression import byte_code_dev, compress_dev  # noqa: E402

dev = torch.device("cuda:0")
PEAKS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM = PEAKS["hbm_gbs"]


def timeit(fn, iters=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def use_stream(ctx):
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    ctx.call("qf_set_stream", _ffi.ptr(st.cuda_stream))
    return st


def row_compress():
    npoly = 16 * 1024 * 1024
    count = npoly * 256
    x = torch.randint(0, 3329, (count,), dtype=torch.int32, device=dev).to(torch.int16)
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    for d in (1, 4, 10, 11):
        ms_c = timeit(lambda: compress_dev(x.data_ptr(), y.data_ptr(), count, 3329, d, st))
        ms_d = timeit(lambda: compress_dev(y.data_ptr(), z.data_ptr(), count, 3329, d, st, decompress=True))
        gb = count * 4 / 1e9
        print(json.dumps({"row": "C5 compression", "d": d, "polys": npoly, "compress_ms": ms_c, "decompress_ms": ms_d,
                          "compress_GBps": gb / ms_c * 1e3, "decompress_GBps": gb / ms_d * 1e3,
                          "frac_of_measured_hbm": gb / ms_c * 1e3 / HBM, "algorithmic_bytes": count * 4,
                          "polys_per_s": npoly / ms_c * 1e3}), flush=True)


def row_encode():
    """ByteEncode_d(Compress_d(.)) / Decompress_d(ByteDecode_d(.)) on the C5 stream: 512 + 32 d algorithmic bytes / polynomial."""
    npoly = 16 * 1024 * 1024
    x = torch.randint(0, 3329, (npoly * 256,), dtype=torch.int32, device=dev).to(torch.int16)
    z = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    for d in (1, 4, 10, 11):
        packed = torch.empty(npoly * 32 * d, dtype=torch.uint8, device=dev)
        ms_e = timeit(lambda: byte_code_dev(x.data_ptr(), packed.data_ptr(), npoly, 3329, d, st))
        ms_d = timeit(lambda: byte_code_dev(packed.data_ptr(), z.data_ptr(), npoly, 3329, d, st, decode=True))
        gb = npoly * (512 + 32 * d) / 1e9
        print(json.dumps({"row": "C5 compress+ByteEncode", "d": d, "polys": npoly, "encode_ms": ms_e, "decode_ms": ms_d,
                          "encode_GBps": gb / ms_e * 1e3, "decode_GBps": gb / ms_d * 1e3,
                          "frac_of_measured_hbm": gb / ms_e * 1e3 / HBM, "frac_of_measured_hbm_decode": gb / ms_d * 1e3 / HBM,
                          "algorithmic_bytes": npoly * (512 + 32 * d), "polys_per_s": npoly / ms_e * 1e3}), flush=True)
        del packed


def row_ring():
    n, q = 256, 3329
    gp = T.GadgetParametersRing.init_default(n, q)
    s = ((2 * 2 * 1.005 * math.sqrt(n) + 1) * 2) * 4
    psf = T.PSFGPVRing(gp, s, 1.005)
    a, td = psf.trap_gen(seed=3)
    psf._install_a(a)
    use_stream(psf.ctx)
    B = 262144
    sig = torch.empty((B, gp.k + 2, n), dtype=torch.int32, device=dev)
    u = torch.empty((B, n), dtype=torch.int64, device=dev)
    fl = torch.empty(B, dtype=torch.uint8, device=dev)
    psf.ctx.call("qf_samp_d_dev", B, 5, 0, _ffi.ptr(sig.data_ptr()))
    ms_d = timeit(lambda: psf.ctx.call("qf_samp_d_dev", B, 5, 0, _ffi.ptr(sig.data_ptr())))
    ms = timeit(lambda: psf.ctx.call("qf_f_a_dev", _ffi.ptr(sig.data_ptr()), B, _ffi.ptr(u.data_ptr()), _ffi.ptr(fl.data_ptr())))
    psf.ctx.call("qf_synchronize")
    assert bool(fl.all())
    actual = (gp.k + 2) * n * 4 + n * 8
    print(json.dumps({"row": "C3 ring f_a", "n": n, "q": q, "B": B, "ms": ms, "evals_per_s": B / ms * 1e3,
                      "algorithmic_bytes_per_target": 7680, "layout_bytes_per_target": actual,
                      "GBps_algorithmic": B * 7680 / ms / 1e6, "GBps_layout": B * actual / ms / 1e6,
                      "frac_of_measured_hbm_layout": B * actual / ms / 1e6 / HBM,
                      "samp_d_ms": ms_d, "samp_d_draws_per_s": B * (gp.k + 2) * n / ms_d * 1e3}), flush=True)
    t0 = time.time()
    psf._install_td(a, td)
    setup = time.time() - t0
    Bp = 32768
    uu = torch.empty((Bp, n), dtype=torch.int64, device=dev)
    assert _ffi.lib().qf_fill_uniform_modq_dev(_ffi.ptr(uu.data_ptr()), uu.numel(), q, 9, None) == 0
    torch.cuda.synchronize()
    e = torch.empty((Bp, gp.k + 2, n), dtype=torch.int32, device=dev)
    ms = timeit(lambda: psf.ctx.call("qf_samp_p_dev", _ffi.ptr(uu.data_ptr()), Bp, 7, 0, _ffi.ptr(e.data_ptr())), iters=3)
    uo = torch.empty_like(uu)
    flp = torch.empty(Bp, dtype=torch.uint8, device=dev)
    psf.ctx.call("qf_f_a_dev", _ffi.ptr(e.data_ptr()), Bp, _ffi.ptr(uo.data_ptr()), _ffi.ptr(flp.data_ptr()))
    psf.ctx.call("qf_synchronize")
    assert torch.equal(uo, uu) and bool(flp.all())
    D = n * (gp.k + 2)
    print(json.dumps({"row": "C3 ring samp_p", "D": D, "B": Bp, "ms": ms, "preimages_per_s": Bp / ms * 1e3,
                      "trapdoor_setup_s": setup, "flop_per_target": D * D + 2 * n * D,
                      "norm2_ratio": float((e.double() ** 2).sum((1, 2)).mean().item() / (D * s * s / (2 * math.pi)))}),
          flush=True)


def row_f_a():
    n, q = 256, 2**24
    gp = T.GadgetParameters.init_default(n, q)
    s = float(math.ceil((math.sqrt(gp.m_bar) + 1.0) * math.sqrt(5.0) * math.log2(n)))
    psf = T.PSFGPV(gp, s)
    rng = np.random.default_rng(1)
    a = rng.integers(0, q, (n, gp.m), dtype=np.int64)
    psf._install_a(a)
    use_stream(psf.ctx)
    B = 65536
    sig = torch.empty((B, gp.m), dtype=torch.int32, device=dev)
    u = torch.empty((B, n), dtype=torch.int64, device=dev)
    fl = torch.empty(B, dtype=torch.uint8, device=dev)
    ms_d = timeit(lambda: psf.ctx.call("qf_samp_d_dev", B, 5, 0, _ffi.ptr(sig.data_ptr())))
    ms = timeit(lambda: psf.ctx.call("qf_f_a_dev", _ffi.ptr(sig.data_ptr()), B, _ffi.ptr(u.data_ptr()), _ffi.ptr(fl.data_ptr())))
    psf.ctx.call("qf_synchronize")
    assert bool(fl.all())
    chk = (sig[:8].cpu().numpy().astype(object) @ a.T.astype(object)) % q
    assert np.array_equal(chk.astype(np.int64), u[:8].cpu().numpy())
    print(json.dumps({"row": "C2 f_a classical", "n": n, "m": gp.m, "B": B, "ms": ms, "evals_per_s": B / ms * 1e3,
                      "useful_int_ops_per_target": 2 * n * gp.m, "TOPs_useful": B * 2 * n * gp.m / ms / 1e9,
                      "sigma_bytes_per_target": gp.m * 4, "GBps_sigma": B * gp.m * 4 / ms / 1e6,
                      "samp_d_ms": ms_d, "samp_d_draws_per_s": B * gp.m / ms_d * 1e3}), flush=True)


def row_pert(n, q, r, s, B, label, dense=None):
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFPerturbation(gp, r, s)
    t0 = time.time()
    a, td = psf.trap_gen(seed=4, dense_sqrt_sigma_2=dense)
    psf._install_a(a)
    psf._install_td(a, td)
    setup = time.time() - t0
    use_stream(psf.ctx)
    u = torch.empty((B, n), dtype=torch.int64, device=dev)
    assert _ffi.lib().qf_fill_uniform_modq_dev(_ffi.ptr(u.data_ptr()), u.numel(), q, 9, None) == 0
    torch.cuda.synchronize()
    e = torch.empty((B, gp.m), dtype=torch.int32, device=dev)
    ms = timeit(lambda: psf.ctx.call("qf_samp_p_dev", _ffi.ptr(u.data_ptr()), B, 7, 0, _ffi.ptr(e.data_ptr())), iters=3)
    uo = torch.empty_like(u)
    fl = torch.empty(B, dtype=torch.uint8, device=dev)
    psf.ctx.call("qf_f_a_dev", _ffi.ptr(e.data_ptr()), B, _ffi.ptr(uo.data_ptr()), _ffi.ptr(fl.data_ptr()))
    psf.ctx.call("qf_synchronize")
    assert torch.equal(uo, u) and bool(fl.all())
    m = gp.m
    print(json.dumps({"row": label, "n": n, "q": q, "m": m, "r": r, "s": s, "B": B, "ms": ms,
                      "sqrt_sigma_2": "dense m x m (reference form)" if td[1] is not None else "block-structured (backend)",
                      "preimages_per_s": B / ms * 1e3, "key_setup_s": setup,
                      "ops_per_target": m * (m + 1) + 2 * n * m + 2 * gp.m_bar * n * gp.k,
                      "norm2_ratio": float((e.double() ** 2).sum(1).mean().item() / (m * (s * r) ** 2 / (2 * math.pi)))}),
          flush=True)


def pert_s(n, q):
    # s = ceil(1.1 sqrt(5 (s1(R)^2 + 1) + 1)), s1(R) ~ (sqrt(m_bar) + sqrt(nk)) / sqrt(2)
    gp = T.GadgetParameters.init_default(n, q)
    s1 = (math.sqrt(gp.m_bar) + math.sqrt(n * gp.k)) / math.sqrt(2)
    return float(math.ceil(1.15 * math.sqrt(5 * (s1 * s1 + 1) + 1)))


ROWS = {
    "compress": row_compress,
    "encode": row_encode,
    "ring": row_ring,
    "f_a": row_f_a,
    "pert_c1": lambda: row_pert(8, 64, 3.0, 25.0, 262144, "C1 PSFPerturbation n=8 q=64 r=3 s=25"),
    "pert_256": lambda: row_pert(256, 2**24, 8.0, pert_s(256, 2**24), 15616, "PSFPerturbation n=256 q=2^24 r=8"),
    "pert_256_dense": lambda: row_pert(256, 2**24, 8.0, pert_s(256, 2**24), 15616, "PSFPerturbation n=256 q=2^24 r=8", True),
    "pert_c4_dense": lambda: row_pert(512, 2**32 - 5, 9.0, pert_s(512, 2**32 - 5), 11264,
                                      "C4 PSFPerturbation n=512 q=2^32-5 r=9 (one GPU shard)", True),
    # C4 per-GPU shard: n = 512, q = 2^32 - 5, k = 32, m = 32849 (the 4 Mi targets are split over 8 GPUs)
    "pert_c4": lambda: row_pert(512, 2**32 - 5, 9.0, pert_s(512, 2**32 - 5), 11264, "C4 PSFPerturbation n=512 q=2^32-5 r=9 (one GPU shard)"),
}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(ROWS)):
        try:
            ROWS[name]()
        except Exception as ex:  # keep going: one row failing must not hide the others
            print(json.dumps({"row": name, "error": repr(ex)}), flush=True)
