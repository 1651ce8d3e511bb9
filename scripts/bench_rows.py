#!/usr/bin/env python
"""Device-resident timings of the SURVEY 8 rows that are not the headline bench (CUDA events on the library's stream,
>= 3 warm-ups, working sets larger than L2): C1 README PSFPerturbation, C3 ring f_a / samp_p, C4 PSFPerturbation (one GPU
shard, TrapGen's A_bar R timed separately), C5 compression and Compress+ByteEncode, classical f_a.  Each row function
returns one dict with its own `roofline` entry; `bench.py` embeds them under `extra.rows`, and
    python scripts/bench_rows.py [rows...]
prints one JSON object per row."""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import torch  # noqa: E402

import tools_b200 as T  # noqa: E402
from tools_b200 import _ffi  # noqa: E402
from tools_b200.compression import byte_code_dev, compress_dev  # noqa: E402


def load_peaks():
    """Measured roofline denominators (driver-written); fallback figures of B200_PROFILING.md otherwise."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        p["source"] = "measured (MEASURED_PEAKS.json)"
        return p
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


PEAKS = load_peaks()
HBM = PEAKS["hbm_gbs"]


def profile_traffic(name):
    """dram read+write bytes per launch of a kernel from the committed ncu summary profiles/traffic_r2.json
    (scripts/ncu_summary.py --json), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic_r2.json")))
        return d.get(name)
    except Exception:
        return None


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


def use_stream(ctx, dev):
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    ctx.call("qf_set_stream", _ffi.ptr(st.cuda_stream))
    return st


def hbm_roofline(kernel, bytes_per_launch, ms, traffic_key=None, note=None):
    gbs = bytes_per_launch / ms / 1e6
    t = profile_traffic(traffic_key) if traffic_key else None
    r = {"kernel": kernel, "bound": "hbm", "achieved": gbs, "peak": HBM, "unit": "GB/s", "frac": gbs / HBM,
         "traffic": (t or {}).get("dram_bytes") if isinstance(t, dict) else t, "peak_source": PEAKS["source"],
         "algorithmic_bytes_per_launch": bytes_per_launch}
    if note:
        r["note"] = note
    return r


# ---------------------------------------------------------------------------------------------------------------------
def row_compress(dev, npoly=16 * 1024 * 1024, ds=(1, 4, 10, 11)):
    """C5: LossyCompressionFIPS203 on 16 Mi polynomials of degree 256 mod 3329; 4 B/coefficient (u16 in + u16 out)."""
    count = npoly * 256
    x = torch.randint(0, 3329, (count,), dtype=torch.int32, device=dev).to(torch.int16)
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    for d in ds:
        ms_c = timeit(lambda: compress_dev(x.data_ptr(), y.data_ptr(), count, 3329, d, st))
        ms_d = timeit(lambda: compress_dev(y.data_ptr(), z.data_ptr(), count, 3329, d, st, decompress=True))
        # round-trip bound of the reference's own test (lossy_compression_fips203.rs:281-326) on a slice
        sl = slice(0, 1 << 20)
        dist = (z[sl].to(torch.int32) - x[sl].to(torch.int32)).abs()
        dist = torch.minimum(dist, 3329 - dist)
        assert int(dist.max().item()) <= 2 ** max(12 - d - 1, 0)
        rows.append({"row": f"C5 lossy_compress / lossy_decompress d={d}", "polys": npoly, "d": d,
                     "compress_ms": ms_c, "decompress_ms": ms_d, "value": count * 4 / ms_c / 1e6, "unit": "GB/s",
                     "decompress_GBps": count * 4 / ms_d / 1e6, "polys_per_s": npoly / ms_c * 1e3,
                     "roofline": hbm_roofline("compress_u16_kernel", count * 4, ms_c, "compress_u16_kernel"),
                     "roofline_decompress": hbm_roofline("compress_u16_kernel<decompress>", count * 4, ms_d)})
    del x, y, z
    return rows


def row_encode(dev, npoly=16 * 1024 * 1024, ds=(1, 4, 10, 11)):
    """C5 in its wire form: ByteEncode_d(Compress_d(.)) / Decompress_d(ByteDecode_d(.)); 512 + 32 d B/polynomial."""
    x = torch.randint(0, 3329, (npoly * 256,), dtype=torch.int32, device=dev).to(torch.int16)
    z = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    for d in ds:
        packed = torch.empty(npoly * 32 * d, dtype=torch.uint8, device=dev)
        ms_e = timeit(lambda: byte_code_dev(x.data_ptr(), packed.data_ptr(), npoly, 3329, d, st))
        ms_d = timeit(lambda: byte_code_dev(packed.data_ptr(), z.data_ptr(), npoly, 3329, d, st, decode=True))
        b = npoly * (512 + 32 * d)
        rows.append({"row": f"C5 Compress+ByteEncode / ByteDecode+Decompress d={d}", "polys": npoly, "d": d, "encode_ms": ms_e,
                     "decode_ms": ms_d, "value": b / ms_e / 1e6, "unit": "GB/s", "decode_GBps": b / ms_d / 1e6,
                     "roofline": hbm_roofline("byte_encode_kernel", b, ms_e, "byte_encode_kernel"),
                     "roofline_decode": hbm_roofline("byte_decode_kernel", b, ms_d)})
        del packed
    del x, z
    return rows


def ring_s(n):
    return ((2 * 2 * 1.005 * math.sqrt(n) + 1) * 2) * 4  # gpv_ring.rs:296-298


def row_ring(dev, B=1 << 20, Bp=131072):
    """C3: PSFGPVRing over X^256 + 1 mod 3329: batched ring f_a (1 Mi targets) and samp_p."""
    n, q = 256, 3329
    gp = T.GadgetParametersRing.init_default(n, q)
    s = ring_s(n)
    psf = T.PSFGPVRing(gp, s, 1.005, device=dev.index or 0)
    a, td = psf.trap_gen(seed=3)
    psf._install_a(a)
    use_stream(psf.ctx, dev)
    sig = torch.empty((B, gp.k + 2, n), dtype=torch.int32, device=dev)
    u = torch.empty((B, n), dtype=torch.int64, device=dev)
    fl = torch.empty(B, dtype=torch.uint8, device=dev)
    psf.ctx.call("qf_samp_d_dev", B, 5, 0, _ffi.ptr(sig.data_ptr()))
    ms = timeit(lambda: psf.ctx.call("qf_f_a_dev", _ffi.ptr(sig.data_ptr()), B, _ffi.ptr(u.data_ptr()), _ffi.ptr(fl.data_ptr())))
    psf.ctx.call("qf_synchronize")
    assert bool(fl.all())
    # bit-exact spot check against the exact negacyclic product (host integers)
    ah = np.asarray(a, dtype=object)
    sg = sig[:2].cpu().numpy().astype(object)
    for b in range(2):
        acc = np.zeros(n, dtype=object)
        for j in range(gp.k + 2):
            full = np.convolve(ah[j], sg[b, j])
            full = np.concatenate([full, np.zeros(2 * n - full.size, dtype=object)])
            acc = acc + full[:n] - full[n:2 * n]
        assert (acc % q).astype(np.int64).tolist() == u[b].cpu().numpy().tolist()
    layout = (gp.k + 2) * n * 4 + n * 8
    rows = [{"row": "C3 ring f_a (X^256+1, q=3329)", "targets": B, "ms": ms, "value": B / ms * 1e3, "unit": "evals/s",
             "roofline": hbm_roofline("f_a_fused_kernel (dense rot^-(a) contraction, tcgen05)", B * 7680, ms, "f_a_fused_ring",
                                      note="7680 B/target algorithmic (SURVEY 8d: 14 x 256 int16 in + 256 u16 out); the ABI "
                                           f"moves int32 sigma and int64 u = {layout} B/target ({B * layout / ms / 1e6:.0f} GB/s)")}]
    del sig, u, fl
    t0 = time.time()
    psf._install_td(a, td)
    setup = time.time() - t0
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
    ops = D * D + 2 * n * D
    rows.append({"row": "C3 ring samp_p", "targets": Bp, "D": D, "ms": ms, "value": Bp / ms * 1e3, "unit": "preimages/s",
                 "trapdoor_setup_s": setup, "checks": {"a_e_equals_u_all": True, "check_domain_all": True,
                 "mean_norm2_over_D_s2_2pi": float((e.double() ** 2).sum((1, 2)).mean().item() / (D * s * s / (2 * math.pi)))},
                 "roofline": {"kernel": "gemm_f64_kernel + np_diag_kernel (D = 3584 nearest plane)", "bound": "tensor",
                              "achieved": Bp * ops / ms / 1e9, "peak": None, "unit": "TFLOP/s (fp64, SURVEY 8d: D^2 + 2 n D per target)",
                              "frac": None, "traffic": None}})
    psf.ctx.close()
    return rows


def pert_s(n, q):
    # s = ceil(1.1 sqrt(5 (s1(R)^2 + 1) + 1)), s1(R) ~ (sqrt(m_bar) + sqrt(nk)) / sqrt(2)
    gp = T.GadgetParameters.init_default(n, q)
    s1 = (math.sqrt(gp.m_bar) + math.sqrt(n * gp.k)) / math.sqrt(2)
    return float(math.ceil(1.15 * math.sqrt(5 * (s1 * s1 + 1) + 1)))


def row_pert(dev, n, q, r, s, B, label, dense=None, i8_peak=None, steps=3):
    """PSFPerturbation::samp_p (mp_perturbation.rs:304-336) device resident + TrapGen timed separately."""
    gp = T.GadgetParameters.init_default(n, q)
    psf = T.PSFPerturbation(gp, r, s, device=dev.index or 0)
    psf.trap_gen(seed=3, dense_sqrt_sigma_2=False)  # untimed: first launches load the kernels' modules
    psf.ctx.call("qf_profile", 1)
    t0 = time.time()
    a, td = psf.trap_gen(seed=4, dense_sqrt_sigma_2=dense)
    t_trapgen = time.time() - t0
    # the A_bar R contraction inside TrapGen (gadget_classical.rs:66): per-launch CUDA-event time of the tcgen05 kernel
    gms, gfl, gln = _ffi.C.c_double(), _ffi.C.c_double(), _ffi.C.c_uint64()
    ims, iops, iiss, iln = _ffi.C.c_double(), _ffi.C.c_double(), _ffi.C.c_double(), _ffi.C.c_uint64()
    psf.ctx.call("qf_profile_read", _ffi.C.byref(gms), _ffi.C.byref(gfl), _ffi.C.byref(gln), _ffi.C.byref(ims),
                 _ffi.C.byref(iops), _ffi.C.byref(iiss), _ffi.C.byref(iln))
    abar_r = {"ms": ims.value, "ops": iops.value, "TOPs": iops.value / ims.value / 1e9 if ims.value else None,
              "launches": int(iln.value)}
    t0 = time.time()
    psf._install_a(a)
    psf._install_td(a, td)
    setup = time.time() - t0
    psf.ctx.call("qf_profile_read", None, None, None, None, None, None, None)
    use_stream(psf.ctx, dev)
    u = torch.empty((B, n), dtype=torch.int64, device=dev)
    assert _ffi.lib().qf_fill_uniform_modq_dev(_ffi.ptr(u.data_ptr()), u.numel(), q, 9, None) == 0
    torch.cuda.synchronize()
    e = torch.empty((B, gp.m), dtype=torch.int32, device=dev)
    for _ in range(3):
        psf.ctx.call("qf_samp_p_dev", _ffi.ptr(u.data_ptr()), B, 7, 0, _ffi.ptr(e.data_ptr()))
    psf.ctx.call("qf_synchronize")
    psf.ctx.call("qf_profile_read", None, None, None, None, None, None, None)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        psf.ctx.call("qf_samp_p_dev", _ffi.ptr(u.data_ptr()), B, 7, (i + 1) * B, _ffi.ptr(e.data_ptr()))
    ev1.record()
    psf.ctx.call("qf_synchronize")
    ms = ev0.elapsed_time(ev1) / steps
    psf.ctx.call("qf_profile_read", _ffi.C.byref(gms), _ffi.C.byref(gfl), _ffi.C.byref(gln), _ffi.C.byref(ims),
                 _ffi.C.byref(iops), _ffi.C.byref(iiss), _ffi.C.byref(iln))
    psf.ctx.call("qf_profile", 0)
    uo = torch.empty_like(u)
    fl = torch.empty(B, dtype=torch.uint8, device=dev)
    psf.ctx.call("qf_f_a_dev", _ffi.ptr(e.data_ptr()), B, _ffi.ptr(uo.data_ptr()), _ffi.ptr(fl.data_ptr()))
    psf.ctx.call("qf_synchronize")
    assert torch.equal(uo, u) and bool(fl.all())
    fa_ms = timeit(lambda: psf.ctx.call("qf_f_a_dev", _ffi.ptr(e.data_ptr()), B, _ffi.ptr(uo.data_ptr()), _ffi.ptr(fl.data_ptr())))
    m = gp.m
    ops_target = m * (m + 1) + 2 * n * m + 2 * gp.m_bar * n * gp.k  # SURVEY 8d per-target figure
    roof = {"kernel": "gemm_i8_kernel (tcgen05 kind::i8): x2 = sqrt(Sigma_2) g (fixed point), v = u - A p, e = p + [R;I] z",
            "bound": "tensor", "unit": "TOP/s", "traffic": None,
            "achieved": iops.value / ims.value / 1e9 if ims.value else None, "peak": i8_peak,
            "frac": (iops.value / ims.value / 1e9 / i8_peak) if (ims.value and i8_peak) else None,
            "achieved_issued": iiss.value / ims.value / 1e9 if ims.value else None,
            "kernel_share_of_step": ims.value / (ms * steps) if ms else None, "launches": int(iln.value),
            "survey_ops_per_target": ops_target, "survey_TOPs": B * ops_target / ms / 1e9,
            "peak_source": "int8 probe measured in this run (qf_probe_i8_peak, sustained)" if i8_peak else None}
    row = {"row": label, "n": n, "q": q, "m": m, "r": r, "s": s, "targets_per_step": B, "ms": ms,
           "value": B / ms * 1e3, "unit": "preimages/s",
           "sqrt_sigma_2": "dense m x m (reference form)" if td[1] is not None else "block-structured (backend)",
           "trap_gen_s": t_trapgen, "trap_gen_abar_r": abar_r, "trapdoor_install_s": setup,
           "f_a": {"value": B / fa_ms * 1e3, "unit": "evals/s", "ms": fa_ms},
           "checks": {"a_e_equals_u_all": True, "check_domain_all": True,
                      "mean_norm2_over_m_s2r2_2pi": float((e.double() ** 2).sum(1).mean().item() / (m * (s * r) ** 2 / (2 * math.pi)))},
           "roofline": roof}
    psf.ctx.close()
    return row


def row_f_a(dev, B=262144):
    """C2 classical f_a on samp_d vectors (device resident)."""
    n, q = 256, 2**24
    gp = T.GadgetParameters.init_default(n, q)
    s = float(math.ceil((math.sqrt(gp.m_bar) + 1.0) * math.sqrt(5.0) * math.log2(n)))
    psf = T.PSFGPV(gp, s, device=dev.index or 0)
    rng = np.random.default_rng(1)
    a = rng.integers(0, q, (n, gp.m), dtype=np.int64)
    psf._install_a(a)
    use_stream(psf.ctx, dev)
    sig = torch.empty((B, gp.m), dtype=torch.int32, device=dev)
    u = torch.empty((B, n), dtype=torch.int64, device=dev)
    fl = torch.empty(B, dtype=torch.uint8, device=dev)
    ms_d = timeit(lambda: psf.ctx.call("qf_samp_d_dev", B, 5, 0, _ffi.ptr(sig.data_ptr())))
    ms = timeit(lambda: psf.ctx.call("qf_f_a_dev", _ffi.ptr(sig.data_ptr()), B, _ffi.ptr(u.data_ptr()), _ffi.ptr(fl.data_ptr())))
    psf.ctx.call("qf_synchronize")
    assert bool(fl.all())
    chk = (sig[:8].cpu().numpy().astype(object) @ a.T.astype(object)) % q
    assert np.array_equal(chk.astype(np.int64), u[:8].cpu().numpy())
    row = {"row": "C2 classical f_a + check_domain", "targets": B, "ms": ms, "value": B / ms * 1e3, "unit": "evals/s",
           "samp_d_draws_per_s": B * gp.m / ms_d * 1e3,
           "roofline": hbm_roofline("f_a_fused_kernel (digit split fused into the tcgen05 contraction)", B * (gp.m * 4 + n * 8), ms,
                                    "f_a_fused_kernel", note="sigma int32 in + u int64 out; 2 n m = %d useful int ops/target = %.0f TOP/s"
                                    % (2 * n * gp.m, B * 2 * n * gp.m / ms / 1e9))}
    psf.ctx.close()
    return row


def c4_params():
    n, q = 512, 2**32 - 5
    return n, q, 9.0, pert_s(n, q)


ROWS = {
    "compress": lambda dev: row_compress(dev),
    "encode": lambda dev: row_encode(dev),
    "ring": lambda dev: row_ring(dev),
    "f_a": lambda dev: [row_f_a(dev)],
    "pert_c1": lambda dev: [row_pert(dev, 8, 64, 3.0, 25.0, 262144, "C1 README PSFPerturbation n=8 q=64 r=3 s=25")],
    "pert_256": lambda dev: [row_pert(dev, 256, 2**24, 8.0, pert_s(256, 2**24), 15616, "PSFPerturbation n=256 q=2^24 r=8")],
    "pert_c4": lambda dev: [row_pert(dev, *c4_params(), 15104, "C4 PSFPerturbation n=512 q=2^32-5 k=32 r=9 (one GPU shard)")],
    "pert_c4_dense": lambda dev: [row_pert(dev, *c4_params(), 7552, "C4 PSFPerturbation, dense sqrt(Sigma_2) (reference form)", True)],
}

if __name__ == "__main__":
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    for name in (sys.argv[1:] or list(ROWS)):
        try:
            for r_ in ROWS[name](device):
                print(json.dumps(r_), flush=True)
        except Exception as ex:  # keep going: one row failing must not hide the others
            print(json.dumps({"row": name, "error": repr(ex)}), flush=True)
