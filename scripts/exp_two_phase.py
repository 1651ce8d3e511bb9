"""Experiment (CPU, numpy): the two-phase form of GPV08 SampleD over the G-trapdoor short basis
S = [[R S', I + R W],[S', W]]  (columns: first nk = [R;I] S', then m_bar = [I + R W; W]).

Phase 1 (coordinates m-1 .. nk): centre -x with x = [R g; g], g = bits(u): its GSO coordinates vanish there, so
   c'_i = - sum_{j>i} U_ij z_j  (block U_22 only).
Between the phases the residual centre is reduced modulo the sub-lattice L1 = [R;I] S' Z^nk (the law of the output is
invariant under such shifts):  c1 = -[z2 + R g3; g3],  g3 = bits((u - A_bar z2) mod q).
Phase 2 (coordinates nk-1 .. 0): T_1 = Mt_1 c1, c'_i = T_1[i] - sum_{i<j<nk} U_ij z_j.
Output e = [z2 + R e_bot; e_bot], e_bot = g3 + S' z1.

Checks A e = u and E||e||^2, and prints the magnitudes that size the fixed-point digit counts.
Usage: python scripts/exp_two_phase.py n [targets]"""
import sys
import time

import numpy as np

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 8
q, k = 1 << 24, 24
logn = int(np.ceil(np.log2(n)))
mb, nk = n * k + logn * logn, n * k
m = mb + nk
rng = np.random.default_rng(1)
R = (rng.integers(0, 2, (mb, nk)) - rng.integers(0, 2, (mb, nk))).astype(np.float64)
Abar = rng.integers(0, q, (n, mb), dtype=np.int64)
neg = (-Abar) % q
W = np.zeros((nk, mb))
for t in range(k):
    W[t::k, :] = (neg >> t) & 1
Sk = np.zeros((k, k))
for j in range(k):
    Sk[j, j] = 2
for i in range(k - 1):
    Sk[i + 1, i] = -1
Sp = np.kron(np.eye(n), Sk)[:, ::-1]
S = np.zeros((m, m))
S[:mb, :nk] = R @ Sp
S[mb:, :nk] = Sp
S[:mb, nk:] = np.eye(mb) + R @ W
S[mb:, nk:] = W
G = np.kron(np.eye(n, dtype=np.int64), (1 << np.arange(k, dtype=np.int64))[None, :])
A = np.concatenate([Abar, (G - (Abar @ R.astype(np.int64))) % q], axis=1) % q
assert not ((A @ S.astype(np.int64)) % q).any(), "S is not a basis of the q-ary kernel lattice"
t0 = time.time()
Q, Rq = np.linalg.qr(S)
print("qr", round(time.time() - t0, 2), "s", flush=True)
d = np.abs(np.diag(Rq))
U = Rq / np.diag(Rq)[:, None]
Mt = (Q / np.diag(Rq)[None, :]).T  # row i = b~_i / ||b~_i||^2
s = float(np.ceil(d.max() * 5.7))
width = s / d
print(f"m {m}  s {s}  ||b~|| first nk [{d[:nk].min():.3f},{d[:nk].max():.3f}]  last m_bar [{d[nk:].min():.4f},{d[nk:].max():.4f}]")
print(f"widths: first nk [{width[:nk].min():.2f},{width[:nk].max():.2f}]  last m_bar [{width[nk:].min():.0f},{width[nk:].max():.0f}]")
U11, U22, U12 = np.triu(U[:nk, :nk], 1), np.triu(U[nk:, nk:], 1), U[:nk, nk:]
Mt1 = Mt[:nk]
print(f"max|U11| {np.abs(U11).max():.3f}  max|U22| {np.abs(U22).max():.4f}  max|U12| {np.abs(U12).max():.2f}")
print(f"Mt_1 top (nk x m_bar): max {np.abs(Mt1[:, :mb]).max():.2e} rms {np.sqrt((Mt1[:, :mb] ** 2).mean()):.2e};  "
      f"bottom (nk x nk): max {np.abs(Mt1[:, mb:]).max():.2e} rms {np.sqrt((Mt1[:, mb:] ** 2).mean()):.2e}")
rowmax_top = np.abs(Mt1[:, :mb]).max(1)
rowmax_all = np.abs(Mt1).max(1)
print(f"row max of Mt_1: top part [{rowmax_top.min():.2e},{rowmax_top.max():.2e}]  whole row [{rowmax_all.min():.2e},{rowmax_all.max():.2e}]")
print(f"row max of |U22| rows: [{np.abs(U22).max(1)[:-1].min():.2e},{np.abs(U22).max(1).max():.2e}]; U11 rows [{np.abs(U11).max(1)[:-1].min():.2e},{np.abs(U11).max(1).max():.2e}]")


def dgauss(c, w):
    # rounded Gaussian stands in for the exact sampler: sizes and moments are what this experiment looks at
    return np.rint(c + rng.standard_normal(c.shape) * w / np.sqrt(2 * np.pi))


us = rng.integers(0, q, (nt, n), dtype=np.int64)
# phase 1
Z2 = np.zeros((mb, nt))
C2 = np.zeros((mb, nt))
for i in range(mb - 1, -1, -1):
    Z2[i] = dgauss(C2[i], width[nk + i])
    if i:
        C2[:i] -= np.outer(U22[:i, i], Z2[i])
z2 = Z2.T.astype(np.int64)  # nt x mb
h = (us - z2 @ Abar.T) % q
g3 = np.zeros((nt, nk))
for t in range(k):
    g3[:, t::k] = (h >> t) & 1
ytop = Z2.T + g3 @ R.T
y = np.concatenate([ytop, g3], axis=1)  # nt x m
T1 = -(Mt1 @ y.T)  # nk x nt
# reference value of the same centres: the one-pass recursion would have T_1 = t_1(-x) - U_12 z2 up to integer shifts
Z1 = np.zeros((nk, nt))
C1 = T1.copy()
for i in range(nk - 1, -1, -1):
    Z1[i] = dgauss(C1[i], width[i])
    if i:
        C1[:i] -= np.outer(U11[:i, i], Z1[i])
ebot = g3 + Z1.T @ Sp.T
etop = Z2.T + ebot @ R.T
e = np.concatenate([etop, ebot], axis=1)
ei = np.rint(e).astype(np.int64)
print("A e = u:", bool(((ei @ A.T) % q == us).all()))
print("||e||^2 / (m s^2 / 2pi):", np.round((e * e).sum(1) / (m * s * s / (2 * np.pi)), 3))
print(f"max|z2| {np.abs(Z2).max():.0f} (2^{np.log2(np.abs(Z2).max()):.1f})  max|y_top| {np.abs(ytop).max():.0f}  "
      f"max|T_1| {np.abs(T1).max():.1f}  max|z1| {np.abs(Z1).max():.0f}  max|e_bot| {np.abs(ebot).max():.0f}  max|e_top| {np.abs(etop).max():.0f}")
blk = 256
print("phase-2 blocks: width range, max|z1|, max|T_1|")
for b0 in range(0, nk, blk):
    sl = slice(b0, min(nk, b0 + blk))
    print(f"  {b0:6d} [{width[sl].min():7.2f},{width[sl].max():7.2f}]  {np.abs(Z1[sl]).max():6.0f}  {np.abs(T1[sl]).max():8.1f}")
print("phase-1 blocks: width range, max|z2|")
for b0 in range(0, mb, blk):
    sl = slice(b0, min(mb, b0 + blk))
    print(f"  {nk + b0:6d} [{width[nk:][sl].min():9.0f},{width[nk:][sl].max():9.0f}]  {np.abs(Z2[sl]).max():8.0f}")
