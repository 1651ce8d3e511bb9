"""Experiment (CPU, numpy): size of the nearest-plane coefficients z and of the centres T when the particular solution
of A x = u is the short gadget preimage x = [R g; g] (g = base-b digits of u) instead of the pivot-column solution with
entries in [0, q).  Prints, per 256-block of coordinates: the width s/||b~_i|| range, max |U| of the rows, max |z| and
max |T| seen over a few targets.  Usage: python scripts/exp_short_centre.py n [targets]"""
import sys
import time

import numpy as np

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 4
q, k = 1 << 24, 24
logn = int(np.ceil(np.log2(n)))
mb, nk = n * k + logn * logn, n * k
m = mb + nk
s = 1428.0 if n == 256 else float(np.ceil(np.sqrt(m) * np.log2(m) * 1.0))  # only the relative sizes matter
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
Sp = np.kron(np.eye(n), Sk)[:, ::-1]  # column-reversed (b^k = q)
S = np.zeros((m, m))
S[:mb, :nk] = R @ Sp
S[mb:, :nk] = Sp
S[:mb, nk:] = np.eye(mb) + R @ W
S[mb:, nk:] = W
t0 = time.time()
Rq = np.linalg.qr(S, mode="r")
print("qr", time.time() - t0, "s", flush=True)
d = np.abs(np.diag(Rq))
U = Rq / np.diag(Rq)[:, None]  # unit upper triangular: U_ij = <b_j, b~_i>/||b~_i||^2
width = s / d
print("m", m, "s", s, "||b~|| min/max", d.min(), d.max(), "width min/max", width.min(), width.max())
# centres: short solution x = [R g; g]: c = -x, t = U (S^-1 c)
us = rng.integers(0, q, (nt, n), dtype=np.int64)
g = np.zeros((nt, nk))
for t in range(k):
    g[:, t::k] = (us >> t) & 1
x = np.concatenate([g @ R.T, g], axis=1)  # nt x m
# GSO coordinates of c = -x:  t = D^-1 B~^t c = D^-1 Q^t c * ... use triangular solve: S y = c -> t = U y
y = np.linalg.solve(S, -x.T)  # m x nt
T = U @ y
print("short centre: max |T| first nk", np.abs(T[:nk]).max(), " last m_bar", np.abs(T[nk:]).max(), " max |y|", np.abs(y).max())
# randomized nearest plane (continuous-Gaussian rounding is enough for sizes)
Z = np.zeros((m, nt))
C = T.copy()
for i in range(m - 1, -1, -1):
    ci = C[i]
    zi = np.rint(ci + rng.standard_normal(nt) * width[i] / np.sqrt(2 * np.pi))
    Z[i] = zi
    if i:
        C[:i] -= np.outer(U[:i, i], zi)
e = x.T + S @ Z
print("||e||^2/(m s^2/2pi)", (e * e).sum(0) / (m * s * s / (2 * np.pi)))
blk = 256
print("block  width[min,max]  max|U row|  max|z|  max|C| (centres incl. updates)")
for b0 in range(0, m, blk):
    sl = slice(b0, min(m, b0 + blk))
    print(f"{b0:6d}  [{width[sl].min():8.2f},{width[sl].max():8.2f}]  {np.abs(np.triu(U, 1)[sl]).max():8.2f}  "
          f"{np.abs(Z[sl]).max():8.0f}  {np.abs(T[sl]).max():10.2f}")
zabs = np.abs(Z)
print("fraction of coordinates with |z| > 127:", (zabs > 127).mean(), " > 32767:", (zabs > 32767).mean())
np.save(f"/tmp/exp_width_{n}.npy", width)
