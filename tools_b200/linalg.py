"""Per-key floating-point setup (not per target): GSO of a short basis
(MatQ::gso, gpv.rs:91) and sqrt(Sigma_2) (compute_sqrt_sigma_2, mp_perturbation.rs:111-139) in
float64.  Uses torch on the GPU when one is present (setup plumbing), numpy otherwise."""
import math

import numpy as np


def _torch_cuda():
    try:
        import torch

        if torch.cuda.is_available():
            return torch
    except Exception:
        pass
    return None


def gso(basis: np.ndarray) -> np.ndarray:
    """Unnormalised Gram-Schmidt of the COLUMNS: b~_i = R_ii * Q[:, i] from a Householder QR."""
    b = np.asarray(basis, dtype=np.float64)
    torch = _torch_cuda()
    if torch is not None and b.shape[0] >= 512:
        t = torch.from_numpy(b).cuda()
        qm, rm = torch.linalg.qr(t)
        out = (qm * torch.diagonal(rm)[None, :]).cpu().numpy()
        del t, qm, rm
        torch.cuda.empty_cache()
        return out
    qm, rm = np.linalg.qr(b)
    return qm * np.diag(rm)[None, :]


def compute_sqrt_sigma_2(r_mat: np.ndarray, s: float, r: float, base: int, sigma: np.ndarray = None) -> np.ndarray:
    """mp_perturbation.rs:111-139: lower Cholesky factor of
    Sigma_2 = r^2/(2 pi) (Sigma - (b^2+1) T T^t - I),  T = [R; I];  Sigma defaults to s^2 I."""
    rm = np.asarray(r_mat, dtype=np.float64)
    m_bar, nk = rm.shape
    m = m_bar + nk
    c = float(base * base + 1)
    torch = _torch_cuda()
    if torch is not None and m >= 512:
        t = torch.cat([torch.from_numpy(rm).cuda(), torch.eye(nk, dtype=torch.float64, device="cuda")], 0)
        sp = -c * (t @ t.T)
        if sigma is None:
            sp.diagonal().add_(s * s)
        else:
            sp += torch.from_numpy(np.asarray(sigma, dtype=np.float64)).cuda()
        sp.diagonal().sub_(1.0)
        sp *= (r * r) / (2.0 * math.pi)
        out = torch.linalg.cholesky(sp).cpu().numpy()
        del t, sp
        torch.cuda.empty_cache()
        return out
    t = np.vstack([rm, np.eye(nk)])
    sp = -c * (t @ t.T)
    sp += (s * s) * np.eye(m) if sigma is None else np.asarray(sigma, dtype=np.float64)
    sp -= np.eye(m)
    sp *= (r * r) / (2.0 * math.pi)
    return np.linalg.cholesky(sp)
