"""Per-key floating-point setup (not per target): GSO of a short basis
(MatQ::gso, gpv.rs:91) in float64.  Uses torch on the GPU when one is present (setup plumbing), numpy otherwise.
sqrt(Sigma_2) (compute_sqrt_sigma_2, mp_perturbation.rs:111-139) is the library's own blocked Cholesky
(qf_compute_sqrt_sigma_2)."""

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
