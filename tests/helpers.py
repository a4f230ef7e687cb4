"""Shared parity metrics."""
import numpy as np


def subspace_sin(U, V):
    """sin of the largest principal angle between span(U) and span(V) (orthonormal columns)."""
    if U.shape[1] != V.shape[1]:
        s = np.linalg.svd(U.T @ V, compute_uv=False)
        return float(np.sqrt(max(0.0, 1.0 - min(s.min(), 1.0) ** 2)))
    return float(np.linalg.norm(V - U @ (U.T @ V), 2))     # ||(I - U U^T) V||_2: accurate for small angles


def col_alignment(U, V):
    """max_i (1 - |u_i . v_i|): per-vector agreement up to sign."""
    return float(np.max(1.0 - np.abs(np.sum(U * V, axis=0))))


def rel_sigma_err(S, Sref):
    s, r = np.diag(S) if S.ndim == 2 else S, np.diag(Sref) if Sref.ndim == 2 else Sref
    return float(np.max(np.abs(s - r) / np.abs(r)))


def recon_err(A, U, S, V):
    return float(np.linalg.norm(A - U @ S @ V.T) / np.linalg.norm(A))
