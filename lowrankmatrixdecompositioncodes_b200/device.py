"""Call the device layer on torch CUDA tensors.  torch supplies device memory and stream interop only.

Column-major convention: an m x n column-major matrix with leading dimension m is a contiguous torch tensor of
shape (n, m) (its transpose view is the mathematical matrix).  `cm(t)` below returns (ptr, rows, cols, ld) for such
a tensor; `new_cm(m, n)` allocates one."""
import ctypes as C

import torch

from . import native


def stream():
    """The library's compute stream as a torch ExternalStream (events recorded under it see our kernels)."""
    return torch.cuda.ExternalStream(native.dev().rsvd_b200_stream())


def new_cm(m, n, device="cuda"):
    return torch.empty((n, m), dtype=torch.float64, device=device)


def from_numpy_cm(a, device="cuda"):
    """numpy (m, n) -> column-major device tensor (shape (n, m))."""
    return torch.from_numpy(a).to(device=device, dtype=torch.float64).t().contiguous()


def to_numpy(t_cm):
    """column-major device tensor (n, m) -> numpy (m, n)."""
    return t_cm.t().cpu().numpy()


def ptr(t):
    assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
    return t.data_ptr()


def gemm(ta, tb, m, n, k, A, lda, B, ldb, Cm, ldc, alpha=1.0, beta=0.0):
    rc = native.dev().rsvd_b200_gemm(ta.encode(), tb.encode(), m, n, k, alpha, ptr(A), lda, ptr(B), ldb, beta, ptr(Cm), ldc)
    native.check(rc)


def svd_rand(A_cm, k, p, vnum=1, q=2, s=1, seed=777, omega=None):
    """A_cm: (n, m) tensor = column-major m x n.  Returns U_cm (k, m), S (k,), V_cm (k, n) on the device."""
    n, m = A_cm.shape
    U = new_cm(m, k)
    V = new_cm(n, k)
    S = torch.empty(k, dtype=torch.float64, device=A_cm.device)
    rc = native.dev().rsvd_b200_svd_rand_dev(ptr(A_cm), m, n, m, k, p, vnum, q, s, seed,
                                             ptr(omega) if omega is not None else None, ptr(U), m, ptr(S), ptr(V), n)
    native.check(rc)
    return U, S, V
