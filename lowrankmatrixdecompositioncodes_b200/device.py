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


class ordered:
    """Stream ordering between torch and the library: the library launches on its own non-blocking stream, so work queued on
    torch's current stream (the producers of the inputs) must be finished before the library reads, and torch must not read
    the outputs before the library has written them.  Both are expressed as event waits between the two streams — no host
    synchronisation."""

    def __enter__(self):
        self.lib_stream = stream()
        self.lib_stream.wait_stream(torch.cuda.current_stream())
        return self

    def __exit__(self, *exc):
        torch.cuda.current_stream().wait_stream(self.lib_stream)
        return False


def gemm(ta, tb, m, n, k, A, lda, B, ldb, Cm, ldc, alpha=1.0, beta=0.0):
    with ordered():
        rc = native.dev().rsvd_b200_gemm(ta.encode(), tb.encode(), m, n, k, alpha, ptr(A), lda, ptr(B), ldb, beta, ptr(Cm), ldc)
    native.check(rc)


def svd_rand(A_cm, k, p, vnum=1, q=2, s=1, seed=777, omega=None):
    """A_cm: (n, m) tensor = column-major m x n.  Returns U_cm (k, m), S (k,), V_cm (k, n) on the device."""
    n, m = A_cm.shape
    U = new_cm(m, k)
    V = new_cm(n, k)
    S = torch.empty(k, dtype=torch.float64, device=A_cm.device)
    with ordered():
        rc = native.dev().rsvd_b200_svd_rand_dev(ptr(A_cm), m, n, m, k, p, vnum, q, s, seed,
                                                 ptr(omega) if omega is not None else None, ptr(U), m, ptr(S), ptr(V), n)
    native.check(rc)
    return U, S, V


class DeviceMatrix:
    """A column-major m x n matrix in library-owned device memory (rsvd_b200_load_binary_dev).  `free()` releases it."""

    def __init__(self, ptr_, m, n):
        self.ptr, self.m, self.n = ptr_, m, n

    def free(self):
        if self.ptr:
            native.dev().rsvd_b200_dev_free(self.ptr)
            self.ptr = None

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.n, self.m), "typestr": "<f8", "data": (self.ptr, False), "version": 2}

    def to_torch(self):
        """an owned (n, m) torch tensor (column-major convention of this module) with the same contents"""
        native.dev().rsvd_b200_sync()
        if self.m * self.n == 0:
            return new_cm(self.m, self.n)
        return torch.as_tensor(self, device="cuda").clone()


def load_binary(path, index_bits=32):
    """Reference-format binary matrix file -> device memory, no host copy of the matrix (include/rsvd_b200.h)."""
    p, m, n = C.c_void_p(), C.c_longlong(), C.c_longlong()
    rc = native.dev().rsvd_b200_load_binary_dev(path.encode(), index_bits, C.byref(p), C.byref(m), C.byref(n))
    native.check(rc)
    return DeviceMatrix(p.value, m.value, n.value)


def store_binary(path, A_ptr, lda, m, n, index_bits=32):
    native.check(native.dev().rsvd_b200_store_binary_dev(path.encode(), index_bits, A_ptr, lda, m, n))
