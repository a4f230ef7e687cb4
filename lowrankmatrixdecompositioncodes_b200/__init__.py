"""lowrankmatrixdecompositioncodes_b200 — Python host-side mirror of the B200-native RSVDPACK hot path.

The product is the C-ABI (include/*.h): librsvd_b200.so (hand-written sm_100a kernels) and the two drop-in
C API libraries librsvd_b200_api{32,64}.so.  This package only binds them with ctypes:

  native.dev        device layer (rsvd_b200.h), device pointers in/out
  api.Api(bits)     the reference's C API (low_rank_svd_rand_decomp_fixed_rank, ...), numpy in/out through the
                    same mat/vec structs a C driver would use
  device            helpers to call the device layer on torch CUDA tensors (torch = memory + streams only)

There is no CPU fallback: importing works anywhere (so the ABI can be inspected), every compute call fails
loudly without a B200.
"""
from . import native  # noqa: F401
from .api import Api  # noqa: F401

__all__ = ["native", "Api"]
