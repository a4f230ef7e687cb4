"""Phase timings of one end-to-end API call at BASELINE configs[1] (host mat in pinned memory): RSVD_B200_VERBOSE=1 makes the
host layer print alloc / upload+device / alloc+download; RSVD_B200_VERBOSE=3 adds the device phases from CUDA events."""
import ctypes as C, os, sys, time
os.environ.setdefault("RSVD_B200_VERBOSE", "1")
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lowrankmatrixdecompositioncodes_b200 as pkg
from lowrankmatrixdecompositioncodes_b200 import native
lib = native.dev(); lib.rsvd_b200_init(0)
m, n, k, p = 50000, 20000, 500, 20
api = pkg.Api(32)
A = torch.randn((n, m), dtype=torch.float64, device="cuda")
M = api.lib.matrix_new(m, n)
native.check(lib.rsvd_b200_d2h(C.cast(M.contents.d, C.c_void_p), A.data_ptr(), m * n))
del A; torch.cuda.empty_cache()
lib.rsvd_b200_set_option(b"verbose", int(os.environ["RSVD_B200_VERBOSE"]))
for it in range(3):
    U, S, V = api.PM(), api.PM(), api.PM(); fr = api.I(0)
    t0 = time.perf_counter()
    api.lib.low_rank_svd_rand_decomp_fixed_rank(M, k, p, 1, 2, 1, C.byref(fr), C.byref(U), C.byref(S), C.byref(V))
    print("call %d: %.1f ms" % (it, (time.perf_counter() - t0) * 1e3), flush=True)
    for x in (U, S, V): api.lib.matrix_delete(x)
