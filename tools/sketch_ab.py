"""tools/sketch_ab.py — the fused sketch Y = A * Omega (gemm_tma_kernel<.,PHILOX,.>) with and without 2-CTA clusters sharing
the generated Omega stages, timed with CUDA events on the library's stream (BASELINE configs[1] shape by default), plus a
bit-exact comparison of the two results and of the plain NN pass for reference.
    python tools/sketch_ab.py [m n l]"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from lowrankmatrixdecompositioncodes_b200 import device as D, native  # noqa: E402

m, n, l = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (50000, 20000, 520)
lib = native.dev()
assert lib.rsvd_b200_init(0) == 0
st = D.stream()
A = torch.randn((n, m), dtype=torch.float64, device="cuda")
B = torch.randn((l, n), dtype=torch.float64, device="cuda")
Y = {0: torch.empty((l, m), dtype=torch.float64, device="cuda"), 1: torch.empty((l, m), dtype=torch.float64, device="cuda")}
torch.cuda.synchronize()


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    lib.rsvd_b200_sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
    for _ in range(reps):
        fn()
    with torch.cuda.stream(st):
        e1.record()
    lib.rsvd_b200_sync()
    return e0.elapsed_time(e1) / reps


flops = 2.0 * m * n * l
lib.rsvd_b200_set_option(b"verbose", 2)
for nc in (1, 0):
    lib.rsvd_b200_set_option(b"no_sketch_cluster", nc)
    native.check(lib.rsvd_b200_sketch(b"N", m, l, n, A.data_ptr(), m, 777, 1, n, 0, Y[nc].data_ptr(), m))
lib.rsvd_b200_set_option(b"verbose", 0)
for nc in (1, 0, 1, 0):
    lib.rsvd_b200_set_option(b"no_sketch_cluster", nc)
    t = timed(lambda: native.check(lib.rsvd_b200_sketch(b"N", m, l, n, A.data_ptr(), m, 777, 1, n, 0, Y[nc].data_ptr(), m)))
    print("sketch %s: %.3f ms  %.2f TFLOP/s" % ("single CTAs   " if nc else "2-CTA clusters", t, flops / t / 1e9), flush=True)
t = timed(lambda: D.gemm("N", "N", m, l, n, A, m, B, n, Y[1], m))
print("plain NN pass (stored B): %.3f ms  %.2f TFLOP/s" % (t, flops / t / 1e9))
lib.rsvd_b200_set_option(b"no_sketch_cluster", 0)
native.check(lib.rsvd_b200_sketch(b"N", m, l, n, A.data_ptr(), m, 777, 1, n, 0, Y[0].data_ptr(), m))
lib.rsvd_b200_set_option(b"no_sketch_cluster", 1)
native.check(lib.rsvd_b200_sketch(b"N", m, l, n, A.data_ptr(), m, 777, 1, n, 0, Y[1].data_ptr(), m))
lib.rsvd_b200_sync()
print("cluster result bit-identical to the single-CTA result:", bool(torch.equal(Y[0], Y[1])))
# left sketch (ID): C = A^T * Omega' with the linear index running along the columns
Yt = {0: torch.empty((l, n), dtype=torch.float64, device="cuda"), 1: torch.empty((l, n), dtype=torch.float64, device="cuda")}
for nc in (1, 0):
    lib.rsvd_b200_set_option(b"no_sketch_cluster", nc)
    t = timed(lambda: native.check(lib.rsvd_b200_sketch(b"T", n, l, m, A.data_ptr(), m, 777, l, 1, 0, Yt[nc].data_ptr(), n)))
    print("left sketch %s: %.3f ms  %.2f TFLOP/s" % ("single CTAs   " if nc else "2-CTA clusters", t, flops / t / 1e9), flush=True)
lib.rsvd_b200_sync()
print("left sketch bit-identical:", bool(torch.equal(Yt[0], Yt[1])))
lib.rsvd_b200_set_option(b"no_sketch_cluster", 0)
