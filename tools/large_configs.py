"""tools/large_configs.py — BASELINE configs[2] (blockrand QB, tolerance mode) and configs[3] (two-sided ID + CUR) at
(scaled) benchmark sizes on one B200, device resident, matrix generated in HBM; prints times and streamed error checks.
    python tools/large_configs.py c3 [rows] | c4 [rows] | abi64"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import lowrankmatrixdecompositioncodes_b200 as pkg  # noqa: E402
from lowrankmatrixdecompositioncodes_b200 import device as D, native  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
lib = native.dev()
assert lib.rsvd_b200_init(local) == 0
what = sys.argv[1] if len(sys.argv) > 1 else "c3"
if world > 1:   # row-partitioned run under torchrun: one rank per GPU, NCCL for the n x l / l x l sums
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ident = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = C.create_string_buffer(128)
        native.check(lib.rsvd_b200_comm_unique_id(buf))
        ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    ident = ident.cuda()
    dist.broadcast(ident, 0)
    native.check(lib.rsvd_b200_comm_init(rank, world, bytes(ident.cpu().numpy().tobytes())))


def gen(m, n, r=1280, lo=-3.0, noise=1e-8, seed=0, m_global=None, row0=0):
    """Rows [row0, row0 + m) of the m_global x n matrix A = X diag(sigma) W^T + noise, sigma = logspace(1, lo, r), built in HBM in
    column slabs (torch only generates data).  Every rank draws the SAME global X and noise and keeps its rows, so the matrix —
    and therefore every index set — is identical for any number of GPUs."""
    mg = m if m_global is None else m_global
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    X = torch.randn((mg, r), dtype=torch.float64, device="cuda", generator=g)[row0:row0 + m].clone() / mg ** 0.5
    W = torch.randn((n, r), dtype=torch.float64, device="cuda", generator=g) / n ** 0.5
    sig = torch.logspace(1, lo, r, dtype=torch.float64, device="cuda")
    A = torch.empty((n, m), dtype=torch.float64, device="cuda")
    step = 256
    for j0 in range(0, n, step):
        j1 = min(n, j0 + step)
        torch.matmul(W[j0:j1] * sig, X.t(), out=A[j0:j1])
        if noise:
            A[j0:j1].add_(torch.randn((j1 - j0, mg), dtype=torch.float64, device="cuda", generator=g)[:, row0:row0 + m], alpha=noise)
    del X, W
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return A, sig


def digest(t):
    """order-sensitive checksum of an index vector (to compare runs on different GPU counts in the logs)"""
    v = t.long()
    w = torch.arange(1, v.numel() + 1, device=v.device, dtype=torch.long)
    return int(((v * w) % 1000003).sum().item() % 1000003)


def sync():
    lib.rsvd_b200_sync()
    torch.cuda.synchronize()


if what == "c3":
    # BASELINE configs[2]: blockrand QB, autorank TOL = 1e-6 (ABSOLUTE Frobenius norm, RRA:1773-1775), kstep = 200, on a matrix
    # whose residual really drops below the tolerance: exact rank 600 + a 1e-12 noise floor (||noise||_F = 1e-12 sqrt(mn) = 1e-7)
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    n, kstep, q, s = 50000, 200, 2, 1
    A, sig = gen(m, n, r=600, lo=-2.0, noise=1e-12)
    normA = lib.rsvd_b200_frob_norm(A.data_ptr(), m, m, n)
    tol = 1e-6
    cap = 2000
    Q = torch.zeros((cap, m), dtype=torch.float64, device="cuda")
    B = torch.zeros((n, cap), dtype=torch.float64, device="cuda")
    fr = C.c_longlong(0)
    print("C3: randQB_pb_new tolerance mode on %d x %d (%.1f GB), kstep=%d q=%d, ||A||_F=%.4f, TOL=%.1e (absolute)" % (m, n, 8e-9 * m * n, kstep, q, normA, tol), flush=True)
    lib.rsvd_b200_set_option(b"verbose", int(os.environ.get("RSVD_B200_VERBOSE", "1")))     # 3: per-phase CUDA-event times of every block step
    sync()
    t0 = time.time()
    native.check(lib.rsvd_b200_randqb_dev(A.data_ptr(), m, n, m, kstep, 0, tol, q, s, 777, Q.data_ptr(), m, B.data_ptr(), cap, C.byref(fr)))
    sync()
    dt = time.time() - t0
    lib.rsvd_b200_set_option(b"verbose", 0)
    f = int(fr.value)
    res = lib.rsvd_b200_frob_norm(A.data_ptr(), m, m, n)     # A now holds the residual A - QB
    nblk = f // kstep
    passes = nblk * (2 * q + 3)
    peak = max(lib.rsvd_b200_fp64_peak_tflops(4000, 0), lib.rsvd_b200_fp64_peak_tflops(4000, 1))
    bound = (2 * q + 3) * 2.0 * m * n * kstep / (peak * 1e12)
    print("   frank=%d  time %.3f s = %.3f s per block step  ||A-QB||_F=%.4g (< TOL: %s)  %.1f TFLOP/s over %d width-%d passes" %
          (f, dt, dt / nblk, res, res < tol, passes * 2.0 * m * n * kstep / dt / 1e12, passes, kstep), flush=True)
    print("   per-step bound 7 passes x 2mn*kstep / measured FP64 peak (%.1f TFLOP/s) = %.3f s; measured / bound = %.3f" % (peak, bound, dt / nblk / bound), flush=True)
    Qf = Q[:f]
    print("   ||Q^T Q - I||_max = %.2e" % (Qf @ Qf.t() - torch.eye(f, dtype=torch.float64, device="cuda")).abs().max().item())

elif what == "c4":
    mg = int(sys.argv[2]) if len(sys.argv) > 2 else 200000      # GLOBAL rows
    n, k, p, q, s = 50000, 1000, 20, 2, 1
    r0, m = native.row_partition(mg, world, rank)
    lib.rsvd_b200_set_option(b"row0", r0)
    lib.rsvd_b200_set_option(b"m_global", mg)
    A, sig = gen(m, n, r=1536, lo=-2.0, m_global=mg, row0=r0)
    if rank == 0:
        print("C4: id_two_sided_rand + cur_rand on %d x %d (%.1f GB) over %d GPU(s), k=%d p=%d q=%d" % (mg, n, 8e-9 * mg * n, world, k, p, q), flush=True)
    Icol = torch.empty(n, dtype=torch.float64, device="cuda"); Irow = torch.empty(mg, dtype=torch.float64, device="cuda")
    T = torch.empty((n - k, k), dtype=torch.float64, device="cuda"); Sm = torch.empty((mg - k, k), dtype=torch.float64, device="cuda")
    sync()
    t0 = time.time()
    native.check(lib.rsvd_b200_id_two_sided_rand_dev(A.data_ptr(), m, n, m, k, p, q, s, 777, Icol.data_ptr(), Irow.data_ptr(), T.data_ptr(), k, Sm.data_ptr(), k))
    sync()
    dt = time.time() - t0
    flops = (1 + 2 * q) * 2.0 * mg * n * (k + p)
    t0 = time.time()
    lib.rsvd_b200_set_option(b"verbose", 3 if rank == 0 else 0)
    native.check(lib.rsvd_b200_id_two_sided_rand_dev(A.data_ptr(), m, n, m, k, p, q, s, 777, Icol.data_ptr(), Irow.data_ptr(), T.data_ptr(), k, Sm.data_ptr(), k))
    sync()
    lib.rsvd_b200_set_option(b"verbose", 0)
    dt2 = time.time() - t0
    if rank == 0:
        print("   two-sided ID: %.3f s first call, %.3f s second call  (%.1f TFLOP/s aggregate over the %d sketch/power passes alone; GEMM floor at 36.9 TFLOP/s per GPU: %.3f s)"
              % (dt, dt2, flops / dt2 / 1e12, 1 + 2 * q, flops / 36.9e12 / world), flush=True)
        print("   digests (identical for every GPU count): Icol %d  Irow %d  Icol[:6] %s  Irow[:6] %s" % (digest(Icol), digest(Irow), Icol[:6].long().tolist(), Irow[:6].long().tolist()), flush=True)
    ic, ir = Icol.long(), Irow.long()
    assert torch.equal(torch.sort(ic).values, torch.arange(n, device="cuda")) and torch.equal(torch.sort(ir).values, torch.arange(mg, device="cuda"))
    if world > 1:   # replicated outputs must be identical on every rank
        ref_ic = ic.clone(); dist.broadcast(ref_ic, 0)
        assert torch.equal(ref_ic, ic), "Icol differs between ranks"
    # column-ID error on a row sample:  A(rows, Icol) ~ A(rows, Icol[:k]) [I T]
    rows = torch.randperm(m, device="cuda")[:4096]
    As = A.t()[rows]                                   # 4096 x n
    Ck = As[:, ic[:k]]
    err = (As[:, ic[k:]] - Ck @ T.t()).norm() / As.norm()
    opt = torch.sqrt((sig[k:] ** 2).sum()) / torch.sqrt((sig ** 2).sum())
    if rank == 0:
        print("   column ID rel. error on 4096 sampled local rows: %.4e  (optimal rank-%d: %.4e)" % (err.item(), k, opt.item()), flush=True)
    del T, Sm, Ck
    torch.cuda.empty_cache()       # hand torch's cached blocks back: at 160 GB every GB counts
    Cm = D.new_cm(m, k); Um = D.new_cm(k, k); Rm = D.new_cm(k, n)
    sync()
    t0 = time.time()
    native.check(lib.rsvd_b200_cur_rand_dev(A.data_ptr(), m, n, m, k, p, q, s, 777, Cm.data_ptr(), m, Um.data_ptr(), k, Rm.data_ptr(), k))
    sync()
    dt = time.time() - t0
    Cs = Cm.t()[rows]                                  # 4096 x k
    err = (As - Cs @ Um.t() @ Rm.t()).norm() / As.norm()
    t0 = time.time()
    native.check(lib.rsvd_b200_cur_rand_dev(A.data_ptr(), m, n, m, k, p, q, s, 777, Cm.data_ptr(), m, Um.data_ptr(), k, Rm.data_ptr(), k))
    sync()
    dt2 = time.time() - t0
    if rank == 0:
        print("   CUR (includes its own two-sided ID): %.3f s first call, %.3f s second call" % (dt, dt2), flush=True)
        print("   CUR rel. error on the row sample: %.4e" % err.item(), flush=True)

elif what == "abi64":
    # int64 ABI end to end with m*n > 2^31 elements (the reason multi_core_mkl_code_64bit exists)
    m, n, k, p = 72000, 30000, 200, 20
    api = pkg.Api(64)
    M = api.lib.matrix_new(m, n)
    A, sig = gen(m, n, r=512, lo=-3.0)
    native.check(lib.rsvd_b200_d2h(C.cast(M.contents.d, C.c_void_p), A.data_ptr(), m * n))
    del A
    torch.cuda.empty_cache()
    print("abi64: %d x %d = %.3g elements (> 2^31), %.1f GB host matrix" % (m, n, float(m) * n, 8e-9 * m * n), flush=True)
    Um, Sm, Vm = api.PM(), api.PM(), api.PM()
    fr = api.I(0)
    t0 = time.time()
    api.lib.low_rank_svd_rand_decomp_fixed_rank(M, k, p, 1, 2, 1, C.byref(fr), C.byref(Um), C.byref(Sm), C.byref(Vm))
    dt = time.time() - t0
    api.check()
    S = api.from_mat(Sm, free=False)
    rel = np.max(np.abs(np.diag(S) - sig[:k].cpu().numpy()) / sig[:k].cpu().numpy())
    api.lib.use_low_rank_svd_for_approximation(M, Um, Sm, Vm)
    print("   API call %.3f s; max rel deviation of sigma from the construction: %.2e; percent error %.4f" % (dt, rel, api.lib.rsvd_b200_api_last_percent_error()), flush=True)
if rank == 0:
    print("status:", lib.rsvd_b200_status(), lib.rsvd_b200_last_error())
if world > 1:
    lib.rsvd_b200_comm_destroy()
    dist.destroy_process_group()
