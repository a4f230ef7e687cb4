"""tools/dist_check.py — multi-GPU parity check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py
Every rank holds a row block of the same seeded matrix; results of the row-partitioned device pipelines are compared with
the single-process oracle (compiled reference if present, else numpy twin) on rank 0's copy of the full matrix."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from lowrankmatrixdecompositioncodes_b200 import device as D, native  # noqa: E402
from oracle import ref_lib, rsvd_numpy as O  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
lib = native.dev()
assert lib.rsvd_b200_init(local) == 0
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ident = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    buf = C.create_string_buffer(128)
    native.check(lib.rsvd_b200_comm_unique_id(buf))
    ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
ident = ident.cuda()
dist.broadcast(ident, 0)
native.check(lib.rsvd_b200_comm_init(rank, world, bytes(ident.cpu().numpy().tobytes())))
ok = True


def report(name, cond, detail):
    global ok
    ok = ok and bool(cond)
    if rank == 0:
        print("%-46s %s  %s" % (name, "OK  " if cond else "FAIL", detail), flush=True)


def gather_rows(t_cm, m_loc, cols):
    """row-sharded column-major (cols, m_loc) tensors -> full numpy (m, cols) on every rank"""
    parts = [torch.empty_like(t_cm) for _ in range(world)]
    sizes = [None] * world
    dist.all_gather_object(sizes, m_loc)
    outs = []
    for r in range(world):
        buf = torch.empty((cols, sizes[r]), dtype=torch.float64, device="cuda")
        if r == rank:
            buf.copy_(t_cm)
        dist.broadcast(buf, r)
        outs.append(buf.t().cpu().numpy())
    return np.vstack(outs)


L = ref_lib.RefLib(32) if ref_lib.available(32) else None
CASES = [(2000, 1500, 100, 20, 2, 1, "gap"), (3001, 900, 40, 10, 2, 1, "logspace")]
if os.environ.get("DIST_CHECK_WIDE"):      # k + p beyond 2048: Jacobi above 2048 columns, sharded pivoted QR with 2100 rows
    CASES = [(6000, 5000, 2100, 20, 2, 1, "logspace")]
for (m, n, k, p, q, s, spec) in CASES:
    A, sig = O.make_matrix(m, n, spec, seed=3, k=k, tail=1e-7)
    r0, rows = native.row_partition(m, world, rank)
    lib.rsvd_b200_set_option(b"row0", r0)
    lib.rsvd_b200_set_option(b"m_global", m)
    A_loc = D.from_numpy_cm(np.ascontiguousarray(A[r0:r0 + rows, :]))
    torch.cuda.synchronize()
    # ---- SVD ----
    U, S, V = D.svd_rand(A_loc, k, p, 1, q, s, seed=777)
    lib.rsvd_b200_sync()
    Ufull = gather_rows(U, rows, k)
    Sn, Vn = S.cpu().numpy(), D.to_numpy(V)
    if L is not None:
        Ur, Sr, Vr = L.svd_rand(A, k, p, 1, q, s, seed=777)
    else:
        Ur, Sr, Vr = O.low_rank_svd_rand_decomp_fixed_rank(A, k, p, 1, q, s, 777)
    rel = np.max(np.abs(Sn - np.diag(Sr)) / np.diag(Sr))
    e, er = np.linalg.norm(A - (Ufull * Sn) @ Vn.T) / np.linalg.norm(A), np.linalg.norm(A - Ur @ Sr @ Vr.T) / np.linalg.norm(A)
    report("svd_rand %dx%d world=%d (%s)" % (m, n, world, spec), rel < 1e-10 and abs(e - er) <= 0.01 * er,
           "max rel sigma err %.2e recon %.6e (ref %.6e) ||UtU-I|| %.1e" % (rel, e, er, np.abs(Ufull.T @ Ufull - np.eye(k)).max()))
    # ---- two-sided ID + CUR ----
    Icol = torch.empty(n, dtype=torch.float64, device="cuda")
    Irow = torch.empty(m, dtype=torch.float64, device="cuda")
    T = torch.empty((n - k, k), dtype=torch.float64, device="cuda")
    Sm = torch.empty((m - k, k), dtype=torch.float64, device="cuda")
    native.check(lib.rsvd_b200_id_two_sided_rand_dev(A_loc.data_ptr(), rows, n, rows, k, p, q, s, 777, Icol.data_ptr(), Irow.data_ptr(),
                                                     T.data_ptr(), k, Sm.data_ptr(), k))
    lib.rsvd_b200_sync()
    if L is not None:
        Icr, Irr, Tr, Smr = L.id_two_sided_rand(A, k, p, q, s, seed=777)
    else:
        Icr, Irr, Tr, Smr = O.id_two_sided_rand_decomp_fixed_rank(A, k, p, q, s, 777)
    report("id_two_sided %dx%d world=%d" % (m, n, world),
           np.array_equal(Icol.cpu().numpy(), Icr) and np.array_equal(Irow.cpu().numpy(), Irr) and np.abs(D.to_numpy(T) - Tr).max() < 1e-10
           and np.abs(D.to_numpy(Sm) - Smr).max() < 1e-10,
           "Icol eq %s Irow eq %s max|T-Tref| %.2e max|S-Sref| %.2e" % (np.array_equal(Icol.cpu().numpy(), Icr), np.array_equal(Irow.cpu().numpy(), Irr),
                                                                       np.abs(D.to_numpy(T) - Tr).max(), np.abs(D.to_numpy(Sm) - Smr).max()))
    Cm = D.new_cm(rows, k); Um = D.new_cm(k, k); Rm = D.new_cm(k, n)
    native.check(lib.rsvd_b200_cur_rand_dev(A_loc.data_ptr(), rows, n, rows, k, p, q, s, 777, Cm.data_ptr(), rows, Um.data_ptr(), k, Rm.data_ptr(), k))
    lib.rsvd_b200_sync()
    Cfull = gather_rows(Cm, rows, k)
    if L is not None:
        Cr, Uc, Rr = L.cur_rand(A, k, p, q, s, seed=777)
    else:
        Cr, Uc, Rr = O.cur_rand_decomp_fixed_rank(A, k, p, q, s, 777)
    report("cur %dx%d world=%d" % (m, n, world), np.array_equal(Cfull, Cr) and np.array_equal(D.to_numpy(Rm), Rr) and
           np.abs(D.to_numpy(Um) - Uc).max() <= 1e-6 * np.abs(Uc).max(),   # U solves an (R R^T) system: cond^2-sensitive
           "C eq %s R eq %s max|U-Uref|/max|U| %.2e" % (np.array_equal(Cfull, Cr), np.array_equal(D.to_numpy(Rm), Rr), np.abs(D.to_numpy(Um) - Uc).max() / np.abs(Uc).max()))
    # ---- blocked QB (rank mode and device-side tolerance mode) ----
    for (kstep, nstep, tol) in [(20, 4, 0.0), (20, 0, float(np.linalg.norm(A)) * 0.3)]:
        Aw = A_loc.clone()
        cap = (min(m, n) // kstep) * kstep if nstep <= 0 else kstep * nstep
        Qd = torch.zeros((cap, rows), dtype=torch.float64, device="cuda")
        Bd = torch.zeros((n, cap), dtype=torch.float64, device="cuda")
        fr = C.c_longlong(0)
        torch.cuda.synchronize()
        native.check(lib.rsvd_b200_randqb_dev(Aw.data_ptr(), rows, n, rows, kstep, nstep, tol, q, s, 777, Qd.data_ptr(), rows, Bd.data_ptr(), cap, C.byref(fr)))
        lib.rsvd_b200_sync()
        f = int(fr.value)
        Qfull = gather_rows(Qd[:f].contiguous(), rows, f)
        Bn = Bd.t().cpu().numpy()[:f, :]
        if L is not None:
            frr, Qr, Br = L.randQB_pb_new(A, kstep, nstep, tol, q, s, seed=777)
        else:
            frr, Qr, Br = O.randQB_pb_new(A, kstep, nstep, tol, q, s, 777)
        d = np.linalg.norm(Qfull @ Bn - Qr @ Br) / np.linalg.norm(A)
        report("randQB kstep=%d nstep=%d tol=%.3g world=%d" % (kstep, nstep, tol, world), f == frr and d < 1e-11, "frank %d (ref %d) ||QB-QrBr||/||A|| %.2e" % (f, frr, d))

# ---- a numerically SINGULAR row-partitioned panel: exact rank 30 sketched with l = 60 columns (TSQR with explicit Q) ----------
m, n, k, p = 4000, 700, 40, 20
rng = np.random.default_rng(4)
A = (np.linalg.qr(rng.standard_normal((m, 30)))[0] * np.logspace(0, -2, 30)) @ np.linalg.qr(rng.standard_normal((n, 30)))[0].T
r0, rows = native.row_partition(m, world, rank)
lib.rsvd_b200_set_option(b"row0", r0)
lib.rsvd_b200_set_option(b"m_global", m)
A_loc = D.from_numpy_cm(np.ascontiguousarray(A[r0:r0 + rows, :]))
torch.cuda.synchronize()
U, S, V = D.svd_rand(A_loc, k, p, 1, 2, 1, seed=777)
lib.rsvd_b200_sync()
Ufull, Sn, Vn = gather_rows(U, rows, k), S.cpu().numpy(), D.to_numpy(V)
strue = np.logspace(0, -2, 30)
rel = np.max(np.abs(Sn[:30] - strue) / strue)
e = np.linalg.norm(A - (Ufull * Sn) @ Vn.T) / np.linalg.norm(A)
report("svd_rand of an exact rank-30 matrix, l=60, world=%d" % world, rel < 1e-10 and Sn[30:].max() < 1e-12 and e < 1e-12 and np.abs(Ufull.T @ Ufull - np.eye(k)).max() < 1e-12,
       "max rel sigma err (first 30) %.2e  max sigma beyond the rank %.1e  recon %.1e  ||UtU-I|| %.1e  qr path %d" % (
           rel, Sn[30:].max(), e, np.abs(Ufull.T @ Ufull - np.eye(k)).max(), lib.rsvd_b200_get_option(b"last_qr_path")))

t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK", "PASS" if t.item() > 0 else "FAIL", flush=True)
lib.rsvd_b200_comm_destroy()
dist.destroy_process_group()
sys.exit(0 if t.item() > 0 else 1)
