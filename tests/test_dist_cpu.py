"""CPU tests of the N>1 host-side logic over a world_size-2 gloo group: the C row-partition function, which quantities
the row-partitioned pipeline reduces (oracle/dist_twin.py restates csrc/device/pipeline.cu rank by rank), and the global
row/column indexing of Omega.  The device kernels themselves need a GPU (tools/dist_check.py runs the same comparison on
2 B200s through NCCL)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lowrankmatrixdecompositioncodes_b200 import native
    from oracle import dist_twin, rsvd_numpy as O
    from helpers import subspace_sin

    def allreduce(x):
        t = torch.from_numpy(x)
        dist.all_reduce(t)       # in place (shares memory with x)

    out = {}
    m, n, k, p, q, s = 1003, 600, 30, 10, 2, 1
    A, sig = O.make_matrix(m, n, "gap", seed=2, k=k, tail=1e-7)
    r0, rows = native.row_partition(m, world, rank)
    A_loc = np.ascontiguousarray(A[r0:r0 + rows])
    U_loc, S, V = dist_twin.svd_rand_sharded(A_loc, k, p, q, s, 777, allreduce)
    parts = [None] * world
    dist.all_gather_object(parts, (r0, U_loc))
    U = np.vstack([u for _, u in sorted(parts, key=lambda t: t[0])])
    Ur, Sr, Vr = O.low_rank_svd_rand_decomp_fixed_rank(A, k, p, 1, q, s, 777)
    out["sigma"] = float(np.max(np.abs(S - np.diag(Sr)) / np.diag(Sr)))
    out["sinU"], out["sinV"] = subspace_sin(U, Ur), subspace_sin(V, Vr)
    out["orth"] = float(np.abs(U.T @ U - np.eye(k)).max())
    I, T = dist_twin.id_rand_sharded(A_loc, r0, m, k, p, q, s, 777, allreduce)
    Ir, Tr = O.id_rand_decomp_fixed_rank(A, k, p, q, s, 777)
    out["pivots_equal"] = bool(np.array_equal(I, Ir))
    out["T"] = float(np.abs(T - Tr).max())
    out["rows"] = (r0, rows)
    ret[rank] = out
    dist.destroy_process_group()


def test_row_partitioned_pipeline_matches_single_process_twin():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(180)
        assert pr.exitcode == 0
    assert sorted(ret.keys()) == [0, 1]
    assert ret[0]["rows"] == (0, 512) and ret[1]["rows"] == (512, 491)     # 16-row aligned blocks covering all rows
    for r in range(world):
        o = ret[r]
        assert o["sigma"] < 1e-10 and o["sinU"] < 1e-6 and o["sinV"] < 1e-6 and o["orth"] < 1e-10
        assert o["pivots_equal"] and o["T"] < 1e-9
