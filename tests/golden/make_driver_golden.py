"""Generates tests/golden/driver{2,4}_ref.txt: the numeric result lines printed by the reference's OWN drivers 2 and 4
(/root/reference/multi_core_mkl_code/driver_multi_core_mkl{2,4}.c) linked with the reference itself (oracle/_ref/refdrv,
built by oracle/build_ref.sh).  The GPU test runs the same driver sources relinked against the B200 libraries and compares.
Run from the repo root in the build container:  python tests/golden/make_driver_golden.py      (about 2-3 minutes of CPU)."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import rsvd_numpy as O  # noqa: E402

KEEP = re.compile(r"output rank is|percent_error|percent error|norm\(|size of|sizes of")


def driver4_matrix():
    A, _ = O.make_matrix(1000, 1000, "logspace", seed=3)
    return A


def run_driver(exe, which, tmp):
    if which == 4:   # loads ../data/A_mat_1kx1k.bin (driver_multi_core_mkl4.c:17)
        os.makedirs(os.path.join(tmp, "data"), exist_ok=True)
        O.write_matrix_binary(driver4_matrix(), os.path.join(tmp, "data", "A_mat_1kx1k.bin"), 32)
        cwd = os.path.join(tmp, "run")
    else:            # appends to timings/driver_multi_core_mkl2.txt (driver_multi_core_mkl2.c:24)
        cwd = tmp
        os.makedirs(os.path.join(tmp, "timings"), exist_ok=True)
    os.makedirs(cwd, exist_ok=True)
    out = subprocess.run([exe], cwd=cwd, capture_output=True, text=True, timeout=3600)
    return out.returncode, [l.strip() for l in out.stdout.splitlines() if KEEP.search(l)]


if __name__ == "__main__":
    for which in (2, 4):
        exe = os.path.join(ROOT, "oracle", "_ref", "refdrv", "driver_multi_core_mkl%d" % which)
        with tempfile.TemporaryDirectory() as tmp:
            rc, lines = run_driver(exe, which, tmp)
        with open(os.path.join(ROOT, "tests", "golden", "driver%d_ref.txt" % which), "w") as f:
            f.write("# exit code %d (driver 4: the reference aborts inside cur_decomp_fixed_rank_or_prec with 'free(): invalid pointer')\n" % rc)
            f.write("\n".join(lines) + "\n")
        print("driver %d: rc=%d, %d lines" % (which, rc, len(lines)))
