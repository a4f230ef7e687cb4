"""bench.py contract (reference arm, runs on the host cores — no GPU): exactly one JSON line on stdout carrying the keys the
driver reads.  The default reference arm runs BASELINE configs[1] in full for every warm-up and timed step (about 20 s per call
on the GPU box's host); the test passes --rows to keep the CPU suite short."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    from oracle import ref_lib
    if not ref_lib.available(32):
        pytest.skip("compiled reference not present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--rows", "2500"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["unit"] == "TFLOP/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None
    assert d["steps"] == 2 and d["warmup"] == 1 and len(d["step_s"]) == 2          # honest counts: every call was made
    assert d["config"]["rows_per_gpu"] == 2500 and d["config"]["global_shape"] == [2500, 20000]
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config(1, 2500)                                # both arms print the same config object
