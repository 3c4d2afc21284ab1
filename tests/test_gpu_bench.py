"""GPU smoke test of bench.py itself: the driver's measurement script must run end to end (sampling leg, end-to-end leg,
TF32 leg, training leg with geometry prefetch + FlatAdam, CPU-oracle baseline and parity self-check, reference arm) and
print one JSON line with the contract's keys.  Small configuration so that it takes seconds."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--batch", "4", "--points",
                        "512", "--preset", "SPVD_TINY", "--cpu-baseline-steps", "1", *extra],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_bench_runs_end_to_end_on_a_small_configuration():
    d = _run("--train-steps", "3")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "parity_rel_err", "train"):
        assert k in d, k
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["roofline"]["bound"] == "tensor" and 0 < d["roofline"]["frac"] < 1
    assert d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["parity"]["ok"] and d["parity_rel_err"] < 1e-3
    assert d["train"]["ms_per_step"] > 0 and d["train"]["steps"] == 3


def test_reference_arm_prints_the_same_metric():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert d["impl"] == "reference" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["value"] == d["value"]
