"""The CPU-side contract of bench.py: `--impl reference` prints ONE JSON line with the keys the driver reads, runs without a GPU,
and non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra=None, extra_args=()):
    env = dict(os.environ, **(env_extra or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-sample", "6", "--cpu-procs", "2", *extra_args], capture_output=True, text=True, timeout=300, env=env,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [ln for ln in out.stdout.splitlines() if ln.strip()]


import pytest


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_reference_arm_line(kind):
    if kind == "reference":
        sys.path.insert(0, ROOT)
        from oracle.ref_import import reference_available

        if not reference_available(travel_only=True):
            pytest.skip("no reference install under baseline/_ref")
    lines = run(extra_args=("--cpu-kind", kind))
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "GP/s" and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] == 2
    assert d["e2e"] == {"value": d["value"], "unit": "GP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_reference_arm_other_ranks_are_silent():
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
