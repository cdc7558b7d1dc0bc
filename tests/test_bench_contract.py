"""CPU: the driver-facing contract of bench.py that can be checked without a GPU — the reference arm
(`--impl reference`: the unmodified reference where /root/reference exists, its NumPy port elsewhere, on a bounded sample) prints ONE JSON line with the agreed keys,
and non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "full256", "--steps", "1",
                           "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)


def test_reference_arm_prints_one_json_line_with_contract_keys():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "Gcells/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    from oracle import ref_harness

    kind = "reference" if ref_harness.reference_available() else "port"  # the unmodified reference where it exists
    assert d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["value"] > 0


def test_reference_arm_is_silent_on_other_ranks():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""
