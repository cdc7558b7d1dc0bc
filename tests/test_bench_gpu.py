"""GPU: bench.py end to end on the small workload — the JSON line carries every key of the contract, the closed-form
parity check passes through both the resident and the host-streamed step, and the secondary measurements run."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
def test_bench_small_workload_contract_and_parity(cuda_device):
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--workload", "full256", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "parity_check", "clocks", "extras"):
        assert key in d, key
    assert d["gpu_launches"] > 0 and d["value"] > 0 and d["dtype"] == "f64"
    pc = d["parity_check"]
    assert pc["checked"] and pc["passed"] and pc["bitwise_repeatable"], pc
    assert pc["max_rel_err"] <= 1e-12 and pc["host_step_max_rel_err"] <= 1e-12, pc
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["frac"] > 0 and r["peak"] > 0  # (256^3 is L2-sized: fractions mean something at 1024^3)
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 4 * 8 * 256**3 and e["d2h_bytes_per_step"] > 0 and e["h2d_ceiling_gbs"] > 0
    x = d["extras"]
    assert x["c5_series"]["streamed_vs_single_chunk_max_rel_err"] <= 1e-12
    assert x["c3_profiles512"]["profiles_xyz_ms"] > 0 and x["c2_from_amr_256"]["ms"] > 0 and x["prolong_512"]["ms_slowest_rank"] > 0
