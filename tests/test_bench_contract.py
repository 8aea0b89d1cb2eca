"""bench.py contract on CPU: the reference arm (`--impl reference`, the oracle port timed on the host cores -- the one place besides
the cpu_baseline leg where bench.py may execute oracle/) prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--lattice", "4,4,4,4", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "SU(3) Wilson HMC MD steps/s" and d["unit"] == "MD steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "MD steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["dtype"] == "f64"


def test_product_fails_loudly_without_a_gpu():
    """No CPU fallback: without a usable GPU the backend constructor raises instead of computing anything."""
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present")
    sys.path.insert(0, os.path.join(ROOT, "gaugefields.jl_b200"))
    import gfb200
    import pytest

    with pytest.raises(Exception):
        gfb200.B200Backend(ngpu=1)
