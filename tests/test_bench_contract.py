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
    assert d["cpu_baseline"]["same_config"] is True and "full lattice 4x4x4x4" in d["cpu_baseline"]["sample"]


def test_reference_arm_uses_all_cores_even_under_torchrun_env():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the arm must set and report the thread count it really used
    (round-1 VERDICT: 'reports cores 32 while running on one thread')."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "flow32", "--lattice", "4,4,4,4", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["unit"] == "flow steps/s"


def test_reference_arm_bounds_its_sample_and_says_so():
    """A lattice the oracle cannot finish in the budget is sampled on a t-sub-volume and flagged same_config = false."""
    sys.path.insert(0, ROOT)
    import bench

    value, ms, cb = bench.cpu_steps_per_s("md64", (16, 16, 16, 64), 6.0, 1.0, 1, 0, budget_s=0.05)
    assert cb["same_config"] is False and "sub-volume" in cb["sample"] and cb["sample_sites"] < 16 * 16 * 16 * 64
    assert abs(value * ms * 1e-3 - 1.0) < 1e-9  # ms_per_step is the scaled-to-full-lattice time of one step


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
