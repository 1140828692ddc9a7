"""CPU-side check of the bench contract's reference arm: `bench.py --impl reference` (the C oracle port on the host cores; the Rust crate
cannot be built here) prints ONE JSON line with the keys the driver reads, on a bounded sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-log2", "10"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "leaves/sec tree build" and d["unit"] == "leaves/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "2^10 users" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "leaves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_work_model_of_the_roofline():
    """bench.py::executed_mac32: windows per blinding and MAC32 per padding node for the instantiated comb windows."""
    sys.path.insert(0, ROOT)
    import bench
    leaf24, pad24, merge24 = bench.executed_mac32(24, 24)
    leaf26, pad26, _ = bench.executed_mac32(26, 24)
    assert round(pad24) == 7411 and round(merge24) == 2947       # 10 mixed adds + 1 product + compress share; full add + compress share
    assert abs(pad24 - pad26 - 7 * 72) < 1e-6 and abs(leaf24 - leaf26 - 7 * 72) < 1e-6  # one mixed addition (7 field products) fewer at 26 bits
    assert abs(bench.executed_mac32(15, 24)[1] - pad24 - 6 * 7 * 72) < 1e-6  # 17 windows instead of 11
