"""bench.py prints ONE JSON line with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def run(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       timeout=timeout)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_contract():
    """--impl reference runs the reference's CPU decoder (or the C port) on the host cores"""
    d = run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "frames/s" and d["value"] > 0 and d["value"] == d["e2e"]["value"] == d["cpu_baseline"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["cpu_baseline"]["cores"] == os.cpu_count() and "workload" in d["config"]
    assert 0.1 < d["wer"] < 0.3


@pytest.mark.gpu
def test_gpu_arm_contract():
    d = run(["--steps", "3", "--warmup", "3", "--frames", "262144", "--e2e-frames", "65536", "--cpu-seconds", "1"])
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks", "fused_monte_carlo", "alu", "smem", "issue", "fixed_point"} <= set(d)
    assert d["gpu_launches"] == 3 and d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"] < 1 and d["roofline"]["peak"] > 1000
    assert d["e2e"]["h2d_bytes_per_step"] == 65536 * 63 * 4 and d["e2e"]["value"] > 0
    assert d["cpu_baseline"]["value"] > 0 and d["value"] > 1000 * d["cpu_baseline"]["value"]
    assert abs(d["wer"] - 0.18) < 0.01 and abs(d["fused_monte_carlo"]["wer"] - 0.18) < 0.01
    # round 2: the fixed-point decoder next to the headline, the host-side ceiling of the e2e path, and pipe counts that
    # are withheld (null) unless the committed ncu capture is stamped with the hash of the sources in use
    fx = d["fixed_point"]
    assert fx["value"] > 1.3 * d["value"] and fx["speedup_edge_iterations"] > 1.4 and 0.15 < fx["wer"] < 0.23
    assert 0 < d["e2e"]["frac_of_h2d_ceiling"] <= 1.05 and d["e2e"]["h2d_ceiling_gbs"] > 10
    import bench
    rec, current = bench.ncu_capture("K2 ms_cyclic BCH(63,36) NMS 4 dB resident")
    if rec is not None and current:
        assert 0 < d["smem"]["frac"] < 1 and 0 < d["issue"]["frac"] < 1 and d["roofline"]["traffic"] > 0
    else:
        assert d["smem"]["frac"] is None and d["issue"]["frac"] is None and d["roofline"]["traffic"] is None
    # the other BASELINE.json configurations ride in the same line (driver-visible): sum-product BCH(15,7), BCH(127,64) on
    # H / redundant H / multiple bases, RS(255,223) with its roofline, BCH(255,131)
    table = d["baseline_configs"]
    names = " | ".join(t["config"] for t in table)
    for key in ("configs[0] BCH(15,7)", "configs[2] BCH(127,64)", "redundant H", "8 bases", "configs[3] RS(255,223)", "configs[4] BCH(255,131)"):
        assert key in names, key
    assert all(t["value"] > 0 and t["n_gpus"] == 1 for t in table)
    rs = [t for t in table if "RS(255,223)" in t["config"]][0]
    assert rs["roofline"]["bytes_per_codeword"] == 511 and 0 < rs["roofline"]["frac"] < 1 and rs["e2e"]["value"] > 0
    assert rs["cpu_baseline"]["value"] > 0 and any("cpu_baseline" in t for t in table if "BCH(255,131)" in t["config"])
    ref = run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert ref["config"]["workload"] == d["config"]["workload"] and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
