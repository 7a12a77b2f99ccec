"""The C++ drop-in layer (include/cc/codes.h, include/cc/simulation.h) and the reference-style CLI
(tools/benchmark.cc), built by channelcoding_b200.build.build_tools()."""
import os
import subprocess

import pytest

from channelcoding_b200 import build as _build


@pytest.fixture(scope="module")
def bins():
    _build.build()
    outs = _build.build_tools()
    return {os.path.basename(p): p for p in outs}


def test_host_part(bins):
    r = subprocess.run([bins["host_layer_test"], "--host"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


def test_cli_usage_without_selection(bins):
    """an empty selection prints the usage like benchmark.c++:434-437 -- but first the catalogue has to be
    constructed, which needs the GPU: on a CPU-only box the program must fail loudly, not fall back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([bins["benchmark"], "--k", "5"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "no usable CUDA device" in (r.stdout + r.stderr)


@pytest.mark.gpu
def test_gpu_part(bins, tmp_path):
    r = subprocess.run([bins["host_layer_test"], "--gpu", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cli_bitflip_and_awgn(bins, tmp_path, kat):
    """benchmark --simulation bitflip reproduces Table 3 counts; --simulation awgn writes the sweep log"""
    r = subprocess.run([bins["benchmark"], "--simulation", "bitflip", "--k", "5", "--dmin", "7", "--algorithm", "ms",
                        "--algorithm", "scms2", "--algorithm", "bm", "--errors", "3", "--out", str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    for tag, key in (("MS", "MS"), ("SCMS2", "SCMS2"), ("BM", "BM")):
        lines = open(tmp_path / ("(31, 16, 7)-%s.log" % tag)).read().splitlines()
        assert lines[0] == " errors                   wer"
        for w in range(4):
            row = kat["bitflip_31_16_7"][str(w)]
            assert abs(float(lines[1 + w].split()[1]) - row[key] / row["patterns"]) < 1e-12
    r = subprocess.run([bins["benchmark"], "--k", "6", "--dmin", "7", "--algorithm", "nms", "--max-samples", "100000",
                        "--seed", "3", "--out", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = open(tmp_path / "(63, 45, 7)-NMS.log").read().splitlines()
    assert lines[0] == "   ebno                   wer" and len(lines) >= 10
    wers = [float(x.split()[1]) for x in lines[1:]]
    assert wers[0] > 0.3 and wers[-1] < 1e-3


@pytest.mark.gpu
def test_python_sweep_equals_cli_sweep(bins, tmp_path):
    """the same sweep driven from C++ (cc::awgn_simulation via the CLI, three pool threads sharing one
    context) and from Python (channelcoding_b200.simulation, the multi-GPU driver with world size 1) must
    write byte-identical "<name>.log" files: same schedule, same Philox frames, same counters"""
    import channelcoding_b200 as cc
    from channelcoding_b200 import simulation
    cli_dir = tmp_path / "cli"
    py_dir = tmp_path / "py"
    cli_dir.mkdir()
    py_dir.mkdir()
    r = subprocess.run([bins["benchmark"], "--k", "5", "--dmin", "7", "--algorithm", "ms", "--algorithm", "nms",
                        "--algorithm", "scms2", "--max-samples", "200000", "--seed", "11", "--threads", "3", "--out",
                        str(cli_dir)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    ctx = cc.Context(0)
    code = ctx.bch(5, dmin=7)
    for tag, variant, alpha in (("MS", "MS", 1.0), ("NMS", "NMS", 0.8), ("SCMS2", "SCMS2", 1.0)):
        name = code.to_string(tag)
        res = simulation.awgn_sweep(simulation.gpu_point_fn(code, variant, alpha, seed=11), name, code.rate, cap=200000,
                                    log_dir=str(py_dir))
        assert open(py_dir / (name + ".log")).read() == open(cli_dir / (name + ".log")).read(), tag
        assert res[0]["ebno"] == 1.0 and res[0]["frames"] == 10000 and res[-1]["wer"] < res[0]["wer"]
    ctx.close()


def test_uncoded_cli_without_gpu_is_loud(bins, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([bins["uncoded"], "--l", "63", "--out", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no usable CUDA device" in (r.stdout + r.stderr)


@pytest.mark.gpu
def test_uncoded_program(bins, tmp_path):
    """simulation/uncoded.c++ on the GPU: "<l>-uncoded.log" with the sweep schedule of rate 0.5 and the word error
    rate of l uncoded BPSK symbols, 1 - (1 - Q(1/sigma))^l"""
    import math
    r = subprocess.run([bins["uncoded"], "--l", "63", "--seed", "5", "--max-samples", "400000", "--out", str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = open(tmp_path / "63-uncoded.log").read().splitlines()
    assert lines[0] == "   ebno                   wer"
    pts = [tuple(map(float, x.split())) for x in lines[1:]]
    assert pts[0][0] == 1.0 and len(pts) >= 10   # one step above the rounded-down limit of rate 0.5 (0.188 dB)
    for eb, wer in pts:
        sigma = 1.0 / math.sqrt(2 * 0.5 * 10 ** (eb / 10))
        q = 0.5 * math.erfc(1.0 / (sigma * math.sqrt(2)))
        expect = 1 - (1 - q) ** 63
        assert abs(wer - expect) < 5 * math.sqrt(expect * (1 - expect) / 1e4) + 1e-4, (eb, wer, expect)


@pytest.mark.gpu
def test_bitflips_program(bins, tmp_path, kat):
    """simulation/bitflips.c++ on the GPU: nine decoders of BCH(31,16,7), every pattern of 0..3 flipped bits, each log
    equals the counts the reference produced (tests/golden/kat.json, Table 3 of the report)"""
    r = subprocess.run([bins["bitflips"], "--errors", "3", "--out", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.splitlines()[:9] == ["(31, 16, 7)-" + t for t in ("BM", "PGZ", "EUKLID", "MS", "NMS", "OMS", "SCMS1", "SCMS2", "2DNMS")]
    for tag in ("BM", "PGZ", "EUKLID", "MS", "NMS", "OMS", "SCMS1", "SCMS2", "2DNMS"):
        lines = open(tmp_path / ("(31, 16, 7)-%s.log" % tag)).read().splitlines()
        assert lines[0] == " errors                   wer" and len(lines) == 5
        for w in range(4):
            row = kat["bitflip_31_16_7"][str(w)]
            assert abs(float(lines[1 + w].split()[1]) - row[tag] / row["patterns"]) < 1e-12, (tag, w)


@pytest.mark.gpu
def test_device_group_on_one_gpu(bins, tmp_path):
    """ccgpu_group / cc::device_group with three members that all sit on device 0: sharding by global frame index and
    the in-library counter merge give exactly the one-device counters, outputs and sweep log"""
    r = subprocess.run([bins["host_layer_test"], "--group", str(tmp_path), "0", "0", "0"], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_device_group_on_all_gpus(bins, tmp_path):
    """the same on every GPU of the box (1 vs N devices give identical counters through the C++ layer)"""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU only")
    r = subprocess.run([bins["host_layer_test"], "--group", str(tmp_path)] + [str(i) for i in range(n)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cli_gpus_device_and_stop_rule(bins, tmp_path):
    """--gpus N shards every point over N devices and must not change a byte of the log; --device picks the first
    device; --stop-rule gf2 changes the decoders' stop test (benchmark.cc used to parse and drop both)"""
    import torch
    n = min(2, torch.cuda.device_count())
    common = ["--k", "5", "--dmin", "7", "--algorithm", "nms", "--max-samples", "200000", "--seed", "5"]
    logs = {}
    for tag, extra in (("one", []), ("gpus", ["--gpus", str(max(n, 1)), "--device", "0"]), ("gf2", ["--stop-rule", "gf2"])):
        d = tmp_path / tag
        d.mkdir()
        r = subprocess.run([bins["benchmark"]] + common + extra + ["--out", str(d)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        logs[tag] = open(d / "(31, 16, 7)-NMS.log").read()
    assert logs["one"] == logs["gpus"]
    assert logs["one"] != logs["gf2"]   # the GF(2) syndrome accepts non-zero codewords: different WER / iteration statistics
    r = subprocess.run([bins["benchmark"]] + common + ["--device", "99", "--out", str(tmp_path)], capture_output=True, text=True,
                       timeout=120)
    assert r.returncode != 0 or "ccgpu_create failed" in (r.stdout + r.stderr)
    r = subprocess.run([bins["benchmark"]] + common + ["--stop-rule", "bogus"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
