"""The reference's OWN classes on top of libccgpu.so: oracle/_ref/integration_test is the reference's
decoder / bitflip_simulation / AWGN loop (compiled from /root/reference by oracle/build_ref.sh) using the
adapter of INTEGRATION.md.  Skipped where oracle/_ref was not built."""
import os
import subprocess

import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "oracle", "_ref", "integration_test")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/integration_test not built")
def test_reference_classes_drive_the_engine(tmp_path, kat):
    r = subprocess.run([BIN, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l and not l.startswith("Using")]
    table = [l for l in lines if len(l.split()) == 2 and l.split()[0].isdigit()]  # "<errors> <wer>" log lines
    for w in range(4):  # Table 3 through the reference's own bitflip_simulation
        row = kat["bitflip_31_16_7"][str(w)]
        assert abs(float(table[w].split()[1]) - row["MS"] / row["patterns"]) < 1e-12
    awgn = [l for l in lines if l.startswith("awgn")][0].split()
    assert 0.55 < int(awgn[1]) / int(awgn[2]) < 0.85  # sigma 0.75 is Eb/N0 = 1.9 dB for (63,36): WER about 0.7
