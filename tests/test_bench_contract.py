"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, and the
product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "lepton")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/lepton not built (needs /root/reference)")
def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "8", "--distinct", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "JPEG MB/s encode+decode" and line["unit"] == "MB/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]
    assert line["encode"]["value"] > line["value"] and line["decode"]["value"] > line["value"]      # value = bytes / (encode s + decode s)


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--images", "2"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
