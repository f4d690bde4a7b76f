"""CPU: the reference arm of bench.py (`--impl reference`: the unmodified reference binary on the host cores) prints the JSON line the driver
expects -- metric / unit of BASELINE.json, `impl`, a `cpu_baseline` of kind "reference" whose value is the line's, an `e2e` object with zero
copy bytes -- and the product arm refuses to run without a GPU instead of falling back to anything on the CPU."""
import json
import os
import subprocess
import sys

import pytest

from conftest import REF_DIR, REPO


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "wtzmo")), reason="reference binary not built")
def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    base = json.load(open(os.path.join(REPO, "BASELINE.json")))
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["unit"] == "Gbp/s" and line["metric"].startswith("aligned Gbp/sec")
    assert str(base.get("metric", "")).lower().split()[0] in line["metric"].lower() or "gbp" in str(base.get("metric", "")).lower()
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["unit"] == line["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--workload", "cfg1", "--steps", "1", "--warmup", "0", "--no-sub", "--no-cpu-baseline"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode != 0                      # no CPU fallback: the product path fails loudly
    assert "reference" not in r.stdout            # and prints no line that could be mistaken for a measurement
