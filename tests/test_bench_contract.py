"""bench.py's reference arm (the unmodified reference on the host cores) needs no GPU: run it on the test-sized workload
and check the JSON line the driver parses."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_driver is built only where /root/reference exists")
def test_reference_arm_prints_the_contract_line(tmp_path):
    env = dict(os.environ, KMX_BENCH_CACHE=str(tmp_path), CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([x for x in r.stdout.splitlines() if x.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "kmers_encoded_per_s" and line["unit"] == "k-mers/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["config"]["workload"] and line["config"]["k"] == 31 and line["gpu_launches"] == 0
    assert line["query"]["value"] > 0


def test_unavailable_reference_is_reported_not_raised(tmp_path, monkeypatch):
    """without the compiled reference the arm must print {"impl": "reference", "unavailable": ...} and exit 0"""
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))          # no oracle/_ref under this root
    import io
    import contextlib
    buf = io.StringIO()

    class A:
        workload, steps, warmup, gpus = "small", 1, 0, 1
    with contextlib.redirect_stdout(buf):
        bench.run_reference(A)
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" in line
