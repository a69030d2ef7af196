"""bench.py contract checks that run without a GPU: the reference arm (`--impl reference`, the oracle on the host cores,
a bounded sample) prints ONE JSON line with the agreed keys; the sm_100a arm's line carries every key the driver reads;
and the product fails loudly when libttvdm_sm100.so is missing."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--height", "64", "--width", "64"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_sm100_arm_line_has_every_contract_key():
    src = (ROOT / "bench.py").read_text()
    body = src[src.index("        line = {"):src.index("        print(json.dumps(line), flush=True)")]
    for key in ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"]:
        assert f'"{key}"' in body, key
    for key in ["bound", "achieved", "peak", "unit", "frac", "traffic"]:
        assert f'"{key}"' in src, key
    for key in ["h2d_bytes_per_step", "d2h_bytes_per_step", "sm_mhz", "sm_max_mhz", "reasons"]:
        assert key in src, key
    assert 'add_argument("--gpus"' in src and "RANK" in src and "LOCAL_RANK" in src and "WORLD_SIZE" in src


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from this_and_that_vdm_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", tmp_path / "libttvdm_sm100.so")
    with pytest.raises(lib.TtvdmError, match="no CPU fallback"):
        lib.load()
    with pytest.raises(lib.TtvdmError):
        lib.launch_count()
