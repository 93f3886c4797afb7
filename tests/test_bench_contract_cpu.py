"""bench.py's reference arm (`--impl reference`: the reference's own CPU implementation of the path, no GPU, none of this repo's
kernels) prints the driver's JSON contract: one line, the headline metric / unit, a positive value, the cpu_baseline and e2e
objects. Runs one bounded step of the real reference modules (oracle/_ref) - or of the oracle port when that directory is absent."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, MAS_CPU_ARM_SECONDS="5")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("VQ-IMG 256^2 images/sec") and d["value"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # under torchrun only rank 0 works: any other rank exits 0 without a line
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                        timeout=120, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), cwd=ROOT)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
