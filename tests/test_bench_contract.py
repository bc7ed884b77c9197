"""bench.py's reference arm runs on the CPU: its one JSON line must carry the keys of the bench contract (the GPU arm builds
its line from the same fields; its shape is checked statically)."""
import ast
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "mpm_particle_substeps_per_sec" and d["unit"] == "particle-substeps/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_gpu_arm_line_has_the_contract_keys():
    """make_line(e2e, cpu) in bench.py: the keyword names of the dict it returns (static: the GPU arm cannot run here)"""
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    fn = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "make_line"][0]
    call = [n for n in ast.walk(fn) if isinstance(n, ast.Return)][0].value
    keys = {k.arg for k in call.keywords}
    assert BASE_KEYS | {"roofline", "clocks"} <= keys, (BASE_KEYS | {"roofline", "clocks"}) - keys
    src = open(os.path.join(ROOT, "bench.py")).read()
    for k in ("bound=", "achieved=", "peak=", "frac=", "traffic="):       # the roofline object
        assert k in src
