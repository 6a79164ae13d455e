"""bench.py contract, CPU side: the reference arm (`--impl reference`) runs on host cores
only and must print ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_json_line():
    out = _run("--impl", "reference", "--workload", "small", "--steps", "2", "--warmup", "1")
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cell_gene_mi_evals_per_s_per_greedy_step"
    assert d["unit"] == "cell*gene/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1
    assert cb["value"] == d["value"] and cb["sample"]
    assert d["port_all_cores"]["cores"] >= 1 and d["port_all_cores"]["value"] > 0
    assert set(d["config"]) == {"workload", "selector", "bins", "k", "l2"}
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    out = _run("--impl", "reference", "--workload", "small", "--steps", "1", "--warmup", "0",
               "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.strip() == ""
