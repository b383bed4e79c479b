"""CPU: the reference arm of bench.py (`--impl reference`: the reference algorithm, i.e. the oracle port, on the host cores)
prints exactly ONE JSON line on stdout with the keys the driver reads; run on a small workload so it takes seconds.
The B200 arm needs a GPU and is exercised on the GPU box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--encoder", "vits", "--size", "70",
                           "--steps", "2", "--warmup", "1", *args], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("images/sec at 518x518 ViT-L fwd+SSI/HDN loss")
    assert d["steps"] == 2 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == dict(value=d["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d["gpu_launches"] == 0 and "unavailable" not in d


def test_reference_arm_runs_on_rank_zero_only():
    r = _run(dict(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])
