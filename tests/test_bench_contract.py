"""The bench.py contract on CPU: the reference arm (`--impl reference`, the reference algorithm on the host cores through the
oracle port) prints exactly ONE JSON line on stdout with the keys the driver parses, for both the DDPM and the VQ-VAE line, and
ranks other than 0 of a multi-rank launch exit 0 without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, env=e,
                       timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "ddpm_unet_train_steps_per_sec" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "configs[1]" in d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_other_ranks_exit_without_output():
    lines = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                 env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"})
    assert lines == []
