"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the reference's own CPU
implementation of the path on the host cores and prints ONE JSON line with the keys the driver reads; under a
multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--quick",
                           "--sample-n", "256", "--steps", "1", "--warmup", "0", "--cores", "2"] + list(args),
                          capture_output=True, text=True, timeout=600, env=env)


def test_reference_arm_prints_one_contract_line():
    out = run({})
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["metric"] == "kkt_factor_solve_fp64_tflops" and d["unit"] == "TFLOP/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 2
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"] > 0.0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["details"]["residual"] < 1e-9
    assert "workload" in d["config"] and "model" not in d["config"]
    # the config is static (what the GPU arm prints too: the driver compares them); run-dependent numbers are elsewhere
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert d["config"] == b.static_config(512, 256, 512, 1)


def test_reference_arm_other_ranks_stay_silent():
    out = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert out.returncode == 0, out.stderr
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]
