"""The reference arm of bench.py (`--impl reference`: the reference's CPU CausalLM path restated, oracle/causal_lm.py) on the
tiny workload: one JSON line with the contract's keys; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny-test", "--steps", "2",
                          "--warmup", "3"], capture_output=True, text=True, env=env, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_the_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decode_tokens_per_s" and d["unit"] == "tokens/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] >= 3 and d["value"] > 0
    assert d["config"]["workload"] == "tiny-test"
    # the line is self-consistent: value == batch / extrapolated full-depth step time, and it says how many steps really ran
    assert abs(d["value"] - d["config"]["batch"] / (d["ms_per_step"] / 1e3)) <= 1e-6 * d["value"]
    assert d["config"]["steps_timed_per_sample"] >= 1 and d["config"]["extrapolated_in_depth"] is True
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}) == []
