"""CPU: `bench.py --impl reference` (the reference algorithm on the host cores) prints ONE JSON line with the
contract's keys; ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
  env = dict(os.environ)
  env.update(env_extra)
  r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                      "--warmup", "1", "--cpu-pairs", "2"], capture_output=True, text=True, env=env, timeout=600)
  assert r.returncode == 0, r.stderr[-2000:]
  return r.stdout.strip()


def test_reference_arm_json_line():
  out = _run({})
  lines = [l for l in out.splitlines() if l.strip()]
  assert len(lines) == 1, out
  d = json.loads(lines[0])
  for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
    assert key in d, key
  assert d["impl"] == "reference" and d["unit"] == "img-pairs/s" and d["value"] > 0
  assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
  assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
  assert "workload" in d["config"]


def test_reference_arm_other_ranks_are_silent():
  assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == ""
