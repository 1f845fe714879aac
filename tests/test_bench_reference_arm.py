"""bench.py's reference arm (--impl reference) runs without a GPU: the numpy restatement of the reference's CPU path
on persistent worker processes.  This pins the JSON contract of that line and that ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "2", "--warmup", "1", "--ref-tasks", "2"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip()


def test_reference_arm_line():
    line = json.loads(_run().splitlines()[-1])
    assert line["impl"] == "reference"
    assert line["unit"] == "crops/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 1
    assert line["config"]["workload"].startswith("cfg1")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # crops per step / time per step is the value: 2 tasks per core and step, a task = one 2-frame shard of cfg1
    crops = (os.cpu_count() or 1) * 2 * 2
    assert cb["cores"] == (os.cpu_count() or 1)
    assert abs(line["value"] - crops / (line["ms_per_step"] * 1e-3)) / line["value"] < 1e-6
    # the two arms name the workload with the same keys (the driver compares the config dicts)
    sys.path.insert(0, ROOT)
    import bench
    from loans_b200 import workloads as W
    assert set(line["config"]) == set(bench.workload_config(W.WORKLOADS["cfg2"], True))


def test_reference_arm_other_ranks_print_nothing():
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == ""
