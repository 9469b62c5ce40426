"""bench.py's CPU-runnable contract: the reference arm (`--impl reference`) prints one JSON line with the keys the driver reads,
runs on rank 0 only under torchrun, and ignores the OMP_NUM_THREADS=1 that torchrun exports to its workers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1",
                           "--warmup", "0", "--cpu-sample-rays", "2e6"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_line_and_rank_gating():
    r = _run({"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "exchange_factor_rays_per_sec" and line["unit"] == "rays/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    n_cores = len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["cores"] == n_cores                      # not the single thread OMP_NUM_THREADS=1 asks for
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("cfg1")
    # under torchrun only rank 0 works and prints
    r1 = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r1.returncode == 0 and r1.stdout.strip() == ""
