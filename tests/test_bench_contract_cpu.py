"""CPU checks of bench.py's contract: the reference arm (the oracle port of the reference's step on the host cores) prints ONE
JSON line with the keys the driver reads, non-zero ranks of a torchrun launch stay silent, and the committed per-launch DRAM
traffic table carries the fields bench.py attaches to its roofline entries."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *flags):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--batch", "8", *flags], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_line_with_the_contract_keys():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "dcgan_train_step_images_per_sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_rank_zero_only():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2") == []


def test_traffic_table_matches_what_bench_reads():
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
        db = json.load(f)
    assert db, "profiles/r02_traffic.json is empty"
    for key, row in db.items():
        assert "[" in key and row["dram_bytes_per_launch"] > 0 and "ncu" in row["source"]
