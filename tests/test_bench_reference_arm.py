"""The reference arm of bench.py (the reference's Analyze loop on host cores) runs without a GPU and
prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "720p",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frame-pairs/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "720p" in line["metric"] and line["config"]["width"] == 1280


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_proportional_shares_keep_the_work_of_the_job():
    """bench.py's e2e leg at N > 1 sizes every rank's share of a step by its measured frame rate
    (DESIGN.md section 8): equal rates give equal shares, the total stays N * frames_per_step up to
    rounding, faster ranks get more, nobody drops below the floor."""
    import bench
    assert bench.proportional_shares([5.0, 5.0], 32) == [32, 32]
    got = bench.proportional_shares([23.0, 23.1, 23.0, 23.1, 35.0, 35.1, 34.9, 35.2], 32)
    assert abs(sum(got) - 8 * 32) <= 4 and got[0] < got[4] and min(got) >= 8
    assert got == sorted(got) or got[:4] == sorted(got[:4])
    assert bench.proportional_shares([1.0, 100.0], 32)[0] == 8
    assert bench.proportional_shares([0.0, 0.0], 32) == [32, 32]
