"""CPU checks of the bench.py contract: the reference arm runs here (the reference's own compiled tracker from oracle/_ref when it was built, else the oracle port)
and prints one JSON line with the agreed keys; the committed default-run JSON of the B200 arm (newest profiles/r*_bench_default.json) carries every key the driver reads."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["higher_is_better"] is True
    import ref
    assert d["cpu_baseline"]["kind"] == ("reference" if ref.available() else "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["warmup"] == 1 and d["steps"] == 1 and {"sequences_per_gpu", "workload"} <= set(d["config"])
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_b200_line_has_every_contract_key():
    newest = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_default.json")))[-1]
    d = json.loads([l for l in open(newest) if l.startswith("{")][-1])
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline"} <= set(d)
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]) and d["roofline"]["bound"] == "hbm"
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]) and d["gpu_launches"] > 0 and d["vs_baseline"] is None
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and "workload" in d["config"] and d["scaling"] == "weak"
    if "keyframe_rate" in d:                                                 # the keyframe-rate leg ran and its device results were identical to the CPU path
        k = d["keyframe_rate"]; assert "error" not in k and k["identical_to_cpu"] is True and k["lidar_front_end"]["sweeps_per_s"] > 0 and k["make_new_traces"]["keyframes_per_s"] > 0
