"""bench.py's output contract (CPU): the reference arm is run here on a tiny sample and its JSON line checked key by
key; the most recent committed record of the CUDA arm (profiles/, written on a B200) is held to the same schema, to
internal consistency (roofline.frac = achieved / peak, e2e carries real copy sizes, launches counted) and to
BASELINE.json's metric."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches")


def _check_base(d):
    for k in BASE_KEYS:
        assert k in d, k
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"].split(" at ")[0] in base["metric"]          # "rays/sec (64 coarse + 128 fine samples)"
    assert d["unit"] == "rays/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None                                 # BASELINE.md publishes no number for this metric
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-rays", "256"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                                    # exactly ONE JSON line on stdout
    d = json.loads(lines[0])
    _check_base(d)
    assert d["impl"] == "reference" and d["dtype"] == "f32" and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "256" in cb["sample"]
    # the unmodified reference is what is timed wherever it is reachable (/root/reference here, oracle/_ref on the GPU box)
    import refimport
    assert cb["kind"] == ("reference" if refimport.available() else "port")
    # same config as the CUDA arm (the bounded sample is described separately)
    assert d["config"]["rays_per_step"] == 640000 and d["rays_timed_per_step"] == 256


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--cpu-rays", "256"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_latest_committed_cuda_record():
    recs = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9][0-9]*_bench_bf16.json")))
    assert recs, "no committed bench record"
    d = json.load(open(recs[-1]))
    _check_base(d)
    assert d["dtype"] == "bf16" and d["n_gpus"] == 1 and d["warmup"] >= 3
    assert d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 640000 * 6 * 4 and d["e2e"]["d2h_bytes_per_step"] == 640000 * 5 * 4
    assert 0.5 * d["value"] < d["e2e"]["value"] <= 1.05 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["frac"] < 1 and r["traffic"] is not None
    # achieved = algorithmic FLOPs per launch / average launch duration
    assert abs(r["achieved"] - r["algorithmic_flop_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e12) < 1e-6 * r["achieved"]
    # value is whole-job rays / time, and the dominant kernel's algorithmic work matches 256 evaluations per ray
    assert abs(d["value"] - d["config"]["rays_per_step"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert abs(r["algorithmic_flop_per_launch"] * r["launches"] / d["steps"] - 640000 * 256 * 1186816) < 1.0
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
