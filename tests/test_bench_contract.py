"""bench.py's reference arm (the oracle port timed on host cores) and its JSON contract -- runs without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "1"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)


def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    # BASELINE.json's metric: "tiles/sec (256x256 b32) ..."
    assert d["impl"] == "reference" and d["metric"] == "tiles_per_sec_256x256_b32" and d["unit"] == "tiles/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_committed_bench_line_carries_the_contract_keys():
    """The bench line committed as evidence (profiles/r2s5_bench.json, produced by `python bench.py` on a B200) has every
    key of the driver's contract, with consistent values."""
    import json
    p = os.path.join(ROOT, "profiles", "r2s5_bench.json")
    d = json.loads(open(p).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "tiles_per_sec_256x256_b32" and d["unit"] == "tiles/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "l2" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - 32 * 1e3 / d["ms_per_step"]) <= 1e-6 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 32 * 256 * 256 * 3 and e["d2h_bytes_per_step"] == 32 * 256 * 256 * 4 and e["value"] > 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is not None and r["achieved"] * 1e12 * d["ms_per_step"] * 1e-3 <= 1.3541e12 * 1.001
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert d["gpu_launches"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    for k in ("single_stream", "parity", "fp32_mode", "slide", "getseg"):
        assert k in d, k
