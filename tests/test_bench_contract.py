"""CPU: the reference arm of bench.py (`--impl reference`, the oracle on the host cores) prints ONE JSON
line with the keys the driver reads, and the committed engine-arm lines (profiles/r2_bench_c*.json, produced
on a B200) carry the roofline / cpu_baseline / e2e objects of the measurement contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["unit"] == "graphs/s" and "workload" in d["config"]
    cb, e2e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert e2e == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_committed_engine_line_has_the_measurement_objects():
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_c1.json")))
    assert BASE_KEYS <= set(d) and d["metric"] == "graphs_per_sec_cgcnn_train_step"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    assert r["fwd_frac"] == r["frac"] and 0 < r["bwd_frac"] < 1          # forward AND backward fractions at top level
    sf = r["smear_fused"]                                                # the fused form in both SURVEY 8(d) accountings
    assert sf["fwd"]["frac_fused_form"] < sf["fwd"]["frac_operator_surface_form"] < 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert e["distances_shipped"]["h2d_bytes_per_step"] < e["h2d_bytes_per_step"]
    assert d["gpu_launches"] > 0 and "clocks" in d and "sm_mhz" in d["clocks"]
    assert d["store_step"]["value"] > 0


def test_committed_lines_cover_every_baseline_config():
    for k, model in ((2, "schnet"), (3, "megnet"), (4, "mpnn")):
        d = json.load(open(os.path.join(ROOT, "profiles", f"r2_bench_c{k}.json")))
        assert BASE_KEYS <= set(d) and d["metric"] == f"graphs_per_sec_{model}_train_step"
        assert d["value"] > 0 and d["roofline"]["frac"] > 0 and f"configs[{k}]" in d["config"]["workload"]
        assert d["e2e"]["value"] > 0 and d["store_step"]["value"] > 0
