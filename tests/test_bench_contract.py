"""bench.py prints ONE JSON line with the contract's keys (CPU: --impl reference; GPU: our arm on config C1)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                      # exactly one line on stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--config", "C1", "--steps", "2", "--warmup", "1"], 300)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "apsp_build_seconds" and d["unit"] == "s" and d["higher_is_better"] is False
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["value"] > 0


@pytest.mark.gpu
def test_our_arm_line_on_c1():
    d = _run(["--config", "C1", "--steps", "3", "--warmup", "3"], 600)
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks"} <= set(d)
    assert d["metric"] == "apsp_build_seconds" and d["unit"] == "s" and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 * d["value"]) < 1e-6 and d["gpu_launches"] > 0
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["fw"]["same_matrix_as_headline"] is True                       # SSSP and Floyd-Warshall agree
    assert d["insertion"]["value"] > 0 and d["insertion"]["roofline"]["bound"] == "hbm"


@pytest.mark.parametrize("fname", ["r02_bench_c3_n1.json", "r02_bench_c3_n2.json", "r02_bench_c3_n4.json", "r02_bench_c3_n8.json",
                                   "r01_bench_c3_n1.json", "r01_bench_c3_n2.json", "r01_bench_c3_n4.json", "r01_bench_c3_n8.json"])
def test_committed_bench_lines_keep_the_contract(fname):
    """The bench lines committed under profiles/ (measured on B200s) carry the contract's keys and are self-consistent."""
    d = json.load(open(os.path.join(ROOT, "profiles", fname)))
    assert BASE_KEYS - {"cpu_baseline"} | {"roofline", "gpu_launches", "clocks"} <= set(d)
    assert d["metric"] == "apsp_build_seconds" and d["unit"] == "s" and d["higher_is_better"] is False and d["vs_baseline"] is None
    assert d["n_gpus"] == int(fname.split("_n")[1].split(".")[0]) and d["warmup"] >= 3
    assert abs(d["ms_per_step"] - 1e3 * d["value"]) < 1e-6 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor", "alu") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] >= 0.95 * d["clocks"]["sm_max_mhz"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] >= d["value"]
    assert d["fw"]["same_matrix_as_headline"] is True and d["checksum"] == 12082473595208.0
    if d["n_gpus"] == 1:
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and r["traffic"] > 0
    if fname.startswith("r02"):
        # round 2: config C5 at every N, sources sharded; at N = 1 the query latencies, the sequential queue and the CPU legs
        assert d["sssp_c5"]["n_gpus"] == d["n_gpus"] and d["sssp_c5"]["matches_scipy_dijkstra"] is True
        if d["n_gpus"] == 1:
            assert d["query_latency"]["device_allocations_during_the_calls"] == 0
            assert d["insertion"]["roofline"]["frac"] >= 0.6
            q = d["insertion"]["sequential_queue"]
            assert q["with_cld_side_effects"]["assigned"] == q["winners_only"]["assigned"] > 0
            assert d["sim"]["cpu_baseline"]["same_objective_on_every_compared_tick"] is True
