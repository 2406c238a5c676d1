"""The committed bench lines (profiles/) carry the keys the measurement contract names: one JSON line per run with metric / value / unit, the
end-to-end number with its copied bytes, the roofline of the dominant kernel, the CPU baseline, the clock samples and the launch count."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
        "e2e", "gpu_launches", "roofline", "clocks"]


def _line(name):
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        pytest.skip(name + " not committed")
    return json.loads(open(p).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["bench_r02s.json", "bench_r02aa.json", "bench_n2_r02z.json", "bench_n4_r02z.json", "bench_n8_r02z.json"])
def test_committed_bench_lines_follow_the_contract(name):
    d = _line(name)
    for k in BASE:
        assert k in d, (name, k)
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], (name, k)
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]            # copies inside the timed region: never the resident number again
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, (name, k)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["gpu_launches"] > 0 and d["n_gpus"] in (1, 2, 4, 8)
    assert d["scaling"] == ("weak" if d["n_gpus"] == 1 else "strong")
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # whole-job throughput: nominal nodes x reads over the step time
    nodes_reads = d["config"]["n_nodes"] * d["config"]["n_reads"]
    assert abs(d["value"] - nodes_reads / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6


def test_full_line_has_cpu_baseline_and_reference_arm():
    d = _line("bench_r02s.json")
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["gpu_same_sample"]["best_nodes_tie_lists_scores_agree"] is True
    ref = _line("bench_ref_r02z.json")
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"]
    assert d["e2e"]["value"] / ref["e2e"]["value"] > 100          # the headline: end to end against the reference's own CPU placement
