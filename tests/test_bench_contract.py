"""CPU: the bench lines committed under profiles/ carry every key of the driver's contract (bench.py prints the same
dictionary), for both arms; and bench.py's reference arm / multi-GPU plumbing is importable without a GPU."""
import ast
import json

from conftest import REPO

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "gpu_launches", "roofline", "clocks", "e2e", "cpu_baseline"]


def _latest(prefix):
    files = sorted((REPO / "profiles").glob(f"r1*_{prefix}.json"))
    assert files, f"no profiles/*_{prefix}.json"
    return json.loads(files[-1].read_text())


def test_our_bench_line_has_the_contract_keys():
    d = _latest("bench_ours")
    for k in REQUIRED:
        assert k in d, k
    assert d["unit"] == "matrices/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64"
    assert d["warmup"] >= 3 and d["gpu_launches"] == 2 * d["steps"]
    assert "workload" in d["config"] and "BASELINE config 2" in d["config"]["workload"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    e = d["e2e"]
    assert e["unit"] == "matrices/s" and e["h2d_bytes_per_step"] > 8e9 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["clocks"]["samples"] >= 1 and "reasons" in d["clocks"]
    # value = matrices per GPU * GPUs / time
    assert abs(d["value"] - d["n_gpus"] * d["config"]["matrices_per_gpu"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6


def test_reference_bench_line_is_marked_and_comparable():
    d = _latest("bench_reference")
    assert d["impl"] == "reference" and d["metric"] == _latest("bench_ours")["metric"] and d["unit"] == "matrices/s"
    assert d["config"]["workload"] == _latest("bench_ours")["config"]["workload"]
    assert "e2e" in d and "cpu_baseline" in d


def test_bench_source_parses_and_keeps_the_oracle_out_of_the_timed_path():
    src = (REPO / "bench.py").read_text()
    tree = ast.parse(src)
    # the oracle is imported in exactly one function: the cpu_baseline leg
    importers = [f.name for f in ast.walk(tree) if isinstance(f, ast.FunctionDef)
                 for n in ast.walk(f) if isinstance(n, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(n)]
    assert set(importers) == {"cpu_baseline"}
