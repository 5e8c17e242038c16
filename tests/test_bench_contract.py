"""CPU: the bench lines committed under profiles/ carry every key of the driver's contract (bench.py prints the same
dictionary), for both arms; and bench.py's reference arm / multi-GPU plumbing is importable without a GPU."""
import ast
import json

from conftest import REPO

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "gpu_launches", "roofline", "clocks", "e2e", "cpu_baseline"]


def _latest(prefix):
    files = sorted((REPO / "profiles").glob(f"r[0-9]*_{prefix}.json"))
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
    assert c["kind"] in ("port", "reference", "openblas") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
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
    # oracle/ is imported in exactly one function: the cpu_baseline leg
    importers = [f.name for f in ast.walk(tree) if isinstance(f, ast.FunctionDef)
                 for n in ast.walk(f) if isinstance(n, (ast.Import, ast.ImportFrom)) and ("oracle" in ast.dump(n) or "cpu_lapack" in ast.dump(n))]
    assert set(importers) == {"cpu_baseline"}
    # the repo's own library is imported in exactly one place: the constructor of the product's arm. The reference arm
    # (--impl reference) never reaches it, so its process maps oracle/_ref only.
    ours = [(c.name, f.name) for c in ast.walk(tree) if isinstance(c, ast.ClassDef) for f in ast.walk(c) if isinstance(f, ast.FunctionDef)
            for n in ast.walk(f) if isinstance(n, (ast.Import, ast.ImportFrom)) and "gputils_b200" in ast.dump(n)]
    assert ours == [("OursArm", "__init__")]
    everywhere = [n for n in ast.walk(tree) if isinstance(n, (ast.Import, ast.ImportFrom)) and "gputils_b200" in ast.dump(n)]
    assert len(everywhere) == 1


def test_torch_input_generator_is_the_counter_based_generator_of_the_survey(oracle):
    """bench.py builds the inputs of BOTH arms with plain torch; it must be the same generator as gpub_fill_* / oracle_np."""
    import importlib.util
    import numpy as np
    import torch
    spec = importlib.util.spec_from_file_location("bench_module", REPO / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    x = bench.gen_uniform(torch.empty(10_001, dtype=torch.float64), -1.0, 1.0, 0x5EED0001, chunk=4096)
    assert np.array_equal(x.numpy(), oracle.fill_uniform(10_001, -1.0, 1.0, 0x5EED0001))
    xf = bench.gen_uniform(torch.empty(777, dtype=torch.float32), -1.0, 1.0, 0x5EED0103)
    assert np.array_equal(xf.numpy(), oracle.fill_uniform(777, -1.0, 1.0, 0x5EED0103, np.float32))
    A = bench.gen_spd(torch.empty((5, 8, 8), dtype=torch.float64), 8.0, 99)
    ref = oracle.fill_spd_batched(8, 5, 8.0, 99)
    assert np.abs(A.numpy().transpose(0, 2, 1) - ref).max() <= 1e-14
