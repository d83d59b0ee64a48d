"""bench.py on the CPU tier: it must at least compile, parse its arguments, and its reference arm (the oracle's C port on
host cores -- no GPU needed) must print ONE JSON line with the keys the driver reads, on the same `config.workload`
string the GPU arm prints."""
import json
import os
import py_compile
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_and_tools_compile():
    for rel in ("bench.py", "__graft_entry__.py", "tools/cli_wall.py", "tools/sweep_variants.py", "tools/update_profiles.py",
                "tools/ncu_summary.py", "tools/rule_probe.py", "tools/few_genes_probe.py", "tools/time_python_reference.py"):
        py_compile.compile(os.path.join(ROOT, rel), doraise=True)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert res.returncode == 0 and "--workload" in res.stdout and "--split" in res.stdout


def test_reference_arm_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c3",
                          "--genes", "2000", "--isolates", "300", "--perms", "40", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-1000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.workload_string("c3", 2000, 300, 1, 40)
