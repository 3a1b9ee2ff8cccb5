"""CPU test of bench.py's reference arm (`--impl reference`): it runs the reference's own CPU code
(oracle/_ref/ref_driver, or the oracle port where the driver is absent) and must print ONE JSON line
with the contract's keys. The GPU arm needs a B200 and is exercised by the driver."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--workload", "ascii"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["config"]["name"] == "ascii" and "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--workload", "ascii"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_pure_python_helpers_do_not_load_the_product_library():
    """The reference arm imports ray_tracying_b200.workloads / .scenes: that must not map librt_b200.so
    (the driver records which native libraries each arm loaded)."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import ray_tracying_b200.workloads, ray_tracying_b200.scenes, ray_tracying_b200.dist\n"
            "assert 'ray_tracying_b200._lib' not in sys.modules and 'ray_tracying_b200.api' not in sys.modules\n"
            "assert 'librt_b200' not in open('/proc/self/maps').read()\n"
            "import ray_tracying_b200 as rt; rt.make_params\n"
            "assert 'librt_b200' in open('/proc/self/maps').read()\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]


def test_reference_arm_sample_is_a_window_of_the_named_workload():
    """The reference arm of a heavy workload renders windows (rows x columns) sized from a ray-count probe."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "mixed100k"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900,
                       env=dict(os.environ, RT_BENCH_CACHE="/tmp/rt_b200_bench"))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][-1])
    assert d["config"]["name"] == "mixed100k" and "windows of" in d["config"]["sample"]
    assert d["rays_per_step"] > 100000 and 0.05 < d["value"] < 1000
