"""CPU: the bench.py contract pieces that do not need a GPU -- the `--impl reference` arm (oracle port on the host cores)
prints one JSON line with the required keys, and non-zero ranks of a torchrun launch of that arm exit without work."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3", "--rays", "256"],
                         capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["warmup"] >= 3
    assert line["config"]["workload"].startswith("delaunay45k_302ktet/256rays/tetra-nerf(128+128,biased,M=512)/eval-forward")
    assert line["config"]["rays_per_step_per_gpu"] == 256 and set(line["config"]) == {"workload", "tetrahedra", "rays_per_step_per_gpu", "parallelism"}
    assert "libtetranerf_b200" not in out.stderr
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_does_not_load_the_cuda_library():
    """the CPU arm times the oracle only: importing bench.py's reference path must not dlopen libtetranerf_b200.so"""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '3', '--rays', '64'];"
            "runpy.run_path(%r, run_name='__main__');"
            "maps = open('/proc/self/maps').read(); assert 'libtetranerf_b200' not in maps, 'CUDA library was loaded'; assert 'liboracle' in maps" % str(ROOT / "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=str(ROOT), env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
