"""bench.py's JSON contract on the CPU side: the reference arm (the reference's compiled CPU filter, or the oracle port, on the host cores) prints one line with
the keys the driver reads; the algorithmic-byte accounting matches SURVEY.md §8(d)."""
import json
import os
import subprocess
import sys

ROOT_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT_ not in sys.path:
    sys.path.insert(0, ROOT_)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1",
                          "--steps", "2", "--warmup", "3", "--cpu-particles", "64"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-updates/sec" and d["unit"] == "particle-updates/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 3
    from oracle import ref_binding               # "reference" where oracle/_ref is built (or buildable), else the port
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_binding.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "C1"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_follow_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    n, c = 1048576, 64
    step, lean = bench.algorithmic_bytes(n, c, True)
    _, generic = bench.algorithmic_bytes(n, c, False)
    assert step == n * (8 * c + 234) == 782237696            # 782 MB per C3 step
    assert lean == n * (8 * c + 82) and generic == n * (8 * c + 97)
