"""bench.py contract pieces that need no GPU: stdout carries exactly ONE JSON line (libraries writing to file descriptor 1
are redirected to stderr), the reference arm answers for configurations it does not time, the clock sampler degrades to
"unavailable" without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_for_untimed_configs():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C3"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-400:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and "unavailable" in d


def test_descriptor_one_is_pointed_at_stderr_for_everyone_else():
    code = ("import sys, os; sys.argv = ['bench.py', '--impl', 'reference', '--config', 'C5']; sys.path.insert(0, %r); "
            "import bench; bench.main(); os.write(1, b'banner from a library\\n')" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-400:]
    assert len([l for l in p.stdout.splitlines() if l.strip()]) == 1 and "banner from a library" in p.stderr


def test_clock_sampler_without_a_gpu():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    out = s.stop(0.0, 1e12)
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}
