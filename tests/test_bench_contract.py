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


def test_clock_sampler_reports_throttle_reasons_inside_the_window():
    """The NVML path: median SM clock and the union of the throttle-reason bits over the samples that fall inside the
    timed region (hw_slowdown 0x8, sw_thermal 0x20, hw_thermal 0x40, sw_power_cap 0x4); samples outside are ignored."""
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.nvml, s.alive, s.mx, s.proc = object(), True, 1965.0, None
    s.rows = [(5.0, 1200.0, 0x8), (10.1, 1965.0, 0x0), (10.2, 1950.0, 0x4), (10.3, 1965.0, 0x0), (10.9, 900.0, 0x40)]
    out = s.stop(10.0, 10.4)
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["reasons"] == ["sw_power_cap"] and out["source"] == "nvml"
    s.rows.append((10.35, 800.0, 0x8 | 0x20))
    s.alive = True
    assert s.stop(10.0, 10.4)["reasons"] == ["hw_slowdown", "sw_power_cap", "sw_thermal_slowdown"]
