"""Host-side pieces of bench.py that must not depend on a GPU being present (CPU box): the clock sampler degrades to an
empty sample set instead of raising, the thread-count helpers stay within the affinity mask, and the reference arm's JSON
line carries the keys the driver reads."""
import importlib.util
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_clock_sampler_without_a_gpu_reports_no_samples():
    b = _bench()
    s = b.ClockSampler(0)
    s.start()
    time.sleep(0.3)
    s.stop_flag = True
    s.join(timeout=10)
    r = s.result()
    assert set(r) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples", "source"}
    assert r["source"] in ("nvml", "nvidia-smi")
    assert r["samples"] == 0 or r["sm_mhz"] > 0  # no driver here: zero samples; on a GPU box: real clocks


def test_host_threads_respects_affinity():
    b = _bench()
    n = b.host_threads()
    assert 1 <= n <= min(64, len(os.sched_getaffinity(0)))


def test_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` on a non-zero rank prints nothing and exits 0 (torchrun launches it on every rank)."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
