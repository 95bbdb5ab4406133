"""Host-side pieces of bench.py that need no GPU: the nvidia-smi clock sampler's parsing (per-GPU medians, throttle reasons)
and the argument defaults the driver relies on."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class FakeProc:
    def terminate(self):
        pass

    def wait(self, timeout=None):
        return 0

    def kill(self):
        pass


def test_clock_sampler_reports_every_gpu_and_the_union_of_reasons():
    b = load_bench()
    s = b.ClockSampler(0, range(2))
    s.proc = FakeProc()
    s.lines = ["0, 1965, 1965, 650.1, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
               "1, 1800, 1965, 990.0, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
               "0, 1950, 1965, 640.0, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
               "1, 1830, 1965, 985.0, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
               "garbage line", "0, [N/A], 1965, 1, 0, x, x, x, x"]
    out = s.stop()
    assert out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 2  # this rank's GPU
    assert out["per_gpu_sm_mhz"] == {"0": 1957.5, "1": 1815.0}
    assert out["reasons"] == ["sw_power_cap"]


def test_clock_sampler_single_gpu_and_missing_tool():
    b = load_bench()
    s = b.ClockSampler(3)
    assert s.stop()["reasons"] == ["nvidia-smi unavailable"]  # never started
    s.proc = FakeProc()
    s.lines = ["3, 1965, 1965, 600, 0, Not Active, Not Active, Not Active, Not Active"]
    out = s.stop()
    assert out["sm_mhz"] == 1965.0 and "per_gpu_sm_mhz" not in out and out["reasons"] == []


def test_argument_defaults():
    b = load_bench()
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        a = b.parse()
    finally:
        sys.argv = argv
    assert a.gpus == 1 and a.warmup >= 3 and a.impl != "reference" and not a.distinct_sweeps and a.batch_pairs == 4096
