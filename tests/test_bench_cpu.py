"""CPU tests of bench.py's host-side pieces (no GPU, no oracle): the clock sampler keeps only the nvidia-smi rows whose own time stamps
fall inside the timed region, and the ncu raw-page reader finds the committed traffic figure of the kernel that ran."""
import datetime
import importlib.util
import os
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    return b


def test_clock_sampler_uses_only_rows_inside_the_timed_region():
    b = _bench()
    cs = b.ClockSampler.__new__(b.ClockSampler)                # no nvidia-smi here: feed it a recorded file
    cs.idx = 0
    cs.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    now = datetime.datetime.now()
    for k in range(-10, 20):                                   # 10 idle rows before the region (120 MHz), then busy rows; one power-capped
        ts = (now + datetime.timedelta(milliseconds=50 * k)).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        cs.f.write("%s, 0, %d, 1965, 600.0, 0x0, Not Active, Not Active, Not Active, %s\n" % (ts, 1965 if k >= 0 else 120, "Active" if k == 3 else "Not Active"))
    cs.f.write("garbage line\n")
    cs.f.flush()

    class P:
        def terminate(self): pass
        def wait(self, timeout=None): pass
    cs.p = P()
    cs.t0, cs.t1 = now, now + datetime.timedelta(milliseconds=400)
    out = cs.stop()
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]
    assert 8 <= out["samples"] <= 10                           # rows 0..8 (+- the 30 ms slack), none of the idle ones


def test_traffic_is_read_from_the_committed_ncu_page_of_the_kernel_that_ran():
    b = _bench()
    for kernel in ("lstm_scan3_kernel", "lstm_scan4_kernel", "lstm_scan5_kernel"):
        t = b.ncu_traffic(kernel)
        assert t is not None and t["bytes"] > 1e8 and t["file"].startswith("profiles/"), (kernel, t)
    assert b.ncu_traffic("no_such_kernel") is None
