"""not-gpu: the bookkeeping of bench.py that decides what the JSON line claims (peak regime, clock windows, traffic lookup)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)

PEAKS = dict(hbm_gbs=6545.9, tflops_sustained=1394.8, tflops_burst=1671.4, sm_max_mhz=1965.0, source="test")


def test_peak_follows_the_clock_regime_the_kernel_was_timed_in():
  burst = dict(sm_mhz=1965.0, sm_max_mhz=1965.0, reasons=[])
  assert bench.regime_peak(PEAKS, burst) == ("burst", 1671.4)
  assert bench.regime_peak(PEAKS, dict(burst, reasons=["sw_power_cap"])) == ("sustained", 1394.8)       # capped: sustained figure
  assert bench.regime_peak(PEAKS, dict(burst, sm_mhz=1560.0)) == ("sustained", 1394.8)                    # clocks well below max
  assert bench.regime_peak(PEAKS, dict(sm_mhz=None, sm_max_mhz=None, reasons=["no clock samples"])) == ("sustained", 1394.8)


def test_clock_windows_are_cut_by_host_time():
  s = bench.ClockSampler(0)
  s.max_mhz, s.source = 1965.0, "test"
  s.rows = [(0.000, 1965.0, set(), 250.0), (0.010, 1965.0, set(), 260.0), (0.020, 1800.0, {"sw_power_cap"}, 990.0),
            (0.030, 1550.0, {"sw_power_cap"}, 1000.0), (0.040, 1545.0, {"sw_power_cap"}, 1001.0)]
  w = s.window(0.0, 0.012)
  assert w["sm_mhz"] == 1965.0 and w["reasons"] == [] and w["samples"] == 2
  w = s.window(0.018, 0.041)
  assert w["sm_mhz"] == 1550.0 and w["reasons"] == ["sw_power_cap"] and w["sm_min_mhz"] == 1545.0 and w["power_w_max"] == 1001.0
  w = s.window(0.0149, 0.0151)                       # shorter than the sampling period: nearest samples, and it says so
  assert w["samples"] == 2 and "note" in w
  e = bench.ClockSampler(0)
  assert e.window(0, 1)["samples"] == 0 and e.window(0, 1)["reasons"] == ["no clock samples"]


def test_traffic_comes_from_the_committed_ncu_capture():
  t = bench.ncu_traffic("hyper_synthesis.layer_2", 24)
  assert t is not None and 1.5e8 < t < 2.0e8          # profiles/r02_ncu_traffic.json: 172.8 MB per launch
  assert bench.ncu_traffic("hyper_synthesis.layer_2", 8) is None and bench.ncu_traffic("no.such.layer", 24) is None
  assert bench.MMA_PASSES == 3
