"""FP64 roofline denominator with its own clock record (MEASURED_PEAKS.json has no FP64 entry).
    python profiles/fp64_peak.py            -> one JSON line: the DFMA-chain microbenchmark of the C ABI
    (ekf_measure_fp64_peak: 8 CTAs/SM x 256 threads x 8 independent DFMA chains), best of 5 launches,
    repeated 5 times, with nvidia-smi SM clocks / throttle reasons sampled during the runs and the spec
    estimate beside it (SMs x 64 DFMA/clk x 2 flop x SM clock). Cross-check with the pipe counter:
    ncu --metrics sm__inst_executed_pipe_fp64.sum,sm__cycles_elapsed.max,gpu__time_duration.sum -k regex:fp64_peak python profiles/fp64_peak.py
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    ekf = bench.load_product()
    s = bench.ClockSampler(0)
    s.start()
    vals = [ekf.measure_fp64_peak(0) for _ in range(5)]
    clocks = s.stop()
    sms = ekf.device_info(0)["sm_count"]
    mhz = clocks.get("sm_mhz") or 0
    print(json.dumps({"fp64_peak_flops": max(vals), "all_runs": vals, "clocks": clocks, "sm_count": sms,
                      "spec_estimate_flops": sms * 64 * 2 * mhz * 1e6,
                      "how": "ekf_measure_fp64_peak: DFMA chains, 2 flop per DFMA, best of 5 launches per call"}))
