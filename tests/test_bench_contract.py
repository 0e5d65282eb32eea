"""Host-side checks of bench.py's contract that need no GPU: the measured-traffic figure quoted in `roofline.traffic`
belongs to the CURRENT kernel sources (profiles/traffic.json is stamped with their hash and bench.py drops a stale
entry -- this test makes a kernel edit without a fresh ncu capture visible here instead of on the GPU box), and the
workload table / extra configs name what BASELINE.json names."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_traffic_entry_of_the_headline_kernel_is_current():
    import bench
    traffic, note = bench.measured_traffic("alarm37:fp64:bnbp_onchip_run")
    assert traffic is not None, note
    # the on-chip kernel moves evidence in and marginals out only: tens of MB per sweep of 1M cases, not the 7.4 GB of a streaming sweep
    assert 1e7 < traffic < 2e8
    src = note.split("profiles/")[1].split(" ")[0]
    assert os.path.exists(os.path.join(ROOT, "profiles", src)), "the capture the figure comes from must be committed under profiles/"


def test_bench_configs_match_baseline_json():
    import bench
    from bayesiannetwork_b200 import synth
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) == 5
    keys = [c[0] for c in bench.EXTRA_CONFIGS]
    assert keys == ["cfg3", "cfg4", "cfg5"]
    for key, workload, total, gpus, precisions, parity, binding in bench.EXTRA_CONFIGS:
        assert workload in synth.WORKLOADS and binding in ("hbm", "fma", "tensor")
        assert total >= 1 << 16 and gpus in (1, 8)
    assert synth.WORKLOADS["alarm37"][1] == 1 << 20 and synth.WORKLOADS["alarm37"][3] == 20
