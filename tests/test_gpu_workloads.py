"""Every bench workload (BASELINE.json configurations, badchimp-cpp_b200/workloads.py) on its own code path --
structured ingest on the device, the step kernels, transfers in reference layout -- against the oracle port on a
small case.  This is the probe bench.py runs before it times anything; N-rank versions run in
tests/multi_gpu_check.py."""
import importlib

import pytest

import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("workload,interior", [("std_case", False), ("trt", False), ("one_phase", False), ("one_phase", True),
                                               ("d2q9_channel", False), ("twophase", False), ("d3q27_dense", False)])
@pytest.mark.parametrize("index", ["compact", "table"])
def test_workload_path_matches_oracle_port(workload, interior, index):
    import torch
    pkg = helpers.load_package()
    ingest = importlib.import_module("badchimp_cpp_b200.ingest")
    multi = importlib.import_module("badchimp_cpp_b200.multi")
    bench_impl = importlib.import_module("badchimp_cpp_b200.bench_impl")
    W = importlib.import_module("badchimp_cpp_b200.workloads")
    wl = W.WORKLOADS[workload]
    form = pkg.capi.INDEX_COMPACT if index == "compact" else pkg.capi.INDEX_TABLE
    res = bench_impl.parity_probe(pkg, ingest, multi, wl, workload, 0, 1, torch.device("cuda", 0), form, "peer", interior)
    assert res["checked_nodes"] > 1000
    if wl["physics"] == "single":
        assert res["bit_exact"] and res["after_10_steps"]["bit_exact"], res
    else:
        assert res["max_rel_f"] <= 1e-12 and res["after_10_steps"]["max_rel_f"] <= 1e-9, res
