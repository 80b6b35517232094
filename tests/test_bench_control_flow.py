"""bench.py's harness on N CPU ranks (gloo) with a do-nothing stand-in for the engine (tests/bench_flow_stub.py): every
rank reaches every collective and rank 0 prints exactly one JSON line with the keys of the contract -- for the default
workload (strong split + secondary weak run + e2e leg) and for the workloads with other decompositions."""
import json
import os
import socket
import subprocess
import sys

import pytest

import helpers

STUB = os.path.join(helpers.ROOT, "tests", "bench_flow_stub.py")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, workload, port):
    env = dict(os.environ, WL=workload)
    if world == 1:
        cmd = [sys.executable, STUB]
        env.update(RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(port), STUB]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.parametrize("world,workload", [(1, "std_case"), (2, "std_case"), (3, "std_case"), (2, "twophase"), (2, "d3q27_dense"), (2, "one_phase")])
def test_every_rank_reaches_every_collective_and_one_line_is_printed(world, workload):
    line = _run(world, workload, _free_port())
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "parity"):
        assert key in line, key
    assert line["n_gpus"] == world and line["e2e"]["value"] is not None
    assert line["config"]["workload_key"] == workload
    if world == 1:
        assert "cpu_baseline" in line and "e2e_from_init_rho" in line
        assert [e["workload_key"] for e in line["other_workloads"]] == ["trt", "one_phase", "one_phase+interior_domains", "twophase",
                                                                        "d2q9_channel", "d3q27_dense"]
        assert all("error" not in e and "skipped" not in e for e in line["other_workloads"]), line["other_workloads"]
    else:
        assert len(line["config"]["nodes_per_rank"]) == world
        assert line["scaling"] == ("weak" if workload == "d3q27_dense" else "strong")
        assert ("weak" in line) == (workload != "d3q27_dense")
        # the stand-in gives odd and even ranks different slab densities: the check is over the whole lattice
        assert line["e2e"]["mean_rho_error"] < 2e-5
