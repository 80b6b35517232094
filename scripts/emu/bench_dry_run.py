"""DEVELOPMENT TOOL (see include/cuda_runtime.h): bench.py's one-GPU default flow on the host model, every workload
shrunk to a few thousand nodes.  Run from the emulation copy:  scripts/emu/run.sh is not needed again after a build --
  cd /tmp/chimp_emu/repo && CHIMP_EMU=1 PYTHONPATH=/root/repo/scripts/emu python /root/repo/scripts/emu/bench_dry_run.py [bench.py arguments]
Nothing it prints is a measurement."""
import importlib
import os
import sys

import emu_plugin  # noqa: F401  (torch device shim, libcudart stand-in)

ROOT = os.getcwd()
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402

helpers.load_package()
W = importlib.import_module("badchimp_cpp_b200.workloads")
SMALL = {"std_case": 32, "trt": 32, "one_phase": 32, "twophase": 32, "d2q9_channel": 96, "d3q27_dense": 16}
for k, s in SMALL.items():
    W.WORKLOADS[k]["size"] = s
bi = importlib.import_module("badchimp_cpp_b200.bench_impl")
bi.EXTRAS_BUDGET_S = 1e9   # the model is slow: no entry is skipped for time
_orig = bi.cpu_baseline_port
bi.cpu_baseline_port = lambda pkg, size=80, seconds_target=12.0: _orig(pkg, size=24, seconds_target=0.5)
sys.argv = ["bench.py", "--steps", "4", "--warmup", "3", "--no-traffic"] + sys.argv[1:]
import bench  # noqa: E402

bench.main()
