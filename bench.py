#!/usr/bin/env python
"""bench.py -- MLUPS of the fused collide-and-stream hot path (see DESIGN.md, section Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload W] [--size S]

One "step" = one full iteration (collide, stream, halo, boundary) over the whole lattice.
Default workload: D3Q19 BGK + Guo force + half-way bounce back in a periodic random sphere
pack 512^3 (porosity ~0.35, sphere radius size/8, seed 1234), BASELINE.json configs[2] geometry run
with the std_case physics of configs[0]; MLUPS counts fluid nodes only.  With --gpus N (under torchrun)
the same pack is split into N z-slabs (strong scaling); the weak-scaling number is reported next to it.
--workload selects the other BASELINE.json configurations (trt, one_phase, d2q9_channel, twophase,
d3q27_dense); every line carries roofline, e2e, clocks and a parity probe against the oracle.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--size", type=int, default=0, help="edge length of the cubic sphere pack (0 = default)")
    ap.add_argument("--index", default="compact", choices=["compact", "table"])
    ap.add_argument("--skip-mask", default="auto", choices=["auto", "on", "off"],
                    help="N=1, compact index, plain single-field step: kernel form that skips the delta words marked as plain runs; "
                         "auto = both forms are tried for a few untimed steps after the warm-up and the faster one is timed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="std_case", choices=["std_case", "trt", "one_phase", "d2q9_channel", "twophase", "d3q27_dense"])
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"], help="N>1: split one lattice (strong) or one block per GPU (weak); default per workload")
    ap.add_argument("--interior-domains", action="store_true", help="one_phase: two interior domains with mass sources (per-step mass-change sum)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity probe against the oracle port before timing")
    ap.add_argument("--no-extra-workloads", action="store_true", help="N=1, default workload: do not time the other BASELINE configurations in the same run")
    ap.add_argument("--no-weak", action="store_true", help="N>1: skip the secondary weak-scaling measurement")
    ap.add_argument("--no-traffic", action="store_true", help="N=1: skip the ncu child run that measures DRAM traffic of one step")
    ap.add_argument("--traffic-probe", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"], help="N>1: halo transport")
    ap.add_argument("--no-balance", dest="balance", action="store_false", help="N>1: equal-thickness z-slabs instead of equal fluid-node counts")
    args = ap.parse_args()
    import helpers
    helpers.load_package()
    import importlib
    bench_impl = importlib.import_module("badchimp_cpp_b200.bench_impl")
    if args.traffic_probe:
        bench_impl.run_traffic_probe(args)
    elif args.impl == "reference":
        bench_impl.run_reference(args)
    else:
        bench_impl.run_b200(args)


if __name__ == "__main__":
    main()
