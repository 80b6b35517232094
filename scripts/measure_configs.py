#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one B200 (not bench lines: reported in
DESIGN.md / profiles).  usage: python scripts/measure_configs.py [d2q9] [d3q27] [twophase] [trt] [onephase]"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402

pkg = helpers.load_package()
ingest = importlib.import_module("badchimp_cpp_b200.ingest")
capi = pkg.capi
PEAK = 6468.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def report(name, n, ms, steps, b_alg, extra=""):
    mlups = n * steps / (ms * 1e-3) / 1e6
    gbs = b_alg * n / (ms / steps * 1e-3) / 1e9
    print(json.dumps({"config": name, "fluid_nodes": n, "steps": steps, "ms_per_step": ms / steps, "MLUPS": mlups,
                      "B_alg": b_alg, "achieved_GBs": gbs, "frac_of_measured_hbm": gbs / PEAK, "note": extra}), flush=True)


def single(name, fluid, lattice, periodic, b_alg, steps=60, trt=None, tau=0.8, force=(1e-6, 0, 0)):
    table, labels, n, n_pad = ingest.build_pull_table(fluid, lattice, periodic)
    del fluid
    lat = capi.lattice_from_device_table(lattice, n, n_pad, 0, table.data_ptr(), labels.data_ptr(), 1, capi.INDEX_COMPACT)
    del table, labels
    torch.cuda.empty_cache()
    lat.init_uniform(1.0)
    lat.step_single(5, tau=tau, force=force, trt=trt)
    lat.synchronize()
    ms = lat.step_timed(steps, tau=tau, force=force, trt=trt)
    rho, _ = lat.download_moments_device_order()
    report(name, n, ms, steps, b_alg, "irregular %.4f, index %.1f B/node, mean rho err %.1e" % (
        lat.irregular_fraction(), lat.index_bytes_per_node(), abs(rho.mean() - 1)))
    lat.close()


def main():
    what = sys.argv[1:] or ["d2q9", "d3q27", "trt", "twophase"]
    if "d2q9" in what:
        g = torch.ones((8192, 8192), dtype=torch.bool, device="cuda")
        g[:, 0] = False
        g[:, -1] = False
        single("D2Q9 SRT Poiseuille channel 8192x8192 (configs[1])", g, "D2Q9", "x", 144.0, force=(1e-7, 0))
    if "d3q27" in what:
        g = torch.ones((512, 512, 512), dtype=torch.bool, device="cuda")
        single("D3Q27 BGK dense periodic 512^3 (configs[4], one GPU)", g, "D3Q27", "xyz", 432.0, steps=30)
    if "trt" in what:
        geo = pkg.geometry.sphere_pack((512,) * 3, 64.0, 0.35, 1234)
        single("D3Q19 TRT sphere pack 512^3 (configs[2] collision)", torch.from_numpy(geo).cuda().bool(), "D3Q19", "xyz", 304.0,
               trt=(0.8, 1.125))
    if "twophase" in what:
        # colour gradient, 2 LbFields, sphere pack 384^3 (configs[3] geometry on one GPU), device-side ingest
        size = int(os.environ.get("TWOPHASE_SIZE", "384"))
        geo = pkg.geometry.sphere_pack((size,) * 3, size / 8.0, 0.35, 1234)
        fluid = torch.from_numpy(geo).cuda().bool()
        table, labels, n, n_pad = ingest.build_pull_table(fluid, "D3Q19", "xyz")
        lat = capi.lattice_from_device_table("D3Q19", n, n_pad, 0, table.data_ptr(), labels.data_ptr(), 2, capi.INDEX_COMPACT)
        del table, labels
        wall_phi = torch.zeros(geo.shape, dtype=torch.float64)          # wettability 0.5: rho0 = rho1 at the wall
        ptable, n_extra, phi_extra = ingest.build_phi_table(fluid, wall_phi, "D3Q19", "xyz")
        lat.set_phi_table_dev(ptable.data_ptr(), n_extra, phi_extra.data_ptr())
        x = torch.arange(size, device="cuda")[:, None, None].expand(size, size, size)
        r0 = (x < size // 2).double()[fluid]
        rho_dev = torch.stack([r0, 1.0 - r0]).contiguous()
        lat.init_equilibrium_dev(rho_dev.data_ptr())
        del ptable, fluid, rho_dev
        torch.cuda.empty_cache()
        args = (1.0, 1.0, 0.01, 1.0, 1e-5, (0, 0, 0), n)
        lat.step_twophase(5, *args)
        lat.synchronize()
        t1 = time.perf_counter()
        lat.step_twophase(60, *args)
        lat.synchronize()
        ms = (time.perf_counter() - t1) * 1e3
        rho, _ = lat.download_moments_device_order()
        report("twophase colour gradient D3Q19 sphere pack %d^3 (configs[3] physics, one GPU)" % size, n, ms, 60, 624.0,
               "host-clock timing of 60 steps (3 launches per step); flux force %.3e" % lat.last_flux_force())


if __name__ == "__main__":
    main()
