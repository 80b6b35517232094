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
        # colour gradient, 2 fields, 384^3 pack: host tables would need the reference-table path, so the
        # two-field lattice is created through the generic builder at a size it handles quickly (128^3)
        size = int(os.environ.get("TWOPHASE_SIZE", "128"))
        geo = pkg.geometry.sphere_pack((size,) * 3, size / 8.0, 0.35, 1234).astype(int)
        t0 = time.time()
        lg = pkg.geometry.LatticeGeometry(geo, "D3Q19", "xyz")
        t = lg.all_ranks()[0]
        x = np.arange(size)[:, None, None] * np.ones(geo.shape)
        rho0 = (x < size / 2).astype(float)
        setup = pkg.cases.two_phase_setup(lg, [t], rho0, 1.0 - rho0, 0.5 * (geo == 0))[0]
        lat = capi.Lattice.from_rank_tables(t, n_fields=2)
        lat.add_halfway_bb(*t.halfway_bb(t.bulk_nodes()))
        lat.set_solid_boundary(setup["solid_bnd"])
        lat.finalize(capi.INDEX_COMPACT)
        lat.set_twophase_density(setup["rho"])
        lat.upload(setup["f0"])
        n = len(t.bulk_nodes())
        args = (1.0, 1.0, 0.01, 1.0, 1e-5, (0, 0, 0), n)
        lat.step_twophase(5, *args)
        lat.synchronize()
        torch.cuda.synchronize()
        setup_s = time.time() - t0
        t1 = time.perf_counter()
        lat.step_twophase(100, *args)
        lat.synchronize()
        ms = (time.perf_counter() - t1) * 1e3
        report("twophase colour gradient D3Q19 sphere pack %d^3 (configs[3] physics)" % size, n, ms, 100, 624.0,
               "host-clock timing of 100 steps (3 launches per step); setup %.1f s" % setup_s)


if __name__ == "__main__":
    main()
