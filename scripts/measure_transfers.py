#!/usr/bin/env python
"""Where the end-to-end cycle of bench.py spends its time: upload of the LbField (reference AoS, pinned host), K steps,
download of rho and vel -- each timed separately on the bench workload (sphere pack 512^3)."""
import ctypes as C
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
size = int(os.environ.get("SIZE", "512"))
geo = pkg.geometry.sphere_pack((size,) * 3, size / 8.0, 0.35, 1234)
table, labels, n, n_pad = ingest.build_pull_table(torch.from_numpy(geo).cuda().bool(), "D3Q19", "xyz")
lat = capi.lattice_from_device_table("D3Q19", n, n_pad, 0, table.data_ptr(), labels.data_ptr(), 1, capi.INDEX_COMPACT)
del table, labels
torch.cuda.empty_cache()
host_f = torch.empty((n + 1, 19), dtype=torch.float64, pin_memory=True)
host_f[:] = torch.from_numpy(pkg.cases.lattice_weights("D3Q19"))[None, :]
host_rho = torch.empty((n + 1,), dtype=torch.float64, pin_memory=True)
host_vel = torch.empty((n + 1, 3), dtype=torch.float64, pin_memory=True)
lib = capi.lib()
p = lat._single_params(0.8, (1e-6, 0, 0), None)
# raw PCIe reference: a plain pinned copy of the same size through torch
dev = torch.empty(host_f.shape, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter(); dev.copy_(host_f, non_blocking=True); torch.cuda.synchronize(); t_raw_h2d = time.perf_counter() - t0
t0 = time.perf_counter(); host_f.copy_(dev, non_blocking=True); torch.cuda.synchronize(); t_raw_d2h = time.perf_counter() - t0
del dev
torch.cuda.empty_cache()
out = {"fluid_nodes": n, "raw_h2d_GBs": host_f.numel() * 8 / t_raw_h2d / 1e9, "raw_d2h_GBs": host_f.numel() * 8 / t_raw_d2h / 1e9}
for rep in range(3):
    t0 = time.perf_counter()
    capi._check(lib.chimp_upload_lbfield(lat.h, C.c_void_p(host_f.data_ptr())))
    t1 = time.perf_counter()
    capi._check(lib.chimp_step_single(lat.h, C.byref(p), C.c_int(20)))
    lat.synchronize()
    t2 = time.perf_counter()
    capi._check(lib.chimp_download_rho(lat.h, C.c_void_p(host_rho.data_ptr()), C.c_int(1)))
    t3 = time.perf_counter()
    capi._check(lib.chimp_download_vel(lat.h, C.c_void_p(host_vel.data_ptr())))
    t4 = time.perf_counter()
    out["rep%d" % rep] = {"upload_s": t1 - t0, "upload_GBs": host_f.numel() * 8 / (t1 - t0) / 1e9, "steps20_s": t2 - t1,
                          "download_rho_s": t3 - t2, "download_vel_s": t4 - t3,
                          "download_GBs": (host_rho.numel() + host_vel.numel()) * 8 / (t4 - t2) / 1e9}
assert abs(float(host_rho[1:].mean()) - 1.0) < 1e-9
print(json.dumps(out, indent=1))
