"""Measurement harness behind bench.py (see DESIGN.md "Measurement").

Workload (N=1): D3Q19 BGK + Guo body force + half-way bounce back (std_case/main.cpp:109-146)
in a periodic random sphere pack, edge 512, sphere radius 64, target porosity 0.35, seed 1234
(BASELINE.json configs[2] geometry).  MLUPS = own fluid nodes x steps / seconds / 1e6.
For N>1 every rank owns one 512^3 block of a 512 x 512 x (512 N) pack (weak scaling, z-slabs).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B_ALG = {"D2Q9": 144.0, "D3Q19": 304.0, "D3Q27": 432.0}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic_gb(lattice, index, n_nodes):
    """DRAM bytes per launch of the step kernel from the committed ncu capture (profiles/traffic.json)"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["%s/%s" % (lattice, index)]
        return t["dram_bytes_per_node"] * n_nodes / 1e9, t["source"]
    except Exception:
        return None, None


def _clock_sampler(stop, out):
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    dev = os.environ.get("LOCAL_RANK", "0")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-i", dev],
                               capture_output=True, text=True, timeout=5)
            parts = [p.strip() for p in r.stdout.strip().split(",")]
            if len(parts) >= 7:
                out.append(parts)
        except Exception:
            pass
        stop.wait(0.1)


def _summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = sorted(float(s[0]) for s in samples)
    reasons = set()
    for s in samples:
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
            if v.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(samples[0][1]), "reasons": sorted(reasons),
            "samples": len(samples)}


def cpu_baseline_port(pkg, size=80, seconds_target=12.0):
    """oracle port (plain-C restatement, 1 core) on a bounded sample of the same workload"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import port
    G = pkg.geometry
    geo = G.sphere_pack((size,) * 3, size / 8.0, 0.35, 1234).astype(int)
    lg = G.LatticeGeometry(geo, "D3Q19", "xyz")
    t = lg.all_ranks()[0]
    bulk = t.bulk_nodes()
    pr = port.PortRank(1, t.neigh, bulk, 1, t.halfway_bb(t.fluid_bnd_nodes()))
    pr.f[:] = pkg.cases.std_case_initial_state(t, np.ones(geo.shape))[0]
    pr.step_std_case(2, tau=0.8, force=(1e-6, 0, 0))
    t0 = time.perf_counter()
    pr.step_std_case(5, tau=0.8, force=(1e-6, 0, 0))
    per = (time.perf_counter() - t0) / 5
    steps = max(5, min(400, int(seconds_target / per)))
    t0 = time.perf_counter()
    pr.step_std_case(steps, tau=0.8, force=(1e-6, 0, 0))
    dt = time.perf_counter() - t0
    return {"value": len(bulk) * steps / dt / 1e6, "unit": "MLUPS", "cores": 1, "kind": "port",
            "sample": "oracle/lb_port.c, D3Q19 BGK sphere pack %d^3 (%d fluid nodes), %d steps, %.1f s" % (size, len(bulk), steps, dt)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref/ref_driver = unmodified
    reference headers, ranks as threads) on the host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    pkg = helpers.load_package()
    G = pkg.geometry
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    cores = os.cpu_count() or 1
    size = args.size or 128
    nranks = 1
    for p in (32, 16, 8, 4, 2):
        if p <= cores and size % p == 0:
            nranks = p
            break
    geo = G.sphere_pack((size,) * 3, size / 8.0, 0.35, 1234).astype(int)
    steps, warm = max(1, min(args.steps, 40)), max(0, min(args.warmup, 5))
    line = {"metric": "MLUPS", "unit": "MLUPS", "impl": "reference", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "std_case D3Q19 BGK+Guo+half-way BB, periodic sphere pack %d^3 porosity~0.35 (bounded sample of the 512^3 workload)" % size}}
    if os.path.exists(driver):
        d = tempfile.mkdtemp(prefix="chimp_ref_")
        os.makedirs(os.path.join(d, "out"))
        lg = G.LatticeGeometry(G.z_slab_rank_map(geo, nranks), "D3Q19", "xyz")
        for t in lg.all_ranks():
            t.write_vtklb(os.path.join(d, "tmp%d.vtklb" % t.my_rank), {"init_rho": np.ones(geo.shape)})
        cmd = [driver, "--case", "std_case", "--lattice", "D3Q19", "--dir", d, "--out", os.path.join(d, "out"),
               "--nranks", str(nranks), "--no-tables", "--no-f", "--time", "--tau", "0.8", "--force", "1e-6,0,0"]
        if warm:
            subprocess.run(cmd + ["--steps", str(warm)], capture_output=True, text=True)
        r = subprocess.run(cmd + ["--steps", str(steps)], capture_output=True, text=True)
        res = json.loads(r.stdout.strip().splitlines()[-1])
        value, kind = res["mlups"], "reference"
        ms = res["loop_seconds"] / steps * 1e3
        sample = "oracle/_ref/ref_driver (unmodified reference headers), %d ranks as threads, %d^3 pack, %d fluid nodes, %d steps" % (
            nranks, size, res["fluid_nodes"], steps)
        import shutil
        shutil.rmtree(d, ignore_errors=True)
    else:
        base = cpu_baseline_port(pkg, size=min(size, 96))
        value, kind, sample, nranks = base["value"], "port", base["sample"], 1
        ms = None
    line.update({"value": value, "ms_per_step": ms,
                 "cpu_baseline": {"value": value, "unit": "MLUPS", "cores": nranks, "kind": kind, "sample": sample},
                 "e2e": {"value": value, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    pkg = helpers.load_package()
    import importlib
    ingest = importlib.import_module("badchimp_cpp_b200.ingest")
    capi = pkg.capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")   # NCCL kernels ahead of the interior blocks
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    size = args.size or 512
    lattice = "D3Q19"
    tau, force = 0.8, (1e-6, 0.0, 0.0)
    if world > 1:
        from . import multi  # noqa
        return multi.run_weak_scaling(args, pkg, ingest, size, lattice, tau, force)

    t_setup = time.perf_counter()
    geo = pkg.geometry.sphere_pack((size,) * 3, size / 8.0, 0.35, 1234)
    fluid = torch.from_numpy(geo).to("cuda").bool()
    table, labels, n, n_pad = ingest.build_pull_table(fluid, lattice, "xyz")
    del fluid
    index_form = capi.INDEX_COMPACT if args.index == "compact" else capi.INDEX_TABLE
    lat = capi.lattice_from_device_table(lattice, n, n_pad, 0, table.data_ptr(), labels.data_ptr(), 1, index_form, local)
    del table, labels
    torch.cuda.empty_cache()
    lat.init_uniform(1.0)
    setup_s = time.perf_counter() - t_setup

    launches0 = capi.lib().chimp_launch_count()
    lat.step_single(args.warmup, tau=tau, force=force)
    lat.synchronize()
    samples, stop = [], threading.Event()
    th = threading.Thread(target=_clock_sampler, args=(stop, samples), daemon=True)
    th.start()
    launches1 = capi.lib().chimp_launch_count()
    ms = lat.step_timed(args.steps, tau=tau, force=force)
    launches2 = capi.lib().chimp_launch_count()
    # keep the GPU busy a little longer so the sampler sees clocks under load
    t_extra = time.perf_counter()
    while len(samples) < 8 and time.perf_counter() - t_extra < 3.0:
        lat.step_single(20, tau=tau, force=force)
        lat.synchronize()
    stop.set()
    th.join()
    mlups = n * args.steps / (ms * 1e-3) / 1e6
    peak, peak_src = _peaks()
    kernel_ms = ms / args.steps
    achieved = B_ALG[lattice] * n / (kernel_ms * 1e-3) / 1e9

    # mass is conserved by collide + stream + bounce back: a size-independent check at full size
    rho = np.zeros(n)
    capi._check(capi.lib().chimp_download_moments_device_order(lat.h, rho.ctypes.data_as(C.c_void_p), None))
    mass_err = abs(rho.sum() / n - 1.0)

    # end to end through the C-ABI with host buffers: upload f (reference AoS layout, pinned),
    # K steps, download rho and vel -- one "write interval" of the reference main
    e2e = None
    try:
        host_f = torch.empty((n + 1, 19), dtype=torch.float64, pin_memory=True)
        w = pkg.cases.lattice_weights(lattice)
        host_f[:] = torch.from_numpy(w)[None, :]
        host_rho = torch.empty((n + 1,), dtype=torch.float64, pin_memory=True)
        host_vel = torch.empty((n + 1, 3), dtype=torch.float64, pin_memory=True)
        lib = capi.lib()
        p = lat._single_params(tau, force, None)
        # one untimed cycle first (first-touch of the pinned pages by the copy engines), like the warm-up steps
        capi._check(lib.chimp_upload_lbfield(lat.h, C.c_void_p(host_f.data_ptr())))
        capi._check(lib.chimp_step_single(lat.h, C.byref(p), C.c_int(1)))
        capi._check(lib.chimp_download_rho(lat.h, C.c_void_p(host_rho.data_ptr()), C.c_int(1)))
        capi._check(lib.chimp_download_vel(lat.h, C.c_void_p(host_vel.data_ptr())))
        t0 = time.perf_counter()
        capi._check(lib.chimp_upload_lbfield(lat.h, C.c_void_p(host_f.data_ptr())))
        capi._check(lib.chimp_step_single(lat.h, C.byref(p), C.c_int(args.steps)))
        capi._check(lib.chimp_download_rho(lat.h, C.c_void_p(host_rho.data_ptr()), C.c_int(1)))
        capi._check(lib.chimp_download_vel(lat.h, C.c_void_p(host_vel.data_ptr())))
        dt = time.perf_counter() - t0
        e2e = {"value": n * args.steps / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": host_f.numel() * 8 / args.steps,
               "d2h_bytes_per_step": (host_rho.numel() + host_vel.numel()) * 8 / args.steps,
               "cycle": "upload LbField (pinned host, reference AoS) + %d steps + download rho, vel; wall clock; one untimed cycle before" % args.steps}
        assert abs(float(host_rho[1:].mean()) - 1.0) < 1e-9
    except Exception as exc:  # pragma: no cover
        e2e = {"value": None, "unit": "MLUPS", "error": str(exc)}

    line = {"metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "std_case physics (D3Q19 BGK + Guo force + half-way bounce back) on the configs[2] geometry: periodic random sphere pack %d^3, R=%d, seed 1234" % (size, size // 8),
                       "fluid_nodes": n, "porosity": n / float(size) ** 3, "index_form": args.index,
                       "l2_policy": "state 2 x %.1f GB >> 126 MB L2 (inputs larger than L2)" % (n * 152 / 1e9),
                       "irregular_tile_fraction": lat.irregular_fraction(),
                       "index_bytes_per_node": lat.index_bytes_per_node(), "setup_seconds": setup_s,
                       "mean_rho_error": mass_err},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _traffic_gb(lattice, args.index, n)[0], "traffic_unit": "GB per launch (ncu dram read+write)",
                         "traffic_source": _traffic_gb(lattice, args.index, n)[1],
                         "algorithmic_gb_per_launch": B_ALG[lattice] * n / 1e9,
                         "peak_source": peak_src, "bytes_per_node": B_ALG[lattice]},
            "e2e": e2e, "gpu_launches": int(launches2 - launches1), "clocks": _summarize_clocks(samples)}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_port(pkg)
    print(json.dumps(line), flush=True)
