"""Measurement harness behind bench.py (see DESIGN.md "Measurement").

Default workload (`--workload std_case`): D3Q19 BGK + Guo body force + half-way bounce back
(std_case/main.cpp:109-146) in a periodic random sphere pack, edge 512, sphere radius 64, porosity ~0.35,
seed 1234 (BASELINE.json configs[2] geometry).  MLUPS = fluid nodes x steps / seconds / 1e6.
With --gpus N the SAME 512^3 pack is split into N z-slabs of equal fluid-node count (strong scaling, the
configuration the ">= 85 % parallel efficiency at 8 GPUs" target is quoted on); the weak-scaling number
(one 512^3 block per GPU) is measured in the same run and reported under "weak".
The other BASELINE.json configurations are selected with --workload (badchimp-cpp_b200/workloads.py).
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
EXTRAS_BUDGET_S = 300.0   # default one-GPU run: wall-clock budget of the secondary workloads (later ones are skipped beyond it)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _committed_traffic(workload, index):
    """DRAM bytes per node per launch from the committed ncu capture (profiles/traffic.json); used only when
    ncu cannot be run next to the bench"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["%s/%s" % (workload, index)]
        return t["dram_bytes_per_node"], t["source"]
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
    """oracle port (plain-C restatement, 1 core) on a bounded sample of the std_case workload"""
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


def _physical_cores():
    try:
        pairs = set()
        phys = core = None
        for line in open("/proc/cpuinfo"):
            if line.startswith("physical id"):
                phys = line.split(":")[1].strip()
            elif line.startswith("core id"):
                core = line.split(":")[1].strip()
            elif not line.strip():
                if phys is not None and core is not None:
                    pairs.add((phys, core))
                phys = core = None
        if pairs:
            return len(pairs)
    except Exception:
        pass
    return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref/ref_driver = unmodified
    reference headers, MPI ranks as threads) on the host cores, bounded sample of the same workload.  The
    rank count is chosen by a short timing over the candidates that divide the sample and fit the cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    pkg = helpers.load_package()
    G = pkg.geometry
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    logical = os.cpu_count() or 1
    try:
        logical = len(os.sched_getaffinity(0))
    except Exception:
        pass
    physical = min(_physical_cores(), logical)
    size = args.size or 128
    geo = G.sphere_pack((size,) * 3, size / 8.0, 0.35, 1234).astype(int)
    steps, warm = max(1, min(args.steps, 40)), max(0, min(args.warmup, 5))
    line = {"metric": "MLUPS", "unit": "MLUPS", "impl": "reference", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "std_case D3Q19 BGK+Guo+half-way BB, periodic sphere pack %d^3 porosity~0.35 (bounded sample of the 512^3 workload: the reference's loader cannot read a 512^3 file, LBvtk.h:194-201)" % size}}
    if os.path.exists(driver):
        tried = {}
        best = None
        for nranks in [p for p in (64, 32, 16, 8, 4, 2, 1) if p <= logical and size % p == 0 and size // p >= 4]:
            d = tempfile.mkdtemp(prefix="chimp_ref_")
            os.makedirs(os.path.join(d, "out"))
            lg = G.LatticeGeometry(G.z_slab_rank_map(geo, nranks) if nranks > 1 else geo, "D3Q19", "xyz")
            for t in lg.all_ranks():
                t.write_vtklb(os.path.join(d, "tmp%d.vtklb" % t.my_rank), {"init_rho": np.ones(geo.shape)})
            cmd = [driver, "--case", "std_case", "--lattice", "D3Q19", "--dir", d, "--out", os.path.join(d, "out"),
                   "--nranks", str(nranks), "--no-tables", "--no-f", "--time", "--tau", "0.8", "--force", "1e-6,0,0"]
            r = subprocess.run(cmd + ["--steps", "4"], capture_output=True, text=True)
            try:
                probe = json.loads(r.stdout.strip().splitlines()[-1])["mlups"]
            except Exception:
                probe = 0.0
            tried[nranks] = round(probe, 2)
            if best is None or probe > best[0]:
                if best is not None:
                    import shutil
                    shutil.rmtree(best[2], ignore_errors=True)
                best = (probe, nranks, d, cmd)
            else:
                import shutil
                shutil.rmtree(d, ignore_errors=True)
            if nranks <= physical // 2 and best[1] > nranks:
                break   # fewer ranks than half the cores only gets slower
        _, nranks, d, cmd = best
        if warm:
            subprocess.run(cmd + ["--steps", str(warm)], capture_output=True, text=True)
        r = subprocess.run(cmd + ["--steps", str(steps)], capture_output=True, text=True)
        res = json.loads(r.stdout.strip().splitlines()[-1])
        value, kind = res["mlups"], "reference"
        ms = res["loop_seconds"] / steps * 1e3
        sample = ("oracle/_ref/ref_driver (unmodified reference headers), %d MPI ranks as threads (best of a 4-step probe over rank counts %s), "
                  "%d^3 pack, %d fluid nodes, %d steps; host has %d physical cores / %d hardware threads available"
                  % (nranks, json.dumps(tried), size, res["fluid_nodes"], steps, physical, logical))
        import shutil
        shutil.rmtree(d, ignore_errors=True)
    else:
        base = cpu_baseline_port(pkg, size=min(size, 96))
        value, kind, sample, nranks = base["value"], "port", base["sample"], 1
        ms = None
    line.update({"value": value, "ms_per_step": ms,
                 "cpu_baseline": {"value": value, "unit": "MLUPS", "cores": nranks, "physical_cores": physical, "kind": kind, "sample": sample},
                 "e2e": {"value": value, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
class _Runner:
    """stepping of one rank's lattice for a workload: the same calls for warm-up, timing and the e2e cycle"""

    def __init__(self, rl, wl):
        self.rl, self.wl, self.lat = rl, wl, rl.lat

    def _tp(self):
        tau0, tau1, sigma, beta, momx, force = self.wl["tp"]
        return (tau0, tau1, sigma, beta, momx, force, self.rl.n_global)

    def step(self, k):
        if self.wl["physics"] == "twophase":
            self.lat.step_twophase(k, *self._tp())
        else:
            self.lat.step_single(k, tau=self.wl["tau"], force=self.wl["force"], trt=self.wl["trt"])

    def timed(self, k):
        if self.wl["physics"] == "twophase":
            return self.lat.step_twophase_timed(k, *self._tp())
        return self.lat.step_timed(k, tau=self.wl["tau"], force=self.wl["force"], trt=self.wl["trt"])


def parity_probe(pkg, ingest, multi, wl, workload, rank, world, device, index_form, halo, interior_domains, steps=10, size=None,
                 skip_mask=False):
    """A small case of the same workload on the same code path (structured ingest, N z-slabs, the same halo
    transport and step kernels) checked against the oracle port of the UNDECOMPOSED geometry on rank 0's host,
    before anything is timed: after one step (the per-step bar: bit-exact for the single-field kernels, <= 1e-12
    relative where a global sum -- two-phase flux controller, mass-change source -- is added up in another order)
    and after `steps` steps (bit-exact, resp. <= 1e-9: rounding differences of the sums are amplified where the
    recolouring subtracts nearly equal numbers)."""
    import torch
    from . import workloads as W
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    scaling = "weak" if wl["scaling"] == "weak" else "strong"
    if size is None:
        size = {"pack": 64, "dense": 24, "channel": 96}[wl["geometry"]]
        if wl["geometry"] == "pack" and scaling == "strong" and size // world < 2:
            size = 4 * world
    rl = W.build(pkg, ingest, multi, wl, size, rank, world, device, scaling, index_form, halo=halo, balance=True,
                 interior_domains=interior_domains, keep_cells=True)
    if skip_mask:
        rl.lat.set_index_skip_mask(True)
    run = _Runner(rl, wl)
    stages = [1, steps]
    payloads, done = [], 0
    for upto in stages:
        run.step(upto - done)
        done = upto
        payloads.append(rl.lat.download()[1:].copy())     # [n, nFields, nQ] rows by the rank's own labels
    mode = rl.lat.peer_mode()
    rl.lat.close()
    payload = (rl.cell_index, payloads)
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(payload, gathered, dst=0)
    else:
        gathered = [payload]
    result = None
    if rank == 0:
        import port
        G = pkg.geometry
        gshape = W.global_shape(wl, size, world, scaling)
        geo = W.full_geometry(wl, gshape).astype(int)
        lg = G.LatticeGeometry(geo, wl["lattice"], W.periodicity(wl))
        tab = lg.all_ranks()[0]
        bulk = tab.bulk_nodes()
        lid = G.LATTICE_ID[wl["lattice"]]
        ones = np.ones(geo.shape)
        if wl["physics"] == "twophase":
            x = np.arange(gshape[0])[:, None, None] * ones
            rho0 = (x < gshape[0] // 2).astype(np.float64)
            setup = pkg.cases.two_phase_setup(lg, [tab], rho0, 1.0 - rho0, 0.5 * (geo == 0))[0]
            pr = port.PortRank(lid, tab.neigh, bulk, 2, tab.halfway_bb(bulk))
            pr.f[:] = setup["f0"]
            pr.rho[:] = setup["rho"]
            tau0, tau1, sigma, beta, momx, force = wl["tp"]
            advance = lambda k: pr.step_twophase(k, setup["solid_bnd"], tau0, tau1, sigma, beta, momx, force, len(bulk))
        elif wl["physics"] == "one_phase":
            fluid = geo.astype(bool)
            near = np.zeros(geo.shape, bool)
            for c in G.BASIS[wl["lattice"]][:-1]:
                near |= np.roll(~fluid, shift=tuple(-int(v) for v in c), axis=(0, 1, 2))
            tags = fluid * (1 + 8 * near)
            interior = np.zeros(geo.shape, dtype=int)
            if interior_domains:
                xs = np.arange(gshape[0])[:, None, None] * np.ones(geo.shape, dtype=int)
                interior[xs < gshape[0] // 4] = 1
                interior[xs >= 3 * gshape[0] // 4] = 2
                interior *= fluid
            s = pkg.cases.one_phase_setup(lg, [tab], {"nodetags": tags, "force": np.ones(geo.shape, dtype=int), "interior_domains": interior})[0]
            pr = port.PortRank(lid, tab.neigh, bulk, 1, None)
            pr.f[:] = s["f0"]
            pr.set_one_phase(s["force_on"], s["interior"], s["add_source"], s["scale"], s["solid_links"], s["press_links"], s["fluid_links"], 1.0)
            advance = lambda k: pr.step_one_phase(k, tau=wl["tau"], force=wl["force"], trt=wl["trt"])
        else:
            pr = port.PortRank(lid, tab.neigh, bulk, 1, tab.halfway_bb(tab.fluid_bnd_nodes()))
            pr.f[:] = pkg.cases.std_case_initial_state(tab, ones)[0]
            advance = lambda k: pr.step_std_case(k, tau=wl["tau"], force=wl["force"], trt=wl["trt"])
        glabel = (np.cumsum(geo.reshape(-1)) * geo.reshape(-1))
        # relative to the population itself, with a floor of 1e-3 of the smallest lattice weight: the second fluid's
        # populations decay to exactly zero away from an interface, and the relative error of a 1e-15 number says nothing
        floor = 1e-3 * float(pkg.cases.lattice_weights(wl["lattice"]).min())
        rel, absd, exact, done = [], [], [], 0
        for si, upto in enumerate(stages):
            advance(upto - done)
            done = upto
            r, a, ex, checked = 0.0, 0.0, True, 0
            for cells, frs in gathered:
                want, fr = pr.f[glabel[cells]], frs[si]
                diff = np.abs(fr - want)
                ex = ex and bool(np.array_equal(fr, want))
                r = max(r, float((diff / np.maximum(np.abs(want), floor)).max()))
                a = max(a, float(diff.max()))
                checked += len(cells)
            assert checked == len(bulk), "parity probe: %d of %d nodes gathered" % (checked, len(bulk))
            rel.append(r); absd.append(a); exact.append(ex)
        nf, nq = payloads[0].shape[1:]
        result = {"against": "oracle port (oracle/lb_port.c) of the undecomposed geometry, rank 0 host", "ranks": world,
                  "case": "%s%s, %s" % (workload, " + interior domains" if interior_domains else "", "x".join(str(v) for v in gshape)),
                  "checked_nodes": checked, "populations": int(checked * nf * nq),
                  "max_rel_f": rel[0], "max_abs_f": absd[0], "bit_exact": exact[0], "steps": 1,
                  "after_%d_steps" % steps: {"max_rel_f": rel[1], "max_abs_f": absd[1], "bit_exact": exact[1]},
                  "halo_transport": rl.halo_mode + (" (fused into the step kernel)" if mode[0] == 2 else ""), "rel_floor": floor}
        if wl["physics"] == "single":
            ok = exact[0] and exact[1]
        else:
            ok = rel[0] <= 1e-12 and rel[1] <= 1e-9
        if not ok:
            raise SystemExit("parity probe FAILED before timing: " + json.dumps(result))
    return result


def _ncu_traffic(args):
    """DRAM bytes of one step launch measured by ncu next to the bench: the same workload is rebuilt in a child
    process under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`, one launch captured after a few
    untimed ones.  Nothing measured under ncu enters a timing."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    bench = os.path.join(ROOT, "bench.py")
    kern = {"twophase": "regex:twoPhaseCollideKernel|phaseMomentsKernel"}.get(args.workload, "regex:collideStreamKernel")
    # the child steps two at a time: launch 0 of a pair is a step without the moment output (like all timed steps but the last)
    count = "2" if args.workload == "twophase" else "1"
    skip = "4" if args.workload == "twophase" else "2"
    cmd = ([ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none", "--print-units", "base", "-k", kern,
           "--launch-skip", skip, "--launch-count", count, "--csv", sys.executable, bench, "--traffic-probe", "--workload", args.workload,
           "--index", args.index, "--size", str(args.size or 0), "--skip-mask", "on" if getattr(args, "skip_mask_selected", False) else "off"] +
          (["--interior-domains"] if args.interior_domains else []))
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
        read = write = 0.0
        found = False
        for line in r.stdout.splitlines():
            cols = [c.strip('"') for c in line.split('","')]
            if len(cols) < 3:
                continue
            for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                if name in cols:
                    k = cols.index(name)
                    unit, val = cols[k + 1], float(cols[k + 2].replace(",", ""))
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
                    if name.endswith("read.sum"):
                        read += val * scale
                    else:
                        write += val * scale
                    found = True
        if not found:
            return None
        return {"read": read, "write": write}
    except Exception:
        return None


def run_traffic_probe(args):
    """child of _ncu_traffic: builds the workload and launches a few steps (ncu captures one)"""
    import importlib
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    pkg = helpers.load_package()
    ingest = importlib.import_module("badchimp_cpp_b200.ingest")
    multi = importlib.import_module("badchimp_cpp_b200.multi")
    W = importlib.import_module("badchimp_cpp_b200.workloads")
    wl = W.WORKLOADS[args.workload]
    torch.cuda.set_device(0)
    index_form = pkg.capi.INDEX_COMPACT if args.index == "compact" else pkg.capi.INDEX_TABLE
    rl = W.build(pkg, ingest, multi, wl, args.size or wl["size"], 0, 1, torch.device("cuda", 0), "strong", index_form,
                 interior_domains=args.interior_domains)
    if args.skip_mask == "on":
        rl.lat.set_index_skip_mask(True)
    run = _Runner(rl, wl)
    for _ in range(3):
        run.step(2)     # per call: one step without and one with the moment output (rho, vel) of the last step
    rl.lat.synchronize()


def _e2e_cycle(pkg, rl, wl, run, steps, total, barrier):
    """the same metric through the C-ABI with HOST buffers: upload the LbField(s) in reference AoS layout from
    pinned memory, K steps, download rho and vel (and phi) -- one write interval of the reference main"""
    import torch
    capi = pkg.capi
    lib = capi.lib()
    lat = rl.lat
    n, nq, nf, nd = rl.n, lat.nq, lat.n_fields, lat.nd
    host_f = torch.empty((n + 1, nf, nq), dtype=torch.float64, pin_memory=True)
    w = torch.from_numpy(pkg.cases.lattice_weights(wl["lattice"]))
    host_f[:] = w[None, None, :] * (0.5 if nf == 2 else 1.0)
    host_rho = torch.empty((n + 1, nf), dtype=torch.float64, pin_memory=True)
    host_vel = torch.empty((n + 1, nd), dtype=torch.float64, pin_memory=True)

    def cycle(k):
        capi._check(lib.chimp_upload_lbfield(lat.h, C.c_void_p(host_f.data_ptr())))
        # N ranks: an upload is a collective state change -- a neighbour that steps on while this rank still uploads
        # would have its first halo stores overwritten by the upload (the barrier is inside the timed region)
        barrier()
        run.step(k)
        capi._check(lib.chimp_download_rho(lat.h, C.c_void_p(host_rho.data_ptr()), C.c_int(nf)))
        capi._check(lib.chimp_download_vel(lat.h, C.c_void_p(host_vel.data_ptr())))

    cycle(1)   # one untimed cycle first (first touch of the pinned pages by the copy engines), like the warm-up steps
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cycle(steps)
    dt = total(time.perf_counter() - t0, "max")
    n_total = total(n, "sum")
    # mass is conserved over the whole lattice, not per slab: the check uses the sum over all ranks.  Nothing in this
    # function may fail on one rank only -- the other ranks would wait in the next collective.
    rho_mean = total(float(host_rho[1:].sum()), "sum") / n_total
    return {"value": n_total * steps / dt / 1e6, "unit": "MLUPS",
            "h2d_bytes_per_step": total(host_f.numel() * 8, "sum") / steps,
            "d2h_bytes_per_step": total((host_rho.numel() + host_vel.numel()) * 8, "sum") / steps,
            "mean_rho_error": abs(rho_mean - 1.0),
            "cycle": "per rank: upload LbField (pinned host, reference AoS) [+ barrier over the ranks when N > 1] + %d steps + download rho, vel; wall clock, max over ranks; one untimed cycle before" % steps}


def _e2e_from_init_rho(pkg, rl, wl, run, steps, device):
    """Secondary end-to-end figure (one GPU, single-field workloads): the data flow of the reference main itself -- the
    host holds the ScalarField init_rho (std_case/main.cpp:62-69), the populations are formed from it on the device
    (f = w_q rho, :92-96 / initiateLbField), K steps, rho and vel come back for output (:149-151).  Only 8 bytes per node
    cross the bus on the way in, instead of the whole LbField of the primary e2e cycle.  Rank-local, no collectives."""
    import torch
    capi = pkg.capi
    lib = capi.lib()
    lat = rl.lat
    n, nd = rl.n, lat.nd
    host_init = torch.ones((n + 1,), dtype=torch.float64, pin_memory=True)     # rows by reference label, row 0 = dummy node
    host_rho = torch.empty((n + 1, 1), dtype=torch.float64, pin_memory=True)
    host_vel = torch.empty((n + 1, nd), dtype=torch.float64, pin_memory=True)
    dev_rho = torch.empty((n,), dtype=torch.float64, device=device)

    def cycle(k):
        # single-rank ingest lattices keep the reference's label order on the device: slot = label - 1
        dev_rho.copy_(host_init[1:], non_blocking=True)
        torch.cuda.synchronize()
        lat.init_equilibrium_dev(dev_rho.data_ptr())
        run.step(k)
        capi._check(lib.chimp_download_rho(lat.h, C.c_void_p(host_rho.data_ptr()), C.c_int(1)))
        capi._check(lib.chimp_download_vel(lat.h, C.c_void_p(host_vel.data_ptr())))

    cycle(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cycle(steps)
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": host_init.numel() * 8 / steps,
            "d2h_bytes_per_step": (host_rho.numel() + host_vel.numel()) * 8 / steps,
            "mean_rho_error": abs(float(host_rho[1:].mean()) - 1.0),
            "cycle": "upload init_rho (pinned host ScalarField) + f = w_q rho on the device + %d steps + download rho, vel; wall clock; one untimed cycle before" % steps}


def _try_skip_mask(pkg, ingest, multi, rl, wl, run, args, workload, device, index_form):
    """One GPU, compact index, plain single-field step: the skip-mask kernel form (chimp_set_index_skip_mask) reads fewer
    index bytes behind one more dependent load, so which form is faster is measured -- a few untimed steps of each after
    the warm-up -- and the winner has to pass its own parity probe before it is timed.  Rank-local; everything is
    reported (config.index_skip_mask); a failure costs the form, not the line."""
    if wl["physics"] != "single" or args.index != "compact" or args.skip_mask == "off":
        return None
    info = {"skipped_word_share": rl.lat.index_skipped_word_fraction(), "mode": args.skip_mask}
    try:
        if args.skip_mask == "auto":
            trial = max(5, min(args.steps, 20))
            for _ in range(2):          # the second round counts: both forms have run once by then
                rl.lat.set_index_skip_mask(False)
                t_plain = run.timed(trial) / trial
                rl.lat.set_index_skip_mask(True)
                t_mask = run.timed(trial) / trial
            info["trial_ms_per_step"] = {"plain": t_plain, "skip_mask": t_mask, "steps_each": trial}
            selected = t_mask < 0.995 * t_plain     # the default form stays unless the other one is measurably faster
        else:
            selected = True
        if selected:
            info["parity"] = parity_probe(pkg, ingest, multi, wl, workload, 0, 1, device, index_form, args.halo, False, skip_mask=True)
        rl.lat.set_index_skip_mask(selected)
        info["selected"] = "skip_mask" if selected else "plain"
    except BaseException as exc:
        rl.lat.set_index_skip_mask(False)
        info["selected"] = "plain"
        info["error"] = str(exc)[:300]
    return info


def _measure(pkg, rl, wl, args, total, barrier, rank, sample_clocks, choose_form=None):
    """warm-up, K timed steps (CUDA events on the engine's stream, max over ranks), clocks under load.  choose_form: called
    between the warm-up and the timed region (one GPU: trial of the index forms, untimed)"""
    import torch
    capi = pkg.capi
    run = _Runner(rl, wl)
    run.step(args.warmup)
    rl.lat.synchronize()
    form = choose_form(run) if choose_form else None
    samples, stop = [], threading.Event()
    th = threading.Thread(target=_clock_sampler, args=(stop, samples), daemon=True)
    if sample_clocks and rank == 0:
        th.start()
    barrier()
    torch.cuda.synchronize()
    l1 = capi.lib().chimp_launch_count()
    ms = run.timed(args.steps)
    rl.lat.synchronize()
    l2 = capi.lib().chimp_launch_count()
    barrier()
    torch.cuda.synchronize()
    ms_max = total(ms, "max")
    if sample_clocks:
        # keep all ranks busy for about two more seconds so that the sampler sees clocks under load; the number of
        # extra steps follows from the all-reduced step time, so every rank steps equally often (peer halos need that)
        extra = int(min(50000, max(100, 2000.0 / max(ms_max / args.steps, 1e-3))))
        run.step(extra)
        rl.lat.synchronize()
        stop.set()
        if rank == 0:
            th.join()
    rho, _ = rl.lat.download_moments_device_order()
    rho_sum = rho.sum() if rl.lat.n_fields == 1 else None
    if rl.lat.n_fields == 2:
        # two fields: the engine keeps rho0 in row 0 and rho1 behind it; total density is conserved
        r2 = np.zeros((rl.n + 1, 2))
        rl.lat.download_rho(r2)
        rho_sum = r2[1:].sum()
    mass_err = abs(total(float(rho_sum), "sum") / total(rl.n, "sum") - 1.0)
    return dict(ms=ms, ms_max=ms_max, launches=int(l2 - l1), samples=samples, mass_err=mass_err, run=run, form=form)


def _extra_workloads(pkg, ingest, multi, W, main_rl, args, device, index_form, peak):
    """One GPU, default workload only: the other BASELINE.json configurations timed in the same run (secondary numbers
    under "other_workloads": same K / W, CUDA events on the engine's stream, state larger than L2), each with its own
    parity probe against the oracle port.  Everything here is rank-local and guarded: a failure costs that entry only.
    The main lattice is reused for the TRT collision (same geometry) and closed afterwards."""
    import torch
    out = []
    t_begin = time.perf_counter()

    def timed_entry(name, rl, wl, extra=None):
        run = _Runner(rl, wl)
        run.step(max(args.warmup, 3))
        rl.lat.synchronize()
        form = None if rl is main_rl else _try_skip_mask(pkg, ingest, multi, rl, wl, run, args, name.split("+")[0], device, index_form)
        ms = run.timed(args.steps)
        rl.lat.synchronize()
        n = rl.n
        achieved = wl["b_alg"] * n / (ms / args.steps * 1e-3) / 1e9
        e = {"workload_key": name, "workload": W.describe(wl, args.size or wl["size"]), "fluid_nodes": n, "steps": args.steps,
             "ms_per_step": ms / args.steps, "value": n * args.steps / (ms * 1e-3) / 1e6, "unit": "MLUPS",
             "roofline": {"bytes_per_node": wl["b_alg"], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak},
             "index_bytes_per_node": rl.lat.index_bytes_per_node(), "irregular_tile_fraction": rl.lat.irregular_fraction()}
        if form:
            e["index_skip_mask"] = form
        if rl.lat.n_fields == 2:
            e["phi_index_bytes_per_node"] = rl.lat.phi_index_bytes_per_node()
        if wl["physics"] == "one_phase":
            e["attribute_bytes_per_node"] = rl.lat.one_phase_attribute_bytes_per_node()
        if extra:
            e.update(extra)
        return e

    def guarded(name, fn):
        if time.perf_counter() - t_begin > EXTRAS_BUDGET_S:
            out.append({"workload_key": name, "skipped": "time budget of the default run used up"})
            return
        try:
            out.append(fn())
        except Exception as exc:  # pragma: no cover
            out.append({"workload_key": name, "error": str(exc)[:300]})
        try:
            torch.cuda.empty_cache()
        except Exception:  # pragma: no cover
            pass

    def probe(name, wl, interior=False):
        try:
            return parity_probe(pkg, ingest, multi, wl, name, 0, 1, device, index_form, args.halo, interior)
        except BaseException as exc:  # a failed probe is reported, not fatal, for a secondary entry
            return {"error": str(exc)[:300]}

    # TRT on the lattice of the main run
    def trt():
        wl = W.WORKLOADS["trt"]
        main_rl.lat.init_uniform(1.0)
        return timed_entry("trt", main_rl, wl, {"parity": probe("trt", wl)})
    guarded("trt", trt)
    try:
        main_rl.lat.close()
        torch.cuda.empty_cache()
    except Exception:  # pragma: no cover
        pass

    def fresh(name, interior=False, ab=None):
        """ab = (environment switch, key): the same workload is timed once more with that switch set to 0 -- the form
        the default replaced (attribute arrays instead of the packed word, the full phi table instead of its derived
        form) -- so that the line carries a measured A/B of the change"""
        def fn():
            wl = W.WORKLOADS[name]
            par = probe(name, wl, interior)
            rl = W.build(pkg, ingest, multi, wl, wl["size"], 0, 1, device, "strong", index_form, interior_domains=interior)
            try:
                e = timed_entry(name + ("+interior_domains" if interior else ""), rl, wl, {"parity": par})
                rho, _ = rl.lat.download_moments_device_order()
                if rl.lat.n_fields == 1:
                    e["mean_rho_error"] = abs(float(rho.mean()) - 1.0)
            finally:
                rl.lat.close()
            if ab and time.perf_counter() - t_begin < 0.8 * EXTRAS_BUDGET_S:
                var, key = ab
                saved = os.environ.get(var)
                os.environ[var] = "0"      # read when a lattice is created
                try:
                    torch.cuda.empty_cache()
                    rl2 = W.build(pkg, ingest, multi, wl, wl["size"], 0, 1, device, "strong", index_form, interior_domains=interior)
                    try:
                        e2 = timed_entry(name, rl2, wl)
                        e[key] = {k: e2[k] for k in ("value", "ms_per_step", "phi_index_bytes_per_node", "attribute_bytes_per_node") if k in e2}
                        e[key]["switch"] = var + "=0"
                    finally:
                        rl2.lat.close()
                except Exception as exc:  # pragma: no cover -- the comparison is optional
                    e[key] = {"error": str(exc)[:200]}
                finally:
                    if saved is None:
                        os.environ.pop(var, None)
                    else:
                        os.environ[var] = saved
            return e
        return fn
    guarded("one_phase", fresh("one_phase", ab=("CHIMP_ATTR_PACKED", "with_attribute_arrays")))
    guarded("one_phase+interior_domains", fresh("one_phase", True))
    guarded("twophase", fresh("twophase", ab=("CHIMP_PHI_DERIVED", "with_phi_table")))
    guarded("d2q9_channel", fresh("d2q9_channel"))
    guarded("d3q27_dense", fresh("d3q27_dense"))
    return out


def run_b200(args):
    import importlib
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    pkg = helpers.load_package()
    ingest = importlib.import_module("badchimp_cpp_b200.ingest")
    multi = importlib.import_module("badchimp_cpp_b200.multi")
    W = importlib.import_module("badchimp_cpp_b200.workloads")
    capi = pkg.capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        import datetime
        # a rank that dies must take the job down quickly, not after the default 10-minute watchdog (5 minutes leave room for a slow first import on a fresh box)
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=300))
    wl = W.WORKLOADS[args.workload]
    if wl["scaling"] == "single" and world > 1:
        raise SystemExit("workload %s is defined on one GPU" % args.workload)
    scaling = args.scaling or ("strong" if wl["scaling"] == "single" else wl["scaling"])
    size = args.size or wl["size"]
    index_form = capi.INDEX_COMPACT if args.index == "compact" else capi.INDEX_TABLE

    def total(x, op="sum"):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()

    # 1. parity of this code path at this rank count, before anything is timed
    parity = None
    if not args.no_parity:
        parity = parity_probe(pkg, ingest, multi, wl, args.workload, rank, world, device, index_form, args.halo, args.interior_domains)
        barrier()

    # 2. the timed workload
    t_setup = time.perf_counter()
    rl = W.build(pkg, ingest, multi, wl, size, rank, world, device, scaling, index_form, halo=args.halo, balance=args.balance,
                 interior_domains=args.interior_domains)
    setup_s = time.perf_counter() - t_setup
    def choose_form(run):
        if world != 1:
            return None
        info = _try_skip_mask(pkg, ingest, multi, rl, wl, run, args, args.workload, device, index_form)
        args.skip_mask_selected = bool(info) and info.get("selected") == "skip_mask"
        return info

    m = _measure(pkg, rl, wl, args, total, barrier, rank, True, choose_form)
    n_total = total(rl.n, "sum")
    per_rank = None
    if world > 1:
        pr = torch.zeros(world, 3, dtype=torch.float64, device=device)
        pr[rank] = torch.tensor([float(rl.n), float(m["ms"]) / args.steps, float(rl.z[1] - rl.z[0])], dtype=torch.float64, device=device)
        dist.all_reduce(pr)
        per_rank = pr.cpu().numpy()
    # no try/except here: an exception caught on one rank only would leave the others waiting in a collective
    e2e = _e2e_cycle(pkg, rl, wl, m["run"], args.steps, total, barrier)
    e2e_rho = None
    if world == 1 and wl["physics"] in ("single", "one_phase"):
        try:    # rank-local and secondary: a failure here must not cost the line
            e2e_rho = _e2e_from_init_rho(pkg, rl, wl, m["run"], args.steps, device)
        except Exception as exc:  # pragma: no cover
            e2e_rho = {"value": None, "unit": "MLUPS", "error": str(exc)}
    irregular, index_bytes = rl.lat.irregular_fraction(), rl.lat.index_bytes_per_node()
    phi_index_bytes = rl.lat.phi_index_bytes_per_node() if rl.lat.n_fields == 2 else None
    attr_bytes = rl.lat.one_phase_attribute_bytes_per_node() if wl["physics"] == "one_phase" else None
    halo_bytes, halo_mode = rl.halo_bytes, rl.halo_mode
    others = None
    if world == 1 and args.workload == "std_case" and not args.size and not args.no_extra_workloads:
        try:
            others = _extra_workloads(pkg, ingest, multi, W, rl, args, device, index_form, _peaks()[0])   # closes rl.lat
        except Exception as exc:  # pragma: no cover -- secondary numbers must never cost the line
            others = [{"error": str(exc)[:300]}]
    else:
        rl.lat.close()
    del rl
    torch.cuda.empty_cache()

    # 3. N > 1, strong scaling: the weak-scaling number of the same physics in the same run (secondary)
    weak = None
    if world > 1 and scaling == "strong" and not args.no_weak and wl["geometry"] in ("pack", "dense"):
        rw = W.build(pkg, ingest, multi, wl, size, rank, world, device, "weak", index_form, halo=args.halo, balance=args.balance,
                     interior_domains=args.interior_domains)
        mw = _measure(pkg, rw, wl, args, total, barrier, rank, False)
        nw = total(rw.n, "sum")
        weak = {"value": nw * args.steps / (mw["ms_max"] * 1e-3) / 1e6, "unit": "MLUPS", "ms_per_step": mw["ms_max"] / args.steps,
                "fluid_nodes": nw, "workload": "one %d^3 block per GPU (%dx%dx%d)" % (size, size, size, size * world),
                "mean_rho_error": mw["mass_err"]}
        rw.lat.close()
        del rw

    if rank == 0:
        peak, peak_src = _peaks()
        kernel_ms = m["ms_max"] / args.steps
        b_alg = wl["b_alg"]
        achieved = b_alg * n_total / world / (kernel_ms * 1e-3) / 1e9
        traffic = None
        traffic_src = None
        if world == 1 and not args.no_traffic:
            t = _ncu_traffic(args)
            if t:
                traffic = (t["read"] + t["write"]) / 1e9
                traffic_src = "ncu dram__bytes_read.sum + dram__bytes_write.sum of the step launch(es) of one step, captured by a child run of this workload next to the bench (read %.3f GB, written %.3f GB)" % (t["read"] / 1e9, t["write"] / 1e9)
        if traffic is None:
            per_node, src = _committed_traffic(args.workload, args.index)
            if per_node:
                traffic, traffic_src = per_node * n_total / world / 1e9, "committed capture: " + src
        gshape = W.global_shape(wl, size, world, scaling)
        config = {"workload": W.describe(wl, size) + ("; %d GPUs: the same %s lattice split into %d z-slabs" % (world, "x".join(map(str, gshape)), world) if world > 1 and scaling == "strong" else "") +
                  ("; one block per GPU, %s in total" % "x".join(map(str, gshape)) if world > 1 and scaling == "weak" else ""),
                  "workload_key": args.workload, "fluid_nodes": n_total, "porosity": n_total / float(np.prod(gshape)), "index_form": args.index,
                  "l2_policy": "state per GPU 2 x %.2f GB >> 126 MB L2 (inputs larger than L2)" % (n_total / world * b_alg / 2 / 1e9),
                  "irregular_tile_fraction": irregular, "index_bytes_per_node": index_bytes, "setup_seconds": setup_s,
                  "mean_rho_error": m["mass_err"]}
        if m.get("form"):
            config["index_skip_mask"] = m["form"]
        if phi_index_bytes is not None:
            config["phi_index_bytes_per_node"] = phi_index_bytes
        if attr_bytes is not None:
            config["attribute_bytes_per_node"] = attr_bytes
        if args.interior_domains:
            config["interior_domains"] = "two interior domains with mass sources: per-step mass-change sum active"
        if world > 1:
            config.update({"parallelism": "z-slab x%d, one process per GPU" % world,
                           "halo_bytes_per_step_per_gpu": halo_bytes,
                           "halo_transport": ("stores into the neighbour GPU's halo slots over NVLink from inside the step kernel, arrival counters (CUDA IPC)" if halo_mode == "peer" else "NCCL send/recv (torch.distributed) between pack and unpack kernels"),
                           "slabs": "balanced by fluid-node count" if args.balance else "equal thickness",
                           "nodes_per_rank": [int(x) for x in per_rank[:, 0]], "ms_per_step_per_rank": [round(float(x), 4) for x in per_rank[:, 1]],
                           "slab_thickness_per_rank": [int(x) for x in per_rank[:, 2]]})
        line = {"metric": "MLUPS", "value": n_total * args.steps / (m["ms_max"] * 1e-3) / 1e6, "unit": "MLUPS", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": kernel_ms, "higher_is_better": True,
                "scaling": scaling if world > 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "traffic_unit": "GB per step per GPU (DRAM read + write)", "traffic_source": traffic_src,
                             "algorithmic_gb_per_launch": b_alg * n_total / world / 1e9, "peak_source": peak_src, "bytes_per_node": b_alg,
                             "per": "GPU (mean)" if world > 1 else "GPU", "frac_of_nominal_8TBs": achieved / 8000.0},
                "e2e": e2e, "gpu_launches": m["launches"], "clocks": _summarize_clocks(m["samples"]), "parity": parity}
        if e2e_rho:
            line["e2e_from_init_rho"] = e2e_rho
        if others:
            line["other_workloads"] = others
        if weak:
            line["weak"] = weak
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_port(pkg)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
