#!/usr/bin/env python
"""N-GPU throughput of the BASELINE.json configurations that are defined on several GPUs (run under torchrun):
  twophase : configs[3] colour gradient D3Q19, sphere pack 384^3 split into N z-slabs (strong scaling), NCCL halos
             of both LbFields + the scalar halo of phi + all-reduced flux force
  d3q27    : configs[4] dense periodic D3Q27 BGK, 512^3 nodes per GPU (weak scaling), peer-store halos
usage: torchrun --nproc-per-node N scripts/measure_multi.py [twophase] [d3q27]   (not bench lines: profiles/ + DESIGN.md)"""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402


def main():
    pkg = helpers.load_package()
    ingest = importlib.import_module("badchimp_cpp_b200.ingest")
    multi = importlib.import_module("badchimp_cpp_b200.multi")
    capi = pkg.capi
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    peak = 6468.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    what = sys.argv[1:] or ["twophase", "d3q27"]
    steps = int(os.environ.get("STEPS", "60"))

    def total(x, op=dist.ReduceOp.SUM):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        return total(e0.elapsed_time(e1), dist.ReduceOp.MAX)

    if "twophase" in what:
        size = int(os.environ.get("TWOPHASE_SIZE", "384"))
        gshape = (size, size, size)
        cuts = [round(k * size / world) for k in range(world + 1)]
        z0, z1 = cuts[rank], cuts[rank + 1]
        if os.environ.get("BALANCE", "1") == "1" and all(cuts[k + 1] - cuts[k] == cuts[1] for k in range(world)):
            # the step time is the maximum over ranks: cut the slabs at equal fluid-node counts
            own = ingest.sphere_pack_slab(gshape, size / 8.0, 0.35, 1234, z0, z1)
            layers = torch.from_numpy(own.reshape(-1, z1 - z0).sum(axis=0).astype(np.int64)).to(dev)
            cuts = multi.balanced_cuts(layers, world)
            z0, z1 = cuts[rank], cuts[rank + 1]
        ext = torch.from_numpy(ingest.sphere_pack_slab(gshape, size / 8.0, 0.35, 1234, z0 - 1, z1 + 1)).to(dev).bool()
        wall_phi = torch.zeros(ext.shape, dtype=torch.float64, device=dev)      # wettability 0.5: rho0 = rho1 at the wall
        slab = ingest.build_slab_tables(ext, "D3Q19", True, wall_phi)
        lat = capi.lattice_from_device_table("D3Q19", slab["n"], slab["n_pad"], slab["n_halo"], slab["table"].data_ptr(),
                                             slab["labels"].data_ptr(), 2, capi.INDEX_COMPACT, local)
        lat.set_phi_table_dev(slab["ptable"].data_ptr(), slab["n_extra"], slab["phi_extra"].data_ptr())
        tp_mode = os.environ.get("CHIMP_HALO", "peer")
        if tp_mode == "peer":
            multi.attach_ring_twophase_peer(lat, slab, rank, world)
        else:
            multi.attach_ring_twophase(lat, slab, rank, world, dev)
        own = ext[:, :, 1:-1]
        x = torch.arange(size, device=dev)[:, None, None].expand(own.shape)
        lab = slab["labels"][: slab["n"]].long() - 1
        r0 = (x < size // 2).double()[own][lab]
        rho_dev = torch.stack([r0, 1.0 - r0]).contiguous()
        lat.init_equilibrium_dev(rho_dev.data_ptr())
        n = slab["n"]
        n_total = int(total(n))
        halo = 8.0 * (2 * (len(slab["faces"]["down"][0]) + len(slab["faces"]["up"][0])) + len(slab["scalar_faces"]["down"][0]) + len(slab["scalar_faces"]["up"][0]))
        del slab, ext, wall_phi, rho_dev, own, x
        torch.cuda.empty_cache()
        args = (1.0, 1.0, 0.01, 1.0, 1e-5, (0, 0, 0), n_total)
        lat.step_twophase(5, *args)
        lat.synchronize()
        # the engine runs on its own streams: time with host clock around a synchronised region, max over ranks
        import time
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lat.step_twophase(steps, *args)
        lat.synchronize()
        ms = total((time.perf_counter() - t0) * 1e3, dist.ReduceOp.MAX)
        rho, _ = lat.download_moments_device_order()
        mass = total(float(rho.sum()))
        if rank == 0:
            mlups = n_total * steps / (ms * 1e-3) / 1e6
            print(json.dumps({"config": "twophase colour gradient D3Q19 sphere pack %d^3 on %d GPUs (configs[3]), z-slabs, %s" % (size, world, "peer-memory halos + mailbox sum (NVLink)" if tp_mode == "peer" else "NCCL halos + all-reduce"),
                              "n_gpus": world, "fluid_nodes": n_total, "steps": steps, "ms_per_step": ms / steps, "MLUPS": mlups,
                              "MLUPS_per_gpu": mlups / world, "B_alg": 624.0, "frac_of_measured_hbm_per_gpu": 624.0 * mlups * 1e6 / world / 1e9 / peak,
                              "halo_bytes_per_step_per_gpu": halo, "scaling": "strong", "slab_cuts": [int(x) for x in cuts],
                              "note": "host clock around %d steps, max over ranks; sum rho0 = %.6f; flux force %.3e" % (steps, mass, lat.last_flux_force())}), flush=True)
        lat.close()
    if "d3q27" in what:
        size = int(os.environ.get("D3Q27_SIZE", "512"))
        ext = torch.ones((size, size, size + 2), dtype=torch.bool, device=dev)
        slab = ingest.build_slab_tables(ext, "D3Q27", True)
        lat = capi.lattice_from_device_table("D3Q27", slab["n"], slab["n_pad"], slab["n_halo"], slab["table"].data_ptr(),
                                             slab["labels"].data_ptr(), 1, capi.INDEX_COMPACT, local)
        mode = os.environ.get("CHIMP_HALO", "peer")
        if mode == "peer":
            multi.attach_ring_peer(lat, slab, rank, world)
        else:
            multi.attach_ring(lat, slab, rank, world, dev)
        n = slab["n"]
        halo = 8.0 * (len(slab["faces"]["down"][0]) + len(slab["faces"]["up"][0]))
        del slab, ext
        torch.cuda.empty_cache()
        lat.init_uniform(1.0)
        force = (1e-6, 0.0, 0.0)
        lat.step_single(5, tau=0.8, force=force)
        lat.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        ms = total(lat.step_timed(steps // 2, tau=0.8, force=force), dist.ReduceOp.MAX)
        dist.barrier()
        rho, _ = lat.download_moments_device_order()
        n_total = int(total(n))
        if rank == 0:
            k = steps // 2
            mlups = n_total * k / (ms * 1e-3) / 1e6
            print(json.dumps({"config": "D3Q27 BGK dense periodic, %d^3 nodes per GPU on %d GPUs (configs[4]), z-slabs, %s halos" % (size, world, mode),
                              "n_gpus": world, "fluid_nodes": n_total, "steps": k, "ms_per_step": ms / k, "MLUPS": mlups, "MLUPS_per_gpu": mlups / world,
                              "B_alg": 432.0, "frac_of_measured_hbm_per_gpu": 432.0 * mlups * 1e6 / world / 1e9 / peak,
                              "halo_bytes_per_step_per_gpu": halo, "scaling": "weak",
                              "note": "CUDA events on the engine stream, max over ranks; mean rho error %.1e" % abs(rho.mean() - 1)}), flush=True)
        lat.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
