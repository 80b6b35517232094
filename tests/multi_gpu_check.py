"""Run under torchrun with N >= 2 GPUs (not collected by pytest; tests/test_gpu_multi.py launches it):
every rank steps its z-slab of a periodic sphere pack with NCCL halos, and also steps the whole
(undecomposed) geometry on its own GPU; rho and u of the slab must equal the corresponding nodes
of the global run bit for bit."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
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
    ok_all = True
    for lattice, nzr, steps in (("D3Q19", 40, 60), ("D3Q27", 24, 30)):
        gshape = (48, 40, nzr * world)
        geo = pkg.geometry.sphere_pack(gshape, 7.0, 0.45, 31)
        # the undecomposed run
        T, L, ng, ngp = ingest.build_pull_table(torch.from_numpy(geo).to(dev).bool(), lattice, "xyz")
        glat = capi.lattice_from_device_table(lattice, ng, ngp, 0, T.data_ptr(), L.data_ptr(), 1, capi.INDEX_COMPACT, local)
        glat.init_uniform(1.0)
        glat.step_single(steps, tau=0.8, force=(1e-5, 2e-6, -1e-6))
        grho, gvel = glat.download_moments_device_order()
        glat.close()
        # my slab with halos
        ext = ingest.sphere_pack_slab(gshape, 7.0, 0.45, 31, rank * nzr - 1, (rank + 1) * nzr + 1)
        assert np.array_equal(ext[:, :, 1:-1], geo[:, :, rank * nzr:(rank + 1) * nzr])
        slab = ingest.build_slab_tables(torch.from_numpy(ext).to(dev).bool(), lattice, True)
        lat = capi.lattice_from_device_table(lattice, slab["n"], slab["n_pad"], slab["n_halo"], slab["table"].data_ptr(),
                                             slab["labels"].data_ptr(), 1, capi.INDEX_COMPACT, local)
        if os.environ.get("CHIMP_HALO", "peer") == "peer":
            multi.attach_ring_peer(lat, slab, rank, world)
        else:
            multi.attach_ring(lat, slab, rank, world, dev)
        lat.init_uniform(1.0)
        lat.step_single(steps, tau=0.8, force=(1e-5, 2e-6, -1e-6))
        rho, vel = lat.download_moments_device_order()
        # slot -> local label -> global slot
        gl = (np.cumsum(geo.reshape(-1)) * geo.reshape(-1)).reshape(gshape)
        own = geo[:, :, rank * nzr:(rank + 1) * nzr].astype(bool)
        glab = gl[:, :, rank * nzr:(rank + 1) * nzr][own]
        lab = slab["labels"][: slab["n"]].cpu().numpy()
        gslot = glab[lab - 1] - 1
        ok = np.array_equal(rho, grho[gslot]) and np.array_equal(vel, gvel[:, gslot])
        print("rank %d %s: slab %d nodes, %d steps, bit-exact vs undecomposed run: %s" % (rank, lattice, slab["n"], steps, ok), flush=True)
        ok_all = ok_all and ok
        lat.close()
    # colour-gradient two-phase run on z-slabs (NCCL population halos + scalar halo of phi + all-reduced flux force)
    # against the undecomposed run: the only difference is the order in which the momentum sum is added up
    nzr, steps = 28, 40
    gshape = (40, 36, nzr * world)
    geo = pkg.geometry.sphere_pack(gshape, 6.0, 0.5, 7).astype(bool)
    xs = np.arange(gshape[0])[:, None, None] * np.ones(gshape)
    rho0 = (xs < gshape[0] / 2).astype(np.float64)
    wall_phi = np.where(geo, 0.0, -0.4)
    n_global = int(geo.sum())
    args = (1.0, 0.8, 0.01, 1.0, 1e-5, (0, 1e-7, 0), n_global)
    fluid = torch.from_numpy(geo).to(dev)
    T, L, ng, ngp = ingest.build_pull_table(fluid, "D3Q19", "xyz")
    glat = capi.lattice_from_device_table("D3Q19", ng, ngp, 0, T.data_ptr(), L.data_ptr(), 2, capi.INDEX_COMPACT, local)
    pt, n_extra, phi_extra = ingest.build_phi_table(fluid, torch.from_numpy(wall_phi), "D3Q19", "xyz")
    glat.set_phi_table_dev(pt.data_ptr(), n_extra, phi_extra.data_ptr())
    r0 = torch.from_numpy(rho0[geo]).to(dev)
    rho_dev = torch.stack([r0, 1.0 - r0]).contiguous()
    glat.init_equilibrium_dev(rho_dev.data_ptr())
    glat.step_twophase(steps, *args)
    grho, gvel = glat.download_moments_device_order()
    gforce = glat.last_flux_force()
    glat.close()
    idx = np.arange(rank * nzr - 1, (rank + 1) * nzr + 1) % gshape[2]
    ext, wext = geo[:, :, idx], wall_phi[:, :, idx]
    slab = ingest.build_slab_tables(torch.from_numpy(ext).to(dev), "D3Q19", True, torch.from_numpy(wext).to(dev))
    lat = capi.lattice_from_device_table("D3Q19", slab["n"], slab["n_pad"], slab["n_halo"], slab["table"].data_ptr(),
                                         slab["labels"].data_ptr(), 2, capi.INDEX_COMPACT, local)
    lat.set_phi_table_dev(slab["ptable"].data_ptr(), slab["n_extra"], slab["phi_extra"].data_ptr())
    if os.environ.get("CHIMP_HALO", "peer") == "peer":
        multi.attach_ring_twophase_peer(lat, slab, rank, world)
    else:
        multi.attach_ring_twophase(lat, slab, rank, world, dev)
    own = geo[:, :, rank * nzr:(rank + 1) * nzr]
    lab = slab["labels"][: slab["n"]].cpu().numpy()
    r0 = torch.from_numpy(rho0[:, :, rank * nzr:(rank + 1) * nzr][own][lab - 1]).to(dev)
    rho_dev = torch.stack([r0, 1.0 - r0]).contiguous()
    lat.init_equilibrium_dev(rho_dev.data_ptr())
    lat.step_twophase(steps, *args)
    rho, vel = lat.download_moments_device_order()
    gl = (np.cumsum(geo.reshape(-1)) * geo.reshape(-1)).reshape(gshape)
    gslot = gl[:, :, rank * nzr:(rank + 1) * nzr][own][lab - 1] - 1
    force = lat.last_flux_force()
    ok = (np.allclose(rho, grho[gslot], rtol=1e-12, atol=1e-15) and np.allclose(vel, gvel[:, gslot], rtol=1e-9, atol=1e-13)
          and abs(force - gforce) <= 1e-9 * abs(gforce))
    print("rank %d twophase D3Q19: slab %d nodes, %d steps, matches undecomposed run: %s (max |d rho| %.2e, max |d u| %.2e, flux force %.6e vs %.6e)"
          % (rank, slab["n"], steps, ok, np.abs(rho - grho[gslot]).max(), np.abs(vel - gvel[:, gslot]).max(), force, gforce), flush=True)
    ok_all = ok_all and ok
    lat.close()
    # the bench workloads on N z-slabs against the ORACLE PORT of the undecomposed geometry (the probe bench.py
    # runs before timing): same ingest, halo transport and step kernels as the timed runs
    bench_impl = importlib.import_module("badchimp_cpp_b200.bench_impl")
    W = importlib.import_module("badchimp_cpp_b200.workloads")
    halo = os.environ.get("CHIMP_HALO", "peer")
    for workload, interior in (("std_case", False), ("trt", False), ("one_phase", False), ("one_phase", True), ("twophase", False),
                               ("d3q27_dense", False)):
        wl = W.WORKLOADS[workload]
        try:
            res = bench_impl.parity_probe(pkg, ingest, multi, wl, workload, rank, world, dev, capi.INDEX_COMPACT, halo, interior)
            ok = True
        except SystemExit as exc:
            res, ok = str(exc), False
        if rank == 0:
            print("workload %s%s on %d ranks vs oracle port: %s" % (workload, " +interior domains" if interior else "", world, res), flush=True)
        ok_all = ok_all and ok
        dist.barrier()
    t = torch.tensor([1.0 if ok_all else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if t.item() != 1.0:
        sys.exit(1)
    if rank == 0:
        print("MULTI_GPU_CHECK_OK", flush=True)


if __name__ == "__main__":
    main()
