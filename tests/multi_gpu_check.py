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
