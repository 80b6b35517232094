"""The BASELINE.json configurations as bench workloads: geometry, physics and decomposition of each,
and the builders that turn one into an engine context per rank (structured ingest on the device).

  std_case     configs[0] physics (D3Q19 BGK + Guo force + half-way bounce back, std_case/main.cpp:109-146) on the
               configs[2] geometry (periodic random sphere pack 512^3, porosity ~0.35); N GPUs: the SAME pack split
               into N z-slabs of equal fluid-node count (strong scaling, rank map of relperm_input.py:16-35)
  trt          the same with the TRT collision (calcOmegaBGKTRT, LBcollision.h:50-75)
  one_phase    configs[2]: std_one_phase loop body (std_one_phase/main.cpp:534-575; force switch, mass source
               attributes) with TRT on the same pack; --interior-domains adds two interior domains, which makes
               the per-step mass-change sum run (main.cpp:520-528)
  d2q9_channel configs[1]: D2Q9 SRT Poiseuille channel 8192 x 8192, one GPU
  twophase     configs[3]: colour-gradient D3Q19, two LbFields, sphere pack 384^3, split into N z-slabs
  d3q27_dense  configs[4]: dense periodic D3Q27 BGK, 512^3 nodes PER GPU (weak scaling)
"""
from __future__ import annotations

import numpy as np
import torch

from . import geometry as G

B_ALG = {"D2Q9": 144.0, "D3Q19": 304.0, "D3Q27": 432.0}

WORKLOADS = {
    "std_case": dict(lattice="D3Q19", geometry="pack", size=512, physics="single", trt=None, scaling="strong", tau=0.8,
                     force=(1e-6, 0.0, 0.0), b_alg=304.0,
                     name="std_case physics (D3Q19 BGK + Guo force + half-way bounce back) on the configs[2] geometry: periodic random sphere pack %(size)d^3, R=%(radius)d, seed 1234"),
    "trt": dict(lattice="D3Q19", geometry="pack", size=512, physics="single", trt=(0.8, 1.125), scaling="strong", tau=0.8,
                force=(1e-6, 0.0, 0.0), b_alg=304.0,
                name="D3Q19 TRT (tau_sym 0.8, Lambda 3/16) + Guo force + half-way bounce back, periodic random sphere pack %(size)d^3, R=%(radius)d, seed 1234"),
    "one_phase": dict(lattice="D3Q19", geometry="pack", size=512, physics="one_phase", trt=(0.8, 1.125), scaling="strong", tau=0.8,
                      force=(1e-6, 0.0, 0.0), b_alg=304.0,
                      name="configs[2]: std_one_phase loop body (force switch + mass-source attributes) D3Q19 TRT, periodic random sphere pack %(size)d^3, R=%(radius)d, seed 1234"),
    "d2q9_channel": dict(lattice="D2Q9", geometry="channel", size=8192, physics="single", trt=None, scaling="single", tau=0.8,
                         force=(1e-7, 0.0, 0.0), b_alg=144.0,
                         name="configs[1]: D2Q9 SRT Poiseuille channel %(size)d x %(size)d (walls at y = 0, ny-1; periodic in x)"),
    "twophase": dict(lattice="D3Q19", geometry="pack", size=384, physics="twophase", trt=None, scaling="strong", b_alg=624.0,
                     tp=(1.0, 1.0, 0.01, 1.0, 1e-5, (0.0, 0.0, 0.0)),
                     name="configs[3]: colour-gradient two-phase D3Q19 (two LbFields, flux-controlled force), periodic random sphere pack %(size)d^3, R=%(radius)d, seed 1234"),
    "d3q27_dense": dict(lattice="D3Q27", geometry="dense", size=512, physics="single", trt=None, scaling="weak", tau=0.8,
                        force=(1e-6, 0.0, 0.0), b_alg=432.0,
                        name="configs[4]: dense periodic D3Q27 BGK, %(size)d^3 nodes per GPU"),
}

PACK_POROSITY, PACK_SEED = 0.35, 1234


def describe(wl, size):
    return wl["name"] % {"size": size, "radius": size // 8}


def global_shape(wl, size, world, scaling):
    if wl["geometry"] == "channel":
        return (size, size)
    return (size, size, size * world) if scaling == "weak" else (size, size, size)


def slab_geometry(wl, ingest, gshape, z0, z1):
    """uint8 [nx, ny, z1 - z0] of the global geometry (z indices wrap periodically)"""
    if wl["geometry"] == "dense":
        return np.ones(gshape[:2] + (z1 - z0,), dtype=np.uint8)
    return ingest.sphere_pack_slab(gshape, gshape[0] / 8.0, PACK_POROSITY, PACK_SEED, z0, z1)


def full_geometry(wl, gshape):
    if wl["geometry"] == "channel":
        g = np.ones(gshape, dtype=np.uint8)
        g[:, 0] = 0
        g[:, -1] = 0
        return g
    if wl["geometry"] == "dense":
        return np.ones(gshape, dtype=np.uint8)
    return G.sphere_pack(gshape, gshape[0] / 8.0, PACK_POROSITY, PACK_SEED).astype(np.uint8)


def periodicity(wl):
    return "x" if wl["geometry"] == "channel" else "xyz"


def one_phase_attributes(cell_x, near_solid, n, interior_domains, nx):
    """host attribute arrays indexed by reference label (row 0 = dummy node): force switch 1 everywhere,
    optionally two interior domains (label 1: x < nx/4, label 2: x >= 3 nx/4).  As in the reference, a node of an
    interior domain carries a mass source only if its tag is a plain fluid tag (std_one_phase/main.cpp:337-346:
    tags < 3, i.e. no solid-interface bit)."""
    force_on = np.ones(n + 1)
    interior = np.zeros(n + 1, dtype=np.int32)
    add = np.zeros(n + 1)
    n_labels = 1
    if interior_domains:
        interior[1:][cell_x < nx // 4] = 1
        interior[1:][cell_x >= 3 * nx // 4] = 2
        add[1:] = ((interior[1:] > 0) & ~near_solid).astype(np.float64)
        n_labels = 3
    return force_on, interior, add, n_labels


def _near_solid(fluid_ext, lattice, wrap_z):
    """bool tensor like fluid_ext: the cell has a solid neighbour (any non-rest direction); z does not wrap for slabs
    (their first / last layers are halo layers that are cut away by the caller)"""
    basis = G.BASIS[lattice]
    near = torch.zeros_like(fluid_ext)
    dims = tuple(range(fluid_ext.dim()))
    for q in range(len(basis) - 1):
        c = [-int(v) for v in basis[q]]
        near |= ~torch.roll(fluid_ext, shifts=c, dims=dims)
    return near


class RankLattice:
    """one rank's engine context of a workload plus what the bench needs to know about it"""

    def __init__(self):
        self.lat = None
        self.n = 0
        self.halo_bytes = 0.0
        self.z = (0, 0)
        self.halo_mode = "none"
        self.n_global = 0
        self.slab_labels = None      # int32 tensor: reference label of each slot (N > 1)
        self.cell_index = None       # flat global cell index of the own fluid cells in label order (parity probes)


def _balanced_range(multi, wl, ingest, gshape, rank, world, device, balance):
    nz = gshape[2]
    cuts = [round(k * nz / world) for k in range(world + 1)]
    if balance and wl["geometry"] == "pack" and world > 1:
        z0, z1 = cuts[rank], cuts[rank + 1]
        own = slab_geometry(wl, ingest, gshape, z0, z1)
        # all ranks must contribute equally long vectors: pad to the thickest slab
        thick = max(cuts[k + 1] - cuts[k] for k in range(world))
        layers = np.zeros(thick, dtype=np.int64)
        layers[: z1 - z0] = own.reshape(-1, z1 - z0).sum(axis=0)
        t = torch.from_numpy(layers).to(device)
        cuts = multi.balanced_cuts(t, world, [cuts[k + 1] - cuts[k] for k in range(world)])
    return cuts


def build(pkg, ingest, multi, wl, size, rank, world, device, scaling, index_form, halo="peer", balance=True,
          interior_domains=False, keep_cells=False):
    """builds this rank's lattice of the workload (state initialised), returns a RankLattice"""
    import torch.distributed as dist
    capi = pkg.capi
    lattice = wl["lattice"]
    two = wl["physics"] == "twophase"
    n_fields = 2 if two else 1
    gshape = global_shape(wl, size, world, scaling)
    out = RankLattice()
    local = device.index if device.index is not None else 0
    if world == 1:
        geo = full_geometry(wl, gshape)
        fluid = torch.from_numpy(geo).to(device).bool()
        table, labels, n, n_pad = ingest.build_pull_table(fluid, lattice, periodicity(wl))
        lat = capi.lattice_from_device_table(lattice, n, n_pad, 0, table.data_ptr(), labels.data_ptr(), n_fields, index_form, local)
        del table, labels
        if two:
            wall_phi = torch.zeros(gshape, dtype=torch.float64)     # wettability 0.5: rho0 = rho1 at the wall
            ptable, n_extra, phi_extra = ingest.build_phi_table(fluid, wall_phi, lattice, "xyz")
            lat.set_phi_table_dev(ptable.data_ptr(), n_extra, phi_extra.data_ptr())
            x = torch.arange(gshape[0], device=device)[:, None, None].expand(gshape)
            r0 = (x < gshape[0] // 2).double()[fluid]
            rho_dev = torch.stack([r0, 1.0 - r0]).contiguous()
            lat.init_equilibrium_dev(rho_dev.data_ptr())
            del ptable, rho_dev, r0, x
        else:
            lat.init_uniform(1.0)
        if wl["physics"] == "one_phase":
            cx = (torch.nonzero(fluid.reshape(-1)).reshape(-1) // int(np.prod(gshape[1:]))).cpu().numpy()
            near = _near_solid(fluid, lattice, True)[fluid].cpu().numpy()
            fo, il, add, nl = one_phase_attributes(cx, near, n, interior_domains, gshape[0])
            scale = np.zeros(nl)
            for l in range(1, nl):
                scale[l] = 1.0 / max(1.0, float(((il == l) & (add > 0)).sum()))
            lat.set_one_phase_attributes(fo, il, add, scale, 1.0)
        if keep_cells:
            out.cell_index = torch.nonzero(fluid.reshape(-1)).reshape(-1).cpu().numpy()
        del fluid
        out.lat, out.n, out.z, out.n_global = lat, n, (0, gshape[-1]), n
        torch.cuda.empty_cache()
        return out

    assert len(gshape) == 3, "only 3-D workloads are decomposed into z-slabs"
    cuts = _balanced_range(multi, wl, ingest, gshape, rank, world, device, balance)
    z0, z1 = cuts[rank], cuts[rank + 1]
    ext = torch.from_numpy(slab_geometry(wl, ingest, gshape, z0 - 1, z1 + 1)).to(device).bool()
    wall_phi = torch.zeros(ext.shape, dtype=torch.float64, device=device) if two else None
    slab = ingest.build_slab_tables(ext, lattice, True, wall_phi)
    lat = capi.lattice_from_device_table(lattice, slab["n"], slab["n_pad"], slab["n_halo"], slab["table"].data_ptr(),
                                         slab["labels"].data_ptr(), n_fields, index_form, local)
    if halo == "peer" and not multi.peer_memory_available(local, world, device):
        halo = "nccl"   # no peer addressing between the GPUs of this node
    faces = slab["faces"]
    if two:
        lat.set_phi_table_dev(slab["ptable"].data_ptr(), slab["n_extra"], slab["phi_extra"].data_ptr())
        if halo == "peer":
            multi.attach_ring_twophase_peer(lat, slab, rank, world)
        else:
            multi.attach_ring_twophase(lat, slab, rank, world, device)
        sf = slab["scalar_faces"]
        out.halo_bytes = 8.0 * (2 * (len(faces["down"][0]) + len(faces["up"][0])) + len(sf["down"][0]) + len(sf["up"][0]))
    else:
        if halo == "peer":
            multi.attach_ring_peer(lat, slab, rank, world)
        else:
            multi.attach_ring(lat, slab, rank, world, device)
        out.halo_bytes = 8.0 * (len(faces["down"][0]) + len(faces["up"][0]))
    n = slab["n"]
    own = ext[:, :, 1:-1]
    lab = slab["labels"][:n].long() - 1
    t = torch.tensor([float(n)], dtype=torch.float64, device=device)
    dist.all_reduce(t)
    out.n_global = int(t.item())
    if two:
        x = torch.arange(gshape[0], device=device)[:, None, None].expand(own.shape)
        r0 = (x < gshape[0] // 2).double()[own][lab]
        rho_dev = torch.stack([r0, 1.0 - r0]).contiguous()
        lat.init_equilibrium_dev(rho_dev.data_ptr())
        del x, r0, rho_dev
    else:
        lat.init_uniform(1.0)
    if wl["physics"] == "one_phase":
        cx = (torch.nonzero(own.reshape(-1)).reshape(-1) // int(own.shape[1] * own.shape[2])).cpu().numpy()
        near = _near_solid(ext, lattice, False)[:, :, 1:-1][own].cpu().numpy()
        fo, il, add, nl = one_phase_attributes(cx, near, n, interior_domains, gshape[0])
        counts = torch.zeros(nl, dtype=torch.float64, device=device)
        for l in range(1, nl):
            counts[l] = float(((il == l) & (add > 0)).sum())
        dist.all_reduce(counts)
        scale = np.zeros(nl)
        scale[1:] = 1.0 / np.maximum(counts.cpu().numpy()[1:], 1.0)
        lat.set_one_phase_attributes(fo, il, add, scale, 1.0)
        if nl > 1:
            multi.attach_allreduce(lat, device)
    if keep_cells:
        # flat global cell index (C-order of the global array) of my fluid cells in label order
        nzs = z1 - z0
        flat = torch.nonzero(own.reshape(-1)).reshape(-1)
        xy, zl = flat // nzs, flat % nzs
        out.cell_index = (xy * gshape[2] + ((zl + z0) % gshape[2])).cpu().numpy()
    out.slab_labels = slab["labels"][:n].cpu().numpy()
    slab.clear()
    del ext, own, lab, wall_phi
    torch.cuda.empty_cache()
    out.lat, out.n, out.z, out.halo_mode = lat, n, (z0, z1), halo
    return out
