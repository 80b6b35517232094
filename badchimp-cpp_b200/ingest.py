"""Structured geometry ingest on the device: voxel array -> pull table, without the
vtklb.py -> ASCII -> LBvtk detour (the reference loader cannot read
files > 2 GiB, LBvtk.h:194-201, so 512^3 cases are unreachable through it).

The numbering is the reference's: own fluid nodes are labelled 1..N in C-order of geo[x,y,z]
(vtklb.py:92-94), device slot i = label - 1.  For std_case semantics (half-way bounce back on
every fluid node that touches a solid, LBhalfwaybb.h:37-63) the pull table is
    T[q][i] = label(pos_i - c_q) - 1      if that cell is fluid
            = -1                          otherwise (own reversed slot).
tests/test_ingest.py checks it against the generic host builder (chimp_build_host) on the
golden geometries.  torch is used as device-memory plumbing only (cumsum / roll / gather).
"""
from __future__ import annotations

import numpy as np
import torch

from . import geometry as G


def build_pull_table(fluid: torch.Tensor, lattice: str, periodic: str = "xyz"):
    """fluid: bool/uint8 tensor [nx, ny(, nz)] (True = fluid), single rank.
    Returns (table int32 [nQ, n_pad], labels int32 [n_pad], n, n_pad) on fluid.device.
    Non-periodic axes are treated as closed (cells outside the array are solid)."""
    basis = G.BASIS[lattice]
    nq, nd = basis.shape
    assert fluid.dim() == nd
    fluid = fluid.bool()
    flat = fluid.reshape(-1)
    n = int(flat.sum().item())
    n_pad = ((n + 31) // 32) * 32
    label = (torch.cumsum(flat, 0, dtype=torch.int32) * flat).reshape(fluid.shape)
    own = torch.nonzero(flat).reshape(-1)
    table = torch.full((nq, n_pad), -1, dtype=torch.int32, device=fluid.device)
    dims = tuple(range(nd))
    for q in range(nq):
        c = [int(x) for x in basis[q]]
        up = torch.roll(label, shifts=c, dims=dims) if any(c) else label
        if any(c):
            for ax, name in enumerate("xyz"[:nd]):
                if name in periodic.lower() or c[ax] == 0:
                    continue
                sl = [slice(None)] * nd
                sl[ax] = 0 if c[ax] > 0 else -1
                up = up.clone() if up is label else up
                up[tuple(sl)] = 0
        table[q, :n] = up.reshape(-1)[own] - 1
        del up
    labels = torch.zeros(n_pad, dtype=torch.int32, device=fluid.device)
    labels[:n] = torch.arange(1, n + 1, dtype=torch.int32, device=fluid.device)
    return table, labels, n, n_pad


def sphere_pack_slab(shape, radius, porosity, seed, z0, z1):
    """the z-range [z0, z1) of geometry.sphere_pack(shape, ...) without building the whole array;
    z indices wrap periodically, so z0 = -1 / z1 = nz + 1 give the halo layers."""
    shape = tuple(int(s) for s in shape)
    nd = len(shape)
    vol = float(np.prod(shape))
    vs = np.pi * radius ** 2 if nd == 2 else 4.0 / 3.0 * np.pi * radius ** 3
    n_sph = max(1, int(round(-np.log(porosity) * vol / vs)))
    rng = np.random.default_rng(seed)
    centres = rng.random((n_sph, nd)) * np.array(shape)
    nzs = z1 - z0
    geo = np.ones(shape[:-1] + (nzs,), dtype=np.uint8)
    r = int(np.ceil(radius)) + 1
    off = np.arange(-r, r + 1)
    nz = shape[-1]
    for c in centres:
        base = np.floor(c).astype(np.int64)
        zs = base[-1] + off
        # slab-local index of each candidate z (periodic images considered)
        hits = []
        for shift in (-nz, 0, nz):
            zl = zs + shift - z0
            ok = (zl >= 0) & (zl < nzs)
            if ok.any():
                hits.append((zs[ok], zl[ok]))
        if not hits:
            continue
        idx = [(base[d] + off) for d in range(nd - 1)]
        d2xy = None
        for d in range(nd - 1):
            dd = (idx[d] - c[d]) ** 2
            sh = [1] * nd
            sh[d] = -1
            d2xy = dd.reshape(sh) if d2xy is None else d2xy + dd.reshape(sh)
        for zg, zl in hits:
            dz = ((zg - c[-1]) ** 2).reshape([1] * (nd - 1) + [-1])
            inside = (d2xy + dz) < radius ** 2
            sub = np.ix_(*([np.mod(idx[d], shape[d]) for d in range(nd - 1)] + [zl]))
            block = geo[sub]
            block[inside] = 0
            geo[sub] = block
    return geo


def build_slab_tables(fluid_ext: torch.Tensor, lattice: str, boundary_first: bool = True, wall_phi_ext: torch.Tensor = None,
                      order: str = "layer"):
    """One rank of a z-slab decomposition (periodic in x, y; z neighbours are other ranks).

    fluid_ext: bool [nx, ny, nz + 2]: the rank's own slab plus one halo layer below (index 0) and
    above (index -1) taken from the neighbour ranks.  Own fluid nodes are labelled 1..N in C-order
    (the reference's per-rank numbering, vtklb.py:92-94).  Device slots: nodes of the first and
    last z-layer first (they are the only halo-coupled ones) when boundary_first.  order = "layer" (default)
    puts the remaining nodes layer by layer (z slowest, each layer in C-order of (x, y)): the sources of 32
    consecutive slots then lie close together in EVERY direction, also next to the two leading layers, so the
    compact index needs (almost) no explicit rows; order = "label" keeps them in label order (z fastest), where
    every tile that touches the second or the second-to-last layer pulls from the far-away leading slots.

    Returns dict(table int32 [nQ, n_pad], labels int32 [n_pad], n, n_pad, n_halo, n_boundary,
    faces = {"down": (send_src, recv_dst), "up": (send_src, recv_dst)}) with int64 slot offsets
    q*stride + slot.  Message order of a face: directions ascending, receiver cells in C-order of (x, y).

    With wall_phi_ext (double [nx, ny, nz + 2], the wall colour (rho0 - rho1)/(rho0 + rho1) at solid cells) the
    colour-gradient tables of a two-field lattice are added: ptable int32 [nQ, n_pad] (phi slot of neighbor(q, n):
    own slot, n_pad + k for wall cell k -- solid cells next to an own fluid cell, the reference's solid boundary
    nodes -- then the ghost cells -- fluid cells of the two halo layers, lower layer first, C-order of (x, y) --
    and the zero slot n_pad + n_extra), phi_extra double [n_extra] (wall colours, zeros for the ghosts) and
    scalar_faces = {"down": (send_src, recv_dst), "up": (...)}: phi slots of my bottom / top layer in C-order,
    which is the order the neighbour numbers its ghosts in (communciateScalarField, LBmonlatmpi.h:181-205).
    """
    basis = G.BASIS[lattice]
    nq, nd = basis.shape
    assert nd == 3 and fluid_ext.dim() == 3
    dev = fluid_ext.device
    fluid_ext = fluid_ext.bool()
    own = fluid_ext[:, :, 1:-1]
    lower, upper = fluid_ext[:, :, 0], fluid_ext[:, :, -1]
    nz = own.shape[2]
    flat = own.reshape(-1)
    n = int(flat.sum().item())
    n_pad = ((n + 31) // 32) * 32
    label = (torch.cumsum(flat, 0, dtype=torch.int32) * flat).reshape(own.shape)
    own_idx = torch.nonzero(flat).reshape(-1)                 # flat cell index of label 1..n
    z_of = own_idx % nz
    if boundary_first:
        is_b = (z_of == 0) | (z_of == nz - 1)
        n_boundary = int(is_b.sum().item())
        if order == "layer":
            # tiers: layer 0, layer nz-1, then layers 1 .. nz-2; inside a tier the C-order of (x, y)
            tier = torch.where(z_of == 0, torch.zeros_like(z_of), torch.where(z_of == nz - 1, torch.ones_like(z_of), z_of + 1))
            key = tier * (own.shape[0] * own.shape[1]) + own_idx // nz
            order_idx = torch.argsort(key, stable=True)
        else:
            order_idx = torch.cat([torch.nonzero(is_b).reshape(-1), torch.nonzero(~is_b).reshape(-1)])  # label-1 per slot
    else:
        order_idx = torch.arange(n, device=dev)
        n_boundary = 0
    order = order_idx
    slot_of_label = torch.full((n + 1,), -1, dtype=torch.int32, device=dev)
    slot_of_label[order + 1] = torch.arange(n, dtype=torch.int32, device=dev)
    slot_grid = slot_of_label[label.long()]                    # -1 on solid cells
    cells_by_slot = own_idx[order]                             # flat own-cell index per slot

    def shift2(a, cx, cy):
        return torch.roll(a, shifts=(cx, cy), dims=(0, 1)) if (cx or cy) else a

    # halo element counts per direction
    recv_mask, counts = {}, [0] * nq
    for q in range(nq):
        cx, cy, cz = (int(v) for v in basis[q])
        if cz == 1:
            m = own[:, :, 0] & shift2(lower, cx, cy)
        elif cz == -1:
            m = own[:, :, -1] & shift2(upper, cx, cy)
        else:
            continue
        recv_mask[q] = m
        counts[q] = int(m.sum().item())
    n_halo = ((max(counts) + 15) // 16) * 16
    stride = n_pad + n_halo
    table = torch.full((nq, n_pad), -1, dtype=torch.int32, device=dev)
    faces = {"down": ([], []), "up": ([], [])}
    for q in range(nq):
        cx, cy, cz = (int(v) for v in basis[q])
        src = torch.roll(slot_grid, shifts=(cx, cy, cz), dims=(0, 1, 2)) if (cx or cy or cz) else slot_grid
        if cz != 0:
            src = src.clone()
            m = recv_mask[q]
            k = torch.cumsum(m.reshape(-1), 0, dtype=torch.int32).reshape(m.shape) - 1
            layer = torch.where(m, n_pad + k, torch.full_like(k, -1))
            if cz == 1:
                src[:, :, 0] = layer
                # I receive direction q from the rank below; the rank above receives it from my top layer
                faces["down"][1].append(q * stride + n_pad + torch.arange(counts[q], device=dev, dtype=torch.int64))
                ms = upper & shift2(own[:, :, -1], cx, cy)
                faces["up"][0].append(q * stride + shift2(slot_grid[:, :, -1], cx, cy)[ms].long())
            else:
                src[:, :, -1] = layer
                faces["up"][1].append(q * stride + n_pad + torch.arange(counts[q], device=dev, dtype=torch.int64))
                ms = lower & shift2(own[:, :, 0], cx, cy)
                faces["down"][0].append(q * stride + shift2(slot_grid[:, :, 0], cx, cy)[ms].long())
        table[q, :n] = src.reshape(-1)[cells_by_slot]
        del src
    labels = torch.zeros(n_pad, dtype=torch.int32, device=dev)
    labels[:n] = (order + 1).to(torch.int32)
    cat = lambda l: torch.cat(l) if l else torch.zeros(0, dtype=torch.int64, device=dev)
    faces = {k: (cat(v[0]), cat(v[1])) for k, v in faces.items()}
    out = dict(table=table, labels=labels, n=n, n_pad=n_pad, n_halo=n_halo, n_boundary=n_boundary, faces=faces,
               stride=stride)
    if wall_phi_ext is not None:
        own_ext = fluid_ext.clone()
        own_ext[:, :, 0] = False
        own_ext[:, :, -1] = False
        near = torch.zeros_like(own_ext)
        for q in range(nq - 1):
            cx, cy, cz = (int(v) for v in basis[q])
            near |= torch.roll(own_ext, shifts=(cx, cy, cz), dims=(0, 1, 2))  # cell has an own fluid neighbour
        wall = near & ~fluid_ext
        ghost = fluid_ext & ~own_ext
        ghost[:, :, 1:-1] = False
        n_wall = int(wall.sum().item())
        n_low, n_up = int(ghost[:, :, 0].sum().item()), int(ghost[:, :, -1].sum().item())
        n_extra = n_wall + n_low + n_up
        slot_ext = torch.full(fluid_ext.shape, n_pad + n_extra, dtype=torch.int32, device=dev)
        slot_ext[:, :, 1:-1] = torch.where(own, slot_grid, slot_ext[:, :, 1:-1])
        slot_ext[wall] = n_pad + torch.arange(n_wall, dtype=torch.int32, device=dev)
        low_slots = n_pad + n_wall + torch.arange(n_low, dtype=torch.int32, device=dev)
        up_slots = n_pad + n_wall + n_low + torch.arange(n_up, dtype=torch.int32, device=dev)
        layer = slot_ext[:, :, 0]
        layer[ghost[:, :, 0]] = low_slots
        layer = slot_ext[:, :, -1]
        layer[ghost[:, :, -1]] = up_slots
        ptable = torch.full((nq, n_pad), n_pad + n_extra, dtype=torch.int32, device=dev)
        ext_cells = (cells_by_slot // nz) * (nz + 2) + (cells_by_slot % nz) + 1   # flat index in the extended array
        for q in range(nq):
            cx, cy, cz = (-int(v) for v in basis[q])                              # value at pos + c_q
            nb = torch.roll(slot_ext, shifts=(cx, cy, cz), dims=(0, 1, 2)) if (cx or cy or cz) else slot_ext
            ptable[q, :n] = nb.reshape(-1)[ext_cells]
            del nb
        phi_extra = torch.zeros(max(n_extra, 1), dtype=torch.float64, device=dev)
        phi_extra[:n_wall] = wall_phi_ext.to(dev).double()[wall]
        out.update(ptable=ptable, n_extra=n_extra, phi_extra=phi_extra, scalar_faces={
            "down": (slot_grid[:, :, 0][own[:, :, 0]].long(), low_slots.long()),
            "up": (slot_grid[:, :, -1][own[:, :, -1]].long(), up_slots.long())})
    return out


def build_phi_table(fluid: torch.Tensor, wall_phi: torch.Tensor, lattice: str, periodic: str = "xyz"):
    """Colour-gradient support tables of a single-rank two-phase lattice (twophase/main_TWOPHASE.cpp):
    ptable[q][i] = phi slot of neighbor(q, n_i) -- own fluid node -> its slot, solid cell next to a
    fluid cell (the reference's solid boundary nodes, LBgeometry.h:37-45) -> n_pad + k, anything else ->
    the zero slot n_pad + n_extra -- and the wall colours phi_extra[k] = wall_phi at those solid cells.
    Non-periodic axes are closed (outside = zero slot)."""
    basis = G.BASIS[lattice]
    nq, nd = basis.shape
    fluid = fluid.bool()
    dev = fluid.device
    flat = fluid.reshape(-1)
    n = int(flat.sum().item())
    n_pad = ((n + 31) // 32) * 32
    dims = tuple(range(nd))
    near = torch.zeros_like(fluid)
    for q in range(nq - 1):
        c = [int(x) for x in basis[q]]
        near |= torch.roll(fluid, shifts=c, dims=dims)   # cell has a fluid neighbour
    wall = near & ~fluid
    n_extra = int(wall.sum().item())
    slot = torch.full(fluid.shape, n_pad + n_extra, dtype=torch.int32, device=dev)
    slot[fluid] = torch.arange(n, dtype=torch.int32, device=dev)
    slot[wall] = n_pad + torch.arange(n_extra, dtype=torch.int32, device=dev)
    own = torch.nonzero(flat).reshape(-1)
    ptable = torch.full((nq, n_pad), n_pad + n_extra, dtype=torch.int32, device=dev)
    for q in range(nq):
        c = [-int(x) for x in basis[q]]                   # value at pos + c_q
        nb = torch.roll(slot, shifts=c, dims=dims) if any(c) else slot
        if any(c):
            for ax, name in enumerate("xyz"[:nd]):
                if name in periodic.lower() or c[ax] == 0:
                    continue
                sl = [slice(None)] * nd
                sl[ax] = 0 if c[ax] > 0 else -1
                nb = nb.clone()
                nb[tuple(sl)] = n_pad + n_extra
        ptable[q, :n] = nb.reshape(-1)[own]
    return ptable, n_extra, wall_phi.to(dev).double()[wall].contiguous()
