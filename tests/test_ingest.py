"""Host logic: the structured (torch) geometry ingest.
* single rank: its pull table equals the table of the generic host builder (chimp_build_host) fed
  with the reference-numbered tables, for D2Q9 / D3Q19 / D3Q27, periodic and closed;
* z-slab ranks: pulling through the slab tables + face exchange lists reproduces the pull of the
  undecomposed periodic geometry (the halo order is the ingest's own, so it is checked by meaning);
* the slab generator of the sphere pack equals slices of the full pack."""
import importlib

import numpy as np
import pytest
import torch

import helpers


def _mods():
    pkg = helpers.load_package()
    return pkg, importlib.import_module("badchimp_cpp_b200.ingest")


@pytest.mark.parametrize("lattice,shape,periodic", [("D3Q19", (14, 12, 16), "xyz"), ("D3Q27", (10, 12, 9), "xyz"),
                                                    ("D2Q9", (20, 16), "xy"), ("D3Q19", (12, 12, 12), "")])
def test_single_rank_table_equals_host_builder(lattice, shape, periodic):
    pkg, ing = _mods()
    G = pkg.geometry
    geo = G.sphere_pack(shape, min(shape) / 5, 0.6, 1).astype(int)
    if periodic == "":
        for ax in range(len(shape)):
            for e in (0, -1):
                s = [slice(None)] * len(shape)
                s[ax] = e
                geo[tuple(s)] = 0
    t = G.LatticeGeometry(geo, lattice, periodic).all_ranks()[0]
    lat = pkg.capi.Lattice.from_rank_tables(t)
    lat.add_halfway_bb(*t.halfway_bb(t.fluid_bnd_nodes()))
    lat.build_host()
    tab, lab, _, info = lat.host_table()
    T, L, n, n_pad = ing.build_pull_table(torch.from_numpy(geo > 0), lattice, periodic)
    assert n == info["n"] and n_pad == info["n_pad"]
    assert np.array_equal(T[:, :n].numpy(), tab)
    assert np.array_equal(L[:n].numpy(), lab)


@pytest.mark.parametrize("n_ranks", [2, 3])
@pytest.mark.parametrize("boundary_first", [False, True])
@pytest.mark.parametrize("lattice", ["D3Q19", "D3Q27"])
def test_slab_tables_reproduce_global_pull(lattice, n_ranks, boundary_first):
    pkg, ing = _mods()
    G = pkg.geometry
    nzr = 6
    shape = (9, 8, nzr * n_ranks)
    geo = torch.from_numpy(G.sphere_pack(shape, 2.6, 0.6, 4)).bool()
    nq = len(G.BASIS[lattice])
    rev = [G.reverse_direction(lattice, q) for q in range(nq)]
    Tg, Lg, ng, _ = ing.build_pull_table(geo, lattice, "xyz")
    rng = np.random.default_rng(0)
    Xg = rng.random((nq, ng))
    pulled_g = np.empty((nq, ng))
    for q in range(nq):
        s = Tg[q, :ng].numpy()
        pulled_g[q] = np.where(s >= 0, Xg[q, np.maximum(s, 0)], Xg[rev[q], np.arange(ng)])
    glabel = (torch.cumsum(geo.reshape(-1), 0) * geo.reshape(-1)).reshape(shape).numpy()
    slabs, X = [], []
    for r in range(n_ranks):
        z = (np.arange(r * nzr - 1, (r + 1) * nzr + 1)) % shape[2]
        ext = geo[:, :, torch.from_numpy(z)]
        sl = ing.build_slab_tables(ext, lattice, boundary_first)
        own = geo[:, :, r * nzr:(r + 1) * nzr].numpy()
        gl = glabel[:, :, r * nzr:(r + 1) * nzr][own]           # global label per own cell in local C-order
        lab = sl["labels"][: sl["n"]].numpy()                   # local label per slot
        gslot = gl[lab - 1] - 1
        x = np.zeros((nq, sl["stride"]))
        x[:, : sl["n"]] = Xg[:, gslot]
        slabs.append((sl, gslot))
        X.append(x)
        if boundary_first:
            halo_users = np.nonzero((sl["table"][:, : sl["n"]].numpy() >= sl["n_pad"]).any(axis=0))[0]
            assert halo_users.size == 0 or halo_users.max() < sl["n_boundary"]
    # face exchange: my "up" message goes to the down-face slots of rank r+1 and vice versa
    for r in range(n_ranks):
        sl = slabs[r][0]
        up, down = (r + 1) % n_ranks, (r - 1) % n_ranks
        X[up].reshape(-1)[slabs[up][0]["faces"]["down"][1].numpy()] = X[r].reshape(-1)[sl["faces"]["up"][0].numpy()]
        X[down].reshape(-1)[slabs[down][0]["faces"]["up"][1].numpy()] = X[r].reshape(-1)[sl["faces"]["down"][0].numpy()]
    for r in range(n_ranks):
        sl, gslot = slabs[r]
        n = sl["n"]
        for q in range(nq):
            s = sl["table"][q, :n].numpy()
            pulled = np.where(s >= 0, X[r][q, np.maximum(s, 0)], X[r][rev[q], np.arange(n)])
            assert np.array_equal(pulled, pulled_g[q, gslot]), "rank %d q %d" % (r, q)


def test_sphere_pack_slab_equals_slices_of_full_pack():
    pkg, ing = _mods()
    full = pkg.geometry.sphere_pack((20, 18, 24), 4.0, 0.5, 3)
    ext = ing.sphere_pack_slab((20, 18, 24), 4.0, 0.5, 3, -1, 25)
    assert np.array_equal(ext[:, :, 1:-1], full)
    assert np.array_equal(ext[:, :, 0], full[:, :, -1]) and np.array_equal(ext[:, :, -1], full[:, :, 0])
    assert np.array_equal(ing.sphere_pack_slab((20, 18, 24), 4.0, 0.5, 3, 6, 12), full[:, :, 6:12])


@pytest.mark.parametrize("cuts", [[0, 5, 12], [0, 4, 8, 12]])
def test_slab_phi_tables_reference_the_same_cells_as_the_single_rank_table(cuts):
    """two-phase z-slabs: for every own node and direction the phi slot (own node / wall cell with its colour /
    ghost cell of a halo layer / zero slot) designates the same global cell as the undecomposed phi table; the
    scalar halo lists pair the sender's layer with the receiver's ghost numbering (C-order of (x, y))"""
    import importlib
    import torch
    pkg = helpers.load_package()
    ingest = importlib.import_module("badchimp_cpp_b200.ingest")
    shape = (10, 9, 12)
    geo = pkg.geometry.sphere_pack(shape, 2.5, 0.6, 5).astype(bool)
    wall = np.random.default_rng(1).random(shape) * 2 - 1
    pt, n_x, pe = ingest.build_phi_table(torch.from_numpy(geo), torch.from_numpy(wall), "D3Q19", "xyz")
    n = int(geo.sum())
    n_pad = ((n + 31) // 32) * 32
    gid = -np.ones(shape, dtype=np.int64)
    gid[geo] = np.arange(n)
    pt, pe = pt.numpy(), pe.numpy()

    def single(i, q):
        s = pt[q, i]
        return ("f", int(s)) if s < n_pad else (("w", float(pe[s - n_pad])) if s < n_pad + n_x else ("z",))

    sends, recvs = {}, {}
    for r in range(len(cuts) - 1):
        z0, z1 = cuts[r], cuts[r + 1]
        idx = np.arange(z0 - 1, z1 + 1) % shape[2]
        ext, wext = geo[:, :, idx], wall[:, :, idx]
        slab = ingest.build_slab_tables(torch.from_numpy(ext), "D3Q19", True, torch.from_numpy(wext))
        ptab, pex, npd, nex = slab["ptable"].numpy(), slab["phi_extra"].numpy(), slab["n_pad"], slab["n_extra"]
        labels = slab["labels"].numpy()[:slab["n"]]
        own = np.argwhere(ext[:, :, 1:-1])
        low, up = np.argwhere(ext[:, :, 0]), np.argwhere(ext[:, :, -1])
        n_wall = nex - len(low) - len(up)

        def cell_of_slot(s):
            x, y, zl = own[labels[s] - 1]
            return int(gid[x, y, (z0 + zl) % shape[2]])

        for slot in range(slab["n"]):
            me = cell_of_slot(slot)
            for q in range(19):
                s = ptab[q, slot]
                if s < npd:
                    got = ("f", cell_of_slot(s))
                elif s < npd + n_wall:
                    got = ("w", float(pex[s - npd]))
                elif s < npd + nex:
                    g = s - npd - n_wall
                    x, y = low[g] if g < len(low) else up[g - len(low)]
                    got = ("f", int(gid[x, y, (z0 - 1) % shape[2] if g < len(low) else z1 % shape[2]]))
                else:
                    got = ("z",)
                assert got == single(me, q), (r, slot, q)
        sf = slab["scalar_faces"]
        sends[r] = {k: [cell_of_slot(int(s)) for s in sf[k][0]] for k in ("down", "up")}
        recvs[r] = {"down": [int(gid[x, y, (z0 - 1) % shape[2]]) for x, y in low], "up": [int(gid[x, y, z1 % shape[2]]) for x, y in up]}
        assert list(sf["down"][1].numpy()) == list(range(npd + n_wall, npd + n_wall + len(low)))
        assert list(sf["up"][1].numpy()) == list(range(npd + n_wall + len(low), npd + nex))
    world = len(cuts) - 1
    for r in range(world):   # what I send down is what rank-1 expects in its "up" ghosts, and vice versa
        assert sends[r]["down"] == recvs[(r - 1) % world]["up"]
        assert sends[r]["up"] == recvs[(r + 1) % world]["down"]
