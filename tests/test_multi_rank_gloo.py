"""N>1 host logic on the CPU: world_size-2 and -3 rings over gloo.  Every rank builds its z-slab
tables with the structured ingest, packs its face populations, exchanges them with
multi.RingHalo (the same object the NCCL path uses) and unpacks; pulling through the slab
table must then reproduce the pull of the undecomposed periodic geometry."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lattice, out_q):
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests"))
    pkg = helpers.load_package()
    ing = importlib.import_module("badchimp_cpp_b200.ingest")
    multi = importlib.import_module("badchimp_cpp_b200.multi")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        G = pkg.geometry
        nzr = 5
        shape = (8, 7, nzr * world)
        geo = torch.from_numpy(G.sphere_pack(shape, 2.4, 0.6, 9)).bool()
        nq = len(G.BASIS[lattice])
        rev = [G.reverse_direction(lattice, q) for q in range(nq)]
        Tg, _, ng, _ = ing.build_pull_table(geo, lattice, "xyz")
        Xg = np.random.default_rng(5).random((nq, ng))
        glabel = (torch.cumsum(geo.reshape(-1), 0) * geo.reshape(-1)).reshape(shape).numpy()
        z = (np.arange(rank * nzr - 1, (rank + 1) * nzr + 1)) % shape[2]
        sl = ing.build_slab_tables(geo[:, :, torch.from_numpy(z)], lattice, True)
        own = geo[:, :, rank * nzr:(rank + 1) * nzr].numpy()
        gl = glabel[:, :, rank * nzr:(rank + 1) * nzr][own]
        gslot = gl[sl["labels"][: sl["n"]].numpy() - 1] - 1
        X = np.zeros((nq, sl["stride"]))
        X[:, : sl["n"]] = Xg[:, gslot]
        f = sl["faces"]
        ring = multi.RingHalo(rank, world, len(f["down"][0]), len(f["down"][1]), len(f["up"][0]), len(f["up"][1]), "cpu")
        ring.send_down[: len(f["down"][0])] = torch.from_numpy(X.reshape(-1)[f["down"][0].numpy()])
        ring.send_up[: len(f["up"][0])] = torch.from_numpy(X.reshape(-1)[f["up"][0].numpy()])
        ring.exchange()
        X.reshape(-1)[f["down"][1].numpy()] = ring.recv_down[: len(f["down"][1])].numpy()
        X.reshape(-1)[f["up"][1].numpy()] = ring.recv_up[: len(f["up"][1])].numpy()
        ok = True
        n = sl["n"]
        for q in range(nq):
            s = sl["table"][q, :n].numpy()
            pulled = np.where(s >= 0, X[q, np.maximum(s, 0)], X[rev[q], np.arange(n)])
            sg = Tg[q, :ng].numpy()
            ref = np.where(sg >= 0, Xg[q, np.maximum(sg, 0)], Xg[rev[q], np.arange(ng)])[gslot]
            ok = ok and np.array_equal(pulled, ref)
        out_q.put((rank, bool(ok), int(n)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,lattice", [(2, "D3Q19"), (3, "D3Q27")])
def test_ring_halo_exchange_over_gloo(world, lattice):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lattice, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] for r in res), res


def _twophase_worker(rank, world, port, out_q):
    """scalar (phi) halo of the two-phase slabs over the ring + the balanced cut planes, CPU / gloo"""
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests"))
    pkg = helpers.load_package()
    ing = importlib.import_module("badchimp_cpp_b200.ingest")
    multi = importlib.import_module("badchimp_cpp_b200.multi")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        G = pkg.geometry
        nzr = 6
        shape = (9, 8, nzr * world)
        geo = G.sphere_pack(shape, 2.5, 0.55, 21).astype(bool)
        wall = np.where(geo, 0.0, 0.25)
        # cut planes at equal fluid-node counts: every rank computes the same cuts from its own layer counts
        layers = torch.from_numpy(geo[:, :, rank * nzr:(rank + 1) * nzr].reshape(-1, nzr).sum(axis=0).astype(np.int64))
        cuts = multi.balanced_cuts(layers, world)
        counts = [int(geo[:, :, cuts[k]:cuts[k + 1]].sum()) for k in range(world)]
        z0, z1 = cuts[rank], cuts[rank + 1]
        idx = np.arange(z0 - 1, z1 + 1) % shape[2]
        sl = ing.build_slab_tables(torch.from_numpy(geo[:, :, idx]), "D3Q19", True, torch.from_numpy(wall[:, :, idx]))
        # phi = a value that identifies the global cell, so that a ghost slot can be checked against its owner
        gid = np.cumsum(geo.reshape(-1)).reshape(shape).astype(np.float64)
        own = geo[:, :, z0:z1]
        lab = sl["labels"][: sl["n"]].numpy()
        n_extra = sl["n_extra"]
        phi = np.zeros(sl["n_pad"] + n_extra + 1)
        phi[: sl["n"]] = gid[:, :, z0:z1][own][lab - 1]
        phi[sl["n_pad"]: sl["n_pad"] + n_extra] = sl["phi_extra"][:n_extra].numpy()
        sf = sl["scalar_faces"]
        ring = multi.RingHalo(rank, world, len(sf["down"][0]), len(sf["down"][1]), len(sf["up"][0]), len(sf["up"][1]), "cpu")
        ring.send_down[: len(sf["down"][0])] = torch.from_numpy(phi[sf["down"][0].numpy()])
        ring.send_up[: len(sf["up"][0])] = torch.from_numpy(phi[sf["up"][0].numpy()])
        ring.exchange()
        phi[sf["down"][1].numpy()] = ring.recv_down[: len(sf["down"][1])].numpy()
        phi[sf["up"][1].numpy()] = ring.recv_up[: len(sf["up"][1])].numpy()
        # every ghost slot now holds the id of the fluid cell of the neighbour's layer it stands for
        low = gid[:, :, (z0 - 1) % shape[2]][geo[:, :, (z0 - 1) % shape[2]]]
        up = gid[:, :, z1 % shape[2]][geo[:, :, z1 % shape[2]]]
        ok = np.array_equal(phi[sf["down"][1].numpy()], low) and np.array_equal(phi[sf["up"][1].numpy()], up)
        # and the table of a boundary node reaches them: neighbour in +z of a top-layer node is that ghost
        pt = sl["ptable"].numpy()
        q_up = [q for q in range(19) if tuple(G.BASIS["D3Q19"][q]) == (0, 0, 1)][0]
        top = np.argwhere(own[:, :, -1])
        for slot in range(sl["n"]):
            cell = np.argwhere(own)[lab[slot] - 1]
            if cell[2] != z1 - z0 - 1:
                continue
            x, y = cell[0], cell[1]
            want = gid[x, y, z1 % shape[2]] if geo[x, y, z1 % shape[2]] else None
            got = phi[pt[q_up, slot]]
            ok = ok and (got == want if want is not None else got in (0.25, 0.0))
        # the sum of one double over the ranks, added in rank order (what the mailboxes do on the GPUs)
        mine = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
        gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, mine)
        total = 0.0
        for t in gathered:
            total = total + float(t[0])
        out_q.put((rank, bool(ok), cuts, counts, total))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_twophase_scalar_ring_and_balanced_cuts_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_twophase_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert all(r[2] == res[0][2] and r[4] == res[0][4] for r in res)          # same cuts, same sum bits on every rank
    cuts, counts = res[0][2], res[0][3]
    assert cuts[0] == 0 and cuts[-1] == 6 * world and all(b > a for a, b in zip(cuts, cuts[1:]))
    assert max(counts) - min(counts) <= 2 * 9 * 8                              # within one z-layer of each other


def _strong_split_worker(rank, world, port, nz, out_q):
    """the cut planes bench.py --gpus N uses for the strong-scaling split (workloads._balanced_range), slabs of unequal
    thickness included (nz not divisible by the rank count)"""
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests"))
    pkg = helpers.load_package()
    ing = importlib.import_module("badchimp_cpp_b200.ingest")
    multi = importlib.import_module("badchimp_cpp_b200.multi")
    W = importlib.import_module("badchimp_cpp_b200.workloads")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        wl = dict(W.WORKLOADS["std_case"])
        gshape = (16, 16, nz)
        cuts = W._balanced_range(multi, wl, ing, gshape, rank, world, torch.device("cpu"), True)
        full = ing.sphere_pack_slab(gshape, gshape[0] / 8.0, W.PACK_POROSITY, W.PACK_SEED, 0, nz)
        counts = [int(full[:, :, cuts[k]:cuts[k + 1]].sum()) for k in range(world)]
        layer_max = int(full.reshape(-1, nz).sum(axis=0).max())
        out_q.put((rank, [int(c) for c in cuts], counts, layer_max))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nz", [(2, 16), (3, 16), (3, 13)])
def test_strong_scaling_cut_planes_over_gloo(world, nz):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_strong_split_worker, args=(r, world, port, nz, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cuts, counts, layer_max = res[0][1], res[0][2], res[0][3]
    assert all(r[1] == cuts for r in res)                                       # every rank computes the same planes
    assert cuts[0] == 0 and cuts[-1] == nz and all(b > a for a, b in zip(cuts, cuts[1:]))
    assert max(counts) - min(counts) <= 2 * layer_max                           # within a z-layer or so of each other
