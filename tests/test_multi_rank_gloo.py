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
