"""SURVEY 8(f1) for host code: chimp_create_from_voxels builds a rank's lattice from a raw voxel array on the host
(no torch), so that a C / C++ main reaches sizes the reference's ASCII route cannot load (LBvtk.h:194-201).
Its host half is checked here, on the CPU, against the two independent producers of the same tables: the
torch ingest (badchimp-cpp_b200/ingest.py) and the generic symbolic builder fed with reference-style tables
(geometry.py == vtklb.py + Grid + Nodes, chimp_build_host); the gpu-marked test runs the lattice against the
oracle port."""
import importlib

import numpy as np
import pytest

import helpers


def _cases():
    rng = np.random.default_rng(77)
    out = []
    for lattice, shape, periodic in (("D3Q19", (9, 7, 11), "xyz"), ("D3Q27", (6, 8, 5), "xyz"), ("D2Q9", (13, 10), "xy"),
                                     ("D3Q19", (8, 9, 7), "xz"), ("D2Q9", (12, 9), "x"), ("D3Q27", (5, 5, 6), ""), ("D3Q19", (33, 4, 4), "y")):
        geo = (rng.random(shape) < 0.7).astype(np.uint8)
        for ax, name in enumerate("xyz"[: len(shape)]):     # closed along the non-periodic axes, as the reference requires
            if name not in periodic:
                sl = [slice(None)] * len(shape)
                sl[ax] = 0
                geo[tuple(sl)] = 0
                sl[ax] = -1
                geo[tuple(sl)] = 0
        out.append((lattice, geo, periodic))
    return out


@pytest.mark.parametrize("lattice,geo,periodic", _cases())
def test_host_voxel_table_equals_torch_ingest_and_generic_builder(lattice, geo, periodic):
    import torch
    pkg = helpers.load_package()
    ingest = importlib.import_module("badchimp_cpp_b200.ingest")
    table, labels, n = pkg.capi.voxel_table_host(lattice, geo, periodic)
    t_table, t_labels, t_n, t_pad = ingest.build_pull_table(torch.from_numpy(geo).bool(), lattice, periodic)
    assert n == t_n and table.shape[1] == t_pad
    assert np.array_equal(table[:, :n], t_table.numpy()[:, :n])
    assert np.array_equal(labels, t_labels.numpy())
    # and the generic builder on reference-style tables (bounce back on every fluid boundary node)
    lg = pkg.geometry.LatticeGeometry(geo.astype(int), lattice, periodic)
    tab = lg.all_ranks()[0]
    lat = pkg.capi.Lattice.from_rank_tables(tab)
    lat.add_halfway_bb(*tab.halfway_bb(tab.fluid_bnd_nodes()))
    lat.build_host(False)
    h_table, h_labels, _, info = lat.host_table()
    assert info["n"] == n and np.array_equal(h_labels, labels[:n])
    assert np.array_equal(h_table, table[:, :n])


@pytest.mark.parametrize("lattice,geo,periodic", [c for c in _cases() if c[2] in ("xyz", "xy") and c[0] != "D3Q27"])
def test_host_voxel_phi_table_equals_torch_ingest(lattice, geo, periodic):
    import torch
    pkg = helpers.load_package()
    ingest = importlib.import_module("badchimp_cpp_b200.ingest")
    rng = np.random.default_rng(5)
    wall_phi = np.where(geo == 0, rng.uniform(-0.9, 0.9, geo.shape), 0.0)
    ptable, n_extra, extra = pkg.capi.voxel_phi_table_host(lattice, geo, wall_phi, periodic)
    t_pt, t_extra_n, t_extra = ingest.build_phi_table(torch.from_numpy(geo).bool(), torch.from_numpy(wall_phi), lattice, periodic)
    n = int(geo.sum())
    assert n_extra == t_extra_n
    assert np.array_equal(ptable[:, :n], t_pt.numpy()[:, :n])
    assert np.array_equal(extra, t_extra.numpy())


def test_voxel_ingest_argument_errors():
    pkg = helpers.load_package()
    with pytest.raises(pkg.capi.ChimpError, match="no fluid"):
        pkg.capi.voxel_table_host("D3Q19", np.zeros((4, 4, 4), dtype=np.uint8))
    with pytest.raises(pkg.capi.ChimpError, match="2-D lattice"):
        pkg.capi.voxel_table_host("D2Q9", np.ones((4, 4, 4), dtype=np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("lattice,geo,periodic", _cases()[:5])
def test_lattice_from_voxels_bit_exact_vs_oracle_port(lattice, geo, periodic):
    pkg = helpers.load_package()
    port = helpers.oracle_port()
    lat = pkg.capi.lattice_from_voxels(lattice, geo, periodic)
    lg = pkg.geometry.LatticeGeometry(geo.astype(int), lattice, periodic)
    tab = lg.all_ranks()[0]
    bulk = tab.bulk_nodes()
    rng = np.random.default_rng(3)
    rho0 = 1.0 + 0.05 * rng.random(geo.shape)
    f0 = pkg.cases.std_case_initial_state(tab, rho0)[0]
    n = len(bulk)
    assert np.array_equal(bulk, np.arange(1, n + 1))          # labels 1..N are the leading rows
    lat.upload(f0[: n + 1])
    F = (1e-6, -2e-6, 5e-7)[: lg.nd]
    lat.step_single(6, tau=0.9, force=F)
    pr = port.PortRank(pkg.geometry.LATTICE_ID[lattice], tab.neigh, bulk, 1, tab.halfway_bb(tab.fluid_bnd_nodes()))
    pr.f[:] = f0
    pr.step_std_case(6, tau=0.9, force=F)
    assert np.array_equal(lat.download()[1:], pr.f[bulk])
    assert np.array_equal(lat.download_rho()[1:, 0], pr.rho[bulk, 0])
    lat.close()


@pytest.mark.gpu
def test_twophase_lattice_from_voxels_vs_oracle_port():
    pkg = helpers.load_package()
    port = helpers.oracle_port()
    geo = pkg.geometry.sphere_pack((14, 12, 16), 3.0, 0.6, 4).astype(np.uint8)
    lg = pkg.geometry.LatticeGeometry(geo.astype(int), "D3Q19", "xyz")
    tab = lg.all_ranks()[0]
    bulk = tab.bulk_nodes()
    x = np.arange(geo.shape[0])[:, None, None] * np.ones(geo.shape)
    rho0 = (x < geo.shape[0] / 2).astype(np.float64)
    wet = 0.3 * (geo == 0)
    s = pkg.cases.two_phase_setup(lg, [tab], rho0, 1.0 - rho0, wet)[0]
    # wall colour from the wall densities (main_TWOPHASE.cpp:173-181, 280-284): rho0 = wet, rho1 = 1 - wet (float32)
    w32 = wet.astype(np.float32)
    wall_phi = ((w32.astype(np.float64)) - (np.float32(1.0) - w32).astype(np.float64)) / ((w32.astype(np.float64)) + (np.float32(1.0) - w32).astype(np.float64))
    lat = pkg.capi.lattice_from_voxels("D3Q19", geo, "xyz", n_fields=2, wall_phi=wall_phi)
    n = len(bulk)
    lat.upload(s["f0"][: n + 1])
    args = (1.0, 0.8, 0.01, 1.0, 1e-5, (0.0, 1e-7, 0.0), n)
    lat.step_twophase(5, *args)
    pr = port.PortRank(1, tab.neigh, bulk, 2, tab.halfway_bb(bulk))
    pr.f[:] = s["f0"]
    pr.rho[:] = s["rho"]
    pr.step_twophase(5, s["solid_bnd"], *args)
    got, want = lat.download()[1:], pr.f[bulk]
    floor = 1e-3 / 36.0
    assert float((np.abs(got - want) / np.maximum(np.abs(want), floor)).max()) <= 1e-9
    lat.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lattice,shape,periodic", [("D3Q19", (20, 16, 18), "xyz"), ("D2Q9", (24, 18), "x")])
def test_cpp_voxel_case_app_reproduces_oracle_port(lattice, shape, periodic, tmp_path):
    """host/apps/voxel_case.cpp: a C++ main on chimp_create_from_voxels (no torch, no .vtklb file) against the oracle port"""
    import json
    import struct
    import subprocess
    from test_host_cpp import build
    pkg = helpers.load_package()
    port = helpers.oracle_port()
    exe = build("voxel_case", link_engine=True)
    if len(shape) == 3:
        geo = pkg.geometry.sphere_pack(shape, 4.0, 0.6, 8).astype(np.uint8)
    else:
        geo = np.ones(shape, dtype=np.uint8)
        geo[:, 0] = 0
        geo[:, -1] = 0
    raw = str(tmp_path / "geo.raw")
    geo.tofile(raw)
    dims = list(shape) + [1] * (3 - len(shape))
    steps, tau, F = 7, 0.8, (1e-6, 2e-7, 0.0)
    out = str(tmp_path / "out.bin")
    r = subprocess.run([exe, lattice] + [str(d) for d in dims] + [raw, periodic, str(steps), repr(tau)] + [repr(x) for x in F] + [out],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    lg = pkg.geometry.LatticeGeometry(geo.astype(int), lattice, periodic)
    tab = lg.all_ranks()[0]
    bulk = tab.bulk_nodes()
    n = len(bulk)
    assert line["fluid_nodes"] == n and line["iterations"] == steps
    pr = port.PortRank(pkg.geometry.LATTICE_ID[lattice], tab.neigh, bulk, 1, tab.halfway_bb(tab.fluid_bnd_nodes()))
    pr.f[:] = pkg.cases.std_case_initial_state(tab, np.ones(geo.shape))[0]
    pr.step_std_case(steps, tau=tau, force=F[: lg.nd])
    data = open(out, "rb").read()
    assert struct.unpack_from("<i", data, 0)[0] == n
    rho = np.frombuffer(data, dtype="<f8", count=n + 1, offset=4)
    vel = np.frombuffer(data, dtype="<f8", count=(n + 1) * lg.nd, offset=4 + 8 * (n + 1)).reshape(n + 1, lg.nd)
    assert np.array_equal(rho[1:], pr.rho[bulk, 0])
    assert np.array_equal(vel[1:], pr.vel[bulk])


@pytest.mark.parametrize("lattice,shape,seed", [("D3Q19", (9, 8, 5), 1), ("D3Q27", (7, 6, 4), 2), ("D3Q19", (12, 10, 2), 3), ("D3Q19", (6, 7, 9), 4)])
def test_host_slab_tables_equal_torch_ingest(lattice, shape, seed):
    """chimp_slab_tables_host (the host half of chimp_create_slab_from_voxels) against ingest.build_slab_tables: pull table,
    labels, slot order (halo-coupled layers first, then layer by layer), halo sizes and the four face lists"""
    import torch
    pkg = helpers.load_package()
    ingest = importlib.import_module("badchimp_cpp_b200.ingest")
    rng = np.random.default_rng(seed)
    nx, ny, nz = shape
    ext = (rng.random((nx, ny, nz + 2)) < 0.7).astype(np.uint8)
    mine = pkg.capi.slab_tables_host(lattice, ext)
    ref = ingest.build_slab_tables(torch.from_numpy(ext).bool(), lattice, True)
    for key in ("n", "n_pad", "n_halo", "n_boundary", "stride"):
        assert mine[key] == ref[key], key
    n = mine["n"]
    assert np.array_equal(mine["labels"], ref["labels"].numpy())
    assert np.array_equal(mine["table"][:, :n], ref["table"].numpy()[:, :n])
    for face in ("down", "up"):
        for k in (0, 1):
            assert np.array_equal(mine["faces"][face][k], ref["faces"][face][k].numpy()), (face, k)


@pytest.mark.gpu
@pytest.mark.parametrize("lattice,world", [("D3Q19", 2), ("D3Q19", 3), ("D3Q27", 2)])
def test_slab_lattices_from_voxels_with_peer_halos_vs_oracle_port(lattice, world):
    """N z-slabs created by chimp_create_slab_from_voxels in one process, connected through peer memory (fused into the
    step kernel), against the oracle port of the undecomposed geometry: bit for bit"""
    from test_gpu_parity import _connect_in_process
    pkg = helpers.load_package()
    port = helpers.oracle_port()
    nzr = 6
    shape = (14, 12, nzr * world)
    geo = pkg.geometry.sphere_pack(shape, 3.5, 0.6, 12).astype(np.uint8)
    lats = []
    for r in range(world):
        idx = np.arange(r * nzr - 1, (r + 1) * nzr + 1) % shape[2]
        lat = pkg.capi.slab_lattice_from_voxels(lattice, geo[:, :, idx], (r - 1) % world, (r + 1) % world)
        lat.init_uniform(1.0)
        lats.append(lat)
    # ring: face 0 of rank r looks down at rank r - 1 (whose face 1 looks up at r), face 1 looks up at rank r + 1
    for r, lat in enumerate(lats):
        down, up = lats[(r - 1) % world], lats[(r + 1) % world]
        lat.connect_peer(0, down.nq * down.plane_stride(), 1, down.recv_dst(1), pointers=down.local_pointers())
        lat.connect_peer(1, up.nq * up.plane_stride(), 0, up.recv_dst(0), pointers=up.local_pointers())
    assert all(lat.peer_mode()[0] == 2 for lat in lats), [lat.peer_mode() for lat in lats]
    steps, tau, F = 6, 0.8, (1e-6, 3e-7, -2e-7)
    for _ in range(steps):
        for lat in lats:
            lat.step_begin(tau=tau, force=F)
        for lat in lats:
            lat.step_end()
    lg = pkg.geometry.LatticeGeometry(geo.astype(int), lattice, "xyz")
    tab = lg.all_ranks()[0]
    bulk = tab.bulk_nodes()
    pr = port.PortRank(pkg.geometry.LATTICE_ID[lattice], tab.neigh, bulk, 1, tab.halfway_bb(tab.fluid_bnd_nodes()))
    pr.f[:] = pkg.cases.std_case_initial_state(tab, np.ones(geo.shape))[0]
    pr.step_std_case(steps, tau=tau, force=F)
    glabel = (np.cumsum(geo.reshape(-1)) * geo.reshape(-1)).reshape(shape)
    checked = 0
    for r, lat in enumerate(lats):
        own = geo[:, :, r * nzr:(r + 1) * nzr].astype(bool)
        gl = glabel[:, :, r * nzr:(r + 1) * nzr][own]          # global label of my cells in local label order
        assert np.array_equal(lat.download()[1:], pr.f[gl]), "rank %d" % r
        checked += len(gl)
        lat.close()
    assert checked == len(bulk)
