"""Parity at BASELINE sizes.

* configs[0] -- std_case, D3Q19 BGK, 128^3 sphere pack -- against dumps the reference itself produces on the fly
  (oracle/_ref/ref_driver = the unmodified reference headers; it travels to the GPU box as a prebuilt binary), as one
  rank and as 8 MPI ranks (threads) whose GPU counterparts are 8 engine contexts exchanging halos through peer stores
  fused into the step kernel: explicit-row tiles, > 20 k tiles per context and the boundary-first ordering at scale.
* the structured-ingest path (the one bench.py times) at 256^3 against the oracle port for a few steps.
Both bit for bit."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

REF_DRIVER = os.path.join(helpers.ROOT, "oracle", "_ref", "ref_driver")


@pytest.mark.skipif(not os.path.exists(REF_DRIVER), reason="needs oracle/_ref/ref_driver (the reference's own headers)")
@pytest.mark.parametrize("nranks", [1, 8])
def test_config0_std_case_128_cube_bit_exact_vs_the_reference_itself(nranks, tmp_path):
    sys.path.insert(0, os.path.join(helpers.ROOT, "oracle"))
    from recfile import read_rec
    from test_gpu_parity import _connect_in_process
    pkg = helpers.load_package()
    G, capi = pkg.geometry, pkg.capi
    size, steps, tau, force = 128, 5, 0.8, (1e-6, 0.0, 0.0)
    geo = G.sphere_pack((size,) * 3, size / 8.0, 0.35, 1234).astype(int)
    lg = G.LatticeGeometry(G.z_slab_rank_map(geo, nranks) if nranks > 1 else geo, "D3Q19", "xyz")
    tabs = lg.all_ranks()
    ones = np.ones(geo.shape)
    os.makedirs(str(tmp_path / "out"))
    for t in tabs:
        t.write_vtklb(str(tmp_path / ("tmp%d.vtklb" % t.my_rank)), {"init_rho": ones})
    cmd = [REF_DRIVER, "--case", "std_case", "--lattice", "D3Q19", "--dir", str(tmp_path), "--out", str(tmp_path / "out"),
           "--nranks", str(nranks), "--steps", str(steps), "--dump", str(steps), "--no-tables", "--tau", repr(tau),
           "--force", ",".join(repr(x) for x in force)]
    subprocess.run(cmd, check=True, capture_output=True, timeout=900)
    lats = []
    for t in tabs:
        lat = capi.Lattice.from_rank_tables(t)
        ss = t.send_side(tabs)
        for k, nr in enumerate(t.neig_ranks):
            lat.add_neighbor(nr, ss[k][0], ss[k][1], ss[k][2], t.recv_nodes[k], t.recv_ndir[k], t.recv_dirs[k])
        lat.add_halfway_bb(*t.halfway_bb(t.fluid_bnd_nodes()))
        lat.finalize(capi.INDEX_COMPACT, nranks > 1)
        lat.upload(pkg.cases.std_case_initial_state(t, ones)[0])
        lats.append(lat)
    if nranks > 1:
        _connect_in_process(lats)
        assert all(lat.peer_mode()[0] == 2 for lat in lats), [lat.peer_mode() for lat in lats]
    for _ in range(steps):          # one step at a time so that the contexts of one GPU advance together
        for lat in lats:
            lat.step_begin(tau=tau, force=force)
        for lat in lats:
            lat.step_end()
    checked = 0
    for r, (lat, t) in enumerate(zip(lats, tabs)):
        rec = read_rec(str(tmp_path / "out" / ("rank%d.rec" % r)))
        bulk = t.bulk_nodes()
        assert np.array_equal(lat.download()[bulk, 0], rec["step%d.f" % steps].reshape(-1, 19)[bulk]), "rank %d" % r
        assert np.array_equal(lat.download_rho()[bulk, 0], rec["step%d.rho" % steps][bulk])
        assert np.array_equal(lat.download_vel()[bulk], rec["step%d.vel" % steps].reshape(-1, 3)[bulk])
        checked += len(bulk)
        lat.close()
    assert checked == int(geo.sum())


def test_ingest_path_256_cube_bit_exact_vs_oracle_port():
    """the bench's own code path (device-side ingest -> compact index -> step kernel -> download in reference layout) on a
    256^3 pack (5.9 M fluid nodes, 185 k tiles) against the oracle port of the same geometry, after 1 and after 4 steps"""
    import torch
    pkg = helpers.load_package()
    ingest = importlib.import_module("badchimp_cpp_b200.ingest")
    multi = importlib.import_module("badchimp_cpp_b200.multi")
    bench_impl = importlib.import_module("badchimp_cpp_b200.bench_impl")
    W = importlib.import_module("badchimp_cpp_b200.workloads")
    res = bench_impl.parity_probe(pkg, ingest, multi, W.WORKLOADS["std_case"], "std_case", 0, 1, torch.device("cuda", 0),
                                  pkg.capi.INDEX_COMPACT, "peer", False, steps=4, size=256)
    assert res["checked_nodes"] > 5_000_000
    assert res["bit_exact"] and res["after_4_steps"]["bit_exact"], res
