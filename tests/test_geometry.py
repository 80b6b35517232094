"""Host logic: the in-memory geometry ingest reproduces, bit for bit, the tables the reference
built for the golden cases (vtklb.py numbering, Nodes types, bulk / boundary lists, half-way
bounce-back link classes, MonLatMpi exchange lists) and writes byte-identical .vtklb files."""
import os

import numpy as np
import pytest

import helpers


@pytest.mark.parametrize("name", helpers.all_golden_names())
def test_tables_match_reference(name):
    g = helpers.Golden(name)
    lg, tabs = helpers.build_tables(g)
    assert len(tabs) == g.nranks
    for r, t in enumerate(tabs):
        nq = lg.nq
        assert int(g.rec(r, "size")[0]) == t.size
        assert np.array_equal(g.rec(r, "neigh").reshape(-1, nq), t.neigh)
        assert np.array_equal(g.rec(r, "pos").reshape(t.size, -1), t.pos)
        assert np.array_equal(g.rec(r, "rank"), t.node_rank)
        assert np.array_equal(g.rec(r, "type"), t.node_type)
        assert np.array_equal(g.rec(r, "bulk"), t.bulk_nodes())
        assert np.array_equal(g.rec(r, "fluidBnd"), t.fluid_bnd_nodes())
        assert np.array_equal(g.rec(r, "solidBnd"), t.solid_bnd_nodes())
        if g.has(r, "bb.node"):
            bnd = t.fluid_bnd_nodes() if g.case == "std_case" else t.bulk_nodes()
            nodes, nb, ng, nd, links = t.halfway_bb(bnd)
            assert np.array_equal(g.rec(r, "bb.node"), nodes)
            assert np.array_equal(g.rec(r, "bb.nBeta"), nb)
            assert np.array_equal(g.rec(r, "bb.nGamma"), ng)
            assert np.array_equal(g.rec(r, "bb.nDelta"), nd)
            assert np.array_equal(g.rec(r, "bb.links"), links.ravel())
        assert int(g.rec(r, "nNeigRanks")[0]) == len(t.neig_ranks)
        ss = t.send_side(tabs)
        for k, nr in enumerate(t.neig_ranks):
            p = "mpi%d." % k
            assert int(g.rec(r, p + "neigRank")[0]) == nr
            assert np.array_equal(g.rec(r, p + "nodesReceived"), t.recv_nodes[k])
            assert np.array_equal(g.rec(r, p + "nDirPerNodeReceived"), t.recv_ndir[k])
            assert np.array_equal(g.rec(r, p + "dirListReceived"), t.recv_dirs[k])
            assert np.array_equal(g.rec(r, p + "nodesToSend"), ss[k][0])
            assert np.array_equal(g.rec(r, p + "nDirPerNodeToSend"), ss[k][1])
            assert np.array_equal(g.rec(r, p + "dirListToSend"), ss[k][2])


@pytest.mark.parametrize("name", ["onephase_d3q19_p1", "onephase_trt_d3q19_p2"])
def test_one_phase_links_match_reference(name):
    g = helpers.Golden(name)
    lg, tabs = helpers.build_tables(g)
    setup = helpers.one_phase_setup(g, lg, tabs)
    for r, pr in enumerate(setup):
        assert np.array_equal(g.rec(r, "solidLinks"), pr["solid_links"].ravel())
        assert np.array_equal(g.rec(r, "pressLinks"), pr["press_links"].ravel())
        assert np.array_equal(g.rec(r, "fluidLinks"), pr["fluid_links"].ravel())
        assert np.array_equal(g.rec(r, "addSource"), pr["add_source"])
        assert np.array_equal(g.rec(r, "scale"), pr["scale"])
        assert np.array_equal(g.rec(r, "interior"), pr["interior"])
        assert len(pr["press_links"]) > 0 and len(pr["fluid_links"]) > 0 and len(pr["solid_links"]) > 0


@pytest.mark.parametrize("name", ["std_d3q19_p1", "std_d3q19_box_p2"])
def test_vtklb_writer_is_byte_identical(name, tmp_path):
    g = helpers.Golden(name)
    lg, tabs = helpers.build_tables(g)
    for r, t in enumerate(tabs):
        path = tmp_path / ("mine%d.vtklb" % r)
        t.write_vtklb(str(path), {"init_rho": g.attr("init_rho")})
        ref = open(os.path.join(helpers.GOLDEN, "%s.tmp%d.vtklb" % (name, r))).read()
        assert open(path).read() == ref


def test_sphere_pack_is_deterministic_and_near_target_porosity():
    pkg = helpers.load_package()
    a = pkg.geometry.sphere_pack((48, 48, 48), 6.0, 0.35, 42)
    b = pkg.geometry.sphere_pack((48, 48, 48), 6.0, 0.35, 42)
    assert np.array_equal(a, b)
    assert abs(a.mean() - 0.35) < 0.08


def test_two_phase_setup_matches_reference_initial_density():
    g = helpers.Golden("twophase_d3q19_p2")
    lg, tabs = helpers.build_tables(g)
    pkg = helpers.load_package()
    setup = pkg.cases.two_phase_setup(lg, tabs, g.attr("rho0"), g.attr("rho1"), g.attr("wettability"))
    for r, t in enumerate(tabs):
        ref = g.rec(r, "rhoInit").reshape(-1, 2)
        # rows the main never initialises differently: compare own bulk and solid boundary rows
        rows = np.concatenate([t.bulk_nodes(), t.solid_bnd_nodes()])
        assert np.array_equal(ref[rows], setup[r]["rho"][rows])


REFERENCE_SCRIPTS = "/root/reference/PythonScripts"


@pytest.mark.skipif(not os.path.isdir(REFERENCE_SCRIPTS), reason="differential test against the reference's own vtklb.py (build container only)")
@pytest.mark.parametrize("seed", range(12))
def test_vtklb_files_equal_the_reference_writer_on_random_geometries(seed, tmp_path):
    """node numbering, neighbour table, ghost / solid-boundary node selection, PROCESSOR pairs and node types of
    random 2-D / 3-D geometries with 1-4 ranks and every periodicity: the product's .vtklb files are byte-identical
    to the ones the reference's PythonScripts/vtklb.py writes for the same input"""
    import contextlib
    import io
    import sys
    sys.path.insert(0, REFERENCE_SCRIPTS)
    from vtklb import vtklb  # noqa: the reference's writer
    pkg = helpers.load_package()
    rng = np.random.default_rng(100 + seed)
    nd = 2 if seed % 3 == 0 else 3
    shape = tuple(int(x) for x in rng.integers(4, 9, size=nd))
    lattice = "D2Q9" if nd == 2 else ("D3Q19" if seed % 2 else "D3Q27")
    periodic = ["", "x", "y", "xy"][seed % 4] if nd == 2 else ["", "x", "yz", "xyz", "z", "xz"][seed % 6]
    fluid = rng.random(shape) < 0.7
    nranks = int(rng.integers(1, 5))
    # ranks as irregular slabs along a random axis (every rank keeps at least one layer)
    ax = int(rng.integers(0, nd))
    cuts = np.sort(rng.choice(np.arange(1, shape[ax]), size=min(nranks - 1, shape[ax] - 1), replace=False)) if nranks > 1 else []
    owner = np.zeros(shape[ax], dtype=int)
    for c in cuts:
        owner[c:] += 1
    sh = [1] * nd
    sh[ax] = -1
    geo = np.where(fluid, owner.reshape(sh) + 1, 0).astype(int)
    present = sorted(set(np.unique(geo)) - {0})
    if present != list(range(1, len(present) + 1)):   # a rank without fluid nodes: renumber densely
        remap = {old: new + 1 for new, old in enumerate(present)}
        geo = np.vectorize(lambda v: remap.get(v, 0))(geo).astype(int)
    if geo.max() == 0:
        pytest.skip("no fluid")
    attr = rng.random(shape)
    basis = lattice if lattice != "D3Q27" else pkg.geometry.BASIS["D3Q27"].astype(int)
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    with contextlib.redirect_stdout(io.StringIO()):
        v = vtklb(geo, basis, periodic, "tmp", str(ref_dir) + "/")
        v.append_data_set("init_rho", attr)
    lg = pkg.geometry.LatticeGeometry(geo, lattice, periodic)
    tabs = lg.all_ranks()
    assert len(tabs) == int(geo.max())
    for r, t in enumerate(tabs):
        mine = tmp_path / ("mine%d.vtklb" % r)
        t.write_vtklb(str(mine), {"init_rho": attr})
        assert open(mine).read() == open(ref_dir / ("tmp%d.vtklb" % r)).read(), (seed, lattice, shape, periodic, r)


REF_DRIVER = os.path.join(helpers.ROOT, "oracle", "_ref", "ref_driver")


@pytest.mark.skipif(not os.path.exists(REF_DRIVER), reason="needs oracle/_ref/ref_driver (the reference's own headers)")
@pytest.mark.parametrize("seed", range(8))
def test_node_tables_equal_the_reference_classes_on_random_geometries(seed, tmp_path):
    """random geometries through the reference's own LBvtk / Grid / Nodes / BndMpi / HalfWayBounceBack (oracle/_ref/ref_driver,
    table dump only) against the in-memory ingest: neighbour lists, node types (incl. the types ghost nodes take from
    their owner), bulk / boundary lists, bounce-back link classes and the MonLatMpi send / receive lists, and against
    the C++ host mirror (host/apps/dump_tables) reading the same files"""
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(helpers.ROOT, "oracle"))
    from recfile import read_rec
    pkg = helpers.load_package()
    rng = np.random.default_rng(500 + seed)
    nd = 2 if seed % 4 == 0 else 3
    lattice = "D2Q9" if nd == 2 else "D3Q19"
    shape = tuple(int(x) for x in rng.integers(5, 10, size=nd))
    periodic = ["xy", "x", "", "y"][seed % 4] if nd == 2 else ["xyz", "xy", "z", "", "xz", "yz"][seed % 6]
    fluid = rng.random(shape) < 0.72
    nranks = int(rng.integers(1, 4))
    owner = np.minimum(np.arange(shape[-1]) * nranks // shape[-1], nranks - 1)
    geo = np.where(fluid, owner.reshape([1] * (nd - 1) + [-1]) + 1, 0).astype(int)
    if sorted(set(np.unique(geo)) - {0}) != list(range(1, nranks + 1)):
        pytest.skip("a rank without fluid nodes")
    lg = pkg.geometry.LatticeGeometry(geo, lattice, periodic)
    tabs = lg.all_ranks()
    for t in tabs:
        t.write_vtklb(str(tmp_path / ("tmp%d.vtklb" % t.my_rank)), {"init_rho": np.ones(shape)})
    out = tmp_path / "out"
    out.mkdir()
    subprocess.run([REF_DRIVER, "--case", "std_case", "--lattice", lattice, "--dir", str(tmp_path), "--out", str(out), "--nranks",
                    str(nranks), "--steps", "0", "--no-f"], check=True, capture_output=True)
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests"))
    import test_host_cpp
    exe = test_host_cpp.build("dump_tables", link_engine=False)
    for r, t in enumerate(tabs):
        rec = read_rec(str(out / ("rank%d.rec" % r)))
        assert np.array_equal(rec["neigh"].reshape(-1, lg.nq), t.neigh)
        assert np.array_equal(rec["type"], t.node_type)
        assert np.array_equal(rec["rank"], t.node_rank)
        assert np.array_equal(rec["bulk"], t.bulk_nodes())
        assert np.array_equal(rec["fluidBnd"], t.fluid_bnd_nodes())
        assert np.array_equal(rec["solidBnd"], t.solid_bnd_nodes())
        nodes, nb, ng, ndl, links = t.halfway_bb(t.fluid_bnd_nodes())
        assert np.array_equal(rec["bb.node"], nodes) and np.array_equal(rec["bb.links"], links.ravel())
        assert np.array_equal(rec["bb.nBeta"], nb) and np.array_equal(rec["bb.nGamma"], ng) and np.array_equal(rec["bb.nDelta"], ndl)
        ss = t.send_side(tabs)
        assert int(rec["nNeigRanks"][0]) == len(t.neig_ranks)
        for k, nr in enumerate(t.neig_ranks):
            p = "mpi%d." % k
            assert int(rec[p + "neigRank"][0]) == nr
            assert np.array_equal(rec[p + "nodesReceived"], t.recv_nodes[k]) and np.array_equal(rec[p + "dirListReceived"], t.recv_dirs[k])
            assert np.array_equal(rec[p + "nodesToSend"], ss[k][0]) and np.array_equal(rec[p + "dirListToSend"], ss[k][2])
        res = subprocess.run([exe, lattice, str(tmp_path / "tmp"), str(r)], capture_output=True, text=True, check=True)
        d = test_host_cpp.parse_dump(res.stdout)
        for key in ("neigh", "type", "rank", "bulk", "fluidBnd", "solidBnd", "bb.node", "bb.nBeta", "bb.nGamma", "bb.nDelta", "bb.links"):
            assert np.array_equal(d[key], rec[key]), key
