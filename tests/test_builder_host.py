"""Host logic of the engine (no GPU): the symbolic table builder behind chimp_build_host.

The pull table T and the halo lists are validated against the oracle: one reference iteration
is run with the oracle port up to the push+swap (skip_boundary), which exposes every node's
post-collision populations X; pulling X through T (and through the neighbours' send lists
for halo slots) must reproduce, bit for bit, the state the oracle reaches after its own ghost
exchange and boundary copies."""
import numpy as np
import pytest

import helpers
from test_oracle import run_port_case


def build_engine_tables(g, lg, tabs, boundary_first):
    pkg = helpers.load_package()
    capi = pkg.capi
    lats = []
    setup = helpers.one_phase_setup(g, lg, tabs) if g.case == "one_phase" else None
    nf = 2 if g.case == "twophase" else 1
    for r, t in enumerate(tabs):
        lat = capi.Lattice.from_rank_tables(t, n_fields=nf)
        ss = t.send_side(tabs)
        for k, nr in enumerate(t.neig_ranks):
            lat.add_neighbor(nr, ss[k][0], ss[k][1], ss[k][2], t.recv_nodes[k], t.recv_ndir[k], t.recv_dirs[k])
        if g.case == "std_case":
            lat.add_halfway_bb(*t.halfway_bb(t.fluid_bnd_nodes()))
        elif g.case == "twophase":
            lat.add_halfway_bb(*t.halfway_bb(t.bulk_nodes()))
        else:
            lat.add_links(capi.LINK_SOLID, setup[r]["solid_links"])
            lat.add_links(capi.LINK_PRESSURE, setup[r]["press_links"])
            lat.add_links(capi.LINK_FLUID_SWAP, setup[r]["fluid_links"])
        lat.build_host(boundary_first)
        lats.append(lat)
    return lats


CASES = [n for n in helpers.all_golden_names() if helpers.Golden(n).case != "twophase"]


@pytest.mark.parametrize("boundary_first", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_pull_table_reproduces_oracle_step(name, boundary_first):
    g = helpers.Golden(name)
    if g.case == "one_phase" and g.nranks > 1:
        pytest.skip("N-rank one_phase oracle replay not available on the CPU")
    port = helpers.oracle_port()
    pkg = helpers.load_package()
    tabs, ranks = run_port_case(g, 1)               # state after one full step
    lg = tabs[0].g
    lats = build_engine_tables(g, lg, tabs, boundary_first)
    # second step, stopped after push + swap: pushed values sit in the neighbours' rows
    a = g.args
    trt = tuple(a["trt"]) if "trt" in a else None
    before = [pr.f.copy() for pr in ranks]
    for pr in ranks:
        if g.case == "std_case":
            pr.step_std_case(1, tau=a.get("tau", 0.8), force=g.force(), trt=trt, skip_boundary=True)
        else:
            pr.step_one_phase(1, tau=a.get("tau", 0.8), force=g.force(), trt=trt, skip_boundary=True)
    nq = lg.nq
    rev = np.array([pkg.geometry.reverse_direction(g.lattice, q) for q in range(nq)])
    w = pkg.cases.lattice_weights(g.lattice)
    basis = pkg.geometry.BASIS[g.lattice]
    X, infos, tables = [], [], []
    for t, pr, lat in zip(tabs, ranks, lats):
        table, labels, pmask, info = lat.host_table()
        stride = info["stride"]
        x = np.zeros((nq, stride))
        for q in range(nq):
            x[q, :info["n"]] = pr.f[t.neigh[labels, q], 0, q]   # X[q][i] = value node i pushed along q
        # anti-bounce-back slots carry the boundary value instead (std_one_phase/main.cpp:168-172)
        for q in range(nq):
            sel = np.nonzero((pmask >> np.uint32(q)) & np.uint32(1))[0]
            if len(sel) == 0:
                continue
            v = pr.vel[labels[sel]]
            u2 = v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]
            if lg.nd == 3:
                u2 = u2 + v[:, 2] * v[:, 2]
            cu = basis[q, 0] * v[:, 0]
            for d in range(1, lg.nd):
                cu = cu + basis[q, d] * v[:, d]
            x[q, sel] = -x[q, sel] + 2 * w[q] * a.get("rhow", 1.0) * (1 + 0.5 * (9.0 * cu * cu - 3.0 * u2))
        X.append(x)
        infos.append(info)
        tables.append((table, labels))
    # halo transport: neighbour's packed send list -> my halo-in slots
    packed = {}
    for r, lat in enumerate(lats):
        for k in range(lat.num_neighbors()):
            nr, src, dst = lat.host_halo_lists(k)
            packed[(r, nr)] = X[r].ravel()[src]
    for r, lat in enumerate(lats):
        for k in range(lat.num_neighbors()):
            nr, src, dst = lat.host_halo_lists(k)
            flat = X[r].ravel()
            flat[dst] = packed[(nr, r)]
    # reference result of the same step
    exch = helpers.exchange_lists(tabs)
    port.exchange_lb_field(ranks, exch, 0)
    for pr in ranks:
        if g.case == "std_case":
            pr.apply_bb(0)
        else:
            pr.apply_one_phase_links()
    for r, (t, pr) in enumerate(zip(tabs, ranks)):
        table, labels = tables[r]
        n = infos[r]["n"]
        assert sorted(labels.tolist()) == t.bulk_nodes().tolist()
        i = np.arange(n)
        for q in range(nq):
            src = table[q]
            pulled = np.where(src >= 0, X[r][q, np.maximum(src, 0)], X[r][rev[q], i])
            assert np.array_equal(pulled, pr.f[labels, 0, q]), "rank %d direction %d" % (r, q)
        if boundary_first and g.nranks > 1:
            nb = infos[r]["n_boundary"]
            assert 0 < nb <= n
            halo_users = np.nonzero((table >= infos[r]["n_pad"]).any(axis=0))[0]
            assert halo_users.max() < nb
    del before


def test_open_boundary_is_rejected():
    pkg = helpers.load_package()
    geo = np.ones((4, 4, 4), dtype=int)
    lg = pkg.geometry.LatticeGeometry(geo, "D3Q19", "xy")     # z is open: pulls through the dummy node
    t = lg.all_ranks()[0]
    lat = pkg.capi.Lattice.from_rank_tables(t)
    lat.add_halfway_bb(*t.halfway_bb(t.fluid_bnd_nodes()))
    with pytest.raises(pkg.capi.ChimpError):
        lat.build_host()


def test_bad_arguments_are_reported():
    pkg = helpers.load_package()
    capi = pkg.capi
    neigh = np.zeros((3, 19), dtype=np.int32)
    with pytest.raises(capi.ChimpError):
        capi.Lattice("D3Q19", neigh, [2, 1])          # not ascending
    neigh[1, 0] = 7
    with pytest.raises(capi.ChimpError):
        capi.Lattice("D3Q19", neigh, [1, 2])          # neighbour out of range


def test_two_rank_one_phase_tables_build():
    """fluid-fluid swap links that touch ghost rows write values no own node ever pulls: the builder
    must accept them (the 2-rank run itself is checked against the reference on the GPU)"""
    g = helpers.Golden("onephase_trt_d3q19_p2")
    lg, tabs = helpers.build_tables(g)
    lats = build_engine_tables(g, lg, tabs, True)
    for lat, t in zip(lats, tabs):
        table, labels, pmask, info = lat.host_table()
        assert info["n"] == len(t.bulk_nodes()) and (pmask != 0).any()
