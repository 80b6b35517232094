"""The reference's library pressure boundaries, PressureBnd<DXQY> and InletOutlet<DXQY> (LBpressurebnd.h:10-88).
No main of the reference calls them; the goldens pbnd_* / inout_* come from the reference's own classes applied after the
bounce back of every std_case step on every k-th fluid boundary node (oracle/ref_driver --pressure-bnd).

CPU: the boundary classes' beta / delta lists equal the product's (Boundary<DXQY> classes pairs like the bounce-back
helper), the oracle port's restatement of both apply() functions reproduces the reference's dumps bit for bit (1 and 2
ranks), and the values the product hands to the engine are the ones the reference stored.
GPU: the engine with those constant links (chimp_add_constant_links) reproduces the same dumps bit for bit."""
import numpy as np
import pytest

import helpers

CASES = ["pbnd_d3q19_p1", "inout_d3q19_p2", "inout_d2q9_p1"]


def _options(g):
    e = [str(x) for x in g.z["extra"]]
    opt = {e[i].lstrip("-"): e[i + 1] for i in range(0, len(e), 2)}
    kind = opt["pressure-bnd"]
    every = int(opt.get("bnd-every", 3))
    rho = float(opt.get("io-rho", 1.02))
    vel = [float(x) for x in opt.get("io-vel", "0.01,-0.005,0.002").split(",")]
    return kind, every, rho, vel


def _bnd_nodes(t, every):
    return t.fluid_bnd_nodes()[::every]


def _rho_bnd(t):
    """ref_driver: rhoBnd(0, n) = 1.0 + 0.01 * (n % 7)"""
    return (1.0 + 0.01 * (np.arange(t.size) % 7)).reshape(-1, 1)


@pytest.mark.parametrize("name", CASES)
def test_boundary_lists_match_the_reference_classes(name):
    g = helpers.Golden(name)
    _, every, _, _ = _options(g)
    lg, tabs = helpers.build_tables(g)
    for r, t in enumerate(tabs):
        nodes, n_beta, n_gamma, n_delta, links = t.halfway_bb(_bnd_nodes(t, every))
        assert np.array_equal(nodes, g.rec(r, "pbnd.nodes"))
        assert np.array_equal(n_beta, g.rec(r, "pbnd.nBeta"))
        assert np.array_equal(n_delta, g.rec(r, "pbnd.nDelta"))
        flat = []
        for b in range(len(nodes)):
            flat += list(links[b, :n_beta[b]]) + list(links[b, n_beta[b] + n_gamma[b]:n_beta[b] + n_gamma[b] + n_delta[b]])
        assert np.array_equal(np.array(flat, dtype=np.int64), g.rec(r, "pbnd.links"))


def _port_ranks(g, tabs):
    port = helpers.oracle_port()
    pkg = helpers.load_package()
    kind, every, rho, vel = _options(g)
    ranks = []
    for t in tabs:
        pr = port.PortRank(pkg.geometry.LATTICE_ID[g.lattice], t.neigh, t.bulk_nodes(), 1, t.halfway_bb(t.fluid_bnd_nodes()))
        pr.f[:] = pkg.cases.std_case_initial_state(t, g.attr("init_rho"))[0]
        pr.set_library_bnd(kind, t.halfway_bb(_bnd_nodes(t, every)), rho_bnd=_rho_bnd(t), rho=rho, vel=vel)
        ranks.append(pr)
    return port, ranks


@pytest.mark.parametrize("name", CASES)
def test_port_reproduces_the_reference_with_library_boundaries(name):
    g = helpers.Golden(name)
    lg, tabs = helpers.build_tables(g)
    port, ranks = _port_ranks(g, tabs)
    exch = helpers.exchange_lists(tabs)
    a = g.args
    done = 0
    for step in [s for s in g.dump if s > 0]:
        for _ in range(step - done):
            if len(ranks) == 1:
                ranks[0].step_std_case(1, tau=a["tau"], force=g.force())
            else:
                for pr in ranks:
                    pr.step_std_case(1, tau=a["tau"], force=g.force(), skip_boundary=True)
                port.exchange_lb_field(ranks, exch, 0)
                for pr in ranks:
                    pr.apply_bb(0)
                    pr.apply_library_bnd(0)
        done = step
        for r, (t, pr) in enumerate(zip(tabs, ranks)):
            bulk = t.bulk_nodes()
            assert np.array_equal(pr.f[bulk], g.f(r, step)[bulk]), "rank %d step %d" % (r, step)
            assert np.array_equal(pr.rho[bulk, 0], g.rec(r, "step%d.rho" % step)[bulk])


@pytest.mark.parametrize("name", CASES)
def test_link_values_are_the_ones_the_reference_stored(name):
    """cases.library_bnd_links forms w[q] * rho resp. the prescribed equilibrium with the reference's operations: after
    step 1 the reference's field holds exactly these numbers at the destinations (last store wins where two boundary
    nodes write the same place)"""
    g = helpers.Golden(name)
    pkg = helpers.load_package()
    kind, every, rho, vel = _options(g)
    lg, tabs = helpers.build_tables(g)
    for r, t in enumerate(tabs):
        node_q, values = pkg.cases.library_bnd_links(t, _bnd_nodes(t, every), kind, rho_bnd=_rho_bnd(t), rho=rho, vel=vel)
        assert len(values) > 10
        last = {}
        for (n, q), v in zip(node_q.tolist(), values.tolist()):
            last[(n, q)] = v
        f1 = g.f(r, 1)
        own = set(t.bulk_nodes().tolist())
        checked = 0
        for (n, q), v in last.items():
            if n in own:
                assert f1[n, 0, q] == v, (n, q)
                checked += 1
        assert checked > 10


@pytest.mark.parametrize("boundary_first", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_pull_table_with_constant_links_reproduces_the_reference_step(name, boundary_first):
    """No GPU: the pull table the engine's host builder makes of bounce back + library boundary (+ the 2-rank exchange),
    evaluated in numpy over what the oracle port's nodes push in their second step, yields the field the reference holds
    after that step's exchange, bounce back and apply() -- the constants sit in halo-in slots nobody sends to."""
    g = helpers.Golden(name)
    pkg = helpers.load_package()
    kind, every, rho, vel = _options(g)
    lg, tabs = helpers.build_tables(g)
    port, ranks = _port_ranks(g, tabs)
    exch = helpers.exchange_lists(tabs)
    a = g.args

    def full_step(skip_last_boundary):
        for pr in ranks:
            pr.step_std_case(1, tau=a["tau"], force=g.force(), skip_boundary=True)
        if skip_last_boundary:
            return
        port.exchange_lb_field(ranks, exch, 0)
        for pr in ranks:
            pr.apply_bb(0)
            pr.apply_library_bnd(0)

    full_step(False)
    full_step(True)                      # second step stopped after push + swap: pushed values sit in the neighbours' rows
    nq = lg.nq
    rev = np.array([pkg.geometry.reverse_direction(g.lattice, q) for q in range(nq)])
    lats, X, tables = [], [], []
    for t, pr in zip(tabs, ranks):
        lat = pkg.capi.Lattice.from_rank_tables(t)
        ss = t.send_side(tabs)
        for k, nr in enumerate(t.neig_ranks):
            lat.add_neighbor(nr, ss[k][0], ss[k][1], ss[k][2], t.recv_nodes[k], t.recv_ndir[k], t.recv_dirs[k])
        lat.add_halfway_bb(*t.halfway_bb(t.fluid_bnd_nodes()))
        node_q, values = pkg.cases.library_bnd_links(t, _bnd_nodes(t, every), kind, rho_bnd=_rho_bnd(t), rho=rho, vel=vel)
        lat.add_constant_links(node_q, values)
        lat.build_host(boundary_first)
        table, labels, pmask, info = lat.host_table()
        x = np.zeros((nq, info["stride"]))
        for q in range(nq):
            x[q, :info["n"]] = pr.f[t.neigh[labels, q], 0, q]
        dst, val = lat.host_constant_links()
        assert len(val) == len(values) and np.array_equal(val, values)
        assert len(set(dst.tolist())) == len(dst) and (dst % info["stride"] >= info["n_pad"]).all()
        x.ravel()[dst] = val
        lats.append(lat); X.append(x); tables.append((table, labels, info))
    packed = {}
    for r, lat in enumerate(lats):
        for k in range(lat.num_neighbors()):
            nr, src, dst = lat.host_halo_lists(k)
            packed[(r, nr)] = X[r].ravel()[src]
    for r, lat in enumerate(lats):
        for k in range(lat.num_neighbors()):
            nr, src, dst = lat.host_halo_lists(k)
            X[r].ravel()[dst] = packed[(nr, r)]
    port.exchange_lb_field(ranks, exch, 0)
    for pr in ranks:
        pr.apply_bb(0)
        pr.apply_library_bnd(0)
    for r, pr in enumerate(ranks):
        table, labels, info = tables[r]
        i = np.arange(info["n"])
        for q in range(nq):
            src = table[q]
            pulled = np.where(src >= 0, X[r][q, np.maximum(src, 0)], X[r][rev[q], i])
            assert np.array_equal(pulled, pr.f[labels, 0, q]), "rank %d direction %d" % (r, q)


def _engine_ranks(g, tabs, boundary_first):
    pkg = helpers.load_package()
    kind, every, rho, vel = _options(g)
    lats = []
    for t in tabs:
        lat = pkg.capi.Lattice.from_rank_tables(t)
        ss = t.send_side(tabs)
        for k, nr in enumerate(t.neig_ranks):
            lat.add_neighbor(nr, ss[k][0], ss[k][1], ss[k][2], t.recv_nodes[k], t.recv_ndir[k], t.recv_dirs[k])
        lat.add_halfway_bb(*t.halfway_bb(t.fluid_bnd_nodes()))
        lat.add_constant_links(*pkg.cases.library_bnd_links(t, _bnd_nodes(t, every), kind, rho_bnd=_rho_bnd(t), rho=rho, vel=vel))
        lat.finalize(1, boundary_first)
        lat.upload(pkg.cases.std_case_initial_state(t, g.attr("init_rho"))[0])
        lats.append(lat)
    return lats


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_with_library_boundaries_bit_exact_vs_reference(name):
    from test_gpu_parity import InProcessRanks
    g = helpers.Golden(name)
    lg, tabs = helpers.build_tables(g)
    lats = _engine_ranks(g, tabs, len(tabs) > 1)
    ranks = InProcessRanks(lats) if len(lats) > 1 else None
    a = g.args
    done = 0
    for step in [s for s in g.dump if s > 0]:
        if ranks:
            ranks.step(step - done, tau=a["tau"], force=g.force())
        else:
            lats[0].step_single(step - done, tau=a["tau"], force=g.force())
        done = step
        for r, (lat, t) in enumerate(zip(lats, tabs)):
            bulk = t.bulk_nodes()
            assert np.array_equal(lat.download()[bulk], g.f(r, step)[bulk]), "rank %d step %d" % (r, step)
            assert np.array_equal(lat.download_rho()[bulk, 0], g.rec(r, "step%d.rho" % step)[bulk])
    for lat in lats:
        lat.close()


@pytest.mark.gpu
def test_gpu_constant_links_survive_upload_and_initialisation():
    """an upload scatters the state through the pull table and an initialisation fills whole planes: both must leave
    the constant links' values in place (a restart in the middle of the run continues to the reference's later dump)"""
    g = helpers.Golden("pbnd_d3q19_p1")
    lg, tabs = helpers.build_tables(g)
    lat, t = _engine_ranks(g, tabs, False)[0], tabs[0]
    a = g.args
    bulk = t.bulk_nodes()
    lat.init_uniform(1.0)                       # whole planes rewritten ...
    lat.upload(g.f(0, 2))                       # ... then the reference's state after step 2 uploaded
    lat.step_single(4, tau=a["tau"], force=g.force())
    assert np.array_equal(lat.download()[bulk], g.f(0, 6)[bulk])
    lat.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,kind", [("pbnd_d3q19_p1", "pressure"), ("inout_d3q19_p2", "inletoutlet")])
def test_cpp_mirror_classes_on_the_engine_match_the_reference(name, kind, tmp_path):
    """host/chimp/LBpressurebnd.h (PressureBnd / InletOutlet of the C++ mirror) handed to GpuLattice::add in the std_case
    application: the reference's dumps bit for bit, one rank and two in-process ranks"""
    import test_host_cpp as H
    exe = H.build("std_case", link_engine=True)
    g = helpers.Golden(name)
    _, tabs = helpers.build_tables(g)
    step = max(g.dump)
    F = g.force()
    deck = tmp_path / "input.dat"
    deck.write_text("<iterations>\n  max %d\n  write %d\n<end>\n<fluid>\n  tau %r\n  bodyforce %r %r %r\n<end>\n"
                    % (step, step, g.args["tau"], F[0], F[1], F[2]))
    prefix = H.write_case_files(g, tabs, tmp_path, {"init_rho": g.attr("init_rho")})
    out = tmp_path / "out.bin"
    import subprocess
    subprocess.run([exe, g.lattice, str(deck), prefix, "0", str(out), str(g.nranks), "-", kind], check=True)
    res = H.read_app_output(out, 19, 3)
    assert len(res) == g.nranks
    for r, (f, rho, vel) in enumerate(res):
        bulk = tabs[r].bulk_nodes()
        assert np.array_equal(f[bulk], g.f(r, step)[bulk, 0])
        assert np.array_equal(rho[bulk], g.rec(r, "step%d.rho" % step)[bulk])
