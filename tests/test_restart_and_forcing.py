"""SURVEY 8(f4) restart files and 8(f2) caller-side global forcing, pinned to the reference itself:

* tests/golden/restart_d3q19_p1.ckpt0.* were written by the reference's own LbField::writeToFile()
  (LBfield.h:378-392) after step 4 of a 10-step std_case run (oracle/gen_golden.py round2); continuing from that
  file must reproduce the reference's dump of step 10 bit for bit -- with the oracle port here, with the CUDA path in
  the gpu-marked tests below.  The reverse direction hands a file written by this repository to the reference's own
  LbField::readFromFile() (ref_driver --restart).
* the fluxForce / capForce records of the goldens are return values of the reference's own calcFluxForceCartDir /
  calcCapNumbForceCartDir (LBglobalforcing.h:8-98) called from ref_driver --global-forcing."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers
from test_oracle import run_port_case

REF_DRIVER = os.path.join(helpers.ROOT, "oracle", "_ref", "ref_driver")
CKPT = os.path.join(helpers.GOLDEN, "restart_d3q19_p1.ckpt0")


def _restart_case():
    g = helpers.Golden("restart_d3q19_p1")
    base = helpers.Golden("std_d3q19_p1")      # same case; its .vtklb file is committed
    assert np.array_equal(g.geo, base.geo) and np.array_equal(g.attr("init_rho"), base.attr("init_rho"))
    return g


def test_reference_restart_file_holds_the_dump_of_that_step():
    pkg = helpers.load_package()
    g = _restart_case()
    f4 = pkg.checkpoint.read_lbfield(CKPT)
    assert np.array_equal(f4, g.f(0, 4))
    assert np.array_equal(pkg.checkpoint.read_scalar_field(CKPT)[:, 0], g.rec(0, "step4.rho"))


def test_port_continues_from_the_reference_restart_file():
    pkg = helpers.load_package()
    port = helpers.oracle_port()
    g = _restart_case()
    lg, tabs = helpers.build_tables(g)
    t = tabs[0]
    pr = port.PortRank(pkg.geometry.LATTICE_ID[g.lattice], t.neigh, t.bulk_nodes(), 1, t.halfway_bb(t.fluid_bnd_nodes()))
    pr.f[:] = pkg.checkpoint.read_lbfield(CKPT)
    pr.step_std_case(6, tau=g.args["tau"], force=g.force())
    bulk = t.bulk_nodes()
    assert np.array_equal(pr.f[bulk], g.f(0, 10)[bulk])
    assert np.array_equal(pr.rho[bulk, 0], g.rec(0, "step10.rho")[bulk])


def test_port_flux_force_equals_the_reference_function():
    port = helpers.oracle_port()
    g = _restart_case()
    for step in (4, 10):
        tabs, ranks = run_port_case(g, step)
        n = len(tabs[0].bulk_nodes())
        want = g.rec(0, "step%d.fluxForce" % step)
        got = [port.flux_force(ranks, 0, d, 1e-5, n) for d in range(3)]
        assert np.array_equal(np.array(got), want), (step, got, want)


def test_port_twophase_forcing_equals_the_reference_functions():
    port = helpers.oracle_port()
    g = helpers.Golden("forcing_twophase_d3q19_p1")
    for step in g.dump:
        tabs, ranks = run_port_case(g, step)
        n = len(tabs[0].bulk_nodes())
        for fld in (0, 1):
            want = g.rec(0, "step%d.fluxForce%d" % (step, fld))
            got = np.array([port.flux_force(ranks, fld, d, 2e-5, n) for d in range(3)])
            assert np.array_equal(got, want), (step, fld, got, want)
        want = g.rec(0, "step%d.capForce" % step)
        got = np.array([port.cap_numb_force(ranks, d, 1e-4, 0.1666666666666666574, 0.1, n) for d in range(3)])
        assert np.array_equal(got, want), (step, got, want)


def _write_restart_for_reference(pkg, tmp_path, f):
    """file in the reference's .lblbf layout + the committed geometry file, where ref_driver expects them"""
    import shutil
    shutil.copy(os.path.join(helpers.GOLDEN, "std_d3q19_p1.tmp0.vtklb"), str(tmp_path / "tmp0.vtklb"))
    pkg.checkpoint.write_lbfield(str(tmp_path / "mine0"), f)
    return str(tmp_path / "mine")


def _reference_continues(tmp_path, prefix, g, steps):
    sys.path.insert(0, os.path.join(helpers.ROOT, "oracle"))
    from recfile import read_rec
    cmd = [REF_DRIVER, "--case", "std_case", "--lattice", "D3Q19", "--dir", str(tmp_path), "--out", str(tmp_path), "--nranks", "1",
           "--steps", str(steps), "--dump", str(steps), "--no-tables", "--tau", repr(g.args["tau"]),
           "--force", ",".join(repr(x) for x in g.force()), "--restart", prefix]
    subprocess.run(cmd, check=True, capture_output=True)
    rec = read_rec(str(tmp_path / "rank0.rec"))
    return rec["step%d.f" % steps].reshape(-1, 1, 19)


@pytest.mark.skipif(not os.path.exists(REF_DRIVER), reason="needs oracle/_ref/ref_driver (the reference's own headers)")
def test_reference_restarts_from_a_file_written_by_the_port(tmp_path):
    """the harness of the gpu-marked reverse test, exercised with the port as the writer"""
    pkg = helpers.load_package()
    g = _restart_case()
    tabs, ranks = run_port_case(g, 4)
    prefix = _write_restart_for_reference(pkg, tmp_path, ranks[0].f)
    f10 = _reference_continues(tmp_path, prefix, g, 6)
    bulk = tabs[0].bulk_nodes()
    assert np.array_equal(f10[bulk], g.f(0, 10)[bulk])


# ---- the CUDA path --------------------------------------------------------------------------------------------
def _engine(pkg, g, t, index_form=1):
    lat = pkg.capi.Lattice.from_rank_tables(t)
    lat.add_halfway_bb(*t.halfway_bb(t.fluid_bnd_nodes()))
    lat.finalize(index_form)
    return lat


@pytest.mark.gpu
@pytest.mark.parametrize("index_form", [0, 1])
def test_gpu_continues_from_the_reference_restart_file(index_form):
    """reference-written .lblbf of step 4 -> upload -> 6 steps on the GPU == the reference's dump of step 10, bit for bit"""
    pkg = helpers.load_package()
    g = _restart_case()
    lg, tabs = helpers.build_tables(g)
    t = tabs[0]
    lat = _engine(pkg, g, t, index_form)
    lat.upload(pkg.checkpoint.read_lbfield(CKPT, expect=(1, 19, t.size)))
    lat.step_single(6, tau=g.args["tau"], force=g.force())
    bulk = t.bulk_nodes()
    assert np.array_equal(lat.download()[bulk], g.f(0, 10)[bulk])
    assert np.array_equal(lat.download_rho()[bulk, 0], g.rec(0, "step10.rho")[bulk])
    assert np.array_equal(lat.download_vel()[bulk], g.rec(0, "step10.vel").reshape(-1, 3)[bulk])
    lat.close()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_DRIVER), reason="needs oracle/_ref/ref_driver (the reference's own headers)")
def test_reference_restarts_from_a_file_written_by_the_gpu(tmp_path):
    """4 steps on the GPU -> .lblbf -> the reference's own LbField::readFromFile() + 6 reference steps == the
    reference's uninterrupted run at step 10 (and the GPU's own step 10), bit for bit"""
    pkg = helpers.load_package()
    g = _restart_case()
    lg, tabs = helpers.build_tables(g)
    t = tabs[0]
    lat = _engine(pkg, g, t)
    lat.upload(pkg.cases.std_case_initial_state(t, g.attr("init_rho"))[0])
    lat.step_single(4, tau=g.args["tau"], force=g.force())
    f4 = lat.download()
    bulk = t.bulk_nodes()
    assert np.array_equal(f4[bulk], g.f(0, 4)[bulk])
    # rows the engine does not own (wall / dummy rows) are whatever the caller's array held: zeros here, while the
    # reference keeps stale pushes there; they are never read before being overwritten (invariants)
    prefix = _write_restart_for_reference(pkg, tmp_path, f4)
    f10 = _reference_continues(tmp_path, prefix, g, 6)
    assert np.array_equal(f10[bulk], g.f(0, 10)[bulk])
    lat.step_single(6, tau=g.args["tau"], force=g.force())
    assert np.array_equal(lat.download()[bulk], f10[bulk])
    lat.close()


@pytest.mark.gpu
def test_gpu_flux_force_equals_the_reference_function():
    """chimp_flux_force against calcFluxForceCartDir run by the reference itself (fixed-shape tree sum on the device
    against the reference's sequential sum: <= 1e-12 relative)"""
    pkg = helpers.load_package()
    g = _restart_case()
    lg, tabs = helpers.build_tables(g)
    t = tabs[0]
    lat = _engine(pkg, g, t)
    lat.upload(pkg.cases.std_case_initial_state(t, g.attr("init_rho"))[0])
    n = len(t.bulk_nodes())
    done = 0
    for step in (4, 10):
        lat.step_single(step - done, tau=g.args["tau"], force=g.force())
        done = step
        want = g.rec(0, "step%d.fluxForce" % step)
        for d in range(3):
            got = lat.flux_force(0, d, 1e-5, n)
            assert abs(got - want[d]) <= 1e-12 * abs(want[d]), (step, d, got, want[d])
    lat.close()


@pytest.mark.gpu
def test_gpu_twophase_forcing_equals_the_reference_functions():
    """chimp_flux_force (both fields) and chimp_capillary_force against calcFluxForceCartDir / calcCapNumbForceCartDir
    run by the reference itself on the same two-phase run"""
    pkg = helpers.load_package()
    g = helpers.Golden("forcing_twophase_d3q19_p1")
    lg, tabs = helpers.build_tables(g)
    t = tabs[0]
    s = pkg.cases.two_phase_setup(lg, tabs, g.attr("rho0"), g.attr("rho1"), g.attr("wettability"))[0]
    lat = pkg.capi.Lattice.from_rank_tables(t, n_fields=2)
    lat.add_halfway_bb(*t.halfway_bb(t.bulk_nodes()))
    lat.set_solid_boundary(s["solid_bnd"])
    lat.finalize(1)
    lat.set_twophase_density(s["rho"])
    lat.upload(s["f0"])
    a = g.args
    n = len(t.bulk_nodes())
    done = 0
    for step in g.dump:
        lat.step_twophase(step - done, a["tau2"][0], a["tau2"][1], a["sigma"], a["beta"], a["momx"], g.force(), n)
        done = step
        for fld in (0, 1):
            want = g.rec(0, "step%d.fluxForce%d" % (step, fld))
            for d in range(3):
                got = lat.flux_force(fld, d, 2e-5, n)
                assert abs(got - want[d]) <= 1e-10 * abs(want[d]), (step, fld, d, got, want[d])
        want = g.rec(0, "step%d.capForce" % step)
        for d in range(3):
            got = lat.capillary_force(d, 1e-4, 0.1666666666666666574, 0.1, n)
            assert abs(got - want[d]) <= 1e-10 * abs(want[d]), (step, d, got, want[d])
    lat.close()
