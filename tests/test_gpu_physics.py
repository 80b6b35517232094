"""GPU tests of longer runs (pytest -m gpu): the 10 000-step bar of north_star, the D2Q9 SRT
Poiseuille channel against the analytic profile (BASELINE.json configs[1], validated at
H = 32 because 8192^2 cannot reach steady state), and size-independent
properties at a larger size (mass conservation, agreement of the two index forms)."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def test_ten_thousand_steps_match_the_oracle_bit_for_bit():
    """macroscopic fields must agree within 1e-10 after 10k steps; they are in fact identical"""
    pkg = helpers.load_package()
    port = helpers.oracle_port()
    geo = pkg.geometry.sphere_pack((14, 12, 16), 3.5, 0.6, 8).astype(int)
    lg = pkg.geometry.LatticeGeometry(geo, "D3Q19", "xyz")
    t = lg.all_ranks()[0]
    f0, _ = pkg.cases.std_case_initial_state(t, np.ones(geo.shape))
    bb = t.halfway_bb(t.fluid_bnd_nodes())
    bulk = t.bulk_nodes()
    lat = pkg.capi.Lattice.from_rank_tables(t)
    lat.add_halfway_bb(*bb)
    lat.finalize(pkg.capi.INDEX_COMPACT)
    lat.upload(f0)
    lat.step_single(10000, tau=0.8, force=(1e-6, 0, 0))
    ref = port.PortRank(1, t.neigh, bulk, 1, bb)
    ref.f[:] = f0
    ref.step_std_case(10000, tau=0.8, force=(1e-6, 0, 0))
    rho, vel = lat.download_rho()[bulk, 0], lat.download_vel()[bulk]
    assert np.allclose(rho, ref.rho[bulk, 0], rtol=1e-10, atol=0)
    assert np.allclose(vel, ref.vel[bulk], rtol=1e-10, atol=1e-20)
    assert np.array_equal(lat.download()[bulk], ref.f[bulk])
    assert np.array_equal(rho, ref.rho[bulk, 0]) and np.array_equal(vel, ref.vel[bulk])


def test_twophase_ten_thousand_steps_within_north_star_tolerance():
    """colour-gradient run of 10 000 steps against the oracle port (pinned bit for bit to the reference's dumps):
    rho0, rho1, u and phi within 1e-10 as north_star asks -- the only arithmetic difference is the order in which the
    flux controller's momentum sum is added up (tree on the GPU, sequential on the CPU)"""
    pkg = helpers.load_package()
    port = helpers.oracle_port()
    shape = (12, 10, 11)
    geo = pkg.geometry.sphere_pack(shape, 3.0, 0.62, 4).astype(int)
    x = np.arange(shape[0])[:, None, None] * np.ones(shape)
    rho0 = (x < shape[0] / 2).astype(float)
    lg = pkg.geometry.LatticeGeometry(geo, "D3Q19", "xyz")
    t = lg.all_ranks()[0]
    s = pkg.cases.two_phase_setup(lg, [t], rho0, 1.0 - rho0, 0.4 * (geo == 0))[0]
    bulk = t.bulk_nodes()
    args = dict(tau0=1.0, tau1=0.8, sigma=0.01, beta=1.0, momx=1e-5, force=(0.0, 1e-7, 0.0))
    lat = pkg.capi.Lattice.from_rank_tables(t, n_fields=2)
    lat.add_halfway_bb(*t.halfway_bb(bulk))
    lat.set_solid_boundary(s["solid_bnd"])
    lat.finalize(pkg.capi.INDEX_COMPACT)
    lat.set_twophase_density(s["rho"])
    lat.upload(s["f0"])
    lat.step_twophase(10000, args["tau0"], args["tau1"], args["sigma"], args["beta"], args["momx"], args["force"], len(bulk))
    ref = port.PortRank(1, t.neigh, bulk, 2, t.halfway_bb(bulk))
    ref.f[:] = s["f0"]
    ref.rho[:] = s["rho"]
    fx = ref.step_twophase(10000, s["solid_bnd"], args["tau0"], args["tau1"], args["sigma"], args["beta"], args["momx"],
                           list(args["force"]), len(bulk))
    assert np.allclose(lat.download_rho()[bulk], ref.rho[bulk], rtol=1e-10, atol=1e-13)
    assert np.allclose(lat.download_vel()[bulk], ref.vel[bulk], rtol=1e-10, atol=1e-13)
    assert np.allclose(lat.download_phase_field()[bulk], ref.cg.reshape(-1)[bulk], rtol=1e-10, atol=1e-13)
    assert np.allclose(lat.download()[bulk], ref.f[bulk], rtol=1e-10, atol=1e-14)
    assert abs(lat.last_flux_force() - fx) <= 1e-8 * abs(fx) + 1e-16
    assert np.isfinite(ref.f[bulk]).all() and ref.rho[bulk].min() > -1e-12


def test_poiseuille_channel_matches_analytic_profile():
    """D2Q9 SRT, periodic in x, walls at y = 0 and y = H+1, body force along x: the steady profile
    is u(y) = F/(2 nu) y' (H - y') with y' measured from the half-way wall; half-way bounce back
    leaves an O(1e-3) relative slip error at tau = 0.8 (the survey measured 5.1e-4 at H = 32)."""
    pkg = helpers.load_package()
    H, nx = 32, 4
    geo = np.ones((nx, H + 2), dtype=int)
    geo[:, 0] = geo[:, -1] = 0
    lg = pkg.geometry.LatticeGeometry(geo, "D2Q9", "x")
    t = lg.all_ranks()[0]
    lat = pkg.capi.Lattice.from_rank_tables(t)
    lat.add_halfway_bb(*t.halfway_bb(t.fluid_bnd_nodes()))
    lat.finalize(pkg.capi.INDEX_COMPACT)
    lat.upload(pkg.cases.std_case_initial_state(t, np.ones(geo.shape))[0])
    tau, F = 0.8, 1e-6
    lat.step_single(40000, tau=tau, force=(F, 0.0))
    vel = lat.download_vel()
    bulk = t.bulk_nodes()
    y = t.pos[bulk, 1].astype(float)          # fluid rows are y = 1..H, walls half-way at 0.5 and H + 0.5
    nu = (tau - 0.5) / 3.0
    yw = y - 0.5
    ua = F / (2.0 * nu) * yw * (H - yw)
    err = np.abs(vel[bulk, 0] - ua).max() / ua.max()
    assert err < 2e-3, err
    assert np.abs(vel[bulk, 1]).max() < 1e-12


@pytest.mark.parametrize("lattice", ["D3Q19", "D3Q27"])
def test_mass_conservation_and_index_form_agreement_at_larger_size(lattice):
    """structured ingest path at 96^3: both index forms give identical moments, and collide +
    stream + bounce back conserve the total mass to rounding"""
    import importlib
    import torch
    pkg = helpers.load_package()
    ingest = importlib.import_module("badchimp_cpp_b200.ingest")
    geo = pkg.geometry.sphere_pack((96, 96, 96), 12.0, 0.4, 77)
    res = []
    for form in (pkg.capi.INDEX_TABLE, pkg.capi.INDEX_COMPACT):
        table, labels, n, n_pad = ingest.build_pull_table(torch.from_numpy(geo).cuda().bool(), lattice, "xyz")
        lat = pkg.capi.lattice_from_device_table(lattice, n, n_pad, 0, table.data_ptr(), labels.data_ptr(), 1, form)
        lat.init_uniform(1.0)
        lat.step_single(200, tau=0.8, force=(1e-5, 0, 0))
        res.append(lat.download_moments_device_order())
        lat.close()
    (rho_a, vel_a), (rho_b, vel_b) = res
    assert np.array_equal(rho_a, rho_b) and np.array_equal(vel_a, vel_b)
    assert abs(rho_a.sum() / len(rho_a) - 1.0) < 1e-12
    assert vel_a[0].mean() > 0


def test_twophase_structured_ingest_equals_reference_table_path():
    """two-field lattice built by the device-side ingest (pull table + phi table + equilibrium init)
    gives bit-identical populations to the one built from the reference's tables"""
    import importlib
    import torch
    pkg = helpers.load_package()
    ingest = importlib.import_module("badchimp_cpp_b200.ingest")
    shape = (26, 22, 24)
    geo = pkg.geometry.sphere_pack(shape, 5.0, 0.5, 13).astype(int)
    x = np.arange(shape[0])[:, None, None] * np.ones(shape)
    rho0 = (x < shape[0] / 2).astype(float)
    wet = 0.25 * (geo == 0)
    lg = pkg.geometry.LatticeGeometry(geo, "D3Q19", "xyz")
    t = lg.all_ranks()[0]
    setup = pkg.cases.two_phase_setup(lg, [t], rho0, 1.0 - rho0, wet)[0]
    a = pkg.capi.Lattice.from_rank_tables(t, n_fields=2)
    a.add_halfway_bb(*t.halfway_bb(t.bulk_nodes()))
    a.set_solid_boundary(setup["solid_bnd"])
    a.finalize(pkg.capi.INDEX_COMPACT)
    a.set_twophase_density(setup["rho"])
    a.upload(setup["f0"])
    n = len(t.bulk_nodes())
    args = (1.0, 0.8, 0.01, 1.0, 1e-5, (0, 1e-7, 0), n)
    a.step_twophase(20, *args)
    fa = a.download()[t.bulk_nodes()]
    # device path
    fluid = torch.from_numpy(geo > 0).cuda()
    table, labels, nd_, n_pad = ingest.build_pull_table(fluid, "D3Q19", "xyz")
    assert nd_ == n
    b = pkg.capi.lattice_from_device_table("D3Q19", n, n_pad, 0, table.data_ptr(), labels.data_ptr(), 2, pkg.capi.INDEX_COMPACT)
    w32 = torch.from_numpy(wet.astype(np.float32))
    wall_phi = ((w32 - (1.0 - w32)) .double() / (w32.double() + (1.0 - w32).double()))
    # the reference computes (rho0 - rho1)/(rho0 + rho1) from float32-parsed wall densities (main_TWOPHASE.cpp:173-181, 280-284)
    r0 = w32.double()
    r1 = (torch.tensor(1.0, dtype=torch.float32) - w32).double()
    wall_phi = (r0 - r1) / (r0 + r1)
    ptable, n_extra, phi_extra = ingest.build_phi_table(fluid, wall_phi, "D3Q19", "xyz")
    b.set_phi_table_dev(ptable.data_ptr(), n_extra, phi_extra.data_ptr())
    rho_dev = torch.stack([torch.from_numpy(rho0)[torch.from_numpy(geo > 0)], torch.from_numpy(1.0 - rho0)[torch.from_numpy(geo > 0)]]).cuda().contiguous()
    b.init_equilibrium_dev(rho_dev.data_ptr())
    b.step_twophase(20, *args)
    b.n_nodes = n + 1
    fb = b.download()[1:]
    assert np.array_equal(fa, fb)


@pytest.mark.parametrize("lattice,shape,periodic", [("D3Q19", (44, 40, 36), "xyz"), ("D3Q19", (30, 34, 28), "x"), ("D2Q9", (300, 260), "xy")])
def test_twophase_step_does_not_depend_on_how_a_run_is_split_into_calls(lattice, shape, periodic):
    """two-phase stepping keeps its flux-controller state on the device between calls: 12 steps in one call and in
    calls of 1 + 4 + 7 steps give the same bits (populations, rho, phi, u, flux force); the int32-table index form
    gives the same fields as the compact one.  The closed box has no periodic far range, the periodic cases have one."""
    pkg = helpers.load_package()
    geo = pkg.geometry.sphere_pack(shape, 5.0, 0.55, 17).astype(int)
    if periodic != "xyz" and periodic != "xy":
        geo[:, 0] = geo[:, -1] = 0
        geo[:, :, 0] = geo[:, :, -1] = 0
    x = np.arange(shape[0]).reshape((-1,) + (1,) * (len(shape) - 1)) * np.ones(shape)
    rho0 = (x < shape[0] / 2).astype(float)
    lg = pkg.geometry.LatticeGeometry(geo, lattice, periodic)
    t = lg.all_ranks()[0]
    setup = pkg.cases.two_phase_setup(lg, [t], rho0, 1.0 - rho0, 0.3 * (geo == 0))[0]
    bulk = t.bulk_nodes()
    args = (1.0, 0.8, 0.01, 1.0, 1e-5, (0, 1e-7, 0), len(bulk))

    def run(chunks, index_form):
        lat = pkg.capi.Lattice.from_rank_tables(t, n_fields=2)
        lat.add_halfway_bb(*t.halfway_bb(bulk))
        lat.set_solid_boundary(setup["solid_bnd"])
        lat.finalize(index_form)
        lat.set_twophase_density(setup["rho"])
        lat.upload(setup["f0"])
        for k in chunks:
            lat.step_twophase(k, *args)
        out = (lat.download()[bulk], lat.download_rho()[bulk], lat.download_phase_field()[bulk], lat.download_vel()[bulk], lat.last_flux_force())
        lat.close()
        return out

    one = run([12], 1)
    split = run([1, 4, 7], 1)
    table = run([12], 0)
    for a, b in zip(one[:4], split[:4]):
        assert np.array_equal(a, b)
    assert one[4] == split[4]
    for a, b in zip(one[:4], table[:4]):
        assert np.allclose(a, b, rtol=1e-13, atol=1e-16)
    assert abs(one[4] - table[4]) <= 1e-12 * abs(one[4])


@pytest.mark.parametrize("lattice,shape,periodic,index_form", [("D3Q19", (36, 30, 32), "xyz", 1), ("D3Q19", (26, 30, 24), "x", 0), ("D2Q9", (200, 160), "xy", 1)])
def test_twophase_derived_phi_index_gives_the_bits_of_the_phi_table(lattice, shape, periodic, index_form, monkeypatch):
    """The collide pass locates phi of neighbor(q, n) (colour gradient, LButilities.h:12-22) either through the full
    table (CHIMP_PHI_DERIVED=0, 4 nQ bytes per node) or through its derived form (default: the pull source of rev(q),
    plus the stored links at solid / ghost / zero slots).  Both must select the same slots: every field is bit-identical
    after 15 steps, and the derived form is the one in use and an order of magnitude smaller."""
    pkg = helpers.load_package()
    geo = pkg.geometry.sphere_pack(shape, 4.0, 0.5, 23).astype(int)
    if periodic == "x":
        geo[:, 0] = geo[:, -1] = 0
        geo[:, :, 0] = geo[:, :, -1] = 0
    x = np.arange(shape[0]).reshape((-1,) + (1,) * (len(shape) - 1)) * np.ones(shape)
    rho0 = (x < shape[0] / 2).astype(float)
    lg = pkg.geometry.LatticeGeometry(geo, lattice, periodic)
    t = lg.all_ranks()[0]
    setup = pkg.cases.two_phase_setup(lg, [t], rho0, 1.0 - rho0, 0.3 * (geo == 0))[0]
    bulk = t.bulk_nodes()
    args = (1.0, 0.8, 0.01, 1.0, 1e-5, (0, 1e-7, 0), len(bulk))
    nq = {"D3Q19": 19, "D2Q9": 9}[lattice]

    def run(derived):
        monkeypatch.setenv("CHIMP_PHI_DERIVED", "1" if derived else "0")   # read when the lattice is created
        lat = pkg.capi.Lattice.from_rank_tables(t, n_fields=2)
        lat.add_halfway_bb(*t.halfway_bb(bulk))
        lat.set_solid_boundary(setup["solid_bnd"])
        lat.finalize(index_form)
        lat.set_twophase_density(setup["rho"])
        lat.upload(setup["f0"])
        lat.step_twophase(15, *args)
        out = (lat.download()[bulk], lat.download_rho()[bulk], lat.download_phase_field()[bulk], lat.download_vel()[bulk], lat.last_flux_force())
        size = lat.phi_index_bytes_per_node()
        lat.close()
        return out, size

    table, table_bytes = run(False)
    derived, derived_bytes = run(True)
    assert table_bytes == 4.0 * nq
    assert 4.0 <= derived_bytes < 0.5 * table_bytes, derived_bytes
    for a, b in zip(table[:4], derived[:4]):
        assert np.array_equal(a, b)
    assert table[4] == derived[4]
