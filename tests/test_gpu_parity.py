"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C-ABI
(libchimp_b200.so via ctypes).  Bars: populations, rho and vel BIT-EXACT against the golden
dumps of the unmodified reference and against the oracle port on larger seeded cases (the
kernels are built with -fmad=false and keep the reference's operation order); the cases
whose global sums are reduced in a different order on the GPU state their tolerance."""
import ctypes as C

import numpy as np
import pytest

import helpers
from test_builder_host import build_engine_tables

pytestmark = pytest.mark.gpu

STD = [n for n in helpers.all_golden_names() if n.startswith(("std_", "trt_"))]


class InProcessRanks:
    """N ranks as N engine contexts on one GPU; the halo transport between the two halves of an
    iteration is a device-to-device copy of the packed buffers (stand-in for NCCL send/recv)."""

    def __init__(self, lats):
        self.lats = lats
        self.rt = C.CDLL("libcudart.so")
        self.rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]

    def step(self, n, **kw):
        for _ in range(n):
            for lat in self.lats:
                lat.step_begin(**kw)
            for lat in self.lats:
                lat.synchronize()
            for r, lat in enumerate(self.lats):
                for k in range(lat.num_neighbors()):
                    nr, ns, nrecv = lat.neighbor_info(k)
                    other = self.lats[nr]
                    ko = [j for j in range(other.num_neighbors()) if other.neighbor_info(j)[0] == r][0]
                    assert other.neighbor_info(ko)[1] == nrecv
                    if nrecv:
                        assert self.rt.cudaMemcpy(lat.recv_buffer_ptr(k), other.send_buffer_ptr(ko), nrecv * 8, 3) == 0
            # device-to-device copies on the legacy stream do not order against the engine's non-blocking streams
            assert self.rt.cudaDeviceSynchronize() == 0
            for lat in self.lats:
                lat.step_end()
            for lat in self.lats:
                lat.synchronize()


@pytest.mark.parametrize("index_form", [0, 1])
@pytest.mark.parametrize("name", [n for n in STD if helpers.Golden(n).nranks == 1])
def test_std_case_single_rank_bit_exact_vs_reference(name, index_form):
    g = helpers.Golden(name)
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    lats = build_engine_tables(g, lg, tabs, False)
    lat, t = lats[0], tabs[0]
    lat.finalize(index_form)
    lat.upload(pkg.cases.std_case_initial_state(t, g.attr("init_rho"))[0])
    a = g.args
    trt = tuple(a["trt"]) if "trt" in a else None
    done = 0
    bulk = t.bulk_nodes()
    for step in [s for s in g.dump if s > 0]:
        lat.step_single(step - done, tau=a.get("tau", 0.8), force=g.force(), trt=trt)
        done = step
        assert np.array_equal(lat.download()[bulk], g.f(0, step)[bulk]), "f differs at step %d" % step
        assert np.array_equal(lat.download_rho()[bulk, 0], g.rec(0, "step%d.rho" % step)[bulk])
        assert np.array_equal(lat.download_vel()[bulk], g.rec(0, "step%d.vel" % step).reshape(t.size, -1)[bulk])


@pytest.mark.parametrize("index_form", [0, 1])
@pytest.mark.parametrize("lattice,shape,periodic", [("D3Q19", (40, 36, 44), "xyz"), ("D3Q27", (30, 28, 26), "xyz"),
                                                    ("D2Q9", (96, 80), "xy"), ("D3Q19", (32, 30, 28), "")])
def test_std_case_bit_exact_vs_oracle_port(lattice, shape, periodic, index_form):
    """larger seeded sphere packs, 25 steps, BGK and TRT, against the oracle port"""
    pkg = helpers.load_package()
    port = helpers.oracle_port()
    geo = pkg.geometry.sphere_pack(shape, min(shape) / 7.0, 0.55, 5).astype(int)
    if periodic == "":
        for ax in range(len(shape)):
            for end in (0, -1):
                s = [slice(None)] * len(shape)
                s[ax] = end
                geo[tuple(s)] = 0
    lg = pkg.geometry.LatticeGeometry(geo, lattice, periodic)
    t = lg.all_ranks()[0]
    rng = np.random.default_rng(2)
    f0, _ = pkg.cases.std_case_initial_state(t, 1.0 + 0.05 * rng.random(geo.shape))
    bb = t.halfway_bb(t.fluid_bnd_nodes())
    bulk = t.bulk_nodes()
    force = (1e-6, -2e-6, 3e-6)[: lg.nd]
    for trt in (None, (0.8, 1.125)):
        lat = pkg.capi.Lattice.from_rank_tables(t)
        lat.add_halfway_bb(*bb)
        lat.finalize(index_form)
        lat.upload(f0)
        lat.step_single(25, tau=0.7, force=force, trt=trt)
        ref = port.PortRank(pkg.geometry.LATTICE_ID[lattice], t.neigh, bulk, 1, bb)
        ref.f[:] = f0
        ref.step_std_case(25, tau=0.7, force=force, trt=trt)
        assert np.array_equal(lat.download()[bulk], ref.f[bulk])
        assert np.array_equal(lat.download_rho()[bulk, 0], ref.rho[bulk, 0])
        assert np.array_equal(lat.download_vel()[bulk], ref.vel[bulk])
        if index_form == 1:
            assert 0.0 <= lat.irregular_fraction() < 0.9
        lat.close()


@pytest.mark.parametrize("boundary_first", [False, True])
@pytest.mark.parametrize("name", [n for n in STD if helpers.Golden(n).nranks > 1])
def test_std_case_n_rank_bit_exact_vs_reference(name, boundary_first):
    """every rank's populations match the reference's N-rank (MPI) run bit for bit"""
    g = helpers.Golden(name)
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    lats = build_engine_tables(g, lg, tabs, boundary_first)
    for lat, t in zip(lats, tabs):
        lat.finalize(1, boundary_first)
        lat.upload(pkg.cases.std_case_initial_state(t, g.attr("init_rho"))[0])
    ranks = InProcessRanks(lats)
    a = g.args
    done = 0
    for step in [s for s in g.dump if s > 0]:
        ranks.step(step - done, tau=a.get("tau", 0.8), force=g.force())
        done = step
        for r, (lat, t) in enumerate(zip(lats, tabs)):
            bulk = t.bulk_nodes()
            assert np.array_equal(lat.download()[bulk], g.f(r, step)[bulk]), "rank %d step %d" % (r, step)
            assert np.array_equal(lat.download_rho()[bulk, 0], g.rec(r, "step%d.rho" % step)[bulk])


@pytest.mark.parametrize("index_form", [0, 1])
def test_one_phase_vs_reference(index_form):
    """std_one_phase loop: masked force, mass-conservation source, solid / anti-bounce-back pressure /
    fluid-fluid links.  The per-label mass change is a tree sum on the GPU (sequential on the CPU),
    so the source differs by rounding order: tolerance 1e-12 relative on f, as north_star states;
    in practice the populations come out bit-identical or within one ulp."""
    g = helpers.Golden("onephase_d3q19_p1")
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    setup = helpers.one_phase_setup(g, lg, tabs)[0]
    lat = build_engine_tables(g, lg, tabs, False)[0]
    t = tabs[0]
    lat.finalize(index_form)
    lat.set_one_phase_attributes(setup["force_on"], setup["interior"], setup["add_source"], setup["scale"],
                                 g.args.get("rhow", 1.0))
    lat.upload(setup["f0"])
    bulk = t.bulk_nodes()
    done = 0
    for step in [s for s in g.dump if s > 0]:
        lat.step_single(step - done, tau=g.args["tau"], force=g.force())
        done = step
        got, ref = lat.download()[bulk], g.f(0, step)[bulk]
        assert np.allclose(got, ref, rtol=1e-12, atol=0.0)
        assert np.allclose(lat.download_rho()[bulk, 0], g.rec(0, "step%d.rho" % step)[bulk], rtol=1e-12, atol=0)
        assert np.allclose(lat.download_vel()[bulk], g.rec(0, "step%d.vel" % step).reshape(t.size, -1)[bulk],
                           rtol=1e-9, atol=1e-18)
        mass = lat.download_mass_change(len(setup["scale"]))
        assert np.allclose(mass, g.rec(0, "step%d.massChange" % step), rtol=1e-9, atol=1e-16)


@pytest.mark.parametrize("name,trt", [("onephase_d3q19_p1", False), ("onephase_d3q19_p1", True)])
def test_one_phase_packed_attribute_word_gives_the_bits_of_the_four_arrays(name, trt, monkeypatch):
    """The one_phase step kernel reads force switch, source switch, interior label and pressure-link mask either from
    four arrays (24 B/node, CHIMP_ATTR_PACKED=0) or from one packed word (default when the switches are 0 / 1, as in
    the reference's input).  Same products, same bits: populations, rho, u and the mass change agree exactly, with
    solid / pressure / fluid-fluid links and interior-domain sources active."""
    g = helpers.Golden(name)
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    setup = helpers.one_phase_setup(g, lg, tabs)[0]
    bulk = tabs[0].bulk_nodes()

    def run(packed):
        monkeypatch.setenv("CHIMP_ATTR_PACKED", "1" if packed else "0")    # read when the lattice is created
        lat = build_engine_tables(g, lg, tabs, False)[0]
        lat.finalize(1)
        lat.set_one_phase_attributes(setup["force_on"], setup["interior"], setup["add_source"], setup["scale"], g.args.get("rhow", 1.0))
        lat.upload(setup["f0"])
        kw = dict(tau=g.args["tau"], force=g.force())
        if trt:
            kw["trt"] = (g.args["tau"], 1.125)
        lat.step_single(12, **kw)
        out = (lat.download()[bulk], lat.download_rho()[bulk], lat.download_vel()[bulk], lat.download_mass_change(len(setup["scale"])))
        size = lat.one_phase_attribute_bytes_per_node()
        lat.close()
        return out, size

    arrays, arrays_bytes = run(False)
    packed, packed_bytes = run(True)
    assert (arrays_bytes, packed_bytes) == (24.0, 4.0)
    for a, b in zip(arrays, packed):
        assert np.array_equal(a, b)


def test_one_phase_without_interior_domains_is_bit_exact():
    """with no interior domains the mass source vanishes and no reduction enters: bit-exact vs the oracle"""
    pkg = helpers.load_package()
    port = helpers.oracle_port()
    g = helpers.Golden("onephase_d3q19_p1")
    lg, tabs = helpers.build_tables(g)
    attrs = {k: g.attr(k) for k in ("nodetags", "force", "interior_domains")}
    attrs["interior_domains"] = np.zeros_like(attrs["interior_domains"])
    setup = pkg.cases.one_phase_setup(lg, tabs, attrs)[0]
    t = tabs[0]
    for trt in (None, (0.8, 1.125)):
        lat = pkg.capi.Lattice.from_rank_tables(t)
        lat.add_links(pkg.capi.LINK_SOLID, setup["solid_links"])
        lat.add_links(pkg.capi.LINK_PRESSURE, setup["press_links"])
        lat.add_links(pkg.capi.LINK_FLUID_SWAP, setup["fluid_links"])
        lat.finalize(1)
        lat.set_one_phase_attributes(setup["force_on"], setup["interior"], setup["add_source"], setup["scale"], 1.0)
        lat.upload(setup["f0"])
        lat.step_single(12, tau=0.8, force=(0, 0, 1e-5), trt=trt)
        ref = port.PortRank(1, t.neigh, t.bulk_nodes(), 1)
        ref.set_one_phase(setup["force_on"], setup["interior"], setup["add_source"], setup["scale"],
                          setup["solid_links"], setup["press_links"], setup["fluid_links"], 1.0)
        ref.f[:] = setup["f0"]
        ref.step_one_phase(12, tau=0.8, force=(0, 0, 1e-5), trt=trt)
        bulk = t.bulk_nodes()
        assert np.array_equal(lat.download()[bulk], ref.f[bulk])
        lat.close()


def test_round_trip_upload_download():
    pkg = helpers.load_package()
    g = helpers.Golden("std_d3q19_p1")
    lg, tabs = helpers.build_tables(g)
    lat = build_engine_tables(g, lg, tabs, False)[0]
    lat.finalize(1)
    rng = np.random.default_rng(0)
    f = rng.random((tabs[0].size, 1, 19))
    lat.upload(f)
    out = lat.download()
    bulk = tabs[0].bulk_nodes()
    assert np.array_equal(out[bulk], f[bulk])


@pytest.mark.parametrize("index_form", [0, 1])
@pytest.mark.parametrize("name", ["twophase_d3q19_p1", "twophase_d2q9_p1"])
def test_twophase_vs_reference(name, index_form):
    """colour-gradient two-phase loop (2 LbFields).  The flux-control force comes from a global
    x-momentum sum that is a tree sum on the GPU (sequential on the CPU): F_x agrees to ~1e-13
    relative and enters f scaled by ~1e-6, so populations are compared at 1e-12 relative
    (north_star's per-step bar); rho0, rho1, phi of the step are computed before the force and
    are bit-exact on the first step."""
    g = helpers.Golden(name)
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    t = tabs[0]
    setup = pkg.cases.two_phase_setup(lg, tabs, g.attr("rho0"), g.attr("rho1"), g.attr("wettability"))[0]
    lat = pkg.capi.Lattice.from_rank_tables(t, n_fields=2)
    lat.add_halfway_bb(*t.halfway_bb(t.bulk_nodes()))
    lat.set_solid_boundary(setup["solid_bnd"])
    lat.finalize(index_form)
    lat.set_twophase_density(setup["rho"])
    lat.upload(setup["f0"])
    a = g.args
    bulk = t.bulk_nodes()
    done = 0
    for step in [s for s in g.dump if s > 0]:
        lat.step_twophase(step - done, a["tau2"][0], a["tau2"][1], a["sigma"], a["beta"], a["momx"], g.force(), len(bulk))
        done = step
        ref_f = g.f(0, step, 2)[bulk]
        got = lat.download()[bulk]
        assert np.allclose(got, ref_f, rtol=1e-12, atol=1e-300)
        ref_rho = g.rec(0, "step%d.rho" % step).reshape(-1, 2)[bulk]
        got_rho = lat.download_rho()[bulk]
        assert np.allclose(got_rho, ref_rho, rtol=1e-12, atol=0)
        ref_cg = g.rec(0, "step%d.cg" % step)
        got_cg = lat.download_phase_field()
        assert np.allclose(got_cg[bulk], ref_cg[bulk], rtol=1e-10, atol=1e-14)
        sb = setup["solid_bnd"]
        assert np.array_equal(got_cg[sb], ref_cg[sb])
        assert np.allclose(lat.download_vel()[bulk], g.rec(0, "step%d.vel" % step).reshape(t.size, -1)[bulk],
                           rtol=1e-9, atol=1e-16)
        fx = float(g.rec(0, "step%d.forceX" % step)[0])
        assert abs(lat.last_flux_force() - fx) <= 1e-10 * abs(fx)
        if step == 1:
            assert np.array_equal(got_rho, ref_rho)


class ThreadedRanks:
    """N ranks as N engine contexts on one GPU, each driven by its own host thread, so that the
    production code path (chimp_step_* with transport callbacks) runs unchanged; the callbacks
    rendezvous on a barrier and move data with device copies -- a stand-in for NCCL / MPI."""

    def __init__(self, lats):
        import threading
        self.lats = lats
        self.n = len(lats)
        self.barrier = threading.Barrier(self.n)
        self.rt = C.CDLL("libcudart.so")
        self.rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        self.rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
        self.shared = [None] * self.n
        for r, lat in enumerate(lats):
            lat.set_exchange_callback(self._exchange(r, False))
            lat.set_scalar_exchange_callback(self._exchange(r, True))
            lat.set_allreduce_callback(self._allreduce(r))

    def _peer_index(self, other, r):
        return [j for j in range(other.num_neighbors()) if other.neighbor_info(j)[0] == r][0]

    def _exchange(self, r, scalar):
        lat = self.lats[r]

        def cb(stream):
            assert self.rt.cudaStreamSynchronize(stream) == 0
            self.barrier.wait()
            for k in range(lat.num_neighbors()):
                nr = lat.neighbor_info(k)[0]
                other = self.lats[nr]
                ko = self._peer_index(other, r)
                if scalar:
                    cnt = lat.scalar_neighbor_info(k)[1]
                    src, dst = other.scalar_send_buffer_ptr(ko), lat.scalar_recv_buffer_ptr(k)
                else:
                    cnt = lat.neighbor_info(k)[2]
                    src, dst = other.send_buffer_ptr(ko), lat.recv_buffer_ptr(k)
                if cnt:
                    assert self.rt.cudaMemcpy(dst, src, cnt * 8, 3) == 0
            # the copies ran on the legacy stream, the unpack kernels follow on the engine's non-blocking stream: the
            # data must have landed (and the peers' send buffers been read) before anybody goes on
            assert self.rt.cudaDeviceSynchronize() == 0
            self.barrier.wait()
        return cb

    def _allreduce(self, r):
        def cb(dev, count, stream):
            assert self.rt.cudaStreamSynchronize(stream) == 0
            mine = np.zeros(count)
            assert self.rt.cudaMemcpy(mine.ctypes.data_as(C.c_void_p), dev, count * 8, 2) == 0
            self.shared[r] = mine
            self.barrier.wait()
            total = self.shared[0].copy()
            for k in range(1, self.n):      # rank order, like the oracle's MPI shim
                total = total + self.shared[k]
            self.barrier.wait()
            assert self.rt.cudaMemcpy(dev, total.ctypes.data_as(C.c_void_p), count * 8, 1) == 0
            assert self.rt.cudaDeviceSynchronize() == 0
        return cb

    def run(self, fn):
        import threading
        errors = []

        def work(r):
            try:
                fn(r, self.lats[r])
            except Exception as exc:  # pragma: no cover
                errors.append((r, exc))
                self.barrier.abort()
        threads = [threading.Thread(target=work, args=(r,)) for r in range(self.n)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors


def test_twophase_two_ranks_vs_reference():
    """2-rank colour-gradient run: scalar ghost exchange of phi between the passes, all-reduced flux
    force, ghost exchange of both LbFields.  Tolerance as in the single-rank test (tree sums)."""
    g = helpers.Golden("twophase_d3q19_p2")
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    setup = pkg.cases.two_phase_setup(lg, tabs, g.attr("rho0"), g.attr("rho1"), g.attr("wettability"))
    lats = []
    for r, t in enumerate(tabs):
        lat = pkg.capi.Lattice.from_rank_tables(t, n_fields=2)
        ss = t.send_side(tabs)
        for k, nr in enumerate(t.neig_ranks):
            lat.add_neighbor(nr, ss[k][0], ss[k][1], ss[k][2], t.recv_nodes[k], t.recv_ndir[k], t.recv_dirs[k])
        lat.add_halfway_bb(*t.halfway_bb(t.bulk_nodes()))
        lat.set_solid_boundary(setup[r]["solid_bnd"])
        lat.finalize(1, True)
        lat.set_twophase_density(setup[r]["rho"])
        lat.upload(setup[r]["f0"])
        lats.append(lat)
    ranks = ThreadedRanks(lats)
    a = g.args
    n_global = sum(len(t.bulk_nodes()) for t in tabs)
    done = 0
    for step in [s for s in g.dump if s > 0]:
        ranks.run(lambda r, lat: lat.step_twophase(step - done, a["tau2"][0], a["tau2"][1], a["sigma"], a["beta"], a["momx"],
                                                   g.force(), n_global))
        done = step
        for r, (lat, t) in enumerate(zip(lats, tabs)):
            bulk = t.bulk_nodes()
            assert np.allclose(lat.download()[bulk], g.f(r, step, 2)[bulk], rtol=1e-12, atol=1e-300), "rank %d step %d" % (r, step)
            assert np.allclose(lat.download_rho()[bulk], g.rec(r, "step%d.rho" % step).reshape(-1, 2)[bulk], rtol=1e-12, atol=0)
            assert np.allclose(lat.download_phase_field()[bulk], g.rec(r, "step%d.cg" % step)[bulk], rtol=1e-10, atol=1e-14)
            fx = float(g.rec(r, "step%d.forceX" % step)[0])
            assert abs(lat.last_flux_force() - fx) <= 1e-10 * abs(fx)


def test_one_phase_trt_two_ranks_vs_reference():
    """2-rank std_one_phase loop with TRT, link boundaries and the all-reduced mass-conservation source"""
    g = helpers.Golden("onephase_trt_d3q19_p2")
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    setup = helpers.one_phase_setup(g, lg, tabs)
    lats = build_engine_tables(g, lg, tabs, True)
    for r, lat in enumerate(lats):
        lat.finalize(1, True)
        s = setup[r]
        lat.set_one_phase_attributes(s["force_on"], s["interior"], s["add_source"], s["scale"], g.args.get("rhow", 1.0))
        lat.upload(s["f0"])
    ranks = ThreadedRanks(lats)
    trt = tuple(g.args["trt"])
    done = 0
    for step in [s for s in g.dump if s > 0]:
        ranks.run(lambda r, lat: lat.step_single(step - done, force=g.force(), trt=trt))
        done = step
        for r, (lat, t) in enumerate(zip(lats, tabs)):
            bulk = t.bulk_nodes()
            assert np.allclose(lat.download()[bulk], g.f(r, step)[bulk], rtol=1e-12, atol=0), "rank %d step %d" % (r, step)
            assert np.allclose(lat.download_rho()[bulk, 0], g.rec(r, "step%d.rho" % step)[bulk], rtol=1e-12, atol=0)
            mass = lat.download_mass_change(len(setup[r]["scale"]))
            assert np.allclose(mass, g.rec(r, "step%d.massChange" % step), rtol=1e-9, atol=1e-16)


@pytest.mark.parametrize("shape,lattice,periodic", [((1, 1, 1), "D3Q19", "xyz"), ((3, 1, 2), "D3Q27", "xyz"), ((5, 3), "D2Q9", "xy"),
                                                    ((33, 1, 1), "D3Q19", "xyz")])
def test_tiny_and_ragged_lattices(shape, lattice, periodic):
    """edge cases: a single periodic node (every neighbour is the node itself), node counts that are
    not a multiple of the 32-node tile, one-node-thick domains"""
    pkg = helpers.load_package()
    port = helpers.oracle_port()
    geo = np.ones(shape, dtype=int)
    lg = pkg.geometry.LatticeGeometry(geo, lattice, periodic)
    t = lg.all_ranks()[0]
    rng = np.random.default_rng(1)
    f0, _ = pkg.cases.std_case_initial_state(t, 1.0 + 0.1 * rng.random(shape))
    bb = t.halfway_bb(t.fluid_bnd_nodes())
    bulk = t.bulk_nodes()
    for form in (0, 1):
        lat = pkg.capi.Lattice.from_rank_tables(t)
        lat.add_halfway_bb(*bb)
        lat.finalize(form)
        lat.upload(f0)
        lat.step_single(7, tau=0.9, force=(1e-5, 2e-5, -1e-5)[: lg.nd])
        ref = port.PortRank(pkg.geometry.LATTICE_ID[lattice], t.neigh, bulk, 1, bb)
        ref.f[:] = f0
        ref.step_std_case(7, tau=0.9, force=(1e-5, 2e-5, -1e-5)[: lg.nd])
        assert np.array_equal(lat.download()[bulk], ref.f[bulk])
        lat.close()


def test_lattice_without_fluid_nodes_is_rejected():
    pkg = helpers.load_package()
    geo = np.zeros((4, 4, 4), dtype=int)
    geo[0, 0, 0] = 1
    t = pkg.geometry.LatticeGeometry(geo, "D3Q19", "xyz").all_ranks()[0]
    with pytest.raises(pkg.capi.ChimpError):
        pkg.capi.Lattice("D3Q19", t.neigh, np.zeros(0, dtype=np.int32)).finalize()


def test_step_argument_errors():
    pkg = helpers.load_package()
    g = helpers.Golden("std_d2q9_channel")
    lg, tabs = helpers.build_tables(g)
    lat = build_engine_tables(g, lg, tabs, False)[0]
    with pytest.raises(pkg.capi.ChimpError):
        lat.step_single(1)                       # not finalized
    lat.finalize(1)
    with pytest.raises(pkg.capi.ChimpError):
        lat.step_twophase(1, 1.0, 1.0, 0.01, 1.0, 1e-5, (0, 0, 0), 10)   # one-field lattice
    with pytest.raises(pkg.capi.ChimpError):
        lat.finalize(1)                          # twice


def _connect_in_process(lats):
    for r, lat in enumerate(lats):
        for k in range(lat.num_neighbors()):
            nr = lat.neighbor_info(k)[0]
            other = lats[nr]
            ko = [j for j in range(other.num_neighbors()) if other.neighbor_info(j)[0] == r][0]
            lat.connect_peer(k, other.nq * other.plane_stride(), ko, other.recv_dst(ko), pointers=other.local_pointers())


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", ["std_d3q19_p3", "std_d3q27_p2", "std_d2q9_pack_p2"])
def test_peer_halos_bit_exact_vs_reference(name, fused, monkeypatch):
    """peer path: every rank stores its outgoing populations straight into the neighbours' halo-in slots
    and publishes an arrival counter (no pack / transport / unpack) -- from inside the step kernel, whose first
    blocks hold the halo-coupled nodes (fused, the default), or from separate push launches (CHIMP_PEER_FUSED=0).
    N contexts in one process connect through raw device pointers, N processes through CUDA IPC handles
    (tests/multi_gpu_check.py)"""
    monkeypatch.setenv("CHIMP_PEER_FUSED", "1" if fused else "0")
    g = helpers.Golden(name)
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    lats = build_engine_tables(g, lg, tabs, True)
    for lat, t in zip(lats, tabs):
        lat.finalize(1, True)
        lat.upload(pkg.cases.std_case_initial_state(t, g.attr("init_rho"))[0])
    _connect_in_process(lats)
    for lat in lats:
        assert lat.peer_mode() == ((2, "") if fused else (1, "switched off (CHIMP_PEER_FUSED=0)")), lat.peer_mode()
    lib = pkg.capi.lib()
    a = g.args
    done = 0
    for step in [s for s in g.dump if s > 0]:
        l0 = lib.chimp_launch_count()
        for _ in range(step - done):
            for lat in lats:
                lat.step_begin(tau=a.get("tau", 0.8), force=g.force())
            for lat in lats:
                lat.step_end()
        if fused:   # one launch per rank and step: exchange and counters live inside the step kernel
            assert lib.chimp_launch_count() - l0 == (step - done) * len(lats)
        done = step
        for r, (lat, t) in enumerate(zip(lats, tabs)):
            bulk = t.bulk_nodes()
            assert np.array_equal(lat.download()[bulk], g.f(r, step)[bulk]), "rank %d step %d" % (r, step)


def test_peer_that_never_arrives_raises_an_error_instead_of_hanging(monkeypatch):
    """bounded device-side waits: a rank whose neighbour stops stepping gets an error from the next synchronising
    call (CHIMP_PEER_TIMEOUT_MS), not a hung GPU"""
    monkeypatch.setenv("CHIMP_PEER_TIMEOUT_MS", "300")
    g = helpers.Golden("std_d3q19_p3")
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    lats = build_engine_tables(g, lg, tabs, True)
    for lat, t in zip(lats, tabs):
        lat.finalize(1, True)
        lat.upload(pkg.cases.std_case_initial_state(t, g.attr("init_rho"))[0])
    _connect_in_process(lats)
    lats[0].step_single(1, tau=0.8, force=g.force())     # step 1 needs nothing from the neighbours
    lats[0].synchronize()
    lats[0].step_single(1, tau=0.8, force=g.force())     # step 2 waits for their step 1, which never comes
    with pytest.raises(pkg.capi.ChimpError, match="did not arrive"):
        lats[0].synchronize()


def test_mass_flux_through_pressure_nodes_and_flux_force():
    """the reductions a main performs either side of the path, formed on the device.
    Mass flux (std_one_phase/main.cpp:607-619): products on the GPU, added in list order on the host --
    bit-identical to the same loop over the downloaded fields, and equal to the reference's dump within the
    tolerance of the tree-summed mass source.  Flux force (LBglobalforcing.h:8-33): tree sum vs sequential."""
    g = helpers.Golden("onephase_d3q19_p1")
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    setup = helpers.one_phase_setup(g, lg, tabs)[0]
    lat = build_engine_tables(g, lg, tabs, False)[0]
    t = tabs[0]
    lat.finalize(1)
    lat.set_one_phase_attributes(setup["force_on"], setup["interior"], setup["add_source"], setup["scale"], g.args.get("rhow", 1.0))
    lat.upload(setup["f0"])
    step = int(max(g.dump))
    lat.step_single(step, tau=g.args["tau"], force=g.force())
    nodes, phase = setup["press_nodes"], setup["press_phase"]
    assert len(nodes) > 0 and set(np.unique(phase)) <= {0, 1}
    local, q, change = pkg.cases.one_phase_mass_flux(lat, setup)

    def sequential(rho, vel):
        acc = [0.0, 0.0]
        for n, p in zip(nodes, phase):
            acc[p] += vel[n, 2] * rho[n]
        return np.array(acc)

    assert np.array_equal(local, sequential(lat.download_rho()[:, 0], lat.download_vel()))
    ref = sequential(g.rec(0, "step%d.rho" % step), g.rec(0, "step%d.vel" % step).reshape(t.size, -1))
    assert np.allclose(local, ref, rtol=1e-9, atol=1e-18)
    assert np.array_equal(q, 0.5 * local)
    # nodes that are not own nodes contribute the zero-initialised rows of the reference fields
    assert np.array_equal(lat.node_list_flux([0, t.size - 1], [0, 1], 2), np.zeros(2))
    with pytest.raises(pkg.capi.ChimpError):
        lat.node_list_flux(nodes, phase + 5, 2)
    # flux force: 2 * (fixed - mean of sum_q c_qd f_q)
    f = lat.download()
    bulk = t.bulk_nodes()
    c = pkg.geometry.BASIS["D3Q19"]
    for d in range(3):
        mean = 0.0
        for n in bulk:
            mean += float(np.sum(f[n, 0, :] * c[:, d]))
        want = 2 * (1e-5 - mean / len(bulk))
        got = lat.flux_force(0, d, 1e-5, len(bulk))
        assert abs(got - want) <= 1e-12 * max(abs(want), 1e-300) + 1e-20


def test_twophase_two_ranks_over_peer_memory_vs_reference():
    """2-rank colour-gradient run with every exchange over peer memory: phi faces and population faces stored into
    the neighbour's arrays, momentum sum through the mailboxes (added in rank order), arrival counters instead of
    callbacks.  Both contexts live in this process and connect through raw device pointers; separate processes
    connect through CUDA IPC (tests/multi_gpu_check.py)."""
    g = helpers.Golden("twophase_d3q19_p2")
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    setup = pkg.cases.two_phase_setup(lg, tabs, g.attr("rho0"), g.attr("rho1"), g.attr("wettability"))
    lats = []
    for r, t in enumerate(tabs):
        lat = pkg.capi.Lattice.from_rank_tables(t, n_fields=2)
        ss = t.send_side(tabs)
        for k, nr in enumerate(t.neig_ranks):
            lat.add_neighbor(nr, ss[k][0], ss[k][1], ss[k][2], t.recv_nodes[k], t.recv_ndir[k], t.recv_dirs[k])
        lat.add_halfway_bb(*t.halfway_bb(t.bulk_nodes()))
        lat.set_solid_boundary(setup[r]["solid_bnd"])
        lat.finalize(1, True)
        lat.set_twophase_density(setup[r]["rho"])
        lat.upload(setup[r]["f0"])
        lats.append(lat)
    world = len(lats)
    for r, lat in enumerate(lats):
        for k in range(lat.num_neighbors()):
            nr = lat.neighbor_info(k)[0]
            other = lats[nr]
            ko = [j for j in range(other.num_neighbors()) if other.neighbor_info(j)[0] == r][0]
            lat.connect_peer(k, other.nq * other.plane_stride(), ko, other.recv_dst(ko), pointers=other.local_pointers())
            lat.connect_peer_scalar(k, other.host_scalar_recv_slots(ko), pointer=other.local_pointers_twophase()[0])
    for r, lat in enumerate(lats):
        lat.connect_world(r, world, pointers=[l.local_pointers_twophase()[1] for l in lats])
    a = g.args
    n_global = sum(len(t.bulk_nodes()) for t in tabs)
    done = 0
    for step in [s for s in g.dump if s > 0]:
        for _ in range(step - done):   # one step at a time so that the contexts advance together
            for lat in lats:
                lat.step_twophase(1, a["tau2"][0], a["tau2"][1], a["sigma"], a["beta"], a["momx"], g.force(), n_global)
        done = step
        for r, (lat, t) in enumerate(zip(lats, tabs)):
            bulk = t.bulk_nodes()
            assert np.allclose(lat.download()[bulk], g.f(r, step, 2)[bulk], rtol=1e-12, atol=1e-300), "rank %d step %d" % (r, step)
            assert np.allclose(lat.download_rho()[bulk], g.rec(r, "step%d.rho" % step).reshape(-1, 2)[bulk], rtol=1e-12, atol=0)
            assert np.allclose(lat.download_phase_field()[bulk], g.rec(r, "step%d.cg" % step)[bulk], rtol=1e-10, atol=1e-14)
            fx = float(g.rec(r, "step%d.forceX" % step)[0])
            assert abs(lat.last_flux_force() - fx) <= 1e-10 * abs(fx)
    assert lats[0].last_flux_force() == lats[1].last_flux_force()   # rank-order sum: identical bits on every rank


def test_capillary_number_force_and_error_paths_of_the_caller_side_reductions():
    """calcCapNumbForceCartDir (LBglobalforcing.h:35-98) as a device reduction against the same loop over the
    downloaded fields; argument errors of the 8(f2) entry points"""
    g = helpers.Golden("twophase_d3q19_p1")
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    t = tabs[0]
    setup = pkg.cases.two_phase_setup(lg, tabs, g.attr("rho0"), g.attr("rho1"), g.attr("wettability"))[0]
    lat = pkg.capi.Lattice.from_rank_tables(t, n_fields=2)
    lat.add_halfway_bb(*t.halfway_bb(t.bulk_nodes()))
    lat.set_solid_boundary(setup["solid_bnd"])
    lat.finalize(1)
    lat.set_twophase_density(setup["rho"])
    lat.upload(setup["f0"])
    a = g.args
    bulk = t.bulk_nodes()
    lat.step_twophase(5, a["tau2"][0], a["tau2"][1], a["sigma"], a["beta"], a["momx"], g.force(), len(bulk))
    f, rho = lat.download(), lat.download_rho()
    c = pkg.geometry.BASIS["D3Q19"]
    nu0, nu1, sig = 1.0 / 6.0, 0.1, 0.01 * 1e-3
    for d in range(3):
        s = np.zeros(4)
        for n in bulk:
            m = float(np.sum(f[n, 0, :] * c[:, d]))
            p0, p1 = rho[n, 0] / (rho[n, 0] + rho[n, 1]), rho[n, 1] / (rho[n, 0] + rho[n, 1])
            s += [p0 * m, p1 * m, p0, p1]
        s /= len(bulk)
        want = 2 * (sig - (s[0] * nu0 + s[1] * nu1)) / (s[2] * nu0 + s[3] * nu1)
        got = lat.capillary_force(d, sig, nu0, nu1, len(bulk))
        assert abs(got - want) <= 1e-10 * abs(want) + 1e-18
    for bad in (lambda: lat.capillary_force(3, sig, nu0, nu1, len(bulk)), lambda: lat.capillary_force(0, sig, nu0, nu1, 0),
                lambda: lat.flux_force(2, 0, 1e-5, len(bulk)), lambda: lat.flux_force(0, 5, 1e-5, len(bulk)),
                lambda: lat.node_list_flux([1, 2], [0, 1], 2, field_no=7), lambda: lat.node_list_flux([1, 2], [0, 1], 2, component=4),
                lambda: lat.connect_world(0, 2, pointers=[0, 0]), lambda: lat.add_scalar_halo_face(0, [0], [0])):
        with pytest.raises(pkg.capi.ChimpError):
            bad()
    one = pkg.capi.Lattice.from_rank_tables(t)
    one.add_halfway_bb(*t.halfway_bb(t.fluid_bnd_nodes()))
    one.finalize(1)
    with pytest.raises(pkg.capi.ChimpError):
        one.capillary_force(0, sig, nu0, nu1, len(bulk))     # one-field lattice
    with pytest.raises(pkg.capi.ChimpError):
        one.ipc_handles_twophase()
