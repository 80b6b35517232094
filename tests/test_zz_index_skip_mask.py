"""Skip-mask form of the compact index (chimp_set_index_skip_mask, kernels.cuh IDX_COMPACT_MASK): the plain single-field
step kernel reads a tile's bases first and leaves out every delta word the tile's mask marks as a plain run.  Same
tables, same sources: the populations must be the bits of the goldens / the oracle port, whether the form is on from the
start or switched in the middle of a run.  (Last file of the suite on purpose: the form is off by default.)"""
import numpy as np
import pytest

import helpers
from test_builder_host import build_engine_tables

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["std_d3q19_p1", "trt_d3q19_p1", "std_d2q9_channel", "std_d3q27_p2", "std_d3q19_p3"])
def test_skip_mask_form_bit_exact_vs_reference(name):
    from test_gpu_parity import InProcessRanks
    g = helpers.Golden(name)
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    lats = build_engine_tables(g, lg, tabs, g.nranks > 1)
    for lat, t in zip(lats, tabs):
        lat.finalize(1, g.nranks > 1)
        before = lat.index_bytes_per_node()
        lat.set_index_skip_mask(True)
        assert lat.index_bytes_per_node() <= before
        lat.upload(pkg.cases.std_case_initial_state(t, g.attr("init_rho"))[0])
    ranks = InProcessRanks(lats) if g.nranks > 1 else None
    a = g.args
    trt = tuple(a["trt"]) if "trt" in a else None
    done = 0
    for step in [s for s in g.dump if s > 0]:
        if ranks:
            ranks.step(step - done, tau=a.get("tau", 0.8), force=g.force())
        else:
            lats[0].step_single(step - done, tau=a.get("tau", 0.8), force=g.force(), trt=trt)
        done = step
        for r, (lat, t) in enumerate(zip(lats, tabs)):
            bulk = t.bulk_nodes()
            assert np.array_equal(lat.download()[bulk], g.f(r, step)[bulk]), "rank %d step %d" % (r, step)
            assert np.array_equal(lat.download_rho()[bulk, 0], g.rec(r, "step%d.rho" % step)[bulk])
    for lat in lats:
        lat.close()


@pytest.mark.parametrize("lattice,shape,periodic", [("D3Q19", (40, 36, 44), "xyz"), ("D2Q9", (300, 280), "x"), ("D3Q27", (30, 26, 34), "xz")])
def test_skip_mask_form_on_larger_lattices_and_switched_mid_run(lattice, shape, periodic):
    """larger seeded cases against the oracle port: many tiles whose delta words are plain runs (the share is reported and
    must be substantial in the open channel), the form switched on after a few steps, off again, moments included"""
    pkg = helpers.load_package()
    port = helpers.oracle_port()
    geo = pkg.geometry.sphere_pack(shape, 6.0, 0.6, 5).astype(int) if lattice != "D2Q9" else np.ones(shape, dtype=int)
    if lattice == "D2Q9":
        geo[:, 0] = geo[:, -1] = 0
    for ax, name in enumerate("xyz"[:len(shape)]):
        if name not in periodic:
            sl = [slice(None)] * len(shape)
            sl[ax] = 0
            geo[tuple(sl)] = 0
            sl[ax] = -1
            geo[tuple(sl)] = 0
    lg = pkg.geometry.LatticeGeometry(geo, lattice, periodic)
    t = lg.all_ranks()[0]
    bulk = t.bulk_nodes()
    bb = t.halfway_bb(t.fluid_bnd_nodes())
    f0 = pkg.cases.std_case_initial_state(t, np.ones(geo.shape))[0]
    force = (1e-6, 2e-7, 0.0) if lattice == "D2Q9" else (1e-6, 2e-7, -3e-7)
    lat = pkg.capi.Lattice.from_rank_tables(t)
    lat.add_halfway_bb(*bb)
    lat.finalize(1)
    lat.upload(f0)
    ref = port.PortRank(pkg.geometry.LATTICE_ID[lattice], t.neigh, bulk, 1, bb)
    ref.f[:] = f0
    share = lat.index_skipped_word_fraction()
    assert 0.0 <= share <= 1.0
    if lattice == "D2Q9":
        assert share > 0.5, share           # an open channel is almost all plain runs
    for on, steps in ((False, 3), (True, 5), (False, 2), (True, 4)):
        lat.set_index_skip_mask(on)
        lat.step_single(steps, tau=0.7, force=force)
        ref.step_std_case(steps, tau=0.7, force=force)
        assert np.array_equal(lat.download()[bulk], ref.f[bulk]), (on, steps)
        assert np.array_equal(lat.download_rho()[bulk, 0], ref.rho[bulk, 0])
        assert np.array_equal(lat.download_vel()[bulk], ref.vel[bulk])
    lat.close()
