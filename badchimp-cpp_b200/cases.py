"""Host-side setup of the three target mains, restated from the reference's main() bodies
(everything that happens before the time loop; the loop itself runs on the GPU).

std_case      src/std_case/main.cpp:62-96          (init_rho attribute, f = w_q * rho)
std_one_phase src/std_one_phase/main.cpp:27-126, 253-345, 452-474   (tags, link lists, mass-source scale)
twophase      src/twophase/main_TWOPHASE.cpp:81-87, 150-205        (float32 density attributes, wettability)
"""
from __future__ import annotations

import numpy as np

from . import geometry as G


def lattice_weights(lattice: str) -> np.ndarray:
    """w[] of the reference's lattice structs (LBd2q9.h:26-32, LBd3q19.h:25-32; D3Q27 per contract)"""
    if lattice == "D2Q9":
        w0, w1, w2 = 16.0 / 36.0, 4.0 / 36.0, 1.0 / 36.0
        return np.array([w1, w2, w1, w2, w1, w2, w1, w2, w0])
    if lattice == "D3Q19":
        w0, w1, w2 = 12.0 / 36.0, 2.0 / 36.0, 1.0 / 36.0
        return np.array([w1] * 3 + [w2] * 6 + [w1] * 3 + [w2] * 6 + [w0])
    if lattice == "D3Q27":
        w0, w1, w2, w3 = 64.0 / 216.0, 16.0 / 216.0, 4.0 / 216.0, 1.0 / 216.0
        half = [w1] * 3 + [w2] * 6 + [w3] * 4
        return np.array(half + half + [w0])
    raise ValueError(lattice)


def std_case_initial_state(tab: G.RankTables, init_rho):
    """f(0,q,n) = w[q]*rho(0,n) on bulk nodes, zero elsewhere (std_case/main.cpp:62-96).
    Returns f [size, 1, nQ] and rho [size]."""
    g = tab.g
    rho = tab.attribute(g.pad_attribute(np.asarray(init_rho, dtype=np.float64)))
    f = np.zeros((tab.size, 1, g.nq))
    bulk = tab.bulk_nodes()
    f[bulk, 0, :] = lattice_weights(g.lattice)[None, :] * rho[bulk, None]
    return f, rho


def one_phase_setup(lg: G.LatticeGeometry, tabs, attrs):
    """Per rank: node tags, link lists, force switch, interior-domain labels, mass-source
    markers, and the global 1/count scale per interior domain (main.cpp:253-345, 452-458)."""
    nqnz = lg.nq - 1
    pad = {k: lg.pad_attribute(np.asarray(v)) for k, v in attrs.items()}
    per_rank = []
    global_max = 0
    for t in tabs:
        tags = t.attribute(pad["nodetags"]).astype(np.int64)
        tags[0] = -1  # Nodes::nodeTag_ default (LBnodes.h:119)
        interior = t.attribute(pad["interior_domains"]).astype(np.int32)
        force_on = t.attribute(pad["force"]).astype(np.float64)
        mine = t.node_rank == t.my_rank
        mine[0] = False
        if mine.any():
            global_max = max(global_max, int(interior[mine].max()))
        per_rank.append(dict(tags=tags, interior=interior, force_on=force_on, mine=mine))
    counts = np.zeros(global_max + 1)
    for t, pr in zip(tabs, per_rank):
        add = np.zeros(t.size)
        sel = pr["mine"] & (pr["interior"] > 0) & (pr["tags"] < 3)
        add[sel] = 1.0
        pr["add_source"] = add
        np.add.at(counts, pr["interior"][sel], 1.0)
    scale = np.zeros(global_max + 1)
    scale[1:] = 1.0 / counts[1:] if global_max else scale[1:]
    for t, pr in zip(tabs, per_rank):
        tags, mine = pr["tags"], pr["mine"]
        n = np.arange(t.size)

        def links(bit, want, need_phase=None):
            flagged = mine & (((tags >> bit) & 1) == 1)
            if need_phase is not None:
                flagged &= (tags & 3) == need_phase
            nodes = n[flagged]
            nb = t.neigh[nodes, :nqnz]
            hit = (tags[nb] & 3) == want
            rows, q = np.nonzero(hit)
            qrev = np.array([G.reverse_direction(lg.lattice, int(x)) for x in range(lg.nq)])
            return np.stack([nodes[rows], qrev[q], nb[rows, q], q], axis=1).astype(np.int32)

        pr["solid_links"] = links(3, 0)
        pr["press_links"] = links(4, 3)
        pr["fluid_links"] = links(2, 2, need_phase=1)
        pr["scale"] = scale
        # findPressureFluidNodes (main.cpp:80-93) and the phase each contributes to (:610)
        press = n[mine & (((tags >> 4) & 1) == 1)]
        pr["press_nodes"] = press.astype(np.int32)
        pr["press_phase"] = ((tags[press] & 3) - 1).astype(np.int32)
        # the initial state: rho = 1, u = 0, f = calcfeq(1, 0, 0) = w_q on bulk (main.cpp:441-474)
        f = np.zeros((t.size, 1, lg.nq))
        f[t.bulk_nodes(), 0, :] = lattice_weights(lg.lattice)[None, :]
        pr["f0"] = f
    return per_rank


def two_phase_setup(lg: G.LatticeGeometry, tabs, rho0, rho1, wettability):
    """rho(2,size) and f(2,size) at t=0 (main_TWOPHASE.cpp:150-205).  The density attributes are
    read through `float` there, so values pass through float32 and the wall value 1-val is
    float32 arithmetic."""
    p0 = lg.pad_attribute(np.asarray(rho0, dtype=np.float64))
    p1 = lg.pad_attribute(np.asarray(rho1, dtype=np.float64))
    pw = lg.pad_attribute(np.asarray(wettability, dtype=np.float64))
    w = lattice_weights(lg.lattice)
    out = []
    for t in tabs:
        rho = np.zeros((t.size, 2))
        rho[:, 0] = t.attribute(p0).astype(np.float32).astype(np.float64)
        rho[:, 1] = t.attribute(p1).astype(np.float32).astype(np.float64)
        wet = t.attribute(pw).astype(np.float32)
        sb = t.node_type == 1
        sb[0] = False
        rho[sb, 0] = wet[sb].astype(np.float64)
        rho[sb, 1] = (np.float32(1.0) - wet[sb]).astype(np.float64)
        rho[0, :] = 0.0
        f = np.zeros((t.size, 2, lg.nq))
        bulk = t.bulk_nodes()
        for fld in range(2):
            # initiateLbField (LBinitiatefield.h:52-56) with u = 0: w*rho*(1.0 + 0 + 4.5*(0 - 0))
            f[bulk, fld, :] = (w[None, :] * rho[bulk, fld, None]) * 1.0
        out.append(dict(rho=rho, f0=f, solid_bnd=t.solid_bnd_nodes()))
    return out


def one_phase_mass_flux(lat, setup, old=(0.0, 0.0)):
    """std_one_phase/main.cpp:607-633: mass flux of each phase through this rank's pressure-boundary
    nodes, q_s = 0.5 * sum vel_z * rho, and its relative change since the previous write.  The products
    are formed on the device from the moments of the last step; across ranks the caller adds the two
    local sums (MPI_Allreduce at :619)."""
    phase = setup["press_phase"]
    if len(phase) and (phase.min() < 0 or phase.max() > 1):
        raise ValueError("Fluid phase = %d in write mass flux" % int(phase[(phase < 0) | (phase > 1)][0]))
    local = lat.node_list_flux(setup["press_nodes"], phase, 2, 0, 2)
    q = 0.5 * local
    change = (q - np.asarray(old)) / (q + 1e-15)
    return local, q, change


def _c_dot_all(lattice: str, v):
    """cDotAll of the lattice structs (LBd2q9.h / LBd3q19.h:129-153): for every direction the signed components of v
    added up in axis order -- the first non-zero component starts the sum, so no 0.0 enters it"""
    basis = G.BASIS[lattice]
    out = np.zeros(len(basis))
    for q, c in enumerate(basis):
        s, first = 0.0, True
        for d, cq in enumerate(c):
            if cq == 0:
                continue
            if first:
                s, first = (v[d] if cq > 0 else -v[d]), False
            else:
                s = s + v[d] if cq > 0 else s - v[d]
        out[q] = s
    return out


def library_bnd_links(tab: G.RankTables, bnd_nodes, kind, rho_bnd=None, rho=1.0, vel=(0.0, 0.0, 0.0), field_no=0):
    """The stores PressureBnd<DXQY>::apply ("pressure", LBpressurebnd.h:19-41) or InletOutlet<DXQY>::apply
    ("inletoutlet", :51-88) perform on the boundary nodes `bnd_nodes` (classed by Boundary<DXQY>, LBboundary.h:109-160:
    the same beta / gamma / delta pairs as the bounce-back helper, isSolid being !isFluid), as the argument of
    Lattice.add_constant_links: node_q [n, 2] = (grid.neighbor(q, node), q) in the reference's store order, and the
    stored values -- w[q] * rho_bnd[node, field_no], or rho * w[q] * (1 + c2Inv cu + c4Inv0_5 (cu^2 - c2 u^2))."""
    lattice = tab.g.lattice
    w = lattice_weights(lattice)
    nodes, n_beta, n_gamma, n_delta, links = tab.halfway_bb(np.asarray(bnd_nodes, dtype=np.int32))
    nd = 2 if lattice == "D2Q9" else 3
    v = [float(x) for x in list(vel)[:nd]]
    u_sq = v[0] * v[0] + v[1] * v[1]
    if nd == 3:
        u_sq = u_sq + v[2] * v[2]
    cu = _c_dot_all(lattice, v)
    c2inv, c4inv0_5, c2 = 3.0, 4.5, 1.0 / 3.0

    def value(q, node):
        if kind == "pressure":
            return w[q] * float(rho_bnd[node, field_no])
        return rho * w[q] * (1.0 + c2inv * cu[q] + c4inv0_5 * (cu[q] * cu[q] - c2 * u_sq))

    node_q, values = [], []
    for b, node in enumerate(nodes):
        for q in links[b, :n_beta[b]]:
            node_q.append((int(tab.neigh[node, q]), int(q)))
            values.append(value(q, node))
        for q in links[b, n_beta[b] + n_gamma[b]:n_beta[b] + n_gamma[b] + n_delta[b]]:
            r = G.reverse_direction(lattice, int(q))
            node_q.append((int(tab.neigh[node, q]), int(q)))
            values.append(value(q, node))
            node_q.append((int(tab.neigh[node, r]), r))
            values.append(value(r, node))
    return np.array(node_q, dtype=np.int32).reshape(-1, 2), np.array(values, dtype=np.float64)
