"""oracle/port.py -- TEST INFRASTRUCTURE: ctypes front end of oracle/lb_port.c (the plain-C
restatement of the reference's CPU algorithm) plus the numpy restatement of the reference's
ghost exchange (MonLatMpi::communicateLbField / communicateScalarField, LBmonlatmpi.h:181-297)
so that N-rank cases can be replayed in one process.  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liblbport.so")
        src = os.path.join(_HERE, "lb_port.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-std=c11", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src, "-lm"])
        _LIB = C.CDLL(so)
    return _LIB


class PortTables(C.Structure):
    _fields_ = [("lattice", C.c_int), ("n_nodes", C.c_int), ("n_fields", C.c_int), ("neigh", C.c_void_p),
                ("n_bulk", C.c_int), ("bulk", C.c_void_p), ("n_bb", C.c_int), ("bb_node", C.c_void_p),
                ("bb_nbeta", C.c_void_p), ("bb_ngamma", C.c_void_p), ("bb_ndelta", C.c_void_p), ("bb_links", C.c_void_p)]


class OnePhase(C.Structure):
    _fields_ = [("force_on", C.c_void_p), ("interior", C.c_void_p), ("add_source", C.c_void_p), ("n_labels", C.c_int),
                ("scale", C.c_void_p), ("mass_change", C.c_void_p), ("n_solid", C.c_int), ("n_press", C.c_int),
                ("n_fluid", C.c_int), ("solid_links", C.c_void_p), ("press_links", C.c_void_p),
                ("fluid_links", C.c_void_p), ("rho_w", C.c_double)]


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class PortRank:
    """One rank's reference-layout state: f AoS [(nFields*nQ)*node + nQ*field + q] etc."""

    def __init__(self, lattice_id, neigh, bulk, n_fields=1, bb=None):
        self.lattice = lattice_id
        self.neigh = _i32(neigh)
        self.n_nodes, self.nq = self.neigh.shape
        self.nd = 2 if lattice_id == 0 else 3
        self.bulk = _i32(bulk)
        self.n_fields = n_fields
        if bb is None:
            z = np.zeros(0, dtype=np.int32)
            bb = (z, z, z, z, z)
        self.bb = [_i32(x) for x in bb]
        self.f = np.zeros((self.n_nodes, n_fields, self.nq))
        self.ftmp = np.zeros_like(self.f)
        self.rho = np.zeros((self.n_nodes, n_fields))
        self.vel = np.zeros((self.n_nodes, self.nd))
        self.cg = np.zeros(self.n_nodes)
        self._keep = []

    def tables(self):
        t = PortTables(self.lattice, self.n_nodes, self.n_fields, _p(self.neigh), len(self.bulk), _p(self.bulk),
                       len(self.bb[0]), _p(self.bb[0]), _p(self.bb[1]), _p(self.bb[2]), _p(self.bb[3]), _p(self.bb[4]))
        return t

    def set_library_bnd(self, kind, lists, rho_bnd=None, rho=1.0, vel=(0.0, 0.0, 0.0)):
        """PressureBnd ("pressure": rho_bnd[n_nodes, n_fields]) or InletOutlet ("inletoutlet": rho, vel) of
        LBpressurebnd.h:10-88 on the boundary nodes of `lists` (nodes, nBeta, nGamma, nDelta, links), applied after
        the bounce back of every std_case step"""
        self.lib_bnd = (kind, [_i32(x) for x in lists], None if rho_bnd is None else _f64(rho_bnd), float(rho),
                        _f64(list(vel) + [0.0] * (3 - len(vel))))

    def apply_library_bnd(self, fld=0):
        kind, l, rho_bnd, rho, vel = self.lib_bnd
        t = self.tables()
        if kind == "pressure":
            lib().port_pressure_bnd_apply(C.byref(t), _p(self.f), C.c_int(fld), C.c_int(len(l[0])), _p(l[0]), _p(l[1]), _p(l[2]),
                                          _p(l[3]), _p(l[4]), _p(rho_bnd))
        else:
            lib().port_inlet_outlet_apply(C.byref(t), _p(self.f), C.c_int(fld), C.c_int(len(l[0])), _p(l[0]), _p(l[1]), _p(l[2]),
                                          _p(l[3]), _p(l[4]), C.c_double(rho), _p(vel))

    def step_std_case(self, n_steps, tau=0.8, force=(0, 0, 0), trt=None, skip_boundary=False):
        if getattr(self, "lib_bnd", None) is not None and not skip_boundary:
            bnd, self.lib_bnd = self.lib_bnd, None
            try:
                for _ in range(n_steps):    # the library boundary follows the bounce back of every step
                    self.step_std_case(1, tau, force, trt)
                    self.lib_bnd = bnd
                    self.apply_library_bnd()
                    self.lib_bnd = None
            finally:
                self.lib_bnd = bnd
            return
        F = _f64(list(force) + [0.0] * (3 - len(force)))
        ts, ta = trt if trt else (0.0, 0.0)
        t = self.tables()
        lib().port_step_std_case(C.byref(t), _p(self.f), _p(self.ftmp), _p(self.rho), _p(self.vel), C.c_int(1 if trt else 0),
                                 C.c_double(tau), C.c_double(ts), C.c_double(ta), _p(F), C.c_int(n_steps),
                                 C.c_int(1 if skip_boundary else 0))

    def apply_bb(self, fld=0):
        t = self.tables()
        lib().port_apply_bb(C.byref(t), _p(self.f), C.c_int(fld))

    def set_one_phase(self, force_on, interior, add_source, scale, solid_links, press_links, fluid_links, rho_w=1.0):
        self.op_arrays = dict(force_on=_f64(force_on), interior=_i32(interior), add_source=_f64(add_source),
                              scale=_f64(scale), mass=np.zeros(len(scale)), solid=_i32(solid_links).reshape(-1, 4),
                              press=_i32(press_links).reshape(-1, 4), fluid=_i32(fluid_links).reshape(-1, 4))
        a = self.op_arrays
        self.op = OnePhase(_p(a["force_on"]), _p(a["interior"]), _p(a["add_source"]), len(a["scale"]), _p(a["scale"]),
                           _p(a["mass"]), len(a["solid"]), len(a["press"]), len(a["fluid"]), _p(a["solid"]),
                           _p(a["press"]), _p(a["fluid"]), rho_w)

    def step_one_phase(self, n_steps, tau=0.8, force=(0, 0, 0), trt=None, skip_boundary=False):
        F = _f64(list(force) + [0.0] * (3 - len(force)))
        ts, ta = trt if trt else (0.0, 0.0)
        t = self.tables()
        lib().port_step_one_phase(C.byref(t), C.byref(self.op), _p(self.f), _p(self.ftmp), _p(self.rho), _p(self.vel),
                                  C.c_int(1 if trt else 0), C.c_double(tau), C.c_double(ts), C.c_double(ta), _p(F),
                                  C.c_int(n_steps), C.c_int(1 if skip_boundary else 0))

    def apply_one_phase_links(self):
        t = self.tables()
        lib().port_apply_one_phase_links(C.byref(t), C.byref(self.op), _p(self.f), _p(self.vel))

    def step_twophase(self, n_steps, solid_bnd, tau0, tau1, sigma, beta, momx, force, n_fluid_global, skip_boundary=False):
        F = _f64(list(force) + [0.0] * (3 - len(force)))
        sb = _i32(solid_bnd)
        t = self.tables()
        lib().port_step_twophase(C.byref(t), C.c_int(len(sb)), _p(sb), _p(self.f), _p(self.ftmp), _p(self.rho), _p(self.vel),
                                 _p(self.cg), C.c_double(tau0), C.c_double(tau1), C.c_double(sigma), C.c_double(beta),
                                 C.c_double(momx), _p(F), C.c_longlong(n_fluid_global), C.c_int(n_steps),
                                 C.c_int(1 if skip_boundary else 0))
        return float(F[0])


def flux_force(ranks, field_no, cart_dir, fixed_flux, n_nodes_global):
    """calcFluxForceCartDir (LBglobalforcing.h:8-33) over all ranks: per-rank sequential sums, added in rank order
    (the in-process MPI shim of oracle/_ref adds them in rank order too)"""
    lib().port_flux_sum.restype = C.c_double
    total = 0.0
    for pr in ranks:
        t = pr.tables()
        total += lib().port_flux_sum(C.byref(t), _p(pr.f), C.c_int(field_no), C.c_int(cart_dir))
    total /= n_nodes_global
    return 2 * (fixed_flux - total)


def cap_numb_force(ranks, cart_dir, sigma_cap_numb, nu0, nu1, n_nodes_global):
    """calcCapNumbForceCartDir (LBglobalforcing.h:35-98) over all ranks"""
    sums = np.zeros(4)
    for pr in ranks:
        t = pr.tables()
        out = np.zeros(4)
        rho = _f64(pr.rho)
        lib().port_cap_numb_sums(C.byref(t), _p(pr.f), _p(rho), C.c_int(cart_dir), _p(out))
        sums += out
    sums /= n_nodes_global
    return 2 * (sigma_cap_numb - (sums[0] * nu0 + sums[1] * nu1)) / (sums[2] * nu0 + sums[3] * nu1)


def exchange_lb_field(ranks, exch, fld=0):
    """MonLatMpi::communicateLbField (LBmonlatmpi.h:236-297) for all rank pairs.
    exch[r] = list over neighbours of dict(rank, send_nodes, send_ndir, send_dirs, recv_nodes, recv_ndir, recv_dirs).
    Sends read ghost-node slots, receives write real-node slots, so order does not matter."""
    msgs = {}
    for r, pr in enumerate(ranks):
        for e in exch[r]:
            nodes = np.repeat(e["send_nodes"], e["send_ndir"])
            q = e["send_dirs"]
            ghost = pr.neigh[nodes, q]
            msgs[(r, e["rank"])] = pr.f[ghost, fld, q].copy()
    for r, pr in enumerate(ranks):
        for e in exch[r]:
            nodes = np.repeat(e["recv_nodes"], e["recv_ndir"])
            q = e["recv_dirs"]
            real = pr.neigh[nodes, q]
            pr.f[real, fld, q] = msgs[(e["rank"], r)]


def exchange_scalar(ranks, exch, field_name="cg"):
    """MonLatMpi::communicateScalarField (LBmonlatmpi.h:181-205)"""
    msgs = {}
    for r, pr in enumerate(ranks):
        for e in exch[r]:
            msgs[(r, e["rank"])] = getattr(pr, field_name)[e["send_nodes"]].copy()
    for r, pr in enumerate(ranks):
        for e in exch[r]:
            getattr(pr, field_name)[e["recv_nodes"]] = msgs[(e["rank"], r)]
