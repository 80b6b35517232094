"""ctypes binding of the C-ABI (include/chimp_b200.h -> libchimp_b200.so).

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible
when a lattice is created, the call raises.  Host arrays are numpy arrays in the reference's
layouts (LbField AoS [(nFields*nQ)*node + nQ*field + q], reference node labels).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import geometry as G

_HERE = os.path.dirname(os.path.abspath(__file__))
# CHIMP_LIB selects an alternative build of the same library (tuning experiments only)
LIB_PATH = os.environ.get("CHIMP_LIB") or os.path.join(_HERE, "libchimp_b200.so")
_lib = None

BGK, TRT = 0, 1
INDEX_TABLE, INDEX_COMPACT = 0, 1
LINK_SOLID, LINK_PRESSURE, LINK_FLUID_SWAP = 0, 1, 2


class ChimpError(RuntimeError):
    pass


class SingleParams(C.Structure):
    _fields_ = [("collision", C.c_int), ("tau", C.c_double), ("tau_sym", C.c_double), ("tau_anti", C.c_double),
                ("force", C.c_double * 3)]


class TwoPhaseParams(C.Structure):
    _fields_ = [("tau0", C.c_double), ("tau1", C.c_double), ("sigma", C.c_double), ("beta", C.c_double),
                ("momx", C.c_double), ("force", C.c_double * 3), ("n_fluid_global", C.c_longlong)]


EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)

# every symbol include/chimp_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "chimp_last_error", "chimp_version", "chimp_launch_count", "chimp_lattice_nq", "chimp_lattice_nd",
    "chimp_lattice_c", "chimp_lattice_w", "chimp_lattice_reverse", "chimp_create", "chimp_add_halfway_bb",
    "chimp_add_links", "chimp_add_constant_links", "chimp_add_neighbor", "chimp_set_solid_boundary", "chimp_build_host", "chimp_finalize",
    "chimp_create_from_device_table", "chimp_destroy", "chimp_upload_lbfield", "chimp_download_lbfield",
    "chimp_download_rho", "chimp_download_vel", "chimp_set_one_phase_attributes", "chimp_step_single",
    "chimp_set_twophase_density", "chimp_step_twophase", "chimp_download_phase_field", "chimp_last_flux_force",
    "chimp_num_neighbors", "chimp_neighbor_info", "chimp_send_buffer_dev", "chimp_recv_buffer_dev",
    "chimp_set_exchange_callback", "chimp_set_stream", "chimp_synchronize", "chimp_num_own_nodes",
    "chimp_host_table_info", "chimp_host_table", "chimp_host_halo_lists", "chimp_host_constant_links", "chimp_irregular_fraction",
    "chimp_index_bytes_per_node", "chimp_phi_index_bytes_per_node", "chimp_set_index_skip_mask", "chimp_index_skipped_word_fraction", "chimp_one_phase_attribute_bytes_per_node", "chimp_plane_stride", "chimp_step_timed", "chimp_step_twophase_timed", "chimp_peer_mode", "chimp_voxel_table_host", "chimp_create_from_voxels",
    "chimp_voxel_phi_table_host", "chimp_set_phi_table_from_voxels", "chimp_slab_tables_host", "chimp_create_slab_from_voxels",
    "chimp_halo_face_recv_count", "chimp_halo_face_recv_list", "chimp_init_uniform",
    "chimp_download_moments_device_order", "chimp_step_begin", "chimp_step_end", "chimp_download_mass_change",
    "chimp_set_halo_buffers", "chimp_halo_stream", "chimp_add_halo_face", "chimp_set_boundary_count", "chimp_set_scalar_exchange_callback",
    "chimp_set_allreduce_callback", "chimp_scalar_neighbor_info", "chimp_init_equilibrium_dev", "chimp_set_phi_table_dev", "chimp_flux_force", "chimp_capillary_force", "chimp_node_list_flux", "chimp_add_scalar_halo_face", "chimp_ipc_handles_twophase", "chimp_local_pointers_twophase", "chimp_connect_peer_scalar", "chimp_connect_world", "chimp_host_scalar_halo_lists", "chimp_ipc_handles", "chimp_local_pointers", "chimp_connect_peer", "chimp_scalar_send_buffer_dev", "chimp_scalar_recv_buffer_dev",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ChimpError("%s is missing: build it with __graft_entry__.build() (nvcc, sm_100a); "
                             "there is no CPU fallback" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.chimp_last_error.restype = C.c_char_p
        l.chimp_launch_count.restype = C.c_longlong
        l.chimp_lattice_w.restype = C.c_double
        l.chimp_last_flux_force.restype = C.c_double
        l.chimp_last_flux_force.argtypes = [C.c_void_p]
        l.chimp_irregular_fraction.restype = C.c_double
        l.chimp_irregular_fraction.argtypes = [C.c_void_p]
        l.chimp_index_bytes_per_node.restype = C.c_double
        l.chimp_index_bytes_per_node.argtypes = [C.c_void_p]
        l.chimp_index_skipped_word_fraction.restype = C.c_double
        l.chimp_index_skipped_word_fraction.argtypes = [C.c_void_p]
        l.chimp_one_phase_attribute_bytes_per_node.restype = C.c_double
        l.chimp_one_phase_attribute_bytes_per_node.argtypes = [C.c_void_p]
        l.chimp_phi_index_bytes_per_node.restype = C.c_double
        l.chimp_phi_index_bytes_per_node.argtypes = [C.c_void_p]
        l.chimp_plane_stride.restype = C.c_longlong
        l.chimp_plane_stride.argtypes = [C.c_void_p]
        l.chimp_send_buffer_dev.restype = C.c_void_p
        l.chimp_send_buffer_dev.argtypes = [C.c_void_p, C.c_int]
        l.chimp_recv_buffer_dev.restype = C.c_void_p
        l.chimp_recv_buffer_dev.argtypes = [C.c_void_p, C.c_int]
        for name in ("chimp_scalar_send_buffer_dev", "chimp_scalar_recv_buffer_dev"):
            getattr(l, name).restype = C.c_void_p
            getattr(l, name).argtypes = [C.c_void_p, C.c_int]
        l.chimp_halo_stream.restype = C.c_void_p
        l.chimp_halo_stream.argtypes = [C.c_void_p]
        l.chimp_set_halo_buffers.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        l.chimp_destroy.restype = None
        l.chimp_destroy.argtypes = [C.c_void_p]
        _lib = l
    return _lib


def _check(rc):
    if rc != 0:
        raise ChimpError(lib().chimp_last_error().decode())


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Lattice:
    """One rank's engine context (chimp_lattice)."""

    def __init__(self, lattice: str, neigh, bulk, n_fields=1, device=-1):
        self.lattice = lattice
        self.nq = len(G.BASIS[lattice])
        self.nd = G.BASIS[lattice].shape[1]
        neigh = _i32(neigh)
        bulk = _i32(bulk)
        self.n_nodes = neigh.shape[0]
        self.n_fields = n_fields
        self.h = C.c_void_p()
        _check(lib().chimp_create(C.byref(self.h), G.LATTICE_ID[lattice], C.c_int(self.n_nodes), _p(neigh),
                                  C.c_int(len(bulk)), _p(bulk), C.c_int(n_fields), C.c_int(device)))
        self._cb = None

    @classmethod
    def from_rank_tables(cls, tab: G.RankTables, n_fields=1, device=-1):
        return cls(tab.g.lattice, tab.neigh, tab.bulk_nodes(), n_fields, device)

    def close(self):
        if self.h:
            lib().chimp_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- boundary objects ---------------------------------------------------------------
    def add_halfway_bb(self, nodes, n_beta, n_gamma, n_delta, links):
        nodes, n_beta, n_gamma, n_delta, links = map(_i32, (nodes, n_beta, n_gamma, n_delta, links))
        _check(lib().chimp_add_halfway_bb(self.h, C.c_int(len(nodes)), _p(nodes), _p(n_beta), _p(n_gamma), _p(n_delta),
                                          _p(links)))

    def add_links(self, kind, links4):
        links4 = _i32(links4).reshape(-1, 4)
        _check(lib().chimp_add_links(self.h, C.c_int(kind), C.c_int(len(links4)), _p(links4)))

    def add_constant_links(self, node_q, values):
        """PressureBnd / InletOutlet (LBpressurebnd.h:10-88): f(q_k, node_k) = values_k after every step's boundary phase;
        node_q is [n, 2] (destination node label, direction)"""
        nq = np.ascontiguousarray(node_q, dtype=np.int32).reshape(-1, 2)
        v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        assert len(nq) == len(v)
        _check(lib().chimp_add_constant_links(self.h, C.c_int(len(v)), _p(nq), _p(v)))

    def add_neighbor(self, rank, send_nodes, send_ndir, send_dirs, recv_nodes, recv_ndir, recv_dirs):
        a = [_i32(x) for x in (send_nodes, send_ndir, send_dirs, recv_nodes, recv_ndir, recv_dirs)]
        _check(lib().chimp_add_neighbor(self.h, C.c_int(rank), C.c_int(len(a[0])), _p(a[0]), _p(a[1]), _p(a[2]),
                                        C.c_int(len(a[3])), _p(a[3]), _p(a[4]), _p(a[5])))

    def set_solid_boundary(self, nodes):
        nodes = _i32(nodes)
        _check(lib().chimp_set_solid_boundary(self.h, C.c_int(len(nodes)), _p(nodes)))

    def build_host(self, boundary_first=False):
        """runs only the host-side table builder (no CUDA); used by the CPU tests"""
        _check(lib().chimp_build_host(self.h, C.c_int(1 if boundary_first else 0)))

    def finalize(self, index_form=INDEX_COMPACT, boundary_first=False):
        _check(lib().chimp_finalize(self.h, C.c_int(index_form), C.c_int(1 if boundary_first else 0)))

    def host_table(self):
        """(table [nQ, n], labels [n], pmask [n], info dict) of the host builder"""
        info = (C.c_longlong * 6)()
        _check(lib().chimp_host_table_info(self.h, info))
        n, n_pad, n_halo, stride, n_boundary, _ = [int(x) for x in info]
        table = np.zeros((self.nq, n), dtype=np.int32)
        labels = np.zeros(n, dtype=np.int32)
        pmask = np.zeros(n, dtype=np.uint32)
        _check(lib().chimp_host_table(self.h, _p(table), _p(labels), _p(pmask)))
        return table, labels, pmask, dict(n=n, n_pad=n_pad, n_halo=n_halo, stride=stride, n_boundary=n_boundary)

    def host_halo_lists(self, k):
        """(send_src, recv_dst) slot offsets (q*stride + slot) of neighbour k"""
        rank = C.c_int()
        ns, nr = C.c_longlong(), C.c_longlong()
        _check(lib().chimp_neighbor_info(self.h, C.c_int(k), C.byref(rank), C.byref(ns), C.byref(nr)))
        ns_f, nr_f = ns.value // self.n_fields, nr.value // self.n_fields
        src = np.zeros(ns_f, dtype=np.int64)
        dst = np.zeros(nr_f, dtype=np.int64)
        _check(lib().chimp_host_halo_lists(self.h, C.c_int(k), _p(src), _p(dst)))
        return rank.value, src, dst

    def host_constant_links(self):
        """(dst slot offsets q*stride + slot, values) of the constant links, in registration order"""
        n = lib().chimp_host_constant_links(self.h, None, None)
        if n < 0:
            raise ChimpError(lib().chimp_last_error().decode())
        dst, val = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.float64)
        lib().chimp_host_constant_links(self.h, _p(dst), _p(val))
        return dst, val

    def host_scalar_recv_slots(self, k):
        """ghost phi slots of neighbour k, in the order its values arrive"""
        ns, nr = self.scalar_neighbor_info(k)
        dst = np.zeros(nr, dtype=np.int64)
        _check(lib().chimp_host_scalar_halo_lists(self.h, C.c_int(k), None, _p(dst)))
        return dst

    # ---- state --------------------------------------------------------------------------
    def upload(self, f_aos):
        f_aos = _f64(f_aos)
        assert f_aos.size == self.n_nodes * self.n_fields * self.nq
        _check(lib().chimp_upload_lbfield(self.h, _p(f_aos)))

    def download(self, out=None):
        if out is None:
            out = np.zeros((self.n_nodes, self.n_fields, self.nq))
        assert out.flags["C_CONTIGUOUS"] and out.size == self.n_nodes * self.n_fields * self.nq
        _check(lib().chimp_download_lbfield(self.h, _p(out)))
        return out

    def download_rho(self, out=None):
        if out is None:
            out = np.zeros((self.n_nodes, self.n_fields))
        _check(lib().chimp_download_rho(self.h, _p(out), C.c_int(out.shape[1])))
        return out

    def download_vel(self, out=None):
        if out is None:
            out = np.zeros((self.n_nodes, self.nd))
        _check(lib().chimp_download_vel(self.h, _p(out)))
        return out

    def download_phase_field(self, out=None):
        if out is None:
            out = np.zeros(self.n_nodes)
        _check(lib().chimp_download_phase_field(self.h, _p(out)))
        return out

    def set_one_phase_attributes(self, force_on, interior, add_source, scale, rho_w=1.0):
        force_on, add_source, scale = _f64(force_on), _f64(add_source), _f64(scale)
        interior = _i32(interior)
        _check(lib().chimp_set_one_phase_attributes(self.h, _p(force_on), _p(interior), _p(add_source),
                                                    C.c_int(len(scale)), _p(scale), C.c_double(rho_w)))

    def set_twophase_density(self, rho2):
        rho2 = _f64(rho2)
        assert rho2.size == self.n_nodes * 2
        _check(lib().chimp_set_twophase_density(self.h, _p(rho2)))

    # ---- stepping -------------------------------------------------------------------------
    def step_single(self, n_steps, tau=0.8, force=(0.0, 0.0, 0.0), trt=None):
        p = SingleParams()
        p.collision = TRT if trt else BGK
        p.tau = tau
        p.tau_sym, p.tau_anti = trt if trt else (0.0, 0.0)
        F = list(force) + [0.0] * (3 - len(force))
        p.force[0], p.force[1], p.force[2] = F[0], F[1], F[2]
        _check(lib().chimp_step_single(self.h, C.byref(p), C.c_int(n_steps)))

    def step_twophase(self, n_steps, tau0, tau1, sigma, beta, momx, force, n_fluid_global):
        p = TwoPhaseParams()
        p.tau0, p.tau1, p.sigma, p.beta, p.momx = tau0, tau1, sigma, beta, momx
        F = list(force) + [0.0] * (3 - len(force))
        p.force[0], p.force[1], p.force[2] = F[0], F[1], F[2]
        p.n_fluid_global = int(n_fluid_global)
        _check(lib().chimp_step_twophase(self.h, C.byref(p), C.c_int(n_steps)))

    def _twophase_params(self, tau0, tau1, sigma, beta, momx, force, n_fluid_global):
        p = TwoPhaseParams()
        p.tau0, p.tau1, p.sigma, p.beta, p.momx = tau0, tau1, sigma, beta, momx
        F = list(force) + [0.0] * (3 - len(force))
        p.force[0], p.force[1], p.force[2] = F[0], F[1], F[2]
        p.n_fluid_global = int(n_fluid_global)
        return p

    def step_twophase_timed(self, n_steps, tau0, tau1, sigma, beta, momx, force, n_fluid_global):
        """runs n_steps and returns their device time in ms (CUDA events on the engine's own stream)"""
        p = self._twophase_params(tau0, tau1, sigma, beta, momx, force, n_fluid_global)
        ms = C.c_double()
        _check(lib().chimp_step_twophase_timed(self.h, C.byref(p), C.c_int(n_steps), C.byref(ms)))
        return ms.value

    def _single_params(self, tau, force, trt):
        p = SingleParams()
        p.collision = TRT if trt else BGK
        p.tau = tau
        p.tau_sym, p.tau_anti = trt if trt else (0.0, 0.0)
        F = list(force) + [0.0] * (3 - len(force))
        p.force[0], p.force[1], p.force[2] = F[0], F[1], F[2]
        return p

    def step_begin(self, tau=0.8, force=(0.0, 0.0, 0.0), trt=None, store_moments=True):
        p = self._single_params(tau, force, trt)
        _check(lib().chimp_step_begin(self.h, C.byref(p), C.c_int(1 if store_moments else 0)))

    def step_end(self):
        _check(lib().chimp_step_end(self.h))

    def download_mass_change(self, n_labels):
        out = np.zeros(n_labels)
        _check(lib().chimp_download_mass_change(self.h, _p(out)))
        return out

    def flux_force(self, field_no, cart_dir, fixed_flux, n_nodes_global):
        """calcFluxForceCartDir (LBglobalforcing.h:8-33)"""
        out = C.c_double(0.0)
        _check(lib().chimp_flux_force(self.h, C.c_int(field_no), C.c_int(cart_dir), C.c_double(fixed_flux),
                                      C.c_longlong(n_nodes_global), C.byref(out)))
        return out.value

    def capillary_force(self, cart_dir, sigma_cap_numb, nu0, nu1, n_nodes_global):
        """calcCapNumbForceCartDir (LBglobalforcing.h:35-98)"""
        out = C.c_double(0.0)
        _check(lib().chimp_capillary_force(self.h, C.c_int(cart_dir), C.c_double(sigma_cap_numb), C.c_double(nu0), C.c_double(nu1),
                                           C.c_longlong(n_nodes_global), C.byref(out)))
        return out.value

    def node_list_flux(self, nodes, bins, n_bins, field_no=0, component=2):
        """sum over the list of vel(component, n) * rho(field, n) per bin, in list order
        (mass flux through the pressure nodes, std_one_phase/main.cpp:607-619)"""
        nodes, bins = _i32(nodes), _i32(bins)
        out = np.zeros(n_bins)
        _check(lib().chimp_node_list_flux(self.h, C.c_int(len(nodes)), _p(nodes), _p(bins), C.c_int(n_bins), C.c_int(field_no),
                                          C.c_int(component), _p(out)))
        return out

    def set_halo_buffers(self, k, send_ptr, recv_ptr):
        _check(lib().chimp_set_halo_buffers(self.h, C.c_int(k), C.c_void_p(send_ptr), C.c_void_p(recv_ptr)))

    def add_halo_face(self, rank, send_src, recv_dst):
        send_src = np.ascontiguousarray(send_src, dtype=np.int64)
        recv_dst = np.ascontiguousarray(recv_dst, dtype=np.int64)
        if not hasattr(self, "_faces") or self._faces is None:
            self._faces = []
        self._faces.append((send_src, recv_dst))
        _check(lib().chimp_add_halo_face(self.h, C.c_int(rank), C.c_longlong(len(send_src)), _p(send_src),
                                         C.c_longlong(len(recv_dst)), _p(recv_dst)))

    def add_scalar_halo_face(self, k, send_src, recv_dst, send_ptr=None, recv_ptr=None):
        send_src = np.ascontiguousarray(send_src, dtype=np.int64)
        recv_dst = np.ascontiguousarray(recv_dst, dtype=np.int64)
        _check(lib().chimp_add_scalar_halo_face(self.h, C.c_int(k), C.c_longlong(len(send_src)), _p(send_src),
                                                C.c_longlong(len(recv_dst)), _p(recv_dst), C.c_void_p(send_ptr), C.c_void_p(recv_ptr)))

    def set_boundary_count(self, n):
        _check(lib().chimp_set_boundary_count(self.h, C.c_int(n)))

    def ipc_handles(self):
        buf = (C.c_ubyte * 192)()
        _check(lib().chimp_ipc_handles(self.h, buf))
        return bytes(buf)

    def local_pointers(self):
        out = (C.c_void_p * 3)()
        _check(lib().chimp_local_pointers(self.h, out))
        return [int(x) if x else 0 for x in out]

    def recv_dst(self, k):
        """slot offsets (q*plane_stride + slot) of my receive list for neighbour / face k"""
        faces = getattr(self, "_faces", None)
        if faces:
            return faces[k][1]
        lib().chimp_halo_face_recv_count.restype = C.c_longlong
        cnt = int(lib().chimp_halo_face_recv_count(self.h, C.c_int(k)))
        if cnt > 0:      # structured-ingest face registered by the library itself (chimp_create_slab_from_voxels)
            dst = np.zeros(cnt, dtype=np.int64)
            _check(lib().chimp_halo_face_recv_list(self.h, C.c_int(k), _p(dst)))
            return dst
        return self.host_halo_lists(k)[2]

    def connect_peer(self, k, peer_field_stride, peer_face, peer_dst, handles=None, pointers=None):
        peer_dst = np.ascontiguousarray(peer_dst, dtype=np.int64)
        hb = (C.c_ubyte * 192).from_buffer_copy(handles) if handles is not None else None
        pp = (C.c_void_p * 3)(*pointers) if pointers is not None else None
        _check(lib().chimp_connect_peer(self.h, C.c_int(k), hb, C.c_int(1 if pointers is not None else 0), pp,
                                        C.c_longlong(peer_field_stride), C.c_int(peer_face), C.c_longlong(len(peer_dst)),
                                        _p(peer_dst)))

    def ipc_handles_twophase(self):
        buf = (C.c_ubyte * 128)()
        _check(lib().chimp_ipc_handles_twophase(self.h, buf))
        return bytes(buf)

    def local_pointers_twophase(self):
        out = (C.c_void_p * 2)()
        _check(lib().chimp_local_pointers_twophase(self.h, out))
        return [int(x) if x else 0 for x in out]

    def connect_peer_scalar(self, k, peer_phi_dst, handle=None, pointer=None):
        """scalar (phi) halo of face k over peer memory: the neighbour's phi array (IPC handle, or raw pointer when
        both lattices live in this process) and the ghost slot each of my outgoing colours lands in"""
        peer_phi_dst = np.ascontiguousarray(peer_phi_dst, dtype=np.int64)
        hb = (C.c_ubyte * 64).from_buffer_copy(handle) if handle is not None else None
        _check(lib().chimp_connect_peer_scalar(self.h, C.c_int(k), hb, C.c_int(1 if pointer is not None else 0),
                                               C.c_void_p(pointer), C.c_longlong(len(peer_phi_dst)), _p(peer_phi_dst)))

    def connect_world(self, rank, world, handles=None, pointers=None):
        """mailboxes of all ranks (rank order) for the global momentum sum"""
        hb = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles)) if handles is not None else None
        pp = (C.c_void_p * world)(*pointers) if pointers is not None else None
        _check(lib().chimp_connect_world(self.h, C.c_int(rank), C.c_int(world), hb, C.c_int(1 if pointers is not None else 0), pp))

    def peer_mode(self):
        """(mode, why): 0 no peer halos, 1 separate push launches, 2 fused into the step kernel"""
        buf = C.create_string_buffer(256)
        mode = int(lib().chimp_peer_mode(self.h, buf, C.c_int(256)))
        return mode, buf.value.decode()

    def plane_stride(self):
        return int(lib().chimp_plane_stride(self.h))

    def halo_stream(self):
        return lib().chimp_halo_stream(self.h)

    def init_equilibrium_dev(self, rho_ptr):
        _check(lib().chimp_init_equilibrium_dev(self.h, C.c_void_p(rho_ptr)))

    def set_phi_table_dev(self, ptable_ptr, n_extra, phi_extra_ptr):
        _check(lib().chimp_set_phi_table_dev(self.h, C.c_void_p(ptable_ptr), C.c_int(n_extra), C.c_void_p(phi_extra_ptr)))

    def init_uniform(self, rho=1.0):
        _check(lib().chimp_init_uniform(self.h, C.c_double(rho)))

    def download_moments_device_order(self):
        n = self.num_own_nodes()
        rho = np.zeros(n)
        vel = np.zeros((self.nd, n))
        _check(lib().chimp_download_moments_device_order(self.h, _p(rho), _p(vel)))
        return rho, vel

    def last_flux_force(self):
        return float(lib().chimp_last_flux_force(self.h))

    def step_timed(self, n_steps, tau=0.8, force=(0.0, 0.0, 0.0), trt=None):
        """runs n_steps and returns the device time in ms of the collide-stream launches
        (CUDA events on the engine's own stream)"""
        p = SingleParams()
        p.collision = TRT if trt else BGK
        p.tau = tau
        p.tau_sym, p.tau_anti = trt if trt else (0.0, 0.0)
        F = list(force) + [0.0] * (3 - len(force))
        p.force[0], p.force[1], p.force[2] = F[0], F[1], F[2]
        ms = C.c_double()
        _check(lib().chimp_step_timed(self.h, C.byref(p), C.c_int(n_steps), C.byref(ms)))
        return ms.value

    # ---- halo / streams -------------------------------------------------------------------
    def num_neighbors(self):
        return int(lib().chimp_num_neighbors(self.h))

    def neighbor_info(self, k):
        rank = C.c_int()
        ns, nr = C.c_longlong(), C.c_longlong()
        _check(lib().chimp_neighbor_info(self.h, C.c_int(k), C.byref(rank), C.byref(ns), C.byref(nr)))
        return rank.value, ns.value, nr.value

    def send_buffer_ptr(self, k):
        return lib().chimp_send_buffer_dev(self.h, C.c_int(k))

    def recv_buffer_ptr(self, k):
        return lib().chimp_recv_buffer_dev(self.h, C.c_int(k))

    def set_exchange_callback(self, fn):
        """fn(stream_ptr:int) -> None is called once per step on the halo stream"""
        def tramp(_user, stream):
            try:
                fn(int(stream) if stream else 0)
                return 0
            except Exception as exc:  # pragma: no cover
                print("exchange callback failed:", exc)
                return 1
        self._cb = EXCHANGE_FN(tramp)
        _check(lib().chimp_set_exchange_callback(self.h, self._cb, None))

    def set_scalar_exchange_callback(self, fn):
        """fn(stream_ptr) moves the packed phi values of every neighbour (scalar halo)"""
        def tramp(_user, stream):
            try:
                fn(int(stream) if stream else 0)
                return 0
            except Exception as exc:  # pragma: no cover
                print("scalar exchange callback failed:", exc)
                return 1
        self._cb_scalar = EXCHANGE_FN(tramp)
        _check(lib().chimp_set_scalar_exchange_callback(self.h, self._cb_scalar, None))

    def set_allreduce_callback(self, fn):
        """fn(dev_ptr, count, stream_ptr) sums `count` doubles at dev_ptr over all ranks, in place"""
        def tramp(_user, dev, count, stream):
            try:
                fn(int(dev), int(count), int(stream) if stream else 0)
                return 0
            except Exception as exc:  # pragma: no cover
                print("allreduce callback failed:", exc)
                return 1
        self._cb_allreduce = ALLREDUCE_FN(tramp)
        _check(lib().chimp_set_allreduce_callback(self.h, self._cb_allreduce, None))

    def scalar_neighbor_info(self, k):
        ns, nr = C.c_longlong(), C.c_longlong()
        _check(lib().chimp_scalar_neighbor_info(self.h, C.c_int(k), C.byref(ns), C.byref(nr)))
        return ns.value, nr.value

    def scalar_send_buffer_ptr(self, k):
        return lib().chimp_scalar_send_buffer_dev(self.h, C.c_int(k))

    def scalar_recv_buffer_ptr(self, k):
        return lib().chimp_scalar_recv_buffer_dev(self.h, C.c_int(k))

    def set_stream(self, ptr):
        _check(lib().chimp_set_stream(self.h, C.c_void_p(ptr)))

    def synchronize(self):
        _check(lib().chimp_synchronize(self.h))

    def num_own_nodes(self):
        return int(lib().chimp_num_own_nodes(self.h))

    def irregular_fraction(self):
        return float(lib().chimp_irregular_fraction(self.h))

    def index_bytes_per_node(self):
        return float(lib().chimp_index_bytes_per_node(self.h))

    def set_index_skip_mask(self, on=True):
        """compact index, plain single-field step: kernel form that skips the delta words marked as plain runs"""
        _check(lib().chimp_set_index_skip_mask(self.h, C.c_int(1 if on else 0)))

    def index_skipped_word_fraction(self):
        return float(lib().chimp_index_skipped_word_fraction(self.h))

    def one_phase_attribute_bytes_per_node(self):
        """one_phase lattices: 4 = attributes packed into one word per node, 24 = four arrays"""
        return float(lib().chimp_one_phase_attribute_bytes_per_node(self.h))

    def phi_index_bytes_per_node(self):
        """two-field lattices: bytes per node of the index behind the colour gradient (76 = table, ~6 = derived form)"""
        return float(lib().chimp_phi_index_bytes_per_node(self.h))


def _periodic_mask(periodic: str):
    return sum(1 << k for k, name in enumerate("xyz") if name in periodic.lower())


def voxel_table_host(lattice: str, voxels, periodic="xyz"):
    """host half of the voxel ingest (no CUDA): (table int32 [nQ, n_pad], labels int32 [n_pad], n)"""
    v = np.ascontiguousarray(voxels, dtype=np.uint8)
    nx, ny, nz = (list(v.shape) + [1])[:3]
    n, n_pad = C.c_int(), C.c_int()
    args = (C.c_int(G.LATTICE_ID[lattice]), C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(v), C.c_int(_periodic_mask(periodic)))
    _check(lib().chimp_voxel_table_host(*args, C.byref(n), C.byref(n_pad), None, None))
    nq = len(G.BASIS[lattice])
    table = np.zeros((nq, n_pad.value), dtype=np.int32)
    labels = np.zeros(n_pad.value, dtype=np.int32)
    _check(lib().chimp_voxel_table_host(*args, C.byref(n), C.byref(n_pad), _p(table), _p(labels)))
    return table, labels, n.value


def voxel_phi_table_host(lattice: str, voxels, wall_phi, periodic="xyz"):
    """host half of the colour-gradient tables: (ptable int32 [nQ, n_pad], n_extra, phi_extra [n_extra])"""
    v = np.ascontiguousarray(voxels, dtype=np.uint8)
    w = _f64(wall_phi)
    nx, ny, nz = (list(v.shape) + [1])[:3]
    args = (C.c_int(G.LATTICE_ID[lattice]), C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(v), C.c_int(_periodic_mask(periodic)))
    n_extra = C.c_int()
    _check(lib().chimp_voxel_phi_table_host(*args, _p(w), C.byref(n_extra), None, None))
    n = int(np.count_nonzero(v))
    n_pad = ((n + 31) // 32) * 32
    ptable = np.zeros((len(G.BASIS[lattice]), n_pad), dtype=np.int32)
    extra = np.zeros(max(n_extra.value, 1))
    _check(lib().chimp_voxel_phi_table_host(*args, _p(w), C.byref(n_extra), _p(ptable), _p(extra)))
    return ptable, n_extra.value, extra[: n_extra.value]


def slab_tables_host(lattice: str, voxels_ext):
    """host half of the z-slab voxel ingest (no CUDA): dict like ingest.build_slab_tables returns (numpy arrays)"""
    v = np.ascontiguousarray(voxels_ext, dtype=np.uint8)
    nx, ny, nze = v.shape
    info = (C.c_longlong * 8)()
    args = (C.c_int(G.LATTICE_ID[lattice]), C.c_int(nx), C.c_int(ny), C.c_int(nze - 2), _p(v))
    _check(lib().chimp_slab_tables_host(*args, info, None, None, None, None, None, None))
    n, n_pad, n_halo, n_boundary, sd, rd, su, ru = [int(x) for x in info]
    nq = len(G.BASIS[lattice])
    table = np.zeros((nq, n_pad), dtype=np.int32)
    labels = np.zeros(n_pad, dtype=np.int32)
    lists = [np.zeros(k, dtype=np.int64) for k in (sd, rd, su, ru)]
    _check(lib().chimp_slab_tables_host(*args, info, _p(table), _p(labels), *[_p(a) for a in lists]))
    return dict(table=table, labels=labels, n=n, n_pad=n_pad, n_halo=n_halo, n_boundary=n_boundary, stride=n_pad + n_halo,
                faces={"down": (lists[0], lists[1]), "up": (lists[2], lists[3])})


def slab_lattice_from_voxels(lattice: str, voxels_ext, rank_down, rank_up, index_form=INDEX_COMPACT, device=-1):
    """chimp_create_slab_from_voxels: one rank's z-slab with its two halo faces registered"""
    v = np.ascontiguousarray(voxels_ext, dtype=np.uint8)
    nx, ny, nze = v.shape
    obj = Lattice.__new__(Lattice)
    obj.lattice = lattice
    obj.nq = len(G.BASIS[lattice])
    obj.nd = G.BASIS[lattice].shape[1]
    obj.n_nodes = int(np.count_nonzero(v[:, :, 1:-1])) + 1
    obj.n_fields = 1
    obj._cb = None
    obj._faces = None
    obj.h = C.c_void_p()
    _check(lib().chimp_create_slab_from_voxels(C.byref(obj.h), G.LATTICE_ID[lattice], C.c_int(nx), C.c_int(ny), C.c_int(nze - 2), _p(v),
                                               C.c_int(1), C.c_int(index_form), C.c_int(device), C.c_int(rank_down), C.c_int(rank_up)))
    return obj


def lattice_from_voxels(lattice: str, voxels, periodic="xyz", n_fields=1, index_form=INDEX_COMPACT, device=-1, wall_phi=None):
    """chimp_create_from_voxels: the route a C / C++ main takes from a raw voxel array to a lattice"""
    v = np.ascontiguousarray(voxels, dtype=np.uint8)
    nx, ny, nz = (list(v.shape) + [1])[:3]
    obj = Lattice.__new__(Lattice)
    obj.lattice = lattice
    obj.nq = len(G.BASIS[lattice])
    obj.nd = G.BASIS[lattice].shape[1]
    obj.n_nodes = int(np.count_nonzero(v)) + 1
    obj.n_fields = n_fields
    obj._cb = None
    obj.h = C.c_void_p()
    _check(lib().chimp_create_from_voxels(C.byref(obj.h), G.LATTICE_ID[lattice], C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(v),
                                          C.c_int(_periodic_mask(periodic)), C.c_int(n_fields), C.c_int(index_form), C.c_int(device)))
    if wall_phi is not None:
        w = _f64(wall_phi)
        _check(lib().chimp_set_phi_table_from_voxels(obj.h, C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(v),
                                                     C.c_int(_periodic_mask(periodic)), _p(w)))
    return obj


def lattice_from_device_table(lattice: str, n_bulk, n_pad, n_halo, table_ptr, label_ptr, n_fields=1,
                              index_form=INDEX_COMPACT, device=-1):
    """structured-ingest path: pull table already resident on the device (int32 [nQ][n_pad])"""
    obj = Lattice.__new__(Lattice)
    obj.lattice = lattice
    obj.nq = len(G.BASIS[lattice])
    obj.nd = G.BASIS[lattice].shape[1]
    obj.n_nodes = n_bulk + 1   # labels 1..N are the leading rows of a reference-layout field (row 0 = dummy node)
    obj.n_fields = n_fields
    obj._cb = None
    obj.h = C.c_void_p()
    _check(lib().chimp_create_from_device_table(C.byref(obj.h), G.LATTICE_ID[lattice], C.c_int(n_bulk), C.c_int(n_pad),
                                                C.c_int(n_halo), C.c_void_p(table_ptr), C.c_void_p(label_ptr),
                                                C.c_int(n_fields), C.c_int(index_form), C.c_int(device)))
    return obj
