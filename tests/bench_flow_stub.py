"""TEST INFRASTRUCTURE (launched by tests/test_bench_control_flow.py, never by the product): bench.py's harness
(badchimp-cpp_b200/bench_impl.run_b200) on N CPU ranks over gloo with a DO-NOTHING stand-in for the engine and for
torch.cuda.  Nothing is computed and nothing is measured; the point is the order of the collectives: every rank must
reach every all-reduce / gather / barrier, whatever its slab looks like.  (Round 2 lost its GPU budget to a rank-local
assertion between two collectives of the e2e leg; the stand-in hands every rank a different slab density on purpose.)"""
import ctypes as C
import importlib
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.distributed as dist
import helpers
pkg = helpers.load_package()
capi = pkg.capi
G = pkg.geometry

class FakeLat:
    def __init__(self, lattice, n, n_fields):
        self.lattice, self.nq, self.nd = lattice, len(G.BASIS[lattice]), G.BASIS[lattice].shape[1]
        self.n, self.n_nodes, self.n_fields, self.h = n, n + 1, n_fields, C.c_void_p(1)
        self._faces = []
    def init_uniform(self, rho): pass
    def init_equilibrium_dev(self, p): pass
    def set_phi_table_dev(self, *a): pass
    def set_one_phase_attributes(self, fo, il, add, scale, rw): assert len(fo) == self.n + 1
    def set_allreduce_callback(self, fn): pass
    def add_halo_face(self, rank, src, dst): self._faces.append((np.asarray(src), np.asarray(dst)))
    def add_scalar_halo_face(self, *a): pass
    def set_boundary_count(self, n): pass
    def ipc_handles(self): return b"\0" * 192
    def ipc_handles_twophase(self): return b"\0" * 128
    def plane_stride(self): return self.n + 64
    def connect_peer(self, k, stride, face, dst, handles=None, pointers=None): assert len(dst) == len(self._faces[k][0]), (len(dst), len(self._faces[k][0]))
    def connect_peer_scalar(self, *a, **kw): pass
    def connect_world(self, *a, **kw): pass
    def step_single(self, k, **kw): pass
    def step_twophase(self, k, *a): assert len(a) == 7
    def step_timed(self, k, **kw): return 2.5 * k * (1 + 0.01 * dist.get_rank())
    def step_twophase_timed(self, k, *a): return 3.7 * k
    def synchronize(self): pass
    def download_moments_device_order(self): return np.ones(self.n), np.zeros((self.nd, self.n))
    def download_rho(self, out=None):
        if out is None: out = np.zeros((self.n_nodes, self.n_fields))
        out[:] = 1.0 / self.n_fields; return out
    def download(self): return np.zeros((self.n_nodes, self.n_fields, self.nq))
    def irregular_fraction(self): return 0.01
    def index_bytes_per_node(self): return 23.5
    def set_index_skip_mask(self, on=True): self.skip = bool(on)
    def index_skipped_word_fraction(self): return 0.6
    def phi_index_bytes_per_node(self): return 6.0
    def one_phase_attribute_bytes_per_node(self): return 4.0
    def peer_mode(self): return (2, "")
    def _single_params(self, *a): return capi.SingleParams()
    def close(self): pass

capi.lattice_from_device_table = lambda lattice, n, n_pad, n_halo, t, l, nf=1, form=1, dev=-1: FakeLat(lattice, n, nf)

class FakeLib:
    n_rows = 0
    def chimp_launch_count(self): return 0
    def __getattr__(self, name):
        def f(h, *a):
            if name == "chimp_download_rho":
                ptr, nf = a[0], a[1].value
                # a slab's own mean density differs from 1 (this is what tripped the 8-GPU run): rank-dependent value
                v = 1.0 + (1e-5 if dist.get_rank() % 2 else -1e-5)
                np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(FakeLib.n_rows * nf,))[:] = v / nf
            return 0
        return f
capi.lib = lambda: FakeLib()

def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.empty_cache = lambda: None
    torch.cuda.device_count = lambda: world
    torch.cuda.can_device_access_peer = lambda a, b: True
    _e, _o = torch.empty, torch.ones
    def strip(kw):
        kw.pop("pin_memory", None)
        return kw
    torch.empty = lambda *a, **kw: _e(*a, **strip(kw))
    torch.ones = lambda *a, **kw: _o(*a, **strip(kw))
    real_device = torch.device
    class DevShim:
        def __call__(self, *a, **kw): return real_device("cpu")
    torch.device = DevShim()
    _init = dist.init_process_group
    dist.init_process_group = lambda backend, **kw: _init("gloo", timeout=kw.get("timeout"))
    if world == 1:
        dist.get_rank = lambda: 0
    bench_impl = importlib.import_module("badchimp_cpp_b200.bench_impl")
    W = importlib.import_module("badchimp_cpp_b200.workloads")
    for k, wl in W.WORKLOADS.items():
        wl["size"] = {"pack": 32, "dense": 16, "channel": 64}[wl["geometry"]]
    bench_impl.parity_probe = lambda *a, **kw: {"stub": True}
    bench_impl._clock_sampler = lambda stop, out: None
    bench_impl._ncu_traffic = lambda args: None
    bench_impl.cpu_baseline_port = lambda pkg, **kw: {"stub": True}
    orig_build = W.build
    def build(*a, **kw):
        rl = orig_build(*a, **kw); FakeLib.n_rows = rl.n + 1; return rl
    W.build = build
    class Args: pass
    args = Args()
    args.workload = os.environ.get("WL", "std_case")
    args.scaling, args.size, args.index, args.halo, args.balance = None, 0, "compact", "peer", True
    args.interior_domains, args.no_parity, args.no_weak, args.no_traffic, args.no_cpu_baseline, args.no_extra_workloads = False, False, False, False, False, False
    args.steps, args.warmup, args.gpus, args.skip_mask = 20, 5, world, "auto"
    bench_impl.run_b200(args)

main()
