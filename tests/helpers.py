"""Shared helpers of the test-suite: load golden vectors (tests/golden, generated from the
unmodified reference by oracle/gen_golden.py), build the product's tables for the same case
and wrap the oracle port (oracle/port.py).  Test infrastructure only."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG_DIR = os.path.join(ROOT, "badchimp-cpp_b200")


def load_package():
    """the package directory name contains a hyphen, so it is loaded by path"""
    name = "badchimp_cpp_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def oracle_port():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import port  # noqa
    return port


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.geo = self.z["geo"]
        self.lattice = str(self.z["lattice"])
        self.periodic = str(self.z["periodic"])
        self.case = str(self.z["case"])
        self.nranks = int(self.z["nranks"])
        self.dump = [int(s) for s in self.z["dump"]]
        self.args = self._parse_args([str(a) for a in self.z["args"]])

    @staticmethod
    def _parse_args(args):
        out = {}
        i = 0
        while i < len(args):
            key = args[i].lstrip("-")
            val = args[i + 1]
            out[key] = [float(x) for x in val.split(",")] if "," in val else float(val)
            i += 2
        return out

    def rec(self, rank, key):
        return self.z["r%d.%s" % (rank, key)]

    def has(self, rank, key):
        return ("r%d.%s" % (rank, key)) in self.z.files

    def attr(self, name):
        return self.z["attr." + name]

    def f(self, rank, step, n_fields=1):
        nq = {"D2Q9": 9, "D3Q19": 19, "D3Q27": 27}[self.lattice]
        return self.rec(rank, "step%d.f" % step).reshape(-1, n_fields, nq)

    def force(self):
        F = self.args.get("force", [0.0, 0.0, 0.0])
        return list(F) + [0.0] * (3 - len(F))


def all_golden_names():
    """goldens that carry the reference's integer tables and field dumps (the round-2 restart_* / forcing_* goldens hold
    field dumps or function values only and have their own tests, tests/test_restart_and_forcing.py; the pbnd_* / inout_*
    goldens run the reference's library pressure boundaries on top of std_case, tests/test_library_bnd.py)"""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith(("restart_", "forcing_", "pbnd_", "inout_")))


def build_tables(g: Golden):
    pkg = load_package()
    lg = pkg.geometry.LatticeGeometry(g.geo, g.lattice, g.periodic)
    return lg, lg.all_ranks()


def exchange_lists(tabs):
    """per rank: list over neighbours (ascending rank) of the MonLatMpi lists"""
    out = []
    for t in tabs:
        ss = t.send_side(tabs)
        out.append([dict(rank=nr, send_nodes=ss[k][0], send_ndir=ss[k][1], send_dirs=ss[k][2],
                         recv_nodes=t.recv_nodes[k], recv_ndir=t.recv_ndir[k], recv_dirs=t.recv_dirs[k])
                    for k, nr in enumerate(t.neig_ranks)])
    return out


def one_phase_setup(g: Golden, lg, tabs):
    """host-side setup of std_one_phase/main.cpp:253-345,452-458 from the attribute arrays"""
    pkg = load_package()
    return pkg.cases.one_phase_setup(lg, tabs, {k: g.attr(k) for k in ("nodetags", "force", "interior_domains")})
