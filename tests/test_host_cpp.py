"""The C++ host mirror of the reference's API (badchimp-cpp_b200/host/chimp/*.h).
CPU part: its LBvtk / Grid / Nodes / HalfWayBounceBack / BndMpi (message-free handshake) reproduce the
reference's integer tables for the golden .vtklb files written by the reference's own vtklb.py.
GPU part: the std_case application built on it (host/apps/std_case.cpp, the reference main's
structure with the node loop replaced by the engine call) reproduces the reference's populations
bit for bit, single rank and two ranks."""
import os
import struct
import subprocess

import numpy as np
import pytest

import helpers

HOST = os.path.join(helpers.PKG_DIR, "host")


def build(app, link_engine):
    out = os.path.join(HOST, "apps", app)
    src = out + ".cpp"
    deps = [src] + [os.path.join(HOST, "chimp", f) for f in os.listdir(os.path.join(HOST, "chimp"))]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-o", out, src]
    if link_engine:
        cmd += ["-I/usr/local/cuda/include", "-L" + helpers.PKG_DIR, "-lchimp_b200", "-L/usr/local/cuda/lib64", "-lcudart",
                "-Wl,-rpath," + helpers.PKG_DIR]
    subprocess.run(cmd, check=True)
    return out


def parse_dump(text):
    out = {}
    for line in text.splitlines():
        parts = line.split()
        out[parts[0]] = np.array([int(x) for x in parts[2:]], dtype=np.int64)
        assert len(out[parts[0]]) == int(parts[1])
    return out


@pytest.mark.parametrize("name", ["std_d3q19_p1", "std_d3q19_box_p2"])
def test_host_mirror_tables_match_reference(name):
    exe = build("dump_tables", link_engine=False)
    g = helpers.Golden(name)
    prefix = os.path.join(helpers.GOLDEN, name + ".tmp")
    for r in range(g.nranks):
        res = subprocess.run([exe, g.lattice, prefix, str(r)], capture_output=True, text=True, check=True)
        d = parse_dump(res.stdout)
        for key in ("neigh", "type", "rank", "bulk", "fluidBnd", "solidBnd", "bb.node", "bb.nBeta", "bb.nGamma",
                    "bb.nDelta", "bb.links"):
            assert np.array_equal(d[key], g.rec(r, key)), key
        for k in range(int(g.rec(r, "nNeigRanks")[0])):
            for sub in ("neigRank", "nodesToSend", "nDirPerNodeToSend", "dirListToSend", "nodesReceived",
                        "nDirPerNodeReceived", "dirListReceived"):
                key = "mpi%d.%s" % (k, sub)
                assert np.array_equal(d[key], g.rec(r, key)), key


def read_app_output(path, nq, nd):
    out = []
    with open(path, "rb") as fh:
        data = fh.read()
    p = 0
    while p < len(data):
        (sz,) = struct.unpack_from("<i", data, p)
        p += 4
        f = np.frombuffer(data, dtype="<f8", count=sz * nq, offset=p).reshape(sz, nq)
        p += 8 * sz * nq
        rho = np.frombuffer(data, dtype="<f8", count=sz, offset=p)
        p += 8 * sz
        vel = np.frombuffer(data, dtype="<f8", count=sz * nd, offset=p).reshape(sz, nd)
        p += 8 * sz * nd
        out.append((f, rho, vel))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["std_d3q19_p1", "std_d3q19_box_p2"])
def test_std_case_app_matches_reference(name, tmp_path):
    exe = build("std_case", link_engine=True)
    g = helpers.Golden(name)
    _, tabs = helpers.build_tables(g)
    step = max(g.dump)
    F = g.force()
    deck = tmp_path / "input.dat"
    deck.write_text("<iterations>\n  max %d\n  write %d\n<end>\n<fluid>\n  tau %r\n  bodyforce %r %r %r\n<end>\n"
                    % (step, step, g.args["tau"], F[0], F[1], F[2]))
    out = tmp_path / "out.bin"
    prefix = os.path.join(helpers.GOLDEN, name + ".tmp")
    subprocess.run([exe, g.lattice, str(deck), prefix, "0", str(out), str(g.nranks)], check=True)
    res = read_app_output(out, 19, 3)
    assert len(res) == g.nranks
    for r, (f, rho, vel) in enumerate(res):
        bulk = tabs[r].bulk_nodes()
        assert np.array_equal(f[bulk], g.f(r, step)[bulk, 0])
        assert np.array_equal(rho[bulk], g.rec(r, "step%d.rho" % step)[bulk])
        assert np.array_equal(vel[bulk], g.rec(r, "step%d.vel" % step).reshape(-1, 3)[bulk])
