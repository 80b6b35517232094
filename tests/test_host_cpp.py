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
                "-Wl,-rpath," + helpers.PKG_DIR, "-pthread"]
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


def write_case_files(g, tabs, tmp_path, attributes):
    """geometry files for the reference-format reader, written by the product's byte-identical .vtklb writer"""
    for r, t in enumerate(tabs):
        t.write_vtklb(str(tmp_path / ("tmp%d.vtklb" % r)), attributes)
    return str(tmp_path / "tmp")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["onephase_d3q19_p1", "onephase_trt_d3q19_p2"])
def test_std_one_phase_app_matches_reference(name, tmp_path):
    """host/apps/std_one_phase.cpp (structure of std_one_phase/main.cpp on the C++ mirror + engine) against the
    reference's loop; tolerances as in test_one_phase_vs_reference (tree-summed mass change)"""
    exe = build("std_one_phase", link_engine=True)
    g = helpers.Golden(name)
    pkg = helpers.load_package()
    lg, tabs = helpers.build_tables(g)
    names = ("nodetags", "domains", "force", "interior_domains", "normal_x", "normal_y", "normal_z")
    prefix = write_case_files(g, tabs, tmp_path, {k: g.attr(k) for k in names})
    step = max(g.dump)
    F = g.force()
    coll = ("tausym %r\n  tauanti %r" % tuple(g.args["trt"])) if "trt" in g.args else ("tau %r" % g.args["tau"])
    deck = tmp_path / "input.dat"
    # the reference loop runs i = 0..max: max = step - 1 gives `step` iterations; one write interval
    deck.write_text("<iterations>\n  max %d\n  write %d\n<end>\n<fluid>\n  %s\n  bodyforce %r %r %r\n  rhow %r\n<end>\n"
                    % (step - 1, step - 1, coll, F[0], F[1], F[2], g.args.get("rhow", 1.0)))
    out = tmp_path / "out.bin"
    subprocess.run([exe, str(deck), prefix, "0", str(out), str(g.nranks)], check=True)
    data = open(out, "rb").read()
    p = 0
    setup = helpers.one_phase_setup(g, lg, tabs)
    flux = np.zeros(2)
    for r, t in enumerate(tabs):
        (sz,) = struct.unpack_from("<i", data, p); p += 4
        assert sz == t.size
        f = np.frombuffer(data, "<f8", sz * 19, p).reshape(sz, 19); p += 8 * sz * 19
        rho = np.frombuffer(data, "<f8", sz, p); p += 8 * sz
        vel = np.frombuffer(data, "<f8", sz * 3, p).reshape(sz, 3); p += 8 * sz * 3
        (nl,) = struct.unpack_from("<i", data, p); p += 4
        mass = np.frombuffer(data, "<f8", nl, p); p += 8 * nl
        bulk = t.bulk_nodes()
        assert np.allclose(f[bulk], g.f(r, step)[bulk, 0], rtol=1e-12, atol=0)
        assert np.allclose(rho[bulk], g.rec(r, "step%d.rho" % step)[bulk], rtol=1e-12, atol=0)
        assert np.allclose(mass, g.rec(r, "step%d.massChange" % step), rtol=1e-9, atol=1e-16)
        gv = g.rec(r, "step%d.vel" % step).reshape(sz, 3)
        gr = g.rec(r, "step%d.rho" % step)
        for n, ph in zip(setup[r]["press_nodes"], setup[r]["press_phase"]):
            flux[ph] += gv[n, 2] * gr[n]
    assert p == len(data)
    # the .flux file of main.cpp:620-630: the last block holds the fluxes after `step` iterations
    lines = open(str(out) + ".flux").read().splitlines()
    assert lines[-3] == "PLOT AT ITERATION: %d" % (step - 1)
    q1 = float(lines[-2].split()[2]); q2 = float(lines[-1].split()[2])
    assert np.allclose([q1, q2], 0.5 * flux, rtol=1e-5, atol=1e-12)   # printed with 6 significant digits


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["twophase_d3q19_p1", "twophase_d3q19_p2", "twophase_d2q9_p1"])
def test_twophase_app_matches_reference(name, tmp_path):
    """host/apps/twophase.cpp against the reference's colour-gradient loop, one and two ranks (in-process
    ranks: scalar halo of phi, all-reduced flux force, population halos of both fields)"""
    exe = build("twophase", link_engine=True)
    g = helpers.Golden(name)
    lg, tabs = helpers.build_tables(g)
    prefix = write_case_files(g, tabs, tmp_path, {k: g.attr(k) for k in ("rho0", "rho1", "wettability", "source")})
    step = max(g.dump)
    a, F = g.args, g.force()
    nd = 2 if g.lattice == "D2Q9" else 3
    nq = 9 if g.lattice == "D2Q9" else 19
    deck = tmp_path / "input.dat"
    deck.write_text("outdir test\n<iterations>\n  max %d\n  write %d\n<end>\n<fluid>\n  tau %r %r\n  sigma %r\n  beta %r\n  momx %r\n"
                    "  bodyforce %s\n<end>\n" % (step - 1, step - 1, a["tau2"][0], a["tau2"][1], a["sigma"], a["beta"], a["momx"],
                                                " ".join(repr(x) for x in F[:nd])))
    out = tmp_path / "out.bin"
    subprocess.run([exe, g.lattice, str(deck), prefix, "0", str(out), str(g.nranks)], check=True)
    data = open(out, "rb").read()
    p = 0
    for r, t in enumerate(tabs):
        (sz,) = struct.unpack_from("<i", data, p); p += 4
        assert sz == t.size
        f = np.frombuffer(data, "<f8", sz * 2 * nq, p).reshape(sz, 2, nq); p += 8 * sz * 2 * nq
        rho = np.frombuffer(data, "<f8", sz * 2, p).reshape(sz, 2); p += 8 * sz * 2
        vel = np.frombuffer(data, "<f8", sz * nd, p).reshape(sz, nd); p += 8 * sz * nd
        cg = np.frombuffer(data, "<f8", sz, p); p += 8 * sz
        bulk = t.bulk_nodes()
        assert np.allclose(f[bulk], g.f(r, step, 2)[bulk], rtol=1e-12, atol=1e-300)
        assert np.allclose(rho[bulk], g.rec(r, "step%d.rho" % step).reshape(-1, 2)[bulk], rtol=1e-12, atol=0)
        assert np.allclose(cg[bulk], g.rec(r, "step%d.cg" % step)[bulk], rtol=1e-10, atol=1e-14)
        assert np.allclose(vel[bulk], g.rec(r, "step%d.vel" % step).reshape(sz, -1)[bulk], rtol=1e-9, atol=1e-16)
    assert p == len(data)
    last = open(str(out) + ".force.dat").read().splitlines()[-1].split()
    fx = float(g.rec(0, "step%d.forceX" % step)[0])
    assert int(last[0]) == step - 1 and abs(float(last[1]) - fx) <= 1e-10 * abs(fx)


def _without_timestamp(data):
    return b"\n".join(l for l in data.split(b"\n") if not l.startswith(b"<!-- Created"))


def compare_vtk_tree(gold, mine):
    n = 0
    for root, _, files in os.walk(gold):
        for f in files:
            rel = os.path.relpath(os.path.join(root, f), gold)
            a = open(os.path.join(gold, rel), "rb").read()
            b = open(os.path.join(mine, rel), "rb").read()
            if rel.endswith(".pvtu"):   # the header carries the wall-clock time of the run
                a, b = _without_timestamp(a), _without_timestamp(b)
            assert a == b, rel
            n += 1
    return n


@pytest.mark.parametrize("name,nrho,fname,geo", [("vtk_std_d3q19_p2", 1, "lb_run", ""), ("vtk_twophase_d2q9_p1", 2, "fluid", "geo"),
                                                  ("vtk_ascii_d2q9_p1", 1, "lb_run", "ascii")])
def test_vtk_output_is_byte_identical_to_reference(name, nrho, fname, geo, tmp_path):
    """Output<LT> of the host mirror writes the files the reference's Output/VTK classes write
    (goldens produced by the reference's own io/Output.h through oracle/ref_driver --vtk): .vtu pieces with the
    voxel / pixel mesh and raw appended arrays, .pvtu index with 100-byte piece records, 2 ranks and 2-D padding;
    the third case is the ASCII format (Output<LT, double, VTK::ASCII>)"""
    exe = build("vtk_write", link_engine=False)
    g = helpers.Golden(name)
    lg, tabs = helpers.build_tables(g)
    prefix = write_case_files(g, tabs, tmp_path, {k[5:]: g.z[k] for k in g.z.files if k.startswith("attr.")})
    step = max(g.dump)
    for r in range(g.nranks):
        fb = tmp_path / ("fields%d.bin" % r)
        fb.write_bytes(g.rec(r, "step%d.rho" % step).tobytes() + g.rec(r, "step%d.vel" % step).tobytes())
        subprocess.run([exe, g.lattice, prefix, str(r), str(g.nranks), str(fb), str(nrho), str(tmp_path / "out"), fname, str(step)]
                       + ([geo] if geo else []), check=True)
    assert compare_vtk_tree(os.path.join(helpers.GOLDEN, name + ".vtk"), str(tmp_path / "out")) >= 2


@pytest.mark.gpu
def test_std_case_app_writes_reference_vtk_files(tmp_path):
    """engine -> download -> Output: the std_case application's VTK files equal the reference's byte for byte
    (its populations and moments are bit-exact)"""
    exe = build("std_case", link_engine=True)
    g = helpers.Golden("vtk_std_d3q19_p2")
    lg, tabs = helpers.build_tables(g)
    prefix = write_case_files(g, tabs, tmp_path, {"init_rho": g.attr("init_rho")})
    step = max(g.dump)
    F = g.force()
    deck = tmp_path / "input.dat"
    deck.write_text("<iterations>\n  max %d\n  write %d\n<end>\n<fluid>\n  tau %r\n  bodyforce %r %r %r\n<end>\n"
                    % (step, step, g.args["tau"], F[0], F[1], F[2]))
    subprocess.run([exe, g.lattice, str(deck), prefix, "0", str(tmp_path / "out.bin"), str(g.nranks), str(tmp_path / "vtk")], check=True)
    assert compare_vtk_tree(os.path.join(helpers.GOLDEN, g.name + ".vtk"), str(tmp_path / "vtk")) == 3


def _read_cpu_loop(path, n_fields, nq, nd):
    data = open(path, "rb").read()
    (sz,) = struct.unpack_from("<i", data, 0)
    p = 4
    f = np.frombuffer(data, dtype="<f8", count=sz * n_fields * nq, offset=p).reshape(sz, n_fields, nq)
    p += 8 * sz * n_fields * nq
    rho = np.frombuffer(data, dtype="<f8", count=sz * n_fields, offset=p).reshape(sz, n_fields)
    p += 8 * sz * n_fields
    vel = np.frombuffer(data, dtype="<f8", count=sz * nd, offset=p).reshape(sz, nd)
    return f, rho, vel


@pytest.mark.parametrize("name", ["std_d3q19_p1", "trt_d3q19_p1", "std_d2q9_channel", "twophase_d3q19_p1", "twophase_d2q9_p1"])
def test_host_mirror_per_node_functions_reproduce_the_reference(name, tmp_path):
    """calcRho / calcVel / calcOmegaBGK[TRT] / calcDeltaOmegaF[TRT] / calcDeltaOmegaST / calcDeltaOmegaRC / grad /
    vecNorm / initiateLbField / propagateTo / swapData / HalfWayBounceBack::apply of the host mirror
    (host/chimp/LBcollision.h, LBfield.h, LBhalfwaybb.h), driven by the reference's loop bodies in
    host/apps/cpu_loop.cpp on the CPU: populations, rho and u equal the reference's own dumps bit for bit"""
    exe = os.path.join(HOST, "apps", "cpu_loop")
    src = exe + ".cpp"
    deps = [src] + [os.path.join(HOST, "chimp", f) for f in os.listdir(os.path.join(HOST, "chimp"))]
    if not os.path.exists(exe) or any(os.path.getmtime(exe) < os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-o", exe, src], check=True)
    g = helpers.Golden(name)
    lg, tabs = helpers.build_tables(g)
    t = tabs[0]
    attrs = {k[5:]: g.z[k] for k in g.z.files if k.startswith("attr.")}
    t.write_vtklb(str(tmp_path / "tmp0.vtklb"), attrs)
    step = max(g.dump)
    a = g.args
    F = g.force()[: lg.nd]
    out = str(tmp_path / "out.bin")
    if g.case == "twophase":
        kind, nf = "twophase", 2
        params = [repr(a["tau2"][0]), repr(a["tau2"][1]), repr(a["sigma"]), repr(a["beta"]), repr(a["momx"])] + [repr(x) for x in F[1:]]
    elif "trt" in a:
        kind, nf = "trt", 1
        params = [repr(a["trt"][0]), repr(a["trt"][1])] + [repr(x) for x in F]
    else:
        kind, nf = "std", 1
        params = [repr(a.get("tau", 0.8))] + [repr(x) for x in F]
    subprocess.run([exe, kind, g.lattice, str(tmp_path / "tmp"), out, str(step)] + params, check=True, capture_output=True)
    f, rho, vel = _read_cpu_loop(out, nf, lg.nq, lg.nd)
    bulk = t.bulk_nodes()
    assert np.array_equal(f[bulk], g.f(0, step, nf)[bulk])
    assert np.array_equal(rho[bulk], g.rec(0, "step%d.rho" % step).reshape(-1, nf)[bulk])
    assert np.array_equal(vel[bulk], g.rec(0, "step%d.vel" % step).reshape(-1, lg.nd)[bulk])


@pytest.mark.parametrize("lattice", ["D2Q9", "D3Q19"])
def test_host_mirror_mass_source_terms_and_equilibrium(lattice, tmp_path):
    """calcDeltaOmegaQ / calcDeltaOmegaQTRT (LBcollision.h:79-121) and calcfeq (LButilities.h:74-91) of the host mirror
    against the same expressions evaluated term by term in IEEE doubles"""
    exe = os.path.join(HOST, "apps", "cpu_loop")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-o", exe, exe + ".cpp"], check=True)
    pkg = helpers.load_package()
    basis = pkg.geometry.BASIS[lattice]
    w = pkg.cases.lattice_weights(lattice)
    tau, ts, ta, src, rho = 0.83, 0.77, 1.19, 3.1e-4, 1.0123
    u = [0.0123, -0.0231, 0.0072][: basis.shape[1]]
    r = subprocess.run([exe, "unitq", lattice, repr(tau), repr(ts), repr(ta), repr(src), repr(rho)] + [repr(x) for x in u],
                       capture_output=True, text=True, check=True)
    got = np.array([[float(x) for x in line.split()] for line in r.stdout.strip().splitlines()])
    u2 = u[0] * u[0] + u[1] * u[1]
    if len(u) == 3:
        u2 = u2 + u[2] * u[2]
    c2 = 1.0 / 3.0
    for q in range(len(basis)):
        cu, first = 0.0, True
        for d in range(len(u)):
            c = int(basis[q, d])
            if c:
                term = u[d] if c > 0 else -u[d]
                cu = term if first else cu + term
                first = False
        e = (1.0 + 3.0 * cu + 4.5 * (cu * cu - c2 * u2))
        assert got[q, 0] == (1 - 0.5 / tau) * src * w[q] * e
        assert got[q, 1] == src * w[q] * ((1 - 0.5 / ta) * 3.0 * cu + (1 - 0.5 / ts) * (1.0 + 4.5 * (cu * cu - c2 * u2)))
        assert got[q, 2] == rho * w[q] * e
