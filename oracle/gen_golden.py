"""oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Run in the build container only (needs
/root/reference): generates the golden vectors under tests/golden/ from the UNMODIFIED
reference -- geometry files by the reference's PythonScripts/vtklb.py, time stepping by
oracle/_ref/ref_driver (reference headers + thin harness, see oracle/Makefile).

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz and *.vtklb

Every golden holds the case definition (geo, lattice, periodic, parameters), the integer
tables the reference built (per rank) and raw f / rho / vel (/ cg) dumps after a few steps.
"""
import contextlib
import io
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("CHIMP_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "PythonScripts"))
sys.path.insert(0, HERE)
import importlib.util

_spec = importlib.util.spec_from_file_location("chimp_geometry", os.path.join(ROOT, "badchimp-cpp_b200", "geometry.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)
from recfile import read_rec  # noqa: E402
from vtklb import vtklb  # noqa: E402  (the reference's own writer)

DRIVER = os.path.join(HERE, "_ref", "ref_driver")
OUT = os.path.join(ROOT, "tests", "golden")


def run_reference(name, geo, lattice, periodic, case, steps, dump, args, attributes, keep_vtklb=False, checkpoint=False, vtk=False,
                  vtk_ascii=False, checkpoint_at=None, extra=()):
    d = tempfile.mkdtemp(prefix="golden_")
    os.makedirs(os.path.join(d, "out"))
    basis = lattice if lattice != "D3Q27" else G.BASIS["D3Q27"].astype(int)
    with contextlib.redirect_stdout(io.StringIO()):
        v = vtklb(np.asarray(geo, dtype=int), basis, periodic, "tmp", d + "/")
        for aname, aval in attributes.items():
            v.append_data_set(aname, aval)
    nranks = int(np.max(geo))
    cmd = [DRIVER, "--case", case, "--lattice", lattice, "--dir", d, "--out", os.path.join(d, "out"), "--nranks",
           str(nranks), "--steps", str(steps), "--dump", ",".join(str(s) for s in dump)] + [str(a) for a in args]
    if checkpoint:
        cmd += ["--checkpoint", os.path.join(OUT, name + ".ckpt")]
    if checkpoint_at is not None:   # the reference's own writeToFile() after that step; the run goes on
        cmd += ["--checkpoint", os.path.join(OUT, name + ".ckpt"), "--checkpoint-at", str(checkpoint_at)]
    cmd += list(extra)
    if vtk:  # the reference's own Output<LT>::write (io/Output.h, io/VTK.h) after the last step
        vdir = os.path.join(OUT, name + ".vtk")
        shutil.rmtree(vdir, ignore_errors=True)
        cmd += ["--vtk", vdir + "/"] + (["--vtk-ascii"] if vtk_ascii else [])
    subprocess.run(cmd, check=True, capture_output=True)
    gold = {"geo": np.asarray(geo, dtype=np.int32), "lattice": lattice, "periodic": periodic, "case": case,
            "steps": steps, "dump": np.array(dump), "args": np.array([str(a) for a in args]), "nranks": nranks}
    if any(str(e).startswith("--pressure-bnd") for e in extra):
        gold["extra"] = np.array([str(e) for e in extra])
    for aname, aval in attributes.items():
        gold["attr." + aname] = np.asarray(aval)
    for r in range(nranks):
        rec = read_rec(os.path.join(d, "out", "rank%d.rec" % r))
        for k, val in rec.items():
            gold["r%d.%s" % (r, k)] = val
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **gold)
    if keep_vtklb:
        for r in range(nranks):
            shutil.copy(os.path.join(d, "tmp%d.vtklb" % r), os.path.join(OUT, "%s.tmp%d.vtklb" % (name, r)))
    shutil.rmtree(d)
    print("golden", name, "ranks", nranks, "records", len(gold))


def one_phase_attributes(shape, seed):
    """tags as documented at std_one_phase/main.cpp:253-259"""
    nx, ny, nz = shape
    fluid = G.sphere_pack(shape, 2.6, 0.75, seed).astype(bool)
    fluid[:, :, 0] = False
    fluid[:, :, -1] = False
    phase = np.where(np.arange(nx)[:, None, None] < nx // 2, 1, 2) * np.ones(shape, dtype=int)
    tags = np.where(fluid, phase, 0)
    tags[:, :, 0] = 3
    tags[:, :, -1] = 3
    basis = G.BASIS["D3Q19"][:-1]

    def shifted(a, c):  # periodic in x,y; z never wraps onto fluid because both ends are ghosts
        return np.roll(a, shift=(-c[0], -c[1], -c[2]), axis=(0, 1, 2))

    low = tags & 3
    near_solid = np.zeros(shape, bool)
    near_press = np.zeros(shape, bool)
    near_other = np.zeros(shape, bool)
    for c in basis:
        nb = shifted(low, c)
        near_solid |= nb == 0
        near_press |= nb == 3
        near_other |= (nb == 1) | (nb == 2)
        near_other &= True
    other = np.zeros(shape, bool)
    for c in basis:
        nb = shifted(low, c)
        other |= fluid & (nb != low) & ((nb == 1) | (nb == 2))
    tags = tags + fluid * (4 * other + 8 * near_solid + 16 * near_press)
    geo = fluid.astype(int)
    interior = np.zeros(shape, dtype=int)
    interior[1:4, 2:6, 4:9] = 1
    interior[nx - 3:, :, 3:7] = 2
    interior *= fluid
    force = np.ones(shape, dtype=int)
    force[:, :, 1:3] = 0
    rng = np.random.default_rng(seed + 1)
    attrs = {"nodetags": tags.astype(int), "domains": phase * fluid, "force": force, "interior_domains": interior,
             "normal_x": rng.random(shape), "normal_y": rng.random(shape), "normal_z": rng.random(shape)}
    return geo, attrs


def vtk_goldens():
    """small cases whose VTK files (written by the reference's Output class) are committed next to the dumps"""
    shape = (6, 5, 7)
    pack = G.sphere_pack(shape, 1.8, 0.7, 3).astype(int)
    run_reference("vtk_std_d3q19_p2", G.z_slab_rank_map(pack, 2), "D3Q19", "xyz", "std_case", 3, [3],
                  ["--tau", 0.8, "--force", "1e-5,2e-6,-3e-6"], {"init_rho": np.ones(shape)}, vtk=True)
    chan = np.ones((5, 6), dtype=int)
    chan[:, 0] = chan[:, -1] = 0
    run_reference("vtk_ascii_d2q9_p1", chan, "D2Q9", "x", "std_case", 2, [2], ["--tau", 0.8, "--force", "1e-5,0,0"],
                  {"init_rho": np.ones(chan.shape)}, vtk=True, vtk_ascii=True)
    pack2d = G.sphere_pack((9, 8), 2.0, 0.7, 9).astype(int)
    x2 = np.arange(9)[:, None] * np.ones((9, 8))
    r2 = (x2 < 4).astype(float)
    run_reference("vtk_twophase_d2q9_p1", pack2d, "D2Q9", "xy", "twophase", 3, [3],
                  ["--tau2", "1.0,1.0", "--sigma", 0.02, "--beta", 0.9, "--momx", 2e-5, "--force", "0,0,0"],
                  {"rho0": r2, "rho1": 1.0 - r2, "wettability": 0.3 * (pack2d == 0), "source": np.zeros((9, 8), dtype=int)}, vtk=True)


def restart_and_forcing_goldens():
    """round 2: a restart file written by the reference in the middle of a run (f4), and the reference's own
    calcFluxForceCartDir / calcCapNumbForceCartDir (LBglobalforcing.h:8-98) evaluated at dump steps (f2)"""
    shape = (12, 10, 14)
    pack = G.sphere_pack(shape, 3.2, 0.62, 11).astype(int)
    ones = np.ones(shape)
    rho_init = 1.0 + 0.02 * np.sin(2 * np.pi * np.arange(shape[0]) / shape[0])[:, None, None] * ones
    # same case as std_d3q19_p1 (whose .vtklb file is committed): checkpoint after step 4, dumps at 4 and 10
    run_reference("restart_d3q19_p1", pack, "D3Q19", "xyz", "std_case", 10, [4, 10], ["--tau", 0.8, "--force", "1e-6,2e-7,-3e-7"],
                  {"init_rho": rho_init}, checkpoint_at=4, extra=["--no-tables", "--global-forcing", "--fixed-flux", "1e-5"])
    x = np.arange(shape[0])[:, None, None] * np.ones(shape)
    rho0 = (x < shape[0] / 2).astype(float)
    tp_attrs = {"rho0": rho0, "rho1": 1.0 - rho0, "wettability": 0.5 * (pack == 0), "source": np.zeros(shape, dtype=int)}
    tp_args = ["--tau2", "1.0,0.8", "--sigma", 0.01, "--beta", 1.0, "--momx", 1e-5, "--force", "0,1e-7,0"]
    run_reference("forcing_twophase_d3q19_p1", pack, "D3Q19", "xyz", "twophase", 5, [1, 3, 5], tp_args, tp_attrs,
                  extra=["--no-tables", "--no-f", "--global-forcing", "--fixed-flux", "2e-5", "--cap-numb", "1e-4,0.1666666666666666574,0.1"])


def library_bnd_goldens():
    """round 2: the reference's own PressureBnd / InletOutlet classes (LBpressurebnd.h:10-88; no caller in its mains)
    applied after the bounce back of every std_case step on every third fluid boundary node (ref_driver --pressure-bnd)"""
    shape = (12, 10, 14)
    pack = G.sphere_pack(shape, 3.2, 0.62, 11).astype(int)
    ones = np.ones(shape)
    F3 = "--force", "1e-6,2e-7,-3e-7"
    run_reference("pbnd_d3q19_p1", pack, "D3Q19", "xyz", "std_case", 6, [1, 2, 6], ["--tau", 0.8, *F3], {"init_rho": ones},
                  extra=["--pressure-bnd", "pressure", "--bnd-every", "3"])
    run_reference("inout_d3q19_p2", G.z_slab_rank_map(pack, 2), "D3Q19", "xyz", "std_case", 6, [1, 6], ["--tau", 0.8, *F3], {"init_rho": ones},
                  extra=["--pressure-bnd", "inletoutlet", "--bnd-every", "3", "--io-rho", "1.02", "--io-vel", "0.01,-0.005,0.002"])
    pack2 = G.sphere_pack((16, 12), 2.5, 0.7, 5).astype(int)
    run_reference("inout_d2q9_p1", pack2, "D2Q9", "xy", "std_case", 8, [1, 8], ["--tau", 0.7, "--force", "1e-6,-2e-6,0"],
                  {"init_rho": np.ones(pack2.shape)},
                  extra=["--pressure-bnd", "inletoutlet", "--bnd-every", "2", "--io-rho", "0.98", "--io-vel", "0.02,0.01,0"])


def main():
    os.makedirs(OUT, exist_ok=True)
    subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)
    if sys.argv[1:] == ["vtk"]:
        vtk_goldens()
        return
    if sys.argv[1:] == ["library-bnd"]:
        library_bnd_goldens()
        return
    if sys.argv[1:] == ["round2"]:
        restart_and_forcing_goldens()
        return
    vtk_goldens()
    restart_and_forcing_goldens()
    library_bnd_goldens()
    shape = (12, 10, 14)
    pack = G.sphere_pack(shape, 3.2, 0.62, 11).astype(int)
    ones = np.ones(shape)
    rho_init = 1.0 + 0.02 * np.sin(2 * np.pi * np.arange(shape[0]) / shape[0])[:, None, None] * ones
    F3 = "--force", "1e-6,2e-7,-3e-7"
    run_reference("std_d3q19_p1", pack, "D3Q19", "xyz", "std_case", 10, [0, 1, 2, 10], ["--tau", 0.8, *F3],
                  {"init_rho": rho_init}, keep_vtklb=True)
    run_reference("std_d3q19_p3", G.z_slab_rank_map(pack, 3), "D3Q19", "xyz", "std_case", 10, [1, 2, 10],
                  ["--tau", 0.8, *F3], {"init_rho": rho_init})
    box = pack.copy()
    box[0] = box[-1] = 0
    box[:, 0] = box[:, -1] = 0
    box[:, :, 0] = box[:, :, -1] = 0
    run_reference("std_d3q19_box_p2", G.z_slab_rank_map(box, 2), "D3Q19", "", "std_case", 6, [1, 6],
                  ["--tau", 0.9, *F3], {"init_rho": rho_init}, keep_vtklb=True)
    chan = np.ones((8, 12), dtype=int)
    chan[:, 0] = chan[:, -1] = 0
    run_reference("std_d2q9_channel", chan, "D2Q9", "x", "std_case", 20, [1, 2, 20], ["--tau", 0.8, "--force", "1e-6,0,0"],
                  {"init_rho": np.ones(chan.shape)}, checkpoint=True)
    pack2 = G.sphere_pack((16, 12), 2.5, 0.7, 5).astype(int)
    run_reference("std_d2q9_pack_p2", G.z_slab_rank_map(pack2, 2), "D2Q9", "xy", "std_case", 8, [1, 8],
                  ["--tau", 0.7, "--force", "1e-6,-2e-6,0"], {"init_rho": np.ones(pack2.shape)})
    run_reference("std_d3q27_p2", G.z_slab_rank_map(pack, 2), "D3Q27", "xyz", "std_case", 6, [1, 2, 6],
                  ["--tau", 0.8, *F3], {"init_rho": rho_init})
    run_reference("trt_d3q19_p1", pack, "D3Q19", "xyz", "std_case", 6, [1, 6], ["--trt", "0.8,1.125", *F3],
                  {"init_rho": rho_init})
    geo1, attrs1 = one_phase_attributes((10, 8, 14), 21)
    run_reference("onephase_d3q19_p1", geo1, "D3Q19", "xy", "one_phase", 8, [1, 2, 8],
                  ["--tau", 0.8, "--force", "0,0,1e-5", "--rhow", 1.0], attrs1)
    run_reference("onephase_trt_d3q19_p2", G.z_slab_rank_map(geo1, 2), "D3Q19", "xy", "one_phase", 6, [1, 6],
                  ["--trt", "0.8,1.125", "--force", "0,0,1e-5", "--rhow", 1.0], attrs1)
    x = np.arange(shape[0])[:, None, None] * np.ones(shape)
    rho0 = (x < shape[0] / 2).astype(float)
    tp_attrs = {"rho0": rho0, "rho1": 1.0 - rho0, "wettability": 0.5 * (pack == 0), "source": np.zeros(shape, dtype=int)}
    tp_args = ["--tau2", "1.0,0.8", "--sigma", 0.01, "--beta", 1.0, "--momx", 1e-5, "--force", "0,1e-7,0"]
    run_reference("twophase_d3q19_p1", pack, "D3Q19", "xyz", "twophase", 8, [1, 2, 8], tp_args, tp_attrs)
    run_reference("twophase_d3q19_p2", G.z_slab_rank_map(pack, 2), "D3Q19", "xyz", "twophase", 6, [1, 6], tp_args, tp_attrs)
    pack2d = G.sphere_pack((14, 12), 2.5, 0.7, 9).astype(int)
    x2 = np.arange(14)[:, None] * np.ones((14, 12))
    r2 = (x2 < 7).astype(float)
    run_reference("twophase_d2q9_p1", pack2d, "D2Q9", "xy", "twophase", 6, [1, 6],
                  ["--tau2", "1.0,1.0", "--sigma", 0.02, "--beta", 0.9, "--momx", 2e-5, "--force", "0,0,0"],
                  {"rho0": r2, "rho1": 1.0 - r2, "wettability": 0.3 * (pack2d == 0), "source": np.zeros((14, 12), dtype=int)})


if __name__ == "__main__":
    main()
