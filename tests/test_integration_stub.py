"""INTEGRATION.md's claim, compiled and run: the binding stub (include/reference_binding/LBgpu.h -- the code block of
INTEGRATION.md section 1) over the reference's own Grid / Nodes / LbField / HalfWayBounceBack classes, and the
reference's three target mains with their node loops switched to the engine (oracle/integration_std_case.cpp,
integration_one_phase.cpp, integration_twophase.cpp), built against the unmodified headers under /root/reference/src.
The binaries travel to the GPU box (oracle/_ref/), run the golden cases there through the C-ABI, and must reproduce the
dumps of the reference's own CPU loops (bit for bit for std_case; within the tolerances of the global sums otherwise)."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

import helpers

STUB = os.path.join(helpers.ROOT, "include", "reference_binding", "LBgpu.h")
BINARY = os.path.join(helpers.ROOT, "oracle", "_ref", "integration_std_case")
BINARY_ONE_PHASE = os.path.join(helpers.ROOT, "oracle", "_ref", "integration_one_phase")
BINARY_TWOPHASE = os.path.join(helpers.ROOT, "oracle", "_ref", "integration_twophase")
REFERENCE = "/root/reference/src/lbsolver"


def test_integration_md_shows_the_compiled_stub_verbatim():
    doc = open(os.path.join(helpers.ROOT, "INTEGRATION.md")).read()
    block = doc.split("```cpp\n", 1)[1].split("```", 1)[0]
    assert block == open(STUB).read()


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="needs the reference tree (build container only)")
def test_stub_and_switched_main_compile_against_the_reference_headers():
    lib = os.path.join(helpers.PKG_DIR, "libchimp_b200.so")
    if not os.path.exists(lib):
        pytest.skip("engine library not built yet")
    r = subprocess.run(["make", "-C", os.path.join(helpers.ROOT, "oracle"), "-B", "integration"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert all(os.path.exists(b) for b in (BINARY, BINARY_ONE_PHASE, BINARY_TWOPHASE)), r.stdout + r.stderr
    assert "error" not in (r.stdout + r.stderr).lower()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BINARY), reason="needs oracle/_ref/integration_std_case (built where the reference tree is)")
def test_reference_main_switched_to_the_engine_reproduces_the_reference_run(tmp_path):
    sys.path.insert(0, os.path.join(helpers.ROOT, "oracle"))
    from recfile import read_rec
    g = helpers.Golden("std_d3q19_p1")
    shutil.copy(os.path.join(helpers.GOLDEN, "std_d3q19_p1.tmp0.vtklb"), str(tmp_path / "tmp0.vtklb"))
    out = tmp_path / "out"
    os.makedirs(str(out))
    F = g.force()
    # 10 iterations in write intervals of 4 (4 + 4 + 2): the engine is entered three times
    # (relative paths, like the reference main's "./../output/": its PVTU writer keeps piece names in a 101-byte buffer)
    r = subprocess.run([BINARY, ".", "out", "10", "4", repr(g.args["tau"]), repr(F[0]), repr(F[1]), repr(F[2])],
                       capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("PLOT AT ITERATION") == 3
    rec = read_rec(str(out / "rank0.rec"))
    bulk = g.rec(0, "bulk")
    assert np.array_equal(rec["step10.f"].reshape(-1, 19)[bulk], g.f(0, 10)[bulk, 0])
    assert np.array_equal(rec["step10.rho"][bulk], g.rec(0, "step10.rho")[bulk])
    assert np.array_equal(rec["step10.vel"].reshape(-1, 3)[bulk], g.rec(0, "step10.vel").reshape(-1, 3)[bulk])


def _run_switched_main(binary, g, tmp_path, params):
    sys.path.insert(0, os.path.join(helpers.ROOT, "oracle"))
    from recfile import read_rec
    lg, tabs = helpers.build_tables(g)
    t = tabs[0]
    attrs = {k[5:]: g.z[k] for k in g.z.files if k.startswith("attr.")}
    t.write_vtklb(str(tmp_path / "tmp0.vtklb"), attrs)       # byte-identical to the reference's vtklb.py (tests/test_geometry.py)
    os.makedirs(str(tmp_path / "out"))
    r = subprocess.run([binary, ".", "out"] + params, capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    return t, read_rec(str(tmp_path / "out" / "rank0.rec"))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BINARY_TWOPHASE), reason="needs oracle/_ref/integration_twophase (built where the reference tree is)")
def test_reference_twophase_main_switched_to_the_engine(tmp_path):
    """main_TWOPHASE.cpp's set-up by the reference's own classes, loop on the GPU: populations <= 1e-12 relative (the
    flux controller's global sum is a tree sum on the device), phi <= 1e-10, flux force <= 1e-10"""
    g = helpers.Golden("twophase_d3q19_p1")
    a, F, step = g.args, g.force(), max(g.dump)
    t, rec = _run_switched_main(BINARY_TWOPHASE, g, tmp_path, [str(step), repr(a["tau2"][0]), repr(a["tau2"][1]), repr(a["sigma"]), repr(a["beta"]),
                                                                repr(a["momx"]), repr(F[1]), repr(F[2])])
    bulk = t.bulk_nodes()
    s = "step%d." % step
    assert np.allclose(rec[s + "f"].reshape(-1, 2, 19)[bulk], g.f(0, step, 2)[bulk], rtol=1e-12, atol=1e-300)
    assert np.allclose(rec[s + "rho"].reshape(-1, 2)[bulk], g.rec(0, s + "rho").reshape(-1, 2)[bulk], rtol=1e-12, atol=0)
    assert np.allclose(rec[s + "cg"][bulk], g.rec(0, s + "cg")[bulk], rtol=1e-10, atol=1e-14)
    assert np.allclose(rec[s + "vel"].reshape(-1, 3)[bulk], g.rec(0, s + "vel").reshape(-1, 3)[bulk], rtol=1e-9, atol=1e-16)
    fx = float(g.rec(0, s + "forceX")[0])
    assert abs(float(rec[s + "forceX"][0]) - fx) <= 1e-10 * abs(fx)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BINARY_ONE_PHASE), reason="needs oracle/_ref/integration_one_phase (built where the reference tree is)")
def test_reference_one_phase_main_switched_to_the_engine(tmp_path):
    """std_one_phase/main.cpp's set-up (tags, link finders, force switch, interior domains, mass-source scale) by the
    reference's own classes, loop on the GPU: populations <= 1e-12 relative (per-domain mass change is a tree sum on the
    device), mass change per domain, and the mass flux through the pressure nodes (main.cpp:606-618) from the fields"""
    g = helpers.Golden("onephase_d3q19_p1")
    a, F, step = g.args, g.force(), max(g.dump)
    t, rec = _run_switched_main(BINARY_ONE_PHASE, g, tmp_path, [str(step), "bgk", repr(a["tau"]), "0", repr(F[0]), repr(F[1]), repr(F[2]),
                                                                 repr(a.get("rhow", 1.0))])
    bulk = t.bulk_nodes()
    s = "step%d." % step
    assert np.allclose(rec[s + "f"].reshape(-1, 19)[bulk], g.f(0, step)[bulk, 0], rtol=1e-12, atol=0.0)
    ref_rho, ref_vel = g.rec(0, s + "rho"), g.rec(0, s + "vel").reshape(-1, 3)
    assert np.allclose(rec[s + "rho"][bulk], ref_rho[bulk], rtol=1e-12, atol=0)
    assert np.allclose(rec[s + "vel"].reshape(-1, 3)[bulk], ref_vel[bulk], rtol=1e-9, atol=1e-18)
    assert np.allclose(rec[s + "massChange"], g.rec(0, s + "massChange"), rtol=1e-9, atol=1e-16)
    # mass flux per phase over the pressure nodes in list order, from the reference's own rho and vel of that step
    setup = helpers.one_phase_setup(g, *helpers.build_tables(g))[0]
    want = np.zeros(2)
    for n, ph in zip(setup["press_nodes"], setup["press_phase"]):
        want[ph] += ref_vel[n, 2] * ref_rho[n]
    assert np.allclose(rec[s + "massFlux"], want, rtol=1e-9, atol=1e-18)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BINARY), reason="needs oracle/_ref/integration_std_case (built where the reference tree is)")
def test_reference_pressure_bnd_object_handed_to_the_engine(tmp_path):
    """the reference's OWN PressureBnd<D3Q19> object (LBpressurebnd.h:10-41; no main of the reference uses it) handed to
    the binding stub's GpuLattice::add in the switched std_case main: the dumps of the reference's CPU loop with
    pressureBnd.apply() after the bounce back (golden pbnd_d3q19_p1), bit for bit"""
    g = helpers.Golden("pbnd_d3q19_p1")
    F = g.force()
    t, rec = _run_switched_main(BINARY, g, tmp_path, ["6", "4", repr(g.args["tau"]), repr(F[0]), repr(F[1]), repr(F[2]), "pressure"])
    bulk = t.bulk_nodes()
    assert np.array_equal(rec["step6.f"].reshape(-1, 19)[bulk], g.f(0, 6)[bulk, 0])
    assert np.array_equal(rec["step6.rho"][bulk], g.rec(0, "step6.rho")[bulk])
