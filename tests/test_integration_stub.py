"""INTEGRATION.md's claim, compiled and run: the binding stub (include/reference_binding/LBgpu.h -- the code block of
INTEGRATION.md section 1) over the reference's own Grid / Nodes / LbField / HalfWayBounceBack classes, and the
reference's std_case main with its node loop switched to the engine (oracle/integration_std_case.cpp), built against
the unmodified headers under /root/reference/src.  The binary travels to the GPU box (oracle/_ref/), runs the golden
case std_d3q19_p1 there through the C-ABI, and must reproduce the dump of the reference's own CPU loop bit for bit."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

import helpers

STUB = os.path.join(helpers.ROOT, "include", "reference_binding", "LBgpu.h")
BINARY = os.path.join(helpers.ROOT, "oracle", "_ref", "integration_std_case")
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
    assert r.returncode == 0 and os.path.exists(BINARY), r.stdout + r.stderr
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
