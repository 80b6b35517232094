"""The C-ABI library loads without a GPU and exports every symbol include/chimp_b200.h
declares; lattice constants match the reference's structs; compute entry points fail loudly
when no CUDA device is present (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers


def header_symbols():
    src = open(os.path.join(helpers.ROOT, "include", "chimp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(chimp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    pkg = helpers.load_package()
    lib = pkg.capi.lib()
    declared = header_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert sorted(pkg.capi.SYMBOLS) == declared


def test_lattice_constants_match_reference_structs():
    pkg = helpers.load_package()
    lib = pkg.capi.lib()
    for name, lid in pkg.geometry.LATTICE_ID.items():
        basis = pkg.geometry.BASIS[name]
        nq = lib.chimp_lattice_nq(lid)
        assert nq == len(basis) and lib.chimp_lattice_nd(lid) == basis.shape[1]
        w = pkg.cases.lattice_weights(name)
        for q in range(nq):
            assert [lib.chimp_lattice_c(lid, q, d) for d in range(basis.shape[1])] == basis[q].tolist()
            assert lib.chimp_lattice_w(lid, q) == w[q]
            r = lib.chimp_lattice_reverse(lid, q)
            assert np.array_equal(basis[r], -basis[q])
        assert abs(w.sum() - 1.0) < 1e-15


def test_reference_header_values():
    """literal values of LBd3q19.h:25-40 / LBd2q9.h:26-39"""
    pkg = helpers.load_package()
    w19 = pkg.cases.lattice_weights("D3Q19")
    assert w19[0] == 2.0 / 36.0 and w19[3] == 1.0 / 36.0 and w19[18] == 12.0 / 36.0
    w9 = pkg.cases.lattice_weights("D2Q9")
    assert w9[0] == 4.0 / 36.0 and w9[1] == 1.0 / 36.0 and w9[8] == 16.0 / 36.0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pkg = helpers.load_package()
    geo = np.ones((4, 4, 4), dtype=int)
    t = pkg.geometry.LatticeGeometry(geo, "D3Q19", "xyz").all_ranks()[0]
    lat = pkg.capi.Lattice.from_rank_tables(t)
    with pytest.raises(pkg.capi.ChimpError, match="no CUDA device"):
        lat.finalize()
