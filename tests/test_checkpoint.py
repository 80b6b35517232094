"""Restart files in the reference's raw formats (.lblbf / .lbsca / .lbvec, LBfield.h:102-138,233-276,
378-421): files written by the reference's own writeToFile() (tests/golden/std_d2q9_channel.ckpt0.*,
produced through oracle/_ref/ref_driver --checkpoint) are read back to the golden fields, and writing
the same fields reproduces the reference's files byte for byte."""
import os

import numpy as np

import helpers


def test_reference_checkpoint_files_round_trip(tmp_path):
    pkg = helpers.load_package()
    ck = pkg.checkpoint
    g = helpers.Golden("std_d2q9_channel")
    prefix = os.path.join(helpers.GOLDEN, "std_d2q9_channel.ckpt0")
    step = max(g.dump)
    f = ck.read_lbfield(prefix)
    rho = ck.read_scalar_field(prefix)
    vel = ck.read_vector_field(prefix)
    assert np.array_equal(f, g.f(0, step))
    assert np.array_equal(rho[:, 0], g.rec(0, "step%d.rho" % step))
    assert np.array_equal(vel[:, 0, :], g.rec(0, "step%d.vel" % step).reshape(-1, 2))
    out = str(tmp_path / "mine")
    ck.write_lbfield(out, f)
    ck.write_scalar_field(out, rho)
    ck.write_vector_field(out, vel)
    for ext in (".lblbf", ".lbsca", ".lbvec"):
        assert open(out + ext, "rb").read() == open(prefix + ext, "rb").read(), ext


def test_readers_refuse_truncated_files_and_foreign_headers(tmp_path):
    """like the reference ("No data read!", LBfield.h:124-131,260-268,406-413): a header that does not match the
    caller's field, or a file shorter than its header says, is an error -- never a silent reshape"""
    import pytest
    pkg = helpers.load_package()
    ck = pkg.checkpoint
    prefix = os.path.join(helpers.GOLDEN, "std_d2q9_channel.ckpt0")
    f = ck.read_lbfield(prefix)
    n_nodes, n_fields, nq = f.shape
    assert ck.read_lbfield(prefix, expect=(n_fields, nq, n_nodes)).shape == f.shape
    with pytest.raises(ValueError, match="no data read"):
        ck.read_lbfield(prefix, expect=(n_fields, 19, n_nodes))
    with pytest.raises(ValueError, match="no data read"):
        ck.read_scalar_field(prefix, expect=(2, n_nodes))
    with pytest.raises(ValueError, match="no data read"):
        ck.read_vector_field(prefix, expect=(1, 3, n_nodes))
    for ext, reader in ((".lblbf", ck.read_lbfield), (".lbsca", ck.read_scalar_field), (".lbvec", ck.read_vector_field)):
        raw = open(prefix + ext, "rb").read()
        cut = str(tmp_path / "cut")
        open(cut + ext, "wb").write(raw[:-16])
        with pytest.raises(ValueError, match="truncated"):
            reader(cut)
        open(cut + ext, "wb").write(raw[:6])
        with pytest.raises(ValueError, match="bad header"):
            reader(cut)
