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
