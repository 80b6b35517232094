"""Checkpoint / restart files in the reference's raw formats (src/lbsolver/LBfield.h:102-138,
233-276, 378-421): native-endian int32 header, then the field's AoS doubles.

    .lbsca   int nFields, int nNodes,           double[nFields*nNodes]          (ScalarField)
    .lbvec   int nFields, int nD, int nNodes,   double[nFields*nD*nNodes]       (VectorField)
    .lblbf   int nFields, int nQ, int nNodes,   double[nFields*nQ*nNodes]       (LbField)

Arrays are node-major [node, field, component] exactly as chimp_upload_lbfield / chimp_download_*
take and return them, so a state written by the CPU code restarts on the GPU and vice versa.

Readers take an optional `expect` tuple with the header the caller's lattice needs ((nFields, nQ, nNodes),
(nFields, nNodes), (nFields, nD, nNodes)); like the reference, which refuses a file whose header does not match the
field it is read into ("No data read!", LBfield.h:124-131, 260-268, 406-413), a mismatch is an error -- as is a
truncated file."""
from __future__ import annotations

import numpy as np


def write_lbfield(path_prefix, f):
    f = np.ascontiguousarray(f, dtype=np.float64)
    n_nodes, n_fields, nq = f.shape
    with open(path_prefix + ".lblbf", "wb") as fh:
        np.array([n_fields, nq, n_nodes], dtype=np.int32).tofile(fh)
        f.tofile(fh)


def _read(path, n_header, expect, what):
    with open(path, "rb") as fh:
        head = np.fromfile(fh, dtype=np.int32, count=n_header)
        if head.size != n_header or (head <= 0).any():
            raise ValueError("%s file %s: bad header %s" % (what, path, head.tolist()))
        header = tuple(int(x) for x in head)
        if expect is not None and header != tuple(int(x) for x in expect):
            raise ValueError("%s file %s holds %s, the field needs %s: no data read" % (what, path, header, tuple(expect)))
        count = int(np.prod([int(x) for x in head], dtype=np.int64))
        data = np.fromfile(fh, dtype=np.float64, count=count)
    if data.size != count:
        raise ValueError("truncated %s file %s: %d of %d values" % (what, path, data.size, count))
    return header, data


def read_lbfield(path_prefix, expect=None):
    (n_fields, nq, n_nodes), data = _read(path_prefix + ".lblbf", 3, expect, "LbField")
    return data.reshape(n_nodes, n_fields, nq)


def write_scalar_field(path_prefix, s):
    s = np.ascontiguousarray(s, dtype=np.float64)
    if s.ndim == 1:
        s = s[:, None]
    n_nodes, n_fields = s.shape
    with open(path_prefix + ".lbsca", "wb") as fh:
        np.array([n_fields, n_nodes], dtype=np.int32).tofile(fh)
        s.tofile(fh)


def read_scalar_field(path_prefix, expect=None):
    (n_fields, n_nodes), data = _read(path_prefix + ".lbsca", 2, expect, "ScalarField")
    return data.reshape(n_nodes, n_fields)


def write_vector_field(path_prefix, v, n_fields=1):
    v = np.ascontiguousarray(v, dtype=np.float64)
    if v.ndim == 2:
        v = v[:, None, :]
    n_nodes, n_fields, nd = v.shape
    with open(path_prefix + ".lbvec", "wb") as fh:
        np.array([n_fields, nd, n_nodes], dtype=np.int32).tofile(fh)
        v.tofile(fh)


def read_vector_field(path_prefix, expect=None):
    (n_fields, nd, n_nodes), data = _read(path_prefix + ".lbvec", 3, expect, "VectorField")
    return data.reshape(n_nodes, n_fields, nd)
