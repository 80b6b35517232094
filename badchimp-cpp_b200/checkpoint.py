"""Checkpoint / restart files in the reference's raw formats (src/lbsolver/LBfield.h:102-138,
233-276, 378-421): native-endian int32 header, then the field's AoS doubles.

    .lbsca   int nFields, int nNodes,           double[nFields*nNodes]          (ScalarField)
    .lbvec   int nFields, int nD, int nNodes,   double[nFields*nD*nNodes]       (VectorField)
    .lblbf   int nFields, int nQ, int nNodes,   double[nFields*nQ*nNodes]       (LbField)

Arrays are node-major [node, field, component] exactly as chimp_upload_lbfield / chimp_download_*
take and return them, so a state written by the CPU code restarts on the GPU and vice versa."""
from __future__ import annotations

import numpy as np


def write_lbfield(path_prefix, f):
    f = np.ascontiguousarray(f, dtype=np.float64)
    n_nodes, n_fields, nq = f.shape
    with open(path_prefix + ".lblbf", "wb") as fh:
        np.array([n_fields, nq, n_nodes], dtype=np.int32).tofile(fh)
        f.tofile(fh)


def read_lbfield(path_prefix):
    with open(path_prefix + ".lblbf", "rb") as fh:
        n_fields, nq, n_nodes = np.fromfile(fh, dtype=np.int32, count=3)
        data = np.fromfile(fh, dtype=np.float64, count=int(n_fields) * int(nq) * int(n_nodes))
    if data.size != int(n_fields) * int(nq) * int(n_nodes):
        raise ValueError("truncated LbField file " + path_prefix + ".lblbf")
    return data.reshape(n_nodes, n_fields, nq)


def write_scalar_field(path_prefix, s):
    s = np.ascontiguousarray(s, dtype=np.float64)
    if s.ndim == 1:
        s = s[:, None]
    n_nodes, n_fields = s.shape
    with open(path_prefix + ".lbsca", "wb") as fh:
        np.array([n_fields, n_nodes], dtype=np.int32).tofile(fh)
        s.tofile(fh)


def read_scalar_field(path_prefix):
    with open(path_prefix + ".lbsca", "rb") as fh:
        n_fields, n_nodes = np.fromfile(fh, dtype=np.int32, count=2)
        data = np.fromfile(fh, dtype=np.float64, count=int(n_fields) * int(n_nodes))
    return data.reshape(n_nodes, n_fields)


def write_vector_field(path_prefix, v, n_fields=1):
    v = np.ascontiguousarray(v, dtype=np.float64)
    if v.ndim == 2:
        v = v[:, None, :]
    n_nodes, n_fields, nd = v.shape
    with open(path_prefix + ".lbvec", "wb") as fh:
        np.array([n_fields, nd, n_nodes], dtype=np.int32).tofile(fh)
        v.tofile(fh)


def read_vector_field(path_prefix):
    with open(path_prefix + ".lbvec", "rb") as fh:
        n_fields, nd, n_nodes = np.fromfile(fh, dtype=np.int32, count=3)
        data = np.fromfile(fh, dtype=np.float64, count=int(n_fields) * int(nd) * int(n_nodes))
    return data.reshape(n_nodes, n_fields, nd)
