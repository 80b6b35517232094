"""oracle/recfile.py -- TEST INFRASTRUCTURE: reader for the tagged-record dumps written by
oracle/ref_driver.cpp (format: b"REC1" | u32 namelen | name | dtype char | u64 count | payload)."""
import struct
import numpy as np


def read_rec(path):
    out = {}
    with open(path, "rb") as fh:
        data = fh.read()
    p = 0
    while p < len(data):
        assert data[p:p + 4] == b"REC1", "bad record magic at %d" % p
        (ln,) = struct.unpack_from("<I", data, p + 4)
        name = data[p + 8:p + 8 + ln].decode()
        p += 8 + ln
        typ = chr(data[p])
        (cnt,) = struct.unpack_from("<Q", data, p + 1)
        p += 9
        if typ == "i":
            arr = np.frombuffer(data, dtype="<i4", count=cnt, offset=p)
            p += 4 * cnt
        elif typ == "d":
            arr = np.frombuffer(data, dtype="<f8", count=cnt, offset=p)
            p += 8 * cnt
        else:
            raise ValueError("unknown record type %r" % typ)
        out[name] = arr.copy()
    return out
