"""Reader / writer of the flat binary tools/reference_fixtures/dump_golden.f90 produces:
    int32 1 (endianness probe) | records: char[24] name | int32 kind (1 int32, 2 float64) | int32 count | data
Arrays are in Fortran order.  The writer exists so that the consuming test can be exercised here (with oracle-made
content, clearly marked) although the reference itself cannot be built in this image."""
import numpy as np


def read_fixture(path):
    raw = open(path, "rb").read()
    probe = np.frombuffer(raw[:4], dtype="<i4")[0]
    if probe == 1:
        bo = "<"
    elif np.frombuffer(raw[:4], dtype=">i4")[0] == 1:
        bo = ">"                     # gfortran -fconvert=swap (the reference's own flags)
    else:
        raise ValueError("not a dump_golden file")
    pos, out = 4, {}
    while pos < len(raw):
        name = raw[pos:pos + 24].decode("ascii").strip()
        kind, count = np.frombuffer(raw[pos + 24:pos + 32], dtype=bo + "i4")
        pos += 32
        dt = np.dtype(bo + ("i4" if kind == 1 else "f8"))
        out[name] = np.frombuffer(raw[pos:pos + int(count) * dt.itemsize], dtype=dt).astype(dt.newbyteorder("="))
        pos += int(count) * dt.itemsize
    return out


def write_fixture(path, records):
    with open(path, "wb") as f:
        np.array([1], dtype="<i4").tofile(f)
        for name, a in records.items():
            a = np.asarray(a)
            kind = 1 if a.dtype.kind in "iu" else 2
            f.write(name.encode("ascii").ljust(24))
            np.array([kind, a.size], dtype="<i4").tofile(f)
            a.astype("<i4" if kind == 1 else "<f8").ravel(order="F").tofile(f)
