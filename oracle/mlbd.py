"""TEST INFRASTRUCTURE — reader/writer for the "MLBD" dump container used by oracle/ref_harness.cpp.

Record layout: u32 name_len | name | u32 dtype (0=f64,1=u32,2=i32,3=u8) | u32 ndim | u64 dims[ndim] | raw data.
"""
import struct
import numpy as np

_DT = {0: np.float64, 1: np.uint32, 2: np.int32, 3: np.uint8}
_CODE = {np.dtype(np.float64): 0, np.dtype(np.uint32): 1, np.dtype(np.int32): 2, np.dtype(np.uint8): 3}


def read(path):
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    o = 0
    while o < len(buf):
        (nl,) = struct.unpack_from("<I", buf, o); o += 4
        name = buf[o:o + nl].decode(); o += nl
        dt, nd = struct.unpack_from("<II", buf, o); o += 8
        dims = struct.unpack_from("<%dQ" % nd, buf, o); o += 8 * nd
        n = int(np.prod(dims)) if nd else 1
        dtype = np.dtype(_DT[dt])
        arr = np.frombuffer(buf, dtype=dtype, count=n, offset=o).reshape(dims).copy()
        o += n * dtype.itemsize
        out[name] = arr
    return out


def write(path, arrays):
    with open(path, "wb") as f:
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb))); f.write(nb)
            f.write(struct.pack("<II", _CODE[a.dtype], a.ndim))
            f.write(struct.pack("<%dQ" % a.ndim, *a.shape))
            f.write(a.tobytes())
