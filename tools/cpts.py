"""Tiny numpy reader for the .cpts scene container (include/cpts_format.h)."""
import struct
import numpy as np

DT = {"config": np.int32, "camera": np.float32, "obj_idxs": np.int32, "emitter_prims": np.int32, "d65_spd": np.float32,
      "grid_bricks": np.int32, "grid_data": np.float32}


def read_cpts(path):
    data = open(path, "rb").read()
    assert data[:8] == b"CPTSCN01", "bad magic"
    n, = struct.unpack_from("<I", data, 8)
    off = 16
    out = {}
    for _ in range(n):
        name = data[off:off + 24].split(b"\0")[0].decode()
        elem, _r, count = struct.unpack_from("<IIQ", data, off + 24)
        off += 40
        nbytes = elem * count
        raw = data[off:off + nbytes]
        if name in DT:
            arr = np.frombuffer(raw, dtype=DT[name]).copy()
        elif name in ("cache_nodes",):
            arr = np.frombuffer(raw, dtype=np.uint32).reshape(-1, 4).copy()
        elif name in ("bsdfs", "emitters", "objects", "tex_desc", "media", "grids"):
            arr = np.frombuffer(raw, dtype=np.uint8).reshape(count, elem).copy()
        else:
            arr = np.frombuffer(raw, dtype=np.float32).reshape(count, elem // 4).copy()
        out[name] = arr
        off += nbytes + (16 - nbytes % 16) % 16
    return out
