"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liboracle.so")


def build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.oracle_load.restype = C.c_void_p
        L.oracle_load.argtypes = [C.c_char_p]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_config.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.oracle_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_render_megakernel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_rng.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.oracle_float_to_half.restype = C.c_uint16
        L.oracle_float_to_half.argtypes = [C.c_float]
        L.oracle_half_to_float.restype = C.c_float
        L.oracle_half_to_float.argtypes = [C.c_uint16]
        for name in ("oracle_closest_hit", "oracle_closest_hit_bruteforce"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_unoccluded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
        _lib = L
    return _lib


class OracleScene:
    def __init__(self, cpts_path):
        self._h = lib().oracle_load(str(cpts_path).encode())
        if not self._h:
            raise RuntimeError("oracle: " + lib().oracle_last_error().decode())
        cfg = (C.c_int32 * 32)()
        lib().oracle_config(self._h, cfg)
        self.width, self.height, self.spp, self.max_depth = cfg[0], cfg[1], cfg[2], cfg[3]

    def render(self, spp, seed_offset=0, first_accum=1, max_depth=0, rows=None, sum_image=None, rules=0):
        """Returns (sum image H x W x 4 float32, stats dict).  Mean = sum / spp.
        rules: 0 wavefront termination rules, 1 megakernel "pt", 2 volumetric "vpt" (participating media)."""
        if sum_image is None:
            sum_image = np.zeros((self.height, self.width, 4), dtype=np.float32)
        stats = (C.c_uint64 * 3)()
        r0, r1 = rows if rows else (0, 0)
        if rules:
            rc = lib().oracle_render_megakernel(self._h, rules, first_accum, spp, seed_offset, max_depth, r0, r1,
                                                C.c_void_p(sum_image.ctypes.data), stats)
            if rc:
                raise ValueError("oracle: rules must be 0, 1 or 2")
        else:
            lib().oracle_render(self._h, first_accum, spp, seed_offset, max_depth, r0, r1,
                                C.c_void_p(sum_image.ctypes.data), stats)
        nrows = (r1 - r0) if rows else self.height
        samples = nrows * self.width * spp
        return sum_image, {"samples": samples, "closest_rays": int(stats[0]), "shadow_rays": int(stats[1]),
                           "shaded": int(stats[2]), "rays": samples + int(stats[0]) + int(stats[1])}

    def closest_hit(self, o, d, brute=False):
        o = np.asarray(o, np.float32); d = np.asarray(d, np.float32); out = np.zeros(4, np.float32)
        fn = lib().oracle_closest_hit_bruteforce if brute else lib().oracle_closest_hit
        prim = fn(self._h, C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data), C.c_void_p(out.ctypes.data))
        return prim, out[:3].copy()

    def unoccluded(self, o, d, max_dist):
        o = np.asarray(o, np.float32); d = np.asarray(d, np.float32)
        return bool(lib().oracle_unoccluded(self._h, C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data), float(max_dist)))

    def close(self):
        if self._h:
            lib().oracle_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rng_stream(seed, offset, n):
    out = np.zeros(n, dtype=np.uint32)
    lib().oracle_rng(int(seed), int(offset), int(n), C.c_void_p(out.ctypes.data))
    return out
