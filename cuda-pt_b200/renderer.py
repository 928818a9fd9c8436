"""Host-side mirror of the reference's Python entry surface for the wavefront path.

`PythonRenderer(scene_path, device_id, seed_offset)` with `render() / variance() / counter() /
avg_frame_time() / info() / release()` follows `pyrender.PythonRenderer`
(reference: src/pyrender/python_bind.cu:61-70, src/pyrender/python_render.cu:53-148):
one more sample per `render()` call, a FRESH `[H, W, 4] float32` CUDA tensor holding the running
mean (gamma applied iff the XML asks for it), `variance()` -> `[H, W, 1]`.

Everything is computed by libcpt_b200.so through its C ABI (include/cpt_b200.h); torch is only used
to own the returned device memory.
"""
import ctypes as C
import time

import numpy as np

from . import _lib
from ._lib import CptError, check, lib

CFG = dict(width=0, height=1, spp=2, max_depth=3, max_diffuse=4, max_specular=5, max_transmit=6, gamma=7,
           renderer=8, num_bsdfs=9, num_emitters=10, envmap_id=11, num_objects=12, num_prims=13, num_nodes=14,
           num_cache=15, cache_level=16, max_node_num=17, use_sbvh=18)


class Scene:
    """Flat host scene (XML loader or .cpts container)."""

    def __init__(self, path):
        self._h = C.c_void_p()
        if isinstance(path, (bytes, bytearray, memoryview)):  # a .cpts container image in memory
            buf = bytes(path)
            check(lib().cpt_b200_scene_load_memory(buf, len(buf), C.byref(self._h)))
            self.path = "<memory>"
            return
        p = str(path).encode()
        if str(path).endswith(".cpts"):
            check(lib().cpt_b200_scene_load(p, C.byref(self._h)))
        else:
            check(lib().cpt_b200_scene_load_xml(p, C.byref(self._h)))
        self.path = str(path)

    @property
    def config(self):
        arr = (C.c_int32 * 32)()
        check(lib().cpt_b200_scene_config(self._h, arr))
        return {k: int(arr[i]) for k, i in CFG.items()}

    @property
    def camera(self):
        arr = (C.c_float * 18)()
        check(lib().cpt_b200_scene_camera(self._h, arr))
        return np.array(arr, dtype=np.float32)

    def set_film(self, width, height):
        check(lib().cpt_b200_scene_set_film(self._h, int(width), int(height)))

    def save(self, path):
        check(lib().cpt_b200_scene_save(self._h, str(path).encode()))

    def close(self):
        if self._h:
            lib().cpt_b200_scene_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class WavefrontRenderer:
    """Thin object over a cpt_b200_renderer handle (C++ surface: TracerBase, tracer_base.cuh:29-160)."""

    def __init__(self, scene, device=0, seed_offset=0):
        self._h = C.c_void_p()
        self.scene = scene if isinstance(scene, Scene) else Scene(scene)
        check(lib().cpt_b200_create(self.scene._h, int(device), int(seed_offset), C.byref(self._h)))
        self.device = int(device)
        self.width = lib().cpt_b200_width(self._h)
        self.height = lib().cpt_b200_height(self._h)

    # -- options
    def set_option(self, key, value):
        check(lib().cpt_b200_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key):
        return int(lib().cpt_b200_get_option(self._h, key.encode()))

    def enable_variance(self, on=True):
        check(lib().cpt_b200_enable_variance(self._h, int(bool(on))))

    def set_seed_offset(self, v):
        check(lib().cpt_b200_set_seed_offset(self._h, int(v)))

    def update_camera(self, cam18):
        arr = (C.c_float * 18)(*[float(x) for x in cam18])
        check(lib().cpt_b200_update_camera(self._h, arr))

    def reset(self):
        check(lib().cpt_b200_reset(self._h))

    # -- rendering
    def render_raw(self, gamma=False):
        """One more sample; returns the device pointer (int) of the running-mean buffer."""
        p = C.c_void_p()
        check(lib().cpt_b200_render_raw(self._h, int(bool(gamma)), C.byref(p)))
        return p.value

    def render_batch(self, spp, gamma=False):
        p = C.c_void_p()
        check(lib().cpt_b200_render_batch(self._h, int(spp), int(bool(gamma)), C.byref(p)))
        return p.value

    def render_to_host(self, spp, gamma=False, out=None):
        """Render `spp` samples and copy the running mean into a host array (H, W, 4) float32.

        `out` may be a numpy array or a (pinned) torch CPU tensor; it is written in place."""
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.float32)
        ptr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        check(lib().cpt_b200_render_to_host(self._h, int(spp), int(bool(gamma)), C.c_void_p(ptr)))
        return out

    def render_offline(self, num_iter, gamma=True):
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        check(lib().cpt_b200_render_offline(self._h, int(num_iter), int(bool(gamma)), C.c_void_p(out.ctypes.data)))
        return out

    def download_output(self):
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        check(lib().cpt_b200_download_output(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def download_variance(self):
        out = np.empty((self.height, self.width, 1), dtype=np.float32)
        check(lib().cpt_b200_download_variance(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def debug_state(self, which):
        """Per-pixel path-state array of sample 0 of the last batch (parity probe)."""
        shape, dt = {"rng": ((2,), np.uint32), "pdf": ((), np.float32)}.get(which, ((4,), np.float32))
        out = np.empty((self.height, self.width) + shape, dtype=dt)
        check(lib().cpt_b200_debug_state(self._h, which.encode(), C.c_void_p(out.ctypes.data)))
        return out

    def output_ptr(self):
        p = C.c_void_p()
        check(lib().cpt_b200_output(self._h, C.byref(p)))
        return p.value

    def accumulator_ptr(self):
        p = C.c_void_p()
        check(lib().cpt_b200_accumulator(self._h, C.byref(p)))
        return p.value

    def variance_ptr(self):
        p = C.c_void_p()
        check(lib().cpt_b200_variance(self._h, C.byref(p)))
        return p.value

    def counter(self):
        return int(lib().cpt_b200_counter(self._h))

    def set_counter(self, v):
        check(lib().cpt_b200_set_counter(self._h, int(v)))

    def stats(self):
        arr = (C.c_uint64 * 8)()
        check(lib().cpt_b200_stats(self._h, arr))
        keys = ["samples", "closest_rays", "shadow_rays", "shaded", "launches", "device_us", "shadow_traced"]
        d = {k: int(arr[i]) for i, k in enumerate(keys)}
        d["rays"] = d["samples"] + d["closest_rays"] + d["shadow_rays"]
        return d

    def kernel_times(self):
        """Per-kernel device time (needs set_option("kernel_timing", 1)): {name: (ms, launches, units)}."""
        ms, ln, un = (C.c_double * 5)(), (C.c_uint64 * 5)(), (C.c_uint64 * 5)()
        check(lib().cpt_b200_kernel_times(self._h, ms, ln, un))
        return {k: {"ms": float(ms[i]), "launches": int(ln[i]), "units": int(un[i])}
                for i, k in enumerate(["raygen", "shade", "trace", "splat", "shadow"])}

    def queue_counts(self, n=64):
        """(paths, shadow) per bounce of the last batch: segments shaded / extended and shadow rays traversed."""
        a, b = (C.c_int32 * n)(), (C.c_int32 * n)()
        k = lib().cpt_b200_queue_counts(self._h, a, b, n)
        if k < 0:
            raise CptError(lib().cpt_b200_last_error().decode("utf-8", "replace"))
        return list(a[:k]), list(b[:k])

    def info(self):
        buf = C.create_string_buffer(2048)
        check(lib().cpt_b200_info(self._h, buf, 2048))
        return buf.value.decode()

    def stream(self):
        return int(lib().cpt_b200_stream(self._h))

    def sync(self):
        check(lib().cpt_b200_sync(self._h))

    def close(self):
        if self._h:
            lib().cpt_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DevView:
    """__cuda_array_interface__ view over renderer-owned device memory (zero copy into torch)."""

    def __init__(self, ptr, shape, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def device_tensor(ptr, shape, device, owner=None):
    """torch view (no copy) of a renderer-owned fp32 device buffer."""
    import torch
    return torch.as_tensor(_DevView(ptr, shape, owner), device=f"cuda:{device}")


class PythonRenderer:
    """Drop-in for `pyrender.PythonRenderer` (reference: src/pyrender/python_bind.cu:61-70)."""

    def __init__(self, scene_path, device_id, seed_offset):
        self._r = WavefrontRenderer(scene_path, device_id, seed_offset)
        cfg = self._r.scene.config
        # RendererType (src/core/enums.cuh:26-37): this library serves the path tracers; the reference's other integrators
        # (light tracer, depth / BVH-cost visualisers, ...) are out of scope and are refused rather than silently replaced
        served = {0: "megakernel-pt rules", 8: "megakernel-pt rules", 1: "wavefront rules", 6: "volumetric-pt rules"}
        if cfg["renderer"] not in served:
            self._r.close()
            raise ValueError(f"cuda-pt_b200: renderer type {cfg['renderer']} of '{scene_path}' is not a path tracer this library serves "
                             "(pt / ptd / wfpt / vpt)")
        self._rules = served[cfg["renderer"]]
        self._gamma = bool(cfg["gamma"])
        self._r.enable_variance(True)  # python_render.cu:93,99: initialize_var_buffer()
        self._times = []
        self.device_id = int(device_id)

    def render(self):
        import torch
        t0 = time.perf_counter()
        ptr = self._r.render_raw(self._gamma)
        self._times = (self._times + [(time.perf_counter() - t0) * 1e3])[-32:]
        view = device_tensor(ptr, (self._r.height, self._r.width, 4), self.device_id, self._r)
        return self._copy(view)  # fresh allocation per call, like gpu_ndarray_deep_copy (python_render.cu:32-51)

    def _copy(self, view):
        """Deep copy of a renderer-owned buffer that is COMPLETE when this returns (the reference uses a synchronous
        cudaMemcpy, python_render.cu:45).  The renderer works on its own non-blocking stream, so a clone merely enqueued on
        torch's current stream could still be pending when the next render() overwrites the buffer."""
        import torch
        out = view.clone()
        torch.cuda.current_stream(self.device_id).synchronize()
        return out

    def variance(self):
        ptr = self._r.variance_ptr()
        if not ptr:
            import torch
            return torch.empty(0)
        return self._copy(device_tensor(ptr, (self._r.height, self._r.width, 1), self.device_id, self._r))

    def counter(self):
        return self._r.counter()

    def avg_frame_time(self):
        return float(sum(self._times) / len(self._times)) if self._times else 0.0

    def info(self):
        print(self._r.info())

    def release(self):
        self._r.close()
