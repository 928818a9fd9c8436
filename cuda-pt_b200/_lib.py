"""ctypes binding of libcpt_b200.so (include/cpt_b200.h) + the in-tree nvcc build recipe.

The library is the product: there is no Python/CPU fallback.  Importing this module
without a built library raises; creating a renderer without a CUDA device raises.
"""
import ctypes as C
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.environ.get("CPTB200_LIB", os.path.join(PKG_DIR, "libcpt_b200.so"))  # override: A/B builds of the kernels
SOURCES = ["wavefront.cu", "capi.cu", "scene_host.cpp", "scene_xml.cpp", "image_io.cpp", "jpeg_read.cpp", "hdr_read.cpp", "nvdb_read.cpp"]
NVCC_FLAGS = [
    "-O3", "--use_fast_math", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(REPO_ROOT, "include", f) for f in os.listdir(os.path.join(REPO_ROOT, "include"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA/C++ source of the package for sm_100a into libcpt_b200.so (in-tree)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("CPTB200_NVCC_EXTRA", "").split() + ["-I" + os.path.join(REPO_ROOT, "include"), "-I" + CSRC]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH, "-lz", "-ldl"]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
            "cuda-pt_b200 has no CPU or pure-Python fallback.")
    L = C.CDLL(LIB_PATH)
    vp, cp, i32, i64, u64 = C.c_void_p, C.c_char_p, C.c_int, C.c_int64, C.c_uint64
    fp = C.POINTER(C.c_float)
    sig = {
        "cpt_b200_last_error": (cp, []),
        "cpt_b200_version": (cp, []),
        "cpt_b200_scene_load_xml": (i32, [cp, C.POINTER(vp)]),
        "cpt_b200_scene_load": (i32, [cp, C.POINTER(vp)]),
        "cpt_b200_scene_load_memory": (i32, [vp, u64, C.POINTER(vp)]),
        "cpt_b200_scene_save": (i32, [vp, cp]),
        "cpt_b200_scene_config": (i32, [vp, C.POINTER(C.c_int32)]),
        "cpt_b200_scene_set_film": (i32, [vp, i32, i32]),
        "cpt_b200_scene_camera": (i32, [vp, fp]),
        "cpt_b200_scene_free": (None, [vp]),
        "cpt_b200_create": (i32, [vp, i32, i32, C.POINTER(vp)]),
        "cpt_b200_destroy": (None, [vp]),
        "cpt_b200_render_raw": (i32, [vp, i32, C.POINTER(vp)]),
        "cpt_b200_render_batch": (i32, [vp, i32, i32, C.POINTER(vp)]),
        "cpt_b200_render_to_host": (i32, [vp, i32, i32, vp]),
        "cpt_b200_render_offline": (i32, [vp, i32, i32, vp]),
        "cpt_b200_enable_variance": (i32, [vp, i32]),
        "cpt_b200_variance": (i32, [vp, C.POINTER(vp)]),
        "cpt_b200_download_variance": (i32, [vp, vp]),
        "cpt_b200_output": (i32, [vp, C.POINTER(vp)]),
        "cpt_b200_accumulator": (i32, [vp, C.POINTER(vp)]),
        "cpt_b200_download_output": (i32, [vp, vp]),
        "cpt_b200_debug_state": (i32, [vp, cp, vp]),
        "cpt_b200_counter": (i32, [vp]),
        "cpt_b200_set_counter": (i32, [vp, i32]),
        "cpt_b200_reset": (i32, [vp]),
        "cpt_b200_set_seed_offset": (i32, [vp, i32]),
        "cpt_b200_update_camera": (i32, [vp, fp]),
        "cpt_b200_width": (i32, [vp]),
        "cpt_b200_height": (i32, [vp]),
        "cpt_b200_set_option": (i32, [vp, cp, i64]),
        "cpt_b200_get_option": (i64, [vp, cp]),
        "cpt_b200_stats": (i32, [vp, C.POINTER(u64)]),
        "cpt_b200_kernel_times": (i32, [vp, C.POINTER(C.c_double), C.POINTER(u64), C.POINTER(u64)]),
        "cpt_b200_queue_counts": (i32, [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), i32]),
        "cpt_b200_info": (i32, [vp, C.c_char_p, i32]),
        "cpt_b200_stream": (u64, [vp]),
        "cpt_b200_sync": (i32, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._signatures = sig
    _lib = L
    return L


class CptError(RuntimeError):
    pass


def check(code):
    if code != 0:
        raise CptError(lib().cpt_b200_last_error().decode("utf-8", "replace"))
