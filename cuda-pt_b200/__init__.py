"""cuda-pt_b200 — B200-native (sm_100a) wavefront path-tracing hot path behind the entry surface of
Enigmatisms/cuda-pt (`PythonRenderer`, `pt <xml>`).  See DESIGN.md / INTEGRATION.md.

The directory name carries a hyphen (fixed by the project layout); import it as `cuda_pt_b200`
(repo-root shim) or through importlib."""
from ._lib import CptError, build, lib  # noqa: F401
from .renderer import PythonRenderer, Scene, WavefrontRenderer, device_tensor  # noqa: F401

__all__ = ["PythonRenderer", "Scene", "WavefrontRenderer", "device_tensor", "build", "lib", "CptError"]
