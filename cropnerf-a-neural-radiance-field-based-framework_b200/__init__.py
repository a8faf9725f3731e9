"""cropnerf_b200 -- B200-native per-ray rendering hot path of FruitNeRF / CropNeRF.

Hand-written sm_100a CUDA kernels behind a C ABI (``include/cropnerf_b200.h``, ``libcropnerf_b200.so``) and a Python
host layer that mirrors the nerfstudio Field / Sampler / Renderer interfaces the reference plugs into
(``/root/reference/crop_nerf/fruit_nerf``).  See DESIGN.md and INTEGRATION.md.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (ctypes binding; loads lazily, fails loudly when the library is missing)
