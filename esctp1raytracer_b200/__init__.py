"""B200 (sm_100a) renderer for the per-pixel hot path of pg42819/EscTp1RayTracer.

Only what the path needs: ``csrc/`` (CUDA kernels + C ABI, built into
``libtracer_cuda.so``), ``api`` (host mirror of the reference interface),
``scenes`` (synthetic workloads), ``dist`` (row-band sharding across GPUs).
"""
from .api import (RNG_EXPLICIT, RNG_HASH, RNG_MT19937, Camera, Frame, MultiRenderer, Renderer, ResidentScene, Scene,  # noqa: F401
                  band_row_count, hash_faceids, mt19937_faceids, write_ppm)
from ._lib import TracerError  # noqa: F401
