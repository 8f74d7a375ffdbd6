"""yuki_b200 — B200 (sm_100a) backend for yuki's per-pixel rendering hot path.

Layout: `csrc/` holds the CUDA kernels, the host-side scene assembly and the C ABI (include/yuki_gpu.h);
`capi` is the ctypes binding, `api` the host-side mirror of the reference interface (Scene / Camera /
Renderer / film_tiles), `desc` the plain scene description, `scenes` the synthetic benchmark scenes.
"""
from .desc import (CameraParameters, FilmSettings, IntegratorType, Light, Material, Mesh, SamplerType, SceneDesc, Texture,  # noqa: F401
                   Transform)
