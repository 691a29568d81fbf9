"""rendertoy3c_b200 — B200-native (sm_100a) wavefront path tracer behind rendertoy3o's operator surface.

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI of include/rt3.h), api.py (host-side
mirror of the reference's device-scene operators) and scenes.py (synthetic inputs for BASELINE.json's
configs).  The CPU oracle is test infrastructure under oracle/ and is never imported from here.
"""
from . import scenes  # noqa: F401
from .api import Context, Rt3Error, camera_rays, camera_uvw, load_library, make_settings  # noqa: F401
