"""luisa-compute-rs_b200 — host-side mirror of luisa-compute-rs's runtime / rtx interface over the
B200 ray-tracing device (liblc_b200.so, C ABI in include/lc_b200_api.h).

The directory name carries a hyphen (it is the reference's name); import it through the
top-level shim `luisa_compute_rs_b200`.
"""
import os as _os

# A device uses more than eight streams (the user's, its own copy / build lanes, torch's).  With the CUDA default of eight hardware queues
# they alias, and a stream parked on a timeline value that has not been signalled yet (Event.wait before Event.signal, legal in the
# reference: cpu/resource.rs:10-44) can then hold up the very stream that is to signal it.  Effective only before the CUDA context exists.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _abi  # noqa: E402
from . import ir
from .runtime import BindlessArray, Buffer, BufferView, Context, Device, Event, LuisaError, Shader, Stream, Texture
from .rtx import (Accel, AccelBuildRequest, AccelOption, AccelUsageHint, CommittedHit, Curve, CurveBasis, HitType, Index, Mesh, ProceduralPrimitive, Ray, SurfaceCandidateFilter, SurfaceHit, INVALID,
                  affine_from_mat4, hit_valid, make_rays, offset_ray_origin)

__all__ = ["Accel", "AccelBuildRequest", "AccelOption", "AccelUsageHint", "CommittedHit", "HitType", "SurfaceCandidateFilter", "Buffer", "BufferView", "Context", "Device", "Event",
           "Index", "INVALID", "LuisaError", "Mesh", "Ray", "Stream", "SurfaceHit", "affine_from_mat4", "hit_valid", "make_rays",
           "offset_ray_origin", "_abi", "ir", "ProceduralPrimitive", "Curve", "CurveBasis", "BindlessArray", "Shader", "Texture"]
