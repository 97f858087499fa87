"""B200-native `fgfa depth` (FlatGFA node depth), host-side Python mirror.

The compute path lives in ``pollen_b200/lib/libflatgfa.so`` (hand-written sm_100a CUDA
kernels behind the C ABI of ``include/fgfa_depth.h`` and ``include/flatgfa.h``).  This
package only binds that ABI with ctypes and mirrors the reference's operator interface
for the path (reference: flatgfa/src/ops/depth.rs:15-82) so tests read like the
reference's own.  There is no CPU fallback: importing works anywhere, computing without
the built library or without a CUDA device raises.
"""
from .binding import (  # noqa: F401
    DepthError,
    DepthPlan,
    MultiDepth,
    FlatBED,
    FlatGFA,
    SegDepth,
    device_count,
    interval_depth_steps,
    lib,
    path_depth_steps,
    seg_depth,
    seg_depth_steps,
    seg_depth_with_uniq,
    seg_depth_with_uniq_steps,
    tokenize_steps,
    window_depth_steps,
)

__all__ = [
    "DepthError",
    "DepthPlan",
    "MultiDepth",
    "FlatBED",
    "FlatGFA",
    "SegDepth",
    "device_count",
    "interval_depth_steps",
    "lib",
    "path_depth_steps",
    "seg_depth",
    "seg_depth_steps",
    "seg_depth_with_uniq",
    "seg_depth_with_uniq_steps",
    "tokenize_steps",
    "window_depth_steps",
]
