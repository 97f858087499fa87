"""ctypes binding of libflatgfa.so and the Python mirror of the depth op.

Names follow the reference: ``seg_depth_with_uniq`` / ``seg_depth`` / ``SegDepth``
(flatgfa/src/ops/depth.rs:15,45,61) and the flatgfa-c accessors
(flatgfa-c/src/lib.rs:62-172).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libflatgfa.so")
_lib = None

FGFA_OK = 0
FGFA_ERR_INVALID_ARG = -1
FGFA_ERR_BAD_MAGIC = -2
FGFA_ERR_TRUNCATED = -3
FGFA_ERR_SPAN_OOB = -4
FGFA_ERR_SEG_OOB = -5
FGFA_ERR_CUDA = -6
FGFA_ERR_NOMEM = -7
FGFA_ERR_NO_DEVICE = -8
FGFA_ERR_TOO_LARGE = -9
FGFA_ERR_PARSE = -10


class DepthError(RuntimeError):
    """A failed call into the depth engine; ``code`` is the FGFA_ERR_* value."""

    def __init__(self, code: int, detail: str = ""):
        self.code = code
        msg = lib().fgfa_strerror(code).decode()
        if detail:
            msg += ": " + detail
        super().__init__(msg)


class _String(C.Structure):  # flatgfa_string_t (flatgfa-c/src/lib.rs:36-40)
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_int)]


class _Handle(C.Structure):  # flatgfa_handle_t (flatgfa-c/src/lib.rs:140-144)
    _fields_ = [("segment_id", C.c_uint32), ("is_forward", C.c_bool)]


EXPORTS = {
    # include/flatgfa.h
    "flatgfa_parse": (C.c_void_p, [C.c_char_p]),
    "flatgfa_free": (None, [C.c_void_p]),
    "flatgfa_get_segment_count": (C.c_uint32, [C.c_void_p]),
    "flatgfa_get_seq": (_String, [C.c_void_p, C.c_uint32]),
    "flatgfa_path_count": (C.c_uint32, [C.c_void_p]),
    "flatgfa_get_path_name": (_String, [C.c_void_p, C.c_uint32]),
    "flatgfa_get_path_step_count": (C.c_uint32, [C.c_void_p, C.c_uint32]),
    "flatgfa_get_step": (C.c_bool, [C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(_Handle)]),
    "flatgfa_load": (C.c_void_p, [C.c_char_p]),
    "flatgfa_seg_depth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "flatgfa_format_seg_depth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "flatgfa_path_depth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "flatgfa_format_path_depth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "flatgfa_dump": (C.c_int, [C.c_void_p, C.c_char_p]),
    "flatgfa_format_gfa": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "flatgfa_parse_mem": (C.c_void_p, [C.c_void_p, C.c_size_t]),
    "flatgfa_image_size": (C.c_size_t, [C.c_void_p]),
    "flatgfa_dump_mem": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "flatbed_parse_mem": (C.c_void_p, [C.c_void_p, C.c_size_t]),
    "flatbed_make_windows": (C.c_void_p, [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64]),
    "flatbed_free": (None, [C.c_void_p]),
    "flatbed_entry_count": (C.c_uint64, [C.c_void_p]),
    "flatbed_get_entry": (C.c_bool, [C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "flatgfa_interval_depth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "flatgfa_window_depth": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "flatgfa_bed_depth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "flatgfa_last_error": (C.c_char_p, []),
    # include/fgfa_depth.h
    "fgfa_strerror": (C.c_char_p, [C.c_int]),
    "fgfa_last_error": (C.c_char_p, []),
    "fgfa_device_count": (C.c_int, []),
    "fgfa_depth_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_size_t]),
    "fgfa_depth_plan_destroy": (None, [C.c_void_p]),
    "fgfa_depth_plan_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_depth_plan_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_depth_plan_feed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_depth_plan_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_depth_plan_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fgfa_depth_plan_use_bitmap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "fgfa_depth_plan_run_stream_only": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_exchange_uniq_depth": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]),
    "fgfa_exchange_recv_bytes": (C.c_size_t, [C.c_int, C.c_uint32]),
    "fgfa_exchange_push": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "fgfa_exchange_reduce": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "fgfa_depth_plan_set_uniq_width": (C.c_int, [C.c_void_p, C.c_int]),
    "fgfa_depth_multi_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int]),
    "fgfa_depth_multi_destroy": (None, [C.c_void_p]),
    "fgfa_depth_multi_partition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_depth_multi_upload": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fgfa_depth_multi_device_steps": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "fgfa_depth_multi_run": (C.c_int, [C.c_void_p, C.c_int]),
    "fgfa_depth_multi_sync": (C.c_int, [C.c_void_p]),
    "fgfa_depth_multi_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_depth_multi_run_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_depth_multi_result_device": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
    "fgfa_depth_multi_last_error": (C.c_char_p, []),
    "fgfa_lpt_partition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]),
    "fgfa_depth_plan_set_engine": (C.c_int, [C.c_void_p, C.c_int]),
    "fgfa_depth_plan_engine": (C.c_int, [C.c_void_p]),
    "fgfa_depth_plan_autotune": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_depth_plan_set_probe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_depth_plan_launches": (C.c_uint32, [C.c_void_p, C.c_int]),
    "fgfa_depth_plan_scratch_bytes": (C.c_size_t, [C.c_void_p]),
    "fgfa_depth_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_seg_depth_with_uniq_steps": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "fgfa_release_workspace": (None, []),
    "fgfa_tokenizer_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32]),
    "fgfa_tokenizer_spans": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "fgfa_tokenizer_parse": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "fgfa_tokenizer_device_steps": (C.c_void_p, [C.c_void_p]),
    "fgfa_tokenizer_destroy": (None, [C.c_void_p]),
    "fgfa_path_depth_steps": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_depth_plan_path_sums": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_interval_scratch_bytes": (C.c_size_t, [C.c_uint64, C.c_uint64]),
    "fgfa_path_offsets_device": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "fgfa_make_windows_device": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgfa_interval_depth_device": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "fgfa_interval_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fgfa_interval_depth_steps": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "fgfa_window_depth_steps": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fgfa_free": (None, [C.c_void_p]),
    "fgfa_flatgfa_counts": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fgfa_seg_depth_with_uniq": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "fgfa_seg_depth": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
}


def lib() -> C.CDLL:
    """Load libflatgfa.so (built in-tree by ``make`` / ``__graft_entry__.build()``).

    Fails loudly if the library is missing: the product path has no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(
                f"{_LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()); "
                "pollen_b200 has no CPU fallback"
            )
        handle = C.CDLL(_LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def device_count() -> int:
    return int(lib().fgfa_device_count())


def _check(rc: int) -> None:
    if rc != FGFA_OK:
        raise DepthError(rc, lib().fgfa_last_error().decode())


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def seg_depth_with_uniq_steps(steps, span_start, span_end, n_segs: int, want_uniq: bool = True):
    """seg_depth_with_uniq (depth.rs:15-39) over raw host arrays: ``steps`` are Handle
    words, ``span_start/span_end`` each path's half-open range.  Returns (depth, uniq)
    as uint64 arrays (uniq is None if ``want_uniq`` is false)."""
    steps, span_start, span_end = _u32(steps), _u32(span_start), _u32(span_end)
    depth = np.empty(n_segs, dtype=np.uint64)
    uniq = np.empty(n_segs, dtype=np.uint64) if want_uniq else None
    _check(
        lib().fgfa_seg_depth_with_uniq_steps(
            steps.ctypes.data, steps.size, span_start.ctypes.data, span_end.ctypes.data,
            span_start.size, n_segs, depth.ctypes.data, uniq.ctypes.data if want_uniq else None,
        )
    )
    return depth, uniq


def path_depth_steps(steps, span_start, span_end, seg_len, path_ids=None):
    """path_depth (depth.rs:88-113) over raw host arrays.  Returns (lengths u64, weighted
    sums u64, mean depths f64) for the queried path ids (default: all paths in order)."""
    steps, span_start, span_end, seg_len = _u32(steps), _u32(span_start), _u32(span_end), _u32(seg_len)
    ids = None if path_ids is None else _u32(path_ids)
    n = span_start.size if ids is None else ids.size
    lengths = np.empty(n, dtype=np.uint64)
    weighted = np.empty(n, dtype=np.uint64)
    means = np.empty(n, dtype=np.float64)
    _check(
        lib().fgfa_path_depth_steps(
            steps.ctypes.data, steps.size, span_start.ctypes.data, span_end.ctypes.data, span_start.size,
            seg_len.ctypes.data, seg_len.size, None if ids is None else ids.ctypes.data, n,
            lengths.ctypes.data, weighted.ctypes.data, means.ctypes.data,
        )
    )
    return lengths, weighted, means


def interval_depth_steps(steps, span_start, span_end, seg_len, path: int, win_start, win_end):
    """interval_depth / bed_depth (window_depth.rs:176-180, 203-211) over raw host arrays: mean
    depth (f64, bit-identical to the reference) of each [win_start, win_end) along path ``path``."""
    steps, span_start, span_end, seg_len = _u32(steps), _u32(span_start), _u32(span_end), _u32(seg_len)
    ws = np.ascontiguousarray(win_start, dtype=np.uint64)
    we = np.ascontiguousarray(win_end, dtype=np.uint64)
    out = np.empty(ws.size, dtype=np.float64)
    _check(
        lib().fgfa_interval_depth_steps(
            steps.ctypes.data, steps.size, span_start.ctypes.data, span_end.ctypes.data, span_start.size,
            seg_len.ctypes.data, seg_len.size, path, ws.ctypes.data, we.ctypes.data, ws.size, out.ctypes.data,
        )
    )
    return out


def window_depth_steps(steps, span_start, span_end, seg_len, path: int, window_size: int):
    """window_depth (window_depth.rs:183-197) over raw host arrays.  Returns (win_start u64,
    win_end u64, depths f64, path_length)."""
    steps, span_start, span_end, seg_len = _u32(steps), _u32(span_start), _u32(span_end), _u32(seg_len)
    out, m, total = C.c_void_p(), C.c_uint64(), C.c_uint64()
    _check(
        lib().fgfa_window_depth_steps(
            steps.ctypes.data, steps.size, span_start.ctypes.data, span_end.ctypes.data, span_start.size,
            seg_len.ctypes.data, seg_len.size, path, window_size, C.byref(out), C.byref(m), C.byref(total),
        )
    )
    try:
        depths = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_double)), shape=(m.value,)).copy() if m.value else np.empty(0)
    finally:
        lib().fgfa_free(out)
    ws = np.arange(m.value, dtype=np.uint64) * np.uint64(window_size)
    we = np.minimum(ws + np.uint64(window_size), np.uint64(total.value))
    return ws, we, depths, total.value


def tokenize_steps(text: bytes, fields, sequential_max: int, others=None):
    """GPU step-list tokenizer: ``fields`` is a list of (offset, length) of step-list text inside
    ``text``; names 1..sequential_max map to name-1, ``others`` is {name: id}.  Returns
    (steps u32, span_start, span_end); raises DepthError(FGFA_ERR_PARSE) outside the strict grammar."""
    buf = np.frombuffer(text, dtype=np.uint8)
    off = np.ascontiguousarray([f[0] for f in fields], dtype=np.uint64)
    ln = np.ascontiguousarray([f[1] for f in fields], dtype=np.uint64)
    h = C.c_void_p()
    _check(lib().fgfa_tokenizer_create(C.byref(h), buf.ctypes.data if buf.size else None, buf.size,
                                       off.ctypes.data, ln.ctypes.data, len(fields)))
    try:
        start = np.empty(len(fields), dtype=np.uint32)
        end = np.empty(len(fields), dtype=np.uint32)
        n = C.c_uint64()
        _check(lib().fgfa_tokenizer_spans(h, start.ctypes.data, end.ctypes.data, C.byref(n)))
        steps = np.empty(n.value, dtype=np.uint32)
        others = others or {}
        names = np.ascontiguousarray(list(others.keys()), dtype=np.uint64)
        ids = np.ascontiguousarray(list(others.values()), dtype=np.uint32)
        _check(lib().fgfa_tokenizer_parse(h, sequential_max, names.ctypes.data if names.size else None,
                                          ids.ctypes.data if ids.size else None, names.size,
                                          steps.ctypes.data if steps.size else None))
        return steps, start, end
    finally:
        lib().fgfa_tokenizer_destroy(h)


def seg_depth_steps(steps, span_start, span_end, n_segs: int) -> np.ndarray:
    """seg_depth (depth.rs:45-56) over raw host arrays."""
    return seg_depth_with_uniq_steps(steps, span_start, span_end, n_segs, want_uniq=False)[0]


class FlatBED:
    """A list of named intervals (``flatbed_t``; flatbed.rs:19-33): a parsed BED file or the
    equally sized windows of ``Windows`` (window_depth.rs:20-58)."""

    def __init__(self, handle):
        if not handle:
            raise DepthError(FGFA_ERR_INVALID_ARG, lib().flatgfa_last_error().decode())
        self._h = C.c_void_p(handle)

    @classmethod
    def parse(cls, bed_text: bytes) -> "FlatBED":  # BEDParser::parse_mem, flatbed.rs:126-131
        buf = C.create_string_buffer(bed_text, len(bed_text))
        return cls(lib().flatbed_parse_mem(C.cast(buf, C.c_void_p), len(bed_text)))

    @classmethod
    def windows(cls, name: bytes, start: int, end: int, size: int) -> "FlatBED":  # Windows::as_bed
        buf = C.create_string_buffer(name, len(name))
        return cls(lib().flatbed_make_windows(C.cast(buf, C.c_void_p), len(name), start, end, size))

    def close(self) -> None:
        if self._h:
            lib().flatbed_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(lib().flatbed_entry_count(self._h))

    def entries(self):
        """[(name bytes, start, end)] in file order."""
        out = []
        s, a, b = _String(), C.c_uint64(), C.c_uint64()
        for i in range(len(self)):
            assert lib().flatbed_get_entry(self._h, i, C.byref(s), C.byref(a), C.byref(b))
            out.append((C.string_at(s.data, s.len) if s.len else b"", a.value, b.value))
        return out


class FlatGFA:
    """A graph handle (``flatgfa_t``): parsed from GFA text or mapped from a .flatgfa file."""

    def __init__(self, handle: int):
        if not handle:
            raise DepthError(FGFA_ERR_INVALID_ARG, lib().flatgfa_last_error().decode())
        self._h = C.c_void_p(handle)

    @classmethod
    def parse(cls, gfa_path: str) -> "FlatGFA":  # flatgfa_parse, lib.rs:62-68
        return cls(lib().flatgfa_parse(os.fsencode(gfa_path)))

    @classmethod
    def parse_bytes(cls, gfa_text: bytes) -> "FlatGFA":  # flatgfa-py parse_bytes, lib.rs:69-71
        buf = C.create_string_buffer(gfa_text, len(gfa_text))
        return cls(lib().flatgfa_parse_mem(C.cast(buf, C.c_void_p), len(gfa_text)))

    def image(self) -> np.ndarray:
        """The graph's .flatgfa image (``file::dump``, file.rs:290-307) as a uint8 array."""
        n = int(lib().flatgfa_image_size(self._h))
        out = np.empty(n, dtype=np.uint8)
        _check(lib().flatgfa_dump_mem(self._h, out.ctypes.data, n))
        return out

    @classmethod
    def load(cls, flatgfa_path: str) -> "FlatGFA":  # memfile::map_file + file::view
        return cls(lib().flatgfa_load(os.fsencode(flatgfa_path)))

    def close(self) -> None:
        if self._h:
            lib().flatgfa_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def segment_count(self) -> int:
        return int(lib().flatgfa_get_segment_count(self._h))

    @property
    def path_count(self) -> int:
        return int(lib().flatgfa_path_count(self._h))

    def seq(self, segment_id: int) -> Optional[bytes]:
        s = lib().flatgfa_get_seq(self._h, segment_id)
        return C.string_at(s.data, s.len) if s.data else None

    def path_name(self, path_index: int) -> Optional[bytes]:
        s = lib().flatgfa_get_path_name(self._h, path_index)
        return C.string_at(s.data, s.len) if s.data else None

    def path_step_count(self, path_index: int) -> int:
        return int(lib().flatgfa_get_path_step_count(self._h, path_index))

    def step(self, path_index: int, step_index: int) -> Optional[Tuple[int, bool]]:
        out = _Handle()
        ok = lib().flatgfa_get_step(self._h, path_index, step_index, C.byref(out))
        return (int(out.segment_id), bool(out.is_forward)) if ok else None

    def dump(self, flatgfa_path: str) -> None:
        _check(lib().flatgfa_dump(self._h, os.fsencode(flatgfa_path)))

    def format_gfa(self) -> bytes:
        """The graph as GFA text (print.rs:144-152)."""
        out, n = C.c_void_p(), C.c_size_t()
        _check(lib().flatgfa_format_gfa(self._h, C.byref(out), C.byref(n)))
        try:
            return C.string_at(out, n.value)
        finally:
            C.CDLL(None).free(out)

    def seg_depth_with_uniq(self) -> Tuple[np.ndarray, np.ndarray]:
        n = self.segment_count
        depth = np.empty(n, dtype=np.uint64)
        uniq = np.empty(n, dtype=np.uint64)
        _check(lib().flatgfa_seg_depth(self._h, depth.ctypes.data, uniq.ctypes.data))
        return depth, uniq

    def seg_depth(self) -> np.ndarray:
        depth = np.empty(self.segment_count, dtype=np.uint64)
        _check(lib().flatgfa_seg_depth(self._h, depth.ctypes.data, None))
        return depth

    def path_depth(self, path_ids=None):
        """``ops::depth::path_depth`` (depth.rs:88-113): (lengths u64, mean depths f64)."""
        ids = None if path_ids is None else _u32(path_ids)
        n = self.path_count if ids is None else int(ids.size)
        lengths = np.empty(n, dtype=np.uint64)
        means = np.empty(n, dtype=np.float64)
        _check(lib().flatgfa_path_depth(self._h, None if ids is None else ids.ctypes.data, n, lengths.ctypes.data, means.ctypes.data))
        return lengths, means

    def format_path_depth(self, lengths, means, path_ids=None) -> bytes:
        ids = None if path_ids is None else _u32(path_ids)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint64)
        means = np.ascontiguousarray(means, dtype=np.float64)
        out, n = C.c_void_p(), C.c_size_t()
        _check(lib().flatgfa_format_path_depth(self._h, None if ids is None else ids.ctypes.data, lengths.size,
                                               lengths.ctypes.data, means.ctypes.data, C.byref(out), C.byref(n)))
        try:
            return C.string_at(out, n.value)
        finally:
            C.CDLL(None).free(out)

    def interval_depth(self, bed: "FlatBED") -> np.ndarray:
        """``ops::window_depth::bed_depth`` (window_depth.rs:203-211): one f64 per entry."""
        out = np.empty(len(bed), dtype=np.float64)
        _check(lib().flatgfa_interval_depth(self._h, bed._h, out.ctypes.data))
        return out

    def window_depth(self, path_name: str, window_size: int) -> bytes:
        """``fgfa window-depth PATH SIZE`` (cmds.rs:477-496): the IntervalDepth table."""
        out, n = C.c_void_p(), C.c_size_t()
        _check(lib().flatgfa_window_depth(self._h, path_name.encode(), window_size, C.byref(out), C.byref(n)))
        try:
            return C.string_at(out, n.value)
        finally:
            C.CDLL(None).free(out)

    def bed_depth(self, bed_text: bytes) -> bytes:
        """``fgfa depth -b BED`` (cmds.rs:246-255): the IntervalDepth table for a BED file's text."""
        out, n = C.c_void_p(), C.c_size_t()
        buf = C.create_string_buffer(bed_text, len(bed_text))
        _check(lib().flatgfa_bed_depth(self._h, C.cast(buf, C.c_void_p), len(bed_text), C.byref(out), C.byref(n)))
        try:
            return C.string_at(out, n.value)
        finally:
            C.CDLL(None).free(out)

    def format_seg_depth(self, depth: np.ndarray, uniq: np.ndarray) -> bytes:
        depth = np.ascontiguousarray(depth, dtype=np.uint64)
        uniq = np.ascontiguousarray(uniq, dtype=np.uint64)
        out, n = C.c_void_p(), C.c_size_t()
        _check(lib().flatgfa_format_seg_depth(self._h, depth.ctypes.data, uniq.ctypes.data, C.byref(out), C.byref(n)))
        try:
            return C.string_at(out, n.value)
        finally:
            C.CDLL(None).free(out)


def seg_depth_with_uniq(gfa: FlatGFA) -> Tuple[np.ndarray, np.ndarray]:
    """``ops::depth::seg_depth_with_uniq`` (flatgfa/src/ops/depth.rs:15)."""
    return gfa.seg_depth_with_uniq()


def seg_depth(gfa: FlatGFA) -> np.ndarray:
    """``ops::depth::seg_depth`` (flatgfa/src/ops/depth.rs:45)."""
    return gfa.seg_depth()


class SegDepth:
    """``ops::depth::SegDepth`` (flatgfa/src/ops/depth.rs:61-82): the printable table."""

    def __init__(self, gfa: FlatGFA, depths: np.ndarray, uniq_depths: np.ndarray):
        self.gfa, self.depths, self.uniq_depths = gfa, depths, uniq_depths

    def emit(self) -> bytes:
        return self.gfa.format_seg_depth(self.depths, self.uniq_depths)


def exchange_uniq_depth(n_ranks, rank, bitmaps, rows, partial_depths, final_depths, final_uniqs, n_segs, stream=0,
                        multicast_base=0, off_partial=0, off_final_depth=0, off_final_uniq=0):
    """One fused popcount + exchange launch (see fgfa_exchange_uniq_depth)."""
    P = C.c_void_p * n_ranks
    r = np.ascontiguousarray(rows, dtype=np.uint32)
    _check(lib().fgfa_exchange_uniq_depth(n_ranks, rank, P(*bitmaps), r.ctypes.data, P(*partial_depths),
                                          P(*final_depths), P(*final_uniqs), n_segs, multicast_base or None,
                                          off_partial, off_final_depth, off_final_uniq, stream or None))


def exchange_recv_bytes(n_ranks: int, n_segs: int) -> int:
    """Bytes of one rank's receive buffer for the push form of the exchange."""
    return int(lib().fgfa_exchange_recv_bytes(int(n_ranks), int(n_segs)))


def exchange_push(n_ranks, rank, bitmap, rows, partial_depth, recv_bufs, n_segs, stream=0, partial_uniq=0):
    """Kernel P (see fgfa_exchange_push): this rank's u8 uniq counts + partial depth into the slice owners' slots.
    ``partial_uniq``: device pointer to u8 counts that already exist (then ``bitmap`` / ``rows`` are ignored)."""
    P = C.c_void_p * n_ranks
    _check(lib().fgfa_exchange_push(int(n_ranks), int(rank), bitmap or None, int(rows), partial_depth or None,
                                    partial_uniq or None, P(*recv_bufs), int(n_segs), stream or None))


def exchange_reduce(n_ranks, rank, recv_buf, final_depths, final_uniqs, n_segs, stream=0, multicast_base=0,
                    off_final_depth=0, off_final_uniq=0):
    """Kernel R (see fgfa_exchange_reduce): add the slots of this rank's slice, store the result to every rank."""
    P = C.c_void_p * n_ranks
    _check(lib().fgfa_exchange_reduce(int(n_ranks), int(rank), recv_buf, P(*final_depths), P(*final_uniqs), int(n_segs),
                                      multicast_base or None, int(off_final_depth), int(off_final_uniq), stream or None))


def lpt_partition_c(span_start, span_end, n_parts: int) -> np.ndarray:
    """fgfa_lpt_partition: part index of every path (the C++ partition the multi-GPU ABI uses)."""
    s, e = _u32(span_start), _u32(span_end)
    out = np.empty(s.size, np.uint32)
    _check(lib().fgfa_lpt_partition(s.ctypes.data, e.ctypes.data, int(s.size), int(n_parts), out.ctypes.data))
    return out


class MultiDepth:
    """fgfa_depth_multi_*: one process driving several GPUs (whole paths per device, partial
    [depth | uniq] combined by NCCL or by kernel X over peer memory)."""

    EXCHANGES = {"nccl": 0, "peer": 1}

    def __init__(self, devices, span_start, span_end, n_segs: int, n_steps: int, exchange: str = "nccl"):
        self.span_start, self.span_end = _u32(span_start), _u32(span_end)
        self.n_segs, self.n_steps = int(n_segs), int(n_steps)
        self.devices = np.ascontiguousarray(devices, np.int32)
        if exchange == "nccl" and self.devices.size > 1:
            # the library dlopens libnccl.so.2 on first use; in a process that will also import torch,
            # torch's bundled NCCL (same SONAME, newer) must be the one that gets loaded
            try:
                import torch  # noqa: F401
            except ImportError:
                pass
        h = C.c_void_p()
        rc = lib().fgfa_depth_multi_create(C.byref(h), self.devices.ctypes.data, int(self.devices.size),
                                           self.span_start.ctypes.data, self.span_end.ctypes.data,
                                           int(self.span_start.size), self.n_segs, self.n_steps, self.EXCHANGES[exchange])
        if rc:
            raise DepthError(rc, lib().fgfa_depth_multi_last_error().decode())
        self._h = h

    def _ck(self, rc):
        if rc:
            raise DepthError(rc, lib().fgfa_depth_multi_last_error().decode())

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().fgfa_depth_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def partition(self):
        owner = np.empty(self.span_start.size, np.uint32)
        steps = np.empty(self.devices.size, np.uint64)
        self._ck(lib().fgfa_depth_multi_partition(self._h, owner.ctypes.data, steps.ctypes.data))
        return owner, steps

    def upload(self, h_steps) -> None:
        h = _u32(h_steps)
        self._ck(lib().fgfa_depth_multi_upload(self._h, h.ctypes.data))

    def run(self, with_uniq: bool = True) -> None:
        self._ck(lib().fgfa_depth_multi_run(self._h, 1 if with_uniq else 0))

    def sync(self) -> None:
        self._ck(lib().fgfa_depth_multi_sync(self._h))

    def download(self, with_uniq: bool = True):
        d = np.empty(self.n_segs, np.uint64)
        u = np.empty(self.n_segs, np.uint64) if with_uniq else None
        self._ck(lib().fgfa_depth_multi_download(self._h, d.ctypes.data, u.ctypes.data if with_uniq else None))
        return d, u

    def run_host(self, h_steps, with_uniq: bool = True):
        h = _u32(h_steps)
        d = np.empty(self.n_segs, np.uint64)
        u = np.empty(self.n_segs, np.uint64) if with_uniq else None
        self._ck(lib().fgfa_depth_multi_run_host(self._h, h.ctypes.data, d.ctypes.data, u.ctypes.data if with_uniq else None))
        return d, u

    def run_host_ptr(self, h_ptr: int, d_ptr: int, u_ptr: int) -> None:
        """Raw-pointer form (pinned torch tensors / numpy buffers owned by the caller)."""
        self._ck(lib().fgfa_depth_multi_run_host(self._h, h_ptr, d_ptr, u_ptr or None))


class DepthPlan:
    """Device-resident form: a plan (chunk table + seen-bitmap scratch) run on device
    buffers.  Buffers are torch CUDA tensors (uint32 viewed as int32 is fine: only
    ``data_ptr()`` is used); torch is plumbing for memory and streams here."""

    def __init__(self, span_start, span_end, n_segs: int, n_steps: int, bitmap_budget_bytes: int = 0):
        self.span_start, self.span_end = _u32(span_start), _u32(span_end)
        self.n_paths, self.n_segs, self.n_steps = int(self.span_start.size), int(n_segs), int(n_steps)
        h = C.c_void_p()
        _check(
            lib().fgfa_depth_plan_create(
                C.byref(h), self.span_start.ctypes.data, self.span_end.ctypes.data,
                self.n_paths, self.n_segs, self.n_steps, bitmap_budget_bytes,
            )
        )
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().fgfa_depth_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _ptr(t) -> Optional[int]:
        """Device pointer of a tensor-like (``data_ptr()``), or a raw integer address."""
        if t is None:
            return None
        return int(t) if isinstance(t, int) else int(t.data_ptr())

    def run(self, d_steps, d_depth, d_uniq=None, stream: int = 0) -> None:
        """Enqueue memset + kernel A (+ kernel B) on ``stream``; asynchronous."""
        _check(lib().fgfa_depth_plan_run(self._h, self._ptr(d_steps), self._ptr(d_depth), self._ptr(d_uniq), stream or None))

    def begin(self, d_depth, stream: int = 0) -> None:
        _check(lib().fgfa_depth_plan_begin(self._h, self._ptr(d_depth), stream or None))

    def feed(self, d_steps, path_lo: int, path_hi: int, d_depth, d_uniq=None, stream: int = 0) -> None:
        _check(lib().fgfa_depth_plan_feed(self._h, self._ptr(d_steps), path_lo, path_hi, self._ptr(d_depth), self._ptr(d_uniq), stream or None))

    def finish(self, d_uniq=None, stream: int = 0) -> None:
        _check(lib().fgfa_depth_plan_finish(self._h, self._ptr(d_uniq), stream or None))

    def status(self, stream: int = 0) -> None:
        """Synchronise and raise DepthError if a run saw an out-of-range segment id."""
        _check(lib().fgfa_depth_plan_status(self._h, stream or None))

    def use_bitmap(self, ptr: int, nbytes: int) -> None:
        """Keep the seen-bitmap in caller-provided (peer-mapped) device memory."""
        _check(lib().fgfa_depth_plan_use_bitmap(self._h, ptr, nbytes))

    def run_stream_only(self, d_steps, depth_ptr: int, stream: int = 0) -> None:
        """memset(depth) + kernel A only (partial depth + seen-bits); no popcount."""
        _check(lib().fgfa_depth_plan_run_stream_only(self._h, self._ptr(d_steps), depth_ptr, stream or None))

    @property
    def bitmap_row_bytes(self) -> int:
        return ((((self.n_segs + 31) // 32) + 31) // 32 * 32) * 4

    def set_uniq_width(self, nbytes: int) -> None:
        """uniq counters as u32 (4, default) or u8 (1; plans of <= 255 paths)."""
        _check(lib().fgfa_depth_plan_set_uniq_width(self._h, nbytes))

    def set_probe(self, before_event: int, after_event: int) -> None:
        """Record the given CUDA events around the next run's step-stream kernel."""
        _check(lib().fgfa_depth_plan_set_probe(self._h, before_event or None, after_event or None))

    def launches(self, with_uniq: bool = True) -> int:
        return int(lib().fgfa_depth_plan_launches(self._h, 1 if with_uniq else 0))

    ENGINES = ("stream", "window")

    @property
    def engine(self) -> str:
        """"window" (segment-major shared-memory windows) or "stream" (path-major L2 reductions)."""
        return self.ENGINES[int(lib().fgfa_depth_plan_engine(self._h))]

    def set_engine(self, name: str) -> None:
        _check(lib().fgfa_depth_plan_set_engine(self._h, self.ENGINES.index(name)))

    def autotune(self, d_steps, stream: int = 0) -> str:
        """Sample the resident pool once and keep the engine that suits it; returns its name."""
        _check(lib().fgfa_depth_plan_autotune(self._h, self._ptr(d_steps), stream or None))
        return self.engine

    @property
    def scratch_bytes(self) -> int:
        return int(lib().fgfa_depth_plan_scratch_bytes(self._h))
