"""Test-side access to the ORACLE (oracle/depth_oracle.c) and a tiny independent GFA
reader.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may use this; the product never does."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def oracle():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "libdepth_oracle.so")
        h = C.CDLL(path)
        h.oracle_seg_depth_with_uniq.restype = C.c_int
        h.oracle_seg_depth_with_uniq.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        h.oracle_seg_depth_with_uniq_parallel.restype = C.c_int
        h.oracle_seg_depth_with_uniq_parallel.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        h.oracle_seg_depth.restype = C.c_int
        h.oracle_seg_depth.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        h.oracle_file_seg_depth_with_uniq.restype = C.c_int
        h.oracle_file_seg_depth_with_uniq.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        h.oracle_file_seg_names.restype = C.c_int
        h.oracle_file_seg_names.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        h.oracle_emit_seg_depth.restype = C.c_int64
        h.oracle_emit_seg_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint64]

        h.oracle_path_depth.restype = C.c_int
        h.oracle_path_depth.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        h.oracle_format_float.restype = C.c_int
        h.oracle_format_float.argtypes = [C.c_double, C.c_int, C.c_char_p, C.c_size_t]

        h.oracle_window_count.restype = C.c_uint64
        h.oracle_window_count.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        h.oracle_make_windows.restype = None
        h.oracle_make_windows.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        h.oracle_path_length.restype = C.c_uint64
        h.oracle_path_length.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        h.oracle_interval_depth.restype = C.c_int
        h.oracle_interval_depth.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32,
                                            C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        h.oracle_parse_bed.restype = C.c_int64
        h.oracle_parse_bed.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        h.oracle_emit_interval_depth.restype = C.c_int64
        h.oracle_emit_interval_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                                 C.c_char_p, C.c_uint64]

        class View(C.Structure):
            _fields_ = [(n, C.c_uint64) for n in ("n_segs", "n_paths", "n_steps", "segs_off", "paths_off", "steps_off")]

        h.View = View
        h.oracle_view.restype = C.c_int
        h.oracle_view.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(View)]
        _LIB = h
    return _LIB


def spans_of(start, end):
    s = np.empty(2 * len(start), dtype=np.uint32)
    s[0::2] = start
    s[1::2] = end
    return s


def depth_with_uniq(steps, start, end, n_segs):
    """oracle_seg_depth_with_uniq -> (rc, depth u64, uniq u64)."""
    steps = np.ascontiguousarray(steps, dtype=np.uint32)
    sp = spans_of(start, end)
    d = np.empty(n_segs, dtype=np.uint64)
    u = np.empty(n_segs, dtype=np.uint64)
    rc = oracle().oracle_seg_depth_with_uniq(steps.ctypes.data, steps.size, sp.ctypes.data, len(start), n_segs, d.ctypes.data, u.ctypes.data)
    return rc, d, u


def depth_with_uniq_parallel(steps, start, end, n_segs, n_threads):
    """The path-parallel CPU variant (NOT the reference's algorithm; reported beside the baseline)."""
    steps = np.ascontiguousarray(steps, dtype=np.uint32)
    sp = spans_of(start, end)
    d = np.empty(n_segs, dtype=np.uint64)
    u = np.empty(n_segs, dtype=np.uint64)
    rc = oracle().oracle_seg_depth_with_uniq_parallel(steps.ctypes.data, steps.size, sp.ctypes.data, len(start), n_segs,
                                                       d.ctypes.data, u.ctypes.data, n_threads)
    return rc, d, u


def depth_only(steps, start, end, n_segs):
    steps = np.ascontiguousarray(steps, dtype=np.uint32)
    sp = spans_of(start, end)
    d = np.empty(n_segs, dtype=np.uint64)
    rc = oracle().oracle_seg_depth(steps.ctypes.data, steps.size, sp.ctypes.data, len(start), n_segs, d.ctypes.data)
    return rc, d


def path_depth(steps, start, end, seg_len, path_ids=None):
    """oracle_path_depth -> (rc, lengths u64, means f64)."""
    steps = np.ascontiguousarray(steps, dtype=np.uint32)
    seg_len = np.ascontiguousarray(seg_len, dtype=np.uint32)
    sp = spans_of(start, end)
    ids = None if path_ids is None else np.ascontiguousarray(path_ids, dtype=np.uint32)
    n = len(start) if ids is None else ids.size
    lengths = np.empty(n, dtype=np.uint64)
    means = np.empty(n, dtype=np.float64)
    rc = oracle().oracle_path_depth(steps.ctypes.data, steps.size, sp.ctypes.data, len(start), seg_len.ctypes.data,
                                    seg_len.size, None if ids is None else ids.ctypes.data, n,
                                    lengths.ctypes.data, means.ctypes.data)
    return rc, lengths, means


def interval_depth(steps, start, end, seg_len, path, win_start, win_end):
    """oracle_interval_depth -> (rc, depths f64)."""
    steps = np.ascontiguousarray(steps, dtype=np.uint32)
    seg_len = np.ascontiguousarray(seg_len, dtype=np.uint32)
    sp = spans_of(start, end)
    ws = np.ascontiguousarray(win_start, dtype=np.uint64)
    we = np.ascontiguousarray(win_end, dtype=np.uint64)
    out = np.empty(ws.size, dtype=np.float64)
    rc = oracle().oracle_interval_depth(steps.ctypes.data, steps.size, sp.ctypes.data, len(start), seg_len.ctypes.data,
                                        seg_len.size, path, ws.ctypes.data, we.ctypes.data, ws.size, out.ctypes.data)
    return rc, out


def windows(start, end, size):
    """oracle_window_count + oracle_make_windows -> (win_start u64, win_end u64)."""
    m = int(oracle().oracle_window_count(start, end, size))
    ws = np.empty(m, dtype=np.uint64)
    we = np.empty(m, dtype=np.uint64)
    oracle().oracle_make_windows(start, end, size, ws.ctypes.data, we.ctypes.data)
    return ws, we


def path_length(steps, start, end, seg_len, path):
    steps = np.ascontiguousarray(steps, dtype=np.uint32)
    seg_len = np.ascontiguousarray(seg_len, dtype=np.uint32)
    sp = spans_of(start, end)
    return int(oracle().oracle_path_length(steps.ctypes.data, sp.ctypes.data, path, seg_len.ctypes.data))


def parse_bed(text: bytes):
    """oracle_parse_bed -> None where the reference panics, else [(name bytes, start, end)]."""
    buf = np.frombuffer(text, dtype=np.uint8)
    cap = text.count(b"\n") + 1
    cols = [np.empty(cap, dtype=np.uint64) for _ in range(4)]
    n = oracle().oracle_parse_bed(buf.ctypes.data if buf.size else None, buf.size, *[c.ctypes.data for c in cols], cap)
    if n < 0:
        return None
    return [(text[int(cols[0][i]):int(cols[0][i] + cols[1][i])], int(cols[2][i]), int(cols[3][i])) for i in range(n)]


def emit_interval_depth(entries, depths) -> bytes:
    """window_depth.rs:163-174 through oracle_emit_interval_depth; entries = [(name bytes, start, end)]."""
    names = b"".join(e[0] for e in entries)
    off = np.cumsum([0] + [len(e[0]) for e in entries[:-1]], dtype=np.uint64) if entries else np.empty(0, np.uint64)
    ln = np.array([len(e[0]) for e in entries], dtype=np.uint64)
    s = np.array([e[1] for e in entries], dtype=np.uint64)
    t = np.array([e[2] for e in entries], dtype=np.uint64)
    d = np.ascontiguousarray(depths, dtype=np.float64)
    cap = 64 + sum(len(e[0]) + 500 for e in entries)
    out = C.create_string_buffer(cap)
    nb = np.frombuffer(names, dtype=np.uint8)
    k = oracle().oracle_emit_interval_depth(nb.ctypes.data if nb.size else None, off.ctypes.data, ln.ctypes.data, s.ctypes.data,
                                            t.ctypes.data, d.ctypes.data, len(entries), out, cap)
    assert k >= 0
    return out.raw[:k]


def format_float(x: float, digits: int = 2) -> str:
    buf = C.create_string_buffer(600)
    n = oracle().oracle_format_float(x, digits, buf, 600)
    assert n >= 0
    return buf.value.decode()


def emit_path_depth(path_names, lengths, means) -> bytes:
    """depth.rs:146-158 with the oracle's format_float."""
    out = ["#path\tstart\tend\tmean.depth"]
    for name, ln, m in zip(path_names, lengths, means):
        out.append(f"{name}\t0\t{int(ln)}\t{format_float(float(m))}")
    return ("\n".join(out) + "\n").encode()


def file_depth(image: bytes):
    """Oracle over a .flatgfa image -> (rc, names, depth, uniq)."""
    o = oracle()
    buf = np.frombuffer(image, dtype=np.uint8)
    v = o.View()
    rc = o.oracle_view(buf.ctypes.data, buf.size, C.byref(v))
    if rc:
        return rc, None, None, None
    n = int(v.n_segs)
    names = np.empty(n, dtype=np.uint64)
    d = np.empty(n, dtype=np.uint64)
    u = np.empty(n, dtype=np.uint64)
    rc = o.oracle_file_seg_depth_with_uniq(buf.ctypes.data, buf.size, d.ctypes.data, u.ctypes.data)
    if rc:
        return rc, None, None, None
    o.oracle_file_seg_names(buf.ctypes.data, buf.size, names.ctypes.data)
    return 0, names, d, u


def emit(names, depth, uniq) -> bytes:
    n = len(names)
    cap = 64 + 64 * n
    out = C.create_string_buffer(cap)
    names = np.ascontiguousarray(names, dtype=np.uint64)
    depth = np.ascontiguousarray(depth, dtype=np.uint64)
    uniq = np.ascontiguousarray(uniq, dtype=np.uint64)
    k = oracle().oracle_emit_seg_depth(names.ctypes.data, depth.ctypes.data, uniq.ctypes.data, n, out, cap)
    assert k >= 0
    return out.raw[:k]


def read_gfa(text: str):
    """Independent minimal GFA reader for pinning the oracle (not the product parser):
    segment pool index = order of S lines (flatgfa/src/parse.rs:138-141), steps in P-line
    order, Handle = index << 1 | (orientation == '-') (flatgfa/src/flatgfa.rs:192-198).
    Returns (names, steps, start, end, path_names)."""
    names, index = [], {}
    for line in text.split("\n"):
        f = line.split("\t")
        if f[0] == "S":
            index[f[1]] = len(names)
            names.append(int(f[1]))
    steps, start, end, pnames = [], [], [], []
    for line in text.split("\n"):
        f = line.split("\t")
        if f[0] == "P":
            pnames.append(f[1])
            start.append(len(steps))
            for tok in f[2].split(","):
                if tok:
                    steps.append((index[tok[:-1]] << 1) | (1 if tok[-1] == "-" else 0))
            end.append(len(steps))
    return (np.array(names, dtype=np.uint64), np.array(steps, dtype=np.uint32),
            np.array(start, dtype=np.uint32), np.array(end, dtype=np.uint32), pnames)
